"""GPU probe (development aid): time the LM pieces at config-1 scale on synthetic state built by the
oracle port's producer / init (CPU) and moved to the device."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-super_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import super_oracle as so
from super_b200 import synth, ops, lm
from gpu_util import to_device_state, device_maps, camera

H, W, step = 480, 640, 32
opt = so.default_opt(height=H, width=W, mesh_step_size=step)
tex = synth.texture(H, W)
t0 = time.time()
nd1 = so.preprocess(opt, synth.frame_inputs(1, H, W, tex=tex))
sf = so.init_surfels(opt, nd1, so.build_graph(opt, nd1))
nd2 = so.preprocess(opt, synth.frame_inputs(2, H, W, tex=tex))
print("cpu init", time.time() - t0, "N", len(sf.points), "J", sf.ED.num, flush=True)
d = to_device_state(sf)
maps = device_maps(nd2, H, W)
cam = camera(nd2, H, W)

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

J = sf.ED.num
ws = lm.LMWorkspace(J, "cuda")
ws.partials = torch.zeros(ops.data_loss_blocks(len(sf.points)), dtype=torch.float64, device="cuda")
ops.lm_begin(ws.state, ws.beta, ws.best)
order = ops.tuple_order(d.knn_indices)
res = {}
res["tuple_order_ms"] = timeit(lambda: ops.tuple_order(d.knn_indices))
res["zero_A_ms"] = timeit(lambda: ws.A.zero_())
res["jtj_sorted_ms"] = timeit(lambda: ops.data_term_jtj(d.points, d.knn_indices, d.knn_w, order, d.ED.points, ws.beta, maps[0], maps[1], cam, 1.0, ws.A, ws.g))
res["jtj_natural_ms"] = timeit(lambda: ops.data_term_jtj(d.points, d.knn_indices, d.knn_w, None, d.ED.points, ws.beta, maps[0], maps[1], cam, 1.0, ws.A, ws.g))
res["loss_ms"] = timeit(lambda: ops.data_term_loss(d.points, d.knn_indices, d.knn_w, d.ED.points, ws.beta, maps[0], maps[1], cam, 1.0, ws.partials))
res["reg_ms"] = timeit(lambda: ops.reg_terms(d.ED.points, d.ED.knn_indices, ws.beta, 10.0, 1.0, True, True, ws.A, ws.g))
ws.A.zero_(); ws.g.zero_()
ops.data_term_jtj(d.points, d.knn_indices, d.knn_w, order, d.ED.points, ws.beta, maps[0], maps[1], cam, 1.0, ws.A, ws.g)
ops.reg_terms(d.ED.points, d.ED.knn_indices, ws.beta, 10.0, 1.0, True, True, ws.A, ws.g)
ops.lm_damp(ws.state, ws.A)
res["cholesky_ex_ms"] = timeit(lambda: torch.linalg.cholesky_ex(ws.A, check_errors=False))
L, info = torch.linalg.cholesky_ex(ws.A, check_errors=False)
res["cholesky_solve_ms"] = timeit(lambda: torch.cholesky_solve(ws.g, L))
res["warp_update_ms"] = timeit(lambda: ops.warp_update(d.points.clone(), d.norms.clone(), d.knn_indices, d.knn_w, d.ED.points.clone(), d.ED.norms.clone(), ws.beta))
res["lm_10it_ms"] = timeit(lambda: lm.lm_solve(d, maps, cam, opt, ws=ws, order=order), n=5, warm=2)
beta, ws = lm.lm_solve(d, maps, cam, opt, ws=ws, order=order)
st = ws.state.read()
res["losses"] = st["loss"].tolist(); res["accept"] = st["accept"].tolist()
# band structure of A
A = ws.A.cpu(); nz = (A.abs() > 0)
rows, cols = nz.nonzero(as_tuple=True)
res["scalar_half_bandwidth"] = int((rows - cols).abs().max())
res["nnz_lower"] = int(nz.sum())
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_lm.json"), "w"), indent=1)
