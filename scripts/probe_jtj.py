"""GPU probe: data_jtj_kernel / data_loss time against the number of surfels (row count overridden on the device),
dense vs band target -- separates the per-chunk cost from the fixed cost of a launch."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-super_b200"))
import torch
from oracle import super_oracle as so
from super_b200 import synth, ops, lm, engine

H, W, step = 480, 640, 32
opt = so.default_opt(height=H, width=W, mesh_step_size=step)
tex = synth.texture(H, W)
trk = engine.Tracker(opt, device="cuda:0")
for t in (1, 2, 3):
    fr = synth.frame_inputs(t, H, W, tex=tex)
    trk.step(torch.from_numpy(fr["depth"]).cuda(), torch.from_numpy(fr["color"]).cuda(), torch.from_numpy(fr["K"]),
             torch.from_numpy(fr["inv_K"]), fr["time"])
trk._refresh_bound()
n_full = trk.n_bound
frame = trk.frames[trk._fi]
sfv = trk.view(n_full)
ws, band = trk.ws, trk.band

def timeit(fn, n=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

res = {}
for n in (n_full, 200000, 100000, 50000, 25000, 3000):
    n = min(n, n_full)
    nd = torch.tensor([n], dtype=torch.int32, device="cuda")
    order = ops.tuple_order(sfv.knn_indices, nd, trk.ED.node_pos, trk.block_bw)
    v = trk.view(n_full)
    res[n] = {
        "jtj_band_us": round(timeit(lambda: ops.data_term_jtj(v.points, v.knn_indices, v.knn_w, order, trk.ED.points, ws.beta, frame.vmap, frame.nmap, frame.cam, 1.0, None, None, n_dev=nd, band=band)), 1),
        "jtj_band_exact_rows_us": round(timeit(lambda: ops.data_term_jtj(v.points[:n], v.knn_indices[:n], v.knn_w[:n], order, trk.ED.points, ws.beta, frame.vmap, frame.nmap, frame.cam, 1.0, None, None, n_dev=nd, band=band)), 1),
        "loss_us": round(timeit(lambda: ops.data_term_loss(v.points, v.knn_indices, v.knn_w, trk.ED.points, ws.beta, frame.vmap, frame.nmap, frame.cam, 1.0, ws.partials, n_dev=nd)), 1),
    }
    print(n, res[n], flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_jtj.json"), "w"), indent=1)
