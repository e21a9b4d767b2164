"""north_star's in-frame study: surfel-sharded J^T J with an all-reduce, 2 GPUs against 1 GPU.

    python scripts/study_sharded_jtj.py                                   # 1 GPU: the unsharded stage times
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/study_sharded_jtj.py                                      # 2 GPUs: each rank assembles half of the surfels

Both ranks hold the whole (replicated, bitwise identical) tracker state.  Per LM iteration a rank assembles J^T J / -J^T r
over ITS half of the visiting order into the fixed-point store, the stores are summed with ncclAllReduce (int64 sum: the
result stays order-independent), both ranks solve redundantly, each evaluates the loss over its half of the surfels and
the two partial sums are all-reduced before the accept/reject step.  Stage-by-stage loop (lm.lm_solve's step-wise form)
on both sides, so that the comparison is like for like.  Writes gpurun_out/study_sharded_jtj_<N>gpu.json on rank 0.
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-super_b200"))
import torch
import torch.distributed as dist
from oracle import super_oracle as so
from super_b200 import engine, lm, ops, synth
from super_b200.lib import call, ptr, stream

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H, W = 480, 640
opt = so.default_opt(height=H, width=W, mesh_step_size=32)
tex = synth.texture(H, W)
trk = engine.Tracker(opt, device=f"cuda:{local}")
frames = [synth.frame_inputs(t, H, W, tex=tex) for t in range(1, 9)]
dev = trk.dev
for f in frames[:5]:
    trk.step(torch.from_numpy(f["depth"]).to(dev), torch.from_numpy(f["color"]).to(dev), torch.from_numpy(f["K"]),
             torch.from_numpy(f["inv_K"]), f["time"])
# one more frame, by hand and sharded
f = frames[5]
fr = engine.preprocess(opt, torch.from_numpy(f["depth"]).to(dev), torch.from_numpy(f["color"]).to(dev), torch.from_numpy(f["K"]),
                       torch.from_numpy(f["inv_K"]), f["time"], frame=trk.next_frame())
trk._refresh_bound()
n = trk.num_surfels()
b, ed, band, cam = trk.cur, trk.ED, trk.band, fr.cam
order = trk._order[:n].contiguous()
lo, hi = rank * n // world, (rank + 1) * n // world
ws = lm.LMWorkspace(ed.num, dev)
nb = ops.data_loss_blocks(hi - lo)
partials = torch.zeros(nb, dtype=torch.float64, device=dev)
loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
ev = lambda: torch.cuda.Event(enable_timing=True)
T = {"assemble": [], "allreduce_store": [], "solve": [], "loss": [], "allreduce_loss": [], "iteration": []}
ops.lm_begin(ws.state, ws.beta, ws.best)
band.info.zero_()
for it in range(10):
    e = [ev() for _ in range(6)]
    e[0].record()
    band.store.zero_()
    st = band.store
    nab = band.n * band.ldab
    call("sb_data_term_jtj", ptr(b.points), ptr(b.knn_idx), ptr(b.knn_w), ptr(order[lo:hi]), hi - lo, None, ptr(ed.points),
         ptr(ws.beta), ed.num, ptr(fr.vmap), ptr(fr.nmap), cam.H, cam.W, cam.c, 1.0, ptr(st[:nab]), band.ldab, band.bw,
         ptr(band.node_pos), ptr(band.overflow), ptr(st[nab:]), None, band.fx_shift, band.fx_gshift, stream())
    if rank == 0:
        ops.reg_terms(ed.points, ed.knn_indices, ws.beta, 10.0, 1.0, True, True, band=band)
    e[1].record()
    if world > 1:
        dist.all_reduce(band.fx[0], op=dist.ReduceOp.SUM)
    e[2].record()
    band.finalize()
    ops.band_solve_step(band, ws.state, ws.beta, 148)
    e[3].record()
    call("sb_data_term_loss", ptr(b.points[lo:hi]), ptr(b.knn_idx[lo:hi]), ptr(b.knn_w[lo:hi]), hi - lo, None, ptr(ed.points),
         ptr(ws.beta), ed.num, ptr(fr.vmap), ptr(fr.nmap), cam.H, cam.W, cam.c, 1.0, ptr(partials), nb, stream())
    torch.sum(partials, dim=0, keepdim=True, out=loss_sum)
    e[4].record()
    if world > 1:
        dist.all_reduce(loss_sum, op=dist.ReduceOp.SUM)
    ops.lm_decide_reg(ws.state, loss_sum, ed.points, ed.knn_indices, 10.0, 1.0, True, True, ws.beta, ws.best)
    e[5].record()
    torch.cuda.synchronize()
    for k, (i0, i1) in zip(("assemble", "allreduce_store", "solve", "loss", "allreduce_loss"), ((0, 1), (1, 2), (2, 3), (3, 4), (4, 5))):
        T[k].append(e[i0].elapsed_time(e[i1]) * 1e3)
    T["iteration"].append(e[0].elapsed_time(e[5]) * 1e3)
st_ = ws.state.read()
med = {k: float(torch.tensor(v[2:]).median()) for k, v in T.items()}       # skip the first two (cold) iterations
out = {"n_gpus": world, "rank": rank, "surfels": n, "shard": [lo, hi], "store_bytes": int(band.fx[0].numel() * 8),
       "us_median_per_stage": med, "loss_trace": [float(x) for x in st_["loss"]], "accept": [int(x) for x in st_["accept"]]}
if world > 1:
    t = torch.tensor([med["iteration"]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["us_iteration_max_over_ranks"] = float(t)
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"study_sharded_jtj_{world}gpu.json"), "w"), indent=1)
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
