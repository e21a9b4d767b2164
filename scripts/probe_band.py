"""GPU probe: time sb_band_solve alone (n=1862, bw=300) for cluster sizes and with phases disabled."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-super_b200"))
import torch
from super_b200 import ops, lib
n, bw = 1862, 300
g = torch.Generator().manual_seed(0)
AB = torch.randn((n, bw + 1), generator=g, dtype=torch.float64)
AB[:, bw] = AB.abs().sum(1) * 2 + 1.0
rhs = torch.randn(n, generator=g, dtype=torch.float64)
band = ops.Band(n, bw, None, "cuda")
ABd, rd = AB.cuda(), rhs.cuda()
def run(cs):
    band.AB.copy_(ABd); band.g.copy_(rd)
    ops.band_solve(band, None, cs)
def timeit(cs, n_it=10):
    for _ in range(3): run(cs)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n_it):
        band.AB.copy_(ABd); band.g.copy_(rd)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.band_solve(band, None, cs); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n_it * 1e3
res = {}
for flags, name in ((0, "full"), (1, "no_trailing_update"), (2, "no_backsub"), (3, "no_update_no_backsub"), (7, "memory+sync only"), (6, "update only (no panel math/backsub)")):
    lib.load().sb_band_debug(flags)
    for cs in (1, 4, 8, 16):
        res[f"{name}/c{cs}"] = round(timeit(cs), 1)
lib.load().sb_band_debug(0)
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_band.json"), "w"), indent=1)
