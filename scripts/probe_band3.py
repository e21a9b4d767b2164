"""GPU probe: time the v3 band solver alone against v2 and the dense library solve."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-super_b200"))
import torch
from super_b200 import ops, lib
n = int(os.environ.get("N", "1862"))
res = {}
for bw in [int(x) for x in os.environ.get("BW", "300,370").split(",")]:
    g = torch.Generator().manual_seed(0)
    AB = torch.randn((n, bw + 1), generator=g, dtype=torch.float64)
    AB[:, bw] = AB.abs().sum(1) * 2 + 1.0
    rhs = torch.randn(n, generator=g, dtype=torch.float64)
    band = ops.Band(n, bw, None, "cuda")
    ABd, rd = AB.cuda(), rhs.cuda()
    def timeit(cs, variant, n_it=10):
        for _ in range(3):
            band.AB.copy_(ABd); band.g.copy_(rd); ops.band_solve(band, None, cs, variant=variant)
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(n_it):
            band.AB.copy_(ABd); band.g.copy_(rd)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.band_solve(band, None, cs, variant=variant); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return round(tot / n_it * 1e3, 1)
    for cs in (148,):
        res[f"v3 bw{bw}/c{cs}"] = timeit(cs, 3)
        res[f"v4 two-sided bw{bw}/c{cs}"] = timeit(cs, 4)
    if lib.load().sb_band2_fits(n, bw):
        res[f"v2 bw{bw}/c64"] = timeit(64, 2)
print(json.dumps(res, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_band3.json"), "w"), indent=1)

# per-phase cycle counters (debug flag 4), with (4) and without (5) the trailing-update work, and without back substitution (6)
import numpy as np
L = lib.load()
bw = int(os.environ.get("PBW", "370"))
g = torch.Generator().manual_seed(0)
AB = torch.randn((n, bw + 1), generator=g, dtype=torch.float64); AB[:, bw] = AB.abs().sum(1) * 2 + 1.0
band = ops.Band(n, bw, None, "cuda"); ABd = AB.cuda(); rd = torch.randn(n, generator=g, dtype=torch.float64).cuda()
NP = (n + 31) // 32
for flags in (4,):
    L.sb_band3_debug(flags)
    for _ in range(2):
        band.AB.copy_(ABd); band.g.copy_(rd); ops.band_solve(band, None, 128, variant=3); torch.cuda.synchronize()
    off = int(L.sb_band3_prof_offset(n, bw))
    prof = band.ws3[off:off + 512].cpu().numpy().view(np.int64)
    names = ["potrf-end->top", "wait BAR_A", "trsm", "syrk", "potrf", "io wait upd"]
    print("flags", flags, {nm: int(v) // NP for nm, v in zip(names, prof[:6])}, "cycles/panel | P total", int(prof[9] - prof[8]),
          "| R fwd end - P end", int(prof[10] - prof[9]), "| backsub", int(prof[11] - prof[10]),
          "\n   io warps [release diag, arm+wait Lx, store Lx, release rows, spin upd, stage]:", [[int(v) // NP for v in prof[12 + 6 * w:18 + 6 * w]] for w in range(3)],
          "| warp1 busy", int(prof[30]) // NP, "[wait prev block, S products, T_b rows, M products, publish]", [int(v) // NP for v in prof[31:36]],
          "\n   block-end times after BAR_A [potrf], [T], [M+publish]:", [[int(v) // NP for v in prof[36 + 4 * w:40 + 4 * w]] for w in range(3)], "| potrf parts [load, chain+T, dmma update]", [int(v) // NP for v in prof[48:51]],
          "\n   U (rank 3) per panel [wait upd(p-1), operand loads, wait diag(p), Linv load + products + stores, signal, idle]:", [int(v) // NP for v in prof[52:58]])
L.sb_band3_debug(0)
