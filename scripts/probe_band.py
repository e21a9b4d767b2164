"""GPU probe: time the band solvers alone (n=1862, bw=300)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-super_b200"))
import torch
from super_b200 import ops, lib
n, bw = 1862, int(os.environ.get("BW", "300"))
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
res = {}
L = lib.load()
for flags, name in ((0, "v2 full"), (1, "v2 P chain + R + backsub (no U work)"), (2, "v2 no backsub"), (3, "v2 P chain + R only")):
    L.sb_band2_debug(flags)
    for cs in (8, 16, 32, 64):
        res[f"{name}/c{cs}"] = timeit(cs, 2)
L.sb_band2_debug(0)
res["v1 full/c16"] = timeit(16, 1)
# dense library reference on the same matrix size
A = torch.randn(n, n, dtype=torch.float64, device="cuda"); A = A @ A.t() + n * torch.eye(n, dtype=torch.float64, device="cuda")
b = rd[:, None].clone()
def lib_solve():
    Lc, info = torch.linalg.cholesky_ex(A, check_errors=False); return torch.cholesky_solve(b, Lc)
for _ in range(3): lib_solve()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): lib_solve()
e1.record(); torch.cuda.synchronize()
res["library dense potrf+potrs"] = round(e0.elapsed_time(e1) / 10 * 1e3, 1)
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_band.json"), "w"), indent=1)

# per-phase cycle counters of the P role (debug flag 4), with and without U work
import numpy as np
for flags in (4, 5):
    L.sb_band2_debug(flags)
    band.AB.copy_(ABd); band.g.copy_(rd); ops.band_solve(band, None, int(os.environ.get('CS', '32')), variant=2); torch.cuda.synchronize()
    NP = (n + 31) // 32
    off = n * band.ldab * 8 + 3 * NP * 4 + 64
    off = (off + 7) & ~7
    prof = band.ws2[off:off + 192].cpu().numpy().view(np.int64)
    un = ["wait upd[p-1]", "wait diag[p]", "rows(after)", "signal rows", "-", "tiles", "signal upd", "loop top"]
    print("   U(rank2):", {nm: int(v) // NP for nm, v in zip(un, prof[16:24])})
    print("   publish path (tid 32): store", prof[8] // NP, "bar", prof[9] // NP, "red.release", prof[10] // NP, "| warp0 trsm", prof[11] // NP)
    names = ["loop top+damp", "warp0 trsm", "barrier wait after trsm||publish", "syrk", "potrf(+stage next)", "-"]
    print("flags", flags, {nm: round(float(v) / NP) for nm, v in zip(names, prof)}, "cycles per panel; total", int(prof[:6].sum()))
L.sb_band2_debug(0)
