"""Does replaying the solve's launches from a CUDA graph shorten the gaps between the dependent kernels?  Ten two-sided
solves (restore + 5 launches each) issued directly vs captured once and replayed (torch.cuda.CUDAGraph = stream capture)."""
import os, sys
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-super_b200"))
import torch
from super_b200 import ops
n, bw = 1862, 320
g = torch.Generator().manual_seed(0)
AB = torch.randn((n, bw + 1), generator=g, dtype=torch.float64); AB[:, bw] = AB.abs().sum(1) * 2 + 1.0
band = ops.Band(n, bw, None, "cuda"); ABd = AB.cuda(); rd = torch.randn(n, generator=g, dtype=torch.float64).cuda()

def ten():
    for _ in range(10):
        band.AB.copy_(ABd); band.g.copy_(rd)
        ops.band_solve(band, None, 148, variant=4)

def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / 10)
    return [round(t, 1) for t in ts]

ten(); torch.cuda.synchronize()
print("direct launches, us per (restore + solve):", timed(ten))
x_direct = band.g.clone()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    ten(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        ten()
torch.cuda.synchronize()
print("graph replay,    us per (restore + solve):", timed(gr.replay))
print("same result:", bool(torch.equal(x_direct, band.g)))
