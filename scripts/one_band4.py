import os, sys
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-super_b200"))
import torch
from super_b200 import ops
n, bw = 1862, 370
g = torch.Generator().manual_seed(0)
AB = torch.randn((n, bw + 1), generator=g, dtype=torch.float64); AB[:, bw] = AB.abs().sum(1) * 2 + 1.0
band = ops.Band(n, bw, None, "cuda"); ABd = AB.cuda(); rd = torch.randn(n, generator=g, dtype=torch.float64).cuda()
for _ in range(4):
    band.AB.copy_(ABd); band.g.copy_(rd); ops.band_solve(band, None, 148, variant=4)
torch.cuda.synchronize()
