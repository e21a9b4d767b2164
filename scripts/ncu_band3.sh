#!/bin/bash
# ncu --set full capture of the v3 banded Cholesky kernel alone (n=1862, bw=370: the C1 normal equations' shape)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/one_band3.py <<'PY'
import os, sys
ROOT = os.environ["GRAFT_REPO_ROOT"] if "GRAFT_REPO_ROOT" in os.environ else os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-super_b200"))
import torch
from super_b200 import ops
n, bw = 1862, 370
g = torch.Generator().manual_seed(0)
AB = torch.randn((n, bw + 1), generator=g, dtype=torch.float64); AB[:, bw] = AB.abs().sum(1) * 2 + 1.0
band = ops.Band(n, bw, None, "cuda"); ABd = AB.cuda(); rd = torch.randn(n, generator=g, dtype=torch.float64).cuda()
for _ in range(4):
    band.AB.copy_(ABd); band.g.copy_(rd); ops.band_solve(band, None, 148, variant=3)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:band_chol3_kernel -s 2 -c 1 -f -o gpurun_out/prof_band3 python /tmp/one_band3.py > gpurun_out/ncu_band3.log 2>&1
tail -3 gpurun_out/ncu_band3.log
