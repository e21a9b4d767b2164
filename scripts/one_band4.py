"""A few two-sided solves (sb_band_solve4) for an ncu launch list; N/BW/VARIANT from the environment.
Prints the event-timed microseconds per solve (not under the profiler: run it twice)."""
import os, sys
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-super_b200"))
import torch
from super_b200 import ops
n, bw = int(os.environ.get("N", 1862)), int(os.environ.get("BW", 320))
variant = int(os.environ.get("VARIANT", 4))
reps = int(os.environ.get("REPS", 4))
g = torch.Generator().manual_seed(0)
AB = torch.randn((n, bw + 1), generator=g, dtype=torch.float64); AB[:, bw] = AB.abs().sum(1) * 2 + 1.0
band = ops.Band(n, bw, None, "cuda"); ABd = AB.cuda(); rd = torch.randn(n, generator=g, dtype=torch.float64).cuda()
ts = []
for _ in range(reps):
    band.AB.copy_(ABd); band.g.copy_(rd)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.band_solve(band, None, 148, variant=variant); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"n={n} bw={bw} variant={variant}: us per solve {[round(t, 1) for t in ts]}")
if variant == 4 and os.environ.get("STAGES", "1") == "1":
    import ctypes, numpy as np
    from super_b200 import lib
    l = lib.load()
    for flags, label in ((256, "bulk-copy back substitution"), (256 | 128, "register-prefetch back substitution")):
        l.sb_band3_debug(flags)
        acc = np.zeros(5)
        for _ in range(reps):
            band.AB.copy_(ABd); band.g.copy_(rd)
            ops.band_solve(band, None, 148, variant=4)
            out = (ctypes.c_float * 5)()
            l.sb_band4_stage_ms(out)
            acc += np.array(out[:]) * 1e3
        print(f"  stages us [reverse, both ends, combine+memset, middle, back substitution] ({label}): {np.round(acc / reps, 1).tolist()}")
    l.sb_band3_debug(0)
