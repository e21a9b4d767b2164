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
    e0.record(); ops.band_solve(band, None, int(os.environ.get("CS", 148)), variant=variant); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(f"n={n} bw={bw} variant={variant}: us per solve {[round(t, 1) for t in ts]}")
if variant == 4 and os.environ.get("STAGES", "1") == "1":
    import ctypes, numpy as np
    from super_b200 import lib
    l = lib.load()
    for flags, label in ((256, "cluster back substitution"), (256 | 512, "bulk-copy back substitution"), (256 | 128, "register-prefetch back substitution"), (256, "cluster back substitution")):
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
if variant == 4 and os.environ.get("PROF4", "0") == "1":
    from super_b200 import lib
    l = lib.load()
    need = max(bw, 64); m = (n - need) // 64
    while m > 0 and n - 64 * m < need: m -= 1
    nA = 32 * m + (n - 64 * m)
    off = int(l.sb_band3_prof_offset(nA, bw))
    l.sb_band3_debug(1024)
    for _ in range(3):
        band.AB.copy_(ABd); band.g.copy_(rd); ops.band_solve(band, None, 148, variant=4)
    torch.cuda.synchronize()
    pr = band.ws4[off: off + 1024].view(torch.int64).cpu().numpy()
    t0 = pr[56]
    print("  cluster back substitution, leader cycles since entry: after init+cluster.sync", pr[57] - t0, "| given panels done", pr[60] - t0,
          "| chain done", pr[58] - t0, "| cta sync", pr[59] - t0, "| exit", pr[61] - t0)
    print("  leader warp 0 totals over its panels: wait for x_{k+1}", pr[64], "| L(k+1,k)^T x + wait pg", pr[65], "| inverse product + x stores", pr[66])
    l.sb_band3_debug(0)

if variant == 4 and os.environ.get("PROFP", "0") == "1":
    # role counters (debug 4) of the TOP instance of the two-sided solve (debug 2048), per eliminated panel
    import numpy as np
    from super_b200 import lib
    l = lib.load()
    need = max(bw, 64); m = (n - need) // 64
    while m > 0 and n - 64 * m < need: m -= 1
    nA = 32 * m + (n - 64 * m)
    off = int(l.sb_band3_prof_offset(nA, bw))
    l.sb_band3_debug(4 | 2048)
    for _ in range(3):
        band.AB.copy_(ABd); band.g.copy_(rd); ops.band_solve(band, None, 148, variant=4)
    torch.cuda.synchronize()
    prof = band.ws4[off: off + 1024].view(torch.int64).cpu().numpy()
    gs = prof[64:69].astype(np.float64) / np.maximum(prof[69:74], 1)
    print("  hot path (globaltimer, ns after the Cholesky end of panel p; counts", prof[69:74].tolist(), "): U owner of (p+2,p+1) starts polling the inverse %.0f | has it %.0f | tile in the mailbox %.0f | pivot CTA has staged it %.0f" % tuple(gs[1:5] - gs[0]))
    names = ["Cholesky end -> barrier 7 reached", "wait at barrier 7", "wait for block column 0 of D (final step)", "-", "Cholesky"]
    print("  top instance, pivot CTA, cycles per eliminated panel (m = %d):" % m, {nm: int(v) // m for nm, v in zip(names, prof[:5])},
          "| R forward substitution done at", int(prof[10] - prof[8]), "after P's start | P total", int(prof[9] - prof[8]), "=", int(prof[9] - prof[8]) // m, "per panel",
          "\n   Cholesky parts [replica loads (block 0 since the panel's barrier), chain + T_b, urgent trailing DMMA + barrier]", [int(v) // m for v in prof[48:51]],
          "\n   inverse builder: total", int(prof[30]) // m, "[wait cols<b, S products, wait T_b, M products, publish]", [int(v) // m for v in prof[31:36]],
          "| inverse builder's barrier-7 arrival after the Cholesky end", int(prof[37]) // m, "| warps 2..7: (panels in which they reached barrier 7 AFTER the Cholesky's end, mean delay)", [(int(v) >> 32, (int(v) & 0xffffffff) // max(1, int(v) >> 32)) for v in prof[38:44]], "| row b published since its own loop top", [int(v) // m for v in prof[44:48]],
          "\n   compute warp 3 since the panel's barrier [T_1 seen, arrived at the staging barrier, substitution loop done, T_2 seen]", [int(v) // m for v in prof[58:62]],
          "| final step in warp 3 since T_3 seen [all four Y_3 exchanged, D value ready]", [int(v) // m for v in prof[62:64]], "| I/O warp 7 [staging begins, staged]", [int(v) // m for v in prof[6:8]],
          "\n   U (rank 3; it owns a tile in few panels, so its waits are mostly idle time) [wait upd(p-1), operand loads, inverse tile by value, products + stores, signal, idle]:", [int(v) // m for v in prof[52:58]])
    l.sb_band3_debug(0)
