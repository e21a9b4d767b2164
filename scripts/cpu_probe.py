"""Host time of the per-frame calls (C1 workload): how long the CPU needs to issue a frame, and sb_lm_frame's share."""
import os, sys, time
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-super_b200"))
import torch
from oracle import super_oracle as so
from super_b200 import engine, synth, lm, lib
H, W = 480, 640
opt = so.default_opt(height=H, width=W, mesh_step_size=32)
tex = synth.texture(H, W)
trk = engine.Tracker(opt, device="cuda:0")
fr = [synth.frame_inputs(t, H, W, tex=tex) for t in range(1, 40)]
dd = [torch.from_numpy(f["depth"]).cuda() for f in fr]
dc = torch.from_numpy(fr[0]["color"]).cuda()
K, iK = torch.from_numpy(fr[0]["K"]), torch.from_numpy(fr[0]["inv_K"])
acc = {"lm_frame": 0.0, "refresh": 0.0, "n": 0}
calls = []
orig = lm.lm_frame
def timed(*a, **k):
    t0 = time.perf_counter(); r = orig(*a, **k); dt = time.perf_counter() - t0; acc["lm_frame"] += dt; calls.append(round(1e6 * dt)); acc["n"] += 1; return r
lm.lm_frame = timed
orig_rb = engine.Tracker._refresh_bound
def timed_rb(self):
    t0 = time.perf_counter(); r = orig_rb(self); acc["refresh"] += time.perf_counter() - t0; return r
engine.Tracker._refresh_bound = timed_rb
for i in range(0, 8):
    trk.step(dd[i], dc, K, iK, fr[i]["time"])
torch.cuda.synchronize()
acc.update(lm_frame=0.0, refresh=0.0, n=0)
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(8, 38):
    trk.step(dd[i], dc, K, iK, fr[i]["time"])
e1.record(); torch.cuda.synchronize()
wall = time.perf_counter() - t0
print("frames 30: GPU ms/frame %.3f | wall ms/frame %.3f | host in sb_lm_frame (capture + update + launch) ms/frame %.3f | host waiting in _refresh_bound ms/frame %.3f"
      % (e0.elapsed_time(e1) / 30, 1e3 * wall / 30, 1e3 * acc["lm_frame"] / 30, 1e3 * acc["refresh"] / 30))
print("host us per sb_lm_frame call, last 30 frames:", calls[-30:])
