import os, sys
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-super_b200"))
import torch
from oracle import super_oracle as so
from super_b200 import engine, synth
H, W, step = 480, 640, 32
opt = so.default_opt(height=H, width=W, mesh_step_size=step, use_derived_gradient=False, mesh_face=True, optimizer="Adam")
tex = synth.texture(H, W)
trk = engine.Tracker(opt, device="cuda:0")
for t in range(1, 5):
    fr = synth.frame_inputs(t, H, W, tex=tex)
    trk.step(torch.from_numpy(fr["depth"]).cuda(), torch.from_numpy(fr["color"]).cuda(), torch.from_numpy(fr["K"]), torch.from_numpy(fr["inv_K"]), fr["time"])
torch.cuda.synchronize()
