import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-super_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
from oracle import super_oracle as so
from super_b200 import synth, engine
for (H, W, step, speed) in ((96, 128, 16, 3.0), (480, 640, 32, 1.0)):
    opt = so.default_opt(height=H, width=W, mesh_step_size=step)
    f = synth.frame_inputs(1, H, W, speed=speed)
    nd = so.preprocess(opt, f)
    fr = engine.preprocess(opt, torch.from_numpy(f["depth"]).cuda(), torch.from_numpy(f["color"]).cuda(),
                           torch.from_numpy(f["K"]), torch.from_numpy(f["inv_K"]), f["time"])
    valid = fr.valid.cpu()
    print(H, W, "valid equal", torch.equal(valid, nd.valid), int(valid.sum()), int(nd.valid.sum()))
    both = valid & nd.valid
    pts = fr.vmap.cpu()[:, :3].double()
    ref = torch.zeros_like(pts); ref[nd.valid] = nd.points
    d = (pts - ref)[both]
    rel = d.abs() / ref[both].abs().clamp_min(1e-30)
    print(" points max abs diff per comp", d.abs().max(0).values.tolist(), "max rel", rel.max(0).values.tolist(),
          "n differing", (d != 0).sum(0).tolist())
    nref = torch.zeros_like(pts); nref[nd.valid] = nd.norms
    print(" norms max diff", (fr.nmap.cpu()[:, :3].double() - nref)[both].abs().max().item())
    print(" inv_K", f["inv_K"][:3, :3].tolist())
