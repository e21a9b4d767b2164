"""GPU probe: BASELINE.json configs 3 (step 16, LM vs Adam) and 5 (1280x1024) -- frames/s of the device tracker."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-super_b200"))
import torch
from oracle import super_oracle as so
from super_b200 import engine, synth

def run(H, W, step, lm, optimizer="Adam", frames=12, warm=3):
    opt = so.default_opt(height=H, width=W, mesh_step_size=step, use_derived_gradient=lm, mesh_face=not lm, optimizer=optimizer)
    tex = synth.texture(H, W)
    trk = engine.Tracker(opt, device="cuda:0")
    fr = [synth.frame_inputs(t, H, W, tex=tex) for t in range(1, frames + 2)]
    dd = [torch.from_numpy(f["depth"]).cuda() for f in fr]
    dc = torch.from_numpy(fr[0]["color"]).cuda()
    K, iK = torch.from_numpy(fr[0]["K"]), torch.from_numpy(fr[0]["inv_K"])
    trk.step(dd[0], dc, K, iK, fr[0]["time"])
    for i in range(1, 1 + warm):
        trk.step(dd[i], dc, K, iK, fr[i]["time"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(1 + warm, frames + 1):
        trk.step(dd[i], dc, K, iK, fr[i]["time"])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (frames - warm)
    out = {"H": H, "W": W, "step": step, "solver": "LM" if lm else optimizer, "J": int(trk.ED.num), "surfels": trk.num_surfels(),
           "ms_per_frame": round(ms, 3), "fps": round(1e3 / ms, 1), "band": None if trk.band is None else int(trk.band.bw)}
    if lm:
        st = trk.ws.state.read()
        out["last_losses"] = [float(x) for x in st["loss"][-3:]]
        out["failed"] = int(st["failed"])
    else:
        out["last_loss"] = trk.gf_ws.read_trace(10)[-1]
    return out

res = []
for cfg in ((480, 640, 32, True), (480, 640, 32, False), (480, 640, 16, True), (480, 640, 16, False), (1024, 1280, 32, True)):
    try:
        r = run(*cfg)
    except Exception as e:
        r = {"cfg": cfg, "error": repr(e)[:300]}
    print(json.dumps(r), flush=True)
    res.append(r)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_configs.json"), "w"), indent=1)
