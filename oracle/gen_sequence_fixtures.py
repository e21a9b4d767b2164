"""TEST INFRASTRUCTURE -- free-running sequence fixtures at BASELINE.json's OWN sizes, precomputed on the CPU by the
oracle port (which is pinned to the unmodified reference: oracle/validate_port.py), so that `-m gpu` tests can check the
device tracker per frame without spending minutes of CPU per frame on the GPU box.

    python -m oracle.gen_sequence_fixtures c2 [n_tracked=200]     BASELINE config 2: 640x480, step 32, LM, bench sequence
    python -m oracle.gen_sequence_fixtures c3a [3]                config 3 (LM):   640x480, step 16 (J ~ 1.1 k, n ~ 7.9 k)
    python -m oracle.gen_sequence_fixtures c3b [3]                config 3 (Adam): 640x480, step 16, autograd path
    python -m oracle.gen_sequence_fixtures c4 [3]                 config 4: Semantic-SuPer losses, superv2, SGD, 640x480
    python -m oracle.gen_sequence_fixtures c5 [2]                 config 5: 1280x1024, step 32, LM

Writes tests/golden/seq_<name>.npz:
    meta            json: sizes, option overrides, sequence kind, tracked labels
    depth_crc       (F,)  u32  zlib.crc32 of every input depth image (the GPU test refuses to compare different inputs)
    n_surfels       (F,)  i64  model size after every frame (frame 0 = init)
    loss            (F-1, I) f64  per-iteration loss (LM: trial loss; autograd: total loss before the step)
    accept          (F-1, I) i8   LM accept/reject per iteration (absent for the autograd configs)
    beta            (F-1, J[+1], 7) f64  the frame's result
    track           (F, T, 3) f32 tracked-point reprojections [u, v, 1]  (LM configs), track_id (F, T) i64
F = n_tracked + 1.  The sequence is regenerated from super_b200.synth on the GPU box (same numpy code).
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import time
import zlib

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "python-super_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from oracle import super_oracle as so          # noqa: E402
from oracle import graphfit_oracle as gfo      # noqa: E402
from super_b200 import synth                   # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

SEM_OPT = {"use_derived_gradient": False, "mesh_face": True, "mesh_arap": False, "sf_point_plane": False,
           "optimizer": "SGD", "method": "semantic-super", "data": "superv2", "num_classes": 3,
           "sf_soft_seg_point_plane": True, "sf_hard_seg_point_plane": False, "sf_bn_morph": True,
           "sf_bn_morph_weight": 0.1, "hard_seg": False, "del_seg_classes": [], "disable_ssim_conf": True}

CONFIGS = {
    # name: (H, W, step, option overrides, sequence kind, default tracked frames)
    "c2": (480, 640, 32, {}, "bench", 200),
    "c3a": (480, 640, 16, {}, "drift", 3),
    "c3b": (480, 640, 16, {"use_derived_gradient": False, "mesh_face": True, "optimizer": "Adam"}, "drift", 3),
    "c4": (480, 640, 32, SEM_OPT, "drift_seg", 3),
    "c5": (1024, 1280, 32, {}, "drift", 2),
}


SEG_SPEED = 3.0


def track_labels(H, W, n_frames):
    """Ten labelled pixels on a diagonal (same on every frame): the --tracking_gt_file of the check."""
    pts = np.array([[int((100 + 40 * i) * W / 640), int((80 + 30 * i) * H / 480), 1] for i in range(10)], dtype=np.int64)
    return {f"{t:06d}": pts for t in range(1, n_frames + 1)}


def sequence(kind, H, W, n_frames, data="superv1"):
    """Generator of the host frames.  'bench': bench.py's stationary sequence (oscillating surface, running frame
    time); 'drift': SURVEY 8(d)'s formula with t = 1, 2, ...; 'drift_seg': the same with class scores."""
    tex = synth.texture(H, W)
    if kind == "bench":
        spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
        bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench)
        for i in range(n_frames):
            f = synth.frame_inputs(bench.shape_time(i), H, W, tex=tex)
            f["time"], f["ID"], f["filename"] = float(i + 1), i + 1, f"{i + 1:06d}"
            yield f
    else:
        for t in range(1, n_frames + 1):
            # class boundaries move 3x faster than SURVEY's formula (9 and 6 px per frame at 640 wide) so that the
            # boundary-morph term has surfels to act on in every iteration (its mean over an empty set is NaN)
            yield synth.frame_inputs(t, H, W, data=data, tex=tex, with_seg=(kind == "drift_seg"), seg_speed=SEG_SPEED)


def crc(frame):
    return zlib.crc32(np.ascontiguousarray(frame["depth"]).tobytes()) & 0xffffffff


def run(name, n_tracked=None, threads=None, out=None):
    H, W, step, over, kind, n_def = CONFIGS[name]
    n_tracked = n_def if n_tracked is None else int(n_tracked)
    torch.set_num_threads(threads or os.cpu_count())
    opt = so.default_opt(height=H, width=W, mesh_step_size=step, **over)
    lm = bool(opt.use_derived_gradient)
    F = n_tracked + 1
    gt = track_labels(H, W, F) if lm else None
    d = {"depth_crc": [], "n_surfels": [], "loss": [], "accept": [], "beta": [], "track": [], "track_id": [], "cpu_s": []}
    out = out or os.path.join(GOLDEN, f"seq_{name}.npz")
    meta = {"name": name, "height": H, "width": W, "step": step, "opt": over, "sequence": kind, "tracked": n_tracked,
            "seg_speed": SEG_SPEED if kind == "drift_seg" else None, "lm": lm, "iterations": opt.num_optimize_iterations,
            "oracle": "oracle/super_oracle.py (+ graphfit_oracle.py), free running, exp32_def producer"}

    def save():
        arrs = {"meta": json.dumps(meta), "depth_crc": np.array(d["depth_crc"], dtype=np.uint32),
                "n_surfels": np.array(d["n_surfels"], dtype=np.int64), "loss": np.array(d["loss"], dtype=np.float64),
                "beta": np.array(d["beta"], dtype=np.float64), "cpu_s": np.array(d["cpu_s"], dtype=np.float32)}
        if lm:
            arrs["accept"] = np.array(d["accept"], dtype=np.int8)
            arrs["track"] = np.array(d["track"], dtype=np.float32)
            arrs["track_id"] = np.array(d["track_id"], dtype=np.int64)
            arrs["gt"] = gt["000001"]
        np.savez_compressed(out, **arrs)

    trk = so.Tracker(opt, gt=gt) if lm else None
    sf = None
    for i, fr in enumerate(sequence(kind, H, W, F, data=opt.data)):
        t0 = time.perf_counter()
        d["depth_crc"].append(crc(fr))
        if lm:
            beta = trk.step(fr, trace=True)
            sf = trk.sf
            if beta is not None:
                d["loss"].append([it["loss"] for it in trk.trace])
                d["accept"].append([int(it["accept"]) for it in trk.trace])
                d["beta"].append(beta.numpy().copy())
            d["track"].append(trk.track_rsts[fr["filename"]].numpy().astype(np.float32))
            d["track_id"].append(sf.track_id.numpy().copy())
        else:
            nd = so.preprocess(opt, fr)
            if sf is None:
                sf = so.init_surfels(opt, nd, so.build_graph(opt, nd))
            else:
                tr = []
                dv = gfo.graph_fit(opt, sf, nd, trace=tr)
                so.update(opt, sf, dv)
                so.fuse(opt, sf, nd)
                so.compact(opt, sf, float(fr["time"]))
                d["loss"].append([x["loss"] for x in tr])
                d["beta"].append(dv.numpy().copy())
        d["n_surfels"].append(len(sf.points))
        d["cpu_s"].append(time.perf_counter() - t0)
        meta["J"] = int(sf.ED.num)
        print(f"{name} frame {i}/{n_tracked}: N = {len(sf.points)}, {d['cpu_s'][-1]:.1f} s", flush=True)
        if i % 10 == 0 or i == F - 1:
            save()
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    a = sys.argv[1:]
    run(a[0], a[1] if len(a) > 1 else None, int(a[2]) if len(a) > 2 else None)
