"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Build-container only (imports /root/reference through oracle/ref_shims.py).  The fixtures it
writes are committed; the GPU box and the CPU test-suite read the .npz files, never the reference.

    python -m oracle.gen_golden            # writes tests/golden/lm_128x96.npz (+ manifest)

What a fixture pins (all produced by the reference's own code path, run_super.py flags
--load_depth --mesh_step_size S --sf_point_plane --mesh_rot --mesh_arap --use_derived_gradient):
  * the per-frame producer output (new_data points/norms, stored as float32: they are f32-exact),
  * the full Surfels + ED-node state after the init frame and after every tracked frame,
  * per LM iteration: trial beta, delta, damping u, total and per-term loss, accept flag,
    matched surfel ids and bilinear corner ids (integers) for iterations 0, 1 and the last,
  * the dense normal equations (A, g) of the first iteration of the first tracked frame,
  * surfel / node positions and normals after Surfels.update, and isStable after fuseInputData.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

from oracle import run_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

LM_FLAGS = ["--sf_point_plane", "--mesh_rot", "--mesh_arap", "--use_derived_gradient"]


def pack_state(d, prefix, snap):
    for k, v in snap.items():
        if v.dtype == np.int64 and k != "track_id":
            v = v.astype(np.int32)
        d[f"{prefix}.{k}"] = v


def gen_lm(name, height, width, step, nframes, speed):
    frames = list(range(1, nframes + 1))
    rec = run_reference.run(["--mesh_step_size", str(step)] + LM_FLAGS, frames, height, width,
                            record_matrices=True, speed=speed)
    d = {}
    meta = {"height": height, "width": width, "step": step, "frames": frames, "speed": speed,
            "flags": LM_FLAGS, "reference": "ucsdarclab/Python-SuPer @ /root/reference (unmodified, CPU, shims)"}
    for fr in rec.frames:
        t = fr["t"]
        nd = fr["new_data"]
        for k in ("points", "norms"):
            a32 = nd[k].astype(np.float32)
            assert np.array_equal(a32.astype(np.float64), nd[k]), "new_data not f32-exact"
            d[f"f{t}.nd.{k}"] = a32
        d[f"f{t}.nd.valid"] = np.packbits(nd["valid"])
        d[f"f{t}.nd.radii"] = nd["radii"]
        d[f"f{t}.nd.confs"] = nd["confs"]
        pack_state(d, f"f{t}.state", fr["state"])
        if not fr["lm_iters"]:
            continue
        its = fr["lm_iters"]
        d[f"f{t}.lm.loss"] = np.array([it["loss"] for it in its])
        d[f"f{t}.lm.loss_terms"] = np.array([it["loss_terms"] for it in its])
        d[f"f{t}.lm.u"] = np.array([it["u"] for it in its])
        d[f"f{t}.lm.beta_try"] = np.stack([it["beta_try"] for it in its])
        d[f"f{t}.lm.delta"] = np.stack([it["delta"].reshape(-1, 7) for it in its])
        d[f"f{t}.lm.g"] = np.stack([it["jtl"][:, 0] for it in its])
        d[f"f{t}.lm.A_diag"] = np.stack([np.diagonal(it["jtj"]) for it in its])
        d[f"f{t}.beta"] = fr["beta"]
        grad_passes = fr["data_passes"][0::2]
        d[f"f{t}.lm.M"] = np.array([len(p["ids"]) for p in grad_passes])
        for i in (0, 1, len(its) - 1):
            d[f"f{t}.lm.it{i}.ids"] = grad_passes[i]["ids"]
            d[f"f{t}.lm.it{i}.corners"] = grad_passes[i]["corners"].astype(np.int16)
        if t == frames[1]:
            d[f"f{t}.lm.it0.A"] = its[0]["jtj"]
        au = fr["after_update"]
        d[f"f{t}.update.points"] = au["points"]
        d[f"f{t}.update.norms"] = au["norms"]
        d[f"f{t}.update.ED_points"] = au["ED_points"]
        d[f"f{t}.update.ED_norms"] = au["ED_norms"]
        af = fr["after_fuse"]
        d[f"f{t}.fuse.isStable"] = np.packbits(af["isStable"])
        d[f"f{t}.fuse.N"] = np.array(len(af["isStable"]))
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=json.dumps(meta), **d)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB, {len(d)} arrays")
    for fr in rec.frames[1:]:
        print(f"  frame {fr['t']}: losses {[f'{it['loss']:.4e}' for it in fr['lm_iters']][:4]}... "
              f"N={len(fr['state']['points'])}")


if __name__ == "__main__":
    gen_lm("lm_128x96", 96, 128, 16, 4, 3.0)
