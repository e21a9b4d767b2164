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


def gen_gf(name, height, width, step, nframes, speed, flags, opt_over, semantic=False, seg_speed=None):
    """Autograd optimiser (GraphFit) fixtures: per tracked frame the 10 iterations' deform_verts, loss terms and the
    gradient the optimiser consumed, the returned deform_verts, the state after Surfels.update and after the frame."""
    frames = list(range(1, nframes + 1))
    rec = run_reference.run(["--mesh_step_size", str(step)] + flags, frames, height, width, with_seg=semantic,
                            semantic=semantic, speed=speed, seg_speed=seg_speed)
    d = {}
    meta = {"height": height, "width": width, "step": step, "frames": frames, "speed": speed, "seg_speed": seg_speed,
            "semantic": semantic, "flags": flags, "opt": opt_over,
            "reference": "ucsdarclab/Python-SuPer @ /root/reference (unmodified, CPU, shims)"}
    for fr in rec.frames:
        t = fr["t"]
        nd = fr["new_data"]
        for k in ("points", "norms"):
            a32 = nd[k].astype(np.float32)
            assert np.array_equal(a32.astype(np.float64), nd[k]), "new_data not f32-exact"
            d[f"f{t}.nd.{k}"] = a32
        d[f"f{t}.nd.valid"] = np.packbits(nd["valid"])
        pack_state(d, f"f{t}.state", fr["state"])
        its = fr.get("ag_iters")
        if not its:
            continue
        d[f"f{t}.ag.deform_in"] = np.stack([it["deform_in"] for it in its])
        d[f"f{t}.ag.loss"] = np.array([it["loss"] for it in its])
        d[f"f{t}.ag.grad"] = np.stack([it["grad"] for it in its])
        for k in sorted({k for it in its for k in it["losses"]}):
            d[f"f{t}.ag.losses.{k}"] = np.array([it["losses"].get(k, np.nan) for it in its])
        d[f"f{t}.beta"] = fr["beta"]
        au = fr["after_update"]
        for k in ("points", "norms", "ED_points", "ED_norms"):
            d[f"f{t}.update.{k}"] = au[k]
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=json.dumps(meta), **d)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB, {len(d)} arrays")
    for fr in rec.frames[1:]:
        its = fr["ag_iters"]
        print(f"  frame {fr['t']}: loss {its[0]['loss']:.6e} -> {its[-1]['loss']:.6e}  terms {its[-1]['losses']}  "
              f"N={len(fr['state']['points'])}")


TRACK_PTS = [[20, 20, 1], [40, 30, 1], [64, 48, 1], [90, 60, 1], [110, 80, 1], [0, 0, 1], [30, 70, 0]]


def gen_track(name, height, width, step, nframes, speed):
    """Tracked-point bookkeeping (nodes.py:225-265,443-458,576-599): per frame the tracked surfel ids and the
    recorded reprojections track_rsts[filename] (T,3), for --tracking_gt_file labels given on every frame."""
    frames = list(range(1, nframes + 1))
    pts = np.array(TRACK_PTS, dtype=np.int64)
    gt = {"gt": {f"{t:06d}": pts.copy() for t in frames}}
    rec = run_reference.run(["--mesh_step_size", str(step)] + LM_FLAGS, frames, height, width, speed=speed, tracking_gt=gt)
    sf = rec.models.super.sf
    d = {"gt": pts}
    for fr in rec.frames:
        t = fr["t"]
        d[f"f{t}.track_id"] = fr["state"]["track_id"]
        d[f"f{t}.track_rsts"] = sf.track_rsts[f"{t:06d}"].cpu().numpy()
        d[f"f{t}.N"] = np.array(len(fr["state"]["points"]))
    meta = {"height": height, "width": width, "step": step, "frames": frames, "speed": speed, "flags": LM_FLAGS,
            "reference": "ucsdarclab/Python-SuPer @ /root/reference (unmodified, CPU, shims)"}
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=json.dumps(meta), **d)
    print(f"wrote {path}: {os.path.getsize(path) / 1e3:.1f} KB; ids per frame {[d[f'f{t}.track_id'].tolist() for t in frames]}")


GF_FLAGS = ["--sf_point_plane", "--mesh_rot", "--mesh_arap", "--mesh_face", "--optimizer", "Adam"]
GF_OPT = dict(use_derived_gradient=False, mesh_face=True, optimizer="Adam")
SEM_FLAGS = ["--load_seg", "--seg_dir", "seg", "--disable_ssim_conf", "--sf_soft_seg_point_plane", "--mesh_rot",
             "--mesh_face", "--sf_bn_morph"]
SEM_OPT = dict(use_derived_gradient=False, mesh_face=True, mesh_arap=False, sf_point_plane=False, optimizer="SGD",
               method="semantic-super", data="superv2", num_classes=3, sf_soft_seg_point_plane=True,
               sf_hard_seg_point_plane=False, sf_bn_morph=True, sf_bn_morph_weight=0.1, hard_seg=False,
               del_seg_classes=[], disable_ssim_conf=True)


HARD_FLAGS = ["--load_seg", "--seg_dir", "seg", "--disable_ssim_conf", "--hard_seg", "--sf_hard_seg_point_plane", "--mesh_rot",
              "--mesh_face", "--mesh_arap"]
HARD_OPT = dict(use_derived_gradient=False, mesh_face=True, mesh_arap=True, sf_point_plane=False, optimizer="SGD",
                method="semantic-super", data="superv2", num_classes=3, sf_soft_seg_point_plane=False,
                sf_hard_seg_point_plane=True, sf_bn_morph=False, hard_seg=True, del_seg_classes=[], disable_ssim_conf=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["lm", "gf", "sem", "hard", "track"]
    if "lm" in which:
        gen_lm("lm_128x96", 96, 128, 16, 4, 3.0)
    if "gf" in which:
        gen_gf("gf_128x96", 96, 128, 16, 3, 3.0, GF_FLAGS, GF_OPT)
    if "track" in which:
        gen_track("track_128x96", 96, 128, 16, 4, 3.0)
    if "hard" in which:
        gen_gf("gf_hard_128x96", 96, 128, 16, 3, 3.0, HARD_FLAGS, HARD_OPT, semantic=True, seg_speed=15.0)
    if "sem" in which:
        gen_gf("gf_sem_128x96", 96, 128, 16, 3, 3.0, SEM_FLAGS, SEM_OPT, semantic=True, seg_speed=15.0)
