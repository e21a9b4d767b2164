"""TEST INFRASTRUCTURE ONLY -- pin oracle/super_oracle.py against the unmodified reference.

Runs the reference (oracle/run_reference.py, CPU, shims) and the port on the same synthetic
sequence and compares EVERY array, stage by stage and teacher-forced (the port's pre-stage state is
replaced by the reference's).  Build-container only (needs /root/reference).

    python -m oracle.validate_port --height 120 --width 160 --step 16 --frames 4
"""
from __future__ import annotations

import argparse
import copy
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "python-super_b200"))

from oracle import run_reference, super_oracle as so  # noqa: E402
from super_b200 import synth  # noqa: E402


def state_from_snapshot(snap, opt, time_now):
    """Build the port's state namespace from a reference snapshot (numpy dict)."""
    sf = so.NS()
    for k in so.PER_SURFEL + ("isStable",):
        if k in snap:
            setattr(sf, k, torch.from_numpy(snap[k].copy()))
    ed = so.NS()
    for k, v in snap.items():
        if k.startswith("ED_"):
            setattr(ed, k[3:], torch.from_numpy(v.copy()))
    ed.num = len(ed.points)
    ed.param_num = 7 * ed.num
    sf.ED = ed
    sf.time = time_now
    sf.semantic = (opt.method == "semantic-super")
    sf.track_id = torch.from_numpy(snap["track_id"].copy()) if "track_id" in snap else None
    return sf


def newdata_from_snapshot(nd_snap, K):
    nd = so.NS()
    for k, v in nd_snap.items():
        setattr(nd, k, torch.from_numpy(v.copy()) if isinstance(v, np.ndarray) else v)
    nd.K = torch.from_numpy(K[0].copy())
    return nd


def cmp(name, a, b, tol=0.0, report=None):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    if a.shape != b.shape:
        msg = f"  {name}: SHAPE {a.shape} vs {b.shape}"
        ok = False
    elif a.dtype.kind in "iub" or tol == 0.0:
        nbad = int(np.count_nonzero(a != b)) if a.dtype.kind in "iub" else int(np.count_nonzero(~((a == b) | (np.isnan(a) & np.isnan(b)))))
        err = float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) if a.size else 0.0
        ok = nbad == 0
        msg = f"  {name}: {'bit-exact' if ok else f'{nbad} differ'} (max|d|={err:.3g})"
    else:
        err = float(np.abs(a - b).max()) if a.size else 0.0
        ok = err <= tol
        msg = f"  {name}: max|d|={err:.3g} (tol {tol:g}) {'ok' if ok else 'FAIL'}"
    print(msg)
    if report is not None:
        report.append((name, ok))
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=120)
    ap.add_argument("--width", type=int, default=160)
    ap.add_argument("--step", type=int, default=16)
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--assemble", default="sparse_mm")
    ap.add_argument("--speed", type=float, default=1.0)
    a = ap.parse_args()
    frames = list(range(1, a.frames + 1))
    rec = run_reference.run(["--mesh_step_size", str(a.step), "--sf_point_plane", "--mesh_rot", "--mesh_arap",
                             "--use_derived_gradient"], frames, a.height, a.width, record_matrices=True, speed=a.speed)
    opt = so.default_opt(height=a.height, width=a.width, mesh_step_size=a.step)
    report = []
    tex = synth.texture(a.height, a.width)
    prev_state = None
    free = so.Tracker(opt)
    for fr in rec.frames:
        t = fr["t"]
        print(f"== frame {t}")
        frame = synth.frame_inputs(t, a.height, a.width, tex=tex, speed=a.speed)
        nd = so.preprocess(opt, frame, ref_exp=True)     # with the reference's own float32 exp: everything bit-exact
        for k in ("points", "norms", "colors", "radii", "confs", "valid", "index_map"):
            cmp(f"new_data.{k}", getattr(nd, k), fr["new_data"][k], report=report)
        # with the oracle's DEFINED exponential (exp32_def, what the CUDA path is held to): same validity and points,
        # normals / confidences / radii within float32 rounding of the reference's
        nd_def = so.preprocess(opt, frame)
        for k in ("points", "valid", "index_map"):
            cmp(f"new_data[exp32_def].{k}", getattr(nd_def, k), fr["new_data"][k], report=report)
        cmp("new_data[exp32_def].norms", nd_def.norms, fr["new_data"]["norms"], tol=2.5e-7, report=report)
        cmp("new_data[exp32_def].confs", nd_def.confs.double(), fr["new_data"]["confs"].astype(np.float64), tol=1e-7, report=report)
        cmp("new_data[exp32_def].radii", nd_def.radii, fr["new_data"]["radii"], tol=1e-9, report=report)
        nb = int((nd_def.norms != nd.norms).any(1).sum())
        print(f"  exp32_def vs the reference's exp: {nb} of {len(nd.norms)} normals differ in the last float32 bit")
        nd_ref = newdata_from_snapshot(fr["new_data"], fr["K"])
        if prev_state is None:
            graph = so.build_graph(opt, nd_ref)
            sf = so.init_surfels(opt, nd_ref, graph)
        else:
            sf = state_from_snapshot(prev_state, opt, t - 1)
            trace = []
            beta = so.lm_solve(opt, sf, nd_ref, assemble=a.assemble, trace=trace)
            for i, (it, rit) in enumerate(zip(trace, fr["lm_iters"])):
                dp = fr["data_passes"][2 * i]
                ok_ids = cmp(f"it{i} matched ids", it["ids"].to(torch.int32), dp["ids"], report=report)
                cmp(f"it{i} corners", it["corners"].to(torch.int32), dp["corners"], report=report)
                cmp(f"it{i} A", it["A"], rit["jtj"], tol=1e-9 * float(np.abs(rit["jtj"]).max()), report=report)
                cmp(f"it{i} g", it["g"], rit["jtl"], tol=1e-12, report=report)
                cmp(f"it{i} delta", it["delta"].reshape(-1, 1), rit["delta"], tol=1e-9, report=report)
                rel = abs(it["loss"] - rit["loss"]) / rit["loss"]
                print(f"  it{i} loss {it['loss']:.12e} ref {rit['loss']:.12e} rel {rel:.2e} accept {it['accept']}")
                report.append((f"it{i} loss", rel < 1e-9))
            cmp("beta", beta, fr["beta"], tol=1e-10, report=report)
            # teacher-forced update / fuse / compact
            beta_ref = torch.from_numpy(fr["beta"].copy())
            sf = state_from_snapshot(prev_state, opt, t - 1)
            so.update(opt, sf, beta_ref)
            for k in ("points", "norms"):
                cmp(f"update.{k}", getattr(sf, k), fr["after_update"][k], report=report)
            cmp("update.ED.points", sf.ED.points, fr["after_update"]["ED_points"], report=report)
            cmp("update.ED.norms", sf.ED.norms, fr["after_update"]["ED_norms"], report=report)
            so.fuse(opt, sf, nd_ref)
            for k in so.PER_SURFEL + ("isStable",):
                if k in fr["after_fuse"]:
                    cmp(f"fuse.{k}", getattr(sf, k), fr["after_fuse"][k], report=report)
            so.compact(opt, sf, float(t))
        for k in so.PER_SURFEL + ("isStable",):
            if k in fr["state"]:
                cmp(f"state.{k}", getattr(sf, k), fr["state"][k], report=report)
        for k in ("points", "norms", "radii", "knn_indices", "knn_w", "edge_index", "triangles",
                  "triangles_areas"):
            cmp(f"state.ED.{k}", getattr(sf.ED, k), fr["state"]["ED_" + k], report=report)
        prev_state = fr["state"]
        # free-running port (no teacher forcing): order-invariant statistics only
        free.step(frame)
        print(f"  free-running: N port {len(free.sf.points)} ref {len(fr['state']['points'])}")
    bad = [n for n, ok in report if not ok]
    print(f"\n{len(report) - len(bad)}/{len(report)} checks passed")
    if bad:
        print("FAILED:", bad[:40])
        sys.exit(1)


if __name__ == "__main__":
    main()
