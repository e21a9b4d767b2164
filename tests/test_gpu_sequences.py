"""GPU: free-running sequences at BASELINE.json's own sizes against CPU-precomputed oracle fixtures
(tests/golden/seq_*.npz, written by oracle/gen_sequence_fixtures.py from the port that is pinned to the unmodified
reference), plus run-to-run bitwise reproducibility of the device tracker.

north_star's tolerances, checked on EVERY frame: per-iteration LM residual (loss) 1e-4 relative, node
quaternions/translations (beta) 1e-4, tracked-point reprojection 0.1 px.  Two free-running trackers can take different
accept/reject decisions at a converged iteration (trial loss and best loss then differ by ~1e-9 relative while the two
implementations differ by ~1e-13): from that iteration on their loss traces compare a rejected trial with an accepted
one.  Such iterations are COUNTED and must coincide with a differing decision; everything else is a hard failure.
"""
import json
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import super_oracle as so
from oracle import gen_sequence_fixtures as gsf

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_LOSS, TOL_BETA, TOL_PX = 1e-4, 1e-4, 0.1


def _load(name):
    z = np.load(os.path.join(GOLDEN, f"seq_{name}.npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta


def _run_sequence(name, max_frames=None):
    """Device tracker over the fixture's sequence; returns per-frame comparison rows."""
    from super_b200 import engine
    z, meta = _load(name)
    H, W = meta["height"], meta["width"]
    opt = so.default_opt(height=H, width=W, mesh_step_size=meta["step"], **meta["opt"])
    lm = meta["lm"]
    F = len(z["n_surfels"]) if max_frames is None else min(len(z["n_surfels"]), max_frames + 1)
    trk = engine.Tracker(opt, device="cuda:0")
    if lm:
        gt = {f"{t:06d}": z["gt"] for t in range(1, F + 1)}
        trk.enable_tracking(gt)
    rows = []
    for i, fr in enumerate(gsf.sequence(meta["sequence"], H, W, F, data=opt.data)):
        crc = zlib.crc32(np.ascontiguousarray(fr["depth"]).tobytes()) & 0xffffffff
        assert crc == int(z["depth_crc"][i]), f"frame {i}: synthetic input differs from the one the fixture was made from"
        seg = torch.from_numpy(fr["seg_conf"]).cuda() if "seg_conf" in fr else None
        beta = trk.step(torch.from_numpy(fr["depth"]).cuda(), torch.from_numpy(fr["color"]).cuda(),
                        torch.from_numpy(fr["K"]), torch.from_numpy(fr["inv_K"]), fr["time"], seg_scores=seg,
                        filename=fr["filename"])
        row = {"frame": i, "surfels": trk.num_surfels(), "surfels_ref": int(z["n_surfels"][i])}
        if lm:
            row["track_id_equal"] = bool(np.array_equal(trk.track_id.cpu().numpy(), z["track_id"][i]))
            row["track_err_px"] = float(np.abs(trk.track_rsts[fr["filename"]].cpu().numpy() - z["track"][i]).max())
        if beta is not None:
            ref_loss = z["loss"][i - 1]
            if lm:
                st = trk.ws.state.read()
                loss, acc = st["loss"], st["accept"]
                assert st["failed"] == 0
                row["accept_equal"] = bool(np.array_equal(acc, z["accept"][i - 1]))
                row["first_flip"] = int(np.argmax(acc != z["accept"][i - 1])) if not row["accept_equal"] else -1
            else:
                loss = np.array([x["total"] for x in trk.gf_ws.read_trace(opt.num_optimize_iterations)])
            both_nan = np.isnan(loss) & np.isnan(ref_loss)
            rel = np.where(both_nan, 0.0, np.abs(loss - ref_loss) / np.abs(ref_loss))
            row["loss_rel_err"] = [float(x) for x in rel]
            row["beta_err"] = float(np.abs(beta.cpu().numpy() - z["beta"][i - 1]).max())
        rows.append(row)
    return rows, meta


def _check(rows, meta, name):
    """Hard criteria on every frame + the flip-attributed loss excursions.  Writes the per-frame table to gpurun_out/."""
    lm = meta["lm"]
    flips_seen = False
    excursions, unexplained = [], []
    for r in rows:
        assert abs(r["surfels"] - r["surfels_ref"]) <= max(2, 1e-4 * r["surfels_ref"]), r
        if lm:
            assert r["track_id_equal"], r
            assert r["track_err_px"] < TOL_PX, r
        if "beta_err" not in r:
            continue
        assert r["beta_err"] < TOL_BETA, r
        if lm and not r["accept_equal"]:
            flips_seen = True
        for it, e in enumerate(r["loss_rel_err"]):
            if not (e < TOL_LOSS):                 # also catches NaN on one side only
                # attributable iff this frame's accept/reject pattern differs at or before this iteration
                ok = lm and (not r["accept_equal"]) and r["first_flip"] <= it
                (excursions if ok else unexplained).append((r["frame"], it, e))
    out = {"fixture": name, "frames_tracked": sum("beta_err" in r for r in rows),
           "max_beta_err": max((r["beta_err"] for r in rows if "beta_err" in r), default=0.0),
           "max_track_err_px": max((r.get("track_err_px", 0.0) for r in rows), default=0.0),
           "max_loss_rel_err_outside_flips": max((e for r in rows if "beta_err" in r and r.get("accept_equal", True)
                                                  for e in r["loss_rel_err"]), default=0.0),
           "frames_with_different_accept_pattern": [r["frame"] for r in rows if lm and "beta_err" in r and not r["accept_equal"]],
           "loss_excursions_after_a_flip": excursions, "unexplained_loss_excursions": unexplained,
           "surfel_count_equal_frames": sum(r["surfels"] == r["surfels_ref"] for r in rows), "frames": len(rows),
           "max_surfel_count_diff": max(abs(r["surfels"] - r["surfels_ref"]) for r in rows)}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"summary": out, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", f"seq_{name}_check.json"), "w"))
    print(json.dumps(out))
    assert not unexplained, f"loss outside 1e-4 without a differing accept/reject decision: {unexplained[:5]}"
    return out


def test_config2_200_frames_lm_640x480():
    """BASELINE config 2: 200 tracked frames, 640x480, step 32, LM x10, per-frame tolerance check."""
    mx = os.environ.get("SB_SEQ_MAX_FRAMES")          # development runs only
    rows, meta = _run_sequence("c2", int(mx) if mx else None)
    out = _check(rows, meta, "c2")
    assert mx or out["frames_tracked"] >= 200


def test_config3a_lm_step16_640x480():
    """BASELINE config 3 (LM): J = 1131 nodes, n = 7917, band ~640: the wide-band path of the solver at its own size."""
    rows, meta = _run_sequence("c3a")
    out = _check(rows, meta, "c3a")
    assert meta["J"] > 1000 and out["frames_tracked"] >= 3


def test_config3b_adam_step16_640x480():
    rows, meta = _run_sequence("c3b")
    _check(rows, meta, "c3b")


def test_config4_semantic_640x480():
    rows, meta = _run_sequence("c4")
    _check(rows, meta, "c4")


def test_config5_lm_1280x1024():
    """BASELINE config 5: 1280x1024, ~1.3 M surfels, J = 1209, n = 8463."""
    rows, meta = _run_sequence("c5")
    out = _check(rows, meta, "c5")
    assert out["frames_tracked"] >= 2


@pytest.mark.parametrize("H,W,step,frames", [(96, 128, 16, 6), (480, 640, 32, 8)])
def test_tracker_is_bitwise_reproducible(H, W, step, frames):
    """Two runs of the device tracker over the same frames give bit-identical state: the normal equations are
    accumulated with integer atomics (order-independent), every other reduction has a fixed order."""
    from super_b200 import engine, synth

    def run():
        opt = so.default_opt(height=H, width=W, mesh_step_size=step)
        trk = engine.Tracker(opt, device="cuda:0")
        tex = synth.texture(H, W)
        betas, losses = [], []
        for t in range(1, frames + 1):
            fr = synth.frame_inputs(t, H, W, tex=tex, speed=3.0 if W < 640 else 1.0)
            b = trk.step(torch.from_numpy(fr["depth"]).cuda(), torch.from_numpy(fr["color"]).cuda(),
                         torch.from_numpy(fr["K"]), torch.from_numpy(fr["inv_K"]), fr["time"])
            if b is not None:
                betas.append(b.clone())
                losses.append(trk.ws.state.read()["loss"].copy())
        return betas, losses, trk.snapshot()

    b1, l1, s1 = run()
    b2, l2, s2 = run()
    for x, y in zip(b1, b2):
        assert torch.equal(x, y)
    for x, y in zip(l1, l2):
        assert np.array_equal(x, y)
    assert s1.keys() == s2.keys()
    for k in s1:
        assert torch.equal(s1[k], s2[k]), k


def test_frame_loop_equals_stepwise_loop(monkeypatch):
    """sb_lm_frame (one J^T J pass per iteration, loss from the Gram panels, kept system after a reject) against the
    stage-by-stage loop (J^T J at the current beta + loss-only pass at the trial beta, LM.py:93-117 literally)."""
    from golden_util import Golden
    from gpu_util import to_device_state, device_maps, camera
    from super_b200 import lm, ops
    G = Golden()
    for t in G.frames[1:]:
        sf, nd = G.state(t - 1), G.new_data(t)
        J = sf.ED.num
        d = to_device_state(sf)
        maps = device_maps(nd, G.H, G.W)
        cam = camera(nd, G.H, G.W)
        pos = torch.arange(J, dtype=torch.int32, device="cuda")
        order = ops.tuple_order(d.knn_indices, None, pos, None)
        res = []
        for stepwise in ("0", "1"):
            monkeypatch.setenv("SB_LM_STEPWISE", stepwise)
            band = ops.Band(7 * J, 7 * (J - 1) + 6, pos, "cuda")
            beta, ws = lm.lm_solve(d, maps, cam, G.opt, order=order, band=band, cluster_size=16)
            res.append((beta.clone().cpu(), ws.state.read()))
        (b0, s0), (b1, s1) = res
        assert np.array_equal(s0["accept"], s1["accept"])
        assert np.abs(s0["loss"] - s1["loss"]).max() <= 1e-11 * s1["loss"].max()
        assert np.abs(s0["u_trace"] - s1["u_trace"]).max() == 0.0
        assert (b0 - b1).abs().max() < 1e-11


def test_graph_replay_equals_stream_launches(monkeypatch):
    """The LM loop and the frame tail replayed from CUDA graphs (sb_lm_frame's graph cache, lib.graph_scope /
    Tracker.tail_scope) against the same launches issued one by one on the stream: bit-identical state, frame by frame."""
    from super_b200 import engine, synth
    H, W, step, frames = 480, 640, 32, 7

    def run(graphs):
        monkeypatch.setenv("SB_LM_GRAPH", "1" if graphs else "0")
        monkeypatch.setenv("SB_TAIL_GRAPH", "1" if graphs else "0")
        opt = so.default_opt(height=H, width=W, mesh_step_size=step)
        trk = engine.Tracker(opt, device="cuda:0")
        tex = synth.texture(H, W)
        betas = []
        for t in range(1, frames + 1):
            fr = synth.frame_inputs(t, H, W, tex=tex)
            b = trk.step(torch.from_numpy(fr["depth"]).cuda(), torch.from_numpy(fr["color"]).cuda(),
                         torch.from_numpy(fr["K"]), torch.from_numpy(fr["inv_K"]), fr["time"])
            if b is not None:
                betas.append(b.clone())
        return betas, trk.snapshot()

    b1, s1 = run(True)
    b0, s0 = run(False)
    assert len(b1) == len(b0) == frames - 1
    for x, y in zip(b1, b0):
        assert torch.equal(x, y)
    for k in s0:
        assert torch.equal(s1[k], s0[k]), k


@pytest.mark.parametrize("J,n", [(266, 50000), (35, 4097), (1209, 30000), (65535, 2000)])
def test_tuple_order_groups_like_a_full_key_sort(J, n):
    """sb_tuple_order (keys with ceil(log2(J+1)) bits per node id, radix sort over just those bits) against a stable sort
    of the full (ascending node set) keys: same visiting order, rows beyond the device-side count last."""
    from super_b200 import ops
    g = torch.Generator().manual_seed(J + n)
    idx = torch.stack([torch.randperm(J, generator=g)[:4] for _ in range(n)]).to(torch.int32)
    # many equal node sets in different slot orders, as in the tracker
    idx[n // 2:] = idx[: n - n // 2][:, torch.tensor([2, 0, 3, 1])]
    n_live = n - 37
    n_dev = torch.tensor([n_live], dtype=torch.int32, device="cuda")
    pos = torch.arange(J, dtype=torch.int32, device="cuda")
    order = ops.tuple_order(idx.cuda(), n_dev, pos, None).cpu().long()
    srt = np.sort(idx.numpy().astype(np.int64), axis=1)[:n_live]
    ref = torch.from_numpy(np.lexsort((srt[:, 3], srt[:, 2], srt[:, 1], srt[:, 0])))      # stable, first column primary
    assert torch.equal(order[:n_live], ref)
    assert sorted(order[n_live:].tolist()) == list(range(n_live, n))
