"""GPU parity of the frame pipeline around the LM solve: producer, fusion, compaction, whole frames."""
import numpy as np
import pytest
import torch

from oracle import super_oracle as so
from golden_util import Golden
from gpu_util import frame_from_newdata, tracker_from_state

pytestmark = pytest.mark.gpu

G = Golden()
TRACKED = G.frames[1:]


def _gpu_frame(t):
    from super_b200 import engine
    f = G.frame(t)
    return engine.preprocess(G.opt, torch.from_numpy(f["depth"]).cuda(), torch.from_numpy(f["color"]).cuda(),
                             torch.from_numpy(f["K"]), torch.from_numpy(f["inv_K"]), f["time"])


def _check_producer_bit_exact(fr, nd):
    valid = fr.valid.cpu()
    assert torch.equal(valid, nd.valid), "validity map differs"
    # float32 arithmetic restated instruction for instruction (oracle normals_8 / exp32_def): everything is bit-exact
    assert torch.equal(fr.vmap.cpu()[valid, :3].double(), nd.points), "back-projected points must be bit-exact (f32)"
    assert torch.equal(fr.nmap.cpu()[valid, :3].double(), nd.norms), "normals must be bit-exact (f32)"
    assert torch.equal(fr.confs.cpu()[valid], nd.confs), "confidences must be bit-exact (f32)"
    assert ((fr.radii.cpu()[valid] - nd.radii).abs() <= 4e-16 * nd.radii.abs()).all()     # one f64 divide
    assert float(fr.vmap.cpu()[~valid].abs().max()) == 0.0


@pytest.mark.parametrize("t", G.frames[:2])
def test_producer_matches_oracle(t):
    _check_producer_bit_exact(_gpu_frame(t), G.new_data(t, ref_exp=False))


@pytest.mark.parametrize("data", ["superv1", "superv2"])
def test_producer_bit_exact_fullsize(data):
    """640x480, both datasets' validity rules: the producer's float32 outputs equal the oracle's to the bit."""
    from super_b200 import engine, synth
    H, W = 480, 640
    opt = so.default_opt(height=H, width=W, data=data)
    f = synth.frame_inputs(7, H, W, data=data)
    fr = engine.preprocess(opt, torch.from_numpy(f["depth"]).cuda(), torch.from_numpy(f["color"]).cuda(),
                           torch.from_numpy(f["K"]), torch.from_numpy(f["inv_K"]), f["time"])
    _check_producer_bit_exact(fr, so.preprocess(opt, f))


@pytest.mark.parametrize("t", TRACKED)
def test_fuse_and_compact_match_reference(t):
    """Teacher-forced: reference state after frame t-1, reference beta, reference new_data."""
    from super_b200 import ops
    sf, nd = G.state(t - 1), G.new_data(t)
    trk = tracker_from_state(G.opt, sf)
    fr = frame_from_newdata(nd, G.frame(t), G.H, G.W)
    beta = torch.from_numpy(G[f"f{t}.beta"].copy()).cuda()
    v = trk.view(trk.n_bound)
    ops.warp_update(v.points, v.norms, v.knn_indices, v.knn_w, trk.ED.points, trk.ED.norms, beta, n_dev=trk.cur.n_dev)
    # fuse
    from super_b200 import engine
    import ctypes
    from super_b200.lib import call, ptr, stream
    pr = engine.SbFuseParams(G.opt.th_dist, G.opt.th_cosine_ang, float(t), 0, 0, 0)
    call("sb_fuse", trk.cur.ref(), fr.ref(), ptr(trk.ED.points), ptr(trk.ED.radii), trk.ED.num, ctypes.byref(pr),
         None, 0, ptr(trk.n_tmp), ptr(trk.overflow), ptr(trk.fuse_ws), trk.fuse_ws.numel(), stream())
    n_fused = int(trk.n_tmp.item())
    assert int(trk.overflow.item()) == 0
    assert n_fused == int(G[f"f{t}.fuse.N"]), "number of rows after fusion differs"
    stable = trk.cur.stable[:n_fused].cpu().bool().numpy()
    assert np.array_equal(np.packbits(stable), G[f"f{t}.fuse.isStable"]), "isStable after fusion differs"
    trk.cur.n_dev.copy_(trk.n_tmp)
    trk._compact(fr)
    ref = G.state(t)
    snap = trk.snapshot()
    assert len(snap["points"]) == len(ref.points)
    assert torch.equal(snap["knn_indices"].cpu(), ref.knn_indices)
    for k, tol in (("points", 1e-13), ("norms", 1e-13), ("knn_w", 1e-13), ("radii", 1e-15), ("confs", 1e-6),
                   ("colors", 1e-6), ("time_stamp", 0.0), ("projdata", 1e-4)):
        err = float((snap[k].cpu().double() - getattr(ref, k).double()).abs().max())
        assert err <= tol, f"{k}: max|d| = {err:g}"


def test_free_running_tracker_follows_reference():
    """Whole frames through engine.Tracker (own producer): surfel counts and LM losses per frame."""
    from super_b200 import engine
    trk = engine.Tracker(G.opt)
    for t in G.frames:
        f = G.frame(t)
        beta = trk.step(torch.from_numpy(f["depth"]).cuda(), torch.from_numpy(f["color"]).cuda(),
                        torch.from_numpy(f["K"]), torch.from_numpy(f["inv_K"]), f["time"])
        n_ref = len(G[f"f{t}.state.points"])
        assert trk.num_surfels() == n_ref, f"frame {t}: {trk.num_surfels()} surfels vs reference {n_ref}"
        if beta is not None:
            st = trk.ws.state.read()
            ref_loss = G[f"f{t}.lm.loss"]
            rel = np.abs(st["loss"] - ref_loss) / ref_loss
            assert rel.max() < 1e-4, f"frame {t}: loss rel err {rel.max():.2e}"      # north_star tolerance
            assert np.abs(beta.cpu().numpy() - G[f"f{t}.beta"]).max() < 1e-4
    # ED nodes after the sequence
    t = G.frames[-1]
    assert np.abs(trk.ED.points.cpu().numpy() - G[f"f{t}.state.ED_points"]).max() < 1e-6


def test_hand_over_slow_paths_give_the_same_frames():
    """The per-frame hand-over (engine.Tracker._publish_count / _refresh_bound): a precomputed tuple order that turns
    out too short is redone, and a band of another width is planned and cached -- same losses and beta as the plain run."""
    from super_b200 import engine
    runs = []
    for sabotage in (False, True):
        trk = engine.Tracker(G.opt)
        out = []
        for t in G.frames:
            f = G.frame(t)
            if sabotage and trk._n_event is not None:        # a hand-over is pending (every tracked frame but the first)
                # pretend the order computed after the last compaction covered only a few rows, and that the planned band
                # was another one: the first forces the redo path, the second a (cached) re-plan
                trk._order_rows = 64
                trk._order = trk._order[:64].contiguous()
                trk.band = None
            beta = trk.step(torch.from_numpy(f["depth"]).cuda(), torch.from_numpy(f["color"]).cuda(),
                            torch.from_numpy(f["K"]), torch.from_numpy(f["inv_K"]), f["time"])
            if beta is not None:
                out.append((trk.ws.state.read()["loss"].copy(), beta.cpu().numpy().copy(), trk.num_surfels()))
        if sabotage:
            assert trk._order_redone == len(G.frames) - 2 and trk.band is not None
        runs.append(out)
    for (la, ba, na), (lb, bb, nb) in zip(*runs):
        assert na == nb
        assert np.abs(la - lb).max() <= 1e-9 * np.abs(la).max()      # atomics: summation order differs run to run
        assert np.abs(ba - bb).max() < 1e-9


def test_init_state_matches_reference():
    from super_b200 import engine
    trk = engine.Tracker(G.opt)
    t = G.frames[0]
    f = G.frame(t)
    trk.step(torch.from_numpy(f["depth"]).cuda(), torch.from_numpy(f["color"]).cuda(), torch.from_numpy(f["K"]),
             torch.from_numpy(f["inv_K"]), f["time"])
    ref = G.state(t)
    snap = trk.snapshot()
    assert len(snap["points"]) == len(ref.points)
    assert torch.equal(snap["knn_indices"].cpu(), ref.knn_indices)
    assert torch.equal(snap["points"].cpu(), ref.points)
    assert (snap["knn_w"].cpu() - ref.knn_w).abs().max() < 1e-12
    assert torch.equal(snap["projdata"].cpu(), ref.projdata)
    assert torch.equal(trk.ED.knn_indices.cpu().long(), ref.ED.knn_indices)
    assert (trk.ED.radii.cpu() - ref.ED.radii).abs().max() < 1e-15
    assert (trk.ED.knn_w.cpu() - ref.ED.knn_w).abs().max() < 1e-12
    assert torch.equal(trk.ED.edge_index.cpu(), ref.ED.edge_index)
    assert torch.equal(trk.ED.triangles.cpu(), ref.ED.triangles)


def test_run_super_entry_point_on_disk_sequence(tmp_path):
    """run_super.py (same flags as the reference's) over an on-disk synthetic sequence in the reference's layout
    gives the same state as driving the Tracker directly."""
    import sys, os
    from super_b200 import synth, engine
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "python-super_b200")
    sys.path.insert(0, root)
    import run_super
    H, W = 96, 128
    synth.write_sequence(str(tmp_path), [1, 2, 3], H, W, speed=3.0)
    models = run_super.main(["--model_name", "t", "--data_dir", str(tmp_path), "--start_id", "1", "--end_id", "4",
                             "--height", str(H), "--width", str(W), "--load_depth", "--mesh_step_size", "16",
                             "--sf_point_plane", "--mesh_rot", "--mesh_arap", "--use_derived_gradient"])
    n_entry = models.super.sf.sf_num
    opt = so.default_opt(height=H, width=W, mesh_step_size=16)
    trk = engine.Tracker(opt, device="cuda:0")
    tex = synth.texture(H, W)
    for t in (1, 2, 3):
        fr = synth.frame_inputs(t, H, W, tex=tex, speed=3.0)
        trk.step(torch.from_numpy(fr["depth"]).cuda(), torch.from_numpy(fr["color"]).cuda(),
                 torch.from_numpy(fr["K"]), torch.from_numpy(fr["inv_K"]), fr["time"])
    assert n_entry == trk.num_surfels()
    assert torch.allclose(models.super.sf.points, trk.cur.points[:n_entry], atol=1e-9)


def test_tracked_points_follow_reference():
    """north_star's third criterion: tracked-point reprojections (projdata[track_id], recorded per labelled frame) within
    0.1 px of the reference over the sequence; the tracked surfel ids themselves are integers and must be equal."""
    import json, os
    from golden_util import GOLDEN_DIR
    from super_b200 import engine, synth
    z = np.load(os.path.join(GOLDEN_DIR, "track_128x96.npz"))
    m = json.loads(str(z["meta"]))
    H, W = m["height"], m["width"]
    opt = so.default_opt(height=H, width=W, mesh_step_size=m["step"])
    trk = engine.Tracker(opt, device="cuda:0")
    trk.enable_tracking({f"{t:06d}": z["gt"] for t in m["frames"]})
    tex = synth.texture(H, W)
    for t in m["frames"]:
        fr = synth.frame_inputs(t, H, W, tex=tex, speed=m["speed"])
        trk.step(torch.from_numpy(fr["depth"]).cuda(), torch.from_numpy(fr["color"]).cuda(), torch.from_numpy(fr["K"]),
                 torch.from_numpy(fr["inv_K"]), fr["time"], filename=fr["filename"])
        assert trk.num_surfels() == int(z[f"f{t}.N"])
        assert np.array_equal(trk.track_id.cpu().numpy(), z[f"f{t}.track_id"]), t
        err = np.abs(trk.track_rsts[fr["filename"]].cpu().numpy() - z[f"f{t}.track_rsts"]).max()
        assert err < 0.1, (t, err)
        assert err < 1e-3            # in fact far inside the tolerance
