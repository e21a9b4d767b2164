"""CPU: the oracle port (oracle/super_oracle.py) against the golden vectors produced by the
unmodified reference (oracle/gen_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import super_oracle as so
from golden_util import Golden

G = Golden()
TRACKED = G.frames[1:]


def _close(a, b, tol, what):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype.kind in "iub":
        assert np.array_equal(a, b), f"{what}: {np.count_nonzero(a != b)} integer mismatches"
    else:
        err = np.abs(a.astype(np.float64) - b.astype(np.float64)).max() if a.size else 0.0
        assert err <= tol, f"{what}: max|d| = {err:g} > {tol:g}"


@pytest.mark.parametrize("t", G.frames)
def test_producer_matches_reference(t):
    nd = G.new_data(t)
    _close(nd.points, G[f"f{t}.nd.points"].astype(np.float64), 0.0, "points")
    # normals / confidences: bit-exact with the reference's own float32 exp (torch.exp = MKL vsExp on the CPU); the oracle's
    # default DEFINES that exponential (exp32_def, what the CUDA path is held to) and agrees to float32 rounding.  Every
    # other step (summation orders, FMA placement, sqrt, divide) is spelt out in the port and equals the reference's.
    _close(nd.norms, G[f"f{t}.nd.norms"].astype(np.float64), 0.0, "norms")
    nd_def = G.new_data(t, ref_exp=False)
    assert torch.equal(nd_def.valid, nd.valid) and torch.equal(nd_def.points, nd.points)
    _close(nd_def.norms, G[f"f{t}.nd.norms"].astype(np.float64), 2.5e-7, "norms (exp32_def)")
    _close(nd_def.confs, G[f"f{t}.nd.confs"], 1e-7, "confs (exp32_def)")
    _close(nd_def.radii, G[f"f{t}.nd.radii"], 1e-9, "radii (exp32_def)")
    assert np.array_equal(np.packbits(nd.valid.numpy()), G[f"f{t}.nd.valid"])
    _close(nd.radii, G[f"f{t}.nd.radii"], 1e-15, "radii")
    _close(nd.confs, G[f"f{t}.nd.confs"], 0.0, "confs")


def test_init_state_matches_reference():
    t = G.frames[0]
    nd = G.new_data(t)
    sf = so.init_surfels(G.opt, nd, so.build_graph(G.opt, nd))
    ref = G.state(t)
    for k in ("knn_indices", "isStable"):
        _close(getattr(sf, k), getattr(ref, k).numpy(), 0, k)
    for k in ("points", "norms", "knn_w", "radii", "confs", "colors", "time_stamp", "projdata"):
        _close(getattr(sf, k), getattr(ref, k).numpy(), 1e-12, k)
    for k in ("knn_indices", "edge_index", "triangles"):
        _close(getattr(sf.ED, k), getattr(ref.ED, k).numpy(), 0, "ED." + k)
    for k in ("points", "radii", "knn_w", "triangles_areas"):
        _close(getattr(sf.ED, k), getattr(ref.ED, k).numpy(), 1e-12, "ED." + k)


@pytest.mark.parametrize("t", TRACKED)
def test_lm_trace_matches_reference(t):
    sf, nd = G.state(t - 1), G.new_data(t)
    trace = []
    beta = so.lm_solve(G.opt, sf, nd, assemble="blocks", trace=trace)
    assert len(trace) == len(G[f"f{t}.lm.loss"])
    for i, it in enumerate(trace):
        ref_loss = G[f"f{t}.lm.loss"][i]
        assert abs(it["loss"] - ref_loss) <= 1e-9 * ref_loss, f"it{i} loss {it['loss']} vs {ref_loss}"
        _close(it["beta_try"], G[f"f{t}.lm.beta_try"][i], 1e-10, f"it{i} beta_try")
        _close(it["delta"], G[f"f{t}.lm.delta"][i], 1e-10, f"it{i} delta")
        assert abs(it["u"] - G[f"f{t}.lm.u"][i]) <= 1e-9 * it["u"] + 1e-12   # golden u = A_damped[0,0]-A[0,0]
        assert len(it["ids"]) == G[f"f{t}.lm.M"][i]
        if f"f{t}.lm.it{i}.ids" in G.z.files:     # integer correspondences: bit-exact
            _close(it["ids"].to(torch.int32), G[f"f{t}.lm.it{i}.ids"], 0, f"it{i} ids")
            _close(it["corners"].to(torch.int16), G[f"f{t}.lm.it{i}.corners"], 0, f"it{i} corners")
        _close(torch.diagonal(it["A"]), G[f"f{t}.lm.A_diag"][i], 1e-9 * G[f"f{t}.lm.A_diag"][i].max(), "A diag")
        _close(it["g"][:, 0], G[f"f{t}.lm.g"][i], 1e-11, "g")
        if f"f{t}.lm.it{i}.A" in G.z.files:
            A = G[f"f{t}.lm.it{i}.A"]
            _close(it["A"], A, 1e-11 * np.abs(A).max(), "A")
    _close(beta, G[f"f{t}.beta"], 1e-10, "beta")


def test_lm_sparse_mm_assembly_equals_block_assembly():
    t = TRACKED[0]
    sf, nd = G.state(t - 1), G.new_data(t)
    beta = torch.tensor([[1., 0, 0, 0, 0, 0, 0]], dtype=torch.float64).repeat(sf.ED.num, 1)
    A1, g1, _ = so.lm_normal_equations(G.opt, sf, nd, beta, "sparse_mm")
    A2, g2, _ = so.lm_normal_equations(G.opt, sf, nd, beta, "blocks")
    assert (A1 - A2).abs().max() <= 1e-12 * A1.abs().max()
    assert (g1 - g2).abs().max() <= 1e-13


@pytest.mark.parametrize("t", TRACKED)
def test_update_fuse_compact_match_reference(t):
    sf, nd = G.state(t - 1), G.new_data(t)
    beta = torch.from_numpy(G[f"f{t}.beta"].copy())
    so.update(G.opt, sf, beta)
    _close(sf.points, G[f"f{t}.update.points"], 1e-13, "update.points")
    _close(sf.norms, G[f"f{t}.update.norms"], 1e-13, "update.norms")
    _close(sf.ED.points, G[f"f{t}.update.ED_points"], 1e-13, "update.ED.points")
    _close(sf.ED.norms, G[f"f{t}.update.ED_norms"], 1e-13, "update.ED.norms")
    so.fuse(G.opt, sf, nd)
    assert len(sf.isStable) == int(G[f"f{t}.fuse.N"])
    assert np.array_equal(np.packbits(sf.isStable.numpy()), G[f"f{t}.fuse.isStable"])
    so.compact(G.opt, sf, float(t))
    ref = G.state(t)
    assert len(sf.points) == len(ref.points)
    _close(sf.knn_indices, ref.knn_indices.numpy(), 0, "knn_indices")
    for k in ("points", "norms", "knn_w", "radii", "confs", "colors", "time_stamp", "projdata"):
        tol = 1e-5 if k == "projdata" else 1e-12
        _close(getattr(sf, k), getattr(ref, k).numpy(), tol, k)


# ---- autograd optimiser (GraphFit) fixtures: configs 3b (Adam, LM terms + face) and 4 (semantic, SGD) ----------
@pytest.mark.parametrize("name", ["gf_128x96", "gf_sem_128x96", "gf_hard_128x96"])
def test_graphfit_port_reproduces_reference(name):
    """oracle/graphfit_oracle.py, teacher-forced with the reference's pre-frame state, reproduces the reference's
    per-iteration deform_verts, loss terms, consumed gradient, result and Surfels.update output."""
    import torch
    from oracle import graphfit_oracle as gf
    g = Golden(name)
    for t in g.frames[1:]:
        sf, nd = g.state(t - 1), g.new_data(t)
        tr = []
        dv = gf.graph_fit(g.opt, sf, nd, trace=tr)
        assert np.abs(np.stack([x["deform_in"].numpy() for x in tr]) - g[f"f{t}.ag.deform_in"]).max() < 1e-13
        gref = g[f"f{t}.ag.grad"]
        assert np.abs(np.stack([x["grad"].numpy() for x in tr]) - gref).max() < 1e-10 * np.abs(gref).max()
        for k in [k for k in g.z.files if k.startswith(f"f{t}.ag.losses.")]:
            ref = g[k]
            mine = np.array([x["losses"].get(k.split(".")[-1], np.nan) for x in tr])
            assert np.allclose(mine, ref, rtol=1e-11, atol=0, equal_nan=True), k
        assert np.abs(dv.numpy() - g[f"f{t}.beta"]).max() < 1e-13
        so.update(g.opt, sf, dv)
        for k, v in (("points", sf.points), ("norms", sf.norms), ("ED_points", sf.ED.points), ("ED_norms", sf.ED.norms)):
            assert np.abs(v.numpy() - g[f"f{t}.update.{k}"]).max() < 1e-12, k


def test_tracked_points_port_reproduces_reference():
    """Tracked-point ids and recorded reprojections of the reference (run with --tracking_gt_file) over 4 frames."""
    import json
    import os
    from golden_util import GOLDEN_DIR
    from super_b200 import synth
    z = np.load(os.path.join(GOLDEN_DIR, "track_128x96.npz"))
    m = json.loads(str(z["meta"]))
    H, W = m["height"], m["width"]
    opt = so.default_opt(height=H, width=W, mesh_step_size=m["step"])
    gt = {f"{t:06d}": z["gt"] for t in m["frames"]}
    trk = so.Tracker(opt, gt=gt)
    tex = synth.texture(H, W)
    for t in m["frames"]:
        fr = synth.frame_inputs(t, H, W, tex=tex, speed=m["speed"])
        trk.step(fr)
        assert len(trk.sf.points) == int(z[f"f{t}.N"])
        assert np.array_equal(trk.sf.track_id.numpy(), z[f"f{t}.track_id"])
        assert np.abs(trk.track_rsts[fr["filename"]].numpy() - z[f"f{t}.track_rsts"]).max() < 1e-5


def test_hard_seg_state_port_reproduces_reference():
    """--hard_seg: per-class kNN at init and for appended surfels, class-gated merges, class-pruned graph: the port,
    free running on the fixture's inputs, reproduces the reference's state after every frame."""
    from oracle import graphfit_oracle as gf
    g = Golden("gf_hard_128x96")
    sf = None
    for t in g.frames:
        nd = g.new_data(t)
        if sf is None:
            sf = so.init_surfels(g.opt, nd, so.build_graph(g.opt, nd))
        else:
            dv = gf.graph_fit(g.opt, sf, nd)
            so.update(g.opt, sf, dv); so.fuse(g.opt, sf, nd); so.compact(g.opt, sf, float(t))
        ref = g.state(t)
        assert len(sf.points) == len(ref.points)
        assert torch.equal(sf.knn_indices, ref.knn_indices) and torch.equal(sf.seg, ref.seg)
        assert torch.equal(sf.ED.knn_indices, ref.ED.knn_indices) and torch.equal(sf.ED.triangles, ref.ED.triangles)
        assert (sf.points - ref.points).abs().max() < 1e-12 and (sf.knn_w - ref.knn_w).abs().max() < 1e-12
