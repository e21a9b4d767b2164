"""GPU: the fused GraphFit kernels (csrc/graphfit.cu, through the C ABI) against the golden vectors of the
unmodified reference's autograd optimiser, teacher-forced with the reference's pre-frame state."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from oracle import super_oracle as so
from golden_util import Golden
from gpu_util import to_device_state, device_maps, camera

pytestmark = pytest.mark.gpu


def _device_inputs(g, t):
    from super_b200 import graphfit
    sf, nd = g.state(t - 1), g.new_data(t)
    d = to_device_state(sf)
    d.isStable = sf.isStable.to(torch.uint8).cuda()
    d.ED.triangles = sf.ED.triangles.to(torch.int32).cuda().contiguous()
    d.ED.triangles_areas = sf.ED.triangles_areas.cuda().contiguous()
    maps = device_maps(nd, g.H, g.W)
    cam = camera(nd, g.H, g.W)
    seg = None
    if g.semantic:
        P = g.H * g.W
        trg = torch.zeros((P, g.opt.num_classes), dtype=torch.float64)
        trg[nd.valid] = nd.seg_conf
        seg = NS(sf_seg=sf.seg.to(torch.int32).cuda(), sf_seg_conf=sf.seg_conf.cuda().contiguous(),
                 trg_seg_conf=trg.cuda(), scores=nd.seg_conf_in[0].cuda().contiguous())
        if getattr(g.opt, "sf_bn_morph", False):
            seg.edge_pts, seg.edge_off = graphfit.edge_points(nd.seg_in[0, 0].cuda(), g.opt.num_classes, g.H, g.W)
    return sf, nd, d, maps, cam, seg


@pytest.mark.parametrize("name", ["gf_128x96", "gf_sem_128x96", "gf_hard_128x96"])
def test_graphfit_matches_reference(name):
    """Per-iteration deform_verts, consumed gradient, loss terms and the result against the reference's autograd."""
    from super_b200 import graphfit
    g = Golden(name)
    for t in g.frames[1:]:
        sf, nd, d, maps, cam, seg = _device_inputs(g, t)
        dv, ws = graphfit.graph_fit(d, maps, cam, g.opt, seg=seg, log_grad=True)
        iters = g.opt.num_optimize_iterations
        gref = g[f"f{t}.ag.grad"]
        gmine = ws.grad_log.cpu().numpy()
        # first iteration: same deform_verts (identity) on both sides -> pure kernel-vs-autograd gradient check
        assert np.abs(gmine[0] - gref[0]).max() < 1e-9 * np.abs(gref[0]).max(), "gradient at identity"
        assert np.abs(ws.dv_log.cpu().numpy() - g[f"f{t}.ag.deform_in"]).max() < 1e-10
        assert np.abs(gmine - gref).max() < 1e-6 * np.abs(gref).max()
        tr = ws.read_trace(iters)
        for k in [k for k in g.z.files if k.startswith(f"f{t}.ag.losses.")]:
            ref = g[k]
            mine = np.array([x[k.split(".")[-1]] for x in tr])
            assert np.allclose(mine, ref, rtol=1e-8, atol=1e-18, equal_nan=True), (k, mine, ref)
        assert np.allclose(np.array([x["total"] for x in tr]), g[f"f{t}.ag.loss"], rtol=1e-8, equal_nan=True)
        # north_star tolerance on the result is 1e-4; the kernels are ~1e-12 here
        assert np.abs(dv.cpu().numpy() - g[f"f{t}.beta"]).max() < 1e-10


@pytest.mark.parametrize("name", ["gf_128x96", "gf_sem_128x96"])
def test_update_with_global_row_matches_reference(name):
    from super_b200 import graphfit, ops
    g = Golden(name)
    for t in g.frames[1:]:
        sf = g.state(t - 1)
        d = to_device_state(sf)
        dv = torch.from_numpy(g[f"f{t}.beta"].copy()).cuda()
        ops.warp_update(d.points, d.norms, d.knn_indices, d.knn_w, d.ED.points, d.ED.norms, dv[:-1].contiguous())
        graphfit.update_global(d.points, d.norms, d.ED.points, d.ED.norms, dv)
        for k, v in (("points", d.points), ("norms", d.norms), ("ED_points", d.ED.points), ("ED_norms", d.ED.norms)):
            assert np.abs(v.cpu().numpy() - g[f"f{t}.update.{k}"]).max() < 1e-12, k


def test_edge_points_match_oracle():
    from oracle import graphfit_oracle as gfo
    from super_b200 import graphfit
    g = Golden("gf_sem_128x96")
    nd = g.new_data(g.frames[1])
    ref = gfo.edge_points(g.opt, nd)
    pts, off = graphfit.edge_points(nd.seg_in[0, 0].cuda(), g.opt.num_classes, g.H, g.W)
    off = off.cpu().tolist()
    for c in range(g.opt.num_classes):
        assert torch.equal(pts[off[c]:off[c + 1]].cpu(), ref[c])


def test_free_running_autograd_tracker_follows_oracle():
    """Config 3b shape (Adam, point-plane + ARAP + rot + face), device tracker vs the CPU port, free running:
    per-frame final loss within 1e-4 relative (north_star), deform_verts within 1e-6, equal surfel counts."""
    from oracle import graphfit_oracle as gfo
    from super_b200 import engine, synth
    H, W, step = 96, 128, 16
    opt = so.default_opt(height=H, width=W, mesh_step_size=step, use_derived_gradient=False, mesh_face=True,
                         optimizer="Adam")
    tex = synth.texture(H, W)
    trk = engine.Tracker(opt, device="cuda:0")
    sf = None
    for t in (1, 2, 3, 4):
        fr = synth.frame_inputs(t, H, W, tex=tex, speed=3.0)
        nd = so.preprocess(opt, fr)
        beta = trk.step(torch.from_numpy(fr["depth"]).cuda(), torch.from_numpy(fr["color"]).cuda(),
                        torch.from_numpy(fr["K"]), torch.from_numpy(fr["inv_K"]), fr["time"])
        if sf is None:
            sf = so.init_surfels(opt, nd, so.build_graph(opt, nd))
            continue
        tr = []
        dv = gfo.graph_fit(opt, sf, nd, trace=tr)
        so.update(opt, sf, dv); so.fuse(opt, sf, nd); so.compact(opt, sf, float(fr["time"]))
        mine = trk.gf_ws.read_trace(opt.num_optimize_iterations)
        rel = abs(mine[-1]["total"] - tr[-1]["loss"]) / tr[-1]["loss"]
        assert rel < 1e-4, (t, mine[-1]["total"], tr[-1]["loss"])
        assert (beta.cpu() - dv).abs().max() < 1e-6
        assert trk.num_surfels() == len(sf.points)


def test_semantic_fuse_and_compact_match_reference():
    """Semantic-SuPer state through fusion and compaction, teacher-forced with the reference's pre-frame state and
    deform_verts (gf_sem fixture): class probabilities are merged with the confidences, classes follow the argmax,
    the kNN weights are the JSD-modulated ones (nodes.py:183-189,345-353,466-484,505-511)."""
    import ctypes
    from super_b200 import graphfit, ops
    from super_b200.lib import call, ptr, stream
    from gpu_util import frame_from_newdata, tracker_from_state
    g = Golden("gf_sem_128x96")
    for t in g.frames[1:]:
        sf, nd = g.state(t - 1), g.new_data(t)
        trk = tracker_from_state(g.opt, sf)
        fr = frame_from_newdata(nd, g.frame(t), g.H, g.W)
        dv = torch.from_numpy(g[f"f{t}.beta"].copy()).cuda()
        v = trk.view(trk.n_bound)
        ops.warp_update(v.points, v.norms, v.knn_indices, v.knn_w, trk.ED.points, trk.ED.norms, dv[:-1].contiguous(),
                        n_dev=trk.cur.n_dev)
        graphfit.update_global(v.points, v.norms, trk.ED.points, trk.ED.norms, dv, n_dev=trk.cur.n_dev)
        pr = trk.fuse_params(fr)
        assert pr.semantic_weights == 3 and pr.class_gate == 0          # superv2, no --hard_seg
        call("sb_fuse", trk.cur.ref(), fr.ref(), ptr(trk.ED.points), ptr(trk.ED.radii), trk.ED.num, ctypes.byref(pr),
             None, 0, ptr(trk.n_tmp), ptr(trk.overflow), ptr(trk.fuse_ws), trk.fuse_ws.numel(), stream())
        trk.cur.n_dev.copy_(trk.n_tmp)
        trk._compact(fr)
        ref = g.state(t)
        snap = trk.snapshot()
        assert len(snap["points"]) == len(ref.points)
        assert torch.equal(snap["knn_indices"].cpu(), ref.knn_indices)
        assert torch.equal(snap["seg"].cpu(), ref.seg)
        for k, tol in (("points", 1e-13), ("norms", 1e-13), ("knn_w", 1e-12), ("seg_conf", 1e-13), ("confs", 1e-6)):
            err = float((snap[k].cpu().double() - getattr(ref, k).double()).abs().max())
            assert err <= tol, f"frame {t} {k}: max|d| = {err:g}"


def test_free_running_semantic_tracker_follows_oracle():
    """Config 4 shape end to end on the device (superv2, SGD, soft-seg point-plane + rot + face + boundary-morph,
    semantic kNN weights and fusion) against the CPU port, free running."""
    from oracle import graphfit_oracle as gfo
    from super_b200 import engine, synth
    g = Golden("gf_sem_128x96")
    opt, H, W = g.opt, g.H, g.W
    trk = engine.Tracker(opt, device="cuda:0")
    sf = None
    for t in (1, 2, 3):
        fr = g.frame(t)
        nd = so.preprocess(opt, fr)
        beta = trk.step(torch.from_numpy(fr["depth"]).cuda(), torch.from_numpy(fr["color"]).cuda(),
                        torch.from_numpy(fr["K"]), torch.from_numpy(fr["inv_K"]), fr["time"],
                        seg_scores=torch.from_numpy(fr["seg_conf"]).cuda())
        if sf is None:
            sf = so.init_surfels(opt, nd, so.build_graph(opt, nd))
            assert trk.num_surfels() == len(sf.points)
            assert (trk.cur.knn_w[: len(sf.points)].cpu() - sf.knn_w).abs().max() < 1e-12
            continue
        tr = []
        dv = gfo.graph_fit(opt, sf, nd, trace=tr)
        so.update(opt, sf, dv); so.fuse(opt, sf, nd); so.compact(opt, sf, float(fr["time"]))
        mine = trk.gf_ws.read_trace(opt.num_optimize_iterations)
        rel = abs(mine[-1]["total"] - tr[-1]["loss"]) / tr[-1]["loss"]
        assert rel < 1e-4, (t, mine[-1], tr[-1])
        assert (beta.cpu() - dv).abs().max() < 1e-6
        assert trk.num_surfels() == len(sf.points)


def test_hard_seg_tracker_follows_reference():
    """--hard_seg end to end on the device: per-class kNN (graph, init, appended surfels), class-gated merges,
    class-pruned graph, hard-seg point-plane weights -- state after every frame vs the reference's (teacher-forced by
    construction: the device tracker free-runs from the same inputs and the integers must stay equal)."""
    g = Golden("gf_hard_128x96")
    from super_b200 import engine
    trk = engine.Tracker(g.opt, device="cuda:0")
    for t in g.frames:
        fr = g.frame(t)
        beta = trk.step(torch.from_numpy(fr["depth"]).cuda(), torch.from_numpy(fr["color"]).cuda(),
                        torch.from_numpy(fr["K"]), torch.from_numpy(fr["inv_K"]), fr["time"],
                        seg_scores=torch.from_numpy(fr["seg_conf"]).cuda())
        ref = g.state(t)
        snap = trk.snapshot()
        assert len(snap["points"]) == len(ref.points), t
        assert torch.equal(snap["knn_indices"].cpu(), ref.knn_indices), t
        assert torch.equal(snap["seg"].cpu(), ref.seg), t
        assert (snap["knn_w"].cpu() - ref.knn_w).abs().max() < 1e-6
        if t == g.frames[0]:
            assert torch.equal(trk.ED.knn_indices.cpu().long(), ref.ED.knn_indices)
            assert torch.equal(trk.ED.triangles.cpu(), ref.ED.triangles)
        else:
            assert np.abs(beta.cpu().numpy() - g[f"f{t}.beta"]).max() < 1e-6
