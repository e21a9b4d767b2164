"""GraphFit: call-compatible with /root/reference/super/deform_mesh.py:10-379.

    deform_verts = GraphFit(opt).forward(inputs, src, trg, models)     # (J+1,7) f64 on device, row J = global transform

`src` is a super_b200 Surfels, `trg` an engine.Frame (the dense-map form of the reference's new_data).  The
optimisation runs as fused loss + analytic-gradient kernels (super_b200.graphfit over csrc/graphfit.cu); the
reference's per-iteration renderer call and empty_cache() (deform_mesh.py:294-298,370) have no counterpart.
Dead branches of the reference are not reproduced: sf_corr (optical flow), render_loss ("not in-use", README:168)
and optim == "LM" (undefined names, :329-368)."""
from __future__ import annotations

import torch

from .. import graphfit as _gf


class GraphFit(torch.nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.valid_margin = 1
        self.optim = opt.optimizer
        self.Niter = opt.num_optimize_iterations
        self.ws = None

    def forward(self, inputs, src, trg, models=None):
        if getattr(self.opt, "deform_udpate_method", "super_edg") != "super_edg":
            raise NotImplementedError("opt['deformation_update_method'] is not specified correctly")
        return self.deform_superedg(inputs, src, trg, models)

    def deform_superedg(self, inputs, src, trg, models=None):
        from types import SimpleNamespace as NS
        trk = src._trk
        view = trk.view(trk.n_bound)
        view.isStable = trk.cur.stable[: trk.n_bound]
        view.ED = NS(points=trk.ED.points, knn_indices=trk.ED.knn_indices, knn_w=trk.ED.knn_w,
                     triangles=trk.ED.triangles_i32, triangles_areas=trk.ED.triangles_areas)
        dv, self.ws = _gf.graph_fit(view, (trg.vmap, trg.nmap), trg.cam, self.opt, ws=self.ws, n_dev=trk.cur.n_dev)
        return dv
