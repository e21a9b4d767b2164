"""GraphFit: call-compatible with /root/reference/super/deform_mesh.py:10-379.

    deform_verts = GraphFit(opt).forward(inputs, src, trg, models)     # (J+1,7) f64 on device, row J = global transform

`src` is a super_b200 Surfels, `trg` an engine.Frame (the dense-map form of the reference's new_data).  The
optimisation runs as fused loss + analytic-gradient kernels (super_b200.graphfit over csrc/graphfit.cu); the
reference's per-iteration renderer call and empty_cache() (deform_mesh.py:294-298,370) have no counterpart.
Dead branches of the reference are not reproduced: sf_corr (optical flow), render_loss ("not in-use", README:168)
and optim == "LM" (undefined names, :329-368)."""
from __future__ import annotations

import torch


class GraphFit(torch.nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.valid_margin = 1
        self.optim = opt.optimizer
        self.Niter = opt.num_optimize_iterations

    def forward(self, inputs, src, trg, models=None):
        if getattr(self.opt, "deform_udpate_method", "super_edg") != "super_edg":
            raise NotImplementedError("opt['deformation_update_method'] is not specified correctly")
        return self.deform_superedg(inputs, src, trg, models)

    def deform_superedg(self, inputs, src, trg, models=None):
        return src.solve(trg)
