"""SuPer: per-frame orchestration, call-compatible with /root/reference/super/super.py:11-83.

    models.super(models, inputs)      # once per frame, inputs = one DataLoader item (batch 1)

Same structure as the reference: stage the inputs on the device (super.py:31-34), depth_preprocessing, then either
init_surfels (mesh_encoder -> Surfels -> prepareStableIndexNSwapAllModel) or fusion (LM_Solver.LM | GraphFit ->
Surfels.update -> fuseInputData -> prepareStableIndexNSwapAllModel -> evaluate every save_sample_freq frames).  Every
stage is a handful of launches of libsuper_b200.so; the host waits once per tracked frame (engine.Tracker).
Both optimisers of the reference are available: the derived-gradient LM solver (--use_derived_gradient, super.py:68)
and the autograd-style GraphFit with SGD/Adam (super.py:70-71).
"""
from __future__ import annotations

import torch

from .. import engine
from .nodes import Surfels


class SuPer(torch.nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.sf = None
        self._frames = None
        self._fi = 0
        if opt.use_derived_gradient:
            from .LM import LM_Solver
            self.lm = LM_Solver(opt)
        else:
            from .deform_mesh import GraphFit
            self.graph_fit = GraphFit(opt)

    def _next_frame(self, dev):
        if self._frames is None:
            self._frames = [engine.Frame(self.opt.height, self.opt.width, dev) for _ in range(2)]
        self._fi ^= 1
        return self._frames[self._fi]

    def forward(self, models, inputs):
        dev = torch.device("cuda", torch.cuda.current_device())
        staged = {}
        for key, ipt in inputs.items():                          # super.py:31-34
            if torch.is_tensor(ipt):
                if key == "divterm":
                    staged[key] = float(ipt.reshape(-1)[0])
                elif key in ("K", "inv_K", "stereo_T", "time", "ID"):
                    staged[key] = ipt                            # small host-side parameters
                else:
                    staged[key] = ipt.to(dev, non_blocking=True)
            else:
                staged[key] = ipt
        inputs.update(staged)
        sfdata, inputs = engine.depth_preprocessing(self.opt, models, inputs, frame=self._next_frame(dev))
        self.last_frame = sfdata
        if self.sf is None:                                      # super.py:47-54
            if getattr(self.opt, "deform_udpate_method", "super_edg") == "super_edg" and hasattr(models, "mesh_encoder"):
                sfdata.ED_nodes = models.mesh_encoder(inputs, sfdata)
            self.init_surfels(models, inputs, sfdata)
            deform_param = None
        else:
            deform_param = self.fusion(models, inputs, sfdata)
        return deform_param

    def init_surfels(self, models, inputs, sfdata):
        self.sf = Surfels(self.opt, models, inputs, sfdata)
        if self.opt.phase == "test":
            self.sf.prepareStableIndexNSwapAllModel(inputs, sfdata)

    def fusion(self, models, inputs, sfdata):
        if self.opt.use_derived_gradient:
            deform_param = self.lm.LM(self.sf, inputs, sfdata)
        else:
            deform_param = self.graph_fit(inputs, self.sf, sfdata, models)
            if deform_param is not None:
                deform_param = deform_param.detach()
        if self.opt.phase == "test":
            with self.sf.tail_scope(deform_param):               # the three stages' launches replayed as one CUDA graph
                self.sf.update(deform_param)
                self.sf.fuseInputData(inputs, sfdata)            # fuse the input data into the reference model
                self.sf.prepareStableIndexNSwapAllModel(inputs, sfdata)
            if int(self.sf.time) % int(getattr(self.opt, "save_sample_freq", 10)) == 0:
                self.sf.evaluate()
        else:
            self.sf.update(deform_param)
        return deform_param
