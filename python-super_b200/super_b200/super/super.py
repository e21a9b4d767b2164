"""SuPer: per-frame orchestration, call-compatible with /root/reference/super/super.py:11-83.

    models.super(models, inputs)      # once per frame, inputs = one DataLoader item (batch 1)

Host->device staging of the frame (super.py:31-34), producer, then init or fusion
(LM -> update -> fuse -> compact), all through super_b200.engine.Tracker, i.e. libsuper_b200.so.
Both optimisers of the reference are available: the derived-gradient LM solver (--use_derived_gradient, super.py:68)
and the autograd-style GraphFit with SGD/Adam (super.py:70-71), the latter for the non-semantic configuration.
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
        self._trk = None
        if opt.use_derived_gradient:
            from .LM import LM_Solver
            self.lm = LM_Solver(opt)
        else:
            from .deform_mesh import GraphFit
            self.graph_fit = GraphFit(opt)

    def forward(self, models, inputs):
        dev = torch.device("cuda", torch.cuda.current_device())
        if self._trk is None:
            self._trk = engine.Tracker(self.opt, device=dev)
            if getattr(self.opt, "tracking_gt_file", None):          # nodes.py:96,115-126, utils/utils.py:383-391
                import os
                import numpy as np
                gt = np.load(os.path.join(os.path.expanduser(self.opt.data_dir), self.opt.tracking_gt_file),
                             allow_pickle=True).tolist()["gt"]
                self._trk.enable_tracking({f"{int(k):06d}": np.asarray(v) for k, v in gt.items()})
        staged = {}
        for key, ipt in inputs.items():                          # super.py:31-34
            if torch.is_tensor(ipt):
                if key == "divterm":
                    staged[key] = float(ipt.reshape(-1)[0])
                elif key in ("K", "inv_K", "time", "ID"):
                    staged[key] = ipt                            # small host-side parameters
                else:
                    staged[key] = ipt.to(dev, non_blocking=True)
            else:
                staged[key] = ipt
        inputs.update(staged)
        depth, color = inputs[("depth", 0)], inputs[("color", 0)]
        time = float(torch.as_tensor(inputs["time"]).reshape(-1)[0])
        seg = inputs.get(("seg", 0))
        inval = engine.extra_invalid_mask(self.opt, depth, seg=seg, mask=inputs.get("valid_mask"))
        frame = engine.preprocess(self.opt, depth, color, inputs["K"], inputs["inv_K"], time,
                                  frame=self._trk.next_frame(), inval=inval,
                                  divterm=inputs.get("divterm", 1.0 / (2.0 * 0.6 * 0.6)),
                                  seg_scores=inputs.get(("seg_conf", 0)))
        self.last_frame = frame
        fname = inputs["filename"][0] if "filename" in inputs else f"{int(time):06d}"
        if self._trk.cur is None:
            self._trk.init(frame)
            self.sf = Surfels(self.opt, self._trk)
            deform_param = None
        else:
            deform_param = self.fusion(models, inputs, frame)
        self._trk._track_points(frame, fname)
        return deform_param

    def fusion(self, models, inputs, sfdata):
        return self._trk.track(sfdata)
