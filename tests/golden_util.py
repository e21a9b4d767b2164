"""Helpers to read tests/golden/*.npz (written by oracle/gen_golden.py from the unmodified reference)
and turn them into oracle-port state.  Test infrastructure."""
import json
import os

import numpy as np
import torch

from oracle import super_oracle as so
from super_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
    def __init__(self, name="lm_128x96"):
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.meta = json.loads(str(self.z["meta"]))
        m = self.meta
        self.H, self.W, self.step, self.frames, self.speed = m["height"], m["width"], m["step"], m["frames"], m["speed"]
        self.semantic, self.seg_speed = bool(m.get("semantic", False)), m.get("seg_speed")
        self.opt = so.default_opt(height=self.H, width=self.W, mesh_step_size=self.step, **m.get("opt", {}))
        self.tex = synth.texture(self.H, self.W)

    def __getitem__(self, k):
        return self.z[k]

    def frame(self, t):
        return synth.frame_inputs(t, self.H, self.W, data=self.opt.data, tex=self.tex, speed=self.speed,
                                  with_seg=self.semantic, seg_speed=self.seg_speed)

    def state(self, t):
        """Oracle-port state namespace holding the reference's state after frame t."""
        pre = f"f{t}.state."
        sf, ed = so.NS(), so.NS()
        for k in self.z.files:
            if not k.startswith(pre):
                continue
            name = k[len(pre):]
            v = self.z[k]
            if v.dtype == np.int32:
                v = v.astype(np.int64)
            ten = torch.from_numpy(v.copy())
            if name.startswith("ED_"):
                setattr(ed, name[3:], ten)
            else:
                setattr(sf, name, ten)
        ed.num = len(ed.points)
        ed.param_num = 7 * ed.num
        sf.ED = ed
        sf.time = t
        sf.semantic = self.semantic
        sf.track_id = None
        return sf

    def new_data(self, t, ref_exp=True):
        """new_data of frame t: recomputed by the port's producer (checked against the golden copies
        in test_oracle_golden.py).  ref_exp: with the reference's float32 exp (torch.exp), so that everything downstream
        compares against the reference's golden vectors at rounding level; False = the oracle's defined exp32_def."""
        return so.preprocess(self.opt, self.frame(t), ref_exp=ref_exp)
