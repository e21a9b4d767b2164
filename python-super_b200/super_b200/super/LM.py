"""LM_Solver: call-compatible with /root/reference/super/LM.py:10-122.

    beta = LM_Solver(opt).LM(sf, inputs, new_data, u=10, v=7.5, minimal_loss=1e10)   # (J,7) f64 on device

`sf` is a super_b200 Surfels, `new_data` an engine.Frame (the dense-map form of the reference's
new_data).  A failed factorisation never raises: the loop stops and the last beta is returned
(LM.py:99-103), decided on the device."""
from __future__ import annotations

from .. import lm as _lm


class LM_Solver:
    def __init__(self, opt, convs=None):
        self.opt = opt
        self.phase = opt.phase
        self.ws = None

    def LM(self, sf, inputs, new_data, u=10, v=7.5, minimal_loss=1e10):
        trk = sf._trk
        view = trk.view(trk.n_bound)
        beta, self.ws = _lm.lm_solve(view, (new_data.vmap, new_data.nmap), new_data.cam, self.opt, ws=self.ws,
                                     u=u, v=v, minimal_loss=minimal_loss, n_dev=trk.cur.n_dev)
        return beta
