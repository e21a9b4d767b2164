"""LM_Solver: call-compatible with /root/reference/super/LM.py:10-122.

    beta = LM_Solver(opt).LM(sf, inputs, new_data, u=10, v=7.5, minimal_loss=1e10)   # (J,7) f64 on device

`sf` is a super_b200 Surfels, `new_data` an engine.Frame (the dense-map form of the reference's new_data).  The loop
runs as ONE C call (sb_lm_frame, csrc/lm_frame.cu) on the band path; a failed factorisation never raises: the updates
stop and the last accepted beta is returned (LM.py:99-103), decided on the device.  prepareCostTerm / Solver and the
term objects (`losses`, `lambdas`) are kept for code that drives the solver piecewise like the reference does."""
from __future__ import annotations

import torch

from .loss import ARAPLoss, DataLoss, RotLoss


class LM_Solver:
    def __init__(self, opt, convs=None):
        self.opt = opt
        self.device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        self.losses, self.lambdas = [], []
        if opt.sf_point_plane:
            self.losses.append(DataLoss())
            self.lambdas.append(opt.sf_point_plane_weight)
        if opt.mesh_arap:
            self.losses.append(ARAPLoss())
            self.lambdas.append(opt.mesh_arap_weight)
        if opt.mesh_rot:
            self.losses.append(RotLoss())
            self.lambdas.append(opt.mesh_rot_weight)
        self.phase = opt.phase
        if self.phase == "train":
            self.convs = convs

    @staticmethod
    def Solver(A, b, method="cholesky"):
        """Dense solve of a damped normal-equation system (LM.py:38-51): library call, kept for API parity -- LM() uses
        the banded solver of the CUDA library."""
        if method == "lu":
            LU, piv = torch.linalg.lu_factor(A)
            return torch.linalg.lu_solve(LU, piv, b)
        return torch.cholesky_solve(b, torch.linalg.cholesky(A))

    def prepareCostTerm(self, sf, inputs, new_data, beta, grad=False):
        """Sum of the terms' normal equations (grad) or losses (LM.py:54-79).  Call `term.prepare(sf, new_data)` on
        self.losses first, as LM does."""
        if grad:
            n = sf.ED_nodes.param_num
            jtj = torch.zeros((n, n), dtype=torch.float64, device=beta.device)
            jtl = torch.zeros((n, 1), dtype=torch.float64, device=beta.device)
            for term, lam in zip(self.losses, self.lambdas):
                a, g = term.forward(lam, beta, inputs, new_data, grad=True)
                jtj += a.to_dense().to(torch.float64)
                jtl += g.to(torch.float64)
            return jtj, jtl
        loss = [term.forward(lam, beta, inputs, new_data) for term, lam in zip(self.losses, self.lambdas)]
        return torch.sum(torch.cat([x.to(torch.float64) for x in loss]))

    def LM(self, sf, inputs, new_data, u=10, v=7.5, minimal_loss=1e10):
        return sf.solve(new_data, u=float(u), v=float(v), minimal_loss=float(minimal_loss))
