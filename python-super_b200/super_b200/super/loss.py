"""DataLoss / ARAPLoss / RotLoss / LossTool, call-compatible with /root/reference/super/loss.py:103-505.

    term = DataLoss(); term.prepare(sf, new_data)
    jtj, jtl = term.forward(lambda_, beta, inputs, new_data, grad=True)     # sparse (7J,7J) f64, dense (7J,1) f64
    r2       = term.forward(lambda_, beta, inputs, new_data, grad=False)    # (M,1) squared residuals

as LM_Solver.prepareCostTerm calls them (/root/reference/super/LM.py:54-79).  Each call is one or two launches of the
kernels the device tracker uses (sb_data_term_jtj / sb_data_term_rows / sb_reg_terms / sb_reg_residuals): the COO
Jacobian + torch.sparse.mm of the reference (loss.py:178-205,285-288) is never formed.  `sf` is a Surfels (or anything
with points, knn_indices, knn_w and ED_nodes), `new_data` either an engine.Frame (dense maps) or a reference-style Data
with compact points / norms / valid (+ index_map), `inputs` needs "K" (1,4,4).

The whole-frame fast path does not go through these classes (LM_Solver.LM = sb_lm_frame); they exist so that code
written against the reference's term API keeps working, and they are tested against the oracle
(tests/test_gpu_face.py).
"""
from __future__ import annotations

import torch

from .. import ops
from ..lib import SuperB200Error, call, ptr, stream

F64 = torch.float64


def _maps(new_data, height, width):
    """(vmap, nmap) dense float4 images of a new frame given either form of new_data."""
    if hasattr(new_data, "vmap"):
        return new_data.vmap, new_data.nmap
    return ops.dense_maps(new_data.points, new_data.norms, new_data.valid.reshape(-1), height, width)


def _camera(inputs, new_data):
    if hasattr(new_data, "cam") and new_data.cam is not None:
        return new_data.cam
    color = inputs[("color", 0)]
    return ops.Camera.from_K(torch.as_tensor(inputs["K"]).cpu(), color.shape[-2], color.shape[-1])


def _sym_sparse(A_lower):
    """Lower-triangular dense accumulation -> the symmetric sparse COO matrix prepare_jtj_jtl returns."""
    A = A_lower + torch.tril(A_lower, -1).t()
    return A.to_sparse()


class LossTool:
    """The static helpers other code imports from the reference (loss.py:103-205)."""

    @staticmethod
    def prepare_Jacobian_idx(cost_size, var_idxs, inc_idx):
        """COO index pairs of a block Jacobian (loss.py:178-197): row r of match m touches columns 7 var_idxs[m,k] +
        inc_idx[r]; returns (2, nnz) int64.  Index bookkeeping only (no arithmetic)."""
        var_idxs = var_idxs.long()
        inc_idx = inc_idx.long().reshape(cost_size, -1) if inc_idx.dim() > 1 else inc_idx.long().reshape(1, -1).expand(cost_size, -1)
        M, K = var_idxs.shape
        n_inc = inc_idx.shape[1]
        rows = torch.arange(M * cost_size, device=var_idxs.device).reshape(M, 1, cost_size, 1).expand(M, K, cost_size, n_inc)
        cols = (7 * var_idxs).reshape(M, K, 1, 1) + inc_idx.reshape(1, 1, cost_size, n_inc)
        return torch.stack([rows.reshape(-1), cols.reshape(-1)], dim=0)

    @staticmethod
    def prepare_jtj_jtl(Jacobian, loss):
        """J^T J and -J^T l of a sparse Jacobian (loss.py:200-205) -- for callers that still build one."""
        Jt = torch.transpose(Jacobian, 0, 1)
        return torch.sparse.mm(Jt, Jacobian), -torch.sparse.mm(Jt, loss)


class DataLoss:
    """Projective point-to-plane data term (loss.py:207-290)."""

    def __init__(self):
        return

    def prepare(self, sf, new_data):
        self.n_neighbors = sf.knn_indices.shape[1]
        self.J_size = sf.ED_nodes.param_num
        self.sf_points = sf.points
        self.sf_knn_w = sf.knn_w
        self.sf_knn_indices = sf.knn_indices if sf.knn_indices.dtype == torch.int32 else sf.knn_indices.to(torch.int32)
        self.ed_points = sf.ED_nodes.points
        self.order = None          # kNN-tuple visiting order, computed on the first Jacobian pass

    def forward(self, lambda_, beta, inputs, new_data, grad=False, dldT_only=False):
        if dldT_only:
            raise SuperB200Error("DataLoss.forward(dldT_only=True) has no caller in the reference (loss.py:261-265) and "
                                 "is not provided")
        cam = _camera(inputs, new_data)
        vmap, nmap = _maps(new_data, cam.H, cam.W)
        beta = beta.to(F64).contiguous()
        J = self.ed_points.shape[0]
        pts, idx, w = self.sf_points.contiguous(), self.sf_knn_indices.contiguous(), self.sf_knn_w.contiguous()
        if grad:
            if self.order is None:
                self.order = ops.tuple_order(idx)
            A = torch.zeros((7 * J, 7 * J), dtype=F64, device=pts.device)
            g = torch.zeros((7 * J, 1), dtype=F64, device=pts.device)
            ops.data_term_jtj(pts, idx, w, self.order, self.ed_points, beta, vmap, nmap, cam, float(lambda_), A, g)
            return _sym_sparse(A), g
        matched, _, r, _ = ops.data_term_rows(pts, idx, w, self.ed_points, beta, vmap, nmap, cam, float(lambda_),
                                              want_jrow=False)
        return torch.pow(r[matched], 2).unsqueeze(1)

    @staticmethod
    def autograd_forward(*args, **kwargs):
        raise SuperB200Error("DataLoss.autograd_forward builds an autograd tape over per-surfel temporaries; super_b200 "
                             "evaluates the autograd configuration with fused loss + analytic-gradient kernels behind "
                             "GraphFit.forward (super_b200/super/deform_mesh.py), which is the supported entry point")


class ARAPLoss:
    """As-rigid-as-possible regulariser (loss.py:403-473)."""

    def __init__(self):
        return

    def prepare(self, sfModel, new_data):
        ed = sfModel.ED_nodes
        self.ed_points = ed.points
        self.ED_knn_indices = ed.knn_indices if ed.knn_indices.dtype == torch.int32 else ed.knn_indices.to(torch.int32)
        self.ED_n_neighbors = self.ED_knn_indices.shape[1]
        self.J_size = (ed.num * self.ED_n_neighbors * 3, ed.param_num)

    def forward(self, lambda_, beta, inputs, new_data, grad=False, dldT_only=False):
        beta = beta.to(F64).contiguous()
        J = self.ed_points.shape[0]
        if grad:
            A = torch.zeros((7 * J, 7 * J), dtype=F64, device=beta.device)
            g = torch.zeros((7 * J, 1), dtype=F64, device=beta.device)
            ops.reg_terms(self.ed_points, self.ED_knn_indices, beta, float(lambda_), 0.0, True, False, A, g)
            return _sym_sparse(A), g
        r2 = torch.empty((J * self.ED_n_neighbors * 3, 1), dtype=F64, device=beta.device)
        call("sb_reg_residuals", ptr(self.ed_points), ptr(self.ED_knn_indices), ptr(beta), J, float(lambda_), 0.0, ptr(r2),
             None, stream())
        return r2

    @staticmethod
    def autograd_forward(input, beta):
        raise SuperB200Error("ARAPLoss.autograd_forward: see DataLoss.autograd_forward (GraphFit.forward is the entry point)")


class RotLoss:
    """Unit-quaternion regulariser, float32 like the reference (loss.py:475-505)."""

    def __init__(self):
        return

    def prepare(self, sfModel, new_data):
        ed = sfModel.ED_nodes
        self.ed_points = ed.points
        self.ed_knn = ed.knn_indices if ed.knn_indices.dtype == torch.int32 else ed.knn_indices.to(torch.int32)
        self.J_size = (ed.num, ed.param_num)

    def forward(self, lambda_, beta, inputs, new_data, grad=False):
        beta = beta.to(F64).contiguous()
        J = self.ed_points.shape[0]
        if grad:
            A = torch.zeros((7 * J, 7 * J), dtype=F64, device=beta.device)
            g = torch.zeros((7 * J, 1), dtype=F64, device=beta.device)
            ops.reg_terms(self.ed_points, self.ed_knn, beta, 0.0, float(lambda_), False, True, A, g)
            return _sym_sparse(A).to(torch.float32), g.to(torch.float32)       # the reference's term is float32 (:488-497)
        r2 = torch.empty((J, 1), dtype=torch.float32, device=beta.device)
        call("sb_reg_residuals", ptr(self.ed_points), ptr(self.ed_knn), ptr(beta), J, 0.0, float(lambda_), None, ptr(r2),
             stream())
        return r2

    @staticmethod
    def autograd_forward(beta):
        raise SuperB200Error("RotLoss.autograd_forward: see DataLoss.autograd_forward (GraphFit.forward is the entry point)")
