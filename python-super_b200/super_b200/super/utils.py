"""get_skew / Trans_points / transformQuatT, call-compatible with /root/reference/super/utils.py:4-71.

Same names, argument meaning, shapes and return conventions ((value, Jacobian) with Jacobian = 0 when grad is False);
the arithmetic runs in csrc/face.cu with the device functions the fused LM kernels use (common.cuh quat_rot_ref), i.e.
in the reference's operation order.  CUDA float64 tensors only: there is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from ..lib import SuperB200Error, call, ptr, stream

F64 = torch.float64


def _dev64(t, what):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise SuperB200Error(f"{what}: super_b200 helpers take CUDA tensors (no CPU path)")
    return t.to(F64).contiguous()


def get_skew(inputs):
    """(...,3) -> (...,3,3) skew-symmetric matrices [a]x (utils.py:4-14)."""
    a = _dev64(inputs, "get_skew")
    out = torch.empty(a.shape + (3,), dtype=F64, device=a.device)
    call("sb_get_skew", ptr(a), a.numel() // 3, ptr(out), stream())
    return out


def transformQuatT(v, beta, grad=False, skew_v=None):
    """T(q,b) v = v + 2 qw (qv x v) + 2 qv x (qv x v) [+ b when beta has 7 columns]; q is NOT normalised (utils.py:41-71).
    v (J,...,3), beta broadcastable to v's leading shape with last dim 4 or 7.  Returns (tv, d tv / d q (...,3,4)) with
    grad, (tv, 0) without.  skew_v is accepted for call compatibility (the kernel forms [v]x itself)."""
    v = _dev64(v, "transformQuatT")
    beta = _dev64(beta, "transformQuatT")
    bdim = beta.shape[-1]
    beta = beta.expand(v.shape[:-1] + (bdim,)).contiguous()
    n = v.numel() // 3
    tv = torch.empty_like(v)
    jac = torch.empty(v.shape + (4,), dtype=F64, device=v.device) if grad else None
    call("sb_transform_quat", ptr(v), ptr(beta), n, bdim, ptr(tv), ptr(jac), stream())
    return (tv, jac) if grad else (tv, 0)


def Trans_points(d_surfels, ednodes, beta, surfel_knn_weights, grad=False, skew_v=None):
    """Eq. (10) of the SuPer paper: sum_k w_k [T(q_k,b_k)(p - g_k) + g_k] (utils.py:17-38).
    d_surfels, ednodes (N,K,3); beta (N,K,7); surfel_knn_weights (N,K) or a scalar.  Returns (points (N,3), Jacobian
    (N,K,3,4) scaled by the weights) with grad, (points, 0) without."""
    d = _dev64(d_surfels, "Trans_points")
    g = _dev64(ednodes, "Trans_points")
    b = _dev64(beta, "Trans_points").expand(d.shape[:-1] + (7,)).contiguous()
    N, K = d.shape[0], d.shape[1]
    if np.isscalar(surfel_knn_weights):
        w = torch.full((N, K), float(surfel_knn_weights), dtype=F64, device=d.device)
    else:
        w = _dev64(surfel_knn_weights, "Trans_points").expand(N, K).contiguous()
    out = torch.empty((N, 3), dtype=F64, device=d.device)
    jac = torch.empty((N, K, 3, 4), dtype=F64, device=d.device) if grad else None
    call("sb_trans_points", ptr(d), ptr(g), ptr(b), ptr(w), N, K, ptr(out), ptr(jac), stream())
    return (out, jac) if grad else (out, 0)
