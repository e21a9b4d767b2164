"""pcd2depth / find_knn / KLD / JSD / torch_dilate, call-compatible with /root/reference/utils/utils.py:152-254.

CUDA tensors in, CUDA tensors out, same shapes / dtypes / tuple orders as the reference; every number comes from
libsuper_b200.so (csrc/face.cu, csrc/knn_warp.cu).  pytorch3d, which the reference calls for kNN, is not needed.
"""
from __future__ import annotations

import torch

from .. import ops
from ..lib import SuperB200Error, call, intr_array, ptr, stream

F64, I64 = torch.float64, torch.int64


def _dev(t, what):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise SuperB200Error(f"{what}: super_b200 helpers take CUDA tensors (no CPU path)")
    return t


def pcd2depth(inputs, pcd, round_coords=True, valid_margin=0):
    """Point cloud -> image plane (utils.py:161-184).  inputs: dict with ("color",0) (1,3,H,W) for the image size and "K"
    (1,4,4).  Returns (v, u, coords, valid_proj): v, u rounded int64 (round_coords) or float64; coords = round(v) W +
    round(u); valid_proj on the rounded coordinates with the margin."""
    pcd = _dev(pcd, "pcd2depth").to(F64).contiguous()
    height, width = inputs[("color", 0)].shape[-2:]
    K = torch.as_tensor(inputs["K"]).reshape(-1, 4, 4)[0].cpu()
    intr = intr_array(float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]))
    shape = pcd.shape[:-1]
    n = pcd.numel() // 3
    dev = pcd.device
    coords = torch.empty(shape, dtype=I64, device=dev)
    valid = torch.empty(shape, dtype=torch.uint8, device=dev)
    if round_coords:
        v, u = torch.empty(shape, dtype=I64, device=dev), torch.empty(shape, dtype=I64, device=dev)
        call("sb_pcd2depth", ptr(pcd), n, intr, int(height), int(width), int(valid_margin), None, None, ptr(v), ptr(u),
             ptr(coords), ptr(valid), stream())
    else:
        v, u = torch.empty(shape, dtype=F64, device=dev), torch.empty(shape, dtype=F64, device=dev)
        call("sb_pcd2depth", ptr(pcd), n, intr, int(height), int(width), int(valid_margin), ptr(v), ptr(u), None, None,
             ptr(coords), ptr(valid), stream())
    return v, u, coords, valid.bool()


def find_knn(points1, points2, num_classes=-1, seg1=None, seg2=None, k=20, radius=0.5, method="knn"):
    """k nearest points2 for every points1 (utils.py:212-242): (sqrt distances (N,k) f64 ascending, indices (N,k) i64).
    num_classes > 0: neighbours are searched inside the query's own class (seg1 / seg2), rows of an absent class keep
    1e8 / -1.  Definition (pytorch3d absent): exact f64 squared distance, ties -> lower index."""
    if method != "knn":
        raise NotImplementedError("find_knn(method='ball_query') is never selected by the reference's tracking path")
    p1, p2 = _dev(points1, "find_knn").to(F64), _dev(points2, "find_knn").to(F64)
    q = r = None
    if num_classes > 0:
        q, r = seg1.to(torch.int32).contiguous(), seg2.to(torch.int32).contiguous()
    d, i = ops.knn(p1, p2, int(k), qseg=q, rseg=r)
    return d, i.to(I64)


def _kl(P, Q, eps, dim, jsd):
    P, Q = torch.broadcast_tensors(_dev(P, "KLD").to(F64), _dev(Q, "KLD").to(F64))
    if dim not in (-1, P.dim() - 1):
        P, Q = P.movedim(dim, -1), Q.movedim(dim, -1)
    P, Q = P.contiguous(), Q.contiguous()
    C = P.shape[-1]
    out = torch.empty(P.shape[:-1], dtype=F64, device=P.device)
    call("sb_kld_jsd", ptr(P), ptr(Q), P.numel() // C, C, float(eps), int(jsd), ptr(out), stream())
    return out


def KLD(P, Q, eps=1e-13, dim=-1):
    """KL(P || Q) = sum P log(P / (Q + eps) + eps) (utils.py:244-250)."""
    return _kl(P, Q, eps, dim, False)


def JSD(P, Q, eps=1e-13, dim=-1):
    """0.5 (KL(P || M) + KL(Q || M)), M = (P + Q) / 2 (utils.py:252-254)."""
    return _kl(P, Q, eps, dim, True)


def torch_dilate(inputs, kernel=10, dtype=torch.bool):
    """(B,C,H,W) box-filter dilation, conv2d(padding='same') geometry (utils.py:152-157)."""
    x = _dev(inputs, "torch_dilate")
    B, C, H, W = x.shape
    src = (x > 0).to(torch.uint8).contiguous()
    out = torch.empty_like(src)
    for i in range(B * C):
        call("sb_dilate_box", ptr(src.view(-1, H, W)[i]), H, W, int(kernel), 0, 0, ptr(out.view(-1, H, W)[i]), stream())
    return out.to(dtype)
