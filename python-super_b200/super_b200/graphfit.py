"""The autograd optimiser of the reference (GraphFit.deform_superedg, /root/reference/super/deform_mesh.py:251-379)
on the device: per iteration  point-plane (+seg weights) kernel -> graph regularisers kernel -> [boundary-morph
kernel] -> optimiser step kernel.  deform_verts, its gradient, the optimiser moments and the loss trace stay on the
device; the loop issues no host synchronisation.  No autograd tape, no renderer call, no empty_cache()
(deform_mesh.py:294-298,370 are pure overhead in the reference).
"""
from __future__ import annotations

import torch

from . import lib
from .lib import call, ptr, stream

F64, I32 = torch.float64, torch.int32
OPTIMIZERS = {"SGD": 0, "Adam": 1}
TRACE_COLS = ("total", "face_losses", "arap_loss", "rot_loss", "point_plane_loss", "sf_bn_morph_loss", "morph_count", "_")


class GraphFitWorkspace:
    def __init__(self, J, device):
        n = 7 * (J + 1)
        self.J = J
        self.dv = torch.zeros((J + 1, 7), dtype=F64, device=device)
        self.ident = torch.tensor([[1., 0, 0, 0, 0, 0, 0]], dtype=F64, device=device).repeat(J + 1, 1)
        self.grad = torch.zeros(n, dtype=F64, device=device)
        self.grad_morph = torch.zeros(n, dtype=F64, device=device)
        self.acc = torch.zeros(6, dtype=F64, device=device)
        self.state = torch.zeros(2 * n, dtype=F64, device=device)
        self.trace = torch.zeros((64, 8), dtype=F64, device=device)
        self.grad_log = None

    def read_trace(self, iters):
        """Loss terms per iteration as a list of dicts (synchronises)."""
        t = self.trace[:iters].cpu().numpy()
        return [dict(zip(TRACE_COLS[:6], row[:6])) for row in t]


def edge_points(seg, num_classes, H, W, margin=1, kernel=3):
    """Per-class edge pixels of the new frame's label map (deform_mesh.py:144-162 + utils/utils.py:276-301): class
    pixels with a non-class pixel in their 3x3 window, outside a `kernel`-wide image border and the margin frame.
    seg (H,W) integer CUDA tensor -> (edge_pts (E,2) f64 [x,y] classes concatenated, edge_off (C+1,) i32).
    Index selection only (once per frame); the distances and the loss are computed in sb_gf_morph."""
    import torch.nn.functional as Fn
    seg = seg.reshape(1, 1, H, W)
    pts, off = [], [0]
    for c in range(num_classes):
        mask = seg == c
        near_other = Fn.max_pool2d((~mask).float(), kernel, stride=1, padding=kernel // 2) > 0
        e = (near_other & mask)[0, 0]
        e[:kernel] = False; e[-kernel:] = False; e[:, :kernel] = False; e[:, -kernel:] = False
        ey, ex = e.nonzero(as_tuple=True)
        ok = (ex >= margin) & (ex < W - 1 - margin) & (ey >= margin) & (ey < H - 1 - margin)
        pts.append(torch.stack([ex[ok], ey[ok]], dim=1).to(F64))
        off.append(off[-1] + int(ok.sum()))
    return torch.cat(pts).contiguous(), torch.tensor(off, dtype=I32, device=seg.device)


def graph_fit(sf, maps, cam, opt, ws=None, n_dev=None, seg=None, log_grad=False):
    """sf: points (N,3) f64, knn_indices (N,4) i32, knn_w (N,4) f64, isStable (N,) u8|None, ED (points, knn_indices i32,
    knn_w, triangles (3,F) i32, triangles_areas).  maps: (vmap, nmap).  seg (semantic terms): namespace with sf_seg (N,)
    i32, sf_seg_conf (N,C) f64, trg_seg_conf (P,C) f64, scores (C,H,W) f64, edge_pts, edge_off.
    Returns (deform_verts (J+1,7) f64 -- a view of the workspace, workspace)."""
    lib.load()
    ed = sf.ED
    J = ed.points.shape[0]
    dev = sf.points.device
    if ws is None or ws.J != J:
        ws = GraphFitWorkspace(J, dev)
    if opt.optimizer not in OPTIMIZERS:
        raise NotImplementedError(f"optimizer {opt.optimizer}: the reference's 'LM' branch of GraphFit is dead code "
                                  "(undefined names, deform_mesh.py:329-368)")
    ws.dv.copy_(ws.ident)                                     # identity every frame (deform_mesh.py:268-270)
    ws.grad.zero_(); ws.grad_morph.zero_(); ws.acc.zero_()
    iters = int(opt.num_optimize_iterations)
    if log_grad:
        ws.grad_log = torch.zeros((iters, J + 1, 7), dtype=F64, device=dev)
        ws.dv_log = torch.zeros((iters, J + 1, 7), dtype=F64, device=dev)
    vmap, nmap = maps
    n_cap = sf.points.shape[0]
    hard = bool(getattr(opt, "sf_hard_seg_point_plane", False))
    soft = bool(getattr(opt, "sf_soft_seg_point_plane", False))
    use_pp = bool(opt.sf_point_plane) or hard or soft
    seg_mode = 1 if soft else (2 if hard else 0)
    use_morph = bool(getattr(opt, "sf_bn_morph", False))
    stable = getattr(sf, "isStable", None)
    if stable is not None and stable.dtype != torch.uint8:
        stable = stable.to(torch.uint8)
    C = int(getattr(opt, "num_classes", 0) or 0) if (seg_mode or use_morph) else 0
    tri = getattr(ed, "triangles", None)
    F_ = 0 if tri is None else tri.shape[1]
    for it in range(iters):
        if log_grad:
            ws.dv_log[it].copy_(ws.dv)
        if use_pp:
            call("sb_gf_data", ptr(sf.points), ptr(sf.knn_indices), ptr(sf.knn_w), ptr(stable), n_cap, ptr(n_dev),
                 ptr(ed.points), J, ptr(ws.dv), ptr(vmap), ptr(nmap), cam.H, cam.W, cam.c,
                 float(opt.sf_point_plane_weight), seg_mode, C, ptr(seg.sf_seg) if seg_mode else None,
                 ptr(seg.sf_seg_conf) if seg_mode else None, ptr(seg.trg_seg_conf) if seg_mode else None,
                 ptr(ws.grad), ptr(ws.acc), stream())
        call("sb_gf_reg", ptr(ed.points), ptr(ed.knn_indices), ptr(ed.knn_w), J, ptr(tri),
             ptr(getattr(ed, "triangles_areas", None)), F_, ptr(ws.dv), float(opt.mesh_arap_weight),
             float(opt.mesh_rot_weight), float(opt.mesh_face_weight), int(bool(opt.mesh_arap)), int(bool(opt.mesh_rot)),
             int(bool(opt.mesh_face)), ptr(ws.grad), ptr(ws.acc), stream())
        if use_morph:
            call("sb_gf_morph", ptr(sf.points), ptr(sf.knn_indices), ptr(sf.knn_w), ptr(stable), n_cap, ptr(n_dev),
                 ptr(ed.points), J, ptr(ws.dv), cam.H, cam.W, cam.c, ptr(seg.scores), C, ptr(seg.sf_seg),
                 ptr(seg.edge_pts), ptr(seg.edge_off), ptr(ws.grad_morph), ptr(ws.acc), stream())
        call("sb_gf_step", ptr(ws.dv), ptr(ws.grad), ptr(ws.grad_morph), ptr(ws.acc),
             float(getattr(opt, "sf_bn_morph_weight", 0.0)), int(use_morph), J, OPTIMIZERS[opt.optimizer],
             float(opt.learning_rate), it, ptr(ws.state), ptr(ws.trace), ptr(ws.grad_log[it]) if log_grad else None,
             stream())
    return ws.dv, ws


def update_global(points, norms, ed_points, ed_norms, dv, n_dev=None):
    """The global-row part of Surfels.update (nodes.py:204-205,211-212,219-222); call after ops.warp_update(dv[:J])."""
    J = ed_points.shape[0]
    call("sb_gf_global_update", ptr(points), ptr(norms), points.shape[0], ptr(n_dev), ptr(ed_points), ptr(ed_norms), J,
         ptr(dv[J]), stream())
