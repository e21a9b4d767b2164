"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch float64 + autograd) of the reference's autograd
optimiser GraphFit (/root/reference/super/deform_mesh.py:25-379) and of the autograd forms of its loss
terms (/root/reference/super/loss.py:9-100,293-401,458-473,502-505).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path
(super_b200.graphfit over csrc/graphfit.cu) never does.  Pinned against the unmodified reference by
tests/golden/gf_*.npz (oracle/gen_golden.py, tests/test_oracle_golden.py).

deform_verts: (J+1, 7) float64, row J = the global transform [q_g | t_g]; identity = [1,0,0,0,0,0,0].
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import super_oracle as so

F64 = torch.float64


def deform_source(sf, dv):
    """deform_mesh.py:198-230 -> (new ED positions (J,3), warped stable surfels (Ns,3))."""
    st = sf.isStable
    idx, w = sf.knn_indices[st], sf.knn_w[st]
    g = sf.ED.points[idx]                                        # (Ns,4,3)
    d = sf.points[st][:, None, :] - g
    new_verts = sf.ED.points + dv[:-1, 4:]
    tv, _ = so.quat_apply(d, dv[idx])                            # R(q_k) d + b_k     super/utils.py:53-57
    new_sf = ((tv + g) * w[..., None]).sum(1)                     # Trans_points       super/utils.py:28-33
    new_verts, _ = so.quat_apply(new_verts, dv[-1:, 0:4])
    new_verts = new_verts + dv[-1:, 4:]
    new_sf, _ = so.quat_apply(new_sf, dv[-1:, 0:4])
    new_sf = new_sf + dv[-1:, 4:]
    return new_verts, new_sf


def bilinear_sample(features, v, u, index_map):
    """loss.py:9-100 with fill='zero', grad=False: (sampled (n,C), all-four-corners-valid (n,))."""
    fl_v, ce_v, fl_u, ce_u = torch.floor(v), torch.ceil(v), torch.floor(u), torch.ceil(u)
    nb = torch.stack([fl_v, fl_v, ce_v, ce_v], dim=-1)
    mb = torch.stack([fl_u, ce_u, fl_u, ce_u], dim=-1)
    im = index_map[nb.long(), mb.long()]
    ok = im >= 0
    U = torch.zeros(nb.shape + (features.shape[-1],), dtype=features.dtype)
    U[ok] = features[im[ok]]
    nb = torch.clamp(1 - torch.abs(nb - v[:, None]), min=0)[..., None]
    mb = torch.clamp(1 - torch.abs(mb - u[:, None]), min=0)[..., None]
    out = torch.sum(U * nb * mb, dim=-2)
    return out, ok.all(dim=-1)


def point_plane_loss(opt, nd, new_sf, src_seg=None, src_seg_conf=None, soft_seg=None):
    """DataLoss.autograd_forward, loss_type 'point-plane' (loss.py:293-401)."""
    H, W = opt.height, opt.width
    K = nd.K
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    X, Y, Z = new_sf[:, 0], new_sf[:, 1], new_sf[:, 2] + 1e-8
    u_ = X * fx / Z + cx
    v_ = Y * fy / Z + cy
    m = 1
    ur, vr = torch.round(u_), torch.round(v_)                                 # validity tests the ROUNDED coordinates
    valid = (vr >= m) & (vr < H - 1 - m) & (ur >= m) & (ur < W - 1 - m)      # utils/utils.py:176-182, valid_margin=1
    u, v = u_[valid], v_[valid]
    feats = [nd.points, nd.norms] + ([nd.seg_conf] if src_seg is not None else [])
    samp, sv = bilinear_sample(torch.cat(feats, dim=-1), v, u, nd.index_map)
    o, n = samp[:, 0:3], samp[:, 3:6]
    losses = ((n[sv] * (new_sf[valid][sv] - o[sv])).sum(-1)) ** 2
    if src_seg is not None:
        trg_conf = samp[:, 6:].softmax(1)                                    # softmax of an already-softmaxed map (:357)
        if soft_seg:
            wts = torch.exp(-0.1 * so.jsd(src_seg_conf[valid], trg_conf))
        else:
            wts = (src_seg[valid] == torch.argmax(trg_conf, dim=1)).to(F64)
        losses = losses * wts[sv].detach()
    return losses.sum()


def arap_loss(ed, beta):
    """ARAPLoss.autograd_forward (loss.py:458-473): knn_w-weighted, the rest vector passes through float32."""
    idx = ed.knn_indices
    d = ed.points[:, None, :] - ed.points[idx]
    tv, _ = so.quat_apply(d, beta[idx])
    r = tv - (d.to(torch.float32) + beta[:, None, 4:7])
    return (ed.knn_w * (r ** 2).sum(-1)).sum()


def rot_loss(dv):
    """RotLoss.autograd_forward (loss.py:502-505), all J+1 rows."""
    return ((1.0 - (dv[:, 0:4] ** 2).sum(1)) ** 2).sum()


def face_loss(ed, new_verts):
    """deform_mesh.py:51-60."""
    t = ed.triangles
    c = torch.cross(new_verts[t[1]] - new_verts[t[0]], new_verts[t[2]] - new_verts[t[0]], dim=1)
    a = 0.5 * torch.sqrt((c ** 2).sum(1) + 1e-13)
    return ((a - ed.triangles_areas) ** 2).sum()


def edge_points(opt, nd, margin=1):
    """Per class: edge pixels (x,y) f64 of the new frame's segmentation (deform_mesh.py:144-162)."""
    H, W = opt.height, opt.width
    out = []
    for c in range(opt.num_classes):
        e = so.find_edge_region(nd.seg_in, opt.num_classes, c, 3)
        ey, ex = e[0, 0].nonzero(as_tuple=True)
        ok = (ex >= margin) & (ex < W - 1 - margin) & (ey >= margin) & (ey < H - 1 - margin)
        out.append(torch.stack([ex[ok], ey[ok]], dim=1).to(F64))
    return out


def bn_morph_loss(opt, nd, new_sf, sf_seg, edge_pts):
    """Semantic boundary-morph term (deform_mesh.py:126-194).  Returns None when no surfel qualifies."""
    H, W = opt.height, opt.width
    K = nd.K
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    Z = new_sf[:, 2] + 1e-8
    x = new_sf[:, 0] * fx / Z + cx
    y = new_sf[:, 1] * fy / Z + cy
    grid = torch.stack([x, y], dim=1)
    sg = torch.stack([x / W * 2 - 1, y / H * 2 - 1], dim=1).to(F64)
    new_seg = F.grid_sample(nd.seg_conf_in, sg[None, :, None, :].detach(),
                            align_corners=False)[0, :, :, 0].argmax(0)      # raw (1,C,H,W) f64 scores
    val = (new_seg != sf_seg) & (sg[:, 0] > -1) & (sg[:, 0] < 1) & (sg[:, 1] > -1) & (sg[:, 1] < 1)
    parts = []
    for c in range(opt.num_classes):
        mk = (sf_seg == c) & val
        if mk.any() and len(edge_pts[c]) > 0:
            gm = grid[mk]
            dist, ids = so.knn(gm.detach(), edge_pts[c], 2)
            d2e = torch.minimum(torch.minimum(gm.min(1).values, W - gm[:, 0]), H - gm[:, 1])
            ok = ~torch.any(dist > d2e[:, None], dim=1)
            gt = edge_pts[c][ids]
            l = ((gt[ok] - gm[ok][:, None, :]) ** 2).sum(2).mean(1)
            parts.append(l[l > 15])
    if not parts:
        return None
    return opt.sf_bn_morph_weight * torch.cat(parts).mean()


def get_losses(opt, sf, nd, dv, edge_pts=None):
    """GraphFit.get_losses (deform_mesh.py:25-196) -> (total, dict of terms)."""
    new_verts, new_sf = deform_source(sf, dv)
    losses = {}
    if opt.mesh_face:
        losses["face_losses"] = opt.mesh_face_weight * face_loss(sf.ED, new_verts)
    if opt.mesh_arap:
        losses["arap_loss"] = opt.mesh_arap_weight * arap_loss(sf.ED, dv[:-1])
    if opt.mesh_rot:
        losses["rot_loss"] = opt.mesh_rot_weight * rot_loss(dv)
    hard = bool(getattr(opt, "sf_hard_seg_point_plane", False))
    soft = bool(getattr(opt, "sf_soft_seg_point_plane", False))
    if opt.sf_point_plane or hard or soft:
        if hard or soft:
            st = sf.isStable
            pp = point_plane_loss(opt, nd, new_sf, sf.seg[st], sf.seg_conf[st], soft)
        else:
            pp = point_plane_loss(opt, nd, new_sf)
        losses["point_plane_loss"] = opt.sf_point_plane_weight * pp
    if getattr(opt, "sf_bn_morph", False):
        bm = bn_morph_loss(opt, nd, new_sf, sf.seg[sf.isStable], edge_pts)
        if bm is not None:
            losses["sf_bn_morph_loss"] = bm
    total = sum(losses.values())
    return total, losses


def graph_fit(opt, sf, nd, trace=None):
    """GraphFit.deform_superedg (deform_mesh.py:251-379), optimizer SGD(momentum 0.9) | Adam, fresh per frame.
    Returns deform_verts (J+1,7) detached."""
    J = sf.ED.num
    dv = torch.tensor([[1., 0, 0, 0, 0, 0, 0]], dtype=F64).repeat(J + 1, 1).requires_grad_(True)
    if opt.optimizer == "SGD":
        optim = torch.optim.SGD([dv], lr=opt.learning_rate, momentum=0.9)
    elif opt.optimizer == "Adam":
        optim = torch.optim.Adam([dv], lr=opt.learning_rate)
    else:
        raise NotImplementedError(opt.optimizer)
    edge_pts = None
    if getattr(opt, "sf_bn_morph", False):
        edge_pts = edge_points(opt, nd)
    for _ in range(opt.num_optimize_iterations):
        optim.zero_grad()
        loss, losses = get_losses(opt, sf, nd, dv, edge_pts)
        loss.backward()
        dv.grad[-1] = dv.grad[-1] / J
        if trace is not None:
            trace.append({"deform_in": dv.detach().clone(), "loss": float(loss),
                          "losses": {k: float(x) for k, x in losses.items()}, "grad": dv.grad.clone()})
        optim.step()
    return dv.detach()
