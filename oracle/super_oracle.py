"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch float64 on CPU) of the reference's per-frame
ED tracking loop.  It is the CHECKER for the CUDA path (tests/, __graft_entry__.smoke(), and the
`cpu_baseline` / `--impl reference` legs of bench.py).  The product package never imports it.

Parity status: PINNED against outputs of the unmodified reference (/root/reference) run in the
build container under oracle/ref_shims.py -- see oracle/gen_golden.py (writes tests/golden/*.npz)
and oracle/validate_port.py (full-array comparison).  Not pinned by the reference's own tests
(there are none) and kNN is DEFINED by us (pytorch3d absent: exact f64 squared
distance, ascending, ties -> lower index).

Every function cites the reference lines it restates.  State is a plain namespace with the
reference's tensor layouts.
"""
from __future__ import annotations

import math
from types import SimpleNamespace as NS

import numpy as np
import torch
import torch.nn.functional as F

F64 = torch.float64


# ----------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------
def knn(p1, p2, K, chunk=8192):
    """find_knn -> pytorch3d knn_points (/root/reference/utils/utils.py:212-220).  Returns
    (sqrt(d2) ascending (N,K), idx (N,K) int64).  Definition: d2 = ((dx*dx + dy*dy) + dz*dz) in f64,
    ascending, ties -> lower index."""
    d_out, i_out = [], []
    for s in range(0, p1.shape[0], chunk):
        diff = p1[s:s + chunk, None, :] - p2[None, :, :]
        d2 = diff[..., 0] * diff[..., 0]
        for c in range(1, diff.shape[-1]):
            d2 = d2 + diff[..., c] * diff[..., c]
        ds, idx = torch.sort(d2, dim=1, stable=True)
        d_out.append(ds[:, :K])
        i_out.append(idx[:, :K])
    if not d_out:
        return torch.zeros((0, K), dtype=p1.dtype), torch.zeros((0, K), dtype=torch.long)
    return torch.sqrt(torch.cat(d_out)), torch.cat(i_out)


def knn_class(p1, p2, K, seg1, seg2, num_classes):
    """find_knn with num_classes > 0 (/root/reference/utils/utils.py:222-242, --hard_seg): neighbours are searched
    among the reference points of the query's own class; indices are global; rows of an absent class keep 1e8 / -1."""
    d = 1e8 * torch.ones((len(p1), K), dtype=F64)
    idx = -torch.ones((len(p1), K), dtype=torch.long)
    for c in range(num_classes):
        m1 = seg1 == c
        i2 = (seg2 == c).nonzero(as_tuple=True)[0]
        if int(m1.sum()) == 0 and len(i2) == 0:
            continue
        assert int(m1.sum()) > 0 and len(i2) >= K
        dc, ic = knn(p1[m1], p2[i2], K)
        d[m1] = dc
        idx[m1] = i2[ic]
    return d, idx


def kld(P, Q, eps=1e-13):
    """/root/reference/utils/utils.py:244-250"""
    return (P * (P / (Q + eps) + eps).log()).sum(-1)


def jsd(P, Q, eps=1e-13):
    """/root/reference/utils/utils.py:252-254"""
    M = 0.5 * (P + Q)
    return 0.5 * (kld(P, M, eps) + kld(Q, M, eps))


def project(K, pts, height, width, margin=0):
    """pcd2depth (/root/reference/utils/utils.py:161-184): returns unrounded (v,u), rounded pixel id
    v*W+u (round half to even) and the in-image mask of the ROUNDED coordinates."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]      # float32 0-dim tensors, promoted to f64
    X, Y, Z = pts[..., 0], pts[..., 1], pts[..., 2] + 1e-8
    u_ = X * fx / Z + cx
    v_ = Y * fy / Z + cy
    u = torch.round(u_).long()
    v = torch.round(v_).long()
    ok = (v >= margin) & (v < height - 1 - margin) & (u >= margin) & (u < width - 1 - margin)
    return v_, u_, v * width + u, ok


def quat_apply(v, beta):
    """transformQuatT without Jacobian (/root/reference/super/utils.py:41-57): q NOT normalised;
    translation added when beta has 7 columns."""
    qw, qv = beta[..., 0:1], beta[..., 1:4]
    cp = torch.cross(qv.expand_as(v), v, dim=-1)
    tv = v + 2.0 * qw * cp + 2.0 * torch.cross(qv.expand_as(v), cp, dim=-1)
    if beta.shape[-1] == 7:
        tv = tv + beta[..., 4:7]
    return tv, cp


def quat_jac(v, beta, cp):
    """d[R(q)v]/dq, (...,3,4) (/root/reference/super/utils.py:59-69):
    col 0: 2 (qv x v); cols 1..3: 2[(qv.v) I + qv v^T - 2 v qv^T - qw [v]x]."""
    qw, qv = beta[..., 0], beta[..., 1:4]
    dot = (qv * v).sum(-1)
    eye = torch.eye(3, dtype=v.dtype)
    outer = qv[..., :, None] * v[..., None, :]            # qv v^T
    vx, vy, vz = v[..., 0], v[..., 1], v[..., 2]
    z = torch.zeros_like(vx)
    # get_skew as the reference builds it (/root/reference/super/utils.py:4-14): stack(...,dim=3)
    # of rows [0,a3,-a2],[-a3,0,a1],[a2,-a1,0] along the LAST-but-one... it is the transpose layout:
    skew = torch.stack([torch.stack([z, -vz, vy], -1),
                        torch.stack([vz, z, -vx], -1),
                        torch.stack([-vy, vx, z], -1)], -2)  # standard [v]x
    dqv = 2.0 * (dot[..., None, None] * eye + outer - 2.0 * outer.transpose(-1, -2)
                 - qw[..., None, None] * skew)
    return torch.cat([2.0 * cp[..., :, None], dqv], dim=-1)


def softmax_exp_weights(dists, radii):
    """softmax_k(exp(-d_k / r_k)) (/root/reference/super/nodes.py:167,191,484,511)."""
    return F.softmax(torch.exp(-dists / radii), dim=-1)


# ----------------------------------------------------------------------------------------------
# per-frame input producer (depth_preprocessing, superv1/superv2 + --load_depth)
# ----------------------------------------------------------------------------------------------
def dilate(x, kernel):
    """torch_dilate (/root/reference/utils/utils.py:152-157): box filter > 0, padding='same'."""
    k = torch.ones((1, 1, kernel, kernel), dtype=x.dtype)
    return F.conv2d(x, k, padding="same") > 0


def fma32(a, b, c):
    """Correctly rounded float32 fma(a, b, c) on float32 tensors, machine-independent: the product is exact in
    float64, the sum is rounded to ODD in float64 (TwoSum error term), and the cast to float32 then rounds once."""
    p = a.double() * b.double()
    c = c.double().expand_as(p)
    s = p + c
    bb = s - p
    e = (p - (s - bb)) + (c - bb)
    even = (s.view(torch.int64) & 1) == 0
    fix = even & (e != 0) & torch.isfinite(s)
    toward = torch.where(e > 0, torch.full_like(s, math.inf), torch.full_like(s, -math.inf))
    s = torch.where(fix, torch.nextafter(s, toward), s)
    return s.float()


EXP_LOG2E = 1.4426950408889634        # the literals below are part of the DEFINITION (csrc/preprocess.cu holds the same)
EXP_LN2_HI = 0.693147180369123816490  # ln 2 with the low 21 mantissa bits cleared: n * LN2_HI is exact for |n| < 2^21
EXP_LN2_LO = 1.90821492927058770002e-10
EXP_COEF = [1.0 / math.factorial(k) for k in range(14)]


def exp32_def(x):
    """exp of a float32 tensor, returned as float32.  DEFINED here (and restated instruction for instruction in
    csrc/preprocess.cu exp32_def) because the reference's torch.exp on CPU float32 is MKL VML's vsExp, which is neither
    correctly rounded nor available as source: n = rint(x log2 e), r = (x - n LN2_HI) - n LN2_LO, degree-13 Taylor
    polynomial by Horner in float64 with separate multiplies and adds (no FMA), 2^n p rounded once to float32.  Agrees
    with the correctly rounded exp except where exp(x) lies within ~1e-16 relative of a float32 rounding boundary;
    differs from the reference's vsExp in the last float32 bit on ~1.2 % of arguments (oracle/validate_port.py)."""
    xd = x.double()
    n = torch.round(xd * EXP_LOG2E)
    r = (xd - n * EXP_LN2_HI) - n * EXP_LN2_LO
    p = torch.full_like(r, EXP_COEF[13])
    for k in range(12, -1, -1):
        p = p * r + EXP_COEF[k]
    return torch.ldexp(p, n.to(torch.int32)).float()


def normals_8(points, colors, exp=None):
    """getN with colours (/root/reference/utils/data_loader.py:546-583): ring neighbours in order
    L, LU, U, RU, R, RD, D, DL; h_a = (p_a - p_c) * exp(-mean|c_a - c_c|); N = sum_{a<b} h_a x h_b, normalised.
    float32 throughout.  Every step whose rounding depends on how torch's CPU kernels happen to be compiled is spelt
    out, so that the result is the same on any machine (measured bit-identical to the reference's own ops on the
    build container except for exp, see exp32_def):
      mean over 3 channels   ((a0 + a1) + a2) / 3
      cross product          fma(a1, b2, -(a2 b1)) per component (torch.linalg.cross, CPU float32)
      sum of the 7 terms     torch's 4-accumulator row sum: (((((t0 + t4) + t5) + t6) + t1) + t2) + t3
      F.normalize            n = sqrt(fma(z, z, fma(y, y, x x))) with the correctly rounded sqrt, N / max(n, 1e-12)
    exp: the float32 exponential; None = exp32_def (the definition the CUDA path is held to), torch.exp = the
    reference's own (used by the tests that pin this port against the reference's golden vectors)."""
    exp = exp or exp32_def
    b, h, w, _ = points.shape
    col = F.pad(colors.permute(0, 2, 3, 1), (0, 0, 1, 1, 1, 1), value=float("nan"))
    pts = F.pad(points, (0, 0, 1, 1, 1, 1), value=float("nan"))
    offs = [(0, -1), (-1, -1), (-1, 0), (-1, 1), (0, 1), (1, 1), (1, 0), (1, -1)]   # (dy,dx)
    cc = col[:, 1:-1, 1:-1, :].reshape(-1, 3)
    pc = pts[:, 1:-1, 1:-1, :].reshape(-1, 3)
    hs = []
    for dy, dx in offs:
        ca = col[:, 1 + dy:h + 1 + dy, 1 + dx:w + 1 + dx, :].reshape(-1, 3)
        pa = pts[:, 1 + dy:h + 1 + dy, 1 + dx:w + 1 + dx, :].reshape(-1, 3)
        ad = torch.abs(ca - cc)
        m = ((ad[:, 0:1] + ad[:, 1:2]) + ad[:, 2:3]) / 3.0
        hs.append((pa - pc) * exp(-m))

    def cross(a, bb):
        return torch.stack([fma32(a[:, 1], bb[:, 2], -(a[:, 2] * bb[:, 1])),
                            fma32(a[:, 2], bb[:, 0], -(a[:, 0] * bb[:, 2])),
                            fma32(a[:, 0], bb[:, 1], -(a[:, 1] * bb[:, 0]))], dim=1)
    t = []
    for a in range(7):
        rest = hs[a + 1]
        for bb in range(a + 2, 8):
            rest = rest + hs[bb]
        t.append(cross(hs[a], rest))
    N = (((((t[0] + t[4]) + t[5]) + t[6]) + t[1]) + t[2]) + t[3]
    n2 = fma32(N[:, 2], N[:, 2], fma32(N[:, 1], N[:, 1], N[:, 0] * N[:, 0]))
    nn = torch.from_numpy(np.sqrt(n2.numpy())).clamp_min(1e-12)     # IEEE sqrt (torch.sqrt on CPU float32 is not)
    N = (N / nn[:, None]).reshape(b, h, w, 3)
    return N, ~torch.any(torch.isnan(N), -1)


def find_edge_region(seg, num_classes, class_id, kernel, ignore_img_edge=True):
    """find_edge_region for one class (/root/reference/utils/utils.py:276-301): class pixels with a
    non-class pixel inside the kernel window; `kernel`-wide image border cleared."""
    mask = (seg == class_id)
    not_mask = (~mask).to(torch.float32)
    edge = dilate(not_mask, kernel) & mask
    if ignore_img_edge:
        edge[:, :, 0:kernel] = False
        edge[:, :, -kernel:] = False
        edge[:, :, :, 0:kernel] = False
        edge[:, :, :, -kernel:] = False
    return edge


def preprocess(opt, frame, ref_exp=False):
    """depth_preprocessing (/root/reference/utils/data_loader.py:333-523) for --load_depth inputs.
    `frame`: dict from super_b200.synth.frame_inputs (numpy).  Returns new_data namespace:
    points,norms (Nv,3) f64; colors (Nv,3) f32; radii (Nv,) f64; confs (Nv,) f32; valid (P,) bool;
    index_map (H,W) i64; time; [seg, seg_conf, dist2edge].
    ref_exp: evaluate the two float32 exponentials (normal weights, confidences) with torch.exp like the reference
    instead of exp32_def -- for the golden-vector pin tests only."""
    exp = torch.exp if ref_exp else exp32_def
    H, W = opt.height, opt.width
    depth = torch.from_numpy(frame["depth"]).clone()[None]          # (1,1,H,W) f32
    color = torch.from_numpy(frame["color"])[None]                  # (1,3,H,W) f32
    K = torch.from_numpy(frame["K"])[None]
    inv_K = torch.from_numpy(frame["inv_K"])[None]
    has_seg = "seg_conf" in frame
    if has_seg:
        seg_conf_in = torch.from_numpy(frame["seg_conf"])[None]     # (1,C,H,W) f64 scores
        seg_in = seg_conf_in[0].argmax(0, True).long()[None]        # (1,1,H,W)

    # BackprojectDepth (/root/reference/depth/monodepth2/layers.py:139-167), float32
    xs, ys = np.meshgrid(range(W), range(H), indexing="xy")
    pix = torch.from_numpy(np.stack([xs.reshape(-1), ys.reshape(-1), np.ones(H * W)], 0)
                           .astype(np.float32))[None]
    # The reference does torch.matmul(inv_K[:, :3, :3], pix) in float32.  Its CPU BLAS evaluates each
    # entry as the FMA chain  t = a*x; t = fma(b, y, t); t = fma(c, 1, t)  (measured here, bit for bit:
    # oracle/validate_port.py); the restatement spells that chain out so that it does not depend on the
    # BLAS build of the machine it runs on (fma32: exact float32 fma).
    ik = inv_K[0, :3, :3]
    px, py, p1 = pix[0, 0:1], pix[0, 1:2], pix[0, 2:3]
    t = ik[:, 0:1] * px
    t = fma32(ik[:, 1:2].expand(3, H * W), py.expand(3, H * W), t)
    cam = fma32(ik[:, 2:3].expand(3, H * W), p1.expand(3, H * W), t)[None]
    cam = depth.view(1, 1, -1) * cam
    pcd = cam.reshape(1, 3, H, W).permute(0, 2, 3, 1).clone()

    if opt.data == "superv1":                                        # :374-405
        inval = torch.zeros_like(depth).bool()
        for c in getattr(opt, "del_seg_classes", []):
            inval |= seg_in == c
        if opt.dilate_invalid_kernel > 0:
            inval = ~dilate((~inval).float(), opt.dilate_invalid_kernel)
            inval = dilate(inval.float(), 2 * opt.dilate_invalid_kernel)
        inval |= depth <= 0
        inval |= depth > 1.5
        depth[inval] = float("nan")
        pcd[inval[:, 0]] = float("nan")
    else:                                                            # superv2 :407-432
        inval = torch.zeros(depth.size(), dtype=torch.bool)
        inval |= depth == 0
        inval[:, :, 0:int(0.1 * W)] = True       # quirk: indexes dim 2 = ROWS (Appendix B #10)
        for c in getattr(opt, "del_seg_classes", []):
            inval |= seg_in == c
        depth[inval] = float("nan")
        pcd[inval[0]] = float("nan")

    norms, valid = normals_8(pcd, color, exp)                           # :437-441
    valid &= ~torch.any(torch.isnan(pcd), dim=3)
    Z = -depth
    points = pcd[0].type(F64)
    norms = norms[0].type(F64)
    valid = valid[0]
    points = points[valid]
    norms = norms[valid]
    index_map = -torch.ones((H, W), dtype=torch.long)
    index_map[valid] = torch.arange(int(valid.count_nonzero()))
    radii = Z[0, 0][valid] / (np.sqrt(2) * K[0, 0, 0] * torch.clamp(torch.abs(norms[..., 2]), 0.26, 1.0))
    U, V = torch.meshgrid(torch.arange(W), torch.arange(H), indexing="xy")
    dc2 = (2. * (U / W) - 1.) ** 2 + (2. * (V / H) - 1.) ** 2
    confs = exp(-dc2 * frame["divterm"])
    colors = color[0].permute(1, 2, 0)[valid]
    nd = NS(points=points, norms=norms, colors=colors, radii=radii, confs=confs[valid],
            valid=valid.view(-1), index_map=index_map, valid_map=valid, time=int(frame["filename"]))
    if has_seg:                                                      # :494-518
        seg_conf = seg_conf_in[0].softmax(0).permute(1, 2, 0)
        nd.seg = seg_in[0, 0][valid]
        nd.seg_conf = seg_conf[valid]
        edge_pts = []
        for c in range(opt.num_classes):
            e = find_edge_region(seg_in, opt.num_classes, c, 3)
            ey, ex = e[0, 0].nonzero(as_tuple=True)
            edge_pts.append(torch.stack([ex / W, ey / H], dim=1).type(F64))
        sv, su, _, _ = project(K[0], nd.points, H, W)
        sc = torch.stack([su / W, sv / H], dim=1)
        d2e = torch.zeros_like(nd.radii)
        for c in range(opt.num_classes):
            m = nd.seg == c
            if m.any():
                dd, _ = knn(sc[m], edge_pts[c], 1)
                d2e[m] = dd[:, 0]
        nd.dist2edge = d2e
        nd.seg_in, nd.seg_conf_in = seg_in, seg_conf_in
    nd.K = K[0]
    return nd


# ----------------------------------------------------------------------------------------------
# ED graph (once per sequence)
# ----------------------------------------------------------------------------------------------
def build_graph(opt, nd):
    """init_graph + DirectDeformGraph grid_mesh branch
    (/root/reference/super/graph_encoder.py:11-67,128-193)."""
    H, W, s = opt.height, opt.width, opt.mesh_step_size
    valid = nd.index_map >= 0
    us = torch.arange(0, W - 1, s)
    vs = torch.arange(0, H - 1, s)
    uu, vv = torch.meshgrid(us, vs, indexing="xy")            # (len(vs), len(us)) row-major over (y,x)
    av = valid[vv, uu]
    u, v = uu[av], vv[av]
    nid = -torch.ones((H, W), dtype=torch.long)
    nid[v, u] = torch.arange(len(u))
    vpad = F.pad(valid, (0, s, 0, s), value=False)
    npad = F.pad(nid, (0, s, 0, s), value=-1)

    def node_at(x, y):
        ok = vpad[y, x]
        n = npad[y, x]
        return torch.where(ok, n, torch.full_like(n, -1))

    a = node_at(u, v)
    r = node_at(u + s, v)
    d = node_at(u, v + s)
    rd = node_at(u + s, v + s)
    # per anchor, edges in order (a,r),(a,rd),(a,d),(r,d); faces (a,r,rd),(a,rd,d)
    e = torch.stack([torch.stack([a, r], 1), torch.stack([a, rd], 1),
                     torch.stack([a, d], 1), torch.stack([r, d], 1)], 1).reshape(-1, 2)
    e = e[~torch.any(e < 0, dim=1)]
    f = torch.stack([torch.stack([a, r, rd], 1), torch.stack([a, rd, d], 1)], 1).reshape(-1, 3)
    f = f[~torch.any(f < 0, dim=1)]
    edge_index, triangles = e.t().contiguous(), f.t().contiguous()

    sel = nd.index_map[nid >= 0]                               # row-major over pixels == node order
    g = NS()
    g.points = nd.points[sel].clone()
    g.norms = nd.norms[sel].clone()
    if hasattr(nd, "seg_conf"):
        g.seg_conf = nd.seg_conf[sel].clone()
        g.seg = torch.argmax(g.seg_conf, dim=1)
        if getattr(opt, "hard_seg", False) and opt.mesh_face:
            ie = g.seg[edge_index[0]] == g.seg[edge_index[1]]
            edge_index = edge_index[:, ie]
            it = (g.seg[triangles[0]] == g.seg[triangles[1]]) & (g.seg[triangles[0]] == g.seg[triangles[2]])
            triangles = triangles[:, it]
    J = len(g.points)
    lens = torch.norm(g.points[edge_index[0]] - g.points[edge_index[1]], dim=1)
    radii = []
    for k in range(J):
        radii.append(lens[torch.any(edge_index == k, dim=0)].mean())
    radii = torch.stack(radii)
    bad = torch.isnan(radii)
    if bad.any():
        radii[bad] = radii[~bad].mean()
    ta = torch.linalg.cross(g.points[triangles[1]] - g.points[triangles[0]],
                            g.points[triangles[2]] - g.points[triangles[0]], dim=1)
    g.radii = radii
    g.edge_index, g.triangles = edge_index, triangles
    g.triangles_areas = 0.5 * torch.sqrt((ta ** 2).sum(1) + 1e-13)
    g.edges_lens = torch.linalg.norm(g.points[edge_index[0]] - g.points[edge_index[1]], dim=-1)
    g.num, g.param_num = J, 7 * J
    return g


# ----------------------------------------------------------------------------------------------
# Surfel state
# ----------------------------------------------------------------------------------------------
def semantic_weights(ed, knn_idx, seg_conf, dists, radii):
    """/root/reference/super/nodes.py:183-189 (power_arg = (1/2, 1/2))."""
    P = ed.seg_conf[knn_idx]
    Q = seg_conf[:, None, :]
    return F.softmax(torch.pow(torch.exp(-jsd(P, Q)), 0.5) * torch.pow(torch.exp(-dists / radii), 0.5),
                     dim=-1)


def init_surfels(opt, nd, graph):
    """Surfels.__init__ + update_ed + update_sfed_knn + first compaction
    (/root/reference/super/nodes.py:93-191, super/super.py:60-63)."""
    sf = NS()
    for k in ("points", "norms", "colors", "radii", "confs"):
        setattr(sf, k, getattr(nd, k).clone())
    for k in ("seg", "seg_conf", "dist2edge"):
        if hasattr(nd, k):
            setattr(sf, k, getattr(nd, k).clone())
    sf.time = nd.time
    N = len(sf.points)
    sf.isStable = torch.ones(N, dtype=torch.bool)
    sf.time_stamp = sf.time * torch.ones(N)
    sf.projdata = torch.flip(nd.valid.view(opt.height, opt.width).nonzero(), dims=[-1]).type(torch.float32)
    sf.ED = graph
    sf.semantic = (opt.method == "semantic-super")
    sf.track_id = None
    # update_ed :154-168  (divides by the QUERY node's radius)
    hard = bool(getattr(opt, "hard_seg", False))
    if hard:
        d, idx = knn_class(graph.points, graph.points, opt.num_ED_neighbors + 1, graph.seg, graph.seg, opt.num_classes)
    else:
        d, idx = knn(graph.points, graph.points, opt.num_ED_neighbors + 1)
    d = d[:, 1:] / graph.radii[:, None]
    graph.knn_w = F.softmax(torch.exp(-d), dim=-1)
    graph.knn_indices = idx[:, 1:]
    # update_sfed_knn :170-191
    if hard:
        d, sf.knn_indices = knn_class(sf.points, graph.points, opt.num_neighbors, sf.seg, graph.seg, opt.num_classes)
    else:
        d, sf.knn_indices = knn(sf.points, graph.points, opt.num_neighbors)
    r = graph.radii[sf.knn_indices]
    sf.isStable[~torch.any(d <= r, dim=1)] = False
    if sf.semantic and not getattr(opt, "hard_seg", False):
        sf.knn_w = semantic_weights(graph, sf.knn_indices, sf.seg_conf, d, r)
    else:
        sf.knn_w = softmax_exp_weights(d, r)
    compact(opt, sf, float(nd.time))
    return sf


PER_SURFEL = ("points", "norms", "colors", "confs", "radii", "time_stamp", "knn_indices", "knn_w",
              "projdata", "seg", "seg_conf", "dist2edge")


def compact(opt, sf, time_now):
    """prepareStableIndexNSwapAllModel (/root/reference/super/nodes.py:543-589), state part only."""
    if opt.disable_removing_unstable_surfels:
        return
    sf.isStable = sf.isStable & (time_now - sf.time_stamp < opt.th_time_steps)
    if sf.track_id is not None:
        sf.isStable[sf.track_id[sf.track_id >= 0]] = True
    keep = sf.isStable
    for k in PER_SURFEL:
        if hasattr(sf, k):
            setattr(sf, k, getattr(sf, k)[keep])
    if sf.track_id is not None:
        id_map = -torch.ones(len(keep), dtype=torch.long)
        id_map[keep] = torch.arange(int(keep.count_nonzero()))
        ok = sf.track_id >= 0
        sf.track_id[ok] = id_map[sf.track_id[ok]]
    sf.isStable = sf.isStable[keep]


def warp_points(points, ed_points, knn_idx, knn_w, beta):
    """Trans_points (/root/reference/super/utils.py:17-38) without Jacobian:
    T(p) = sum_k w_k ( R(q_k)(p - g_k) + b_k + g_k )."""
    g = ed_points[knn_idx]
    d = points[:, None, :] - g
    tv, cp = quat_apply(d, beta[knn_idx])
    tv = tv + g
    return torch.sum(knn_w[..., None] * tv, dim=-2), d, cp


def update(opt, sf, beta):
    """Surfels.update (/root/reference/super/nodes.py:193-223).  Quirk kept: b_k is added to the
    normal before normalisation (7-column beta)."""
    ed = sf.ED
    lm = opt.use_derived_gradient
    b_ = beta[sf.knn_indices]
    sf.points, _, _ = warp_points(sf.points, ed.points, sf.knn_indices, sf.knn_w, beta)
    if not lm:
        sf.points = sf.points + beta[-1:, 4:]
    n, _ = quat_apply(sf.norms[:, None, :].repeat(1, sf.knn_indices.shape[1], 1), b_)
    n = torch.sum(sf.knn_w[..., None] * n, dim=-2)
    if not lm:
        n, _ = quat_apply(n, beta[-1:, 0:4])
    sf.norms = F.normalize(n, dim=-1)
    if lm:
        ed.points = ed.points + beta[:, 4:]
        en, _ = quat_apply(ed.norms, beta[:, 0:4])
    else:
        ed.points = ed.points + beta[:-1, 4:]
        ed.points = ed.points + beta[-1:, 4:]
        en, _ = quat_apply(ed.norms, beta[:-1, 0:4])
        en, _ = quat_apply(en, beta[-1:, 0:4])
    ed.norms = F.normalize(en, dim=-1)


# ----------------------------------------------------------------------------------------------
# LM cost terms
# ----------------------------------------------------------------------------------------------
def bilinear_block(v, u, target, index_map, grad):
    """LossTool.bilinear_intrpl_block (/root/reference/super/loss.py:106-157): corners
    (fl_v,fl_u),(fl_v,ce_u),(ce_v,fl_u),(ce_v,ce_u); NaN for out-of-image / index<0 corners;
    weight max(0,1-|y-v|)*max(0,1-|x-u|); gradient order [d/du, d/dv] with sign(corner-coord>=0)."""
    fv, cv, fu, cu = torch.floor(v), torch.ceil(v), torch.floor(u), torch.ceil(u)
    nb = torch.stack([fv, fv, cv, cv], dim=1)
    mb = torch.stack([fu, cu, fu, cu], dim=1)
    h, w = index_map.shape
    nl, ml = nb.long(), mb.long()
    idx = index_map[nl.clamp(0, h - 1), ml.clamp(0, w - 1)]
    ok = (idx >= 0) & (nl >= 0) & (nl < h) & (ml >= 0) & (ml < w)
    U = torch.full(nb.shape + (target.shape[-1],), float("nan"), dtype=F64)
    U[ok] = target[idx[ok]]
    dn = (nb - v[:, None])[..., None]
    dm = (mb - u[:, None])[..., None]
    wn = torch.clamp(1 - torch.abs(dn), min=0)
    wm = torch.clamp(1 - torch.abs(dm), min=0)
    val = torch.sum(U * wn * wm, dim=1)
    if not grad:
        return val, None, (nl, ml)
    sn = torch.where(dn >= 0, 1., -1.)
    sm = torch.where(dm >= 0, 1., -1.)
    g = torch.stack([torch.sum(U * wn * sm, dim=1), torch.sum(U * wm * sn, dim=1)], dim=2)
    return val, g, (nl, ml)


def data_term(opt, sf, nd, beta, lam, grad):
    """DataLoss.prepare + forward (/root/reference/super/loss.py:212-290).
    Returns dict: ids (matched surfel rows), corners (M,4) [fl_v,ce_v,fl_u,ce_u], r (M,) residual,
    and with grad: jrow (M,4,7) = lam * d r / d beta[knn_k]."""
    H, W = opt.height, opt.width
    K = nd.K
    T, d, cp = warp_points(sf.points, sf.ED.points, sf.knn_indices, sf.knn_w, beta)
    v_, u_, coords, _ = project(K, T, H, W)
    nvalid = len(nd.valid)
    pair = nd.valid[coords.clamp(0, nvalid - 1)] & (coords >= 0) & (coords < nvalid)
    ids = pair.nonzero()[:, 0]
    v, u = v_[pair], u_[pair]
    o, dodpi, (nl, ml) = bilinear_block(v, u, nd.points, nd.index_map, grad)
    n, dndpi, _ = bilinear_block(v, u, nd.norms, nd.index_map, grad)
    ok = ~torch.any(torch.isnan(o) | torch.isnan(n), dim=1)
    ids, o, n = ids[ok], o[ok], n[ok]
    Tm = T[ids]
    diff = Tm - o
    r = lam * torch.sum(n * diff, dim=1)
    out = {"ids": ids, "r": r,
           "corners": torch.stack([nl[ok][:, 0], nl[ok][:, 2], ml[ok][:, 0], ml[ok][:, 1]], 1)}
    if not grad:
        return out
    dodpi, dndpi = dodpi[ok], dndpi[ok]
    fx, fy = K[0, 0].double(), K[1, 1].double()
    Z = Tm[:, 2]
    dpi = torch.zeros((len(Tm), 2, 3), dtype=F64)               # dPi_block :161-173 (Z without 1e-8)
    dpi[:, 0, 0] = fx / Z
    dpi[:, 0, 2] = -fx * Tm[:, 0] / (Z * Z)
    dpi[:, 1, 1] = fy / Z
    dpi[:, 1, 2] = -fy * Tm[:, 1] / (Z * Z)
    O = dodpi @ dpi                                             # (M,3,3) do/dT
    Nn = dndpi @ dpi                                            # (M,3,3) dn/dT
    # row vector a^T = n^T (I - O) + diff^T Nn  == d r / d T   (:258-283)
    a = n - torch.einsum("mi,mij->mj", n, O) + torch.einsum("mi,mij->mj", diff, Nn)
    bk = beta[sf.knn_indices[ids]]
    Jq = quat_jac(d[ids], bk, cp[ids])                          # (M,4,3,4)
    w = sf.knn_w[ids]
    jq = torch.einsum("mi,mkij->mkj", a, Jq) * w[..., None]
    jb = a[:, None, :] * w[..., None]
    out["jrow"] = lam * torch.cat([jq, jb], dim=-1)
    return out


def arap_term(sf, beta, lam, grad):
    """ARAPLoss (/root/reference/super/loss.py:403-455): r_(j,k) = lam [R(q_n)(g_j-g_n) + b_n -
    (g_j-g_n) - b_j], n = N_ED(j,k)."""
    ed = sf.ED
    d = ed.points[:, None, :] - ed.points[ed.knn_indices]
    bn = beta[ed.knn_indices]
    tv, cp = quat_apply(d, bn)
    r = lam * (tv - (d + beta[:, None, 4:7]))
    if not grad:
        return r, None
    return r, lam * quat_jac(d, bn, cp)                         # (J,K,3,4) wrt q_n


def rot_term(beta, lam, grad):
    """RotLoss (/root/reference/super/loss.py:475-499) -- float32 by the reference's choice."""
    q = beta[:, 0:4].type(torch.float32)
    r = lam * (1. - torch.sum(torch.pow(q, 2), dim=1, keepdim=True))
    if not grad:
        return r, None
    return r, (-lam * 2 * q)


def lm_normal_equations(opt, sf, nd, beta, assemble="blocks"):
    """LM_Solver.prepareCostTerm(grad=True) (/root/reference/super/LM.py:54-68): dense A = sum J^T J,
    g = -sum J^T r.  assemble='sparse_mm' follows the reference's op sequence (COO Jacobian +
    torch.sparse.mm, loss.py:200-205) and is what the CPU baseline times; 'blocks' accumulates the
    same sums per 7x7 node-pair block with index_add_ (different summation order, ~1e-16 rel)."""
    J = sf.ED.num
    n = 7 * J
    A = torch.zeros((n, n), dtype=F64)
    g = torch.zeros((n, 1), dtype=F64)
    info = {}
    if opt.sf_point_plane:
        dt = data_term(opt, sf, nd, beta, opt.sf_point_plane_weight, True)
        info["data"] = dt
        kn = sf.knn_indices[dt["ids"]]
        jrow, r = dt["jrow"], dt["r"]
        M, Kn = kn.shape
        cols = (kn[:, :, None] * 7 + torch.arange(7)[None, None, :]).reshape(M, -1)
        vals = jrow.reshape(M, -1)
        if assemble == "sparse_mm":
            rows = torch.arange(M)[:, None].expand(M, 7 * Kn)
            good = ~torch.isnan(vals)
            Jm = torch.sparse_coo_tensor(torch.stack([rows[good], cols[good]]), vals[good], (M, n))
            Jt = Jm.t()
            A += torch.sparse.mm(Jt, Jm).to_dense()
            g += -torch.sparse.mm(Jt, r[:, None])
        else:
            outer = vals[:, :, None] * vals[:, None, :]
            lin = (cols[:, :, None] * n + cols[:, None, :]).reshape(-1)
            A.view(-1).index_add_(0, lin, outer.reshape(-1))
            g.view(-1).index_add_(0, cols.reshape(-1), -(vals * r[:, None]).reshape(-1))
    if opt.mesh_arap:
        r, Jq = arap_term(sf, beta, opt.mesh_arap_weight, True)
        lam = opt.mesh_arap_weight
        ed = sf.ED
        Jn, Kn = ed.knn_indices.shape
        nn = ed.knn_indices.reshape(-1)                                  # neighbour node per (j,k)
        jj = torch.arange(Jn)[:, None].expand(Jn, Kn).reshape(-1)
        rows = torch.arange(Jn * Kn * 3).reshape(-1, 3)
        Jq = Jq.reshape(-1, 3, 4)
        # 6 non-zeros per residual row: cols 7n+{0..3}, 7n+4+c (value lam), 7j+4+c (value -lam)
        c3 = torch.arange(3)[None, :]
        cols = torch.cat([(nn[:, None, None] * 7 + torch.arange(4)[None, None, :]).expand(-1, 3, -1),
                          (nn[:, None] * 7 + 4 + c3)[..., None], (jj[:, None] * 7 + 4 + c3)[..., None]], -1)
        vals = torch.cat([Jq, torch.full_like(Jq[..., :1], lam), torch.full_like(Jq[..., :1], -lam)], -1)
        Jm = torch.sparse_coo_tensor(torch.stack([rows[..., None].expand(-1, -1, 6).reshape(-1),
                                                  cols.reshape(-1)]), vals.reshape(-1),
                                     (Jn * Kn * 3, n))
        Jt = Jm.t()
        A += torch.sparse.mm(Jt, Jm).to_dense()
        g += -torch.sparse.mm(Jt, r.reshape(-1, 1))
    if opt.mesh_rot:
        r, Jr = rot_term(beta, opt.mesh_rot_weight, True)               # f32
        rows = torch.arange(J)[:, None].expand(J, 4).reshape(-1)
        cols = (torch.arange(J)[:, None] * 7 + torch.arange(4)[None, :]).reshape(-1)
        Jm = torch.sparse_coo_tensor(torch.stack([rows, cols]), Jr.reshape(-1), (J, n), dtype=torch.float32)
        Jt = Jm.t()
        A += torch.sparse.mm(Jt, Jm).to_dense()                          # f32 product added into f64
        g += -torch.sparse.mm(Jt, r)
    return A, g, info


def lm_cost(opt, sf, nd, beta):
    """LM_Solver.prepareCostTerm(grad=False) (/root/reference/super/LM.py:70-78): one torch.sum over
    the concatenation [data r^2 (f64), arap r^2 (f64), rot r^2 (f32 promoted)]."""
    terms, parts = {}, []
    if opt.sf_point_plane:
        dt = data_term(opt, sf, nd, beta, opt.sf_point_plane_weight, False)
        parts.append(torch.pow(dt["r"][:, None], 2))
        terms["data"] = torch.sum(parts[-1])
        terms["_data"] = dt
    if opt.mesh_arap:
        r, _ = arap_term(sf, beta, opt.mesh_arap_weight, False)
        parts.append(torch.pow(r.reshape(-1, 1), 2))
        terms["arap"] = torch.sum(parts[-1])
    if opt.mesh_rot:
        r, _ = rot_term(beta, opt.mesh_rot_weight, False)
        parts.append(torch.pow(r, 2))                                    # f32
        terms["rot"] = torch.sum(parts[-1])
    return torch.sum(torch.cat(parts)), terms


def lm_solve(opt, sf, nd, u=10.0, v=7.5, minimal_loss=1e10, assemble="blocks", trace=None):
    """LM_Solver.LM (/root/reference/super/LM.py:81-122): beta reset to identity, additive damping
    u (reset to 10 every frame), Cholesky solve, accept iff loss < minimal_loss."""
    J = sf.ED.num
    beta = torch.tensor([[1., 0, 0, 0, 0, 0, 0]], dtype=F64).repeat(J, 1)
    best = beta.clone()
    for it in range(opt.num_optimize_iterations):
        A, g, info = lm_normal_equations(opt, sf, nd, beta, assemble)
        rec = {"beta_in": beta.clone(), "u": u}
        if trace is not None:
            rec["A"], rec["g"] = A.clone(), g.clone()
            if "data" in info:
                rec["ids"], rec["corners"] = info["data"]["ids"], info["data"]["corners"]
        A[torch.arange(7 * J), torch.arange(7 * J)] += u
        try:
            L = torch.linalg.cholesky(A)
            delta = torch.cholesky_solve(g, L).view(-1, 7)
        except RuntimeError:
            break
        beta = beta + delta
        loss, terms = lm_cost(opt, sf, nd, beta)
        rec.update(delta=delta, beta_try=beta.clone(), loss=float(loss),
                   loss_terms={k: float(x) for k, x in terms.items() if not k.startswith("_")})
        if loss < minimal_loss:
            minimal_loss = loss
            u /= v
            best = beta.clone()
            rec["accept"] = True
        else:
            u *= v
            beta = best.clone()
            rec["accept"] = False
        if trace is not None:
            trace.append(rec)
    return beta


# ----------------------------------------------------------------------------------------------
# surfel fusion
# ----------------------------------------------------------------------------------------------
def _merge(opt, sf, i1, src, i2, time_now, add_new, src_is_new):
    """merge_data (/root/reference/super/nodes.py:301-355).  i1: surfel rows; i2: rows of `src`
    (new data or sf itself).  Returns bool mask over the pairs that merged."""
    if len(i1) == 0:
        return torch.zeros(0, dtype=torch.bool)
    p, n, c, r, w = sf.points[i1], sf.norms[i1], sf.colors[i1], sf.radii[i1], sf.confs[i1]
    p2, n2, c2, r2, w2 = src.points[i2], src.norms[i2], src.colors[i2], src.radii[i2], src.confs[i2]
    ok = (torch.linalg.norm(p - p2, dim=-1) < opt.th_dist) & (torch.sum(n * n2, dim=-1) > opt.th_cosine_ang)
    if (getattr(opt, "hard_seg", False) or opt.data == "superv1") and hasattr(sf, "seg") and hasattr(src, "seg"):
        ok &= sf.seg[i1] == src.seg[i2]
    idx = i1[ok]
    w, w2 = w[ok], w2[ok]
    ws = w + w2
    w = w / ws
    w2 = w2 / ws
    pu = w[:, None] * p[ok] + w2[:, None] * p2[ok]
    nu = w[:, None] * n[ok] + w2[:, None] * n2[ok]
    sf.radii[idx] = w * r[ok] + w2 * r2[ok]
    sf.confs[idx] = ws
    sf.points[idx] = pu
    sf.norms[idx] = F.normalize(nu, dim=-1)
    wc, wc2 = w[:, None], w2[:, None]
    if add_new:
        wn = wc2 * 3
        sf.colors[idx] = wc / (wc + wn) * c[ok] + wn / (wc + wn) * c2[ok]
    else:
        sf.colors[idx] = wc * c[ok] + wc2 * c2[ok]
    sf.time_stamp[idx] = time_now
    if hasattr(sf, "seg"):
        sc = wc * sf.seg_conf[i1][ok] + wc2 * src.seg_conf[i2][ok]
        sc = sc / sc.sum(1, keepdim=True)
        sf.seg_conf[idx] = sc
        sf.seg[idx] = torch.argmax(sc, dim=1)
    return ok


def fuse(opt, sf, nd, max_layers=16):
    """Surfels.fuseInputData (/root/reference/super/nodes.py:270-541).  Confidence ties on one pixel
    -> lower surfel index first (the reference's first sort is unstable; Appendix B #9)."""
    H, W = opt.height, opt.width
    P = H * W
    sf.time = nd.time
    valid = nd.valid.clone()
    _, _, coords, inimg = project(nd.K, sf.points, H, W)
    live = inimg & sf.isStable
    o1 = torch.sort(sf.confs, descending=True, stable=True)[1]
    cs, o2 = torch.sort(coords[o1], stable=True)
    order = o1[o2]
    live_s = live[order]
    ids = order[live_s]
    cs = cs[live_s]
    # rank of each live surfel inside its pixel's confidence-ordered list
    layers_val, layers_idx = [], []
    deleted = []
    if len(cs) > 0:
        first = torch.ones(len(cs), dtype=torch.bool)
        first[1:] = cs[1:] != cs[:-1]
        start = torch.cummax(torch.where(first, torch.arange(len(cs)), torch.zeros(len(cs), dtype=torch.long)), 0)[0]
        rank = torch.arange(len(cs)) - start
        nl = min(int(rank.max()) + 1, max_layers)
        for l in range(nl):
            m = rank == l
            vm = torch.zeros(P, dtype=torch.bool)
            im = torch.zeros(P, dtype=torch.long)
            vm[cs[m]] = True
            im[cs[m]] = ids[m]
            layers_val.append(vm)
            layers_idx.append(im)
        over = ids[rank >= max_layers]
        if len(over) > 0:
            deleted.append(over)
    time_now = nd.time
    add_valid = None
    compact_of_pixel = nd.index_map.view(-1)
    if not opt.disable_merging_new_surfels and layers_val:                 # :409-422
        add_valid = valid & ~layers_val[0]
        valid[add_valid] = False
        for vm, im in zip(layers_val, layers_idx):
            if not valid.any():
                break
            sel = valid & vm
            merged = _merge(opt, sf, im[sel], nd, compact_of_pixel[sel], time_now, True, True)
            valid[sel] = ~merged
        add_valid |= valid
    if not opt.disable_merging_exist_surfels and layers_val:               # :424-460
        L = len(layers_val)
        for i in range(L):
            cur = layers_val[i]                      # aliasing kept: cumulative &= mutates layer i
            for j in range(i + 1, L):
                cur &= layers_val[j]
                if not cur.any():
                    continue
                i1, i2 = layers_idx[i][cur], layers_idx[j][cur]
                merged = _merge(opt, sf, i1, sf, i2, time_now, False, False)
                keep = torch.ones_like(cur)
                keep[cur] = ~merged
                layers_val[j] &= keep
                gone = i2[merged]
                deleted.append(gone)
                if sf.track_id is not None:
                    tgt = i1[merged]
                    for k in range(len(sf.track_id)):
                        hit = gone == sf.track_id[k]
                        if hit.any():
                            sf.track_id[k] = tgt[hit][0]
        if deleted:
            dl = torch.unique(torch.cat(deleted))
            if sf.track_id is not None:
                for k in range(len(sf.track_id)):
                    if (dl == sf.track_id[k]).any():
                        sf.track_id[k] = -2
            sf.isStable[dl] = False
    # knn weights of ALL existing surfels with their old indices (:466-484)
    dists = torch.linalg.norm(sf.points[:, None, :] - sf.ED.points[sf.knn_indices], dim=-1)
    radii = sf.ED.radii[sf.knn_indices]
    if sf.semantic:
        sf.knn_w = semantic_weights(sf.ED, sf.knn_indices, sf.seg_conf, dists, radii)
    else:
        sf.knn_w = softmax_exp_weights(dists, radii)
    if not opt.disable_adding_new_surfels and add_valid is not None:       # :486-531
        addc = add_valid[nd.valid]
        if addc.count_nonzero() > 0:
            npts = nd.points[addc]
            if getattr(opt, "hard_seg", False):
                d, nidx = knn_class(npts, sf.ED.points, opt.num_neighbors, nd.seg[addc], sf.ED.seg, opt.num_classes)
            else:
                d, nidx = knn(npts, sf.ED.points, opt.num_neighbors)
            r = sf.ED.radii[nidx]
            ok = torch.any(d <= r, dim=1)
            if sf.semantic and not getattr(opt, "hard_seg", False):
                nw = semantic_weights(sf.ED, nidx, nd.seg_conf[addc], d, r)
            else:
                nw = softmax_exp_weights(d, r)
            k = int(ok.count_nonzero())
            sf.isStable = torch.cat([sf.isStable, torch.ones(k, dtype=torch.bool)])
            sf.knn_w = torch.cat([sf.knn_w, nw[ok]])
            sf.knn_indices = torch.cat([sf.knn_indices, nidx[ok]])
            sf.points = torch.cat([sf.points, npts[ok]])
            for key in ("norms", "colors", "seg", "seg_conf", "radii", "confs", "dist2edge"):
                if hasattr(nd, key) and hasattr(sf, key):
                    setattr(sf, key, torch.cat([getattr(sf, key), getattr(nd, key)[addc][ok]]))
            sf.time_stamp = torch.cat([sf.time_stamp, nd.time * torch.ones(k)])
    v, u, _, _ = project(nd.K, sf.points, H, W)
    sf.projdata = torch.stack([u, v], dim=1).type(torch.float32)


# ----------------------------------------------------------------------------------------------
# tracked points (evaluation bookkeeping that rides on the state)
# ----------------------------------------------------------------------------------------------
def init_track_pts(sf, nd, gt_row, th=0.2):
    """Surfels.init_track_pts (/root/reference/super/nodes.py:225-249).  gt_row (T,3) int [x, y, valid].  Quirks
    kept: the result row is projdata[track_id[k]] AFTER this call's assignment (the loop variable is a view), with
    negative ids indexing from the end (-1 -> last surfel); `gt_id > 0` drops the first valid pixel; a lost id (-2)
    masks surfel N-2."""
    T = len(sf.track_id)
    rst = torch.zeros((T, 3))
    gt = torch.as_tensor(gt_row, dtype=torch.int)
    for k in range(T):
        tid = sf.track_id[k]                       # a view: follows the assignment below, as in the reference
        x, y, v = gt[k]
        gt_id = nd.index_map[y, x]
        if tid < 0 and gt_id > 0 and v == 1:
            dists = torch.linalg.norm(sf.points - nd.points[gt_id], dim=-1)
            inval = sf.track_id[(sf.track_id >= 0) | (sf.track_id == -2)]
            if len(inval) > 0:
                dists[inval.long()] = 1e13
                dists[~sf.isStable] = 1e13
            if torch.min(dists) < th:
                sf.track_id[k] = torch.argmin(dists)
        rst[k, 0:2] = sf.projdata[tid.long()]
        rst[k, 2] = 1
    return rst


def track_points(sf, nd, filename, gt, track_rsts):
    """The tracking part of prepareStableIndexNSwapAllModel (/root/reference/super/nodes.py:594-599) with
    update_track_pts (:251-265).  gt: {filename: (T,3)}; track_rsts: dict filled in place."""
    if (sf.track_id >= 0).count_nonzero() > 0 and filename in gt:                  # update_track_pts
        if filename not in track_rsts:
            track_rsts[filename] = init_track_pts(sf, nd, gt[filename], th=1e-2)
        else:
            for k in range(len(sf.track_id)):
                if sf.track_id[k] >= 0:
                    track_rsts[filename][k, 0:2] = sf.projdata[sf.track_id[k]]
                    track_rsts[filename][k, 2] = 1
    if (sf.track_id == -1).count_nonzero() > 0 and filename in gt:
        track_rsts[filename] = init_track_pts(sf, nd, gt[filename], th=0.2)


# ----------------------------------------------------------------------------------------------
# whole frame
# ----------------------------------------------------------------------------------------------
def default_opt(**kw):
    """The flag defaults of /root/reference/options.py that the path reads."""
    o = NS(method="super", phase="test", use_derived_gradient=True, num_optimize_iterations=10,
           num_ED_neighbors=4, num_neighbors=4, th_dist=0.1, th_cosine_ang=0.4, th_time_steps=30,
           disable_removing_unstable_surfels=False, disable_merging_new_surfels=False,
           disable_merging_exist_surfels=False, disable_adding_new_surfels=False,
           mesh_step_size=32, data="superv1", height=480, width=640, dilate_invalid_kernel=5,
           sf_point_plane=True, sf_point_plane_weight=1.0, mesh_arap=True, mesh_arap_weight=10.0,
           mesh_rot=True, mesh_rot_weight=1.0, mesh_face=False, mesh_face_weight=1.0,
           optimizer="SGD", learning_rate=5e-5)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Tracker:
    """SuPer.forward (/root/reference/super/super.py:23-83) for the LM configuration."""

    def __init__(self, opt, assemble="blocks", gt=None):
        self.opt, self.sf, self.assemble = opt, None, assemble
        self.trace = None
        self.gt, self.track_rsts = gt, {}       # gt: {"%06d": (T,3) int [x,y,valid]}  (--tracking_gt_file)

    def step(self, frame, trace=False):
        nd = preprocess(self.opt, frame)
        self.last_nd = nd
        if self.sf is None:
            graph = build_graph(self.opt, nd)
            self.sf = init_surfels(self.opt, nd, graph)
            if self.gt is not None:              # Surfels.__init__ :124 + the first prepareStable... (super.py:63)
                T = len(next(iter(self.gt.values())))
                self.sf.track_id = -torch.ones(T, dtype=torch.long)
                track_points(self.sf, nd, frame["filename"], self.gt, self.track_rsts)
            return None
        tr = [] if trace else None
        beta = lm_solve(self.opt, self.sf, nd, assemble=self.assemble, trace=tr)
        self.trace = tr
        update(self.opt, self.sf, beta)
        fuse(self.opt, self.sf, nd)
        compact(self.opt, self.sf, float(frame["time"]))
        if self.gt is not None:
            track_points(self.sf, nd, frame["filename"], self.gt, self.track_rsts)
        return beta
