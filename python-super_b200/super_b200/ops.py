"""Thin Python wrappers over the C ABI (one function per entry point of include/super_b200.h).

Tensors are torch CUDA tensors used as device-memory handles; all arithmetic happens in the CUDA
library.  Every wrapper raises if the library is missing (super_b200.lib.load): no fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import lib
from .lib import call, ptr, stream

F64, I32 = torch.float64, torch.int32


def _dev(t):
    if not t.is_cuda:
        raise lib.SuperB200Error("super_b200 ops need CUDA tensors (no CPU path)")
    return t


def as_i32(idx):
    return idx if idx.dtype == I32 else idx.to(I32)


# ---- kNN / weights / warp ----------------------------------------------------------------------
def knn(query, ref, K, n_dev=None, qseg=None, rseg=None):
    """find_knn (/root/reference/utils/utils.py:212-242): (sqrt dists (N,K) f64, idx (N,K) i32).  qseg/rseg (i32):
    the per-class search of --hard_seg."""
    query, ref = _dev(query).contiguous(), _dev(ref).contiguous()
    n, dim = query.shape
    dist = torch.empty((n, K), dtype=F64, device=query.device)
    idx = torch.empty((n, K), dtype=I32, device=query.device)
    call("sb_knn_class", ptr(query), n, ptr(n_dev), ptr(qseg), ptr(ref), ref.shape[0], ptr(rseg), dim, K, ptr(dist),
         ptr(idx), stream())
    return dist, idx


def knn_weights(dist, idx, radii, radius_mode=0, stable=None, n_dev=None):
    """softmax(exp(-d/r)) (+ clears stable[i] when no node is within its radius)."""
    n = dist.shape[0]
    w = torch.empty((n, 4), dtype=F64, device=dist.device)
    call("sb_knn_weights", ptr(dist), ptr(idx), n, ptr(n_dev), ptr(radii), radius_mode, ptr(w), ptr(stable),
         stream())
    return w


def reweight(points, idx, ed_points, radii, w_out=None, n_dev=None):
    n = points.shape[0]
    if w_out is None:
        w_out = torch.empty((n, 4), dtype=F64, device=points.device)
    call("sb_reweight", ptr(points), ptr(idx), n, ptr(n_dev), ptr(ed_points), ptr(radii), ptr(w_out), stream())
    return w_out


def warp_update(points, norms, idx, w, ed_points, ed_norms, beta, n_dev=None):
    """Surfels.update in place (/root/reference/super/nodes.py:193-223, LM form)."""
    call("sb_warp_update", ptr(points), ptr(norms), ptr(idx), ptr(w), points.shape[0], ptr(n_dev),
         ptr(ed_points), ptr(ed_norms), ptr(beta), ed_points.shape[0], stream())


# ---- new-frame maps ------------------------------------------------------------------------------
def dense_maps(nd_points, nd_norms, valid, height, width):
    """Compact (Nv,3) f64 points/normals + valid (P,) -> dense float4 maps (P,4) f32 [x,y,z,valid].
    The values are float32-exact in the reference (data_loader.py:453-454 casts f32 results up)."""
    P = height * width
    vmap = torch.zeros((P, 4), dtype=torch.float32, device=nd_points.device)
    nmap = torch.zeros((P, 4), dtype=torch.float32, device=nd_points.device)
    vmap[valid, :3] = nd_points.to(torch.float32)
    vmap[valid, 3] = 1.0
    nmap[valid, :3] = nd_norms.to(torch.float32)
    return vmap, nmap


# ---- LM data term --------------------------------------------------------------------------------
class Camera:
    def __init__(self, fx, fy, cx, cy, height, width):
        self.fx, self.fy, self.cx, self.cy, self.H, self.W = float(fx), float(fy), float(cx), float(cy), height, width
        self.c = lib.intr_array(fx, fy, cx, cy)

    @staticmethod
    def from_K(K, height, width):
        """K: (4,4) or (1,4,4) float32 tensor/array (values promoted to f64 exactly, like the reference)."""
        K = K.reshape(-1, 4, 4)[0]
        return Camera(float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]), height, width)


def data_term_rows(points, knn_idx, knn_w, ed_points, beta, vmap, nmap, cam, lam, want_jrow=True, n_dev=None):
    n = points.shape[0]
    dev = points.device
    matched = torch.zeros(n, dtype=torch.uint8, device=dev)
    corners = torch.zeros((n, 4), dtype=I32, device=dev)
    r = torch.zeros(n, dtype=F64, device=dev)
    jrow = torch.zeros((n, 28), dtype=F64, device=dev) if want_jrow else None
    call("sb_data_term_rows", ptr(points), ptr(knn_idx), ptr(knn_w), n, ptr(n_dev), ptr(ed_points), ptr(beta),
         ed_points.shape[0], ptr(vmap), ptr(nmap), cam.H, cam.W, cam.c, float(lam), ptr(matched), ptr(corners),
         ptr(r), ptr(jrow), stream())
    return matched.bool(), corners, r, jrow


class Band:
    """Band-storage target of the normal equations in the node order `node_pos` (node id -> position), scalar
    half-bandwidth bw.

    Two representations:
      * `fx`: two int64 FIXED-POINT stores (AB | g; entries are multiples of 2^-fx_shift, 2^-fx_gshift) that the assembly
        kernels add into with integer atomics -- order-independent, hence bitwise reproducible;
      * `AB` (n, ldab) lower band row-major + `g` (n): the f64 work copy the solver factors in place (`finalize()` =
        sb_band_from_fixed).  Reading `AB` / `g` / `to_dense()` after an assembly finalizes first.
    `overflow`: bit 0 an entry fell outside the band, bit 1 an addend fell outside the fixed-point range."""

    FX_SHIFT, FX_GSHIFT = 40, 48

    def __init__(self, n, bw, node_pos, device, fx_shift=None, fx_gshift=None):
        self.n, self.bw = int(n), int(bw)
        self.ldab = self.bw + 1
        self.fx_shift = self.FX_SHIFT if fx_shift is None else int(fx_shift)
        self.fx_gshift = self.FX_GSHIFT if fx_gshift is None else int(fx_gshift)
        m = self.n * self.ldab + self.n
        self.fx = [torch.zeros(m, dtype=torch.int64, device=device) for _ in range(2)]
        self.work = torch.zeros(m, dtype=F64, device=device)
        self._AB = self.work[: self.n * self.ldab].view(self.n, self.ldab)
        self._g = self.work[self.n * self.ldab:]
        self._dirty = False
        self.node_pos = node_pos
        # solver position -> node id (the step folded into the solve scatters x back to beta's node order)
        self.pos_node = None if node_pos is None else torch.argsort(node_pos.long()).to(I32).contiguous()
        self.overflow = torch.zeros(1, dtype=I32, device=device)
        self.dinv = torch.zeros(self.n, dtype=F64, device=device)
        self.info = torch.zeros(1, dtype=I32, device=device)

    @property
    def store(self):
        """The fixed-point store the stand-alone assembly entry points add into (zero it before a new assembly)."""
        self._dirty = True
        return self.fx[0]

    def finalize(self):
        call("sb_band_from_fixed", ptr(self.fx[0]), self.n, self.ldab, self.fx_shift, self.fx_gshift, ptr(self._AB),
             ptr(self._g), stream())
        self._dirty = False

    @property
    def AB(self):
        if self._dirty:
            self.finalize()
        return self._AB

    @property
    def g(self):
        if self._dirty:
            self.finalize()
        return self._g

    def to_dense(self):
        """Symmetric dense matrix in the ORIGINAL node order (tests)."""
        n, bw = self.n, self.bw
        AB = self.AB.cpu()
        A = torch.zeros((n, n), dtype=F64)
        for d in range(bw + 1):                       # diagonal offset i - j = bw - d
            off = bw - d
            if off < n:
                A.diagonal(-off).copy_(AB[off:, d])
        A = A + torch.tril(A, -1).t()
        if self.node_pos is not None:
            pos = self.node_pos.cpu().long()
            sidx = (7 * pos[:, None] + torch.arange(7)[None, :]).reshape(-1)   # original scalar -> permuted
            A = A[sidx][:, sidx]
        return A


def _target(A, g, band):
    """(A, lda, bw, node_pos, overflow, g, fx_shift, fx_gshift) of an assembly call: dense f64 or the band's store."""
    if band is not None:
        st = band.store
        nab = band.n * band.ldab
        return st[:nab], band.ldab, band.bw, band.node_pos, band.overflow, st[nab:], band.fx_shift, band.fx_gshift
    return A, (A.stride(0) if A is not None else 0), -1, None, None, g, -1, -1


def data_term_jtj(points, knn_idx, knn_w, order, ed_points, beta, vmap, nmap, cam, lam, A, g, loss_cur=None,
                  n_dev=None, band=None):
    A, lda, bw, pos, ovf, g, sh, gsh = _target(A, g, band)
    call("sb_data_term_jtj", ptr(points), ptr(knn_idx), ptr(knn_w), ptr(order), points.shape[0], ptr(n_dev),
         ptr(ed_points), ptr(beta), ed_points.shape[0], ptr(vmap), ptr(nmap), cam.H, cam.W, cam.c, float(lam),
         ptr(A), lda, bw, ptr(pos), ptr(ovf), ptr(g), ptr(loss_cur), sh, gsh, stream())


def data_loss_blocks(n_cap):
    return lib.load().sb_data_loss_blocks(int(n_cap))


def data_term_loss(points, knn_idx, knn_w, ed_points, beta, vmap, nmap, cam, lam, partials, n_dev=None):
    call("sb_data_term_loss", ptr(points), ptr(knn_idx), ptr(knn_w), points.shape[0], ptr(n_dev), ptr(ed_points),
         ptr(beta), ed_points.shape[0], ptr(vmap), ptr(nmap), cam.H, cam.W, cam.c, float(lam), ptr(partials),
         partials.numel(), stream())


def data_term_loss_decide(points, knn_idx, knn_w, ed_points, ed_knn, beta, best, vmap, nmap, cam, lam, lam_arap, lam_rot,
                          use_arap, use_rot, partials, state, n_dev=None):
    """Loss-only pass + accept/reject (regularisers' losses included) in one launch."""
    call("sb_data_term_loss_decide", ptr(points), ptr(knn_idx), ptr(knn_w), points.shape[0], ptr(n_dev), ptr(ed_points),
         ptr(beta), ed_points.shape[0], ptr(vmap), ptr(nmap), cam.H, cam.W, cam.c, float(lam), ptr(partials),
         partials.numel(), ptr(state.buf), ptr(ed_knn), float(lam_arap), float(lam_rot), int(use_arap), int(use_rot),
         ptr(beta), ptr(best), stream())


_ORDER_WS = {}


def tuple_order(knn_idx, n_dev=None, node_pos=None, block_bw=None, J=None):
    """Surfel ids sorted by their (ordered) 4-tuple of ED nodes: the visiting order of the J^T J
    kernel, so that a warp's 32 surfels share their node blocks.  Rows beyond *n_dev sort last.
    block_bw (1,) i32: atomically max-ed with the node-block half-bandwidth the tuples need.
    J (number of ED nodes) bounds the key width: 4*ceil(log2(J+1)) bits are sorted (sb_tuple_order)."""
    n = knn_idx.shape[0]
    dev = knn_idx.device
    if J is None:
        J = int(node_pos.shape[0]) if node_pos is not None else 65535
    ws = _ORDER_WS.get(dev)
    if ws is None or ws[0] < n:
        cap = max(n, 1)
        tb = int(lib.load().sb_tuple_order_temp_bytes(cap))
        ws = (cap, torch.empty(cap, dtype=torch.int64, device=dev), torch.empty(cap, dtype=torch.int64, device=dev),
              torch.empty(cap, dtype=I32, device=dev), torch.empty(max(tb, 8), dtype=torch.uint8, device=dev), tb)
        _ORDER_WS[dev] = ws
    order = torch.empty(n, dtype=I32, device=dev)
    call("sb_tuple_order", ptr(knn_idx), n, ptr(n_dev), int(J), ptr(node_pos), ptr(block_bw), ptr(ws[1]), ptr(ws[2]), ptr(ws[3]),
         ptr(order), ptr(ws[4]), ws[5], stream())
    return order


# ---- LM regularisers / controller ----------------------------------------------------------------
def reg_terms(ed_points, ed_knn, beta, lam_arap, lam_rot, use_arap, use_rot, A=None, g=None, loss2=None, band=None):
    A, lda, bw, pos, ovf, g, sh, gsh = _target(A, g, band)
    call("sb_reg_terms", ptr(ed_points), ptr(ed_knn), ptr(beta), ed_points.shape[0], float(lam_arap), float(lam_rot),
         int(use_arap), int(use_rot), ptr(A), lda, bw, ptr(pos), ptr(ovf), ptr(g), ptr(loss2), sh, gsh, stream())


def band_solve(band, u_ptr=None, cluster_size=16, variant=None):
    """(A + u I) x = g in place: band.g <- x (solver node order).  u_ptr: device address of u.
    variant 4 (default): sb_band_solve4 (two-sided: variant 3's factorisation from both ends at once, middle block last);
    variant 3: sb_band_solve3 (one-sided; DMMA products with explicit block inverses, push-style back substitution)."""
    import os
    if variant is None:
        variant = int(os.environ.get("SB_BAND_VARIANT", "4"))
    l = lib.load()
    if cluster_size < 3 or not l.sb_band3_fits(band.n, band.bw):
        raise lib.SuperB200Error(f"band solver does not take n={band.n}, bw={band.bw} on {cluster_size} CTAs")
    if variant == 4 and cluster_size >= 8:
        if getattr(band, "ws4", None) is None:
            band.ws4 = torch.zeros(int(l.sb_band4_workspace_bytes(band.n, band.bw, band.ldab)), dtype=torch.uint8,
                                   device=band.AB.device)
        call("sb_band_solve4", ptr(band.AB), band.ldab, band.n, band.bw, ptr(band.g), u_ptr, ptr(band.dinv),
             ptr(band.info), ptr(band.ws4), band.ws4.numel(), int(cluster_size), stream())
    else:
        if getattr(band, "ws3", None) is None:
            band.ws3 = torch.zeros(int(l.sb_band3_workspace_bytes(band.n, band.bw)), dtype=torch.uint8,
                                   device=band.AB.device)
        call("sb_band_solve3", ptr(band.AB), band.ldab, band.n, band.bw, ptr(band.g), u_ptr, ptr(band.dinv),
             ptr(band.info), ptr(band.ws3), band.ws3.numel(), int(cluster_size), stream())


def band_solve_step(band, state, beta, cluster_size=148):
    """band_solve (two-sided variant) + lm_step in one call: beta += x unless the factorisation failed; the step rides in
    the solve's last kernel.  Returns False if this system/variant has no folded path (caller runs the two calls)."""
    import os
    l = lib.load()
    if int(os.environ.get("SB_BAND_VARIANT", "4")) != 4 or cluster_size < 8 or not l.sb_band3_fits(band.n, band.bw):
        return False
    if getattr(band, "ws4", None) is None:
        band.ws4 = torch.zeros(int(l.sb_band4_workspace_bytes(band.n, band.bw, band.ldab)), dtype=torch.uint8,
                               device=band.AB.device)
    call("sb_band_solve4_step", ptr(band.AB), band.ldab, band.n, band.bw, ptr(band.g), state.buf.data_ptr(),
         ptr(band.dinv), ptr(band.info), ptr(band.ws4), band.ws4.numel(), int(cluster_size),
         state.buf.data_ptr() + state.failed_offset, ptr(beta), ptr(band.pos_node), stream())
    return True


class LMState:
    """Device-resident controller (u, minimal_loss, trace) -- decoded on demand, never during the loop."""

    def __init__(self, device):
        l = lib.load()
        self.nbytes = l.sb_lm_state_bytes()
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        off = (ctypes.c_int * 8)()
        l.sb_lm_state_offsets(off)
        self.off = list(off)
        self.failed_offset = self.off[3]

    def read(self):
        """D2H copy + decode (synchronises): dict(u, minimal_loss, iter, failed, loss, loss_terms, accept, u_trace)."""
        import numpy as np
        h = self.buf.cpu().numpy()
        o = self.off
        it = int(h[o[2]:o[2] + 4].view(np.int32)[0])
        n = min(it, 64)
        return {
            "u": float(h[o[0]:o[0] + 8].view(np.float64)[0]),
            "minimal_loss": float(h[o[1]:o[1] + 8].view(np.float64)[0]),
            "iter": it,
            "failed": int(h[o[3]:o[3] + 4].view(np.int32)[0]),
            "loss": h[o[4]:o[4] + 8 * 64].view(np.float64)[:n].copy(),
            "loss_terms": h[o[5]:o[5] + 8 * 192].view(np.float64).reshape(64, 3)[:n].copy(),
            "accept": h[o[6]:o[6] + 4 * 64].view(np.int32)[:n].copy(),
            "u_trace": h[o[7]:o[7] + 8 * 64].view(np.float64)[:n].copy(),
        }


def lm_begin(state, beta, best, u=10.0, v=7.5, minimal_loss=1e10):
    call("sb_lm_begin", ptr(state.buf), ptr(beta), ptr(best), beta.shape[0], float(u), float(v), float(minimal_loss),
         stream())


def lm_damp(state, A):
    call("sb_lm_damp", ptr(state.buf), ptr(A), A.stride(0), A.shape[0], stream())


def lm_step(state, info, beta, delta, node_pos=None):
    call("sb_lm_step", ptr(state.buf), ptr(info), ptr(beta), ptr(delta), beta.numel(), ptr(node_pos), stream())


def lm_decide(state, partials, loss2, beta, best):
    call("sb_lm_decide", ptr(state.buf), ptr(partials), partials.numel(), ptr(loss2), ptr(beta), ptr(best),
         beta.numel(), stream())


def lm_decide_reg(state, partials, ed_points, ed_knn, lam_arap, lam_rot, use_arap, use_rot, beta, best):
    """lm_decide with the regularisers' losses evaluated inside the same launch."""
    call("sb_lm_decide_reg", ptr(state.buf), ptr(partials), partials.numel(), ptr(ed_points), ptr(ed_knn),
         ed_points.shape[0], float(lam_arap), float(lam_rot), int(use_arap), int(use_rot), ptr(beta), ptr(best),
         beta.numel(), stream())
