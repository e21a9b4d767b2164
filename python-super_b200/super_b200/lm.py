"""Levenberg-Marquardt loop over beta in R^{J x 7} on the device: the B200 restatement of
LM_Solver.LM (/root/reference/super/LM.py:81-122).

Band path (the default, `lm_frame`): ONE C call, sb_lm_frame (csrc/lm_frame.cu), enqueues the whole loop -- per iteration
fixed-point store -> f64 band, two-sided banded Cholesky with the damping read from the device state and the step
beta += delta in its last kernel, then the J^T J pass AT the trial beta (tensor-core Gram panels; ARAP/Rot in the same
launch) whose last block takes the accept/reject decision: the trial loss falls out of that pass and, when the step is
accepted, the pass already is the next iteration's normal equations.  Bitwise reproducible (integer atomics).
Step-by-step path (`lm_solve` with on_iter hooks, the dense cross-check, systems the band solver does not take): zero
-> J^T J -> ARAP/Rot -> [damping -> library Cholesky | band solve] -> step -> loss-only pass -> accept/reject, one C call
per stage.  u, minimal_loss, the failure flag and the loss trace stay on the device (ops.LMState): no host sync.
"""
from __future__ import annotations

import ctypes

import torch

import os

from . import lib, ops
from .lib import call, ptr, stream

F64 = torch.float64


class SbLMFrame(ctypes.Structure):
    """ctypes mirror of SbLMFrame (include/super_b200.h)."""
    _P, _I, _D = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    _fields_ = [("points", _P), ("knn_idx", _P), ("knn_w", _P), ("order", _P), ("n_cap", _I), ("n_dev", _P),
                ("ed_points", _P), ("ed_knn", _P), ("J", _I),
                ("vmap", _P), ("nmap", _P), ("H", _I), ("W", _I), ("intr", _D * 4),
                ("lam_data", _D), ("lam_arap", _D), ("lam_rot", _D), ("use_arap", _I), ("use_rot", _I),
                ("iterations", _I), ("u", _D), ("v", _D), ("minimal_loss", _D),
                ("state", _P), ("beta", _P), ("best", _P),
                ("partials_loss", _P), ("n_partials_loss", _I), ("rows", _P), ("keys", _P), ("row_stride", _I),
                ("rec_vals", _P), ("rec_keys", _P), ("rec_count", _P), ("rec_cap", _I),
                ("n", _I), ("bw", _I), ("ldab", _I), ("node_pos", _P), ("pos_node", _P),
                ("fx_store", _P * 2), ("fx_shift", _I), ("fx_gshift", _I),
                ("AB", _P), ("g", _P), ("band_overflow", _P), ("dinv", _P), ("info", _P),
                ("solver_ws", _P), ("solver_ws_bytes", ctypes.c_longlong), ("n_ctas", _I),
                ("jtj_events", _P), ("n_jtj_events", _I), ("solve_events", _P), ("n_solve_events", _I),
                ("stage_events", _P), ("n_stage_events", _I), ("graph_cache", _P)]


class LMWorkspace:
    """Buffers reused across frames for a given J."""

    def __init__(self, J, device):
        n = 7 * J
        self.J, self.n = J, n
        self.device = device
        self._A = self._g = None       # dense (7J)^2 matrix: allocated only when the dense cross-check path runs
        self.beta = torch.zeros((J, 7), dtype=F64, device=device)
        self.best = torch.zeros((J, 7), dtype=F64, device=device)
        self.loss2 = torch.zeros(2, dtype=F64, device=device)
        self.state = ops.LMState(device)
        self.partials = None
        self.rows = self.keys = None        # Jacobian rows (29, stride) + node-set keys of the frame loop's evaluation pass
        self.graph_handle = ctypes.c_void_p(0)   # sb_lm_frame's instantiated CUDA graph of the frame loop (SbLMFrame.graph_cache)

    def __del__(self):
        try:
            if self.graph_handle:
                lib.load().sb_lm_graph_destroy(ctypes.byref(self.graph_handle))
        except Exception:
            pass

    @property
    def A(self):
        if self._A is None:
            self._A = torch.zeros((self.n, self.n), dtype=F64, device=self.device)
        return self._A

    @property
    def g(self):
        if self._g is None:
            self._g = torch.zeros((self.n, 1), dtype=F64, device=self.device)
        return self._g


def band_frame_ok(band, cluster_size):
    """The one-call frame loop needs the two-sided / one-sided tensor-core band solver for this system."""
    return band is not None and cluster_size >= 8 and bool(lib.load().sb_band3_fits(band.n, band.bw))


def _build_frame_struct(sf, ed, J, n_cap, n_dev, order, vmap, nmap, cam, opt, ws, band, u, v, minimal_loss, cluster_size):
    f = SbLMFrame()
    f.points, f.knn_idx, f.knn_w, f.order = ptr(sf.points), ptr(sf.knn_indices), ptr(sf.knn_w), ptr(order)
    f.n_cap, f.n_dev = n_cap, ptr(n_dev)
    f.ed_points, f.ed_knn, f.J = ptr(ed.points), ptr(ed.knn_indices), J
    f.vmap, f.nmap, f.H, f.W = ptr(vmap), ptr(nmap), cam.H, cam.W
    f.intr = (ctypes.c_double * 4)(cam.fx, cam.fy, cam.cx, cam.cy)
    f.lam_data, f.lam_arap, f.lam_rot = opt.sf_point_plane_weight, opt.mesh_arap_weight, opt.mesh_rot_weight
    f.use_arap, f.use_rot = int(bool(opt.mesh_arap)), int(bool(opt.mesh_rot))
    f.iterations, f.u, f.v, f.minimal_loss = int(opt.num_optimize_iterations), u, v, minimal_loss
    f.state, f.beta, f.best = ptr(ws.state.buf), ptr(ws.beta), ptr(ws.best)
    f.partials_loss, f.n_partials_loss = ptr(ws.partials_frame), ws.partials_frame.numel()
    f.rows, f.keys, f.row_stride = ptr(ws.rows), ptr(ws.keys), ws.keys.numel()
    f.rec_vals, f.rec_keys, f.rec_count, f.rec_cap = ptr(ws.rec_vals), ptr(ws.rec_keys), ptr(ws.rec_count), ws.rec_cap
    f.n, f.bw, f.ldab = band.n, band.bw, band.ldab
    f.node_pos, f.pos_node = ptr(band.node_pos), ptr(band.pos_node)
    f.fx_store = (ctypes.c_void_p * 2)(ptr(band.fx[0]), ptr(band.fx[1]))
    f.fx_shift, f.fx_gshift = band.fx_shift, band.fx_gshift
    f.AB, f.g = ptr(band._AB), ptr(band._g)
    f.band_overflow, f.dinv, f.info = ptr(band.overflow), ptr(band.dinv), ptr(band.info)
    f.solver_ws, f.solver_ws_bytes, f.n_ctas = ptr(band.ws4), band.ws4.numel(), int(cluster_size)
    return f


def lm_frame(sf, maps, cam, opt, ws, band, u=10.0, v=7.5, minimal_loss=1e10, order=None, n_dev=None, cluster_size=148,
             jtj_events=None, solve_events=None, row_capacity=None, stage_events=None):
    """The whole LM loop of one frame in one C call (sb_lm_frame).  Same arguments and result as lm_solve.
    row_capacity: rows to size the Jacobian-row scratch for (the tracker passes its surfel capacity, so that the buffer
    is allocated once per sequence).
    jtj_events: optional list of raw cudaEvent_t handles (2 per J^T J pass: begin, end) recorded around those launches."""
    ed = sf.ED
    J = ed.points.shape[0]
    dev = sf.points.device
    l = lib.load()
    n_cap = sf.points.shape[0]
    nb = max(1024, ops.data_loss_blocks(n_cap))
    if getattr(ws, "partials_frame", None) is None or ws.partials_frame.numel() != nb:
        ws.partials_frame = torch.zeros(nb, dtype=F64, device=dev)
    if ws.keys is None or ws.keys.numel() < n_cap:
        stride = max(n_cap, int(row_capacity or 0))
        stride = (stride + 31) // 32 * 32
        ws.rows = torch.empty((29, stride), dtype=F64, device=dev)
        ws.keys = torch.empty(stride, dtype=torch.int64, device=dev)
        ws.rec_cap = 2 * (stride // 32) + 4096
        ws.rec_vals = torch.empty((ws.rec_cap, 436), dtype=F64, device=dev)
        ws.rec_keys = torch.empty(ws.rec_cap, dtype=torch.int64, device=dev)
        ws.rec_count = torch.zeros(1, dtype=torch.int32, device=dev)
    if getattr(band, "ws4", None) is None:
        band.ws4 = torch.zeros(int(l.sb_band4_workspace_bytes(band.n, band.bw, band.ldab)), dtype=torch.uint8, device=dev)
    vmap, nmap = maps
    # the argument block only depends on persistent buffers and a few scalars: built once per distinct combination (the two
    # input buffers alternate), reused afterwards -- the host has just waited for the previous frame when it gets here, so
    # everything in front of the launch is time the device spends idle
    ck = (ptr(sf.points), ptr(sf.knn_indices), ptr(sf.knn_w), ptr(order), n_cap, ptr(n_dev), ptr(ed.points), ptr(ed.knn_indices),
          ptr(vmap), ptr(nmap), id(band), ptr(band.ws4), float(u), float(v), float(minimal_loss), int(cluster_size),
          ptr(ws.rows), ptr(ws.partials_frame), cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy,
          float(opt.sf_point_plane_weight), float(opt.mesh_arap_weight), float(opt.mesh_rot_weight), bool(opt.mesh_arap),
          bool(opt.mesh_rot), int(opt.num_optimize_iterations), ptr(ws.beta), ptr(ws.state.buf))
    cache = ws.__dict__.setdefault("frame_structs", {})
    f = cache.get(ck) if not (jtj_events or solve_events or stage_events) else None
    if f is None:
        f = _build_frame_struct(sf, ed, J, n_cap, n_dev, order, vmap, nmap, cam, opt, ws, band, u, v, minimal_loss, cluster_size)
        if not (jtj_events or solve_events or stage_events):
            if len(cache) > 16:
                cache.clear()
            cache[ck] = f
    keep = []
    for name, evs in (("jtj", jtj_events), ("solve", solve_events), ("stage", stage_events)):
        if evs:
            arr = (ctypes.c_void_p * len(evs))(*evs)
            keep.append(arr)
            setattr(f, name + "_events", ctypes.cast(arr, ctypes.c_void_p))
            setattr(f, "n_" + name + "_events", len(evs))
    if not keep and os.environ.get("SB_LM_GRAPH", "1") != "0":
        f.graph_cache = ctypes.cast(ctypes.pointer(ws.graph_handle), ctypes.c_void_p)
    call("sb_lm_frame", ctypes.byref(f), stream())
    lib.LAUNCHES += 2 + 8 * f.iterations        # lm_begin, eval, Gram, scatter; per iteration 5 (solve) + eval + Gram + scatter | loss
    band._dirty = False
    return ws.beta, ws



def lm_solve(sf, maps, cam, opt, ws=None, u=10.0, v=7.5, minimal_loss=1e10, order=None, n_dev=None,
             on_iter=None, band=None, cluster_size=16, jtj_events=None, solve_events=None, row_capacity=None,
             stage_events=None):
    """sf: object with points (N,3) f64, knn_indices (N,4) i32, knn_w (N,4) f64 and ED (points, knn_indices i32).
    maps: (vmap, nmap) dense float4 images of the new frame.  Returns beta (J,7) f64 (a view of the
    workspace) -- the same value LM_Solver.LM returns.
    band: ops.Band -> normal equations assembled straight into band storage and solved by sb_band_solve;
    None -> dense A + the library Cholesky (torch.linalg / cuSOLVER), kept as the cross-check path."""
    ed = sf.ED
    J = ed.points.shape[0]
    dev = sf.points.device
    if ws is None or ws.J != J:
        ws = LMWorkspace(J, dev)
    if (on_iter is None and band_frame_ok(band, cluster_size) and bool(opt.sf_point_plane)
            and os.environ.get("SB_LM_STEPWISE", "0") != "1"):
        return lm_frame(sf, maps, cam, opt, ws, band, u, v, minimal_loss, order, n_dev, cluster_size, jtj_events,
                        solve_events, row_capacity, stage_events)
    n_cap = sf.points.shape[0]
    nb = ops.data_loss_blocks(n_cap)
    if ws.partials is None or ws.partials.numel() != nb:
        ws.partials = torch.zeros(nb, dtype=F64, device=dev)
    vmap, nmap = maps
    use_data, use_arap, use_rot = bool(opt.sf_point_plane), bool(opt.mesh_arap), bool(opt.mesh_rot)
    lam_d, lam_a, lam_r = opt.sf_point_plane_weight, opt.mesh_arap_weight, opt.mesh_rot_weight
    if order is None and use_data:
        order = ops.tuple_order(sf.knn_indices, n_dev)
    ops.lm_begin(ws.state, ws.beta, ws.best, u, v, minimal_loss)
    ws.loss2.zero_()
    if band is not None:
        band.info.zero_()
    for it in range(opt.num_optimize_iterations):
        if band is not None:
            band.store.zero_()               # AB and g in one memset (fixed-point store)
        else:
            ws.A.zero_()
            ws.g.zero_()
        if use_data:
            ops.data_term_jtj(sf.points, sf.knn_indices, sf.knn_w, order, ed.points, ws.beta, vmap, nmap, cam,
                              lam_d, ws.A if band is None else None, ws.g if band is None else None, n_dev=n_dev,
                              band=band)
        if use_arap or use_rot:
            ops.reg_terms(ed.points, ed.knn_indices, ws.beta, lam_a, lam_r, use_arap, use_rot,
                          ws.A if band is None else None, ws.g if band is None else None, band=band)
        if on_iter is not None:
            on_iter(it, "normal_equations", ws)
        if band is not None:
            # own banded Cholesky on one thread-block cluster; damping u is read from the device state
            delta = band.g
            if not ops.band_solve_step(band, ws.state, ws.beta, cluster_size):       # step folded into the solve
                ops.band_solve(band, ws.state.buf.data_ptr(), cluster_size)
                ops.lm_step(ws.state, band.info, ws.beta, delta, band.node_pos)
        else:
            ops.lm_damp(ws.state, ws.A)
            L, info = torch.linalg.cholesky_ex(ws.A, check_errors=False)      # reads the lower triangle
            delta = torch.cholesky_solve(ws.g, L)
            ops.lm_step(ws.state, info, ws.beta, delta)
        if on_iter is not None:
            on_iter(it, "step", ws, delta)
        if use_data:     # loss-only pass; its last block runs the accept/reject step (one launch)
            ops.data_term_loss_decide(sf.points, sf.knn_indices, sf.knn_w, ed.points, ed.knn_indices, ws.beta, ws.best,
                                      vmap, nmap, cam, lam_d, lam_a, lam_r, use_arap, use_rot, ws.partials, ws.state,
                                      n_dev=n_dev)
            continue
        ws.partials.zero_()
        if use_arap or use_rot:      # regularisers' losses inside the decide launch
            ops.lm_decide_reg(ws.state, ws.partials, ed.points, ed.knn_indices, lam_a, lam_r, use_arap, use_rot,
                              ws.beta, ws.best)
        else:
            ops.lm_decide(ws.state, ws.partials, ws.loss2, ws.beta, ws.best)
    return ws.beta, ws
