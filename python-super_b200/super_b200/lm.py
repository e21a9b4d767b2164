"""Levenberg-Marquardt loop over beta in R^{J x 7} on the device: the B200 restatement of
LM_Solver.LM (/root/reference/super/LM.py:81-122).

Per iteration, band path (the default): data-term J^T J kernel (tensor-core Gram panels), with the ARAP/Rot kernel
and the memset of the next iteration's band buffer on a side stream under it -> two-sided banded Cholesky with the
damping read from the device state and the step beta += delta in its last kernel -> loss-only pass whose last block
runs the accept/reject step.  Dense path (cross-check): zero A,g -> J^T J -> ARAP/Rot -> damping -> library Cholesky ->
step -> loss-only pass -> accept/reject.  u, minimal_loss, the failure flag and the loss trace stay on the device
(ops.LMState): the loop issues no host sync.
"""
from __future__ import annotations

import torch

from . import ops

F64 = torch.float64


class LMWorkspace:
    """Buffers reused across frames for a given J."""

    def __init__(self, J, device):
        n = 7 * J
        self.J, self.n = J, n
        self.A = torch.zeros((n, n), dtype=F64, device=device)
        self.g = torch.zeros((n, 1), dtype=F64, device=device)
        self.beta = torch.zeros((J, 7), dtype=F64, device=device)
        self.best = torch.zeros((J, 7), dtype=F64, device=device)
        self.loss2 = torch.zeros(2, dtype=F64, device=device)
        self.state = ops.LMState(device)
        self.partials = None
        # the regularisers' normal-equation terms (6 CTAs, ~7 us) run on a side stream under the data term's pass
        self.side = torch.cuda.Stream(device=device)
        self.fork, self.join, self.cleared = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()


def lm_solve(sf, maps, cam, opt, ws=None, u=10.0, v=7.5, minimal_loss=1e10, order=None, n_dev=None,
             on_iter=None, band=None, cluster_size=16):
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
    overlap = use_data and (use_arap or use_rot) and on_iter is None
    cleared_ahead = False
    for it in range(opt.num_optimize_iterations):
        if band is not None:
            if cleared_ahead:                # the side stream cleared the other store during the previous iteration
                band.flip()
                torch.cuda.current_stream().wait_event(ws.cleared)
            else:
                band.store.zero_()           # AB and g in one memset
        else:
            ws.A.zero_()
            ws.g.zero_()
        if overlap:      # fork: both kernels only add into the zeroed A, g
            main = torch.cuda.current_stream()
            ws.fork.record(main)
            ws.side.wait_event(ws.fork)
            with torch.cuda.stream(ws.side):
                ops.reg_terms(ed.points, ed.knn_indices, ws.beta, lam_a, lam_r, use_arap, use_rot, ws.A, ws.g, band=band)
                ws.join.record(ws.side)
                if band is not None:         # last read by the previous iteration's solve, which the fork is behind
                    band.other_store.zero_()
                    ws.cleared.record(ws.side)
                    cleared_ahead = True
        if use_data:
            ops.data_term_jtj(sf.points, sf.knn_indices, sf.knn_w, order, ed.points, ws.beta, vmap, nmap, cam,
                              lam_d, ws.A, ws.g, n_dev=n_dev, band=band)
        if overlap:
            main.wait_event(ws.join)
        elif use_arap or use_rot:
            ops.reg_terms(ed.points, ed.knn_indices, ws.beta, lam_a, lam_r, use_arap, use_rot, ws.A, ws.g, band=band)
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
