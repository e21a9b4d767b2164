"""Device-resident tracker state and the per-frame pipeline (host side, above the C ABI).

This is the B200-native equivalent of SuPer.forward / SuPer.fusion
(/root/reference/super/super.py:23-83): preprocess -> [init | LM -> update -> fuse -> compact],
with every stage a handful of CUDA kernels from libsuper_b200.so.  Row counts live on the device; the
host waits ONCE per tracked frame, for the pinned-memory copy of the row count and of the band width the
next normal equations need (Tracker._publish_count / _refresh_bound).

Layouts are the reference's in capacity-sized buffers; the drop-in classes in
super_b200/super/ expose exact-size views of them.
"""
from __future__ import annotations

import contextlib
import ctypes
import os
from types import SimpleNamespace as NS

import torch

from . import lib, ops, lm, graphfit
from .lib import call, ptr, stream

F64, F32, I32, I64, U8 = torch.float64, torch.float32, torch.int32, torch.int64, torch.uint8


# ---- ctypes mirrors of the structs in include/super_b200.h -------------------------------------------
class SbSurfels(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("points", "norms", "colors", "confs", "radii", "time_stamp",
                                               "knn_idx", "knn_w", "projdata", "stable")] + \
               [("cap", ctypes.c_int), ("n_dev", ctypes.c_void_p), ("seg", ctypes.c_void_p),
                ("seg_conf", ctypes.c_void_p), ("n_classes", ctypes.c_int)]


class SbFrame(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("vmap", "nmap", "radii", "confs", "color")] + \
               [("H", ctypes.c_int), ("W", ctypes.c_int)] + [(n, ctypes.c_double) for n in ("fx", "fy", "cx", "cy")] + \
               [("seg", ctypes.c_void_p), ("seg_conf", ctypes.c_void_p)]


class SbFuseParams(ctypes.Structure):
    _fields_ = [("th_dist", ctypes.c_double), ("th_cos", ctypes.c_double), ("time_now", ctypes.c_float),
                ("disable_merging_new", ctypes.c_int), ("disable_merging_exist", ctypes.c_int),
                ("disable_adding_new", ctypes.c_int), ("class_gate", ctypes.c_int), ("semantic_weights", ctypes.c_int),
                ("ed_seg_conf", ctypes.c_void_p), ("ed_seg", ctypes.c_void_p)]


SURFEL_FIELDS = (("points", 3, F64), ("norms", 3, F64), ("colors", 3, F32), ("confs", 0, F32), ("radii", 0, F64),
                 ("time_stamp", 0, F32), ("knn_idx", 4, I32), ("knn_w", 4, F64), ("projdata", 2, F32),
                 ("stable", 0, U8))


class SurfelBuffers:
    """One capacity-sized set of surfel arrays + its device row counter."""

    def __init__(self, cap, device, n_classes=0):
        self.cap = int(cap)
        for name, width, dt in SURFEL_FIELDS:
            shape = (self.cap, width) if width else (self.cap,)
            setattr(self, name, torch.zeros(shape, dtype=dt, device=device))
        self.n_dev = torch.zeros(1, dtype=I32, device=device)
        self.n_classes = int(n_classes)
        self.seg = torch.zeros(self.cap, dtype=I32, device=device) if n_classes else None          # Semantic-SuPer state
        self.seg_conf = torch.zeros((self.cap, n_classes), dtype=F64, device=device) if n_classes else None
        self.c = SbSurfels(*[ptr(getattr(self, n)) for n, _, _ in SURFEL_FIELDS], self.cap, ptr(self.n_dev),
                           ptr(self.seg), ptr(self.seg_conf), self.n_classes)

    def ref(self):
        return ctypes.byref(self.c)


class Frame:
    """One preprocessed input frame as dense per-pixel images (the SbFrame of the C ABI)."""

    def __init__(self, H, W, device):
        P = H * W
        self.H, self.W, self.P = H, W, P
        self.vmap = torch.zeros((P, 4), dtype=F32, device=device)
        self.nmap = torch.zeros((P, 4), dtype=F32, device=device)
        self.radii = torch.zeros(P, dtype=F64, device=device)
        self.confs = torch.zeros(P, dtype=F32, device=device)
        self.valid_i32 = torch.zeros(P, dtype=I32, device=device)
        self.pcd = torch.zeros((P, 4), dtype=F32, device=device)
        self.color = None
        self.cam = None
        self.time = 0.0
        self.c = None
        self.ED_nodes = None     # set by SuPer.forward on the first frame (the reference's sfdata.ED_nodes)
        self.seg = None          # (P,) i32, (P,C) f64, (C,H,W) f64 scores: semantic inputs (preprocess fills them)
        self.seg_conf = None
        self.scores = None

    def bind(self, color, cam, time):
        self.color, self.cam, self.time = color, cam, float(time)
        self.c = SbFrame(ptr(self.vmap), ptr(self.nmap), ptr(self.radii), ptr(self.confs), ptr(color), self.H,
                         self.W, cam.fx, cam.fy, cam.cx, cam.cy, ptr(self.seg), ptr(self.seg_conf))

    def ref(self):
        return ctypes.byref(self.c)

    @property
    def valid(self):
        return self.vmap[:, 3] != 0


def preprocess(opt, depth, color, K, inv_K, time, frame=None, inval=None, divterm=1.0 / (2.0 * 0.6 * 0.6),
               seg_scores=None):
    """depth_preprocessing (/root/reference/utils/data_loader.py:333-523) on the device.
    depth (H,W) f32, color (3,H,W) f32 CUDA tensors; K, inv_K (4,4) f32 host or device tensors."""
    H, W = opt.height, opt.width
    dev = depth.device
    if frame is None:
        frame = Frame(H, W, dev)
    Kh = K.detach().cpu().reshape(-1, 4, 4)[0] if torch.is_tensor(K) else torch.as_tensor(K).reshape(-1, 4, 4)[0]
    iKh = inv_K.detach().cpu().reshape(-1, 4, 4)[0] if torch.is_tensor(inv_K) else torch.as_tensor(inv_K).reshape(-1, 4, 4)[0]
    ik = (ctypes.c_float * 9)(*[float(iKh[i, j]) for i in range(3) for j in range(3)])
    cam = ops.Camera.from_K(Kh, H, W)
    depth = depth.reshape(H, W).contiguous()
    color = color.reshape(3, H, W).contiguous()
    call("sb_preprocess", ptr(depth), ptr(color), ptr(inval), ik, float(Kh[0, 0]), float(divterm),
         1 if opt.data == "superv2" else 0, H, W, ptr(frame.pcd), ptr(frame.vmap), ptr(frame.nmap), ptr(frame.radii),
         ptr(frame.confs), ptr(frame.valid_i32), stream())
    if seg_scores is not None:           # inputs[("seg_conf", 0)]: (C,H,W) f64 class scores
        sc = seg_scores.reshape(-1, H, W).to(F64).contiguous()
        C = sc.shape[0]
        if frame.seg is None or frame.seg_conf.shape[1] != C:
            frame.seg = torch.zeros(frame.P, dtype=I32, device=dev)
            frame.seg_conf = torch.zeros((frame.P, C), dtype=F64, device=dev)
        frame.scores = sc
        call("sb_seg_maps", ptr(sc), C, H, W, ptr(frame.seg), ptr(frame.seg_conf), stream())
    else:
        frame.seg = frame.seg_conf = frame.scores = None
    frame.bind(color, cam, time)
    return frame


def extra_invalid_mask(opt, depth, seg=None, mask=None):
    """The morphology part of the reference's invalid mask (data_loader.py:374-397): only needed when a
    valid mask is loaded or classes are deleted; otherwise the open/dilate of an all-false mask is a no-op.
    mask: (H,W) bool VALID mask (--load_valid_mask), seg: (H,W) integer labels (--del_seg_classes).  Two launches of
    sb_dilate_box: inval = ~dilate(~inval, k); inval = dilate(inval, 2k)  (:394-397)."""
    if mask is None and not getattr(opt, "del_seg_classes", []):
        return None
    H, W = opt.height, opt.width
    inval = torch.zeros((H, W), dtype=torch.bool, device=depth.device) if mask is None else ~mask.reshape(H, W).bool()
    for c in getattr(opt, "del_seg_classes", []):
        inval |= seg.reshape(H, W) == c
    inval = inval.to(U8).contiguous()
    if opt.data == "superv1" and opt.dilate_invalid_kernel > 0:
        tmp, out = torch.empty_like(inval), torch.empty_like(inval)
        call("sb_dilate_box", ptr(inval), H, W, int(opt.dilate_invalid_kernel), 1, 1, ptr(tmp), stream())
        call("sb_dilate_box", ptr(tmp), H, W, 2 * int(opt.dilate_invalid_kernel), 0, 0, ptr(out), stream())
        inval = out
    return inval


def ssim_confidence(opt, frame, depth, color, K, inv_K, stereo_T, data_range=2.0):
    """The SSIM depth-confidence term of run_semantic_super.py's default configuration (data_loader.py:360-372,477-479):
    confs <- 0.5 confs + 0.5 sigmoid(mean_c SSIM(warp, image)), in place on frame.confs.  Two launches (sb_ssim_conf).
    data_range: skimage (absent here) derives 2.0 from a float image -- parity unpinned, see include/super_b200.h."""
    H, W = opt.height, opt.width
    Kh = torch.as_tensor(K).detach().cpu().reshape(-1, 4, 4)[0].float()
    iKh = torch.as_tensor(inv_K).detach().cpu().reshape(-1, 4, 4)[0].float()
    Th = torch.as_tensor(stereo_T).detach().cpu().reshape(-1, 4, 4)[0].float()
    P = torch.matmul(Kh, Th)[:3, :]
    ik = (ctypes.c_float * 9)(*[float(iKh[i, j]) for i in range(3) for j in range(3)])
    kt = (ctypes.c_float * 12)(*[float(P[i, j]) for i in range(3) for j in range(4)])
    warp = torch.empty((3, H, W), dtype=F32, device=depth.device)
    frame.ssim = torch.empty((H, W), dtype=F32, device=depth.device)
    call("sb_ssim_conf", ptr(depth.reshape(H, W).contiguous()), ptr(color.reshape(3, H, W).contiguous()), ik, kt, H, W,
         float(data_range), ptr(warp), ptr(frame.confs), ptr(frame.ssim), stream())


def dist2edge(opt, frame):
    """Normalised distance of every valid pixel's projection to the nearest edge pixel of its own class
    (data_loader.py:494-517): a per-class nearest-neighbour search in 2-D (sb_knn_class, K = 1).  Stored by the reference
    on new_data and carried onto new surfels; nothing on the tracking path reads it.  Returns (P,) f64 (0 where invalid)."""
    H, W, C = opt.height, opt.width, opt.num_classes
    dev = frame.vmap.device
    pts, off = graphfit.edge_points(frame.seg.view(H, W), C, H, W, margin=-(10 ** 6))       # no margin frame here
    out = torch.zeros(frame.P, dtype=F64, device=dev)
    valid = frame.valid
    n = int(valid.sum())
    if n == 0 or pts.shape[0] == 0:
        return out
    from .utils.utils import pcd2depth
    v, u, _, _ = pcd2depth({("color", 0): torch.empty((1, 3, H, W), device="meta"), "K": _K44(frame.cam)},
                           frame.vmap[valid, :3].to(F64), round_coords=False)
    # divisors as device tensors: torch's CUDA tensor / python-scalar division multiplies by the reciprocal (1 ulp off the
    # IEEE quotient the reference's CPU division gives)
    Wd, Hd = torch.tensor(float(W), dtype=F64, device=dev), torch.tensor(float(H), dtype=F64, device=dev)
    q = torch.stack([u / Wd, v / Hd], dim=1).contiguous()
    # the reference divides the INTEGER pixel coordinates of the edge points, i.e. in float32 (data_loader.py:507)
    r = torch.stack([pts[:, 0].float() / Wd.float(), pts[:, 1].float() / Hd.float()], dim=1).to(F64).contiguous()
    rseg = torch.repeat_interleave(torch.arange(C, device=dev, dtype=I32), (off[1:] - off[:-1]).long())
    d, _ = ops.knn(q, r, 1, qseg=frame.seg[valid].contiguous(), rseg=rseg.contiguous())
    out[valid] = d[:, 0]
    return out


def _K44(cam):
    K = torch.eye(4, dtype=torch.float64)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = cam.fx, cam.fy, cam.cx, cam.cy
    return K[None]


def depth_preprocessing(opt, models, inputs, frame=None, return_valid_map=False):
    """depth_preprocessing (/root/reference/utils/data_loader.py:333-523), same call: returns (data, inputs) with `data`
    the engine.Frame holding points / norms / radii / confs / validity (and seg, seg_conf) as dense device images."""
    depth, color = inputs[("depth", 0)], inputs[("color", 0)]
    time = float(torch.as_tensor(inputs["time"]).reshape(-1)[0]) if "time" in inputs else float(inputs["filename"][0])
    seg = inputs.get(("seg", 0))
    inval = extra_invalid_mask(opt, depth, seg=seg, mask=inputs.get("valid_mask"))
    divterm = inputs.get("divterm", 1.0 / (2.0 * 0.6 * 0.6))
    frame = preprocess(opt, depth, color, inputs["K"], inputs["inv_K"], time, frame=frame, inval=inval,
                       divterm=float(torch.as_tensor(divterm).reshape(-1)[0]), seg_scores=inputs.get(("seg_conf", 0)))
    if hasattr(opt, "disable_ssim_conf") and not opt.disable_ssim_conf:
        if "stereo_T" not in inputs:
            raise lib.SuperB200Error("the SSIM confidence term (default of run_semantic_super.py) needs inputs['stereo_T']; "
                                     "pass --disable_ssim_conf otherwise")
        ssim_confidence(opt, frame, depth, color, inputs["K"], inputs["inv_K"], inputs["stereo_T"])
    if return_valid_map:
        return frame, inputs, frame.valid.view(opt.height, opt.width)
    return frame, inputs


# ---- ED graph (once per sequence) -------------------------------------------------------------------
def build_graph(opt, frame, topology_only=False):
    """init_graph + DirectDeformGraph grid_mesh (/root/reference/super/graph_encoder.py:11-67,128-193) from the dense maps:
    one launch (sb_graph_build, csrc/graph.cu), then update_ed (nodes.py:154-168).  One host read of the three counts."""
    H, W, s = frame.H, frame.W, opt.mesh_step_size
    dev = frame.vmap.device
    l = lib.load()
    G = int(l.sb_graph_anchors(H, W, s))
    C = frame.seg_conf.shape[1] if frame.seg_conf is not None else 0
    prune = bool(C) and bool(getattr(opt, "hard_seg", False)) and bool(opt.mesh_face)     # graph_encoder.py:141-149
    ws = torch.zeros(3 * G, dtype=I32, device=dev)
    points, norms = torch.zeros((G, 3), dtype=F64, device=dev), torch.zeros((G, 3), dtype=F64, device=dev)
    uv = torch.zeros((G, 2), dtype=I32, device=dev)
    seg = torch.zeros(G, dtype=I32, device=dev) if C else None
    seg_conf = torch.zeros((G, C), dtype=F64, device=dev) if C else None
    edges, faces = torch.zeros((4 * G, 2), dtype=I32, device=dev), torch.zeros((2 * G, 3), dtype=I32, device=dev)
    lens, radii = torch.zeros(4 * G, dtype=F64, device=dev), torch.zeros(G, dtype=F64, device=dev)
    areas = torch.zeros(2 * G, dtype=F64, device=dev)
    node_pos, counts = torch.zeros(G, dtype=I32, device=dev), torch.zeros(3, dtype=I32, device=dev)
    call("sb_graph_build", ptr(frame.vmap), ptr(frame.nmap), ptr(frame.seg_conf), C, H, W, s, int(prune), ptr(ws),
         ptr(points), ptr(norms), ptr(uv), ptr(seg), ptr(seg_conf), ptr(edges), ptr(faces), ptr(lens), ptr(radii),
         ptr(areas), ptr(node_pos), ptr(counts), stream())
    J, E, Fc = (int(x) for x in counts.tolist())                # init only: a sync is fine here
    if not topology_only and J < opt.num_ED_neighbors + 1:
        raise lib.SuperB200Error(f"only {J} ED nodes on valid pixels: the graph needs more than num_ED_neighbors")
    g = NS()
    g.points, g.norms, g.radii = points[:J].contiguous(), norms[:J].contiguous(), radii[:J].contiguous()
    g.anchor_uv, g.node_pos = uv[:J].to(I64), node_pos[:J].contiguous()
    g.edge_index = edges[:E].t().to(I64).contiguous()
    g.triangles = faces[:Fc].t().to(I64).contiguous()
    g.triangles_i32 = faces[:Fc].t().contiguous()
    g.edges_lens, g.triangles_areas = lens[:E].contiguous(), areas[:Fc].contiguous()
    g.num, g.param_num = J, 7 * J
    if C:
        g.seg_conf, g.seg_i32 = seg_conf[:J].contiguous(), seg[:J].contiguous()
        g.seg = g.seg_i32.to(I64)
    else:
        g.seg_i32 = None
    if topology_only:
        return g
    # update_ed (/root/reference/super/nodes.py:154-168): K+1 nearest, drop self, weights use the query radius
    hard = bool(getattr(opt, "hard_seg", False)) and g.seg_i32 is not None
    dist, idx = ops.knn(g.points, g.points, opt.num_ED_neighbors + 1, qseg=g.seg_i32 if hard else None,
                        rseg=g.seg_i32 if hard else None)                  # nodes.py:157-163
    g.knn_indices = idx[:, 1:].contiguous()
    g.knn_w = ops.knn_weights(dist[:, 1:].contiguous(), g.knn_indices, g.radii, radius_mode=1)
    span = torch.zeros(1, dtype=I32, device=dev)
    call("sb_graph_pair_span", ptr(g.knn_indices), ptr(g.node_pos), J, g.knn_indices.shape[1], ptr(span), stream())
    g.block_bw_ed = int(span.item())                                        # ARAP pairs (init-time sync)
    return g


# ---- tracker -----------------------------------------------------------------------------------------
class Tracker:
    """Sequence state + per-frame step.  opt: the reference's option namespace (options.py flags)."""

    def __init__(self, opt, device="cuda", capacity_factor=2.5):
        lib.load()
        self.opt = opt
        self.dev = torch.device(device)
        self.H, self.W = opt.height, opt.width
        self.P = self.H * self.W
        self.cap = int(capacity_factor * self.P)
        self.cur = None              # SurfelBuffers holding the state
        self.alt = None              # ping-pong partner for compaction
        self.ED = None
        self.ws = None               # LM workspace
        self.fuse_ws = None
        self.n_bound = 0             # host upper bound on the row count
        self._n_pinned = torch.zeros(1, dtype=I32).pin_memory() if torch.cuda.is_available() else None
        self._n_event = None
        self._order = None           # visiting order of the next frame's J^T J pass (computed after compaction)
        self._bands = {}             # half bandwidth -> ops.Band
        self._order_rows = 0         # rows the precomputed order covers
        self._n_exact = None         # exact row count at the last hand-over, growth since the one before
        self._last_growth = 0
        self._order_redone = 0
        self.n_tmp = torch.zeros(1, dtype=I32, device=self.dev)
        self.overflow = torch.zeros(1, dtype=I32, device=self.dev)
        self.track_id = None
        self.frames = None           # two Frame buffers for step(), allocated on first use
        self._fi = 0
        self.time = None
        self.last_beta = None
        # "band": own cluster Cholesky in band storage (sb_band_solve); "dense": dense A + library Cholesky.
        # The default follows the faster one as measured on B200 (profiles/): see DESIGN.md section 5.
        self.solver = getattr(opt, "solver", os.environ.get("SB_SOLVER", "band"))
        self.cluster_size = int(getattr(opt, "solver_ctas", os.environ.get("SB_SOLVER_CTAS", "148")))
        self.band = None
        self.block_bw = torch.zeros(1, dtype=I32, device=self.dev)
        # pinned hand-over words: block half-bandwidth needed, band.overflow (bit 0 outside the band, bit 1 outside the
        # fixed-point range), sb_fuse's capacity-overflow flag
        self._bw_pinned = torch.zeros(3, dtype=I32).pin_memory() if torch.cuda.is_available() else None
        self.event_sink = None       # bench.py: {"jtj": [], "solve": []} receiving (begin, end) raw cudaEvent_t pairs
        self.events_per_frame = (3, 2)

    # -- row-count / band-width hand-over: ONE host wait per tracked frame -----------------------------------
    # After compaction the surfels' node tuples are final, so the NEXT frame's visiting order and the block
    # half-bandwidth its normal equations need are computed right there (_publish_count) and copied to pinned
    # memory.  track() waits for that copy (_refresh_bound) after the new frame's producer has been enqueued, so
    # the device has work while the host wakes up.  What the wait buys: exact row counts (grids and the tuple
    # sort are not sized for a loose upper bound) and a band that is exactly as wide as the pattern (the solve
    # is 60 % of the frame and scales with bw^2: 255 us at bw 300 against 312 us with a 25 % safety margin).
    def _publish_count(self):
        self._order_and_gather()
        if getattr(self, "_defer_handover", False):
            self._handover_pending = True      # inside tail_scope: after the captured launches have been issued
        else:
            self._handover()

    def _order_and_gather(self):
        self._order = None
        if self.ED is not None and self.ED.node_pos is not None and getattr(self.opt, "use_derived_gradient", True):
            # The host's bound at this point is (rows at frame start + H W); the frame really adds a few thousand rows.
            # Sort the rows that can be expected (last exact count + a margin that follows the last growth); if the
            # exact count turns out larger, _refresh_bound redoes the order (slow path, counted in _order_redone).
            rows = self.n_bound
            if self._n_exact is not None:
                rows = min(rows, self._n_exact + max(16384, 4 * max(0, self._last_growth)))
            self._order_rows = rows
            self._order = ops.tuple_order(self.cur.knn_idx[:rows], self.cur.n_dev, self.ED.node_pos, self.block_bw)
            self._gather_sorted(rows)

    def _handover(self):
        self._n_pinned.copy_(self.cur.n_dev, non_blocking=True)
        self._bw_pinned[0:1].copy_(self.block_bw, non_blocking=True)
        if self.band is not None:
            self._bw_pinned[1:2].copy_(self.band.overflow, non_blocking=True)
        self._bw_pinned[2:3].copy_(self.overflow, non_blocking=True)
        self._n_event = torch.cuda.Event()
        self._n_event.record()

    def _gather_sorted(self, rows):
        """The LM's view of the model in visiting order (points, node tuples, weights: final once the frame is compacted):
        gathered once here, read coalesced by every data-term pass of the next frame."""
        if getattr(self, "_sorted", None) is None:
            self._sorted = NS(points=torch.empty((self.cap, 3), dtype=F64, device=self.dev),
                              knn_idx=torch.empty((self.cap, 4), dtype=I32, device=self.dev),
                              knn_w=torch.empty((self.cap, 4), dtype=F64, device=self.dev))
        b, s = self.cur, self._sorted
        call("sb_gather_sorted", ptr(b.points), ptr(b.knn_idx), ptr(b.knn_w), ptr(self._order), int(rows), ptr(b.n_dev),
             ptr(s.points), ptr(s.knn_idx), ptr(s.knn_w), stream())
        self._sorted_rows = int(rows)

    def _refresh_bound(self):
        if self._n_event is None:
            return
        self._n_event.synchronize()
        self._n_event = None
        n = min(self.cap, int(self._n_pinned[0]))
        self._last_growth = 0 if self._n_exact is None else n - self._n_exact
        self._n_exact = self.n_bound = n
        # Both flags are raised by kernels of the frame that has just been handed over: its beta has been applied and fused
        # by the time the host sees them (the price of running one frame ahead).  The tracker state is not trustworthy
        # after either, so this is fatal rather than a warning.
        if int(self._bw_pinned[2]) != 0:
            raise lib.SuperB200Error(f"surfel capacity ({self.cap} rows) exceeded in the fusion of the previous frame: new "
                                     "surfels were dropped; construct Tracker with a larger capacity_factor")
        if self.band is not None and int(self._bw_pinned[1]) != 0:
            what = []
            if int(self._bw_pinned[1]) & 1:
                what.append("normal-equation entries fell outside the planned band")
            if int(self._bw_pinned[1]) & 2:
                what.append(f"a normal-equation addend exceeded the fixed-point range 2^{62 - self.band.fx_shift} "
                            "(lower Band.FX_SHIFT for such term weights)")
            raise lib.SuperB200Error("previous frame's LM solve: " + "; ".join(what))
        if self._order is not None:
            bwb = int(self._bw_pinned[0])
            if n > self._order_rows:             # more rows than the order covers: redo it (synchronises; rare)
                self._order = ops.tuple_order(self.cur.knn_idx[:n], self.cur.n_dev, self.ED.node_pos, self.block_bw)
                self._order_rows = n
                self._order_redone += 1
                self._gather_sorted(n)
                bwb = int(self.block_bw.item())
            self._plan_band(max(bwb, self.ED.block_bw_ed))

    def _plan_band(self, block_bw_needed):
        """Band storage exactly as wide as the pattern (rounded up to the solver's 32-column tile, which costs the
        solve nothing), or the dense path.  Bands are cached per width: the pattern's width is a running maximum."""
        J = self.ED.num
        bw = 7 * min(J - 1, block_bw_needed) + 6
        bw = min(7 * J - 1, (bw + 31) // 32 * 32)
        if self.solver != "band" or not lib.load().sb_band3_fits(7 * J, bw):
            self.band = None
            return
        if self.band is None or self.band.bw != bw:
            if bw not in self._bands:
                # fixed-point range follows the term weights: entries of J^T J scale with lambda^2 (defaults 1, 10, 1 ->
                # the default exponent; 4x the weight -> 4 bits more range, 4 bits less resolution)
                import math
                o = self.opt
                lam2 = max(float(o.sf_point_plane_weight) ** 2, float(o.mesh_arap_weight) ** 2 if o.mesh_arap else 0.0,
                           float(o.mesh_rot_weight) ** 2 if o.mesh_rot else 0.0, 1e-30)
                extra = max(0, math.ceil(math.log2(lam2 / 100.0)))
                self._bands[bw] = ops.Band(7 * J, bw, self.ED.node_pos, self.dev, fx_shift=ops.Band.FX_SHIFT - extra,
                                           fx_gshift=ops.Band.FX_GSHIFT - extra)
            self.band = self._bands[bw]

    @staticmethod
    def _new_events(n):
        out = []
        for _ in range(n):
            e = ctypes.c_void_p()
            call("sb_event_create", ctypes.byref(e))
            out.append(e.value)
        return out

    def num_surfels(self):
        """Exact row count (synchronises)."""
        return int(self.cur.n_dev.item())

    # -- frame stages -------------------------------------------------------------------------------------
    def next_frame(self):
        if self.frames is None:
            self.frames = [Frame(self.H, self.W, self.dev), Frame(self.H, self.W, self.dev)]
        self._fi ^= 1
        return self.frames[self._fi]

    def __del__(self):
        try:                      # the instantiated graph of the frame tail (tail_scope)
            h = getattr(self, "_tail_graph", None)
            if h:
                lib.load().sb_lm_graph_destroy(ctypes.byref(h))
        except Exception:
            pass

    def reset(self):
        """Forget the sequence but keep every allocation (surfel buffers, band stores, solver workspaces, Jacobian-row
        scratch): the next init_state() starts a new sequence of the same image size without touching the allocator --
        for runs over many short sequences."""
        self._spare = (self.cur, self.alt, self.fuse_ws) if self.cur is not None else getattr(self, "_spare", None)
        self.cur = self.alt = self.ED = None
        self.n_bound = 0
        self._n_event = self._order = None
        self._order_rows = 0
        self._n_exact = None
        self._last_growth = 0
        self.band = None
        self.track_id = None if getattr(self, "gt", None) is None else -torch.ones_like(self.track_id)
        self.track_rsts = {}
        self.time = self.last_beta = None
        self._finished_once = False
        self.overflow.zero_()
        for b in self._bands.values():
            b.overflow.zero_()

    def init(self, frame, filename=None):
        """Surfels.__init__ + the first prepareStableIndexNSwapAllModel (nodes.py:93-191, super.py:60-63)."""
        self.init_state(frame)
        self.finish_frame(frame, filename)

    def init_state(self, frame):
        """Surfels.__init__ (nodes.py:93-149): ED graph, one surfel per valid pixel, update_ed, update_sfed_knn."""
        opt, dev = self.opt, self.dev
        # the reference hands the graph over on the data object (sfdata.ED_nodes = models.mesh_encoder(...), super.py:49-50)
        self.ED = getattr(frame, "ED_nodes", None) or build_graph(opt, frame)
        frame.ED_nodes = None
        self.semantic = frame.seg_conf is not None
        C = frame.seg_conf.shape[1] if self.semantic else 0
        self.sem_weights = self.semantic and getattr(opt, "method", "super") == "semantic-super"
        spare = getattr(self, "_spare", None)
        if spare is not None and spare[0].n_classes == C:
            self.cur, self.alt, self.fuse_ws = spare          # reset(): reuse the previous sequence's buffers
            self.cur.stable.zero_()
        else:
            self.cur, self.alt = SurfelBuffers(self.cap, dev, C), SurfelBuffers(self.cap, dev, C)
            self.fuse_ws = torch.zeros(int(lib.load().sb_fuse_workspace_bytes(self.H, self.W, self.cap)), dtype=U8, device=dev)
        self._spare = None
        self._finished_once = False
        valid = frame.valid
        n = int(valid.sum())                               # init only: a sync is fine here
        if n > self.cap:
            raise lib.SuperB200Error("surfel capacity too small")
        b = self.cur
        b.points[:n] = frame.vmap[valid, :3].to(F64)
        b.norms[:n] = frame.nmap[valid, :3].to(F64)
        b.colors[:n] = frame.color.reshape(3, -1).t()[valid]
        b.confs[:n] = frame.confs[valid]
        b.radii[:n] = frame.radii[valid]
        b.time_stamp[:n] = frame.time
        b.stable[:n] = 1
        b.n_dev.fill_(n)
        hard = self.semantic and bool(getattr(opt, "hard_seg", False))
        if self.semantic:
            b.seg[:n] = frame.seg[valid]
            b.seg_conf[:n] = frame.seg_conf[valid]
        dist, idx = ops.knn(b.points[:n], self.ED.points, opt.num_neighbors, qseg=b.seg[:n] if hard else None,
                            rseg=self.ED.seg_i32 if hard else None)        # nodes.py:172-178
        b.knn_idx[:n] = idx
        b.knn_w[:n] = ops.knn_weights(dist, idx, self.ED.radii, 0, b.stable)      # also clears stable (radius test)
        if self.semantic:
            if self.sem_weights and not getattr(opt, "hard_seg", False):          # nodes.py:183-189
                call("sb_reweight_semantic", ptr(b.points), ptr(b.knn_idx), n, None, ptr(self.ED.points),
                     ptr(self.ED.radii), ptr(self.ED.seg_conf), ptr(b.seg_conf), C, ptr(b.knn_w), stream())
        # projdata of the init frame = pixel coordinates (x,y) of the valid pixels (nodes.py:143-145)
        pix = valid.nonzero()[:, 0]
        b.projdata[:n, 0] = (pix % self.W).to(F32)
        b.projdata[:n, 1] = (pix // self.W).to(F32)
        self.n_bound = n
        self.time = frame.time

    def finish_frame(self, frame, filename=None):
        """Surfels.prepareStableIndexNSwapAllModel (nodes.py:543-599): stability / time-stamp rule, compaction, tracked
        points; then the hand-over of the row count, the next frame's visiting order and the band plan."""
        first = not getattr(self, "_finished_once", False)
        self._compact(frame, keep_projdata=first)
        self._track_points(frame, filename if filename is not None else f"{int(frame.time):06d}")
        if first:                # band plan from the pattern of the initial tuples + ARAP pairs
            self.block_bw.zero_()
        self._publish_count()
        if first:
            self._refresh_bound()
        self._finished_once = True

    def _compact(self, frame, keep_projdata=False):
        opt = self.opt
        proj = self.cur.projdata.clone() if keep_projdata else None
        call("sb_compact", self.cur.ref(), self.alt.ref(), frame.ref(), float(frame.time), int(opt.th_time_steps),
             int(bool(opt.disable_removing_unstable_surfels)), ptr(self.track_id),
             0 if self.track_id is None else self.track_id.numel(), ptr(self.fuse_ws), self.fuse_ws.numel(), stream())
        if keep_projdata:   # init frame: projdata is the integer pixel grid, not a re-projection
            n = int(self.cur.n_dev.item())
            keep = (self.cur.stable[:n] != 0)
            m = int(keep.sum())
            self.alt.projdata[:m] = proj[:n][keep]
        self.cur, self.alt = self.alt, self.cur

    def solve(self, frame, u=10.0, v=7.5, minimal_loss=1e10):
        """The frame's deformation: LM_Solver.LM (LM.py:81-122) or GraphFit.forward (deform_mesh.py:232-379).
        Returns beta (J,7) [LM] or deform_verts (J+1,7) [autograd configuration]."""
        opt = self.opt
        self._refresh_bound()
        sfv = self.view(self.n_bound)
        if getattr(opt, "use_derived_gradient", True):
            order = self._order
            if order is None:
                order = ops.tuple_order(sfv.knn_indices, self.cur.n_dev, self.ED.node_pos, self.block_bw)
            elif getattr(self, "_sorted", None) is not None and self._sorted_rows >= self.n_bound and self.band is not None:
                # the frame loop reads the copies gathered into visiting order (coalesced); same values, same order
                # Row bound handed to the frame loop: rounded up to 16 k rows (the kernels stop at the device-side count; rows
                # behind it are never read), so that sb_lm_frame's argument block repeats from frame to frame and its
                # instantiated graph is simply launched again instead of being re-captured and updated (~90 us of host time
                # during which the device has nothing to do: the host has just waited for the previous frame)
                s = self._sorted
                nb_ = min(self.cap, -(-self.n_bound // 16384) * 16384)
                sfv = NS(points=s.points[:nb_], norms=sfv.norms, knn_indices=s.knn_idx[:nb_], knn_w=s.knn_w[:nb_], ED=self.ED)
                order = None
            jev = sev = tev = None
            if self.event_sink is not None and self.event_sink.get("frames_left", 1 << 30) > 0:
                # per-pass / per-solve events only on the first frames of the timed region: an event record costs ~1 us of
                # stream time, forty of them per frame would sit inside the number they are meant to explain
                self.event_sink["frames_left"] = self.event_sink.get("frames_left", 1 << 30) - 1
                jev, sev = self._new_events(2 * self.events_per_frame[0]), self._new_events(2 * self.events_per_frame[1])
                self.event_sink["jtj"] += list(zip(jev[0::2], jev[1::2]))
                self.event_sink["solve"] += list(zip(sev[0::2], sev[1::2]))
                if "timeline" in self.event_sink and len(self.event_sink["timeline"]) < self.event_sink.get("timeline_frames", 0):
                    tev = self._new_events(5 + 5 * int(opt.num_optimize_iterations))
                    self.event_sink["timeline"].append(tev)
            beta, self.ws = lm.lm_solve(sfv, (frame.vmap, frame.nmap), frame.cam, opt, ws=self.ws, u=u, v=v,
                                        minimal_loss=minimal_loss, n_dev=self.cur.n_dev, order=order, band=self.band,
                                        cluster_size=self.cluster_size, jtj_events=jev, solve_events=sev,
                                        row_capacity=self.cap, stage_events=tev)
        else:
            # autograd configuration of the reference (GraphFit, super.py:70-71): fused loss+gradient kernels
            seg = None
            if self.semantic and (getattr(opt, "sf_soft_seg_point_plane", False) or
                                  getattr(opt, "sf_hard_seg_point_plane", False) or getattr(opt, "sf_bn_morph", False)):
                seg = NS(sf_seg=self.cur.seg[: self.n_bound], sf_seg_conf=self.cur.seg_conf[: self.n_bound],
                         trg_seg_conf=frame.seg_conf, scores=frame.scores, edge_pts=None, edge_off=None)
                if getattr(opt, "sf_bn_morph", False):
                    seg.edge_pts, seg.edge_off = graphfit.edge_points(frame.seg.view(self.H, self.W), opt.num_classes,
                                                                      self.H, self.W)
            sfv.isStable = self.cur.stable[: self.n_bound]
            sfv.ED = NS(points=self.ED.points, knn_indices=self.ED.knn_indices, knn_w=self.ED.knn_w,
                        triangles=self.ED.triangles_i32, triangles_areas=self.ED.triangles_areas)
            beta, self.gf_ws = graphfit.graph_fit(sfv, (frame.vmap, frame.nmap), frame.cam, opt,
                                                  ws=getattr(self, "gf_ws", None), n_dev=self.cur.n_dev, seg=seg)
        self.last_beta = beta
        return beta

    def apply(self, deform):
        """Surfels.update (nodes.py:193-223): warp surfels and nodes by the frame's deformation; a (J+1)-row deformation
        carries the global transform of the autograd configuration in its last row (:204-205,211-212,219-222)."""
        if deform is None:
            return
        J, nb = self.ED.num, self.n_bound
        b = self.cur
        deform = deform.contiguous()
        ops.warp_update(b.points[:nb], b.norms[:nb], b.knn_idx[:nb], b.knn_w[:nb], self.ED.points, self.ED.norms,
                        deform[:J], n_dev=b.n_dev)
        if deform.shape[0] == J + 1:
            graphfit.update_global(b.points[:nb], b.norms[:nb], self.ED.points, self.ED.norms, deform, n_dev=b.n_dev)

    def fuse(self, frame):
        """Surfels.fuseInputData (nodes.py:270-541)."""
        pr = self.fuse_params(frame)
        call("sb_fuse", self.cur.ref(), frame.ref(), ptr(self.ED.points), ptr(self.ED.radii), self.ED.num,
             ctypes.byref(pr), ptr(self.track_id), 0 if self.track_id is None else self.track_id.numel(),
             ptr(self.n_tmp), ptr(self.overflow), ptr(self.fuse_ws), self.fuse_ws.numel(), stream())
        call("sb_copy_i32", ptr(self.cur.n_dev), ptr(self.n_tmp), 1, stream())
        self.n_bound = min(self.cap, self.n_bound + self.P)
        self.time = frame.time

    def track(self, frame, filename=None):
        """SuPer.fusion (/root/reference/super/super.py:66-83): solve -> update -> fuse -> compact."""
        beta = self.solve(frame)
        with self.tail_scope(beta):
            self.apply(beta)
            self.fuse(frame)
            self.finish_frame(frame, filename)
        return beta

    @contextlib.contextmanager
    def tail_scope(self, beta):
        """The frame's tail -- warp/update, fusion, compaction, the next frame's visiting order and its gathered copies: ~25
        dependent launches of this library and nothing else -- replayed as ONE CUDA graph (lib.graph_scope), like the LM loop
        inside sb_lm_frame.  apply / fuse / finish_frame (Surfels.update / fuseInputData / prepareStableIndexNSwapAllModel)
        are called inside it; the hand-over of the row count to the host (pinned copies, torch) follows the graph's launch.
        Plain execution with tracked points (their bookkeeping allocates), the autograd deformation and on the first frames."""
        scoped = (getattr(self, "_finished_once", False) and getattr(self, "gt", None) is None and beta is not None
                  and beta.shape[0] == self.ED.num and getattr(self.opt, "use_derived_gradient", True)
                  and getattr(self, "_sorted", None) is not None and os.environ.get("SB_TAIL_GRAPH", "1") != "0")
        if not scoped:
            yield
            return
        if getattr(self, "_tail_graph", None) is None:
            self._tail_graph = ctypes.c_void_p(0)
        self._defer_handover, self._handover_pending = True, False
        try:
            with lib.graph_scope(self._tail_graph):
                yield
        finally:
            self._defer_handover = False
        if self._handover_pending:
            self._handover()

    def enable_tracking(self, gt):
        """--tracking_gt_file: gt = {"%06d": (T,3) int array [x, y, valid]} (utils/utils.py:383-391).  Tracked surfel ids
        live on the device (nodes.py:124); the recorded reprojections are in self.track_rsts[filename] (T,3) f32."""
        self.gt = {k: torch.as_tensor(v, dtype=I32).to(self.dev).contiguous() for k, v in gt.items()}
        T = next(iter(self.gt.values())).shape[0]
        self.track_id = -torch.ones(T, dtype=I64, device=self.dev)
        self.track_rsts = {}

    def _track_points(self, frame, filename):
        """update_track_pts / init_track_pts (nodes.py:225-265,594-599): one launch, no host sync."""
        if getattr(self, "gt", None) is None or filename not in self.gt:
            return
        b = self.cur
        out = torch.zeros((self.track_id.numel(), 3), dtype=F32, device=self.dev)
        call("sb_track_points", ptr(b.points), ptr(b.stable), ptr(b.projdata), b.cap, ptr(b.n_dev), ptr(frame.vmap),
             self.H, self.W, ptr(self.gt[filename]), self.track_id.numel(), ptr(self.track_id), ptr(out), stream())
        self.track_rsts[filename] = out

    def fuse_params(self, frame):
        """merge thresholds + the semantic switches of fuseInputData (nodes.py:314-316,466-484,505-511)."""
        opt = self.opt
        sem = bool(getattr(self, "semantic", False))
        gate = sem and (bool(getattr(opt, "hard_seg", False)) or opt.data == "superv1")
        semw = sem and bool(getattr(self, "sem_weights", False))
        hard = sem and bool(getattr(opt, "hard_seg", False))
        return SbFuseParams(opt.th_dist, opt.th_cosine_ang, float(frame.time), int(bool(opt.disable_merging_new_surfels)),
                            int(bool(opt.disable_merging_exist_surfels)), int(bool(opt.disable_adding_new_surfels)),
                            int(gate), (1 if semw else 0) | (2 if semw and not hard else 0),
                            ptr(self.ED.seg_conf) if semw else None,
                            ptr(self.ED.seg_i32) if hard else None)

    def view(self, n):
        """Views of the first n rows of the current buffers, in the layouts the LM solver takes."""
        b = self.cur
        return NS(points=b.points[:n], norms=b.norms[:n], knn_indices=b.knn_idx[:n], knn_w=b.knn_w[:n], ED=self.ED)

    def step(self, depth, color, K, inv_K, time, inval=None, seg_scores=None, filename=None):
        """One SuPer.forward: preprocess + (init | track).  Returns beta or None.  seg_scores: (C,H,W) class scores
        (inputs[("seg_conf",0)]) for the Semantic-SuPer configuration."""
        frame = preprocess(self.opt, depth, color, K, inv_K, time, frame=self.next_frame(), inval=inval,
                           seg_scores=seg_scores)
        if filename is None:
            filename = f"{int(time):06d}"
        if self.cur is None:
            self.init(frame, filename)
            return None
        return self.track(frame, filename)

    def snapshot(self):
        """Exact-size copies of the state in the reference's layouts (synchronises)."""
        n = self.num_surfels()
        b = self.cur
        out = {name: getattr(b, name)[:n].clone() for name, _, _ in SURFEL_FIELDS}
        out["knn_indices"] = out.pop("knn_idx").to(I64)
        out["isStable"] = out.pop("stable").bool()
        if b.seg is not None:
            out["seg"] = b.seg[:n].to(I64)
            out["seg_conf"] = b.seg_conf[:n].clone()
        return out
