#!/usr/bin/env python
"""bench.py -- tracked frames/sec of full ED tracking (producer + LM x10 + warp + fusion + compaction)
at 640x480 / mesh_step_size 32 (BASELINE.json config 2) on N B200s, one independent sequence per GPU.

    python bench.py --gpus 1 --steps K --warmup W            # this framework (C-ABI CUDA path)
    python bench.py --impl reference --steps K --warmup W    # reference algorithm on the host CPU (oracle port)
    torchrun ... bench.py --gpus N ...                        # N independent replicas, NCCL only for the metric gather

One JSON line on stdout (rank 0).  A "step" is one tracked frame.  `value` = frames/s with the frame's
inputs already resident in HBM; `e2e` = the same metric through the drop-in SuPer.forward call with
pinned HOST buffers (H2D of depth+colour and a D2H read of beta inside the timed region).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "python-super_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import torch

H, W, STEP, LM_ITERS = 480, 640, 32, 10
METRIC = "tracked frames/sec (ED warp+ICP+LM) at 640x480"
WORKLOAD = ("SuPer LM tracking, synthetic 640x480 deforming-surface depth, mesh_step_size 32, "
            "--sf_point_plane --mesh_rot --mesh_arap --use_derived_gradient, 10 LM iterations/frame")


def make_opt():
    from types import SimpleNamespace as NS
    return NS(method="super", phase="test", use_derived_gradient=True, num_optimize_iterations=LM_ITERS,
              num_ED_neighbors=4, num_neighbors=4, th_dist=0.1, th_cosine_ang=0.4, th_time_steps=30,
              disable_removing_unstable_surfels=False, disable_merging_new_surfels=False,
              disable_merging_exist_surfels=False, disable_adding_new_surfels=False, mesh_step_size=STEP,
              data="superv1", height=H, width=W, dilate_invalid_kernel=5, sf_point_plane=True,
              sf_point_plane_weight=1.0, mesh_arap=True, mesh_arap_weight=10.0, mesh_rot=True, mesh_rot_weight=1.0,
              mesh_face=False, mesh_face_weight=1.0, optimizer="SGD", learning_rate=5e-5)


def shape_time(i):
    """Frame index -> time argument of the synthetic SURFACE (SURVEY 8d formula): a triangle wave 1..21..1 (period 40).
    With the surface time running on (depth drifting by 0.2 % per frame) the reference's fusion rules let the model die
    out -- the CPU port loses 3/4 of its surfels within 120 frames, the device tracker likewise -- and the benchmark
    would time an ever lighter workload; with the oscillation the count stays at its nominal value (3.0e5 +- 3 %)."""
    period = 40
    j = i % period
    return 1 + (j if j < period // 2 else period - j)


def frames_host(n, seed=0):
    """The sequence: oscillating surface, frame TIME running on (i + 1), so that the time-stamp rule of the fusion
    (surfels unseen for th_time_steps frames are removed) stays in play."""
    from super_b200 import synth
    tex = synth.texture(H, W, seed)
    out = []
    for i in range(n):
        f = synth.frame_inputs(shape_time(i), H, W, tex=tex)
        f["time"], f["ID"], f["filename"] = float(i + 1), i + 1, f"{i + 1:06d}"
        out.append(f)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        """Samples from here on count: the process is started before the warm-up frames so that its start-up (NVML
        initialisation, which can stall driver calls for tens of milliseconds) stays out of the timed region."""
        self.t_begin = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter()
        time.sleep(0.15)
        self.proc.terminate()
        t0 = getattr(self, "t_begin", 0.0)
        rows = [r for t, r in self.rows if t0 <= t <= t_end + 0.05] or [r for _, r in self.rows[-2:]]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active")
                                                         for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def solver_report(trk, solve_ms):
    """The LM step solve (the frame's longest kernel; latency-bound pivot chain, not an HBM pass): measured time per
    solve and the FP64 rate of its algorithmic flops n*bw^2 (band Cholesky) against the dense (7J)^3/3 the reference
    pays in cuSOLVER."""
    if not solve_ms or trk.band is None:
        return {"kernel": "library dense Cholesky (torch.linalg / cuSOLVER)", "ms_per_solve": None}
    n, bw = trk.band.n, trk.band.bw
    ms = float(np.mean(solve_ms))
    variant = os.environ.get("SB_BAND_VARIANT", "4")
    kern = ("band_reverse (+ fixed-point store -> f64 band) + band_chol3_dual + band_combine + band_chol3 + band_backsub4 kernels (sb_band_solve4_step_fx: two-sided solve, LM step folded into the last kernel)"
            if variant == "4" else "band_chol3_kernel (sb_band_solve3)")
    return {"kernel": kern, "n": n, "half_bandwidth": bw, "ms_per_solve": ms,
            "solves_timed": len(solve_ms), "band_flops": float(n) * bw * bw, "dense_flops": float(n) ** 3 / 3.0,
            "achieved_gflops_band": float(n) * bw * bw / (ms * 1e-3) / 1e9, "bound": "latency (sequential pivot chain)"}


OTHER_CONFIGS = {
    # BASELINE.json configs other than the headline one: (H, W, mesh_step_size, option overrides, with_seg, data)
    "c3a_lm_step16_640x480": (480, 640, 16, {}, False, "superv1"),
    "c3b_adam_step16_640x480": (480, 640, 16, {"use_derived_gradient": False, "mesh_face": True, "optimizer": "Adam"}, False, "superv1"),
    "c4_semantic_640x480": (480, 640, 32, {"use_derived_gradient": False, "mesh_face": True, "mesh_arap": False,
                                           "sf_point_plane": False, "optimizer": "SGD", "method": "semantic-super",
                                           "data": "superv2", "num_classes": 3, "sf_soft_seg_point_plane": True,
                                           "sf_hard_seg_point_plane": False, "sf_bn_morph": True, "sf_bn_morph_weight": 0.1,
                                           "hard_seg": False, "del_seg_classes": [], "disable_ssim_conf": True}, True, "superv2"),
    "c5_lm_1280x1024": (1024, 1280, 32, {}, False, "superv1"),
}


def probe_config(name, dev, frames=8, warm=3):
    """Frames/s of the device tracker on one of the other BASELINE configs (inputs resident in HBM, CUDA events around
    the timed frames, L2 not flushed).  Parity of each at its own size: tests/test_gpu_sequences.py."""
    from super_b200 import engine, synth
    Hc, Wc, step, over, with_seg, data = OTHER_CONFIGS[name]
    opt = make_opt()
    opt.height, opt.width, opt.mesh_step_size = Hc, Wc, step
    for k, v in over.items():
        setattr(opt, k, v)
    tex = synth.texture(Hc, Wc)
    fr = [synth.frame_inputs(t, Hc, Wc, data=data, tex=tex, with_seg=with_seg, seg_speed=3.0 if with_seg else None)
          for t in range(1, frames + warm + 2)]
    dd = [torch.from_numpy(f["depth"]).to(dev) for f in fr]
    seg = [torch.from_numpy(f["seg_conf"]).to(dev) for f in fr] if with_seg else [None] * len(fr)
    dc = torch.from_numpy(fr[0]["color"]).to(dev)
    K, iK = torch.from_numpy(fr[0]["K"]), torch.from_numpy(fr[0]["inv_K"])
    trk = engine.Tracker(opt, device=dev)
    for i in range(1 + warm):
        trk.step(dd[i], dc, K, iK, fr[i]["time"], seg_scores=seg[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(1 + warm, 1 + warm + frames):
        trk.step(dd[i], dc, K, iK, fr[i]["time"], seg_scores=seg[i])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / frames
    out = {"frames_per_s": 1e3 / ms, "ms_per_frame": ms, "ed_nodes": int(trk.ED.num), "surfels": trk.num_surfels(),
           "frames_timed": frames, "solver": "LM" if opt.use_derived_gradient else opt.optimizer,
           "half_bandwidth": None if trk.band is None else int(trk.band.bw)}
    del trk
    torch.cuda.empty_cache()
    return out


def reduce_over_ranks(total_ms, e2e_ms, info, world, device):
    """The only cross-rank step of the benchmark (replicas are independent sequences): MAX of the timed
    intervals over ranks and a gather of the per-rank summaries.  Backend-agnostic (NCCL on the GPUs, gloo in the
    CPU test tests/test_multiprocess.py)."""
    if world <= 1:
        return total_ms, e2e_ms, None
    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    gathered = [None] * world
    torch.distributed.all_gather_object(gathered, info)
    return float(t[0]), float(t[1]), gathered


def aggregate_value(world, steps, total_ms):
    """Whole-job frames/s: every rank tracked `steps` frames of its own sequence within the slowest rank's time."""
    return world * steps / (total_ms / 1e3)


# ----------------------------------------------------------------------------------------------------
def run_cuda(args, rank, world, local_rank):
    from super_b200 import engine, lib
    from super_b200.super.super import SuPer

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib.load()
    K, Wm = args.steps, args.warmup
    nfr = 1 + Wm + K
    host = frames_host(nfr, seed=rank)                      # one independent sequence per rank
    opt = make_opt()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- arm 1: inputs resident in HBM ----------------------------------------------------------------
    dd = [torch.from_numpy(f["depth"]).to(dev) for f in host]
    dc = torch.from_numpy(host[0]["color"]).to(dev)
    Kt, iKt = torch.from_numpy(host[0]["K"]), torch.from_numpy(host[0]["inv_K"])
    trk = engine.Tracker(opt, device=dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    trk.step(dd[0], dc, Kt, iKt, host[0]["time"])
    for i in range(1, 1 + Wm):
        trk.step(dd[i], dc, Kt, iKt, host[i]["time"])
    n_surf_first = trk.num_surfels()
    barrier()
    sampler.mark_begin()
    lib.LAUNCHES = 0
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    # (begin, end) CUDA events on the launch stream around every J^T J pass and every linear solve of every timed frame, recorded inside sb_lm_frame (SbLMFrame.jtj_events / solve_events)
    trk.event_sink = {"jtj": [], "solve": [], "timeline": [], "timeline_frames": min(8, K), "frames_left": min(8, K)}
    trk.events_per_frame = (LM_ITERS, LM_ITERS)            # every J^T J pass (the first one of a frame is L2-cold) and every solve
    t_wall = time.perf_counter()
    for k in range(K):
        i = 1 + Wm + k
        flush.fill_(k & 0xff)                                # L2 flush, outside the timed events
        ev[k][0].record()
        trk.step(dd[i], dc, Kt, iKt, host[i]["time"])
        ev[k][1].record()
    barrier()
    wall = time.perf_counter() - t_wall
    launches = lib.LAUNCHES
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))

    def ev_ms(pairs):
        out = []
        for a, b in pairs:
            ms = ctypes.c_float()
            lib.call("sb_event_elapsed_ms", a, b, ctypes.byref(ms))
            out.append(ms.value)
        return out
    jt_ms, solve_ms = ev_ms(trk.event_sink["jtj"]), ev_ms(trk.event_sink["solve"])
    # live per-stage times of the LM loop (events after every stage of sb_lm_frame, first frames of the timed region)
    names = ["lm_begin", "eval_decide", "gram", "scatter"]
    for it in range(LM_ITERS):
        names += ["solve"] + (["eval_decide", "gram", "scatter"] if it + 1 < LM_ITERS else ["loss_decide"])
    stage_us = {}
    for tev in trk.event_sink["timeline"]:
        for i, nm in enumerate(names):
            if i + 1 < len(tev):
                stage_us.setdefault(nm, []).append(1e3 * ev_ms([(tev[i], tev[i + 1])])[0])
        for e in tev:
            lib.call("sb_event_destroy", e)
    timeline = {k: {"launches_per_frame": len(v) // max(1, len(trk.event_sink["timeline"])), "us_mean": float(np.mean(v)),
                    "us_per_frame": float(np.sum(v)) / max(1, len(trk.event_sink["timeline"]))} for k, v in stage_us.items()}
    for a, b in trk.event_sink["jtj"] + trk.event_sink["solve"]:
        lib.call("sb_event_destroy", a)
        lib.call("sb_event_destroy", b)
    trk.event_sink = None
    n_surf = trk.num_surfels()
    st = trk.ws.state.read()
    overflow = int(trk.overflow.item())

    # ---- arm 2: end to end through the drop-in SuPer.forward with pinned host buffers ------------------
    pin = []
    for f in host:
        pin.append({("depth", 0): torch.from_numpy(f["depth"])[None].pin_memory(),
                    ("color", 0): torch.from_numpy(f["color"])[None].pin_memory(),
                    "K": torch.from_numpy(f["K"])[None], "inv_K": torch.from_numpy(f["inv_K"])[None],
                    "time": torch.tensor([f["time"]], dtype=torch.float64), "filename": [f["filename"]],
                    "ID": torch.tensor([f["ID"]]), "divterm": torch.tensor([f["divterm"]], dtype=torch.float64)})
    h2d = int(pin[0][("depth", 0)].numel() * 4 + pin[0][("color", 0)].numel() * 4)
    model = SuPer(opt)
    models = type("Models", (), {})()
    models.super = model
    beta_host = torch.zeros((1, 7), dtype=torch.float64).pin_memory()
    model(models, dict(pin[0]))
    for i in range(1, 1 + Wm):
        model(models, dict(pin[i]))
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d2h = 0
    ev_prev = None
    beta_hosts = [beta_host, beta_host]
    use_prefetch = os.environ.get("SB_E2E_PREFETCH", "1") != "0"
    if use_prefetch:
        model.prefetch(pin[1 + Wm])      # like every later frame's: one frame ahead, under the previous frame's kernels
    for k in range(K):
        beta = model(models, dict(pin[1 + Wm + k]))
        if use_prefetch and k + 1 < K:
            model.prefetch(pin[1 + Wm + k + 1])                  # next frame's H2D copies on a side stream (SuPer.prefetch)
        if beta_host.shape != beta.shape:
            beta_host = torch.zeros(beta.shape, dtype=torch.float64).pin_memory()
            beta_hosts = [beta_host, torch.zeros(beta.shape, dtype=torch.float64).pin_memory()]
        # D2H read of the step's result into pinned memory, every step; the host waits for it ONE step later (a consumer that
        # works one frame behind), so that issuing the next frame is not held up by the read -- all reads are complete before
        # the timed region ends
        beta_hosts[k & 1].copy_(beta, non_blocking=True)
        ev_read = torch.cuda.Event()
        ev_read.record()
        if ev_prev is not None:
            ev_prev.synchronize()
        ev_prev = ev_read
        d2h = beta.numel() * 8
    if ev_prev is not None:
        ev_prev.synchronize()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    e2e_wall = time.perf_counter() - t0

    # ---- reduce over ranks: max time, sum frames ----------------------------------------------------------
    total_ms, e2e_ms, gathered = reduce_over_ranks(total_ms, e2e_ms, {"rank": rank, "fps": K / (sum(step_ms) / 1e3),
                                                                     "surfels": n_surf}, world, dev)
    if rank != 0:
        return

    value = aggregate_value(world, K, total_ms)
    N, P = n_surf, H * W
    b_pass = 44 * N + 28 * P                                 # algorithmic bytes of one surfel pass
    pass_avg_ms = float(np.mean(jt_ms)) if jt_ms else None          # evaluation + Gram + scatter launches of one data-term pass
    # the kernel the roofline is quoted on: the evaluation launch, which is the one that reads SURVEY 8(d)'s bytes; its
    # duration = the CUDA-event interval around it inside sb_lm_frame (includes the launch gap in front of it)
    jt_avg_ms = (timeline["eval_decide"]["us_mean"] / 1e3) if "eval_decide" in timeline else pass_avg_ms
    peak = peaks.get("hbm_gbs", 6650.0)
    achieved = (b_pass / 1e9) / (jt_avg_ms / 1e3) if jt_avg_ms else None
    traffic, traffic_note = None, None
    import glob
    tpaths = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_jtj_traffic.json")))
    tpath = tpaths[-1] if tpaths else ""
    if tpath:        # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        traffic_note = (f"ncu capture at {tj['surfels_at_capture']} surfels (algorithmic bytes there "
                        f"{tj['algorithmic_bytes_at_capture']}): {tj['source']}")
    out = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": total_ms / K, "ms_per_lm_iteration": (total_ms / K) / LM_ITERS,
        "ms_per_step_median": float(np.median(step_ms)), "ms_per_step_max": float(np.max(step_ms)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "configs_index": 1 if (H, W) == (480, 640) else 4, "frames_timed": K, "surfels": N, "surfels_first_timed_frame": n_surf_first, "ed_nodes": int(trk.ED.num),
                   "parallelism": f"{world} independent sequence replica(s), no data-path collective",
                   "l2": "256 MiB buffer written between timed steps (outside the per-step CUDA events)",
                   "timing": "sum of per-step CUDA-event intervals on the launch stream, max over ranks"},
        "e2e": {"value": world * K / (e2e_ms / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / K, "wall_ms_per_step": 1e3 * e2e_wall / K,
                "api": "super_b200.super.super.SuPer.forward(models, inputs) with pinned host depth+colour" + ("; SuPer.prefetch(next inputs) starts the next frame's host->device copies on a side stream, inside the timed region" if use_prefetch else "") + "; every step's result is copied to pinned host memory, the host waits for the copy one step later"},
        "gpu_launches": launches,
        "gpu_launches_per_lm_iteration": launches / (K * LM_ITERS),
        "clocks": clocks,
        "roofline": {"kernel": "data_eval_decide_kernel (warp+project+bilinear+residual+28-entry Jacobian row per surfel of one LM iteration; ARAP/Rot blocks and the LM decision ride in the same launch; the Gram and scatter launches behind it read its row buffer)",
                     "data_term_pass_ms": pass_avg_ms,
                     "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_note": traffic_note,
                     "algorithmic_bytes": b_pass, "launch_ms": jt_avg_ms,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s"},
        "solver": solver_report(trk, solve_ms),
        "lm_timeline_live": timeline,
        "lm_trace_last_frame": {"loss": [float(x) for x in st["loss"]], "accept": [int(x) for x in st["accept"]]},
        "wall_s_timed_region": wall, "capacity_overflow": overflow, "tuple_order_redone": int(trk._order_redone),
    }
    if gathered:
        out["per_rank"] = gathered
    if world == 1 and not args.no_configs:
        # the other BASELINE configs, next to the headline (which stays on config 2): driver-visible frames/s
        out["configs"] = {}
        for name in OTHER_CONFIGS:
            try:
                out["configs"][name] = probe_config(name, dev)
            except Exception as e:          # a probe must not take the headline line down
                out["configs"][name] = {"error": repr(e)[:200]}
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(iters=args.cpu_iters)
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------------
def cpu_setup():
    from oracle import super_oracle as so
    opt = so.default_opt(height=H, width=W, mesh_step_size=STEP)
    host = frames_host(2)
    trk = so.Tracker(opt, assemble="sparse_mm")
    trk.step(host[0])
    return so, opt, trk, host


def cpu_one_iteration(so, opt, sf, nd, beta, u):
    """One LM iteration as the reference executes it (LM.py:96-107): normal equations (COO Jacobian +
    torch.sparse.mm), damping, Cholesky solve, loss-only pass."""
    A, g, _ = so.lm_normal_equations(opt, sf, nd, beta, "sparse_mm")
    n = A.shape[0]
    A[torch.arange(n), torch.arange(n)] += u
    L = torch.linalg.cholesky(A)
    delta = torch.cholesky_solve(g, L).view(-1, 7)
    loss, _ = so.lm_cost(opt, sf, nd, beta + delta)
    return beta + delta, float(loss)


def cpu_baseline(iters=3):
    """Oracle port (reference algorithm, torch CPU, all host threads) on a bounded sample: `iters` LM
    iterations of the first tracked frame + one update/fuse/compact, extrapolated to 10 iterations/frame."""
    torch.set_num_threads(os.cpu_count())
    so, opt, trk, host = cpu_setup()
    t0 = time.perf_counter()
    nd = so.preprocess(opt, host[1])
    t_pre = time.perf_counter() - t0
    J = trk.sf.ED.num
    beta = torch.tensor([[1., 0, 0, 0, 0, 0, 0]], dtype=torch.float64).repeat(J, 1)
    u, its = 10.0, []
    for _ in range(iters):
        t0 = time.perf_counter()
        beta, _ = cpu_one_iteration(so, opt, trk.sf, nd, beta, u)
        its.append(time.perf_counter() - t0)
        u /= 7.5
    t0 = time.perf_counter()
    so.update(opt, trk.sf, beta)
    so.fuse(opt, trk.sf, nd)
    so.compact(opt, trk.sf, float(host[1]["time"]))
    t_rest = time.perf_counter() - t0
    t_it = float(np.mean(its))
    frame_s = t_pre + LM_ITERS * t_it + t_rest
    return {"value": 1.0 / frame_s, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{iters} LM iterations ({t_it:.2f} s each) + producer {t_pre:.2f} s + update/fuse/compact "
                      f"{t_rest:.2f} s of tracked frame 1, extrapolated to {LM_ITERS} iterations/frame",
            "ms_per_lm_iteration": 1e3 * t_it}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the Python
    reference cannot travel to the GPU box).  Each step = one LM iteration of a tracked frame."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    so, opt, trk, host = cpu_setup()
    t0 = time.perf_counter()
    nd = so.preprocess(opt, host[1])
    t_pre = time.perf_counter() - t0
    J = trk.sf.ED.num
    ident = torch.tensor([[1., 0, 0, 0, 0, 0, 0]], dtype=torch.float64).repeat(J, 1)
    beta, u = ident.clone(), 10.0
    budget_s = 240.0
    t_start = time.perf_counter()
    for _ in range(args.warmup):
        if time.perf_counter() - t_start > 0.3 * budget_s:
            break
        beta, _ = cpu_one_iteration(so, opt, trk.sf, nd, beta, u)
    times, steps_done = [], 0
    for k in range(args.steps):
        if k % LM_ITERS == 0:
            beta, u = ident.clone(), 10.0
        t0 = time.perf_counter()
        beta, _ = cpu_one_iteration(so, opt, trk.sf, nd, beta, u)
        times.append(time.perf_counter() - t0)
        u /= 7.5
        steps_done += 1
        if time.perf_counter() - t_start > budget_s:       # keep the whole run within a few minutes
            break
    t0 = time.perf_counter()
    so.update(opt, trk.sf, beta)
    so.fuse(opt, trk.sf, nd)
    so.compact(opt, trk.sf, float(host[1]["time"]))
    t_rest = time.perf_counter() - t0
    t_it = float(np.mean(times))
    frame_s = t_pre + LM_ITERS * t_it + t_rest
    value = 1.0 / frame_s
    sample = (f"each step = one LM iteration (normal equations via COO Jacobian + torch.sparse.mm, Cholesky, loss pass) "
              f"of a 640x480 tracked frame on the host CPU; {steps_done} steps timed ({t_it:.2f} s each); frame time = "
              f"producer {t_pre:.2f} s + {LM_ITERS} x iteration + update/fuse/compact {t_rest:.2f} s")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
           "steps": steps_done, "warmup": args.warmup, "ms_per_step": 1e3 * frame_s,
           "ms_per_lm_iteration": 1e3 * t_it, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "configs_index": 1},
           "extrapolated": True,     # frame time = producer + 10 x (timed LM iteration) + update/fuse/compact, not a timed frame
           "cpu_baseline": {"value": value, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2 (default, the headline): 640x480; c5: BASELINE config 5's 1280x1024 sequences, one per GPU")
    ap.add_argument("--cpu-iters", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "c5":       # BASELINE config 5: 1280x1024 (~1.3 M surfels per sequence), one sequence per GPU
        global H, W, METRIC, WORKLOAD
        H, W = 1024, 1280
        METRIC = "tracked frames/sec (ED warp+ICP+LM) at 1280x1024"
        WORKLOAD = WORKLOAD.replace("640x480", "1280x1024")
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_cuda(args, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
