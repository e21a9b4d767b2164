"""TEST INFRASTRUCTURE ONLY -- drive the UNMODIFIED reference (/root/reference) on CPU under the
shims of oracle/ref_shims.py, exactly like /root/reference/run_super.py:13-21, with recording hooks.

Runs only in the build container (needs /root/reference).  Consumers:
  oracle/gen_golden.py    -> tests/golden/*.npz   (committed fixtures)
  oracle/validate_port.py -> full-array comparison of oracle/super_oracle.py against the reference
  bench.py is NOT a consumer (the reference cannot travel to the GPU box).
"""
from __future__ import annotations

import os
import sys
import tempfile
import time

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "python-super_b200"))

from oracle import ref_shims  # noqa: E402


def _t2n(x):
    return x.detach().cpu().numpy().copy() if torch.is_tensor(x) else x


def snapshot_surfels(sf):
    """Copy every array of the reference's Surfels state (layouts: SURVEY.md 8(b))."""
    out = {}
    for k in ("points", "norms", "colors", "confs", "radii", "time_stamp", "knn_indices", "knn_w",
              "projdata", "isStable", "seg", "seg_conf", "dist2edge"):
        if hasattr(sf, k):
            out[k] = _t2n(getattr(sf, k))
    ed = sf.ED_nodes
    for k in ("points", "norms", "radii", "knn_indices", "knn_w", "edge_index", "edges_lens",
              "triangles", "triangles_areas", "seg", "seg_conf"):
        if hasattr(ed, k) and getattr(ed, k) is not None:
            out["ED_" + k] = _t2n(getattr(ed, k))
    if hasattr(sf, "track_id"):
        out["track_id"] = _t2n(sf.track_id)
    return out


def snapshot_newdata(d):
    out = {}
    for k in ("points", "norms", "colors", "radii", "confs", "valid", "index_map", "seg", "seg_conf",
              "dist2edge"):
        if hasattr(d, k):
            out[k] = _t2n(getattr(d, k))
    out["time"] = d.time
    return out


_CURRENT_REC = [None]


class Recorder:
    """Per-frame records filled by the hooks below."""

    def __init__(self, record_matrices=False):
        self.frames = []          # one dict per processed frame
        self.cur = None
        self.record_matrices = record_matrices

    def begin(self, t):
        self.cur = {"t": t, "lm_iters": [], "data_passes": []}
        self.frames.append(self.cur)


def run(argv, frames, height, width, with_seg=False, semantic=False, record_matrices=False,
        amp=None, quiet=True, time_only=False, tracking_gt=None, speed=1.0, seg_speed=None):
    """Run the reference over `frames` (list of ints; first = init frame).  Returns Recorder."""
    ref_shims.install()
    from super_b200 import synth

    tmp = tempfile.mkdtemp(prefix="super_ref_")
    synth.write_sequence(tmp, frames, height, width, with_seg=with_seg, amp=amp, speed=speed, seg_speed=seg_speed)
    if tracking_gt is not None:
        np.save(os.path.join(tmp, "gt.npy"), tracking_gt, allow_pickle=True)

    import options
    from utils.shared_functions import init_dataset, InitNets
    from utils.utils import set_seed
    import super.super as super_mod
    import super.LM as lm_mod
    import super.loss as loss_mod
    import super.nodes as nodes_mod
    import super.deform_mesh as dm_mod

    Opt = options.SemanticSuPerOptions if semantic else options.SuPerOptions
    args = ["--model_name", "golden", "--output_dir", os.path.join(tmp, "out"), "--data_dir", tmp,
            "--start_id", str(frames[0]), "--end_id", str(frames[-1] + 1),
            "--save_sample_freq", "100000", "--height", str(height), "--width", str(width),
            "--load_depth"] + list(argv)
    if tracking_gt is not None:
        args += ["--tracking_gt_file", "gt.npy"]
    opt = Opt().parse(args)
    set_seed(opt.seed)
    torch.set_num_threads(os.cpu_count())

    rec = Recorder(record_matrices)
    _CURRENT_REC[0] = rec

    if not time_only:
        # ---- hooks -------------------------------------------------------------------------
        orig_dp = super_mod.depth_preprocessing

        def dp_hook(opt_, models_, inputs_, **kw):
            data, inputs_ = orig_dp(opt_, models_, inputs_, **kw)
            rec.cur["new_data"] = snapshot_newdata(data)
            rec.cur["K"] = _t2n(inputs_["K"])
            return data, inputs_
        super_mod.depth_preprocessing = dp_hook

        orig_pcd2depth = loss_mod.pcd2depth
        orig_bil = loss_mod.LossTool.bilinear_intrpl_block
        pass_state = {}

        def pcd2depth_hook(inputs_, pcd, round_coords=True, valid_margin=0):
            r = orig_pcd2depth(inputs_, pcd, round_coords=round_coords, valid_margin=valid_margin)
            pass_state["proj"] = r
            return r
        loss_mod.pcd2depth = pcd2depth_hook

        def bil_hook(v, u, target_, index_map=None, grad=False, normalization=False):
            r = orig_bil(v, u, target_, index_map=index_map, grad=grad, normalization=normalization)
            pass_state.setdefault("bil", []).append((v.clone(), u.clone(), r[0].clone()))
            return r
        loss_mod.LossTool.bilinear_intrpl_block = staticmethod(bil_hook)

        orig_data_fwd = loss_mod.DataLoss.forward

        def data_fwd_hook(self, lambda_, beta, inputs_, new_data, grad=False, dldT_only=False):
            pass_state.clear()
            out = orig_data_fwd(self, lambda_, beta, inputs_, new_data, grad=grad, dldT_only=dldT_only)
            v_, u_, coords, _ = pass_state["proj"]
            nv = len(new_data.valid)
            valid_pair = new_data.valid[coords.clamp(0, nv - 1)] & (coords >= 0) & (coords < nv)
            (v, u, pts), (_, _, nrm) = pass_state["bil"]
            ok = ~torch.any(torch.isnan(pts) | torch.isnan(nrm), dim=1)
            ids = valid_pair.nonzero()[:, 0][ok]
            corners = torch.stack([torch.floor(v[ok]), torch.ceil(v[ok]),
                                   torch.floor(u[ok]), torch.ceil(u[ok])], dim=1).to(torch.int32)
            rec.cur["data_passes"].append({"grad": bool(grad), "ids": _t2n(ids.to(torch.int32)),
                                           "corners": _t2n(corners)})
            return out
        loss_mod.DataLoss.forward = data_fwd_hook

        orig_cost = lm_mod.LM_Solver.prepareCostTerm

        def cost_hook(self, sf, inputs_, new_data, beta, grad=False):
            out = orig_cost(self, sf, inputs_, new_data, beta, grad=grad)
            if grad:
                it = {"beta_in": _t2n(beta)}
                if rec.record_matrices:
                    it["jtj"], it["jtl"] = _t2n(out[0]), _t2n(out[1])
                else:
                    it["jtj_diag"], it["jtl"] = _t2n(torch.diagonal(out[0])), _t2n(out[1])
                rec.cur["lm_iters"].append(it)
            else:
                it = rec.cur["lm_iters"][-1]
                it["beta_try"] = _t2n(beta)
                it["loss"] = float(out)
                # per-term losses at the trial beta
                it["loss_terms"] = [float(torch.sum(t.forward(l, beta, inputs_, new_data)))
                                    for t, l in zip(self.losses, self.lambdas)]
                # that extra data pass is ours, not the reference's: drop its record
                if any(isinstance(t, loss_mod.DataLoss) for t in self.losses):
                    rec.cur["data_passes"].pop()
            return out
        lm_mod.LM_Solver.prepareCostTerm = cost_hook

        orig_solver = lm_mod.LM_Solver.Solver

        def solver_hook(A, b, method="cholesky"):
            x = orig_solver(A, b, method=method)
            it = rec.cur["lm_iters"][-1]
            it["delta"] = _t2n(x)
            it["u"] = float(A[0, 0]) - float(it.get("jtj_diag", np.diagonal(it.get("jtj", np.zeros((1, 1)))))[0])
            return x
        lm_mod.LM_Solver.Solver = staticmethod(solver_hook)

        orig_fuse = nodes_mod.Surfels.fuseInputData

        def fuse_hook(self, inputs_, sfdata):
            rec.cur["after_update"] = snapshot_surfels(self)
            r = orig_fuse(self, inputs_, sfdata)
            rec.cur["after_fuse"] = snapshot_surfels(self)
            return r
        nodes_mod.Surfels.fuseInputData = fuse_hook

        orig_super_fusion = super_mod.SuPer.fusion

        def fusion_hook(self, models_, inputs_, sfdata):
            dp = orig_super_fusion(self, models_, inputs_, sfdata)
            rec.cur["beta"] = _t2n(dp)
            return dp
        super_mod.SuPer.fusion = fusion_hook

        # autograd path: record per-iteration loss dicts
        orig_get_losses = dm_mod.GraphFit.get_losses

        def get_losses_hook(self, deform_verts, *a, **kw):
            loss, losses = orig_get_losses(self, deform_verts, *a, **kw)
            rec.cur.setdefault("ag_iters", []).append(
                {"deform_in": _t2n(deform_verts), "loss": float(loss),
                 "losses": {k: float(v) for k, v in losses.items()}})
            return loss, losses
        dm_mod.GraphFit.get_losses = get_losses_hook

        # ... and the gradient the optimiser actually consumes (after grad[-1] /= J)
        for opt_name in ("SGD", "Adam"):
            base = getattr(torch.optim, opt_name)
            if getattr(base, "_recording", False):
                continue

            def make(base_cls):
                class Recording(base_cls):
                    _recording = True

                    def step(self, *a, **kw):
                        p = self.param_groups[0]["params"][0]
                        r = _CURRENT_REC[0]
                        if r is not None and r.cur is not None and r.cur.get("ag_iters"):
                            r.cur["ag_iters"][-1]["grad"] = _t2n(p.grad)
                        return super().step(*a, **kw)
                return Recording
            setattr(torch.optim, opt_name, make(base))

    # ---- run -----------------------------------------------------------------------------------
    loader = init_dataset(opt)
    models = InitNets(opt)
    times = []
    for inputs in loader:
        t = int(inputs["filename"][0])
        rec.begin(t)
        if models.super.sf is not None and not hasattr(models.super.sf, "logger"):
            models.super.sf.logger = ref_shims.make_logger()
        t0 = time.perf_counter()
        models.super(models, inputs)
        times.append(time.perf_counter() - t0)
        rec.cur["wall_s"] = times[-1]
        if not time_only:
            rec.cur["state"] = snapshot_surfels(models.super.sf)
        if not quiet:
            print(f"frame {t}: {times[-1]:.3f}s  N={len(models.super.sf.points)}", flush=True)
    rec.models = models
    rec.opt = opt
    return rec


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--frames", type=int, default=3)
    ap.add_argument("--step", type=int, default=32)
    a = ap.parse_args()
    r = run(["--mesh_step_size", str(a.step), "--sf_point_plane", "--mesh_rot", "--mesh_arap",
             "--use_derived_gradient"], list(range(1, a.frames + 1)), a.height, a.width, quiet=False)
    for f in r.frames:
        print(f["t"], [it["loss"] for it in f["lm_iters"]])
