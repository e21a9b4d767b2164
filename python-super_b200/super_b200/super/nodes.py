"""Surfels: the reference's model container, call-compatible with /root/reference/super/nodes.py:36-803.

    sf = Surfels(opt, models, inputs, data)            # nodes.py:93: one surfel per valid pixel of the first frame
    sf.prepareStableIndexNSwapAllModel(inputs, data)   # :543  stability rule + compaction (+ tracked points, lazy render)
    sf.update(deform); sf.fuseInputData(inputs, data); sf.prepareStableIndexNSwapAllModel(inputs, data)   # super.py:73-78
    sf.update_ed(); sf.update_sfed_knn(); sf.evaluate()

`data` / `sfdata` is the engine.Frame the producer returns (dense per-pixel maps: the device form of the reference's
torch_geometric Data).  The state lives in capacity-sized device buffers (engine.Tracker, which this class extends);
attribute reads return exact-size tensors in the reference's layouts and dtypes -- they synchronise once
to learn the row count, the tracking path itself never does.  Attributes are read-only views: the methods above are the
way to change the model.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from .. import engine, ops
from ..lib import call, ptr, stream

F32, I64 = torch.float32, torch.int64


def evaluate(gt, est, igonored_ids=[], normalize=False):
    """Reprojection error of tracked points (nodes.py:17-34): gt, est (T,3) [x, y, valid]; -1 where gt is not valid."""
    val = (gt[:, 2] == 1)
    if len(igonored_ids) > 0:
        val[np.array(igonored_ids) - 1] = False
    dists = np.linalg.norm(gt[:, 0:2] - est[:, 0:2], axis=1)
    dists[~val] = -1
    if normalize:
        dists /= 480
    return dists


def get_gt(opt):
    """utils/utils.py:360-391: (file contents, gt dict, int keys, str keys, (N,T,3) array) of --tracking_gt_file."""
    data_dir = os.path.expanduser(opt.data_dir)
    if not os.path.exists(data_dir):
        raise ValueError(f"Path {data_dir} does not exist. This is likely an error with args.data_dir configuration.")
    gt_path = os.path.join(data_dir, opt.tracking_gt_file)
    if not os.path.exists(gt_path):
        raise ValueError("Ground truth file does not exist!")
    allpts = np.array(np.load(gt_path, allow_pickle=True)).tolist()
    gt = allpts["gt"]
    gt_intkeys = sorted(int(k) for k in gt.keys())
    gt_strkeys = sorted(f"{int(k):06d}" for k in gt.keys())
    gt = {f"{int(k):06d}": np.asarray(v) for k, v in gt.items()}
    return allpts, gt, gt_intkeys, gt_strkeys, np.stack([gt[k] for k in gt_strkeys], axis=0)


class _EDNodes:
    """ED_nodes attribute bag (graph_encoder.py:185-192) over the tracker's graph; indices come out as int64."""

    def __init__(self, g):
        self._g = g

    def __getattr__(self, k):
        g = object.__getattribute__(self, "_g")
        v = getattr(g, k)
        if k == "knn_indices":
            return v.to(I64)
        return v


class Surfels(engine.Tracker):
    _ROWS = {"points": "points", "norms": "norms", "colors": "colors", "confs": "confs", "radii": "radii",
             "time_stamp": "time_stamp", "knn_w": "knn_w", "projdata": "projdata"}

    def __init__(self, opt, models, inputs, data, capacity_factor=2.5):
        super().__init__(opt, device=data.vmap.device, capacity_factor=capacity_factor)
        self.models = models
        self.evaluate_tracking = getattr(opt, "tracking_gt_file", None) is not None
        self.hard_seg = bool(getattr(opt, "hard_seg", False))
        if getattr(opt, "method", "super") == "semantic-super":
            self.power_arg = (1 / 2, 1 / 2)
        self.output_dir = None
        self.summary_writer = None
        self.tracking_eval_errors = {}
        if opt.phase == "test":
            self.output_dir = os.path.join(getattr(opt, "output_dir", "results"), str(getattr(opt, "model_name", "model")))
            if self.evaluate_tracking:                                   # nodes.py:115-126
                self.track_pts, gt, self.gt_intkeys, self.gt_strkeys, self.gt_array = get_gt(opt)
                self.gt_np = gt
                self.enable_tracking(gt)
                self.track_num = self.gt_array.shape[1]
        self._renderImg = self._renderImg_conf_heat = None
        self._render_inputs = None
        self.init_state(data)

    # ---- attribute views in the reference's layouts ------------------------------------------------------------
    @property
    def ED_nodes(self):
        return _EDNodes(self.ED)

    @property
    def sf_num(self):
        return self.num_surfels()

    @property
    def surfel_num(self):
        return torch.tensor(self.num_surfels())

    @property
    def knn_indices(self):
        return self.cur.knn_idx[: self.num_surfels()].to(I64)

    @property
    def isStable(self):
        return self.cur.stable[: self.num_surfels()].bool()

    @property
    def seg(self):
        return self.cur.seg[: self.num_surfels()].to(I64)

    @property
    def seg_conf(self):
        return self.cur.seg_conf[: self.num_surfels()]

    def __getattr__(self, k):
        rows = type(self)._ROWS
        if k in rows:
            cur = self.__dict__.get("cur")
            if cur is None:
                raise AttributeError(k)
            return getattr(cur, rows[k])[: self.num_surfels()]
        raise AttributeError(k)

    # ---- the reference's methods -----------------------------------------------------------------------------------
    def update(self, deform):
        """nodes.py:193-223."""
        self.apply(deform)

    def fuseInputData(self, inputs, sfdata):
        """nodes.py:270-541."""
        self.fuse(sfdata)

    def prepareStableIndexNSwapAllModel(self, inputs, sfdata):
        """nodes.py:543-627.  The render of :625 is lazy: `renderImg` / `renderImg_conf_heat` render on first read."""
        fname = None
        if isinstance(inputs, dict) and "filename" in inputs:
            fname = inputs["filename"][0] if not isinstance(inputs["filename"], str) else inputs["filename"]
        self.finish_frame(sfdata, fname)
        self._renderImg = self._renderImg_conf_heat = None
        self._render_inputs = inputs

    def update_ed(self):
        """nodes.py:154-168: ED node-node neighbours (K+1 nearest, self dropped) and weights from the CURRENT node
        positions."""
        g, opt = self.ED, self.opt
        hard = self.hard_seg and g.seg_i32 is not None
        dist, idx = ops.knn(g.points, g.points, opt.num_ED_neighbors + 1, qseg=g.seg_i32 if hard else None,
                            rseg=g.seg_i32 if hard else None)
        g.knn_indices = idx[:, 1:].contiguous()
        g.knn_w = ops.knn_weights(dist[:, 1:].contiguous(), g.knn_indices, g.radii, radius_mode=1)

    def update_sfed_knn(self):
        """nodes.py:170-191: surfel -> node neighbours, the radius test on isStable and the (semantic) weights from the
        CURRENT positions.  Invalidates the precomputed visiting order of the J^T J pass."""
        opt, b, n = self.opt, self.cur, self.n_bound
        hard = self.semantic and self.hard_seg
        dist, idx = ops.knn(b.points[:n], self.ED.points, opt.num_neighbors, n_dev=b.n_dev,
                            qseg=b.seg[:n] if hard else None, rseg=self.ED.seg_i32 if hard else None)
        b.knn_idx[:n] = idx
        b.knn_w[:n] = ops.knn_weights(dist, idx, self.ED.radii, 0, b.stable, n_dev=b.n_dev)
        if self.semantic and self.sem_weights and not self.hard_seg:
            C = b.seg_conf.shape[1]
            call("sb_reweight_semantic", ptr(b.points), ptr(b.knn_idx), n, ptr(b.n_dev), ptr(self.ED.points),
                 ptr(self.ED.radii), ptr(self.ED.seg_conf), ptr(b.seg_conf), C, ptr(b.knn_w), stream())
        self._order = None
        self.block_bw.zero_()
        self._publish_count()

    # ---- lazy render (nodes.py:630-650) ---------------------------------------------------------------------------------
    def render_img(self, inputs):
        self._render_inputs = inputs
        self._renderImg = self._renderImg_conf_heat = None

    def _render(self):
        from ..renderer import conf2color
        renderer = getattr(self.models, "renderer", None)
        if renderer is None or self._render_inputs is None:
            return
        n = self.num_surfels()
        view = engine.NS(points=self.cur.points[:n], colors=self.cur.colors[:n], mask=self.cur.stable[:n], n_dev=self.cur.n_dev)
        rad = getattr(self.opt, "renderer_rad", 0.0002)
        self._renderImg = renderer(self._render_inputs, view, colors=view.colors, rad=rad).permute(2, 0, 1).unsqueeze(0)
        heat = conf2color(self.cur.confs[:n])
        self._renderImg_conf_heat = renderer(self._render_inputs, view, colors=heat, rad=rad).permute(2, 0, 1).unsqueeze(0)

    @property
    def renderImg(self):
        if self._renderImg is None:
            self._render()
        return self._renderImg

    @property
    def renderImg_conf_heat(self):
        if self._renderImg_conf_heat is None:
            self._render()
        return self._renderImg_conf_heat

    # ---- evaluation (nodes.py:754-803, utils/utils.py:406-513 without the matplotlib figures) ------------------------------
    def evaluate(self):
        """Reprojection errors of the tracked points against --tracking_gt_file.  Writes, under
        output_dir/model_name: tracking_rst.npy ({filename: (T,3) [u, v, 1]}, the file the reference's evaluation
        scripts read, options.py:260-261) and reprojerr.json (the scalars the reference logs as reprojerr/pythonsuper_*).
        Returns the dict of scalars (None when nothing was tracked)."""
        if not self.evaluate_tracking or len(getattr(self, "track_rsts", {})) == 0:
            return None
        rsts = {k: v.cpu().numpy() for k, v in self.track_rsts.items()}          # one D2H per labelled frame, here only
        for k in rsts.keys() & set(self.gt_strkeys):
            if k not in self.tracking_eval_errors:
                self.tracking_eval_errors[k] = evaluate(self.gt_np[k].astype(np.float64), rsts[k].astype(np.float64))
        if not self.tracking_eval_errors:
            return None
        keys = sorted(self.tracking_eval_errors.keys())
        err = np.stack([self.tracking_eval_errors[k] for k in keys], axis=0)                 # (N, T)
        out = {"time": self.time, "frames": keys, "reprojerr/pythonsuper_mean": float(np.mean(err)),
               "reprojerr/pythonsuper_std": float(np.std(err))}
        valid = err >= 0
        out["per_point_mean"] = [float(np.mean(err[:, i][valid[:, i]])) if valid[:, i].any() else None for i in range(err.shape[1])]
        edge_ids = list(getattr(self.opt, "edge_ids", []) or [])
        if edge_ids:
            sel = np.zeros(err.shape[1], dtype=bool)
            sel[np.array(edge_ids) - 1] = True
            out["reprojerr/pythonsuper_edge_pts_mean"] = float(np.mean(err[:, sel]))
            out["reprojerr/pythonsuper_edge_pts_std"] = float(np.std(err[:, sel]))
        if self.output_dir is not None:
            os.makedirs(self.output_dir, exist_ok=True)
            np.save(os.path.join(self.output_dir, "tracking_rst.npy"), rsts, allow_pickle=True)
            with open(os.path.join(self.output_dir, "reprojerr.json"), "w") as f:
                json.dump(out, f, indent=1)
        return out
