"""Surfels: the reference's state container (/root/reference/super/nodes.py:36-91) as a view over the
device-resident buffers of super_b200.engine.Tracker.  Attribute reads return exact-size tensors in
the reference's layouts and dtypes (they synchronise once to learn the row count)."""
from __future__ import annotations

import torch


class _EDNodes:
    def __init__(self, g):
        self._g = g

    def __getattr__(self, k):
        g = object.__getattribute__(self, "_g")
        v = getattr(g, k)
        if k == "knn_indices":
            return v.to(torch.int64)
        return v


class Surfels:
    _MAP = {"points": "points", "norms": "norms", "colors": "colors", "confs": "confs", "radii": "radii",
            "time_stamp": "time_stamp", "knn_w": "knn_w", "projdata": "projdata"}

    def __init__(self, opt, tracker):
        self.opt = opt
        self._trk = tracker

    @property
    def ED_nodes(self):
        return _EDNodes(self._trk.ED)

    @property
    def time(self):
        return self._trk.time

    @property
    def sf_num(self):
        return self._trk.num_surfels()

    @property
    def surfel_num(self):
        return torch.tensor(self._trk.num_surfels())

    @property
    def knn_indices(self):
        return self._trk.cur.knn_idx[: self._trk.num_surfels()].to(torch.int64)

    @property
    def track_id(self):
        return self._trk.track_id

    @property
    def track_rsts(self):
        return self._trk.track_rsts

    @property
    def seg(self):
        return self._trk.cur.seg[: self._trk.num_surfels()].to(torch.int64)

    @property
    def seg_conf(self):
        return self._trk.cur.seg_conf[: self._trk.num_surfels()]

    @property
    def isStable(self):
        return self._trk.cur.stable[: self._trk.num_surfels()].bool()

    def __getattr__(self, k):
        m = type(self)._MAP
        if k in m:
            trk = object.__getattribute__(self, "_trk")
            return getattr(trk.cur, m[k])[: trk.num_surfels()]
        raise AttributeError(k)

    def update(self, deform):
        """Surfels.update (nodes.py:193-223)."""
        if deform is None:
            return
        from .. import ops
        trk = self._trk
        v = trk.view(trk.n_bound)
        ops.warp_update(v.points, v.norms, v.knn_indices, v.knn_w, trk.ED.points, trk.ED.norms,
                        deform.contiguous(), n_dev=trk.cur.n_dev)
