"""Move oracle-port state (CPU, reference layouts) onto the GPU in the layouts the C ABI takes."""
from types import SimpleNamespace as NS

import torch

from super_b200 import ops


def to_device_state(sf, dev="cuda"):
    d = NS()
    d.points = sf.points.to(dev).contiguous()
    d.norms = sf.norms.to(dev).contiguous()
    d.knn_indices = sf.knn_indices.to(torch.int32).to(dev).contiguous()
    d.knn_w = sf.knn_w.to(dev).contiguous()
    for k in ("colors", "confs", "radii", "time_stamp", "isStable"):
        if hasattr(sf, k):
            setattr(d, k, getattr(sf, k).to(dev).contiguous())
    ed = NS()
    ed.points = sf.ED.points.to(dev).contiguous()
    ed.norms = sf.ED.norms.to(dev).contiguous()
    ed.radii = sf.ED.radii.to(dev).contiguous()
    ed.knn_indices = sf.ED.knn_indices.to(torch.int32).to(dev).contiguous()
    ed.knn_w = sf.ED.knn_w.to(dev).contiguous()
    d.ED = ed
    return d


def device_maps(nd, H, W, dev="cuda"):
    return ops.dense_maps(nd.points.to(dev), nd.norms.to(dev), nd.valid.to(dev), H, W)


def camera(nd, H, W):
    return ops.Camera.from_K(nd.K, H, W)
