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


def frame_from_newdata(nd, frame_np, H, W, dev="cuda"):
    """engine.Frame filled with the ORACLE's (i.e. the reference's) new_data values, so that fusion
    tests are teacher-forced on exact inputs."""
    from super_b200 import engine
    fr = engine.Frame(H, W, dev)
    valid = nd.valid.to(dev)
    fr.vmap[valid, :3] = nd.points.to(dev).float()
    fr.vmap[valid, 3] = 1.0
    fr.nmap[valid, :3] = nd.norms.to(dev).float()
    fr.radii[valid] = nd.radii.to(dev)
    fr.confs[valid] = nd.confs.to(dev)
    color = torch.from_numpy(frame_np["color"]).to(dev).contiguous()
    if hasattr(nd, "seg_conf"):
        C = nd.seg_conf.shape[1]
        fr.seg = torch.zeros(H * W, dtype=torch.int32, device=dev)
        fr.seg_conf = torch.zeros((H * W, C), dtype=torch.float64, device=dev)
        fr.seg[valid] = nd.seg.to(torch.int32).to(dev)
        fr.seg_conf[valid] = nd.seg_conf.to(dev)
        fr.scores = nd.seg_conf_in[0].to(dev).contiguous()
    fr.bind(color, camera(nd, H, W), float(frame_np["time"]))
    return fr


def tracker_from_state(opt, sf, dev="cuda"):
    """engine.Tracker whose device buffers hold an oracle/golden state."""
    from super_b200 import engine, lib
    trk = engine.Tracker(opt, device=dev)
    n = len(sf.points)
    C = sf.seg_conf.shape[1] if hasattr(sf, "seg_conf") else 0
    trk.semantic = C > 0
    trk.sem_weights = trk.semantic and getattr(opt, "method", "super") == "semantic-super"
    trk.cur, trk.alt = engine.SurfelBuffers(trk.cap, trk.dev, C), engine.SurfelBuffers(trk.cap, trk.dev, C)
    trk.fuse_ws = torch.zeros(int(lib.load().sb_fuse_workspace_bytes(trk.H, trk.W, trk.cap)), dtype=torch.uint8, device=dev)
    b = trk.cur
    b.points[:n] = sf.points.to(dev); b.norms[:n] = sf.norms.to(dev); b.colors[:n] = sf.colors.to(dev)
    b.confs[:n] = sf.confs.to(dev); b.radii[:n] = sf.radii.to(dev); b.time_stamp[:n] = sf.time_stamp.to(dev)
    b.knn_idx[:n] = sf.knn_indices.to(torch.int32).to(dev); b.knn_w[:n] = sf.knn_w.to(dev)
    b.projdata[:n] = sf.projdata.to(dev); b.stable[:n] = sf.isStable.to(torch.uint8).to(dev)
    if C:
        b.seg[:n] = sf.seg.to(torch.int32).to(dev); b.seg_conf[:n] = sf.seg_conf.to(dev)
    b.n_dev.fill_(n)
    trk.n_bound = n
    trk.ED = to_device_state(sf).ED
    if C:
        trk.ED.seg_conf = sf.ED.seg_conf.to(dev).contiguous()
    trk.ED.num = sf.ED.num
    trk.ED.node_pos = None
    trk._publish_count()
    return trk
