"""DirectDeformGraph / init_graph, call-compatible with /root/reference/super/graph_encoder.py:11-193.

    graph = models.mesh_encoder(inputs, data)        # grid_mesh ED graph of the first frame (super.py:49-50)

`data` is the producer's engine.Frame.  The graph is built by ONE kernel (sb_graph_build, csrc/graph.cu) + update_ed's
kNN / weight launches; the returned attribute bag has the reference's fields (points, norms, radii, edge_index,
edges_lens, triangles, triangles_areas, num, param_num[, seg, seg_conf]) plus what the device tracker adds
(knn_indices / knn_w of update_ed, node_pos = the band solver's node order).  Only the grid_mesh branch exists: the
ball-pivoting and kNN branches of init_ED_nodes need open3d / are not selected by the run scripts.
"""
from __future__ import annotations

import torch

from .. import engine


def init_graph(valid, step=1):
    """(H,W) bool valid map -> (node mask (H,W) bool, edge_index (2,E) i64, triangles (3,F) i64), graph_encoder.py:11-67.
    Positions are irrelevant for the topology: the kernel is run on a frame whose maps only carry the validity."""
    H, W = valid.shape
    fr = engine.Frame(H, W, valid.device)
    fr.vmap[:, 3] = valid.reshape(-1).to(torch.float32)
    opt = engine.NS(mesh_step_size=int(step), num_ED_neighbors=0, hard_seg=False, mesh_face=False)
    g = engine.build_graph(opt, fr, topology_only=True)
    mask = torch.zeros((H, W), dtype=torch.bool, device=valid.device)
    mask[g.anchor_uv[:, 1], g.anchor_uv[:, 0]] = True
    return mask, g.edge_index, g.triangles


class DirectDeformGraph(torch.nn.Module):
    def __init__(self, opt) -> None:
        super().__init__()
        self.opt = opt

    def forward(self, inputs, data):
        return engine.build_graph(self.opt, data)
