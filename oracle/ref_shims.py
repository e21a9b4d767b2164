"""TEST INFRASTRUCTURE ONLY -- import shims that let the UNMODIFIED reference at /root/reference
run on CPU in this container.  Used only by oracle/gen_golden.py and
oracle/validate_port.py, which run HERE (the build container); /root/reference does not exist on
the GPU box and nothing under tests/, bench.py or the product package imports this file.

What is stubbed (the reference imports these at module scope but they are absent here):
  pytorch3d.ops.knn_points      -> exact f64 brute force, ascending, ties -> lower index
                                   (pytorch3d 0.6.2 is NOT installed: kNN parity is "unpinned",
                                   this stub DEFINES it)
  pytorch3d pulsar Renderer     -> zeros image (output feeds tensorboard only)
  torch_geometric.data.Data     -> attribute bag
  open3d, segmentation_models_pytorch, skimage, matplotlib, opt_einsum -> empty / trivial
and `.cuda()` -> identity, so the reference's hard-coded device moves are no-ops on CPU.
"""
from __future__ import annotations

import logging
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = "/root/reference"


class Data:
    """Stand-in for torch_geometric.data.Data: only attribute get/set, items(), keys() are used
    (/root/reference/super/nodes.py:135,521)."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def items(self):
        return list(self.__dict__.items())

    def keys(self):
        return list(self.__dict__.keys())


def exact_knn(p1: torch.Tensor, p2: torch.Tensor, K: int, chunk: int = 8192):
    """K nearest rows of p2 for each row of p1: squared distance sum((a-b)^2) evaluated in the
    input dtype in x,y,z order, ascending, ties -> lower index (stable sort)."""
    d_out, i_out = [], []
    for s in range(0, p1.shape[0], chunk):
        a = p1[s:s + chunk]
        diff = a[:, None, :] - p2[None, :, :]
        d2 = diff[..., 0] * diff[..., 0]
        for c in range(1, diff.shape[-1]):
            d2 = d2 + diff[..., c] * diff[..., c]
        ds, idx = torch.sort(d2, dim=1, stable=True)
        d_out.append(ds[:, :K])
        i_out.append(idx[:, :K])
    if not d_out:
        return (torch.zeros((0, K), dtype=p1.dtype), torch.zeros((0, K), dtype=torch.long))
    return torch.cat(d_out), torch.cat(i_out)


def _knn_points(p1, p2, K=1, **kw):
    d, i = exact_knn(p1[0], p2[0], K)
    return d[None], i[None], None


def _ball_query(*a, **kw):
    raise NotImplementedError("ball_query is never selected by the reference's call sites")


def _matrix_to_rotation_6d(m):
    return m[..., :2, :].clone().reshape(m.shape[:-2] + (6,))


class _PulsarRenderer(torch.nn.Module):
    def __init__(self, width, height, n, **kw):
        super().__init__()
        self.w, self.h = int(width), int(height)

    def forward(self, *a, **kw):
        return torch.zeros((self.h, self.w, 3), dtype=torch.float32)


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Inject the stub modules and CPU monkey-patches, put /root/reference on sys.path."""
    global _installed
    if _installed:
        return
    _installed = True

    _module("pytorch3d")
    _module("pytorch3d.ops", knn_points=_knn_points, ball_query=_ball_query)
    _module("pytorch3d.transforms", quaternion_to_matrix=None, matrix_to_quaternion=None,
            matrix_to_rotation_6d=_matrix_to_rotation_6d)
    _module("pytorch3d.renderer")
    _module("pytorch3d.renderer.points")
    _module("pytorch3d.renderer.points.pulsar", Renderer=_PulsarRenderer)
    _module("torch_geometric")
    _module("torch_geometric.data", Data=Data)
    _module("open3d")
    _module("segmentation_models_pytorch")
    sk = _module("skimage")
    sk.io = _module("skimage.io", imread=None)
    sk.metrics = _module("skimage.metrics", structural_similarity=None)
    mpl = _module("matplotlib")
    mpl.pyplot = _module("matplotlib.pyplot",
                         get_cmap=lambda name: (lambda x: np.zeros((len(x), 4))))
    _module("opt_einsum", contract=torch.einsum)

    # CPU run: every hard-coded .cuda() / device='cuda' becomes a no-op.
    torch.Tensor.cuda = lambda self, *a, **kw: self
    torch.nn.Module.cuda = lambda self, *a, **kw: self
    torch.cuda.set_device = lambda *a, **kw: None
    torch.cuda.manual_seed = lambda *a, **kw: None
    torch.cuda.manual_seed_all = lambda *a, **kw: None
    torch.cuda.empty_cache = lambda *a, **kw: None
    _orig_tensor = torch.tensor

    def _tensor(*a, **kw):
        dev = kw.get("device", None)
        if dev is not None and "cuda" in str(dev):
            kw["device"] = "cpu"
        return _orig_tensor(*a, **kw)

    torch.tensor = _tensor

    from PIL import Image
    if not hasattr(Image, "ANTIALIAS"):
        Image.ANTIALIAS = Image.LANCZOS
    if not hasattr(np, "bool"):
        np.bool = bool

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def make_logger():
    lg = logging.getLogger("super_ref")
    lg.setLevel(logging.ERROR)
    return lg
