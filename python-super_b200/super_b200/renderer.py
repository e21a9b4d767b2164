"""Surfel splat renderer in the role of the reference's pulsar wrapper (/root/reference/renderer/renderer.py:12-78):

    img = models.renderer(inputs, data, colors=None, view_scale=1.0, rad=0.01, bg_col=torch.tensor([0., 0., 0.]))   # (H,W,3) f32

as Surfels.render_ calls it (/root/reference/super/nodes.py:630-642).  One z-buffer pass + one resolve pass of
csrc/face.cu (sb_render_splats): every surfel is a sphere of radius `rad`, the nearest sphere along a pixel's ray wins --
pulsar's blending with gamma = 1e-5 is that limit to within its 1e-5 blending weight.  pytorch3d is absent from this
stack, so the pixel-level output is NOT pinned against pulsar (its principal point is rounded to whole pixels,
renderer.py:45-46; here the exact K is used); it is pinned against oracle/render_oracle.py, a numpy restatement of the
sphere z-buffer, in tests/test_gpu_face.py.
"""
from __future__ import annotations

import ctypes

import torch

from .lib import SuperB200Error, call, intr_array, ptr, stream

F32 = torch.float32


def conf2color(confs):
    """Surfel confidence in [0,1] -> RGB heat colour (utils/utils.py:308-314 uses matplotlib's 'magma' table; matplotlib
    is not on this stack, so the map is its degree-6 polynomial fit: max deviation ~0.01 per channel, visualisation only)."""
    assert confs.dim() == 1, f"Point condfidences should be of shape (N,), but got {confs.shape}"
    t = confs.to(torch.float64).clamp(0.0, 1.0)[:, None]
    c = torch.tensor([[-0.002136485053939582, -0.000749655052795221, -0.005386127855323933],
                      [0.2516605407371642, 0.6775232436837668, 2.494026599312351],
                      [8.353717279216625, -3.577719514958484, 0.3144679030132573],
                      [-27.66873308576866, 14.26473078096533, -13.64921318813922],
                      [52.17613981234068, -27.94360607168351, 12.94416944238394],
                      [-50.76852536473588, 29.04658282127291, 4.23415299384598],
                      [18.65570506591883, -11.48977351997711, -5.601961508734096]], dtype=torch.float64, device=confs.device)
    out = c[6].expand(t.shape[0], 3)
    for k in range(5, -1, -1):
        out = out * t + c[k]
    return out.clamp(0.0, 1.0)


class Renderer(torch.nn.Module):
    """models.renderer.  `data` needs .points (N,3) and .colors (N,3); optional .mask (N,) u8 (e.g. isStable) and
    .n_dev (device row counter) let the tracker render straight from its capacity buffers."""

    def __init__(self, opt):
        super().__init__()
        self.height, self.width = opt.height, opt.width
        self.gamma = 1.0e-5
        self._zbuf = None

    def forward(self, inputs, data, colors=None, view_scale=1.0, rad=0.01, bg_col=torch.tensor([0.0, 0.0, 0.0]),
                return_depth=False):
        if colors is None:
            colors = data.colors
        points = data.points
        if not points.is_cuda:
            raise SuperB200Error("renderer: CUDA tensors only (no CPU path)")
        H, W = int(self.height * view_scale), int(self.width * view_scale)
        K = torch.as_tensor(inputs["K"]).reshape(-1, 4, 4)[0].cpu().double() * view_scale
        intr = intr_array(float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]))
        dev = points.device
        pts = points.to(torch.float64).contiguous()
        cols = colors.to(F32).contiguous()
        if self._zbuf is None or self._zbuf.numel() != H * W or self._zbuf.device != dev:
            self._zbuf = torch.empty(H * W, dtype=torch.int64, device=dev)
        img = torch.empty((H, W, 3), dtype=F32, device=dev)
        depth = torch.empty((H, W), dtype=F32, device=dev) if return_depth else None
        bg = (ctypes.c_float * 3)(*[float(x) for x in torch.as_tensor(bg_col).reshape(-1)[:3]])
        mask = getattr(data, "mask", None)
        call("sb_render_splats", ptr(pts), ptr(cols), ptr(mask), pts.shape[0], ptr(getattr(data, "n_dev", None)), intr, H, W,
             float(rad), bg, ptr(self._zbuf), ptr(img), ptr(depth), None, stream())
        return (img, depth) if return_depth else img


Pulsar = Renderer          # the name run scripts of the reference import
