"""TEST INFRASTRUCTURE ONLY -- CPU restatements (numpy / torch CPU) of the pieces around the path that have no
counterpart in super_oracle.py: the sphere z-buffer that stands in for pulsar (parity UNPINNED: pytorch3d is absent,
/root/reference/renderer/renderer.py:63-78 cannot be run here), and the SSIM depth confidence
(/root/reference/utils/data_loader.py:360-372,477-479; skimage is absent too, so structural_similarity's published
algorithm -- Wang et al. 2004 as implemented by skimage.metrics with its defaults -- is restated: parity UNPINNED).
Only tests/ import this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def render_spheres(points, colors, K, H, W, rad, bg=(0.0, 0.0, 0.0)):
    """Nearest-sphere-wins rendering: every point is a sphere of radius rad; for each pixel the ray through its centre
    is intersected with the spheres whose projected disc covers it.  Returns (img (H,W,3) f32, depth (H,W) f32,
    index (H,W) i64 with -1 = background).  O(N * footprint), numpy loops: small inputs only."""
    fx, fy, cx, cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    img = np.tile(np.asarray(bg, dtype=np.float32), (H, W, 1))
    depth = np.zeros((H, W), dtype=np.float32)
    index = -np.ones((H, W), dtype=np.int64)
    best = np.full((H, W), np.inf, dtype=np.float32)
    pts = np.asarray(points, dtype=np.float64)
    for i, (X, Y, Z) in enumerate(pts):
        if not (Z > rad):
            continue
        u, v = X * fx / Z + cx, Y * fy / Z + cy
        ru, rv = rad * fx / Z, rad * fy / Z
        x0, x1 = max(0, int(np.ceil(u - ru))), min(W - 1, int(np.floor(u + ru)))
        y0, y1 = max(0, int(np.ceil(v - rv))), min(H - 1, int(np.floor(v + rv)))
        hit = False
        for yy in range(y0, y1 + 1):
            for xx in range(x0, x1 + 1):
                dx, dy = (xx - cx) / fx, (yy - cy) / fy
                dd, dp, pp = dx * dx + dy * dy + 1.0, dx * X + dy * Y + Z, X * X + Y * Y + Z * Z
                disc = dp * dp - dd * (pp - rad * rad)
                if disc < 0:
                    continue
                t = np.float32((dp - np.sqrt(disc)) / dd)
                if not t > 0:
                    continue
                hit = True
                if t < best[yy, xx] or (t == best[yy, xx] and i < index[yy, xx]):
                    best[yy, xx], index[yy, xx] = t, i
        if not hit:
            xr, yr = int(np.rint(u)), int(np.rint(v))
            if 0 <= xr < W and 0 <= yr < H:
                t = np.float32(Z - rad)
                if t < best[yr, xr] or (t == best[yr, xr] and i < index[yr, xr]):
                    best[yr, xr], index[yr, xr] = t, i
    m = index >= 0
    img[m] = np.asarray(colors, dtype=np.float32)[index[m]]
    depth[m] = best[m]
    return img, depth, index


def uniform_filter_reflect(x, size=7):
    """scipy.ndimage.uniform_filter(mode='reflect') on the last two axes of a float64 array."""
    r = size // 2
    xp = np.pad(x, [(0, 0)] * (x.ndim - 2) + [(r, r), (r, r)], mode="symmetric")
    c = np.cumsum(np.cumsum(np.pad(xp, [(0, 0)] * (x.ndim - 2) + [(1, 0), (1, 0)]), axis=-1), axis=-2)
    H, W = x.shape[-2:]
    s = c[..., size:size + H, size:size + W] - c[..., :H, size:size + W] - c[..., size:size + H, :W] + c[..., :H, :W]
    return s / (size * size)


def ssim_map(x, y, data_range=2.0, win=7):
    """structural_similarity(x, y, channel_axis=0, full=True)[1] with skimage's defaults; x, y (C,H,W)."""
    x, y = x.astype(np.float64), y.astype(np.float64)
    NP = win * win
    cov_norm = NP / (NP - 1.0)
    ux, uy = uniform_filter_reflect(x, win), uniform_filter_reflect(y, win)
    uxx, uyy, uxy = uniform_filter_reflect(x * x, win), uniform_filter_reflect(y * y, win), uniform_filter_reflect(x * y, win)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    return ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))


def ssim_confidence(depth, color, K, inv_K, stereo_T, confs, data_range=2.0):
    """data_loader.py:360-372,477-479: depth (H,W) f32, color (3,H,W) f32 torch CPU tensors, confs (H,W) f32.
    Returns (0.5 confs + 0.5 sigmoid(mean_c SSIM), SSIM mean map)."""
    H, W = depth.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W)], 0)
    cam = depth.reshape(1, -1) * (inv_K[:3, :3] @ pix)
    cam = torch.cat([cam, torch.ones(1, H * W)], 0)
    P = (K @ stereo_T)[:3, :]
    cp = P @ cam
    pc = (cp[:2] / (cp[2:3] + 1e-7)).reshape(2, H, W).permute(1, 2, 0).clone()
    pc[..., 0] /= W - 1
    pc[..., 1] /= H - 1
    pc = (pc - 0.5) * 2
    warp = F.grid_sample(color[None], pc[None], mode="bilinear", padding_mode="zeros", align_corners=False)[0]
    s = ssim_map(warp.numpy(), color.numpy(), data_range).mean(0)
    out = 0.5 * confs + 0.5 * torch.sigmoid(torch.from_numpy(s).float())
    return out, s
