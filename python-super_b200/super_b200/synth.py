"""Synthetic deforming-tissue sequences in the reference's on-disk / in-memory shapes.

The reference ships no data; BASELINE.json's configs are all quoted on synthetic
640x480 (or 1280x1024) deforming-surface depth + random texture (+ 3-class seg logits).
The formulas are SURVEY.md Appendix C.3 (the ones the survey's reference numbers used):

    z_t(x,y) = 1.0 + A*sin(2*pi*x/W + 0.05 t)*cos(2*pi*y/H) + 0.002 t        A = 0.05 * W/640
    disp_t   = (1/z_t - 1/max_depth) / (1/min_depth - 1/max_depth)            float32

`disp` is what the reference's loader reads from `<data_dir>/depth/%06d.npy` and turns back
into depth with disp_to_depth(., 0.1, 80)  (/root/reference/utils/data_loader.py:257-266,
/root/reference/depth/monodepth2/layers.py:16-25).  The amplitude is scaled with the image
width so that small parity-test images see the same surface slope per pixel as 640x480.

Nothing here touches the GPU; frames are numpy arrays (the host buffers of the e2e bench).
"""
from __future__ import annotations

import os
import numpy as np

MIN_DEPTH = 0.1
MAX_DEPTH = 80.0

# Hard-coded intrinsics of the reference's datasets (/root/reference/utils/data_loader.py:201-211).
INTRINSICS = {
    "superv1": (883.0, 883.0, 445.06, 190.24),
    "superv2": (768.98551924, 768.98551924, 292.8861567, 291.61479526),
}


def make_K(data: str = "superv1") -> np.ndarray:
    fx, fy, cx, cy = INTRINSICS[data]
    K = np.eye(4, dtype=np.float32)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = fx, fy, cx, cy
    return K


def texture(height: int, width: int, seed: int = 0) -> np.ndarray:
    """Fixed random uint8 RGB texture (H,W,3); same image for every frame."""
    rng = np.random.default_rng(seed)
    return (rng.random((height, width, 3)) * 255).astype(np.uint8)


def depth_at(t: float, height: int, width: int, amp: float | None = None) -> np.ndarray:
    """Ground-truth depth z_t (float64, H x W).  `t` may be scaled by the caller (speed * frame id)."""
    if amp is None:
        amp = 0.05 * width / 640.0
    yy, xx = np.mgrid[0:height, 0:width]
    return (1.0 + amp * np.sin(xx / width * 2 * np.pi + 0.05 * t) * np.cos(yy / height * 2 * np.pi)
            + 0.002 * t)


def disp_at(t: float, height: int, width: int, amp: float | None = None) -> np.ndarray:
    """float32 'disp' image in the loader's convention (see module docstring)."""
    z = depth_at(t, height, width, amp)
    return (((1.0 / z) - 1.0 / MAX_DEPTH) / (1.0 / MIN_DEPTH - 1.0 / MAX_DEPTH)).astype(np.float32)


def seg_logits_at(t: int, height: int, width: int) -> np.ndarray:
    """(3,H,W) float32 class scores with two slowly moving vertical boundaries (Appendix C.3)."""
    _, xx = np.mgrid[0:height, 0:width]
    s = width / 640.0
    b1 = 0.4 * width + 3 * t * s
    b2 = 0.7 * width + 2 * t * s
    sc = 20.0 * s
    return np.stack([-(xx - b1) / sc,
                     -np.abs(xx - (b1 + b2) / 2) / sc + (b2 - b1) / (2 * sc),
                     (xx - b2) / sc]).astype(np.float32)


def write_sequence(data_dir: str, frames, height: int, width: int, seed: int = 0,
                   with_seg: bool = False, amp: float | None = None, speed: float = 1.0,
                   seg_speed: float | None = None) -> None:
    """Write the reference's on-disk layout (/root/reference/utils/data_loader.py:179-199):
    rgb/%06d-left.png (+right), depth/%06d.npy, seg/%06d-left.npy."""
    from PIL import Image
    os.makedirs(os.path.join(data_dir, "rgb"), exist_ok=True)
    os.makedirs(os.path.join(data_dir, "depth"), exist_ok=True)
    if with_seg:
        os.makedirs(os.path.join(data_dir, "seg"), exist_ok=True)
    tex = Image.fromarray(texture(height, width, seed))
    for t in frames:
        tex.save(os.path.join(data_dir, "rgb", f"{t:06d}-left.png"))
        tex.save(os.path.join(data_dir, "rgb", f"{t:06d}-right.png"))
        np.save(os.path.join(data_dir, "depth", f"{t:06d}.npy"), disp_at(t * speed, height, width, amp))
        if with_seg:
            np.save(os.path.join(data_dir, "seg", f"{t:06d}-left.npy"),
                    seg_logits_at(t * (speed if seg_speed is None else seg_speed), height, width))


def frame_inputs(t: int, height: int, width: int, data: str = "superv1", seed: int = 0,
                 with_seg: bool = False, amp: float | None = None, tex: np.ndarray | None = None,
                 speed: float = 1.0, seg_speed: float | None = None) -> dict:
    """One frame as the host-side arrays the loader would produce (before batching):
    color (3,H,W) f32 in [0,1], disp/depth (1,H,W) f32, K, inv_K (4,4) f32, filename, time."""
    if tex is None:
        tex = texture(height, width, seed)
    color = np.ascontiguousarray(tex.transpose(2, 0, 1)).astype(np.float32) / np.float32(255.0)
    disp = disp_at(t * speed, height, width, amp)
    # disp_to_depth: python-double scalars applied to a float32 tensor (scalar cast to f32 per op)
    min_disp, max_disp = 1.0 / MAX_DEPTH, 1.0 / MIN_DEPTH
    scaled = np.float32(min_disp) + np.float32(max_disp - min_disp) * disp
    depth = (np.float32(1.0) / scaled).astype(np.float32)
    K = make_K(data)
    out = {
        "filename": f"{t:06d}", "ID": t, "time": float(t),
        "color": color, "disp": scaled[None], "depth": depth[None],
        "K": K, "inv_K": np.linalg.pinv(K).astype(np.float32),
        "divterm": 1.0 / (2.0 * 0.6 * 0.6),
    }
    if with_seg:
        out["seg_conf"] = seg_logits_at(t * (speed if seg_speed is None else seg_speed), height, width).astype(np.float64)
    return out
