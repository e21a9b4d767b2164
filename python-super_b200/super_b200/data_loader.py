"""Input side of the drop-in: the reference's on-disk layout and per-frame `inputs` dict
(/root/reference/utils/data_loader.py:27-271, hard-coded intrinsics :201-211, disp_to_depth
/root/reference/depth/monodepth2/layers.py:16-25).  Decoding only -- every number the tracking path computes comes
from the CUDA library.  CNN depth / segmentation inference is out of scope (--load_depth / --load_seg inputs)."""
from __future__ import annotations

import os

import numpy as np
import torch

KS = {
    "superv1": np.array([[883.0, 0, 445.06, 0], [0, 883.0, 190.24, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32),
    "superv2": np.array([[768.98551924, 0, 292.8861567, 0], [0, 768.98551924, 291.61479526, 0], [0, 0, 1, 0],
                         [0, 0, 0, 1]], dtype=np.float32),
}


class SuPerDataset(torch.utils.data.Dataset):
    def __init__(self, opt):
        if not opt.load_depth:
            raise NotImplementedError("super_b200 takes precomputed depth (--load_depth): the depth CNNs of the reference "
                                      "(monodepth2 / RAFT-Stereo) are outside the ED tracking path")
        self.opt = opt
        self.ids = list(range(opt.start_id, opt.end_id))
        self.K = KS[opt.data]
        self.inv_K = np.linalg.pinv(self.K).astype(np.float32)

    def __len__(self):
        return len(self.ids)

    def __getitem__(self, i):
        from PIL import Image
        opt, t = self.opt, self.ids[i]
        H, W = opt.height, opt.width
        img = Image.open(os.path.join(opt.data_dir, opt.rgb_dir, f"{t:06d}-left{opt.img_ext}")).convert("RGB")
        if img.size != (W, H):
            img = img.resize((W, H), Image.LANCZOS)
        color = torch.from_numpy(np.asarray(img).transpose(2, 0, 1).astype(np.float32) / np.float32(255.0))
        disp = np.load(os.path.join(opt.data_dir, opt.depth_dir, f"{t:06d}{opt.depth_ext}")).astype(np.float32)
        min_disp, max_disp = 1.0 / opt.max_depth, 1.0 / opt.min_depth
        scaled = np.float32(min_disp) + np.float32(max_disp - min_disp) * disp
        depth = (np.float32(1.0) / scaled).astype(np.float32)
        out = {"filename": f"{t:06d}", "ID": t, "time": float(t), ("color", 0): color, ("color_aug", 0): color,
               ("disp", 0): torch.from_numpy(scaled)[None], ("depth", 0): torch.from_numpy(depth)[None],
               "K": torch.from_numpy(self.K), "inv_K": torch.from_numpy(self.inv_K),
               "divterm": 1.0 / (2.0 * 0.6 * 0.6)}
        if getattr(opt, "load_seg", False):
            p = os.path.join(opt.data_dir, opt.seg_dir, f"{t:06d}-left{opt.seg_ext}")
            conf = torch.as_tensor(np.load(p)).double()
            out[("seg_conf", 0)] = conf
            out[("seg", 0)] = conf.argmax(0, True).long()
        return out


def init_dataset(opt):
    """utils/shared_functions.py:171-176: DataLoader(batch 1, no shuffle, no workers)."""
    return torch.utils.data.DataLoader(SuPerDataset(opt), batch_size=1, shuffle=False, num_workers=0)


class InitNets:
    """utils/shared_functions.py:22-169 reduced to what the tracking path needs: `.super` (SuPer)."""

    def __init__(self, opt):
        from .super.super import SuPer
        self.opt = opt
        self.device = torch.device("cuda", opt.gpu)
        self.super = SuPer(opt)
