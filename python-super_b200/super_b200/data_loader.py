"""Input side of the drop-in: the reference's on-disk layout and per-frame `inputs` dict
(/root/reference/utils/data_loader.py:27-271, hard-coded intrinsics :201-211, disp_to_depth
/root/reference/depth/monodepth2/layers.py:16-25), and InitNets (/root/reference/utils/shared_functions.py:22-176).
Decoding only -- every number the tracking path computes comes from the CUDA library.  CNN depth / segmentation
inference is out of scope (--load_depth / --load_seg inputs).

init_dataset returns a PrefetchLoader: frames are decoded and collated by a background thread into PINNED host buffers
two frames ahead, so that SuPer.forward's `.to(device, non_blocking=True)` copies overlap the previous frame's kernels
(the reference decodes synchronously in the main thread, num_workers=0).
"""
from __future__ import annotations

import os
import queue
import threading

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

    def _color(self, path):
        from PIL import Image
        with open(path, "rb") as f:
            img = Image.open(f).convert("RGB")
        H, W = self.opt.height, self.opt.width
        if img.size != (W, H):
            img = img.resize((W, H), Image.LANCZOS)            # the reference's Image.ANTIALIAS (data_loader.py:64-65)
        return torch.from_numpy(np.asarray(img).transpose(2, 0, 1).astype(np.float32) / np.float32(255.0))

    def _disp(self, path):
        if self.opt.depth_ext == ".png":                       # data_loader.py:254-255
            from PIL import Image
            return np.asarray(Image.open(path)).astype(np.float32)
        return np.load(path).astype(np.float32)

    def __getitem__(self, i):
        opt, t = self.opt, self.ids[i]
        d = os.path.expanduser(opt.data_dir)
        color = self._color(os.path.join(d, opt.rgb_dir, f"{t:06d}-left{opt.img_ext}"))
        disp = self._disp(os.path.join(d, opt.depth_dir, f"{t:06d}{opt.depth_ext}"))
        min_disp, max_disp = 1.0 / opt.max_depth, 1.0 / opt.min_depth
        scaled = np.float32(min_disp) + np.float32(max_disp - min_disp) * disp
        depth = (np.float32(1.0) / scaled).astype(np.float32)
        stereo_T = np.eye(4, dtype=np.float32)
        stereo_T[0, 3] = -0.1                                   # left view, no flip (data_loader.py:125-130)
        out = {"filename": f"{t:06d}", "ID": t, "time": float(t), ("color", 0): color, ("color_aug", 0): color,
               ("disp", 0): torch.from_numpy(scaled)[None], ("depth", 0): torch.from_numpy(depth)[None],
               "K": torch.from_numpy(self.K), "inv_K": torch.from_numpy(self.inv_K), "stereo_T": torch.from_numpy(stereo_T),
               "divterm": 1.0 / (2.0 * 0.6 * 0.6)}
        if getattr(opt, "load_seg", False):
            p = os.path.join(d, opt.seg_dir, f"{t:06d}-left{opt.seg_ext}")
            if p.endswith(".npy"):                              # data_loader.py:229-238
                conf = torch.as_tensor(np.load(p)).double()
                label = conf.argmax(0, True).long()
            else:                                               # label PNG: one-hot scores (:240-246)
                from PIL import Image
                label = torch.from_numpy(np.asarray(Image.open(p)).astype(np.int64))[None]
                conf = torch.nn.functional.one_hot(label[0]).permute(2, 0, 1).double()
            out[("seg_conf", 0)] = conf
            out[("seg", 0)] = label
        if getattr(opt, "load_valid_mask", False):              # data_loader.py:376-383 (cv2.imread(..., 0) as bool)
            from PIL import Image
            m = np.asarray(Image.open(os.path.join(d, opt.valid_mask_dir, f"{t:06d}-left.png")).convert("L"))
            out["valid_mask"] = torch.from_numpy(m != 0)
        return out


class PrefetchLoader:
    """DataLoader(batch 1, no shuffle) (utils/shared_functions.py:171-176) with a decode thread: items are collated and
    their tensors moved to pinned memory `depth` frames ahead of the consumer."""

    def __init__(self, dataset, depth=2, pin=None):
        self.dataset, self.depth = dataset, int(depth)
        self.pin = torch.cuda.is_available() if pin is None else pin

    def __len__(self):
        return len(self.dataset)

    def _collate(self, item):
        out = {}
        for k, v in item.items():
            if torch.is_tensor(v):
                v = v[None]
                out[k] = v.pin_memory() if self.pin else v
            elif isinstance(v, str):
                out[k] = [v]
            elif isinstance(v, float):
                out[k] = torch.tensor([v], dtype=torch.float64)
            elif isinstance(v, int):
                out[k] = torch.tensor([v])
            else:
                out[k] = v
        return out

    def __iter__(self):
        q = queue.Queue(maxsize=self.depth)
        stop = object()

        def work():
            try:
                for i in range(len(self.dataset)):
                    q.put(self._collate(self.dataset[i]))
                q.put(stop)
            except BaseException as e:          # surface decode errors in the consumer thread
                q.put(e)

        threading.Thread(target=work, daemon=True).start()
        while True:
            item = q.get()
            if item is stop:
                return
            if isinstance(item, BaseException):
                raise item
            yield item


def init_dataset(opt):
    """utils/shared_functions.py:171-176."""
    return PrefetchLoader(SuPerDataset(opt))


class InitNets:
    """utils/shared_functions.py:22-169 reduced to what the tracking path needs: `.super` (SuPer), `.mesh_encoder`
    (DirectDeformGraph), `.renderer` (the splat renderer in pulsar's role), `.device`."""

    def __init__(self, opt):
        from .renderer import Renderer
        from .super.graph_encoder import DirectDeformGraph
        from .super.super import SuPer
        self.opt = opt
        self.device = torch.device("cuda", opt.gpu)
        if opt.method in ("super", "semantic-super"):
            self.super = SuPer(opt)
            self.mesh_encoder = DirectDeformGraph(opt)
        if getattr(opt, "renderer", None) is not None:
            self.renderer = Renderer(opt)
        if getattr(opt, "depth_model", None) is not None or getattr(opt, "seg_model", None) is not None:
            raise NotImplementedError("depth / segmentation CNN inference is outside the ED tracking path: precompute and "
                                      "pass --load_depth / --load_seg")
