"""CPU: the input side of the drop-in (super_b200.data_loader) produces the reference's `inputs` schema
(/root/reference/utils/data_loader.py:90-154) from the reference's on-disk layout."""
import numpy as np
import torch

from super_b200 import synth
from super_b200.data_loader import PrefetchLoader, SuPerDataset, init_dataset
from super_b200.options import SemanticSuPerOptions, SuPerOptions


def test_loader_schema_and_values(tmp_path):
    H, W = 48, 64
    synth.write_sequence(str(tmp_path), [1, 2, 3], H, W, with_seg=True)
    opt = SemanticSuPerOptions().parse(["--model_name", "t", "--data_dir", str(tmp_path), "--start_id", "1", "--end_id", "4",
                                        "--height", str(H), "--width", str(W), "--load_depth", "--load_seg",
                                        "--disable_ssim_conf"])
    loader = init_dataset(opt)
    assert isinstance(loader, PrefetchLoader) and len(loader) == 3
    items = list(loader)
    assert [it["filename"] for it in items] == [["000001"], ["000002"], ["000003"]]
    it = items[1]
    assert it[("color", 0)].shape == (1, 3, H, W) and it[("color", 0)].dtype == torch.float32
    assert it[("depth", 0)].shape == (1, 1, H, W) and it[("disp", 0)].shape == (1, 1, H, W)
    assert it["K"].shape == (1, 4, 4) and it["inv_K"].shape == (1, 4, 4) and it["stereo_T"].shape == (1, 4, 4)
    assert it["time"].dtype == torch.float64 and float(it["time"]) == 2.0 and int(it["ID"]) == 2
    assert it[("seg_conf", 0)].shape == (1, 3, H, W) and it[("seg_conf", 0)].dtype == torch.float64
    assert it[("seg", 0)].shape == (1, 1, H, W) and it[("seg", 0)].dtype == torch.int64
    assert float(it["K"][0, 0, 0]) == np.float32(768.98551924)             # superv2 intrinsics (data_loader.py:207-211)
    # the decoded frame equals the in-memory synthetic frame (disp -> depth by disp_to_depth, layers.py:16-25)
    f = synth.frame_inputs(2, H, W, data="superv2", with_seg=True)
    assert torch.equal(it[("depth", 0)][0], torch.from_numpy(f["depth"]))
    assert torch.equal(it[("color", 0)][0], torch.from_numpy(f["color"]))


def test_png_depth_and_valid_mask(tmp_path):
    from PIL import Image
    H, W = 32, 40
    synth.write_sequence(str(tmp_path), [1], H, W)
    (tmp_path / "seg" / "tissue").mkdir(parents=True)
    m = np.zeros((H, W), dtype=np.uint8)
    m[4:20, 5:30] = 255
    Image.fromarray(m).save(tmp_path / "seg" / "tissue" / "000001-left.png")
    d16 = (np.arange(H * W).reshape(H, W) % 200).astype(np.uint8)
    Image.fromarray(d16).save(tmp_path / "depth" / "000001.png")
    opt = SuPerOptions().parse(["--model_name", "t", "--data_dir", str(tmp_path), "--start_id", "1", "--end_id", "2",
                                "--height", str(H), "--width", str(W), "--load_depth", "--depth_ext", ".png",
                                "--load_valid_mask"])
    item = SuPerDataset(opt)[0]
    assert torch.equal(item["valid_mask"], torch.from_numpy(m != 0))
    disp = d16.astype(np.float32)
    scaled = np.float32(1.0 / 80.0) + np.float32(1.0 / 0.1 - 1.0 / 80.0) * disp
    assert torch.equal(item[("disp", 0)][0], torch.from_numpy(scaled))
