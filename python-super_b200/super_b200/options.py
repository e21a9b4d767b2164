"""Command-line surface of run_super.py / run_semantic_super.py: the flag names and defaults are the
drop-in contract (/root/reference/options.py:8-349).  Declared as a table; only the flags the ED
tracking path reads change behaviour, the CNN-related ones are accepted for compatibility."""
from __future__ import annotations

import argparse
import os

_HERE = os.path.dirname(os.path.abspath(__file__))

# (flag, kwargs)   'B' = store_true switch
B = {"action": "store_true"}
_SUPER_FLAGS = [
    ("method", dict(default="super")), ("phase", dict(default="test")),
    ("start_id", dict(type=int, default=4)), ("end_id", dict(type=int, default=521)),
    ("seed", dict(type=int, default=0)), ("use_derived_gradient", B),
    ("save_sample_freq", dict(type=int, default=10)), ("output_dir", dict(default="results")),
    ("model_name", dict(required=True)),
    ("optimizer", dict(default="SGD")), ("learning_rate", dict(type=float, default=5e-5)),
    ("num_optimize_iterations", dict(type=int, default=10)),
    ("num_ED_neighbors", dict(type=int, default=4)), ("num_neighbors", dict(type=int, default=4)),
    ("th_dist", dict(type=float, default=0.1)), ("th_cosine_ang", dict(type=float, default=0.4)),
    ("th_time_steps", dict(type=int, default=30)),
    ("disable_removing_unstable_surfels", B), ("disable_merging_new_surfels", B),
    ("disable_merging_exist_surfels", B), ("disable_adding_new_surfels", B),
    ("normal_model", dict(default="8neighbors")), ("deform_udpate_method", dict(default="super_edg")),
    ("downsample_params", dict(nargs="+", type=float, default=[0.1, 50, 0.1])),
    ("ball_piv_radii", dict(nargs="+", type=float, default=[0.08])),
    ("mesh_step_size", dict(type=int, default=30)),
    ("load_depth", B), ("depth_ext", dict(default=".npy")),
    ("min_depth", dict(type=float, default=0.1)), ("max_depth", dict(type=float, default=80.0)),
    ("depth_model", dict()), ("num_layers", dict(type=int, default=0)),
    ("valid_iters", dict(type=int, default=32)), ("hidden_dims", dict(nargs="+", type=int, default=[128, 128, 128])),
    ("corr_levels", dict(type=int, default=4)), ("corr_radius", dict(type=int, default=4)),
    ("shared_backbone", B), ("n_downsample", dict(type=int, default=2)), ("context_norm", dict(default="batch")),
    ("slow_fast_gru", B), ("n_gru_layers", dict(type=int, default=3)), ("corr_implementation", dict(default="reg")),
    ("mixed_precision", B), ("pretrained_depth_checkpoint_dir", dict()), ("pretrained_encoder_checkpoint_dir", dict()),
    ("post_process", B), ("depth_width_range", dict(nargs="+", type=float, default=[0.02, 0.98])),
    ("depth_filter_kernel_size", dict(type=int, default=-1)), ("weights_init", dict(type=str, default="pretrained")),
    ("optical_flow_model", dict()), ("renderer", dict(default="pulsar")),
    ("renderer_rad", dict(type=float, default=0.0002)),
    ("data_dir", dict(default=os.path.join(_HERE, "v1_520_pairs"))), ("rgb_dir", dict(default="rgb")),
    ("depth_dir", dict(default="depth")), ("seg_dir", dict(default="seg/DeepLabV3+")),
    ("data", dict(default="superv1")), ("height", dict(type=int, default=480)), ("width", dict(type=int, default=640)),
    ("img_ext", dict(default=".png")), ("load_valid_mask", B), ("valid_mask_dir", dict(default="seg/tissue")),
    ("dilate_invalid_kernel", dict(type=int, default=5)),
    ("sf_point_plane", B), ("sf_point_plane_weight", dict(type=float, default=1.0)),
    ("mesh_arap", B), ("mesh_arap_weight", dict(type=float, default=10.0)),
    ("mesh_face", B), ("mesh_face_weight", dict(type=float, default=1.0)),
    ("mesh_rot", B), ("mesh_rot_weight", dict(type=float, default=1.0)),
    ("sf_corr", B), ("sf_corr_match_renderimg", B), ("sf_corr_weight", dict(type=float, default=0.001)),
    ("sf_corr_loss_type", dict(default="point-point")),
    ("gpu", dict(type=int, default=0)), ("tracking_gt_file", dict()),
]

_SEMANTIC_FLAGS = [
    ("load_seg", B), ("seg_ext", dict(default=".npy")), ("seg_model", dict()), ("pretrained_seg_checkpoint_dir", dict()),
    ("seg_num_layers", dict(type=int)), ("hard_seg", B), ("del_seg_classes", dict(nargs="+", type=int, default=[])),
    ("disable_ssim_conf", B), ("num_classes", dict(type=int, default=3)),
    ("edge_ids", dict(nargs="+", type=int, default=[])), ("sf_hard_seg_point_plane", B), ("sf_soft_seg_point_plane", B),
    ("sf_bn_morph", B), ("sf_bn_morph_weight", dict(type=float, default=0.1)),
    ("render_loss", B), ("render_loss_weight", dict(type=float, default=1e-4)),
]
_SEMANTIC_DEFAULTS = dict(method="semantic-super", data="superv2", start_id=0, end_id=151, seg_dir="seg")


def _add(parser, flags):
    for name, kw in flags:
        parser.add_argument("--" + name, **kw)


class SuPerOptions:
    def __init__(self):
        self.parser = argparse.ArgumentParser(description="SuPer options (super_b200)")
        _add(self.parser, _SUPER_FLAGS)

    def parse(self, args=None):
        self.options = self.parser.parse_args(args)
        return self.options


class SemanticSuPerOptions(SuPerOptions):
    def __init__(self):
        super().__init__()
        _add(self.parser, _SEMANTIC_FLAGS)
        self.parser.set_defaults(**_SEMANTIC_DEFAULTS)
