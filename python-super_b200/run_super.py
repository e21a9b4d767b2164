#!/usr/bin/env python
"""Drop-in for /root/reference/run_super.py:10-24: same flags, same per-frame call."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch
from tqdm import tqdm

from super_b200.options import SuPerOptions
from super_b200.data_loader import init_dataset, InitNets


def main(argv=None, options=SuPerOptions):
    opt = options().parse(argv)
    torch.manual_seed(opt.seed)
    torch.cuda.set_device(opt.gpu)
    loader = init_dataset(opt)
    models = InitNets(opt)
    for inputs in tqdm(loader):
        models.super(models, inputs)
    if models.super.sf is not None:          # the reference evaluates every save_sample_freq frames (super.py:79-80);
        models.super.sf.evaluate()           # one more at the end so that the last frames are in tracking_rst.npy
    return models


if __name__ == "__main__":
    main()
