"""Runtime configuration.  The reference reads a module-global ``device`` from configs/config.json at
import time (utils/basic_anchors.py:5-9, nets/rpn.py:11-15, nets/frcnn_training.py:13-17); here it is
one setting, defaulting to the process's current CUDA device (one process per GPU)."""
from __future__ import annotations

import os

import torch

device = os.environ.get("FRCNN_DEVICE")  # e.g. "cuda:0"; None -> current CUDA device


def get_device() -> torch.device:
    if device is not None:
        return torch.device(device)
    return torch.device("cuda", torch.cuda.current_device())
