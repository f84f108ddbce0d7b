"""Drop-in for the reference's utils/basic_anchors.py (same names, arguments, defaults, layouts);
the arithmetic runs in the sm_100a kernels of csrc/boxmath.cu."""
from __future__ import annotations

from .. import config, functional as F


def generate_basic_anchor(base_size=8, ratios=[0.5, 1, 2], anchor_scales=[8, 16, 32]):
    """[len(ratios)*len(anchor_scales), 4] fp32 (x_min, y_min, x_max, y_max) on the configured device
    (reference: utils/basic_anchors.py:11-23)."""
    return F.base_anchors(base_size, ratios, anchor_scales, device=config.get_device())


def enumerate_shifted_anchor(anchor_base, feat_stride, height, width):
    """[height*width*A, 4] fp32; location k = y*width + x, anchor k*A + a
    (reference: utils/basic_anchors.py:27-57)."""
    return F.shifted_anchors(anchor_base, feat_stride, height, width)
