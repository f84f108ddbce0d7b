"""Drop-in for the reference's utils/loc_bbox_iou.py; kernels in csrc/boxmath.cu."""
from __future__ import annotations

from typing import List

from .. import functional as F


def bbox_iou(bbox_a, bbox_b):
    """[n_a, n_b] IoU with the reference's +1e-8 denominator (utils/loc_bbox_iou.py:4-27).
    Raises IndexError when a box tensor does not have 4 columns, as the reference does."""
    return F.bbox_iou(bbox_a, bbox_b)


def loc2bbox(src_bbox, loc):
    """Apply (dx, dy, dw, dh) offsets to boxes (utils/loc_bbox_iou.py:29-61); loc may be [R, 4k]."""
    return F.loc2bbox(src_bbox, loc)


def bbox2loc(src_bbox, dst_bbox):
    """Offsets that take src boxes to dst boxes (utils/loc_bbox_iou.py:63-89)."""
    return F.bbox2loc(src_bbox, dst_bbox)


def xywh2xyxy(anchor: List[List]) -> List[List]:
    """(x, y, w, h) -> (x_min, y_min, x_max, y_max), in place (utils/loc_bbox_iou.py:91-97;
    dataset-preparation helper on Python lists, kept for import compatibility)."""
    anchor[2] += anchor[0]
    anchor[3] += anchor[1]
    return anchor
