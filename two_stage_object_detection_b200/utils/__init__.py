from .basic_anchors import enumerate_shifted_anchor, generate_basic_anchor  # noqa: F401
from .loc_bbox_iou import bbox2loc, bbox_iou, loc2bbox, xywh2xyxy  # noqa: F401
