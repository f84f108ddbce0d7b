"""two_stage_object_detection_b200 -- the proposal-and-RoI hot path of a two-stage detector as
hand-written sm_100a CUDA kernels behind the reference's Python module signatures.

    from two_stage_object_detection_b200.nets import RegionProposalNetwork, HarNetRoIHead, ...
    from two_stage_object_detection_b200.utils import loc2bbox, bbox_iou, ...
    from two_stage_object_detection_b200 import functional   # batched tensor-level ops

The CUDA library (libfrcnn_b200.so, C ABI in include/frcnn_b200.h) is built in-tree by
``python -m two_stage_object_detection_b200.build``; there is no CPU fallback.
"""
__version__ = "0.1.0"
