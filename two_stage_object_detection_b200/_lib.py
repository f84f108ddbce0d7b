"""ctypes binding of libfrcnn_b200.so (include/frcnn_b200.h).

There is no CPU or PyTorch fallback anywhere in this package: if the library is missing or a
tensor is not on a CUDA device the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FRCNN_B200_LIB") or os.path.join(_PKG, "libfrcnn_b200.so")
ABI_VERSION = 2
ERR_UNSUPPORTED = -4  # FRCNN_ERR_UNSUPPORTED
MAX_BASE_ANCHORS = 64

IMG_OK = 0
IMG_PAD_INDEX_ERROR = 1
IMG_SCATTER_INDEX_ERROR = 2


class AnchorSpec(C.Structure):
    _fields_ = [("anchors", C.c_void_p), ("base", C.c_void_p), ("num_base", C.c_int32),
                ("feat_stride", C.c_int32), ("height", C.c_int32), ("width", C.c_int32)]


class ProposalParams(C.Structure):
    _fields_ = [("batch", C.c_int32), ("num_anchors", C.c_int32), ("n_pre_nms", C.c_int32),
                ("n_post_nms", C.c_int32), ("clip_x_max", C.c_float), ("clip_y_max", C.c_float),
                ("min_size", C.c_float), ("nms_thresh", C.c_double), ("score_mode", C.c_int32),
                ("boxes_are_decoded", C.c_int32), ("nms_superblock", C.c_int32)]


class AnchorTargetParams(C.Structure):
    _fields_ = [("batch", C.c_int32), ("num_anchors", C.c_int32), ("max_gt", C.c_int32),
                ("n_sample", C.c_int32), ("pos_iou_thresh", C.c_float), ("neg_iou_thresh", C.c_float),
                ("n_pos", C.c_int32)]


class ProposalTargetParams(C.Structure):
    _fields_ = [("batch", C.c_int32), ("num_roi", C.c_int32), ("max_gt", C.c_int32),
                ("n_sample", C.c_int32), ("pos_per_image", C.c_int32), ("pos_iou_thresh", C.c_float),
                ("neg_iou_thresh_high", C.c_float), ("neg_iou_thresh_low", C.c_float)]


_P = C.c_void_p
_I = C.c_int32
_L = C.c_int64
_F = C.c_float
_D = C.c_double
_Z = C.c_size_t

# name -> (restype, argtypes); every symbol include/frcnn_b200.h declares
SIGNATURES = {
    "frcnn_abi_version": (_I, []),
    "frcnn_last_error": (C.c_char_p, []),
    "frcnn_device_info": (_I, [C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "frcnn_launch_count": (C.c_uint64, []),
    "frcnn_last_roi_kernel": (C.c_char_p, []),
    "frcnn_base_anchors": (_I, [C.POINTER(_F), C.POINTER(_F), _I, C.POINTER(_F), _I, _P, _P]),
    "frcnn_shifted_anchors": (_I, [_P, _I, _I, _I, _I, _P, _P]),
    "frcnn_loc2bbox": (_I, [_P, _P, _L, _I, _P, _P]),
    "frcnn_bbox2loc": (_I, [_P, _P, _L, _P, _P]),
    "frcnn_bbox_iou": (_I, [_P, _P, _L, _L, _P, _P]),
    "frcnn_proposals_workspace_bytes": (_Z, [C.POINTER(ProposalParams)]),
    "frcnn_proposals": (_I, [C.POINTER(ProposalParams), C.POINTER(AnchorSpec), _P, _P, _P, _P, _P, _P, _P,
                             _Z, _P]),
    "frcnn_decode_clip_score": (_I, [C.POINTER(ProposalParams), C.POINTER(AnchorSpec), _P, _P, _P, _P, _P,
                                     _P]),
    "frcnn_topk_workspace_bytes": (_Z, [_I, _I]),
    "frcnn_topk_sorted": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _Z, _P]),
    "frcnn_nms_sorted_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "frcnn_nms_sorted": (_I, [_P, _P, _I, _I, _D, _I, _I, _P, _P, _P, _Z, _P]),
    "frcnn_nms_workspace_bytes": (_Z, [_I]),
    "frcnn_nms": (_I, [_P, _P, _I, _D, _P, _P, _P, _Z, _P]),
    "frcnn_anchor_targets_workspace_bytes": (_Z, [C.POINTER(AnchorTargetParams)]),
    "frcnn_anchor_targets": (_I, [C.POINTER(AnchorTargetParams), C.POINTER(AnchorSpec), _P, _P, _P, _P, _P,
                                  _P, _Z, _P]),
    "frcnn_proposal_targets": (_I, [C.POINTER(ProposalTargetParams), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "frcnn_roi_head_coords": (_I, [_P, _P, _I, _I, _F, _F, _I, _I, _P, _P]),
    "frcnn_roi_workspace_bytes": (_Z, [_I, _I]),
    "frcnn_roi_pool_forward": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _F, _P, _P, _P, _Z, _P]),
    "frcnn_detection_decode": (_I, [_P, _P, _P, _P, C.c_int64, _I, _P, _P, _P, _P, _P]),
    "frcnn_nms_by_class": (_I, [_P, _P, _P, _P, _I, _I, C.c_double, _P, _P, _P]),
    "frcnn_roi_align_mean_workspace_bytes": (_Z, [_I, _I]),
    "frcnn_roi_align_mean_forward": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _F, _I, _I, _P, _P, _Z, _P]),
    "frcnn_roi_pool_mean_forward": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _F, _P, _P, _Z, _P]),
    "frcnn_roi_pool_backward": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "frcnn_roi_align_forward": (_I, [_P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _F, _I, _I, _I, _P, _P, _Z, _P]),
    "frcnn_roi_align_backward": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _I, _I, _P, _P]),
}

_lib = None
_lock = threading.Lock()


class FrcnnError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the library (built in-tree by two_stage_object_detection_b200.build); raises if absent."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise FrcnnError(
                    f"{LIB_PATH} is missing: build it with `python -m two_stage_object_detection_b200.build` "
                    "(this package has no CPU/PyTorch fallback)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
                fn.restype = res
                fn.argtypes = args
            if lib.frcnn_abi_version() != ABI_VERSION:
                raise FrcnnError("libfrcnn_b200.so ABI version mismatch; rebuild")
            _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().frcnn_last_error().decode("utf-8", "replace")
        raise FrcnnError(f"{what} failed (status {rc}): {msg}")


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise FrcnnError("two_stage_object_detection_b200 runs only on CUDA tensors "
                             "(no CPU fallback): got a tensor on " + str(t.device))
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise FrcnnError(f"tensors on different devices: {dev} vs {t.device}")
    if dev is None:
        raise FrcnnError("no CUDA tensor given")
    return dev


def ptr(t) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


_ws_cache: dict = {}


def workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    """Growable per-(device, stream) scratch buffer; reuse is safe because work is stream-ordered."""
    if torch.cuda.is_current_stream_capturing():
        # under CUDA-graph capture the buffer must belong to the graph being captured (its private memory pool):
        # a cached one may come from the pool of a graph that no longer exists
        return torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=dev)
    key = (dev.index, stream_ptr(dev))
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=dev)
        _ws_cache[key] = buf
    return buf


def f32c(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def launch_count() -> int:
    """Kernels this library has launched in this process (bench.py's gpu_launches)."""
    return int(load().frcnn_launch_count())


def last_roi_kernel() -> str:
    """Template instance the last RoI forward call picked (bench.py's roofline.kernel)."""
    return load().frcnn_last_roi_kernel().decode("utf-8", "replace")


def device_info():
    lib = load()
    sm, mj, mn = _I(), _I(), _I()
    check(lib.frcnn_device_info(C.byref(sm), C.byref(mj), C.byref(mn)), "frcnn_device_info")
    return sm.value, mj.value, mn.value
