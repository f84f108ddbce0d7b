"""Tensor-level (batched) entry points over the C ABI.  Every function takes CUDA tensors, allocates
its outputs with torch, enqueues kernels on the current CUDA stream and never synchronises with the
host.  The reference-named per-image wrappers live in ``utils/`` and ``nets/``.
"""
from __future__ import annotations

import ctypes as C
import logging
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
_log = logging.getLogger("two_stage_object_detection_b200")
_warned = set()


def _note_unfused(what: str, why: str) -> None:
    """A fused kernel did not cover this shape and the (still CUDA, still ours) two-step form ran instead:
    say so once per call site and shape class, never silently."""
    if what not in _warned:
        _warned.add(what)
        _log.warning("%s: fused kernel not used (%s); running the two-kernel form", what, why)


from ._lib import AnchorSpec, AnchorTargetParams, ProposalParams, ProposalTargetParams, check, f32c, ptr


def _spec(dev, anchor: Optional[torch.Tensor], base: Optional[torch.Tensor], feat_stride, feat_hw):
    """Build the anchor spec; returns (spec, keepalive)."""
    s = AnchorSpec()
    keep = []
    if anchor is not None:
        a = f32c(anchor).view(-1, 4)
        keep.append(a)
        s.anchors = a.data_ptr()
        s.num_base = 0
    else:
        if base is None or feat_hw is None or feat_stride is None:
            raise ValueError("give either `anchor` or (`base`, `feat_stride`, `feat_hw`)")
        b = f32c(base).view(-1, 4)
        keep.append(b)
        s.anchors = None
        s.base = b.data_ptr()
        s.num_base = b.shape[0]
        s.feat_stride = int(feat_stride)
        s.height, s.width = int(feat_hw[0]), int(feat_hw[1])
    return s, keep


# ------------------------------------------------------------------------------------------------
# anchors / box math
# ------------------------------------------------------------------------------------------------
def base_anchors(base_size=8, ratios=(0.5, 1, 2), anchor_scales=(8, 16, 32), device="cuda") -> torch.Tensor:
    lib = _lib.load()
    dev = torch.device(device)
    if dev.type != "cuda":
        raise _lib.FrcnnError("base_anchors: device must be a CUDA device (no CPU fallback)")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    nr, ns = len(ratios), len(anchor_scales)
    # host does exactly the Python-double part of utils/basic_anchors.py:16-17
    r = (C.c_float * nr)(*[float(np.float32(x)) for x in ratios])
    ir = (C.c_float * nr)(*[float(np.float32(1.0 / x)) for x in ratios])
    sz = (C.c_float * ns)(*[float(np.float32(base_size * s)) for s in anchor_scales])
    out = torch.empty((nr * ns, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.frcnn_base_anchors(r, ir, nr, sz, ns, out.data_ptr(), _lib.stream_ptr(dev)),
              "frcnn_base_anchors")
    return out


def shifted_anchors(anchor_base: torch.Tensor, feat_stride: int, height: int, width: int) -> torch.Tensor:
    lib = _lib.load()
    dev = _lib.require_cuda(anchor_base)
    b = f32c(anchor_base).view(-1, 4)
    out = torch.empty((int(height) * int(width) * b.shape[0], 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.frcnn_shifted_anchors(b.data_ptr(), b.shape[0], int(feat_stride), int(height), int(width),
                                        out.data_ptr(), _lib.stream_ptr(dev)), "frcnn_shifted_anchors")
    return out


def loc2bbox(src_bbox: torch.Tensor, loc: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    dev = _lib.require_cuda(src_bbox, loc)
    if src_bbox.shape[0] == 0:
        return torch.zeros((0, 4), dtype=torch.float32, device=dev)
    s, l = f32c(src_bbox), f32c(loc)
    if s.dim() != 2 or s.shape[1] != 4 or l.dim() != 2 or l.shape[1] % 4 or l.shape[0] != s.shape[0]:
        raise ValueError(f"loc2bbox: bad shapes {tuple(src_bbox.shape)} / {tuple(loc.shape)}")
    out = torch.empty_like(l)
    with torch.cuda.device(dev):
        check(lib.frcnn_loc2bbox(s.data_ptr(), l.data_ptr(), s.shape[0], l.shape[1] // 4, out.data_ptr(),
                                 _lib.stream_ptr(dev)), "frcnn_loc2bbox")
    return out


def bbox2loc(src_bbox: torch.Tensor, dst_bbox: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    dev = _lib.require_cuda(src_bbox, dst_bbox)
    s, d = f32c(src_bbox), f32c(dst_bbox)
    if s.dim() != 2 or s.shape[1] != 4 or s.shape != d.shape:
        raise ValueError(f"bbox2loc: bad shapes {tuple(src_bbox.shape)} / {tuple(dst_bbox.shape)}")
    out = torch.empty_like(s)
    with torch.cuda.device(dev):
        check(lib.frcnn_bbox2loc(s.data_ptr(), d.data_ptr(), s.shape[0], out.data_ptr(), _lib.stream_ptr(dev)),
              "frcnn_bbox2loc")
    return out


def bbox_iou(bbox_a: torch.Tensor, bbox_b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if bbox_a.shape[1] != 4 or bbox_b.shape[1] != 4:
        raise IndexError  # utils/loc_bbox_iou.py:14-16
    lib = _lib.load()
    dev = _lib.require_cuda(bbox_a, bbox_b)
    a, b = f32c(bbox_a), f32c(bbox_b)
    if out is None:
        out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=dev)
    elif (out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev
          or tuple(out.shape) != (a.shape[0], b.shape[0])):
        raise ValueError("bbox_iou: out must be a contiguous float32 [Na,Nb] tensor on the inputs' device")
    with torch.cuda.device(dev):
        check(lib.frcnn_bbox_iou(a.data_ptr(), b.data_ptr(), a.shape[0], b.shape[0], out.data_ptr(),
                                 _lib.stream_ptr(dev)), "frcnn_bbox_iou")
    return out


# ------------------------------------------------------------------------------------------------
# proposal layer
# ------------------------------------------------------------------------------------------------
def _proposal_params(B, N, n_pre_nms, n_post_nms, clip_x_max, clip_y_max, min_size, nms_iou, score_mode,
                     decoded, superblock) -> ProposalParams:
    p = ProposalParams()
    p.batch, p.num_anchors = int(B), int(N)
    p.n_pre_nms, p.n_post_nms = int(n_pre_nms), int(n_post_nms)
    p.clip_x_max, p.clip_y_max = float(clip_x_max), float(clip_y_max)
    p.min_size = float(np.float32(min_size))
    p.nms_thresh = float(nms_iou)
    p.score_mode = int(score_mode)
    p.boxes_are_decoded = int(decoded)
    p.nms_superblock = int(superblock)
    return p


def _conv_layout(dev, loc, score, anchor, base, feat_stride):
    """score_mode 2: loc [B,4A,H,W] / logits [B,2A,H,W] straight from the RPN's 1x1 convs (NCHW)."""
    l, s = f32c(loc), f32c(score)
    if l.dim() != 4 or s.dim() != 4 or l.shape[1] % 4 or s.shape[1] * 2 != l.shape[1] or \
            l.shape[0] != s.shape[0] or l.shape[2:] != s.shape[2:]:
        raise ValueError(f"proposals(layout='nchw'): loc [B,4A,H,W] and logits [B,2A,H,W] expected, got "
                         f"{tuple(loc.shape)} / {tuple(score.shape)}")
    B, A, H, W = l.shape[0], l.shape[1] // 4, l.shape[2], l.shape[3]
    spec, keep = _spec(dev, anchor, base, feat_stride, (H, W))
    spec.num_base, spec.height, spec.width = A, H, W  # the (A, H, W) factorisation of N, explicit anchors or not
    if base is not None and base.shape[0] != A:
        raise ValueError(f"proposals(layout='nchw'): {A} anchors per location in loc, {base.shape[0]} base anchors")
    return l, s, B, A * H * W, spec, keep


def proposals(loc: torch.Tensor, score: torch.Tensor, *, clip_x_max: float, clip_y_max: float,
              n_pre_nms: int, n_post_nms: int, nms_iou: float = 0.7, min_size: float = 16.0,
              anchor: Optional[torch.Tensor] = None, base: Optional[torch.Tensor] = None,
              feat_stride: Optional[int] = None, feat_hw: Optional[Sequence[int]] = None,
              score_is_logits: bool = False, boxes_are_decoded: bool = False, nms_superblock: int = 0,
              layout: str = "nhwc") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Batched ProposalCreator (nets/rpn.py:36-70 for every image of the batch in one pass).

    loc [B,N,4]; score [B,N] fg probabilities (or [B,N,2] logits with ``score_is_logits``).
    ``layout="nchw"``: loc [B,4A,H,W] and logits [B,2A,H,W] are the RPN conv outputs themselves, read in
    place (no permute / contiguous pass, nets/rpn.py:107-113); results are bit-identical.
    Returns (rois [B,n_post,4], roi_src [B,n_post] int32, n_keep [B] int32, status [B] int32); all on
    the device, nothing synchronises.  ``min_size`` is the already scaled value (min_size*scale).
    """
    lib = _lib.load()
    dev = _lib.require_cuda(loc, score)
    if layout == "nchw":
        if boxes_are_decoded:
            raise ValueError("proposals: layout='nchw' reads conv outputs, not decoded boxes")
        l, s, B, N, spec, keep = _conv_layout(dev, loc, score, anchor, base, feat_stride)
        mode = 2
    else:
        l, s = f32c(loc), f32c(score)
        if l.dim() != 3 or l.shape[2] != 4:
            raise ValueError(f"proposals: loc must be [B,N,4], got {tuple(loc.shape)}")
        B, N = l.shape[0], l.shape[1]
        want = (B, N, 2) if score_is_logits else (B, N)
        if tuple(s.shape) != want:
            raise ValueError(f"proposals: score must be {want}, got {tuple(score.shape)}")
        spec, keep = (AnchorSpec(), []) if boxes_are_decoded else _spec(dev, anchor, base, feat_stride, feat_hw)
        mode = 1 if score_is_logits else 0
    p = _proposal_params(B, N, n_pre_nms, n_post_nms, clip_x_max, clip_y_max, min_size, nms_iou,
                         mode, boxes_are_decoded, nms_superblock)
    rois = torch.empty((B, n_post_nms, 4), dtype=torch.float32, device=dev)
    src = torch.empty((B, n_post_nms), dtype=torch.int32, device=dev)
    n_keep = torch.empty((B,), dtype=torch.int32, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    nbytes = lib.frcnn_proposals_workspace_bytes(C.byref(p))
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, nbytes)
        check(lib.frcnn_proposals(C.byref(p), C.byref(spec), l.data_ptr(), s.data_ptr(), rois.data_ptr(),
                                  src.data_ptr(), n_keep.data_ptr(), status.data_ptr(), ws.data_ptr(),
                                  ws.numel(), _lib.stream_ptr(dev)), "frcnn_proposals")
    del keep
    return rois, src, n_keep, status


def decode_clip_score(loc, score, *, clip_x_max, clip_y_max, min_size=16.0, anchor=None, base=None,
                      feat_stride=None, feat_hw=None, score_is_logits=False, boxes_are_decoded=False,
                      layout="nhwc"):
    """Stage 1 only: (boxes [B,N,4] clipped, keys [B,N] uint32-as-int32, fg [B,N])."""
    lib = _lib.load()
    dev = _lib.require_cuda(loc, score)
    if layout == "nchw":
        l, s, B, N, spec, keep = _conv_layout(dev, loc, score, anchor, base, feat_stride)
        mode = 2
    else:
        l, s = f32c(loc), f32c(score)
        B, N = l.shape[0], l.shape[1]
        spec, keep = (AnchorSpec(), []) if boxes_are_decoded else _spec(dev, anchor, base, feat_stride, feat_hw)
        mode = 1 if score_is_logits else 0
    p = _proposal_params(B, N, 0, 0, clip_x_max, clip_y_max, min_size, 0.7, mode, boxes_are_decoded, 0)
    boxes = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
    keys = torch.empty((B, N), dtype=torch.int32, device=dev)
    fg = torch.empty((B, N), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.frcnn_decode_clip_score(C.byref(p), C.byref(spec), l.data_ptr(), s.data_ptr(),
                                          boxes.data_ptr(), keys.data_ptr(), fg.data_ptr(),
                                          _lib.stream_ptr(dev)), "frcnn_decode_clip_score")
    del keep
    return boxes, keys, fg


def topk_sorted(keys: torch.Tensor, boxes: Optional[torch.Tensor], k_cap: int):
    """Stage 2 only: (order [B,k_cap] int32, n_sel [B] int32, sorted_boxes [B,k_cap,4] or None)."""
    lib = _lib.load()
    dev = _lib.require_cuda(keys)
    k = keys.contiguous()
    assert k.dtype == torch.int32 and k.dim() == 2
    B, N = k.shape
    k_cap = int(min(k_cap, N)) if k_cap > 0 else N
    order = torch.empty((B, k_cap), dtype=torch.int32, device=dev)
    n_sel = torch.empty((B,), dtype=torch.int32, device=dev)
    bx = f32c(boxes) if boxes is not None else None
    sb = torch.empty((B, k_cap, 4), dtype=torch.float32, device=dev) if boxes is not None else None
    nbytes = lib.frcnn_topk_workspace_bytes(B, N)
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, nbytes)
        check(lib.frcnn_topk_sorted(k.data_ptr(), ptr(bx), B, N, k_cap, order.data_ptr(), n_sel.data_ptr(),
                                    ptr(sb), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)),
              "frcnn_topk_sorted")
    return order, n_sel, sb


def nms_sorted(sorted_boxes: torch.Tensor, n_sel: torch.Tensor, iou_threshold: float, keep_cap: int,
               superblock: int = 0):
    """Stage 3 only: greedy NMS over score-ordered boxes [B,R,4]; (keep [B,keep_cap] int32, n_keep [B])."""
    lib = _lib.load()
    dev = _lib.require_cuda(sorted_boxes, n_sel)
    sb = f32c(sorted_boxes)
    B, R = sb.shape[0], sb.shape[1]
    ns = n_sel.to(torch.int32).contiguous()
    keep = torch.full((B, keep_cap), -1, dtype=torch.int32, device=dev)
    n_keep = torch.zeros((B,), dtype=torch.int32, device=dev)
    nbytes = lib.frcnn_nms_sorted_workspace_bytes(B, R, keep_cap, superblock)
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, nbytes)
        check(lib.frcnn_nms_sorted(sb.data_ptr(), ns.data_ptr(), B, R, float(iou_threshold), int(keep_cap),
                                   int(superblock), keep.data_ptr(), n_keep.data_ptr(), ws.data_ptr(),
                                   ws.numel(), _lib.stream_ptr(dev)), "frcnn_nms_sorted")
    return keep, n_keep


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """torchvision.ops.nms equivalent (the op the reference imports at nets/rpn.py:7): kept original
    indices, int64, score-descending with stable ties.  Syncs once to size the result."""
    lib = _lib.load()
    dev = _lib.require_cuda(boxes, scores)
    b, s = f32c(boxes).view(-1, 4), f32c(scores).view(-1)
    n = b.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=dev)
    keep = torch.empty((n,), dtype=torch.int64, device=dev)
    n_keep = torch.zeros((1,), dtype=torch.int32, device=dev)
    nbytes = lib.frcnn_nms_workspace_bytes(n)
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, nbytes)
        check(lib.frcnn_nms(b.data_ptr(), s.data_ptr(), n, float(iou_threshold), keep.data_ptr(),
                            n_keep.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev)), "frcnn_nms")
    return keep[: int(n_keep.item())]


# ------------------------------------------------------------------------------------------------
# training targets
# ------------------------------------------------------------------------------------------------
def pad_gt(bboxes: Sequence[torch.Tensor], labels: Optional[Sequence[torch.Tensor]] = None, device=None):
    """Pack per-image GT lists into ([B,Gmax,4], [B,Gmax] int64, n_gt [B] int32)."""
    dev = device if device is not None else bboxes[0].device
    B = len(bboxes)
    g = max([int(b.shape[0]) for b in bboxes] + [1])
    bb = torch.zeros((B, g, 4), dtype=torch.float32, device=dev)
    ll = torch.zeros((B, g), dtype=torch.int64, device=dev)
    n = torch.tensor([int(b.shape[0]) for b in bboxes], dtype=torch.int32, device=dev)
    for i, b in enumerate(bboxes):
        if b.shape[0]:
            bb[i, : b.shape[0]] = b.to(dev).float().view(-1, 4)
            if labels is not None:
                ll[i, : b.shape[0]] = labels[i].to(dev).view(-1).long()
    return bb, ll, n


def anchor_targets(bbox: torch.Tensor, n_gt: torch.Tensor, *, anchor=None, base=None, feat_stride=None,
                   feat_hw=None, n_sample=256, pos_iou_thresh=0.7, neg_iou_thresh=0.3, pos_ratio=0.5,
                   return_argmax=False):
    """Batched AnchorTargetCreator.  bbox [B,Gmax,4], n_gt [B] int32 -> loc [B,N,4], label [B,N] int64."""
    lib = _lib.load()
    dev = _lib.require_cuda(bbox, n_gt)
    bb = f32c(bbox)
    B, G = bb.shape[0], bb.shape[1]
    spec, keep = _spec(dev, anchor, base, feat_stride, feat_hw)
    N = keep[0].shape[0] if anchor is not None else spec.num_base * spec.height * spec.width
    p = AnchorTargetParams()
    p.batch, p.num_anchors, p.max_gt = B, int(N), G
    p.n_sample = int(n_sample)
    p.pos_iou_thresh = float(np.float32(pos_iou_thresh))
    p.neg_iou_thresh = float(np.float32(neg_iou_thresh))
    p.n_pos = int(pos_ratio * n_sample)
    loc = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
    label = torch.empty((B, N), dtype=torch.int64, device=dev)
    argmax = torch.empty((B, N), dtype=torch.int32, device=dev) if return_argmax else None
    ng = n_gt.to(torch.int32).contiguous()
    nbytes = lib.frcnn_anchor_targets_workspace_bytes(C.byref(p))
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, nbytes)
        check(lib.frcnn_anchor_targets(C.byref(p), C.byref(spec), bb.data_ptr(), ng.data_ptr(), loc.data_ptr(),
                                       label.data_ptr(), ptr(argmax), ws.data_ptr(), ws.numel(),
                                       _lib.stream_ptr(dev)), "frcnn_anchor_targets")
    del keep
    return (loc, label, argmax) if return_argmax else (loc, label)


def proposal_targets(roi: torch.Tensor, bbox: torch.Tensor, label: torch.Tensor, n_gt: torch.Tensor, *,
                     n_sample=128, pos_ratio=0.5, pos_iou_thresh=0.5, neg_iou_thresh_high=0.5,
                     neg_iou_thresh_low=0.0):
    """Batched ProposalTargetCreator.  roi [B,R,4], bbox [B,Gmax,4], label [B,Gmax] int64, n_gt [B].
    Returns (sample_roi [B,S,4], gt_loc [B,S,4], gt_label [B,S] int64, n_out [B], status [B])."""
    lib = _lib.load()
    dev = _lib.require_cuda(roi, bbox, label, n_gt)
    r, bb = f32c(roi), f32c(bbox)
    ll = label.detach().to(torch.int64).contiguous()
    B, R, G = r.shape[0], r.shape[1], bb.shape[1]
    p = ProposalTargetParams()
    p.batch, p.num_roi, p.max_gt = B, R, G
    p.n_sample = int(n_sample)
    p.pos_per_image = int(n_sample * pos_ratio)
    p.pos_iou_thresh = float(np.float32(pos_iou_thresh))
    p.neg_iou_thresh_high = float(np.float32(neg_iou_thresh_high))
    p.neg_iou_thresh_low = float(np.float32(neg_iou_thresh_low))
    S = int(n_sample)
    sample = torch.empty((B, S, 4), dtype=torch.float32, device=dev)
    gt_loc = torch.empty((B, S, 4), dtype=torch.float32, device=dev)
    gt_label = torch.empty((B, S), dtype=torch.int64, device=dev)
    n_out = torch.empty((B,), dtype=torch.int32, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    ng = n_gt.to(torch.int32).contiguous()
    with torch.cuda.device(dev):
        check(lib.frcnn_proposal_targets(C.byref(p), r.data_ptr(), bb.data_ptr(), ll.data_ptr(), ng.data_ptr(),
                                         sample.data_ptr(), gt_loc.data_ptr(), gt_label.data_ptr(),
                                         n_out.data_ptr(), status.data_ptr(), _lib.stream_ptr(dev)),
              "frcnn_proposal_targets")
    return sample, gt_loc, gt_label, n_out, status


# ------------------------------------------------------------------------------------------------
# RoI head gather
# ------------------------------------------------------------------------------------------------
def roi_head_coords(rois: torch.Tensor, roi_indices: torch.Tensor, img_size, feat_hw) -> torch.Tensor:
    """nets/classify.py:29-38: rois [n,R,4] + roi_indices [n] -> [n*R,5] feature-map RoIs."""
    lib = _lib.load()
    dev = _lib.require_cuda(rois, roi_indices)
    r = f32c(rois)
    n, R = r.shape[0], r.shape[1]
    idx = roi_indices.detach().to(torch.int32).contiguous().view(-1)
    out = torch.empty((n * R, 5), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.frcnn_roi_head_coords(r.data_ptr(), idx.data_ptr(), n, R, float(img_size[0]),
                                        float(img_size[1]), int(feat_hw[0]), int(feat_hw[1]), out.data_ptr(),
                                        _lib.stream_ptr(dev)), "frcnn_roi_head_coords")
    return out


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


def roi_pool_forward(feat, rois5, output_size, spatial_scale=1.0, with_argmax=False, out=None, rois_per_image=0):
    lib = _lib.load()
    dev = _lib.require_cuda(feat, rois5)
    f, r = f32c(feat), f32c(rois5).view(-1, 5)
    B, Cc, H, W = f.shape
    K = r.shape[0]
    ph, pw = _pair(output_size)
    if out is None:
        out = torch.empty((K, Cc, ph, pw), dtype=torch.float32, device=dev)
    am = torch.empty((K, Cc, ph, pw), dtype=torch.int32, device=dev) if with_argmax else None
    nbytes = lib.frcnn_roi_workspace_bytes(B, K)
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, nbytes)
        check(lib.frcnn_roi_pool_forward(f.data_ptr(), B, Cc, H, W, r.data_ptr(), K, int(rois_per_image), ph, pw,
                                         float(spatial_scale), out.data_ptr(), ptr(am), ws.data_ptr(),
                                         ws.numel(), _lib.stream_ptr(dev)), "frcnn_roi_pool_forward")
    return (out, am) if with_argmax else out


def roi_pool_mean(feat, rois5, output_size, spatial_scale=1.0, rois_per_image=0):
    """mean over the bins of roi_pool(feat, rois5) -> [K,C], in one kernel (the HarDNet head's
    RoIPool + AdaptiveAvgPool2d(1) + Flatten, nets/classify.py:43-46 with models/hardnet.py:203-212).
    Inference only (no autograd).  Shapes the fused kernel does not cover take the two steps."""
    lib = _lib.load()
    dev = _lib.require_cuda(feat, rois5)
    f, r = f32c(feat), f32c(rois5).view(-1, 5)
    B, Cc, H, W = f.shape
    K = r.shape[0]
    ph, pw = _pair(output_size)
    out = torch.empty((K, Cc), dtype=torch.float32, device=dev)
    nbytes = lib.frcnn_roi_workspace_bytes(B, K)
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, nbytes)
        rc = lib.frcnn_roi_pool_mean_forward(f.data_ptr(), B, Cc, H, W, r.data_ptr(), K, int(rois_per_image), ph, pw,
                                             float(spatial_scale), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                             _lib.stream_ptr(dev))
    if rc == _lib.ERR_UNSUPPORTED:
        _note_unfused("roi_pool_mean", lib.frcnn_last_error().decode("utf-8", "replace"))
        return roi_pool_forward(f, r, output_size, spatial_scale, rois_per_image=rois_per_image).mean((2, 3))
    check(rc, "frcnn_roi_pool_mean_forward")
    return out


def roi_align_mean(feat, rois5, output_size, spatial_scale=1.0, sampling_ratio=-1, aligned=False, rois_per_image=0):
    """mean over the bins of roi_align(feat, rois5) -> [K,C] (RoIAlign head + AdaptiveAvgPool2d(1) + Flatten) as a
    separable weighted window sum; inference only.  Uncovered shapes (adaptive sampling grid, maps wider than
    64 pixels) take the two steps."""
    lib = _lib.load()
    dev = _lib.require_cuda(feat, rois5)
    f, r = f32c(feat), f32c(rois5).view(-1, 5)
    B, Cc, H, W = f.shape
    K = r.shape[0]
    ph, pw = _pair(output_size)
    out = torch.empty((K, Cc), dtype=torch.float32, device=dev)
    nbytes = lib.frcnn_roi_align_mean_workspace_bytes(B, K)
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, nbytes)
        rc = lib.frcnn_roi_align_mean_forward(f.data_ptr(), B, Cc, H, W, r.data_ptr(), K, int(rois_per_image), ph, pw,
                                              float(spatial_scale), int(sampling_ratio), int(bool(aligned)),
                                              out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    if rc == _lib.ERR_UNSUPPORTED:
        _note_unfused("roi_align_mean", lib.frcnn_last_error().decode("utf-8", "replace"))
        return roi_align_forward(f, r, output_size, spatial_scale, sampling_ratio, aligned,
                                 rois_per_image=rois_per_image, exact=True).mean((2, 3))
    check(rc, "frcnn_roi_align_mean_forward")
    return out


def roi_align_forward(feat, rois5, output_size, spatial_scale=1.0, sampling_ratio=-1, aligned=False, out=None,
                      rois_per_image=0, exact=False):
    """torchvision.ops.roi_align forward.  ``exact=True``: the reference's operation order, bit-identical to
    torchvision's CPU kernel.  Default: the fast kernel where one exists (sampling_ratio 2, 7x7 / 14x14 bins;
    FMA + merged separable weights, within 1e-5 of the largest tap magnitude), the exact ones elsewhere."""
    lib = _lib.load()
    dev = _lib.require_cuda(feat, rois5)
    f, r = f32c(feat), f32c(rois5).view(-1, 5)
    B, Cc, H, W = f.shape
    K = r.shape[0]
    ph, pw = _pair(output_size)
    if out is None:
        out = torch.empty((K, Cc, ph, pw), dtype=torch.float32, device=dev)
    nbytes = lib.frcnn_roi_workspace_bytes(B, K)
    with torch.cuda.device(dev):
        ws = _lib.workspace(dev, nbytes)
        check(lib.frcnn_roi_align_forward(f.data_ptr(), B, Cc, H, W, r.data_ptr(), K, int(rois_per_image), ph, pw,
                                          float(spatial_scale), int(sampling_ratio), int(bool(aligned)),
                                          int(bool(exact)), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                          _lib.stream_ptr(dev)),
              "frcnn_roi_align_forward")
    return out


class _RoIPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, rois5, output_size, spatial_scale, rois_per_image):
        need = feat.requires_grad
        res = roi_pool_forward(feat, rois5, output_size, spatial_scale, with_argmax=need,
                               rois_per_image=rois_per_image)
        if need:
            out, am = res
            ctx.save_for_backward(am, f32c(rois5).view(-1, 5))
            ctx.shape = tuple(feat.shape)
            ctx.ps = _pair(output_size)
            return out
        return res

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        am, r = ctx.saved_tensors
        B, Cc, H, W = ctx.shape
        dev = grad_out.device
        go = f32c(grad_out)
        gi = torch.zeros((B, Cc, H, W), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.frcnn_roi_pool_backward(go.data_ptr(), am.data_ptr(), r.data_ptr(), r.shape[0], B, Cc, H, W,
                                              ctx.ps[0], ctx.ps[1], gi.data_ptr(), _lib.stream_ptr(dev)),
                  "frcnn_roi_pool_backward")
        return gi, None, None, None, None


class _RoIAlignFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, rois5, output_size, spatial_scale, sampling_ratio, aligned, rois_per_image, exact):
        out = roi_align_forward(feat, rois5, output_size, spatial_scale, sampling_ratio, aligned,
                                rois_per_image=rois_per_image, exact=exact)
        if feat.requires_grad:
            ctx.save_for_backward(f32c(rois5).view(-1, 5))
            ctx.cfg = (tuple(feat.shape), _pair(output_size), float(spatial_scale), int(sampling_ratio),
                       int(bool(aligned)))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        (r,) = ctx.saved_tensors
        (B, Cc, H, W), (ph, pw), scale, sr, al = ctx.cfg
        dev = grad_out.device
        go = f32c(grad_out)
        gi = torch.zeros((B, Cc, H, W), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(lib.frcnn_roi_align_backward(go.data_ptr(), r.data_ptr(), r.shape[0], B, Cc, H, W, ph, pw, scale,
                                               sr, al, gi.data_ptr(), _lib.stream_ptr(dev)),
                  "frcnn_roi_align_backward")
        return gi, None, None, None, None, None, None, None


def roi_pool(feat, rois5, output_size, spatial_scale=1.0, rois_per_image=0):
    """torchvision.ops.roi_pool equivalent with autograd w.r.t. ``feat``.  ``rois_per_image=R`` promises
    that rows [b*R,(b+1)*R) of rois5 belong to image b (skips the device-side bucketing pass)."""
    if feat.requires_grad and torch.is_grad_enabled():
        return _RoIPoolFn.apply(feat, rois5, output_size, spatial_scale, rois_per_image)
    return roi_pool_forward(feat, rois5, output_size, spatial_scale, rois_per_image=rois_per_image)


def roi_align(feat, rois5, output_size, spatial_scale=1.0, sampling_ratio=-1, aligned=False, rois_per_image=0,
              exact=False):
    """torchvision.ops.roi_align equivalent with autograd w.r.t. ``feat`` (``exact``: see roi_align_forward)."""
    if feat.requires_grad and torch.is_grad_enabled():
        return _RoIAlignFn.apply(feat, rois5, output_size, spatial_scale, sampling_ratio, aligned, rois_per_image,
                                 exact)
    return roi_align_forward(feat, rois5, output_size, spatial_scale, sampling_ratio, aligned,
                             rois_per_image=rois_per_image, exact=exact)


# ------------------------------------------------------------------------------------------------
# after the head: detections
# ------------------------------------------------------------------------------------------------
def detection_decode(roi: torch.Tensor, roi_cls_loc: torch.Tensor, roi_score: torch.Tensor,
                     label: Optional[torch.Tensor] = None, check_labels: bool = False):
    """nets/frcnn_training.py:311-320 for any leading shape: roi [...,4], roi_cls_loc [...,4C],
    roi_score [...,C], label [...] int64 or None (None: decode with the best class's loc row).
    Returns (boxes [...,4], cls_score [...], cls_index [...] int64).  ``check_labels=True`` synchronises
    and raises IndexError for a label outside [0,C), as the reference's gather does."""
    lib = _lib.load()
    dev = _lib.require_cuda(roi, roi_cls_loc, roi_score)
    Cn = roi_score.shape[-1]
    lead = tuple(roi_score.shape[:-1])
    r, cl, sc = f32c(roi).view(-1, 4), f32c(roi_cls_loc).view(-1, 4 * Cn), f32c(roi_score).view(-1, Cn)
    T = sc.shape[0]
    if r.shape[0] != T or cl.shape[0] != T:
        raise ValueError("detection_decode: roi / roi_cls_loc / roi_score disagree on the number of rows")
    lab = None if label is None else label.detach().to(torch.int64).contiguous().view(-1)
    boxes = torch.empty((T, 4), dtype=torch.float32, device=dev)
    cs = torch.empty((T,), dtype=torch.float32, device=dev)
    ci = torch.empty((T,), dtype=torch.int64, device=dev)
    bad = torch.empty((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.frcnn_detection_decode(r.data_ptr(), cl.data_ptr(), sc.data_ptr(), ptr(lab), T, int(Cn),
                                         boxes.data_ptr(), cs.data_ptr(), ci.data_ptr(), bad.data_ptr(),
                                         _lib.stream_ptr(dev)), "frcnn_detection_decode")
    if check_labels and lab is not None and int(bad.item()):
        raise IndexError("detection_decode: label outside [0, n_class)")
    return boxes.view(*lead, 4), cs.view(lead), ci.view(lead)


def nms_by_class(boxes: torch.Tensor, scores: torch.Tensor, classes: Optional[torch.Tensor], iou_threshold: float,
                 n_valid: Optional[torch.Tensor] = None):
    """The evaluator's per-class NMS (nets/frcnn_training.py:441-454), batched: boxes [B,R,4], scores [B,R],
    classes [B,R] int64 or None (class-agnostic, multi_inference.py:84), n_valid [B] or None.
    Returns (keep [B,R] int32: kept row indices by (score desc, index asc), -1 padded; n_keep [B])."""
    lib = _lib.load()
    dev = _lib.require_cuda(boxes, scores)
    b, s = f32c(boxes), f32c(scores)
    if b.dim() != 3 or s.shape != b.shape[:2]:
        raise ValueError("nms_by_class expects boxes [B,R,4] and scores [B,R]")
    B, R = s.shape
    cl = None if classes is None else classes.detach().to(torch.int64).contiguous()
    nv = None if n_valid is None else n_valid.to(torch.int32).contiguous()
    keep = torch.empty((B, R), dtype=torch.int32, device=dev)
    n_keep = torch.empty((B,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.frcnn_nms_by_class(b.data_ptr(), s.data_ptr(), ptr(cl), ptr(nv), B, R, float(iou_threshold),
                                    keep.data_ptr(), n_keep.data_ptr(), _lib.stream_ptr(dev))
    if rc != _lib.ERR_UNSUPPORTED:
        check(rc, "frcnn_nms_by_class")
        return keep, n_keep
    # more than 1024 rows per image: the evaluator's own loop (one frcnn_nms per image and class; synchronises)
    _note_unfused("nms_by_class", f"{R} rows per image > 1024")
    keep.fill_(-1)
    for i in range(B):
        n = R if nv is None else int(nv[i])
        ci = torch.zeros(n, dtype=torch.int64, device=dev) if cl is None else cl[i, :n]
        parts = [idx[nms(b[i, idx], s[i, idx], iou_threshold)] for idx in
                 (torch.nonzero(ci == c).flatten() for c in torch.unique(ci).tolist())]
        kept = torch.cat(parts) if parts else torch.zeros(0, dtype=torch.int64, device=dev)
        # merge the per-class lists into (score desc, index asc) order: stable sort of the kept rows by
        # index, then by descending score
        kept = kept.sort().values
        kept = kept[torch.sort(s[i, kept], descending=True, stable=True).indices]
        keep[i, :kept.numel()] = kept.to(torch.int32)
        n_keep[i] = kept.numel()
    return keep, n_keep
