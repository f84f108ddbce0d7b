"""oracle/ref_port.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy fp32 + the C file next to it) of the reference's proposal-and-RoI path.
Nothing in the product package may import this module: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs do, and
only as the checker / the CPU arm.

Every function cites the reference lines it restates (paths relative to /root/reference).
All +,-,*,/ are single fp32 roundings exactly as the reference's ATen CPU kernels do them, so
integer/index outputs are bit-identical; exp()/log() come from numpy instead of SLEEF, so
``decode``/``encode`` agree with the reference to ~1 ulp (tolerance 1e-5 relative, stated in
the tests).

Parity pin: ``tests/test_oracle_golden.py`` checks every function below against
``tests/golden/*.npz`` -- outputs of the real reference callables (and torchvision 0.26.0 CPU
kernels) produced in the build container by ``tests/golden/make_golden.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

import numpy as np

F32 = np.float32
_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lock = threading.Lock()
_lib = None


class OracleIndexError(IndexError):
    """Raised where the reference itself raises IndexError (pad / label-scatter quirks)."""


# --------------------------------------------------------------------------------------------
# C library
# --------------------------------------------------------------------------------------------
def build(force: bool = False) -> str:
    """Compile oracle/frcnn_oracle.c -> oracle/_build/liboracle.so (gcc, no FMA contraction)."""
    src = os.path.join(_HERE, "frcnn_oracle.c")
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
           "-o", _SO, src, "-lm"]
    subprocess.run(cmd, check=True)
    return _SO


def lib():
    global _lib
    with _lock:
        if _lib is None:
            build()
            L = ctypes.CDLL(_SO)
            i64, f32p, i64p, i32p = (ctypes.c_int64, ctypes.POINTER(ctypes.c_float),
                                     ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32))
            L.oracle_nms.restype = i64
            L.oracle_nms.argtypes = [f32p, f32p, i64, ctypes.c_double, i64p]
            L.oracle_argsort_desc_stable.restype = None
            L.oracle_argsort_desc_stable.argtypes = [f32p, i64, i64p]
            L.oracle_proposal_layer.restype = ctypes.c_int
            L.oracle_proposal_layer.argtypes = [f32p, f32p, i64, ctypes.c_float, ctypes.c_float,
                                                ctypes.c_float, i64, i64, ctypes.c_double, f32p,
                                                i64p, i64p]
            L.oracle_proposal_layer_batch.restype = None
            L.oracle_proposal_layer_batch.argtypes = [f32p, f32p, i64, i64, ctypes.c_float,
                                                      ctypes.c_float, ctypes.c_float, i64, i64,
                                                      ctypes.c_double, f32p, i64p, i64p, i32p]
            L.oracle_roi_pool.restype = None
            L.oracle_roi_pool.argtypes = [f32p, i64, i64, i64, i64, f32p, i64, i64, i64,
                                          ctypes.c_float, f32p, i32p]
            L.oracle_roi_align.restype = None
            L.oracle_roi_align.argtypes = [f32p, i64, i64, i64, i64, f32p, i64, i64, i64,
                                           ctypes.c_float, ctypes.c_int, ctypes.c_int, f32p]
            L.oracle_roi_pool_backward.restype = None
            L.oracle_roi_pool_backward.argtypes = [f32p, i32p, f32p, i64, i64, i64, i64, i64, i64, i64, f32p]
            L.oracle_roi_align_backward.restype = None
            L.oracle_roi_align_backward.argtypes = [f32p, f32p, i64, i64, i64, i64, i64, i64, i64,
                                                    ctypes.c_float, ctypes.c_int, ctypes.c_int, f32p]
            L.oracle_max_threads.restype = ctypes.c_int
            L.oracle_set_threads.restype = None
            L.oracle_set_threads.argtypes = [ctypes.c_int]
            _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=F32)


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def set_threads(n: int) -> int:
    """OpenMP threads of the C oracle (bench.py's CPU arm: all host cores, whatever OMP_NUM_THREADS a
    launcher exported).  Returns the count now in effect."""
    lib().oracle_set_threads(int(n))
    return max_threads()


# --------------------------------------------------------------------------------------------
# a1 / a2: anchors
# --------------------------------------------------------------------------------------------
def base_anchors(base_size=8, ratios=(0.5, 1, 2), anchor_scales=(8, 16, 32)) -> np.ndarray:
    """utils/basic_anchors.py:11-23.  Row i_ratio*S + j_scale = (-w/2, -h/2, w/2, h/2) with
    h = size*sqrt(r), w = size*sqrt(1/r), all in fp32 (size = base_size*scale, host product)."""
    out = np.zeros((len(ratios) * len(anchor_scales), 4), dtype=F32)
    for i, r in enumerate(ratios):
        sr = np.sqrt(F32(r))
        sir = np.sqrt(F32(1.0 / r))
        for j, s in enumerate(anchor_scales):
            size = F32(base_size * s)
            h = F32(size * sr)
            w = F32(size * sir)
            out[i * len(anchor_scales) + j] = (-w / F32(2), -h / F32(2), w / F32(2), h / F32(2))
    return out


def shifted_anchors(anchor_base, feat_stride, height, width) -> np.ndarray:
    """utils/basic_anchors.py:27-57.  Location k = y*W + x; anchor[k*A + a] = base[a] + shift."""
    base = _f32(anchor_base)
    xs = (np.arange(width, dtype=np.int64) * int(feat_stride)).astype(F32)
    ys = (np.arange(height, dtype=np.int64) * int(feat_stride)).astype(F32)
    sx = np.tile(xs, height)
    sy = np.repeat(ys, width)
    shift = np.stack([sx, sy, sx, sy], axis=1)  # [K,4]
    return (base[None, :, :] + shift[:, None, :]).reshape(-1, 4).astype(F32)


# --------------------------------------------------------------------------------------------
# a3 / a4 / a5: box math
# --------------------------------------------------------------------------------------------
def decode(src_bbox, loc) -> np.ndarray:
    """loc2bbox, utils/loc_bbox_iou.py:29-61.  loc may be [R, 4k] (strided 0::4 groups)."""
    src = _f32(src_bbox)
    loc = _f32(loc)
    if src.shape[0] == 0:
        return np.zeros((0, 4), dtype=F32)
    w = (src[:, 2] - src[:, 0])[:, None]
    h = (src[:, 3] - src[:, 1])[:, None]
    cx = src[:, 0][:, None] + F32(0.5) * w
    cy = src[:, 1][:, None] + F32(0.5) * h
    dx, dy, dw, dh = loc[:, 0::4], loc[:, 1::4], loc[:, 2::4], loc[:, 3::4]
    ncx = dx * w + cx
    ncy = dy * h + cy
    nw = np.exp(dw) * w
    nh = np.exp(dh) * h
    out = np.zeros_like(loc)
    out[:, 0::4] = ncx - F32(0.5) * nw
    out[:, 1::4] = ncy - F32(0.5) * nh
    out[:, 2::4] = ncx + F32(0.5) * nw
    out[:, 3::4] = ncy + F32(0.5) * nh
    return out


def encode(src_bbox, dst_bbox) -> np.ndarray:
    """bbox2loc, utils/loc_bbox_iou.py:63-89."""
    s = _f32(src_bbox).reshape(-1, 4)
    d = _f32(dst_bbox).reshape(-1, 4)
    w = s[:, 2] - s[:, 0]
    h = s[:, 3] - s[:, 1]
    cx = s[:, 0] + F32(0.5) * w
    cy = s[:, 1] + F32(0.5) * h
    bw = d[:, 2] - d[:, 0]
    bh = d[:, 3] - d[:, 1]
    bcx = d[:, 0] + F32(0.5) * bw
    bcy = d[:, 1] + F32(0.5) * bh
    eps = np.finfo(F32).eps
    w = np.maximum(w, eps)
    h = np.maximum(h, eps)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.stack([(bcx - cx) / w, (bcy - cy) / h, np.log(bw / w), np.log(bh / h)], axis=1)
    return out.astype(F32)


def iou(bbox_a, bbox_b) -> np.ndarray:
    """bbox_iou, utils/loc_bbox_iou.py:4-27: inter / (((area_a + area_b) - inter) + 1e-8f)."""
    a = _f32(bbox_a)
    b = _f32(bbox_b)
    if a.ndim != 2 or b.ndim != 2 or a.shape[1] != 4 or b.shape[1] != 4:
        raise IndexError("bbox_iou expects [n,4] boxes")
    tl = np.maximum(a[:, None, :2], b[None, :, :2])
    br = np.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = np.maximum(br - tl, F32(0))
    inter = wh[..., 0] * wh[..., 1]
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / (((area_a[:, None] + area_b[None, :]) - inter) + F32(1e-8))).astype(F32)


def fg_scores(logits) -> np.ndarray:
    """nets/rpn.py:115-118: softmax over the last (size-2) dim, foreground column."""
    x = _f32(logits)
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return (e[..., 1] / (e[..., 0] + e[..., 1])).astype(F32)


# --------------------------------------------------------------------------------------------
# a8: NMS,  a7: proposal layer
# --------------------------------------------------------------------------------------------
def argsort_desc_stable(scores) -> np.ndarray:
    s = _f32(scores).reshape(-1)
    out = np.empty(s.shape[0], dtype=np.int64)
    lib().oracle_argsort_desc_stable(_p(s, ctypes.c_float), s.shape[0], _p(out, ctypes.c_int64))
    return out


def nms(boxes, scores, iou_threshold) -> np.ndarray:
    """torchvision.ops.nms as called at nets/rpn.py:63."""
    b = _f32(boxes).reshape(-1, 4)
    s = _f32(scores).reshape(-1)
    keep = np.empty(max(b.shape[0], 1), dtype=np.int64)
    k = lib().oracle_nms(_p(b, ctypes.c_float), _p(s, ctypes.c_float), b.shape[0],
                         float(iou_threshold), _p(keep, ctypes.c_int64))
    return keep[:k].copy()


def detection_decode(roi, roi_cls_loc, roi_score, label=None):
    """Post-head decode, nets/frcnn_training.py:311-320: class-specific loc rows gathered by `label`
    (the reference passes gt_roi_label; None = the predicted class, plain inference), loc2bbox, and
    torch.max over the class scores (first index on ties, NaN wins).
    roi [R,4], roi_cls_loc [R,4C], roi_score [R,C] -> (boxes [R,4], cls_score [R], cls_index [R] int64)."""
    r = _f32(roi)
    sc = _f32(roi_score)
    R, C = sc.shape
    loc = _f32(roi_cls_loc).reshape(R, C, 4)
    nan = np.isnan(sc)
    cls_index = np.where(nan.any(1), nan.argmax(1), sc.argmax(1)).astype(np.int64)
    cls_score = sc[np.arange(R), cls_index]
    pick = cls_index if label is None else np.asarray(label).astype(np.int64)
    boxes = decode(r, loc[np.arange(R), pick]) if R else np.zeros((0, 4), F32)
    return boxes, cls_score, cls_index


def nms_by_class(boxes, scores, classes, iou_threshold):
    """The evaluator's per-class NMS, nets/frcnn_training.py:441-454: for every class c, torchvision nms on
    the rows with classes == c.  Returns the kept ORIGINAL row indices of all classes merged, ordered by
    (score desc, index asc) -- the rows of class c, in that order, are the reference's `keep` for c.
    classes=None: one class (multi_inference.py:84)."""
    b = _f32(boxes)
    s = _f32(scores)
    n = b.shape[0]
    cl = np.zeros(n, np.int64) if classes is None else np.asarray(classes).astype(np.int64)
    kept = []
    for c in np.unique(cl):
        idx = np.nonzero(cl == c)[0]
        kept.append(idx[nms(b[idx], s[idx], iou_threshold)])
    kept = np.concatenate(kept) if kept else np.zeros(0, np.int64)
    order = argsort_desc_stable(s)
    rank = np.empty(n, np.int64)
    rank[order] = np.arange(n)
    return kept[np.argsort(rank[kept], kind="stable")].astype(np.int64)


def proposal_limits(mode, n_train_pre_nms=12000, n_train_post_nms=600, n_test_pre_nms=3000,
                    n_test_post_nms=300):
    """nets/rpn.py:37-42: only the exact string "train" selects the train limits."""
    if mode == "train":
        return n_train_pre_nms, n_train_post_nms
    return n_test_pre_nms, n_test_post_nms


def clip_filter(roi, img_size, min_size):
    """nets/rpn.py:47-54.  x clamped to [0,img_size[1]], y to [0,img_size[2]] (as written in the
    reference, whose callers pass (3,H,W)).  Returns (clipped [N,4], keep_index)."""
    r = _f32(roi).copy()
    xmax, ymax = F32(img_size[1]), F32(img_size[2])
    r[:, [0, 2]] = np.minimum(np.maximum(r[:, [0, 2]], F32(0)), xmax)
    r[:, [1, 3]] = np.minimum(np.maximum(r[:, [1, 3]], F32(0)), ymax)
    ms = F32(min_size)
    keep = np.where(((r[:, 2] - r[:, 0]) >= ms) & ((r[:, 3] - r[:, 1]) >= ms))[0]
    return r, keep


def proposal_layer_from_boxes(decoded, score, img_size, scale=1.0, nms_iou=0.7, n_pre_nms=12000,
                              n_post_nms=600, min_size=16, return_extra=False):
    """nets/rpn.py:47-69 on already-decoded boxes (exp-free, hence bit-exact vs the reference)."""
    d = _f32(decoded).reshape(-1, 4)
    s = _f32(score).reshape(-1)
    n = d.shape[0]
    out = np.zeros((n_post_nms, 4), dtype=F32)
    src = np.full((n_post_nms,), -1, dtype=np.int64)
    nk = np.zeros(1, dtype=np.int64)
    rc = lib().oracle_proposal_layer(_p(d, ctypes.c_float), _p(s, ctypes.c_float), n,
                                     float(img_size[1]), float(img_size[2]),
                                     float(F32(min_size * scale)), int(n_pre_nms), int(n_post_nms),
                                     float(nms_iou), _p(out, ctypes.c_float),
                                     _p(src, ctypes.c_int64), _p(nk, ctypes.c_int64))
    if rc != 0:
        raise OracleIndexError("pad index beyond the pre-NMS row count (nets/rpn.py:65-69)")
    if return_extra:
        return out, src, int(nk[0])
    return out


def proposal_layer(loc, score, anchor, img_size, scale=1.0, mode="train", **kw):
    """ProposalCreator.__call__, nets/rpn.py:36-70."""
    limits = {k: kw.pop(k) for k in list(kw) if k.startswith("n_t")}
    n_pre, n_post = proposal_limits(mode, **limits)
    return proposal_layer_from_boxes(decode(anchor, loc), score, img_size, scale=scale,
                                     n_pre_nms=n_pre, n_post_nms=n_post, **kw)


def proposal_layer_batch_from_boxes(decoded, score, img_size, scale, nms_iou, n_pre_nms,
                                    n_post_nms, min_size):
    """OpenMP batch of independent images (cpu_baseline leg).  decoded [B,N,4], score [B,N]."""
    d = _f32(decoded)
    s = _f32(score)
    B, N = s.shape
    out = np.zeros((B, n_post_nms, 4), dtype=F32)
    src = np.full((B, n_post_nms), -1, dtype=np.int64)
    nk = np.zeros(B, dtype=np.int64)
    rc = np.zeros(B, dtype=np.int32)
    lib().oracle_proposal_layer_batch(_p(d, ctypes.c_float), _p(s, ctypes.c_float), B, N,
                                      float(img_size[1]), float(img_size[2]),
                                      float(F32(min_size * scale)), int(n_pre_nms),
                                      int(n_post_nms), float(nms_iou), _p(out, ctypes.c_float),
                                      _p(src, ctypes.c_int64), _p(nk, ctypes.c_int64),
                                      _p(rc, ctypes.c_int32))
    return out, src, nk, rc


# --------------------------------------------------------------------------------------------
# a9: AnchorTargetCreator
# --------------------------------------------------------------------------------------------
def anchor_targets(bbox, anchor, n_sample=256, pos_iou_thresh=0.7, neg_iou_thresh=0.3,
                   pos_ratio=0.5, return_extra=False):
    """AnchorTargetCreator.__call__, nets/frcnn_training.py:29-101 -- quirks kept:
    one anchor per GT (first index on ties), later GT wins collisions, positives capped to the
    FIRST n_pos by index, negatives never subsampled unless n_neg <= 0 (len() of a 1-tuple)."""
    bbox = _f32(bbox).reshape(-1, 4)
    anchor = _f32(anchor).reshape(-1, 4)
    N, G = anchor.shape[0], bbox.shape[0]
    ious = iou(anchor, bbox)
    label = np.full((N,), -1, dtype=np.int64)
    if G == 0:
        argmax = np.zeros((N,), dtype=np.int64)
        max_ious = np.zeros((N,), dtype=F32)
        gt_argmax = np.zeros((0,), dtype=np.int64)
    else:
        argmax = ious.argmax(axis=1).astype(np.int64)
        max_ious = ious[np.arange(N), argmax]
        gt_argmax = ious.argmax(axis=0).astype(np.int64)
        for i in range(G):
            argmax[gt_argmax[i]] = i
    label[max_ious < F32(neg_iou_thresh)] = 0
    label[max_ious >= F32(pos_iou_thresh)] = 1
    if G > 0:
        label[gt_argmax] = 1
    n_pos = int(pos_ratio * n_sample)
    pos = np.where(label == 1)[0]
    pos_len = pos.size
    if pos_len > n_pos:
        label[pos[n_pos:]] = -1
        pos_len = n_pos
    n_neg = n_sample - pos_len
    neg = np.where(label == 0)[0]
    if 1 > n_neg:  # len(torch.where(...)) == 1  (frcnn_training.py:96-99)
        label[neg[n_neg:]] = -1
    if (label > 0).any():
        loc = encode(anchor, bbox[argmax])
    else:
        loc = np.zeros_like(anchor)
    if return_extra:
        return loc, label, argmax, max_ious, gt_argmax
    return loc, label


# --------------------------------------------------------------------------------------------
# a10: ProposalTargetCreator
# --------------------------------------------------------------------------------------------
def proposal_targets(roi, bbox, label, n_sample=128, pos_ratio=0.5, pos_iou_thresh=0.5,
                     neg_iou_thresh_high=0.5, neg_iou_thresh_low=0.0):
    """ProposalTargetCreator.__call__, nets/frcnn_training.py:122-177 -- loc_normalize_std is
    unused by the reference; the label scatter uses ORIGINAL roi indices on the SAMPLED array
    (raises IndexError when one is out of range)."""
    roi = _f32(roi).reshape(-1, 4)
    bbox = _f32(bbox).reshape(-1, 4)
    label = np.asarray(label).reshape(-1).astype(np.int64)
    pos_per_image = int(n_sample * pos_ratio)
    roi = np.concatenate([roi, bbox], axis=0)
    R, G = roi.shape[0], bbox.shape[0]
    m = iou(roi, bbox)
    if G == 0:
        assign = np.zeros((R,), dtype=np.int64)
        max_iou = np.zeros((R,), dtype=F32)
        roi_label = np.zeros((R,), dtype=np.int64)
    else:
        assign = m.argmax(axis=1).astype(np.int64)
        max_iou = m[np.arange(R), assign]
        roi_label = label[assign] + 1
    pos = np.where(max_iou >= F32(pos_iou_thresh))[0]
    if pos.size > pos_per_image:
        pos = pos[:pos_per_image]
    neg = np.where((max_iou < F32(neg_iou_thresh_high)) & (max_iou >= F32(neg_iou_thresh_low)))[0]
    n_neg = n_sample - pos.size
    if neg.size > n_neg:
        neg = neg[:n_neg]
    keep = np.concatenate([pos, neg])
    sample_roi = roi[keep]
    if G == 0:
        return sample_roi, np.zeros_like(sample_roi), roi_label[keep]
    gt_loc = encode(sample_roi, bbox[assign[keep]])
    out_label = roi_label[keep].copy()
    if neg.size and (neg.max() >= out_label.shape[0]):
        raise OracleIndexError("label scatter index out of range (nets/frcnn_training.py:175)")
    out_label[neg] = 0
    return sample_roi, gt_loc, out_label


# --------------------------------------------------------------------------------------------
# a11-a13: RoI head gather
# --------------------------------------------------------------------------------------------
def roi_to_feature_coords(rois, img_size, feat_h, feat_w) -> np.ndarray:
    """nets/classify.py:33-36: x / img_size[1] * W_f ; y / img_size[0] * H_f (two roundings)."""
    r = _f32(rois).reshape(-1, 4)
    out = np.zeros_like(r)
    out[:, [0, 2]] = r[:, [0, 2]] / F32(img_size[1]) * F32(feat_w)
    out[:, [1, 3]] = r[:, [1, 3]] / F32(img_size[0]) * F32(feat_h)
    return out


def roi_pool(feat, rois5, output_size, spatial_scale=1.0, return_argmax=False):
    """torchvision.ops.roi_pool as used by RoIPool at nets/classify.py:43."""
    f = _f32(feat)
    r = _f32(rois5).reshape(-1, 5)
    B, C, H, W = f.shape
    ph, pw = (output_size, output_size) if np.isscalar(output_size) else output_size
    out = np.empty((r.shape[0], C, ph, pw), dtype=F32)
    am = np.empty(out.shape, dtype=np.int32)
    lib().oracle_roi_pool(_p(f, ctypes.c_float), B, C, H, W, _p(r, ctypes.c_float), r.shape[0],
                          ph, pw, float(spatial_scale), _p(out, ctypes.c_float),
                          _p(am, ctypes.c_int32))
    return (out, am) if return_argmax else out


def roi_align(feat, rois5, output_size, spatial_scale=1.0, sampling_ratio=-1, aligned=False):
    """torchvision.ops.roi_align (oracle for the RoIAlign 7x7 configuration)."""
    f = _f32(feat)
    r = _f32(rois5).reshape(-1, 5)
    B, C, H, W = f.shape
    ph, pw = (output_size, output_size) if np.isscalar(output_size) else output_size
    out = np.empty((r.shape[0], C, ph, pw), dtype=F32)
    lib().oracle_roi_align(_p(f, ctypes.c_float), B, C, H, W, _p(r, ctypes.c_float), r.shape[0],
                           ph, pw, float(spatial_scale), int(sampling_ratio), int(bool(aligned)),
                           _p(out, ctypes.c_float))
    return out


def roi_pool_backward(grad_out, argmax, rois5, feat_shape):
    """torchvision RoIPool backward w.r.t. the features: grad_in[b,c,argmax] += grad_out, in the CPU
    kernel's sequential order.  feat_shape = (B,C,H,W)."""
    go = _f32(grad_out)
    am = np.ascontiguousarray(argmax, dtype=np.int32)
    r = _f32(rois5).reshape(-1, 5)
    B, C, H, W = (int(v) for v in feat_shape)
    gi = np.empty((B, C, H, W), dtype=F32)
    lib().oracle_roi_pool_backward(_p(go, ctypes.c_float), _p(am, ctypes.c_int32), _p(r, ctypes.c_float),
                                   r.shape[0], B, C, H, W, go.shape[2], go.shape[3], _p(gi, ctypes.c_float))
    return gi


def roi_align_backward(grad_out, rois5, feat_shape, spatial_scale=1.0, sampling_ratio=-1, aligned=False):
    """torchvision roi_align backward w.r.t. the features (what autograd computes through
    torchvision.ops.roi_align): grad * w_i / count onto the four taps of every sample."""
    go = _f32(grad_out)
    r = _f32(rois5).reshape(-1, 5)
    B, C, H, W = (int(v) for v in feat_shape)
    gi = np.empty((B, C, H, W), dtype=F32)
    lib().oracle_roi_align_backward(_p(go, ctypes.c_float), _p(r, ctypes.c_float), r.shape[0], B, C, H, W,
                                    go.shape[2], go.shape[3], float(spatial_scale), int(sampling_ratio),
                                    int(bool(aligned)), _p(gi, ctypes.c_float))
    return gi


def roi_head_gather(feat, rois, roi_indices, img_size, roi_size=7, spatial_scale=1.0,
                    op="pool", **kw):
    """nets/classify.py:29-43 generalised over the per-image RoI count (reference hard-codes 128):
    rois [n,R,4] image coords, roi_indices [n]; returns the pooled tensor [n*R,C,P,P]."""
    f = _f32(feat)
    r = _f32(rois)
    n, R = r.shape[0], r.shape[1]
    fm = roi_to_feature_coords(r.reshape(-1, 4), img_size, f.shape[2], f.shape[3])
    idx = np.repeat(np.asarray(roi_indices).reshape(-1).astype(F32), R)[:, None]
    r5 = np.concatenate([idx, fm], axis=1)
    if op == "pool":
        return roi_pool(f, r5, roi_size, spatial_scale)
    return roi_align(f, r5, roi_size, spatial_scale, **kw)
