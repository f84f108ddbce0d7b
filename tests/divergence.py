"""Why may two runs of the proposal layer differ?  (test infrastructure, numpy only)

`decode` uses exp(): the reference's SLEEF exp, numpy's and CUDA's expf differ in the last place, so boxes
decoded by two implementations agree to a few units in the last place but not bit for bit.  Everything after
the decode is a chain of hard decisions (min-size filter, greedy NMS with a `> thr` test); a coordinate that
moves by one ulp can flip a decision that sat on its threshold, and one flipped NMS decision shifts every
later output row.  "x % of the rows may differ" is therefore the wrong test.  The right one, implemented here:

  * walk the reference's and the other implementation's candidate lists in lockstep;
  * up to the FIRST decision on which they disagree every output row must agree within the decode tolerance;
  * that first disagreement must be a threshold flip: a min-size test within `size_tol` of min_size, or an NMS
    test whose IoU (against an already kept box) is within `iou_tol` of the threshold -- on both sides;
  * with no disagreement at all, ALL rows must agree.

Arithmetic is fp32 in the order torchvision's nms kernel uses, so the IoUs seen here are the ones
the implementations compared.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def clip_valid(decoded, img_size, min_size):
    """nets/rpn.py:47-54: clamp x to [0,img_size[1]], y to [0,img_size[2]]; valid = both sides >= min_size."""
    roi = np.array(decoded, dtype=F32, copy=True)
    roi[:, [0, 2]] = np.clip(roi[:, [0, 2]], F32(0), F32(img_size[1]))
    roi[:, [1, 3]] = np.clip(roi[:, [1, 3]], F32(0), F32(img_size[2]))
    w, h = roi[:, 2] - roi[:, 0], roi[:, 3] - roi[:, 1]
    return roi, (w >= F32(min_size)) & (h >= F32(min_size)), w, h


def _iou_vs_kept(box, area, kb, ka):
    if not len(kb):
        return np.zeros(0, F32)
    xx1, yy1 = np.maximum(kb[:, 0], box[0]), np.maximum(kb[:, 1], box[1])
    xx2, yy2 = np.minimum(kb[:, 2], box[2]), np.minimum(kb[:, 3], box[3])
    w, h = np.maximum(F32(0), xx2 - xx1), np.maximum(F32(0), yy2 - yy1)
    inter = (w * h).astype(F32)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (inter / ((ka + area).astype(F32) - inter).astype(F32)).astype(F32)


def first_divergence(dec_ref, dec_other, score, img_size, min_size, nms_iou, n_pre, n_post,
                     size_tol=1e-3, iou_tol=1e-4):
    """Returns a dict: {"kind": None | "min_size" | "nms", "rows_equal": number of leading output rows that must
    agree, "detail": str}.  Raises AssertionError when the first disagreement is NOT a threshold flip."""
    roi_r, ok_r, w_r, h_r = clip_valid(dec_ref, img_size, min_size)
    roi_o, ok_o, _, _ = clip_valid(dec_other, img_size, min_size)
    score = np.asarray(score, dtype=F32)
    flipped = np.nonzero(ok_r != ok_o)[0]
    for i in flipped:  # every validity flip must sit on the min-size threshold, on BOTH sides
        w_o, h_o = roi_o[i, 2] - roi_o[i, 0], roi_o[i, 3] - roi_o[i, 1]
        margin = np.inf
        if (w_r[i] >= min_size) != (w_o >= min_size):
            margin = min(margin, max(abs(float(w_r[i]) - min_size), abs(float(w_o) - min_size)))
        if (h_r[i] >= min_size) != (h_o >= min_size):
            margin = min(margin, max(abs(float(h_r[i]) - min_size), abs(float(h_o) - min_size)))
        assert margin <= size_tol, f"anchor {i}: min-size decision differs with margin {margin}"

    def ordered(ok):
        idx = np.nonzero(ok)[0]
        o = np.argsort(-score[idx].astype(np.float64), kind="stable")  # (score desc, index asc); no NaN scores here
        o = idx[o]
        return o[:n_pre] if n_pre > 0 else o

    cand_r, cand_o = ordered(ok_r), ordered(ok_o)
    thr = float(nms_iou)
    cap = max(n_post, 1)
    kept_r, kept_o = np.zeros((cap, 4), F32), np.zeros((cap, 4), F32)
    area_r, area_o = np.zeros(cap, F32), np.zeros(cap, F32)
    nk = 0
    n = min(len(cand_r), len(cand_o))
    for p in range(n):
        if nk >= n_post:
            return {"kind": None, "rows_equal": n_post, "detail": "no disagreement"}
        a, b = int(cand_r[p]), int(cand_o[p])
        if a != b:
            assert a in flipped or b in flipped, f"candidate {p}: anchors {a} / {b} differ without a min-size flip"
            return {"kind": "min_size", "rows_equal": nk,
                    "detail": f"candidate list differs at position {p} (anchor {a} vs {b}: a min-size flip)"}
        br, bo = roi_r[a], roi_o[a]
        ar = F32((br[2] - br[0]) * (br[3] - br[1]))
        ao = F32((bo[2] - bo[0]) * (bo[3] - bo[1]))
        iou_r = _iou_vs_kept(br, ar, kept_r[:nk], area_r[:nk])
        iou_o = _iou_vs_kept(bo, ao, kept_o[:nk], area_o[:nk])
        # torchvision compares the fp32 quotient with the double threshold; NaN (0/0) never suppresses
        sup_r = bool((iou_r.astype(np.float64) > thr).any())
        sup_o = bool((iou_o.astype(np.float64) > thr).any())
        if sup_r != sup_o:
            # the pair whose test flipped: the IoU closest to the threshold among the kept boxes, on the side that
            # suppressed; the same pair's IoU on the other side must be just as close
            sup, keep = (iou_r, iou_o) if sup_r else (iou_o, iou_r)
            j = int(np.argmin(np.where(sup.astype(np.float64) > thr, sup.astype(np.float64) - thr, np.inf)))
            near = max(abs(float(sup[j]) - thr), abs(float(keep[j]) - thr))
            assert near <= iou_tol, f"candidate {p}: NMS decision differs, IoUs {sup[j]} / {keep[j]} vs {thr}"
            return {"kind": "nms", "rows_equal": nk,
                    "detail": f"candidate {p} (anchor {a}): IoU within {near:.2e} of {thr}"}
        if not sup_r:
            kept_r[nk], kept_o[nk], area_r[nk], area_o[nk] = br, bo, ar, ao
            nk += 1
    if len(cand_r) != len(cand_o):
        return {"kind": "min_size", "rows_equal": nk, "detail": "candidate lists differ in length"}
    return {"kind": None, "rows_equal": n_post, "detail": "no disagreement"}
