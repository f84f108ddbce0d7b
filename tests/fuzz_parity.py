"""Randomised parity sweep: CUDA path vs the oracle on random shapes and box distributions (degenerate boxes,
RoIs hanging over the map, channel tails, ragged GT counts).  Not part of the test suite (minutes, not seconds);
run on a GPU box:   python tests/fuzz_parity.py [--iters 200] [--seed 0]
Exits non-zero on the first mismatch and prints the case that produced it."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_port as O  # noqa: E402  (checker only)
from two_stage_object_detection_b200 import functional as F  # noqa: E402

DEV = torch.device("cuda:0")


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def rand_rois(rng, K, B, H, W):
    kind = rng.integers(0, 4)
    c = rng.uniform(-4, max(W, H) + 4, (K, 2))
    if kind == 0:
        wh = rng.uniform(0.2, 8, (K, 2))
    elif kind == 1:
        wh = rng.uniform(W * 0.3, W * 1.3, (K, 2))
    elif kind == 2:
        wh = np.concatenate([rng.uniform(0.2, 10, (K // 2, 2)), rng.uniform(4, W, (K - K // 2, 2))])
    else:
        wh = rng.uniform(0, 3, (K, 2)) * rng.integers(0, 2, (K, 2))  # zero-size and tiny
    r = np.concatenate([rng.integers(0, B, (K, 1)), c - wh / 2, c + wh / 2], 1).astype(np.float32)
    if K > 3:
        r[1, [1, 3]] = r[1, [3, 1]]  # inverted
    if rng.integers(0, 2):
        r[:, 1:] = np.round(r[:, 1:] * 2) / 2  # half-pixel grid: exact .5 rounding cases
    return r


def check_roi(rng, it):
    B, Cc = int(rng.integers(1, 4)), int(rng.integers(1, 13))
    H, W = int(rng.integers(5, 70)), int(rng.integers(5, 70))
    P = int(rng.choice([7, 14, 5, 3]))
    K = int(rng.integers(1, 200))
    feat = rng.standard_normal((B, Cc, H, W)).astype(np.float32)
    if rng.integers(0, 3) == 0:
        feat = np.round(feat * 2) / 2  # ties
    rois = rand_rois(rng, K, B, H, W)
    scale = float(rng.choice([1.0, 0.5, 0.0625 * 16]))
    tag = f"roi it={it} B={B} C={Cc} H={H} W={W} P={P} K={K} scale={scale}"
    ro, ra = O.roi_pool(feat, rois, P, scale, return_argmax=True)
    out, am = F.roi_pool_forward(T(feat), T(rois), P, scale, with_argmax=True)
    assert np.array_equal(N(out), ro) and np.array_equal(N(am), ra), tag + " pool+argmax"
    assert np.array_equal(N(F.roi_pool_forward(T(feat), T(rois), P, scale)), ro), tag + " pool"
    sr = int(rng.choice([1, 2, 3]))
    al = bool(rng.integers(0, 2))
    rl = O.roi_align(feat, rois, P, scale, sr, al)
    assert np.array_equal(N(F.roi_align_forward(T(feat), T(rois), P, scale, sr, al, exact=True)), rl), \
        tag + f" align sr={sr} al={al} (reference order)"
    fast = N(F.roi_align_forward(T(feat), T(rois), P, scale, sr, al, exact=False))  # streaming / FMA variants where they exist
    assert np.abs(fast - rl).max() <= 1e-5 * np.abs(feat).max(), tag + f" align sr={sr} al={al} (fast)"
    if P in (7, 14):
        ref = ro.astype(np.float64).mean((2, 3))
        got = N(F.roi_pool_mean(T(feat), T(rois), P, scale))
        assert np.abs(got - ref).max() <= 1e-5 * max(np.abs(ro).max(), 1e-30), tag + " pool_mean"
    if H <= 64 and W <= 64 and P * sr <= 32:
        ref = rl.astype(np.float64).mean((2, 3))
        got = N(F.roi_align_mean(T(feat), T(rois), P, scale, sr, al))
        assert np.abs(got - ref).max() <= 1e-5 * np.abs(feat).max(), tag + " align_mean"


def rand_boxes(rng, n, size, lo=4, hi=None):
    hi = max(hi or size / 2, lo + 1)
    c = rng.uniform(0, size, (n, 2))
    wh = rng.uniform(lo, hi, (n, 2))
    return np.concatenate([c - wh / 2, c + wh / 2], 1).astype(np.float32)


def check_targets(rng, it):
    H, W = int(rng.integers(2, 40)), int(rng.integers(2, 40))
    S = 16 * max(H, W)
    B = int(rng.integers(1, 4))
    gts = [rand_boxes(rng, int(rng.integers(0, 12)), S, 20, S / 2) for _ in range(B)]
    for g in gts:
        if len(g) > 2 and rng.integers(0, 2):
            g[1] = g[0]  # duplicate GT: later one wins the shared best anchor
    labels = [rng.integers(0, 20, len(g)) for g in gts]
    tag = f"targets it={it} H={H} W={W} B={B} G={[len(g) for g in gts]}"
    bb, ll, n_gt = F.pad_gt([T(g) for g in gts], [T(l) for l in labels], device=DEV)
    base = F.base_anchors(device=DEV)
    loc, label, am = F.anchor_targets(bb, n_gt, base=base, feat_stride=16, feat_hw=(H, W), return_argmax=True)
    anchor = O.shifted_anchors(O.base_anchors(), 16, H, W)
    for b in range(B):
        rl, rlab = O.anchor_targets(gts[b], anchor)[:2]
        assert np.array_equal(N(label[b]), rlab), tag + f" labels b={b}"
        assert np.allclose(N(loc[b]), rl, rtol=1e-5, atol=1e-5, equal_nan=True), tag + f" loc b={b}"
    R = int(rng.integers(1, 300))
    rois = np.stack([rand_boxes(rng, R, S, 8, S / 2) for _ in range(B)])
    s, l, y, n_out, status = F.proposal_targets(T(rois), bb, ll, n_gt)
    for b in range(B):
        try:
            rs, rl2, ry = O.proposal_targets(rois[b], gts[b], labels[b])
        except IndexError:
            assert int(status[b]) == 2, tag + f" status b={b}"
            continue
        k = int(n_out[b])
        assert int(status[b]) == 0 and k == rs.shape[0], tag + f" n_out b={b}"
        assert np.array_equal(N(s[b, :k]), rs) and np.array_equal(N(y[b, :k]), ry), tag + f" sample b={b}"
        assert np.allclose(N(l[b, :k]), rl2, rtol=1e-5, atol=1e-5, equal_nan=True), tag + f" gt_loc b={b}"


def check_proposals(rng, it):
    H, W = int(rng.integers(3, 40)), int(rng.integers(3, 40))
    B = int(rng.integers(1, 4))
    Nn = H * W * 9
    S = 16 * max(H, W)
    n_pre = int(rng.integers(10, 3000))
    n_post = int(rng.integers(1, min(n_pre, 600) + 1))
    loc = (rng.standard_normal((B, Nn, 4)) * rng.choice([0.1, 0.3, 1.0])).astype(np.float32)
    score = rng.uniform(0, 1, (B, Nn)).astype(np.float32)
    if rng.integers(0, 2):
        score = np.round(score * 64) / 64  # many exact ties
    base = F.base_anchors(device=DEV)
    kw = dict(clip_x_max=float(S), clip_y_max=float(S), min_size=16.0)
    tag = f"proposals it={it} H={H} W={W} B={B} n_pre={n_pre} n_post={n_post}"
    boxes, keys, fg = F.decode_clip_score(T(loc), T(score), base=base, feat_stride=16, feat_hw=(H, W), **kw)
    rois, src, n_keep, status = F.proposals(T(loc), T(score), base=base, feat_stride=16, feat_hw=(H, W),
                                            n_pre_nms=n_pre, n_post_nms=n_post, nms_iou=0.7, **kw)
    ref, ref_src, ref_nk, rc = O.proposal_layer_batch_from_boxes(N(boxes), N(fg), (3, S, S), 1.0, 0.7, n_pre, n_post, 16)
    for b in range(B):
        if rc[b]:
            assert int(status[b]) != 0, tag + f" status b={b}"
            continue
        assert int(status[b]) == 0, tag + f" status b={b}"
        assert np.array_equal(N(rois[b]), ref[b]) and np.array_equal(N(src[b]).astype(np.int64), ref_src[b]), tag + f" b={b}"


def check_long_nms(rng, it):
    """Long score-ordered candidate lists (several super-blocks; the adaptive schedule when 2 * cap > 2048) with
    random density, i.e. keep rates from a few percent to nearly all, images of different lengths in one batch."""
    if it % 8:
        return
    B = int(rng.integers(1, 4))
    R = int(rng.integers(2500, 14000))
    cap = int(rng.choice([300, 1100, 2000, 3000]))
    thr = float(rng.choice([0.5, 0.7]))
    n_sel = rng.integers(R // 3, R + 1, B).astype(np.int32)
    batch = np.zeros((B, R, 4), np.float32)
    refs = []
    for b in range(B):
        extent = float(rng.choice([200, 600, 2000, 6000]))
        bx = rand_boxes(rng, int(n_sel[b]), extent, 10, 120)
        batch[b, :n_sel[b]] = bx
        refs.append(O.nms(bx, np.arange(len(bx), 0, -1, dtype=np.float32), thr))
    keep, n_keep = F.nms_sorted(T(batch), T(n_sel), thr, cap)
    for b in range(B):
        k = int(n_keep[b])
        tag = f"long nms it={it} B={B} R={R} cap={cap} thr={thr} b={b}"
        assert k == min(cap, len(refs[b])), tag + f" count {k} vs {len(refs[b])}"
        assert np.array_equal(N(keep[b, :k]).astype(np.int64), refs[b][:k]), tag


def check_detections(rng, it):
    B, R, C = int(rng.integers(1, 4)), int(rng.integers(1, 700)), int(rng.integers(1, 30))
    boxes = np.stack([rand_boxes(rng, R, 300, 5, 150) for _ in range(B)])
    scores = rng.uniform(0, 1, (B, R)).astype(np.float32)
    if rng.integers(0, 2):
        scores = np.round(scores * 32) / 32
    cls = rng.integers(0, C, (B, R))
    thr = float(rng.choice([0.1, 0.3, 0.5, 0.7]))
    tag = f"detections it={it} B={B} R={R} C={C} thr={thr}"
    keep, nk = F.nms_by_class(T(boxes), T(scores), T(cls), thr)
    for b in range(B):
        ref = O.nms_by_class(boxes[b], scores[b], cls[b], thr)
        assert int(nk[b]) == len(ref) and np.array_equal(N(keep[b, :len(ref)]).astype(np.int64), ref), tag + f" b={b}"
    roi = rand_boxes(rng, R, 300, 5, 150)
    cl = (rng.standard_normal((R, C * 4)) * 0.3).astype(np.float32)
    sc = np.round(rng.standard_normal((R, C)) * 4).astype(np.float32) / 4
    lab = rng.integers(0, C, R)
    bx, cs, ci = F.detection_decode(T(roi), T(cl), T(sc), T(lab))
    rb, rs, ri = O.detection_decode(roi, cl, sc, lab)
    assert np.array_equal(N(ci), ri) and np.array_equal(N(cs), rs), tag + " decode class"
    assert np.allclose(N(bx), rb, rtol=1e-5, atol=3e-3), tag + " decode boxes"
    na, nb = int(rng.integers(1, 400)), int(rng.integers(1, 40))
    a, b2 = rand_boxes(rng, na, 200, 0, 80), rand_boxes(rng, nb, 200, 0, 80)
    assert np.array_equal(N(F.bbox_iou(T(a), T(b2))), O.iou(a, b2)), tag + f" iou {na}x{nb}"


def check_align_stream(rng, it):
    """The streaming RoIAlign kernel (7x7 bins, 2x2 sampling grid, channels a multiple of four): ragged images, grouped
    and bucketed RoI lists, degenerate / inverted / out-of-map RoIs, against the oracle at 1e-5 of the largest feature."""
    B, Cc = int(rng.integers(1, 5)), 4 * int(rng.integers(1, 5))
    H, W = int(rng.integers(5, 70)), int(rng.integers(5, 70))
    per = int(rng.integers(1, 90))
    grouped = bool(rng.integers(0, 2))
    K = per * B if grouped else int(rng.integers(1, 300))
    feat = rng.standard_normal((B, Cc, H, W)).astype(np.float32)
    rois = rand_rois(rng, K, B, H, W)
    if grouped:
        rois[:, 0] = np.repeat(np.arange(B), per)
    al = bool(rng.integers(0, 2))
    scale = float(rng.choice([1.0, 0.5, 0.25]))
    tag = f"align-stream it={it} B={B} C={Cc} H={H} W={W} K={K} grouped={grouped} aligned={al} scale={scale}"
    ref = O.roi_align(feat, rois, 7, scale, 2, al)
    got = N(F.roi_align_forward(T(feat), T(rois), 7, scale, 2, al, exact=False, rois_per_image=per if grouped else 0))
    assert F._lib.last_roi_kernel().startswith("roi_align_stream2"), tag + " took " + F._lib.last_roi_kernel()
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(feat).max(), tag


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    checks = [check_roi, check_targets, check_proposals, check_detections, check_long_nms, check_align_stream]
    counts = {c.__name__: 0 for c in checks}
    for it in range(args.iters):
        for c in checks:
            c(rng, it)
            counts[c.__name__] += 1
    torch.cuda.synchronize()
    print("fuzz ok:", counts, "seed", args.seed)


if __name__ == "__main__":
    main()
