"""Micro-benchmarks of the streaming kernels at batches large enough (>= 256 MB per launch) for
an HBM fraction to be meaningful: fused decode-clip-score and the fused IoU + argmax + label + encode of
AnchorTargetCreator.  Algorithmic bytes per unit are the SURVEY 8d figures.  CUDA events, L2 defeated by
the >= 256 MB working set."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from two_stage_object_detection_b200 import functional as F  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 6552.6
if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")):
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]


def timeit(fn, n=int(os.environ.get("MB_ITERS", 20))):
    for _ in range(int(os.environ.get("MB_WARM", 3))):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


H = W = 38
N = H * W * 9
base = F.base_anchors(device=dev)
out = {}
for B in [int(v) for v in os.environ.get("MB_B", "16,512").split(",")]:
    g = torch.Generator().manual_seed(0)
    loc = (torch.randn(B, N, 4, generator=g) * 0.2).to(dev)
    logits = torch.randn(B, N, 2, generator=g).to(dev)
    t = timeit(lambda: F.decode_clip_score(loc, logits, clip_x_max=600, clip_y_max=600, base=base, feat_stride=16,
                                           feat_hw=(H, W), score_is_logits=True))
    # functional.decode_clip_score also writes the fg score ([B,N] fp32) for the caller: 48 B/anchor here
    by = B * N * 48
    out[f"decode_clip_score B={B}"] = dict(ms=t, GBs=by / t / 1e6, frac=by / t / 1e6 / PEAK, bytes=by)
    gts = torch.rand(B, 8, 2, generator=g) * 600
    wh = 50 + torch.rand(B, 8, 2, generator=g) * 200
    bbox = torch.cat([gts - wh / 2, gts + wh / 2], -1).clamp(0, 600).to(dev)
    n_gt = torch.full((B,), 8, dtype=torch.int32, device=dev)
    t = timeit(lambda: F.anchor_targets(bbox, n_gt, base=base, feat_stride=16, feat_hw=(H, W)))
    # 40 B/anchor (anchor 16 in, label int64 8 + loc 16 out) + 16 G.  Anchors are generated in
    # registers here, so the bytes that really move are 24 out + one 4-byte word written and re-read (32).
    by, moved = B * N * 40 + B * 8 * 16, B * N * 32 + B * 8 * 16
    out[f"anchor_targets B={B} G=8"] = dict(ms=t, GBs=by / t / 1e6, frac=by / t / 1e6 / PEAK, bytes=by,
                                            GBs_moved=moved / t / 1e6, frac_moved=moved / t / 1e6 / PEAK)
    for nb in (8, 64):
        na = B * N // 8 if nb == 64 else B * N
        c = torch.rand(na, 2, generator=g) * 600
        a = torch.cat([c - 40, c + 40], 1).to(dev)
        c = torch.rand(nb, 2, generator=g) * 600
        b8 = torch.cat([c - 100, c + 100], 1).to(dev)
        t = timeit(lambda: F.bbox_iou(a, b8))
        by = na * 16 + nb * 16 + na * nb * 4
        out[f"bbox_iou [{na}x{nb}]"] = dict(ms=t, GBs=by / t / 1e6, frac=by / t / 1e6 / PEAK, bytes=by)
for k, v in out.items():
    print(k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items()})
