"""RoIAlign 7x7 sr=2: the selected non-exact kernel against the reference-order kernel, plus timing.
Env: FRCNN_ALIGN_IMPL / FRCNN_ALIGN_THREADS pick the variant (read once per process)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from two_stage_object_detection_b200 import functional as F, _lib
dev = torch.device('cuda:0')
rng = np.random.default_rng(5)
worst = 0.0
for (B, C, H, W, K, grouped) in [(2, 8, 50, 50, 301, False), (3, 12, 38, 38, 64, True), (1, 4, 64, 64, 700, False), (2, 8, 21, 33, 10, True),
                                 (4, 16, 50, 50, 1200, True)]:
    feat = rng.standard_normal((B, C, H, W)).astype(np.float32)
    c = rng.uniform(-4, W + 4, (K, 2)); wh = rng.uniform(0.3, W * 0.9, (K, 2))
    bi = np.repeat(np.arange(B), K // B)[:, None] if grouped else rng.integers(0, B, (K, 1))
    K2 = len(bi)
    rois = np.concatenate([bi, c[:K2] - wh[:K2] / 2, c[:K2] + wh[:K2] / 2], 1).astype(np.float32)
    rois[::17, 3] = rois[::17, 1] - 1.0   # inverted
    rois[3] = [rois[3, 0], -50, -50, -30, -30]  # out of the map
    for al in (False, True):
        ft, rt = torch.from_numpy(feat).to(dev), torch.from_numpy(rois).to(dev)
        ref = F.roi_align_forward(ft, rt, 7, 1.0, 2, al, exact=True, rois_per_image=K2 // B if grouped else 0)
        got = F.roi_align_forward(ft, rt, 7, 1.0, 2, al, exact=False, rois_per_image=K2 // B if grouped else 0)
        err = float((ref - got).abs().max()) / float(np.abs(feat).max())
        worst = max(worst, err)
        print((B, C, H, W, K2, grouped, al), _lib.last_roi_kernel(), f"err {err:.2e}", flush=True)
        assert err <= 1e-5
name = "cfg4"
cfg = bench.WORKLOADS[name]
loc, logits, feat = bench.make_inputs(cfg, 1, device=dev)
B, H, W, C, P, S = cfg["batch"], cfg["H"], cfg["W"], cfg["C"], cfg["P"], cfg["img"]
base = F.base_anchors(device=dev)
rois, *_ = F.proposals(loc, logits, clip_x_max=S, clip_y_max=S, n_pre_nms=cfg["n_pre"], n_post_nms=cfg["n_post"], base=base, feat_stride=16, feat_hw=(H, W), score_is_logits=True)
idx = torch.arange(B, dtype=torch.int32, device=dev)
rois5 = F.roi_head_coords(rois, idx, (S, S), (H, W))
K = B * cfg["n_post"]
pooled = torch.empty((K, C, P, P), device=dev)
ref = F.roi_align_forward(feat, rois5, P, 1.0, 2, False, exact=True, rois_per_image=cfg["n_post"])
for grouped in (cfg["n_post"], 0):
    pooled.fill_(float('nan'))
    F.roi_align_forward(feat, rois5, P, 1.0, 2, False, out=pooled, rois_per_image=grouped)
    err = float((ref - pooled).abs().max()) / float(feat.abs().max())
    print("cfg4 grouped", grouped, _lib.last_roi_kernel(), f"err {err:.2e}")
    assert err <= 1e-5
    for _ in range(5): F.roi_align_forward(feat, rois5, P, 1.0, 2, False, out=pooled, rois_per_image=grouped)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(30): F.roi_align_forward(feat, rois5, P, 1.0, 2, False, out=pooled, rois_per_image=grouped)
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 30
    alg = K * C * P * P * 4 + K * 20 + B * C * H * W * 4
    print(f"TIME grouped={grouped} {_lib.last_roi_kernel()}: {t:.4f} ms {alg / t / 1e6 / 6552.6:.3f} of peak")
