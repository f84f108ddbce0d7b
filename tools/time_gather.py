"""Time the RoI gather of the bench workload in isolation (CUDA events)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from two_stage_object_detection_b200 import functional as F
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg = bench.WORKLOADS[name]; dev = torch.device('cuda:0')
loc, logits, feat = bench.make_inputs(cfg, 1, device=dev)
B, H, W, C, P, S = cfg["batch"], cfg["H"], cfg["W"], cfg["C"], cfg["P"], cfg["img"]
base = F.base_anchors(device=dev)
rois, *_ = F.proposals(loc, logits, clip_x_max=S, clip_y_max=S, n_pre_nms=cfg["n_pre"], n_post_nms=cfg["n_post"], base=base, feat_stride=16, feat_hw=(H, W), score_is_logits=True)
idx = torch.arange(B, dtype=torch.int32, device=dev)
rois5 = F.roi_head_coords(rois, idx, (S, S), (H, W))
K = B * cfg["n_post"]
pooled = torch.empty((K, C, P, P), device=dev)
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
alg = K * C * P * P * 4 + K * 20 + B * C * H * W * 4
for grouped in (0, cfg["n_post"]):
    if cfg["op"] == "pool":
        t = timeit(lambda: F.roi_pool_forward(feat, rois5, P, 1.0, out=pooled, rois_per_image=grouped))
    else:
        t = timeit(lambda: F.roi_align_forward(feat, rois5, P, 1.0, 2, False, out=pooled, rois_per_image=grouped))
    print(f"{name} {cfg['op']} P={P} grouped={grouped}: {t:.3f} ms  {alg / t / 1e6:.0f} GB/s  ({alg / t / 1e6 / 6552.6:.3f} of peak)")
t = timeit(lambda: F.proposals(loc, logits, clip_x_max=S, clip_y_max=S, n_pre_nms=cfg["n_pre"], n_post_nms=cfg["n_post"], base=base, feat_stride=16, feat_hw=(H, W), score_is_logits=True))
print(f"{name} proposals: {t*1e3:.1f} us")
if cfg["op"] == "pool":
    t = timeit(lambda: F.roi_pool_mean(feat, rois5, P, 1.0, rois_per_image=cfg["n_post"]))
    t2 = timeit(lambda: F.roi_pool_forward(feat, rois5, P, 1.0, out=pooled, rois_per_image=cfg["n_post"]).mean((2, 3)))
    print(f"{name} fused pool+mean -> [K,C]: {t:.3f} ms   (pool then .mean((2,3)): {t2:.3f} ms)")
else:
    t = timeit(lambda: F.roi_align_mean(feat, rois5, P, 1.0, 2, False, rois_per_image=cfg["n_post"]))
    t2 = timeit(lambda: F.roi_align_forward(feat, rois5, P, 1.0, 2, False, out=pooled, rois_per_image=cfg["n_post"]).mean((2, 3)))
    print(f"{name} fused align+mean -> [K,C]: {t:.3f} ms   (align then .mean((2,3)): {t2:.3f} ms)")
