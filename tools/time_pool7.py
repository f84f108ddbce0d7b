"""RoIPool 7x7 forward on the feature map / proposals of a bench workload (CUDA events): python tools/time_pool7.py cfg4"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from two_stage_object_detection_b200 import functional as F, _lib
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
cfg = bench.WORKLOADS[name]; dev = torch.device('cuda:0')
loc, logits, feat = bench.make_inputs(cfg, 1, device=dev)
B, H, W, C, S = cfg["batch"], cfg["H"], cfg["W"], cfg["C"], cfg["img"]
base = F.base_anchors(device=dev)
rois, *_ = F.proposals(loc, logits, clip_x_max=S, clip_y_max=S, n_pre_nms=cfg["n_pre"], n_post_nms=cfg["n_post"], base=base, feat_stride=16, feat_hw=(H, W), score_is_logits=True)
idx = torch.arange(B, dtype=torch.int32, device=dev)
rois5 = F.roi_head_coords(rois, idx, (S, S), (H, W))
K = B * cfg["n_post"]
pooled = torch.empty((K, C, 7, 7), device=dev)
fn = lambda: F.roi_pool_forward(feat, rois5, 7, 1.0, out=pooled, rois_per_image=cfg["n_post"])
for _ in range(5): fn()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(30): fn()
b.record(); torch.cuda.synchronize()
t = a.elapsed_time(b) / 30
alg = K * C * 49 * 4 + K * 20 + B * C * H * W * 4
print(f"{name} map {H}x{W} C={C} K={K}: {_lib.last_roi_kernel()}  {t:.4f} ms  {alg / t / 1e6 / 6552.6:.3f} of peak")
fn = lambda: F.roi_pool_mean(feat, rois5, 7, 1.0, rois_per_image=cfg["n_post"])
for _ in range(5): fn()
torch.cuda.synchronize()
a.record()
for _ in range(30): fn()
b.record(); torch.cuda.synchronize()
print(f"{name} fused pool + mean: {_lib.last_roi_kernel()}  {a.elapsed_time(b) / 30:.4f} ms")
