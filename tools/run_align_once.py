"""One RoIAlign forward of the cfg4 workload (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from two_stage_object_detection_b200 import functional as F
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
cfg = bench.WORKLOADS[name]; dev = torch.device('cuda:0')
loc, logits, feat = bench.make_inputs(cfg, 1, device=dev)
B, H, W, C, P, S = cfg["batch"], cfg["H"], cfg["W"], cfg["C"], cfg["P"], cfg["img"]
base = F.base_anchors(device=dev)
rois, *_ = F.proposals(loc, logits, clip_x_max=S, clip_y_max=S, n_pre_nms=cfg["n_pre"], n_post_nms=cfg["n_post"], base=base, feat_stride=16, feat_hw=(H, W), score_is_logits=True)
idx = torch.arange(B, dtype=torch.int32, device=dev)
rois5 = F.roi_head_coords(rois, idx, (S, S), (H, W))
pooled = torch.empty((B * cfg["n_post"], C, P, P), device=dev)
for _ in range(3):
    if cfg["op"] == "pool":
        F.roi_pool_forward(feat, rois5, P, 1.0, out=pooled, rois_per_image=cfg["n_post"])
    else:
        F.roi_align_forward(feat, rois5, P, 1.0, 2, False, out=pooled, rois_per_image=cfg["n_post"])
torch.cuda.synchronize()
