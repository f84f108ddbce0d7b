"""A few cfg3-shaped training RoIPool launches (7x7 + argmax on 128 sampled RoIs per image) for ncu captures / timing."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from two_stage_object_detection_b200 import functional as F, _lib
dev = torch.device('cuda:0')
wl = bench.Workload("cfg3", dev, 0, 1)
for i in range(3):
    wl.step(i)
torch.cuda.synchronize()
loc, logits, feat = wl.sets[0]
rois, *_ = F.proposals(loc, logits, **wl.pkw)
sel, _, _, _, _ = F.proposal_targets(rois, wl.gt_box, wl.gt_lab, wl.n_gt)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for _ in range(50):
    wl.gather(feat, sel)
b.record(); torch.cuda.synchronize()
print("train gather", _lib.last_roi_kernel(), a.elapsed_time(b) / 50, "ms; alg bytes", wl.alg_bytes)
