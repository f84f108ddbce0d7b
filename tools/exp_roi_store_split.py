import os, sys, torch, time
sys.path.insert(0, '/root/repo')
import bench
from two_stage_object_detection_b200 import functional as F
cfg = bench.WORKLOADS['cfg2']; dev = torch.device('cuda:0')
loc, logits, feat = bench.make_inputs(cfg, 1, device=dev)
B,H,W,C,P = 16,38,38,1024,14
base = F.base_anchors(device=dev)
rois, *_ = F.proposals(loc, logits, clip_x_max=600, clip_y_max=600, n_pre_nms=3000, n_post_nms=300, base=base, feat_stride=16, feat_hw=(H,W), score_is_logits=True)
idx = torch.arange(B, dtype=torch.int32, device=dev)
rois5 = F.roi_head_coords(rois, idx, (600,600), (H,W))
pooled = torch.empty((4800, C, P, P), device=dev)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
t = timeit(lambda: pooled.fill_(1.0)); print(f"fill_ 3.85GB: {t:.3f} ms  {pooled.numel()*4/t/1e6:.0f} GB/s")
src = torch.empty_like(pooled)
t = timeit(lambda: pooled.copy_(src)); print(f"copy_ 3.85GB: {t:.3f} ms  {2*pooled.numel()*4/t/1e6:.0f} GB/s (r+w)")
for dbg in (0, 1, 2):
    os.environ['FRCNN_ROI_DEBUG'] = str(dbg)
    t = timeit(lambda: F.roi_pool_forward(feat, rois5, P, 1.0, out=pooled)); print(f"roi_pool debug={dbg}: {t:.3f} ms")
