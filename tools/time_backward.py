"""RoIPool / RoIAlign backward (gradient w.r.t. the features) at the training-config and inference-config sizes.
FRCNN_BACKWARD_DIRECT=1 selects the one-global-atomic-per-element kernels for comparison."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from two_stage_object_detection_b200 import functional as F, _lib
dev = torch.device("cuda:0")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for name, per in (("cfg3", 128), ("cfg4", 300), ("cfg2", 300)):
    cfg = bench.WORKLOADS[name]
    B, H, W, C, S = cfg["batch"], cfg["H"], cfg["W"], cfg["C"], cfg["img"]
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(B, C, H, W, generator=g).to(dev)
    ctr = torch.rand(B * per, 2, generator=g) * W
    wh = 2 + torch.rand(B * per, 2, generator=g) * (W / 2)
    rois = torch.cat([torch.arange(B).repeat_interleave(per)[:, None].float(), ctr - wh / 2, ctr + wh / 2], 1).to(dev)
    lib = _lib.load()
    stream = _lib.stream_ptr(dev)
    for P in (7,) if name != "cfg2" else (7, 14):
        out, am = F.roi_pool_forward(feat, rois, P, 1.0, with_argmax=True, rois_per_image=per)
        go = torch.randn_like(out)
        gi = torch.zeros_like(feat)
        def pool_bw():
            gi.zero_()
            _lib.check(lib.frcnn_roi_pool_backward(go.data_ptr(), am.data_ptr(), rois.data_ptr(), rois.shape[0], B, C, H, W, P, P,
                                                   gi.data_ptr(), stream), "pool backward")
        def align_bw():
            gi.zero_()
            _lib.check(lib.frcnn_roi_align_backward(go.data_ptr(), rois.data_ptr(), rois.shape[0], B, C, H, W, P, P, 1.0, 2, 0,
                                                    gi.data_ptr(), stream), "align backward")
        tz = timeit(lambda: gi.zero_())
        print(f"{name} K={rois.shape[0]} C={C} {H}x{W} P={P}: roi_pool backward {timeit(pool_bw) - tz:.4f} ms, roi_align backward "
              f"{timeit(align_bw) - tz:.4f} ms (zero-fill of grad_in {tz:.4f} ms excluded; grad_out {go.numel() * 4 / 1e6:.0f} MB)")
