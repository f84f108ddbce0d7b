"""Eager vs CUDA-graph replay of one cfg2 step (proposals -> coords -> RoIPool)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from two_stage_object_detection_b200 import functional as F
cfg = bench.WORKLOADS["cfg2"]; dev = torch.device("cuda:0")
B, H, W, C, P, S, n_post = cfg["batch"], cfg["H"], cfg["W"], cfg["C"], cfg["P"], cfg["img"], cfg["n_post"]
sets = [bench.make_inputs(cfg, s, device=dev) for s in range(3)]
base = F.base_anchors(device=dev); idx = torch.arange(B, dtype=torch.int32, device=dev)
pooled = torch.empty((B * n_post, C, P, P), device=dev)
def step(i):
    loc, logits, feat = sets[i % 3]
    rois, *_ = F.proposals(loc, logits, clip_x_max=S, clip_y_max=S, n_pre_nms=cfg["n_pre"], n_post_nms=n_post, base=base, feat_stride=16, feat_hw=(H, W), score_is_logits=True)
    rois5 = F.roi_head_coords(rois, idx, (S, S), (H, W))
    F.roi_pool_forward(feat, rois5, P, 1.0, out=pooled, rois_per_image=n_post)
def timeit(fn, n=60):
    for i in range(6): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print(f"eager: {timeit(step):.4f} ms/step")
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for i in range(3): step(i)
torch.cuda.current_stream().wait_stream(s)
graphs = []
for i in range(3):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): step(i)
    graphs.append(g)
print(f"graph: {timeit(lambda i: graphs[i % 3].replay()):.4f} ms/step")
