"""Time the cluster top-k sort in isolation."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from two_stage_object_detection_b200 import functional as F
dev = torch.device('cuda:0')
def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for B, N, k, dist in ((1, 12996, 3000, "int"), (16, 12996, 3000, "int"), (16, 12996, 3000, "prob"), (32, 22500, 3000, "prob"),
                      (32, 22500, 3000, "peaked"), (16, 12996, 12000, "int"), (8, 36864, 30000, "prob"), (16, 1600, 1600, "int")):
    g = torch.Generator().manual_seed(0)
    if dist == "int":
        keys = torch.randint(1 << 20, (1 << 31) - 1, (B, N), generator=g, dtype=torch.int64).to(torch.int32).to(dev)
    else:  # order-preserving keys of fp32 scores in (0, 1): softmax of random logits, or nearly all scores alike
        sc = torch.sigmoid(torch.randn(B, N, generator=g) * (2.0 if dist == "prob" else 0.01))
        keys = (sc.view(torch.int32) | (-2 ** 31)).to(dev)
    boxes = torch.rand(B, N, 4, generator=g).to(dev)
    t = timeit(lambda: F.topk_sorted(keys, boxes, k))
    print(f"B={B} N={N} k={k} {dist}: {t:.1f} us")
