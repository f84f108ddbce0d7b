import torch, time
dev = torch.device("cuda:0")
for mb in (4, 100, 400):
    h = torch.empty(mb * 1024 * 1024 // 4, dtype=torch.float32).pin_memory()
    d = torch.empty_like(h, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"H2D {mb} MB: {ms:.3f} ms  {mb / 1024 / (ms * 1e-3):.1f} GiB/s  {mb * 1.048576 / ms:.1f} GB/s")
    a.record()
    for _ in range(10):
        h.copy_(d, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"D2H {mb} MB: {ms:.3f} ms  {mb * 1.048576 / ms:.1f} GB/s")
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv"], capture_output=True, text=True).stdout)
print(subprocess.run(["nproc"], capture_output=True, text=True).stdout)
