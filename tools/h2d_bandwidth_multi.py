"""Concurrent pinned host -> device copy rate with one process per GPU (run under torchrun).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_bandwidth_multi.py

Every rank copies a 100 MB pinned slab (the size of one cfg2 step's inputs) to its own GPU, all ranks at the same
time (barrier-aligned), then each rank alone; prints per-rank GB/s and the aggregate.  Tells whether the
end-to-end leg of bench.py at N > 1 is bound by the host (memory / PCIe root complex) or by something we do."""
import os, sys, json
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
MB = 100
h = torch.empty(MB * 1000 * 1000 // 4, dtype=torch.float32).pin_memory()
h.fill_(1.0)
d = torch.empty_like(h, device=dev)

def rate(n=10):
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        d.copy_(h, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    return MB * 1e6 * n / (a.elapsed_time(b) * 1e-3) / 1e9

together = rate()
alone = []
for r in range(world):
    if world > 1:
        dist.barrier()
    v = rate() if r == rank else 0.0
    alone.append(v)
    if world > 1:
        dist.barrier()
t = torch.tensor([together, alone[rank]], device=dev, dtype=torch.float64)
if world > 1:
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
else:
    out = [t]
if rank == 0:
    tog = [float(o[0]) for o in out]
    alo = [float(o[1]) for o in out]
    print(json.dumps({"ranks": world, "mb_per_copy": MB, "concurrent_gbs_per_rank": [round(v, 1) for v in tog],
                      "concurrent_gbs_total": round(sum(tog), 1), "alone_gbs_per_rank": [round(v, 1) for v in alo],
                      "host_cpus": len(os.sched_getaffinity(0))}))
if world > 1:
    dist.destroy_process_group()
