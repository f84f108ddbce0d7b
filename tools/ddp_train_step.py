"""Config-3 style training steps under DDP: one process per GPU, each rank owns its images, the hot path
(proposals, target creators, RoIPool fwd/bwd) runs in the sm_100a kernels, gradients are all-reduced by
stock DDP over NCCL, detections are all-gathered once per step.

    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/ddp_train_step.py [--steps 3]
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from two_stage_object_detection_b200.distributed import all_gather_detections  # noqa: E402
from two_stage_object_detection_b200.nets import FasterRCNNTrainer  # noqa: E402


class TinyExtractor(nn.Module):
    """Stand-in stride-16 backbone (the real ones are out of scope): 3 -> 512 channels."""

    def __init__(self):
        super().__init__()
        self.net = nn.Sequential(nn.Conv2d(3, 32, 3, 4, 1), nn.ReLU(), nn.Conv2d(32, 512, 3, 4, 1), nn.ReLU())

    def forward(self, x):
        return self.net(x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)  # same initial weights everywhere
    model = FasterRCNNTrainer("train", num_classes=20, extractor=TinyExtractor()).to(dev)
    ddp = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.AdamW(ddp.parameters(), lr=1e-3, weight_decay=1e-4)
    g = torch.Generator().manual_seed(100 + rank)  # different data per rank
    for step in range(args.steps):
        imgs = [torch.rand(3, 320, 320, generator=g) for _ in range(args.batch)]
        boxes, labels = [], []
        for _ in range(args.batch):
            c = torch.rand(4, 2, generator=g) * 320
            wh = 40 + torch.rand(4, 2, generator=g) * 120
            boxes.append(torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, 320))
            labels.append(torch.randint(0, 20, (4,), generator=g))
        losses, anchors_pred, *_ = ddp(imgs, boxes, labels)
        opt.zero_grad(set_to_none=True)
        losses[-1].backward()
        opt.step()
        dets, _ = all_gather_detections(anchors_pred.detach())
        total = losses[-1].detach().clone()
        if world > 1:
            dist.all_reduce(total)
        assert torch.isfinite(total), "loss is not finite"
        if rank == 0:
            print(f"step {step}: mean loss {float(total) / world:.4f}, gathered detections {tuple(dets.shape)}", flush=True)
    # DDP keeps the replicas identical
    if world > 1:
        flat = torch.cat([p.detach().flatten() for p in model.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(flat, ref), "replicas diverged"
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print("ddp ok", flush=True)


if __name__ == "__main__":
    main()
