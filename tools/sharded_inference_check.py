"""SURVEY 8e check: the per-image sharded path + one all-gather equals the single-GPU run bit for bit.

    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_inference_check.py

Every rank builds the same seeded batch, runs proposals -> head coordinates -> RoIPool on its contiguous
shard of the images (shard_bounds), the detections are all-gathered (NCCL); rank 0 also runs the whole batch
on its own GPU and compares: gathered RoIs / keep counts identical, its shard of the pooled features identical.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from two_stage_object_detection_b200 import functional as F  # noqa: E402
from two_stage_object_detection_b200.distributed import all_gather_detections, shard_bounds  # noqa: E402


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, C, H, W, P, S = 5, 24, 38, 38, 7, 600  # 5 images: an uneven split over 2 or 4 ranks
    N = H * W * 9
    g = torch.Generator().manual_seed(1234)
    loc = (torch.randn(B, N, 4, generator=g) * 0.2).to(dev)
    logits = torch.randn(B, N, 2, generator=g).to(dev)
    feat = torch.randn(B, C, H, W, generator=g).to(dev)
    base = F.base_anchors(device=dev)
    kw = dict(clip_x_max=S, clip_y_max=S, min_size=16.0, base=base, feat_stride=16, feat_hw=(H, W),
              score_is_logits=True, n_pre_nms=3000, n_post_nms=300, nms_iou=0.7)

    def run(lo, hi):
        rois, src, n_keep, status = F.proposals(loc[lo:hi], logits[lo:hi], **kw)
        idx = torch.arange(hi - lo, dtype=torch.int32, device=dev)
        rois5 = F.roi_head_coords(rois, idx, (S, S), (H, W))
        pooled = F.roi_pool_forward(feat[lo:hi], rois5, P, 1.0, rois_per_image=300)
        assert not status.cpu().numpy().any()
        return rois, n_keep, pooled

    lo, hi = shard_bounds(B, rank, world)
    rois, n_keep, pooled = run(lo, hi)
    all_rois, all_keep = all_gather_detections(rois, n_keep)
    if rank == 0:
        full_rois, full_keep, full_pooled = run(0, B)
        assert torch.equal(all_rois, full_rois), "gathered RoIs differ from the single-GPU run"
        assert torch.equal(all_keep, full_keep), "gathered keep counts differ"
        assert torch.equal(pooled, full_pooled[lo * 300:hi * 300]), "pooled features of the shard differ"
        print(f"sharded ok: world {world}, rois {tuple(all_rois.shape)}, shard of rank 0 = images [{lo},{hi})")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
