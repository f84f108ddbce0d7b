"""Drop-in for the target creators of the reference's nets/frcnn_training.py (:19-177).  Each class
keeps the per-image call signature of the reference and adds ``batched`` for whole batches; the
work runs in csrc/targets.cu.  The reference's quirks are reproduced deliberately.
FasterRCNNTrainer (the loss glue, :179-345) is a caller of this path and is not part of it."""
from __future__ import annotations

import torch

from .. import _lib, functional as F


class AnchorTargetCreator():
    """nets/frcnn_training.py:19-103."""

    def __init__(self, n_sample=256, pos_iou_thresh=0.7, neg_iou_thresh=0.3, pos_ratio=0.5):
        self.n_sample = n_sample
        self.pos_iou_thresh = pos_iou_thresh
        self.neg_iou_thresh = neg_iou_thresh
        self.pos_ratio = pos_ratio

    def _kw(self):
        return dict(n_sample=self.n_sample, pos_iou_thresh=self.pos_iou_thresh,
                    neg_iou_thresh=self.neg_iou_thresh, pos_ratio=self.pos_ratio)

    def batched(self, bbox, n_gt, anchor=None, base=None, feat_stride=None, feat_hw=None):
        """bbox [B,Gmax,4], n_gt [B] -> (loc [B,N,4], label [B,N] int64); no host sync."""
        return F.anchor_targets(bbox, n_gt, anchor=anchor, base=base, feat_stride=feat_stride,
                                feat_hw=feat_hw, **self._kw())

    def __call__(self, bbox, anchor):
        dev = anchor.device
        g = bbox.shape[0]
        bb = bbox.reshape(1, g, 4) if g else torch.zeros((1, 1, 4), dtype=torch.float32, device=dev)
        n_gt = torch.tensor([g], dtype=torch.int32, device=dev)
        loc, label = self.batched(bb.to(dev), n_gt, anchor=anchor)
        return loc[0], label[0]


class ProposalTargetCreator(object):
    """nets/frcnn_training.py:105-177."""

    def __init__(self, n_sample=128, pos_ratio=0.5, pos_iou_thresh=0.5, neg_iou_thresh_high=0.5,
                 neg_iou_thresh_low=0):
        self.n_sample = n_sample
        self.pos_ratio = pos_ratio
        self.pos_roi_per_image = int(self.n_sample * self.pos_ratio)
        self.pos_iou_thresh = pos_iou_thresh
        self.neg_iou_thresh_high = neg_iou_thresh_high
        self.neg_iou_thresh_low = neg_iou_thresh_low

    def _kw(self):
        return dict(n_sample=self.n_sample, pos_ratio=self.pos_ratio, pos_iou_thresh=self.pos_iou_thresh,
                    neg_iou_thresh_high=self.neg_iou_thresh_high, neg_iou_thresh_low=self.neg_iou_thresh_low)

    def batched(self, roi, bbox, label, n_gt):
        """roi [B,R,4], bbox [B,Gmax,4], label [B,Gmax], n_gt [B] ->
        (sample_roi [B,S,4], gt_roi_loc [B,S,4], gt_roi_label [B,S], n_out [B], status [B])."""
        return F.proposal_targets(roi, bbox, label, n_gt, **self._kw())

    def __call__(self, roi, bbox, label, loc_normalize_std=(0.1, 0.1, 0.2, 0.2)):
        # loc_normalize_std is accepted and ignored, exactly as in the reference (:170 is commented out)
        dev = roi.device
        g = bbox.shape[0]
        bb = bbox.reshape(1, g, 4) if g else torch.zeros((1, 1, 4), dtype=torch.float32, device=dev)
        ll = label.reshape(1, g) if g else torch.zeros((1, 1), dtype=torch.int64, device=dev)
        n_gt = torch.tensor([g], dtype=torch.int32, device=dev)
        s, l, y, n_out, status = self.batched(roi.reshape(1, -1, 4), bb.to(dev), ll.to(dev), n_gt)
        n_out, status = (int(v) for v in torch.stack([n_out[0], status[0]]).tolist())
        if status & _lib.IMG_SCATTER_INDEX_ERROR:
            raise IndexError("negative-sample index out of range for the sampled labels "
                             "(the reference raises here too: nets/frcnn_training.py:175)")
        return s[0, :n_out], l[0, :n_out], y[0, :n_out].type_as(label)
