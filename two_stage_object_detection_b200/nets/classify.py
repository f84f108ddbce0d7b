"""Drop-in for the reference's nets/classify.py: the RoI head.  The RoI -> feature-map coordinate map
and the RoIPool / RoIAlign gather run in csrc/roi_ops.cu; the classifier module and the two Linear
layers stay PyTorch, with the reference's parameter names (``cls_loc``, ``score``)."""
from __future__ import annotations

import torch
from torch import nn

from .. import functional as F


def _pair2(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class RoIPool(nn.Module):
    """torchvision.ops.RoIPool stand-in (nets/classify.py:4,17): (input [B,C,H,W], rois [K,5])."""

    def __init__(self, output_size, spatial_scale):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale

    def forward(self, input, rois, rois_per_image=0):
        return F.roi_pool(input, rois, self.output_size, self.spatial_scale, rois_per_image)


class RoIAlign(nn.Module):
    """torchvision.ops.RoIAlign stand-in (the BASELINE RoIAlign 7x7 configuration)."""

    def __init__(self, output_size, spatial_scale, sampling_ratio=-1, aligned=False):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio
        self.aligned = aligned

    def forward(self, input, rois, rois_per_image=0):
        return F.roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned,
                           rois_per_image)


class HarNetRoIHead(nn.Module):
    """nets/classify.py:8-56.  forward(x [n,C,H,W], rois [n,R,4], roi_indices [n], img_size) ->
    (roi_cls_locs [n,R,4*n_class], roi_scores [n,R,n_class]).

    Differences from the reference, all opt-in or supersets: any R per image (the reference
    hard-codes 128 in an expand), ``in_features`` other than 512, ``roi_op="align"``, and
    ``roi_indices=None`` meaning "RoI row i belongs to image i" (what the reference's callers always
    pass, frcnn_training.py:290), which lets the gather skip its per-image bucketing pass."""

    def __init__(self, n_class, roi_size, spatial_scale, classifier, in_features=512, roi_op="pool",
                 sampling_ratio=-1, aligned=False):
        super().__init__()
        self.classifier = classifier
        self.cls_loc = nn.Linear(in_features, n_class * 4)
        self.score = nn.Linear(in_features, n_class)
        if roi_op == "pool":
            self.roi = RoIPool((roi_size, roi_size), spatial_scale)
        elif roi_op == "align":
            self.roi = RoIAlign((roi_size, roi_size), spatial_scale, sampling_ratio, aligned)
        else:
            raise ValueError("roi_op must be 'pool' or 'align'")
        self.fuse_mean = True  # set False to always materialise the pooled tensor

    def gather(self, x, rois, roi_indices, img_size):
        """The hot part: coordinate map + index concat + RoI gather -> [n*R, C, P, P]."""
        rois = rois.view(x.shape[0], -1, 4)
        grouped = 0
        if roi_indices is None:
            roi_indices = torch.arange(x.shape[0], dtype=torch.int32, device=x.device)
            grouped = rois.shape[1]
        indices_and_rois = F.roi_head_coords(rois, roi_indices, img_size, (x.size()[2], x.size()[3]))
        return self.roi(x, indices_and_rois, grouped)

    def _fused_mean_ok(self, x):
        """RoIPool -> classifier collapses to one kernel when the classifier is exactly the reference's
        HarNetClassifier (AdaptiveAvgPool2d(1) + Flatten, models/hardnet.py:203-212) and nothing needs the
        pooled tensor for a backward pass."""
        seq = getattr(self.classifier, "clssifier", None)
        return (self.fuse_mean and isinstance(self.roi, (RoIPool, RoIAlign)) and isinstance(seq, nn.Sequential)
                and len(seq) == 2
                and isinstance(seq[0], nn.AdaptiveAvgPool2d) and tuple(_pair2(seq[0].output_size)) == (1, 1)
                and isinstance(seq[1], nn.Flatten) and not (torch.is_grad_enabled() and x.requires_grad))

    def forward(self, x, rois, roi_indices, img_size):
        n = x.shape[0]
        if self._fused_mean_ok(x):
            rois_ = rois.view(n, -1, 4)
            grouped = 0
            if roi_indices is None:
                roi_indices = torch.arange(n, dtype=torch.int32, device=x.device)
                grouped = rois_.shape[1]
            r5 = F.roi_head_coords(rois_, roi_indices, img_size, (x.size()[2], x.size()[3]))
            if isinstance(self.roi, RoIPool):
                fc7 = F.roi_pool_mean(x, r5, self.roi.output_size, self.roi.spatial_scale, grouped)
            else:
                fc7 = F.roi_align_mean(x, r5, self.roi.output_size, self.roi.spatial_scale, self.roi.sampling_ratio,
                                       self.roi.aligned, grouped)
        else:
            pool = self.gather(x, rois, roi_indices, img_size)
            fc7 = self.classifier(pool)
        roi_cls_locs = self.cls_loc(fc7)
        roi_scores = self.score(fc7)
        roi_cls_locs = roi_cls_locs.view(n, -1, roi_cls_locs.size(1))
        roi_scores = roi_scores.view(n, -1, roi_scores.size(1))
        return roi_cls_locs, roi_scores
