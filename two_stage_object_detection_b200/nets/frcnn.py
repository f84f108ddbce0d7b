"""Drop-in for the reference's nets/frcnn.py inference wrapper (same constructor arguments, forward
modes and return arity).  The reference's own file is stale (it passes two positional 512s to the
RPN and unpacks five RPN outputs); this one wires the same pieces so that it actually runs.  The
backbone is out of scope of the hot path: pass ``extractor`` / ``classifier`` modules (the reference
hard-wires HarDNet-39)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .classify import HarNetRoIHead
from .rpn import RegionProposalNetwork


class GlobalAvgClassifier(nn.Module):
    """models/hardnet.py:203-212 (HarNetClassifier): global average pool + flatten."""

    def __init__(self):
        super().__init__()
        self.clssifier = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten())

    def forward(self, x):
        return self.clssifier(x)


class FasterRCNN(nn.Module):
    def __init__(self, num_classes,
                 mode="training",
                 feat_stride=16,
                 anchor_scales=[8, 16, 32],
                 ratios=[0.5, 1, 2],
                 extractor=None, classifier=None, feat_channels=512, roi_size=7, roi_op="pool"):
        super(FasterRCNN, self).__init__()
        self.feat_stride = feat_stride
        self.extractor = extractor if extractor is not None else nn.Identity()
        self.classifier = classifier if classifier is not None else GlobalAvgClassifier()
        self.rpn = RegionProposalNetwork(
            feat_channels,
            ratios=ratios,
            anchor_scales=anchor_scales,
            feat_stride=self.feat_stride,
            mode=mode
        )
        self.rpn.return_roi_indices = True
        self.head = HarNetRoIHead(
            n_class=num_classes + 1,
            roi_size=roi_size,
            spatial_scale=1,
            classifier=self.classifier,
            in_features=feat_channels,
            roi_op=roi_op,
        )

    @staticmethod
    def _image_index(roi_indices):
        # the RPN hands back one index per RoI [n,R]; the head wants one per image [n]
        if roi_indices is None:
            return None
        return roi_indices[:, 0].to(torch.int32) if roi_indices.dim() == 2 else roi_indices

    def forward(self, x, scale=1., mode="forward"):
        if mode == "forward":
            img_size = x.shape[2:]
            base_feature = self.extractor(x)
            # the proposal layer indexes img_size[1], img_size[2] (nets/rpn.py:47-48)
            _, _, rois, roi_indices, _ = self.rpn.forward(base_feature, (x.shape[1],) + tuple(img_size), scale)
            # our own RPN emits image i's RoIs in row i: tell the head (None) so it skips the bucketing pass
            roi_cls_locs, roi_scores = self.head.forward(base_feature, rois, None, img_size)
            return roi_cls_locs, roi_scores, rois, roi_indices
        elif mode == "extractor":
            return self.extractor.forward(x)
        elif mode == "rpn":
            base_feature, img_size = x
            return self.rpn.forward(base_feature, img_size, scale)
        elif mode == "head":
            base_feature, rois, roi_indices, img_size = x
            return self.head.forward(base_feature, rois, self._image_index(roi_indices), img_size)

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
