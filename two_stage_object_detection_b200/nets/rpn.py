"""Drop-in for the reference's nets/rpn.py: ``ProposalCreator`` and ``RegionProposalNetwork`` with the
same constructor / call signatures and return values, backed by the batched sm_100a proposal
pipeline (csrc/proposals.cu).  The two 1x1 convolutions stay PyTorch/cuDNN as in the reference.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import _lib, config, functional as F
from ..utils.basic_anchors import generate_basic_anchor


class ProposalCreator():
    """nets/rpn.py:17-70.  ``mode == "train"`` selects the train limits, anything else the test ones."""

    def __init__(
        self,
        mode,
        nms_iou=0.7,
        n_train_pre_nms=12000,
        n_train_post_nms=600,
        n_test_pre_nms=3000,
        n_test_post_nms=300,
        min_size=16
    ):
        self.mode = mode
        self.nms_iou = nms_iou
        self.n_train_pre_nms = n_train_pre_nms
        self.n_train_post_nms = n_train_post_nms
        self.n_test_pre_nms = n_test_pre_nms
        self.n_test_post_nms = n_test_post_nms
        self.min_size = min_size
        # strict: raise the IndexError the reference raises when padding runs past the candidate
        # list (costs one tiny device->host read per call); batched() never synchronises.
        self.strict_reference = True

    def limits(self):
        if self.mode == "train":
            return self.n_train_pre_nms, self.n_train_post_nms
        return self.n_test_pre_nms, self.n_test_post_nms

    def batched(self, loc, score, img_size, scale=1., anchor=None, base=None, feat_stride=None, feat_hw=None,
                score_is_logits=False, layout="nhwc"):
        """All images at once: loc [B,N,4], score [B,N] (or logits [B,N,2]); with ``layout="nchw"`` the RPN conv
        outputs themselves, loc [B,4A,H,W] and logits [B,2A,H,W].
        Returns (rois [B,n_post,4], roi_src, n_keep, status) without host synchronisation."""
        n_pre, n_post = self.limits()
        return F.proposals(loc, score, clip_x_max=img_size[1], clip_y_max=img_size[2], n_pre_nms=n_pre,
                           n_post_nms=n_post, nms_iou=self.nms_iou, min_size=self.min_size * scale,
                           anchor=anchor, base=base, feat_stride=feat_stride, feat_hw=feat_hw,
                           score_is_logits=score_is_logits, layout=layout)

    def __call__(self, loc, score, anchor, img_size, scale=1.):
        rois, _, _, status = self.batched(loc.unsqueeze(0), score.reshape(1, -1), img_size, scale, anchor=anchor)
        if self.strict_reference and int(status[0].item()) & _lib.IMG_PAD_INDEX_ERROR:
            raise IndexError("padding index out of range for the pre-NMS proposal list "
                             "(the reference raises here too: nets/rpn.py:65-69)")
        return rois[0]


class RegionProposalNetwork(nn.Module):
    """nets/rpn.py:72-143.  forward(x, img_size, scale) -> (rpn_locs [n,N,4], rpn_scores [n,N,2],
    rois [n,n_post,4], anchor [1,N,4]); parameter names ``score`` / ``loc`` match the reference's
    checkpoints."""

    def __init__(
        self,
        in_channels=512,
        ratios=[0.5, 1, 2],
        anchor_scales=[8, 16, 32],
        feat_stride=16,
        mode="training",
    ):
        super(RegionProposalNetwork, self).__init__()
        self.anchor_base = generate_basic_anchor(anchor_scales=anchor_scales, ratios=ratios)
        n_anchor = self.anchor_base.shape[0]
        self.score = nn.Conv2d(in_channels, n_anchor * 2, 1, 1, 0)
        self.loc = nn.Conv2d(in_channels, n_anchor * 4, 1, 1, 0)
        self.feat_stride = feat_stride
        self.proposal_layer = ProposalCreator(mode)
        # fused_softmax: the decode kernel reads the conv outputs in place (NCHW) and computes softmax(...)[1]
        # itself: neither the permute(0,2,3,1).contiguous() passes (nets/rpn.py:107-113) nor the softmax + slice +
        # copy (:115-118) stand between the convolutions and the proposal layer.
        self.fused_softmax = True
        # The NHWC tensors rpn_locs / rpn_scores are RETURN VALUES (the trainer's losses read them).  They are
        # produced after the proposal kernels have been enqueued, and not at all when this is False
        # (inference callers that only want the RoIs get None in their place).
        self.return_rpn_outputs = True
        # nets/frcnn.py:37,48 unpacks five values (with roi_indices); nets/rpn.py:143 returns four.
        self.return_roi_indices = False
        self._anchor_cache = {}
        self.last_status = None

    def _anchors(self, h, w, device):
        key = (h, w, str(device))
        a = self._anchor_cache.get(key)
        if a is None:
            base = self.anchor_base.to(device)
            a = F.shifted_anchors(base, self.feat_stride, h, w)
            self._anchor_cache = {key: a}
        return a

    def forward(self, x, img_size, scale=1.):
        n, _, h, w = x.shape
        loc_map, score_map = self.loc(x), self.score(x)  # [n,4A,h,w], [n,2A,h,w] as cuDNN writes them
        base = self.anchor_base.to(x.device)
        rpn_locs = rpn_scores = None
        if self.fused_softmax:
            rois, _, _, status = self.proposal_layer.batched(
                loc_map, score_map, img_size, scale, base=base, feat_stride=self.feat_stride, layout="nchw")
        if self.return_rpn_outputs or not self.fused_softmax:
            rpn_locs = loc_map.permute(0, 2, 3, 1).contiguous().view(n, -1, 4)
            rpn_scores = score_map.permute(0, 2, 3, 1).contiguous().view(n, -1, 2)
        if not self.fused_softmax:  # the reference's own sequence of ops, kept for comparison
            score_in = torch.softmax(rpn_scores, dim=-1)[:, :, 1].contiguous()
            rois, _, _, status = self.proposal_layer.batched(
                rpn_locs, score_in, img_size, scale, base=base, feat_stride=self.feat_stride, feat_hw=(h, w))
        self.last_status = status  # per-image FRCNN_IMG_* flags, left on the device
        rois = rois.type_as(x)
        anchor = self._anchors(h, w, x.device).unsqueeze(0)
        if self.return_roi_indices:
            roi_indices = torch.arange(n, device=x.device, dtype=torch.float32).view(n, 1).expand(n, rois.shape[1])
            return rpn_locs, rpn_scores, rois, roi_indices, anchor
        return rpn_locs, rpn_scores, rois, anchor
