"""Drop-in for the target creators of the reference's nets/frcnn_training.py (:19-177).  Each class
keeps the per-image call signature of the reference and adds ``batched`` for whole batches; the
work runs in csrc/targets.cu.  The reference's quirks are reproduced deliberately.
FasterRCNNTrainer (the loss glue, :179-345) is the path's immediate caller; a batched PyTorch counterpart
over these kernels is provided at the bottom."""
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


class FasterRCNNTrainer(torch.nn.Module):
    """Batched counterpart of the reference's loss glue (nets/frcnn_training.py:179-345), the immediate
    caller of the hot path.  Same constructor arguments, same forward signature and return arity, same
    loss definitions -- but every image of the batch is processed (the reference only ever uses
    ``imgs[0]`` and crashes for a second image) and nothing loops in Python over images.

    The backbone is not part of this package: pass ``extractor`` (a module mapping [n,3,H,W] to
    [n,feat_channels,H/16,W/16]; the reference hard-wires HarDNet-39) and optionally ``classifier``.
    PyTorch does the convolutions, Linears and losses; anchors, proposals, both target creators and the
    RoI gather run in the sm_100a kernels.  RoIs and targets are detached."""

    def __init__(self, mode, num_classes, feat_stride=16, anchor_scales=[8, 16, 32], ratios=[0.5, 1, 2],
                 extractor=None, classifier=None, feat_channels=512):
        super().__init__()
        from .classify import HarNetRoIHead
        from .frcnn import GlobalAvgClassifier
        from .rpn import RegionProposalNetwork
        self.feat_stride = feat_stride
        self.rpn_sigma = 1
        self.roi_sigma = 1
        self.n_classes = num_classes
        self.anchor_target_creator = AnchorTargetCreator()
        self.proposal_target_creator = ProposalTargetCreator()
        self.feat_extra = extractor if extractor is not None else torch.nn.Identity()
        self.classifier = classifier if classifier is not None else GlobalAvgClassifier()
        self.rpn = RegionProposalNetwork(feat_channels, ratios=ratios, anchor_scales=anchor_scales,
                                         feat_stride=self.feat_stride, mode=mode)
        self.head = HarNetRoIHead(n_class=num_classes + 1, roi_size=7, spatial_scale=1, classifier=self.classifier,
                                  in_features=feat_channels)
        self.loc_normalize_std = [0.1, 0.1, 0.2, 0.2]
        self.strict_reference = True  # raise the reference's IndexErrors (one small device->host read)

    @staticmethod
    def _loc_loss(pred_loc, gt_loc, gt_label, sigma):
        """_fast_rcnn_loc_loss (:220-238) per image: smooth-L1 over rows with label > 0, summed and divided
        by the number of selected elements (NaN for an image without positives, as in the reference)."""
        pos = (gt_label > 0).unsqueeze(-1)
        s2 = sigma ** 2
        diff = torch.abs(gt_loc - pred_loc).float()
        loss = torch.where(diff < (1. / s2), 0.5 * s2 * diff ** 2, diff - 0.5 / s2)
        # select-then-sum like the reference (:222-223): a non-finite target on a row that is NOT selected
        # (bbox2loc against a zero-width GT gives -inf for every anchor assigned to it) must not reach the sum
        return torch.where(pos, loss, torch.zeros_like(loss)).sum((1, 2)) / (pos.sum((1, 2)) * 4)

    def forward(self, imgs, bboxes, labels, scale=1):
        from torch.nn import functional as TF
        dev = self.rpn.loc.weight.device
        x = imgs if torch.is_tensor(imgs) else torch.stack([i for i in imgs], dim=0)
        x = x.to(dev)
        n = x.shape[0]
        img_size = x.shape[1:]  # (C, H, W): the reference indexes it as such (:252)
        base_feature = self.feat_extra(x)
        rpn_locs, rpn_scores, rois, anchor = self.rpn.forward(x=base_feature, img_size=img_size, scale=scale)
        bb, ll, n_gt = F.pad_gt([b.to(dev) for b in bboxes], [l.to(dev) for l in labels], device=dev)
        h, w = base_feature.shape[2], base_feature.shape[3]
        gt_rpn_loc, gt_rpn_label = self.anchor_target_creator.batched(
            bb, n_gt, base=self.rpn.anchor_base.to(dev), feat_stride=self.feat_stride, feat_hw=(h, w))
        sample_rois, gt_roi_locs, gt_roi_labels, n_out, status = self.proposal_target_creator.batched(
            rois.detach(), bb, ll, n_gt)
        if self.strict_reference:
            flags = torch.cat([status, self.rpn.last_status]).cpu()
            if int(flags.max()) != 0:
                raise IndexError("the reference raises IndexError for this batch "
                                 "(nets/rpn.py:65-69 or nets/frcnn_training.py:175)")
        # RPN losses (:273-277), one value per image, averaged over the batch (:333-338)
        rpn_loc_loss = self._loc_loss(rpn_locs, gt_rpn_loc, gt_rpn_label, self.rpn_sigma)
        ce = TF.cross_entropy(rpn_scores.reshape(-1, 2), gt_rpn_label.reshape(-1), ignore_index=-1,
                              reduction="none").view(n, -1)
        rpn_cls_loss = ce.sum(1) / (gt_rpn_label >= 0).sum(1)
        # head on the sampled RoIs (:289-298); rows past n_out (fewer than n_sample candidates) are padding
        roi_cls_locs, roi_scores = self.head.forward(x=base_feature, rois=sample_rois, roi_indices=None,
                                                     img_size=img_size)
        n_sample = roi_cls_locs.size(1)
        valid = torch.arange(n_sample, device=dev).unsqueeze(0) < n_out.unsqueeze(1)
        lab = torch.where(valid, gt_roi_labels, torch.zeros_like(gt_roi_labels))
        roi_loc = roi_cls_locs.view(n, n_sample, -1, 4).gather(
            2, lab.view(n, n_sample, 1, 1).expand(n, n_sample, 1, 4)).squeeze(2)
        # :311-320 for the whole batch in one kernel: class-specific loc row -> loc2bbox, torch.max over classes
        anchors_pred, classes_score_pred, classes_pred = F.detection_decode(
            sample_rois, roi_cls_locs.detach(), roi_scores.detach(), lab)
        roi_loc_loss = self._loc_loss(roi_loc, gt_roi_locs, torch.where(valid, gt_roi_labels, -torch.ones_like(lab)),
                                      self.roi_sigma)
        ce2 = TF.cross_entropy(roi_scores.reshape(-1, roi_scores.size(2)),
                               torch.where(valid, gt_roi_labels, -torch.ones_like(lab)).reshape(-1),
                               ignore_index=-1, reduction="none").view(n, -1)
        roi_cls_loss = ce2.sum(1) / valid.sum(1)
        losses = [rpn_loc_loss.sum() / n, rpn_cls_loss.sum() / n, roi_loc_loss.sum() / n, roi_cls_loss.sum() / n]
        losses = losses + [sum(losses)]
        # ground truth as the reference returns it (boxes, labels + 1) in the padded batch layout [n,Gmax,...]:
        # rows past an image's n_gt are padding, box (0,0,0,0) with class id 0 = background, never a real class
        # (a calculate_metrics-style consumer must not count them); self.last_n_gt holds the true counts
        self.last_n_gt = n_gt
        real = torch.arange(ll.shape[1], device=dev).unsqueeze(0) < n_gt.unsqueeze(1)
        return losses, anchors_pred, classes_pred, classes_score_pred, bb, torch.where(real, ll + 1, torch.zeros_like(ll))
