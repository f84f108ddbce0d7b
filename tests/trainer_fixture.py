"""Seeded inputs of the trainer parity case, shared by tests/golden/make_golden.py (which runs the
reference on them) and tests/test_gpu_parity.py (which runs the CUDA path on them): nothing large is
stored, both sides regenerate the tensors from the seed with torch's CPU generator."""
import torch

TRAINER_SEEDS = (701, 702)


def trainer_inputs(seed, feat_hw=(20, 20), n_gt=5, num_classes=20, size=320):
    """Everything the trainer parity test needs, regenerated from a seed (nothing big is stored):
    backbone features, RPN / head weights, GT boxes and labels."""
    g = torch.Generator().manual_seed(seed)
    feat = torch.relu(torch.randn(1, 512, *feat_hw, generator=g))
    w = dict(rpn_score_w=torch.randn(18, 512, 1, 1, generator=g) * 0.05, rpn_score_b=torch.zeros(18),
             rpn_loc_w=torch.randn(36, 512, 1, 1, generator=g) * 0.01, rpn_loc_b=torch.zeros(36),
             cls_loc_w=torch.randn((num_classes + 1) * 4, 512, generator=g) * 0.02,
             cls_loc_b=torch.zeros((num_classes + 1) * 4),
             score_w=torch.randn(num_classes + 1, 512, generator=g) * 0.05, score_b=torch.zeros(num_classes + 1))
    c = torch.rand(n_gt, 2, generator=g) * size
    wh = 40 + torch.rand(n_gt, 2, generator=g) * 120
    bbox = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, size)
    label = torch.randint(0, num_classes, (n_gt,), generator=g)
    return feat, w, bbox, label


def load_trainer_weights(trainer, w):
    with torch.no_grad():
        trainer.rpn.score.weight.copy_(w["rpn_score_w"]); trainer.rpn.score.bias.copy_(w["rpn_score_b"])
        trainer.rpn.loc.weight.copy_(w["rpn_loc_w"]); trainer.rpn.loc.bias.copy_(w["rpn_loc_b"])
        trainer.head.cls_loc.weight.copy_(w["cls_loc_w"]); trainer.head.cls_loc.bias.copy_(w["cls_loc_b"])
        trainer.head.score.weight.copy_(w["score_w"]); trainer.head.score.bias.copy_(w["score_b"])


