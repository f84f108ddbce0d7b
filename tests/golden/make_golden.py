"""Generate tests/golden/*.npz by running the REAL reference (read-only at /root/reference) and the
torchvision 0.26.0 CPU kernels it depends on.  Runs only in the build container (the GPU box has
no /root/reference); the resulting fixtures are committed and are what pins the oracle.

    python tests/golden/make_golden.py

Seeds are fixed; re-running reproduces the files byte-for-byte up to npz metadata.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

import torchvision  # noqa: E402
from torchvision.ops import nms as tv_nms, roi_align as tv_roi_align, roi_pool as tv_roi_pool  # noqa: E402

import utils.basic_anchors as ref_anchors  # noqa: E402
import utils.loc_bbox_iou as ref_box  # noqa: E402
import nets.rpn as ref_rpn  # noqa: E402
import nets.classify as ref_cls  # noqa: E402
import nets.frcnn_training as ref_tr  # noqa: E402

ref_anchors.device = ref_rpn.device = ref_tr.device = "cpu"
torch.set_num_threads(8)


def npy(t):
    return t.detach().cpu().numpy()


def save(name, **arrs):
    meta = dict(torch=torch.__version__, torchvision=torchvision.__version__)
    np.savez_compressed(os.path.join(OUT, name + ".npz"),
                        _meta=np.array(repr(meta)), **{k: np.asarray(v) for k, v in arrs.items()})
    sz = os.path.getsize(os.path.join(OUT, name + ".npz"))
    print(f"  wrote {name}.npz  {sz/1024:.0f} KiB")


# ------------------------------------------------------------------------------------------------
def gen_anchors():
    base = ref_anchors.generate_basic_anchor()
    base2 = ref_anchors.generate_basic_anchor(base_size=16, ratios=[0.5, 1, 2, 3], anchor_scales=[4, 8])
    sh_small = ref_anchors.enumerate_shifted_anchor(base, 16, 5, 7)
    sh_full = ref_anchors.enumerate_shifted_anchor(base, 16, 38, 38)
    sh_rect = ref_anchors.enumerate_shifted_anchor(base2, 8, 9, 4)
    a = torch.tensor([[100, 100, 200, 200]], dtype=torch.float32)
    b = torch.tensor([[150, 150, 250, 250]], dtype=torch.float32)
    save("anchors_kat", base=npy(base), base2=npy(base2), shifted_5x7=npy(sh_small),
         shifted_38x38=npy(sh_full), shifted_rect=npy(sh_rect),
         kat_iou=npy(ref_box.bbox_iou(a, b)), kat_loc=npy(ref_box.bbox2loc(a, b)),
         kat_roundtrip=npy(ref_box.loc2bbox(a, ref_box.bbox2loc(a, b))))


def rand_boxes(g, n, size, wmin=4.0, wmax=None):
    wmax = wmax or size / 2
    cx = torch.rand(n, generator=g) * size
    cy = torch.rand(n, generator=g) * size
    w = wmin + torch.rand(n, generator=g) * (wmax - wmin)
    h = wmin + torch.rand(n, generator=g) * (wmax - wmin)
    return torch.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).float()


def gen_boxmath():
    g = torch.Generator().manual_seed(101)
    src = rand_boxes(g, 500, 600)
    loc = torch.randn(500, 4, generator=g) * 0.3
    loc8 = torch.randn(500, 8, generator=g) * 0.3
    dst = rand_boxes(g, 500, 600)
    gts = rand_boxes(g, 7, 600, 50, 250)
    # degenerate rows: zero-width src, identical boxes, inverted dst
    src_d = src.clone()
    src_d[0, 2] = src_d[0, 0]
    src_d[1] = dst[1]
    dst_d = dst.clone()
    dst_d[2, [0, 2]] = dst_d[2, [2, 0]]
    save("boxmath", src=npy(src), loc=npy(loc), loc8=npy(loc8), dst=npy(dst), gts=npy(gts),
         src_d=npy(src_d), dst_d=npy(dst_d),
         decode=npy(ref_box.loc2bbox(src, loc)), decode8=npy(ref_box.loc2bbox(src, loc8)),
         encode=npy(ref_box.bbox2loc(src, dst)), encode_d=npy(ref_box.bbox2loc(src_d, dst_d)),
         iou=npy(ref_box.bbox_iou(src, gts)), iou_self=npy(ref_box.bbox_iou(src[:64], src_d[:64])))


# ------------------------------------------------------------------------------------------------
def proposal_case(name, H, W, mode, loc_std, seed, img=(3, 600, 600), scale=1.0, quantize=None,
                  limits=None, min_size=16, nms_iou=0.7, expect_error=False, score_kind="softmax"):
    g = torch.Generator().manual_seed(seed)
    base = ref_anchors.generate_basic_anchor()
    anchor = ref_anchors.enumerate_shifted_anchor(base, 16, H, W)
    N = anchor.shape[0]
    loc = (torch.randn(N, 4, generator=g) * loc_std).float()
    if score_kind == "softmax":
        score = torch.softmax(torch.randn(N, 2, generator=g), -1)[:, 1].contiguous()
    else:
        score = torch.rand(N, generator=g)
    if quantize:
        score = torch.round(score * quantize) / quantize
    kw = dict(limits or {})
    pc = ref_rpn.ProposalCreator(mode, nms_iou=nms_iou, min_size=min_size, **kw)
    n_pre, n_post = ((pc.n_train_pre_nms, pc.n_train_post_nms) if mode == "train"
                     else (pc.n_test_pre_nms, pc.n_test_post_nms))
    # stage by stage with the reference's own callables / the same torch ops it uses
    decoded = ref_box.loc2bbox(anchor, loc)
    roi = decoded.clone()
    roi[:, [0, 2]] = torch.clamp(roi[:, [0, 2]], min=0, max=img[1])
    roi[:, [1, 3]] = torch.clamp(roi[:, [1, 3]], min=0, max=img[2])
    ms = min_size * scale
    keep = torch.where(((roi[:, 2] - roi[:, 0]) >= ms) & ((roi[:, 3] - roi[:, 1]) >= ms))[0]
    s_valid = score[keep]
    ties = int(s_valid.numel() - torch.unique(s_valid).numel())
    order = torch.argsort(s_valid, descending=True, stable=True)
    if n_pre > 0:
        order = order[:n_pre]
    roi_sorted = roi[keep][order]
    nms_keep = tv_nms(roi_sorted, s_valid[order], nms_iou)
    arrs = dict(H=H, W=W, img_size=np.array(img), scale=scale, mode=np.array(mode), n_pre=n_pre,
                n_post=n_post, min_size=min_size, nms_iou=nms_iou, loc=npy(loc), score=npy(score),
                decoded=npy(decoded), valid_idx=npy(keep).astype(np.int32),
                order=npy(order).astype(np.int32), nms_keep=npy(nms_keep).astype(np.int32),
                ties=ties, expect_error=int(expect_error))
    err = False
    try:
        if ties:
            # tie order of torch.argsort(descending=True) is unspecified: run the
            # reference with a stable argsort substituted, as the survey prescribes
            orig = torch.argsort
            torch.argsort = lambda x, descending=False, **k: orig(x, descending=descending, stable=True)
        try:
            out = pc(loc, score, anchor, img, scale=scale)
        finally:
            if ties:
                torch.argsort = orig
        arrs["roi"] = npy(out)
    except IndexError:
        err = True
    assert err == expect_error, (name, err)
    if not err:
        # cross-check the staged pipeline against the end-to-end reference call
        k = nms_keep
        if len(k) < n_post:
            k = torch.cat([k, torch.arange(n_post - len(k))])
        assert torch.equal(roi_sorted[k[:n_post]], out), name
    print(f"  {name}: N={N} valid={keep.numel()} ties={ties} n_sel={order.numel()} "
          f"nms_keep={nms_keep.numel()} err={err}")
    save(name, **arrs)


def gen_proposals():
    proposal_case("proposal_small_train", 10, 12, "train", 0.2, 1, img=(3, 160, 192),
                  limits=dict(n_train_pre_nms=600, n_train_post_nms=100))
    proposal_case("proposal_small_overlap", 10, 12, "test", 0.05, 2, img=(3, 192, 192),
                  limits=dict(n_test_pre_nms=800, n_test_post_nms=60), nms_iou=0.5)
    proposal_case("proposal_small_ties", 12, 12, "train", 0.15, 3, img=(3, 192, 192), quantize=64,
                  limits=dict(n_train_pre_nms=700, n_train_post_nms=128))
    proposal_case("proposal_small_pad", 8, 8, "test", 0.02, 4, img=(3, 128, 128),
                  limits=dict(n_test_pre_nms=400, n_test_post_nms=200), nms_iou=0.1)
    proposal_case("proposal_small_error", 3, 3, "test", 0.02, 5, img=(3, 48, 48),
                  limits=dict(n_test_pre_nms=400, n_test_post_nms=300), nms_iou=0.05,
                  expect_error=True)
    proposal_case("proposal_small_scale", 9, 14, "anything", 0.3, 6, img=(3, 224, 144), scale=1.5,
                  limits=dict(n_test_pre_nms=300, n_test_post_nms=50), min_size=12)
    proposal_case("proposal_600_test", 38, 38, "test", 0.2, 7)
    proposal_case("proposal_600_train", 38, 38, "train", 0.2, 8)


def gen_proposals_large():
    """The BASELINE configs 4 and 5 at their full sizes through the reference's ProposalCreator:
    800x800 (N = 22 500, 3 000 -> 300) and the proposal-stress size 1024x1024 (N = 36 864, 30 000 -> 2 000,
    ~4 s of torchvision CPU nms).  loc std 0.1 keeps enough overlap that NMS needs several thousand candidates
    to find its 2 000 boxes."""
    proposal_case("proposal_800_test", 50, 50, "test", 0.2, 21, img=(3, 800, 800))
    proposal_case("proposal_1024_stress", 64, 64, "test", 0.1, 22, img=(3, 1024, 1024),
                  limits=dict(n_test_pre_nms=30000, n_test_post_nms=2000))


def gen_nms():
    g = torch.Generator().manual_seed(202)
    arrs = {}
    for i, (n, thr, kind) in enumerate([(300, 0.7, "rand"), (1000, 0.5, "ties"), (257, 0.3, "dup"),
                                        (64, 0.7, "zero"), (1, 0.7, "rand"), (2500, 0.7, "dense")]):
        if kind == "dense":
            b = rand_boxes(g, n, 200, 30, 120)
        else:
            b = rand_boxes(g, n, 400, 8, 150)
        s = torch.rand(n, generator=g)
        if kind == "ties":
            s = torch.round(s * 20) / 20
        if kind == "dup":
            b[n // 2:] = b[: n - n // 2].clone()
            s[n // 2:] = s[: n - n // 2].clone()
        if kind == "zero":
            b[::3, 2] = b[::3, 0]
            b[1::7] = 0
        k = tv_nms(b, s, thr)
        arrs[f"boxes{i}"] = npy(b)
        arrs[f"scores{i}"] = npy(s)
        arrs[f"thr{i}"] = thr
        arrs[f"keep{i}"] = npy(k).astype(np.int32)
        print(f"  nms case {i} ({kind}): n={n} keep={k.numel()}")
    arrs["n_cases"] = 6
    save("nms", **arrs)


# ------------------------------------------------------------------------------------------------
def gen_anchor_targets():
    base = ref_anchors.generate_basic_anchor()
    arrs = {}
    cases = [("small", 10, 12, 8, 192), ("one_gt", 10, 12, 1, 192), ("no_gt", 6, 6, 0, 96),
             ("full", 38, 38, 8, 600), ("many_pos", 16, 16, 40, 256), ("dup_gt", 10, 12, 6, 192)]
    for ci, (name, H, W, G, S) in enumerate(cases):
        g = torch.Generator().manual_seed(300 + ci)
        anchor = ref_anchors.enumerate_shifted_anchor(base, 16, H, W)
        if name == "many_pos":
            # GT boxes that coincide with anchors -> >128 positives, exercises the first-k cap
            idx = torch.randperm(anchor.shape[0], generator=g)[:G]
            bbox = anchor[idx].clone()
            bbox = torch.cat([bbox, bbox + 1.0, bbox - 1.0, bbox + 2.0, bbox - 2.0])
        else:
            bbox = rand_boxes(g, G, S, 40, S / 2).clamp(0, S) if G else torch.zeros(0, 4)
        if name == "dup_gt":
            bbox[3] = bbox[1]  # two GTs share their best anchor: later GT wins
            bbox[5] = bbox[0]
        atc = ref_tr.AnchorTargetCreator()
        loc, label = atc(bbox, anchor)
        arrs[f"{name}_H"] = H
        arrs[f"{name}_W"] = W
        arrs[f"{name}_bbox"] = npy(bbox)
        arrs[f"{name}_loc"] = npy(loc)
        arrs[f"{name}_label"] = npy(label).astype(np.int8)
        print(f"  anchor_targets {name}: N={anchor.shape[0]} G={bbox.shape[0]} pos={(label==1).sum().item()} "
              f"neg={(label==0).sum().item()} ign={(label==-1).sum().item()}")
    # non-default hyper-parameters incl. the n_neg <= 0 slice quirk
    g = torch.Generator().manual_seed(399)
    anchor = ref_anchors.enumerate_shifted_anchor(base, 16, 10, 12)
    bbox = rand_boxes(g, 12, 192, 40, 96).clamp(0, 192)
    atc = ref_tr.AnchorTargetCreator(n_sample=8, pos_iou_thresh=0.5, neg_iou_thresh=0.2, pos_ratio=1.0)
    loc, label = atc(bbox, anchor)
    arrs["custom_H"], arrs["custom_W"] = 10, 12
    arrs["custom_bbox"], arrs["custom_loc"] = npy(bbox), npy(loc)
    arrs["custom_label"] = npy(label).astype(np.int8)
    arrs["custom_params"] = np.array([8, 0.5, 0.2, 1.0])
    print(f"  anchor_targets custom: pos={(label==1).sum().item()} neg={(label==0).sum().item()}")
    arrs["names"] = np.array([c[0] for c in cases])
    save("anchor_targets", **arrs)


def gen_proposal_targets():
    arrs = {}
    names = []

    def case(name, roi, bbox, label, expect_error=None, **kw):
        ptc = ref_tr.ProposalTargetCreator(**kw)
        names.append(name)
        arrs[f"{name}_roi"], arrs[f"{name}_bbox"] = npy(roi), npy(bbox)
        arrs[f"{name}_label"] = npy(label)
        arrs[f"{name}_params"] = np.array([kw.get("n_sample", 128), kw.get("pos_ratio", 0.5),
                                           kw.get("pos_iou_thresh", 0.5),
                                           kw.get("neg_iou_thresh_high", 0.5),
                                           kw.get("neg_iou_thresh_low", 0)], dtype=np.float64)
        try:
            s, l, y = ptc(roi, bbox, label)
            assert not expect_error, name
            arrs[f"{name}_sample_roi"], arrs[f"{name}_gt_loc"] = npy(s), npy(l)
            arrs[f"{name}_gt_label"] = npy(y)
            arrs[f"{name}_error"] = 0
            print(f"  proposal_targets {name}: R={roi.shape[0]} G={bbox.shape[0]} out={s.shape[0]} "
                  f"labels>0={(y>0).sum().item()}")
        except IndexError:
            assert expect_error is not False, name
            arrs[f"{name}_error"] = 1
            print(f"  proposal_targets {name}: IndexError (reference raises)")

    g = torch.Generator().manual_seed(400)
    bbox = rand_boxes(g, 8, 600, 50, 250).clamp(0, 600)
    label = torch.randint(0, 20, (8,), generator=g)
    roi = rand_boxes(g, 600, 600, 16, 300).clamp(0, 600)
    case("rand600", roi, bbox, label, expect_error=False)
    # proposals concentrated on the GTs: many positives (first 64 kept) and the stale-index scatter
    jit = bbox[torch.randint(0, 8, (300,), generator=g)] + torch.randn(300, 4, generator=g) * 6
    roi2 = torch.cat([jit[:40], roi[:200], jit[40:90], roi[200:260]])
    case("mixed", roi2, bbox, label)
    case("no_gt", roi[:300], torch.zeros(0, 4), torch.zeros(0, dtype=torch.int64))
    case("tiny", roi[:20], bbox[:2], label[:2])
    case("lowthr", roi[:150], bbox, label, neg_iou_thresh_low=0.05)
    case("lowthr_ok", roi, bbox, label, neg_iou_thresh_low=0.01)
    case("small_sample", roi2[:200], bbox, label, n_sample=32, pos_ratio=0.25)
    # positives first, 200 of them: neg indices >= len(keep) -> IndexError in the reference
    jit2 = bbox[torch.randint(0, 8, (200,), generator=g)] + torch.randn(200, 4, generator=g) * 2
    case("scatter_error", torch.cat([jit2, roi[:300]]), bbox, label, expect_error=True)
    arrs["names"] = np.array(names)
    save("proposal_targets", **arrs)


# ------------------------------------------------------------------------------------------------
def gen_roi():
    g = torch.Generator().manual_seed(500)
    feat = torch.randn(2, 8, 20, 24, generator=g)
    feat[0, 0, 3:6, 4:9] = 0.75  # plateaus: first-max-wins argmax
    feat[1, 3] = torch.relu(feat[1, 3])
    K = 60
    r = rand_boxes(g, K, 24, 1, 16)
    bi = torch.randint(0, 2, (K, 1), generator=g).float()
    rois = torch.cat([bi, r], 1)
    rois[0, 1:] = torch.tensor([-5.0, -3.0, 4.0, 6.0])       # partly outside (negative)
    rois[1, 1:] = torch.tensor([20.0, 15.0, 40.0, 30.0])     # beyond the map
    rois[2, 1:] = torch.tensor([10.0, 10.0, 5.0, 4.0])       # inverted
    rois[3, 1:] = torch.tensor([100.0, 100.0, 120.0, 130.0])  # fully outside
    rois[4, 1:] = torch.tensor([2.5, 3.5, 2.5, 3.5])         # degenerate point, .5 rounding
    rois[5, 1:] = torch.tensor([0.0, 0.0, 23.0, 19.0])       # whole map
    arrs = dict(feat=npy(feat), rois=npy(rois))
    for P in (7, 14, 3):
        for sc in (1.0, 0.5):
            out = tv_roi_pool(feat, rois, (P, P), sc)
            arrs[f"pool_P{P}_s{sc}"] = npy(out)
    out, am = torch.ops.torchvision.roi_pool(feat, rois, 1.0, 7, 7)
    arrs["pool_argmax_P7_s1.0"] = npy(am).astype(np.int32)
    assert torch.equal(out, tv_roi_pool(feat, rois, (7, 7), 1.0))
    for P in (7, 2):
        for sr in (-1, 2, 3):
            for al in (False, True):
                for sc in (1.0, 0.25):
                    out = tv_roi_align(feat, rois, (P, P), sc, sr, al)
                    arrs[f"align_P{P}_sr{sr}_al{int(al)}_s{sc}"] = npy(out)
    save("roi_ops", **arrs)

    # the RoI head's coordinate map + gather (nets/classify.py:19-56), 128 RoIs as the reference needs
    import models.hardnet as ref_hd
    torch.manual_seed(501)
    head = ref_cls.HarNetRoIHead(n_class=3, roi_size=7, spatial_scale=1, classifier=ref_hd.HarNetClassifier())
    x = torch.randn(1, 512, 10, 12, generator=g)
    rr = rand_boxes(g, 128, 160, 10, 100).clamp(0, 192).view(1, 128, 4)
    captured = {}
    h = head.roi.register_forward_hook(lambda m, i, o: captured.update(rois5=i[1].detach().clone(), pool=o.detach().clone()))
    outs = {}
    for tag, img in (("chw", (3, 160, 192)), ("hw", (160, 192))):
        with torch.no_grad():
            locs, scores = head(x, rr, torch.zeros(1, dtype=torch.int32), img)
        outs[f"{tag}_img"] = np.array(img)
        outs[f"{tag}_rois5"] = npy(captured["rois5"])
        outs[f"{tag}_pool_sum"] = npy(captured["pool"].double().sum((2, 3))).astype(np.float64)
        outs[f"{tag}_pool_head"] = npy(captured["pool"][:16, :32])
        outs[f"{tag}_cls_locs"], outs[f"{tag}_scores"] = npy(locs), npy(scores)
    h.remove()
    save("roi_head", x=npy(x[:, :, :, :]).astype(np.float32), rois=npy(rr),
         w_loc=npy(head.cls_loc.weight), b_loc=npy(head.cls_loc.bias),
         w_score=npy(head.score.weight), b_score=npy(head.score.bias), **outs)


def gen_roi_large():
    """torchvision roi_pool 7x7 / 14x14 and roi_align 7x7 (sampling_ratio 2) on the feature-map sizes of the
    BASELINE configs (38x38, 50x50, 64x64; C = 8 keeps the fixture small), RoIs in feature coordinates sized
    like proposals (1 px .. most of the map, some hanging over the border)."""
    arrs = {}
    for H in (38, 50, 64):
        g = torch.Generator().manual_seed(700 + H)
        feat = torch.randn(2, 8, H, H, generator=g)
        feat[1] = torch.relu(feat[1])
        K = 48
        r = rand_boxes(g, K, H, 1.0, H * 0.9)
        r[::9] += 3.0    # some hang over the right / bottom border
        r[4::9] -= 3.0   # ... or the left / top one
        bi = torch.randint(0, 2, (K, 1), generator=g).float()
        rois = torch.cat([bi, r], 1)
        arrs[f"feat{H}"], arrs[f"rois{H}"] = npy(feat), npy(rois)
        for P in (7, 14):
            out, am = torch.ops.torchvision.roi_pool(feat, rois, 1.0, P, P)
            assert torch.equal(out, tv_roi_pool(feat, rois, (P, P), 1.0))
            arrs[f"pool{H}_P{P}"] = npy(out)
            arrs[f"argmax{H}_P{P}"] = npy(am).astype(np.int32)
        arrs[f"align{H}_P7_sr2"] = npy(tv_roi_align(feat, rois, (7, 7), 1.0, 2, False))
        arrs[f"align{H}_P7_sr2_al"] = npy(tv_roi_align(feat, rois, (7, 7), 1.0, 2, True))
        print(f"  roi_large {H}x{H}: K={K}")
    save("roi_large", **arrs)


def gen_roi_backward():
    """Gradients w.r.t. the features through torchvision.ops.roi_pool / roi_align (CPU autograd): the oracle
    of frcnn_roi_pool_backward / frcnn_roi_align_backward.  Same feature map and RoIs as roi_ops.npz (border,
    inverted, outside, degenerate RoIs included)."""
    src = np.load(os.path.join(OUT, "roi_ops.npz"))
    feat0, rois = torch.from_numpy(src["feat"]), torch.from_numpy(src["rois"])
    g = torch.Generator().manual_seed(800)
    arrs = {}
    for P in (7, 14):
        feat = feat0.clone().requires_grad_(True)
        out = tv_roi_pool(feat, rois, (P, P), 1.0)
        go = torch.randn(out.shape, generator=g)
        (gi,) = torch.autograd.grad(out, feat, go)
        arrs[f"pool_P{P}_go"], arrs[f"pool_P{P}_gi"] = npy(go), npy(gi)
    for sr in (2, -1):
        for al in (False, True):
            for sc in (1.0, 0.5):
                feat = feat0.clone().requires_grad_(True)
                out = tv_roi_align(feat, rois, (7, 7), sc, sr, al)
                go = torch.randn(out.shape, generator=g)
                (gi,) = torch.autograd.grad(out, feat, go)
                tag = f"align_sr{sr}_al{int(al)}_s{sc}"
                arrs[f"{tag}_go"], arrs[f"{tag}_gi"] = npy(go), npy(gi)
    save("roi_backward", **arrs)


def gen_rpn_forward():
    """RegionProposalNetwork.forward post-conv glue (nets/rpn.py:107-143) on a tiny feature map."""
    torch.manual_seed(600)
    rpn = ref_rpn.RegionProposalNetwork(in_channels=16, mode="test")
    rpn.proposal_layer = ref_rpn.ProposalCreator("test", n_test_pre_nms=500, n_test_post_nms=40)
    x = torch.randn(1, 16, 9, 11)
    with torch.no_grad():
        # large-ish loc/score weights so proposals are non-trivial
        rpn.loc.weight.mul_(3.0)
        rpn.score.weight.mul_(4.0)
        locs, scores, rois, anchor = rpn(x, (3, 144, 176), 1.0)
    fg = torch.softmax(scores, -1)[:, :, 1]
    assert torch.unique(fg).numel() == fg.numel(), "tie-free fixture expected" 
    print(f"  rpn_forward: rois {tuple(rois.shape)} unique fg {torch.unique(fg).numel()}/{fg.numel()}")
    save("rpn_forward", x=npy(x), w_loc=npy(rpn.loc.weight), b_loc=npy(rpn.loc.bias),
         w_score=npy(rpn.score.weight), b_score=npy(rpn.score.bias), img_size=np.array((3, 144, 176)),
         rpn_locs=npy(locs), rpn_scores=npy(scores), fg=npy(fg), rois=npy(rois), anchor=npy(anchor))


sys.path.insert(0, os.path.dirname(OUT))
from trainer_fixture import TRAINER_SEEDS, load_trainer_weights, trainer_inputs  # noqa: E402


def gen_trainer():
    """FasterRCNNTrainer.forward (nets/frcnn_training.py:240-345), batch of one (all the reference can do),
    with the HarDNet extractor replaced by a fixed feature tensor."""
    class Fixed(torch.nn.Module):
        def __init__(self, feat):
            super().__init__()
            self.feat = feat

        def forward(self, x):
            return self.feat

    arrs = {"seeds": np.array(TRAINER_SEEDS)}
    trainer = ref_tr.FasterRCNNTrainer(mode="train", num_classes=20)
    for seed in TRAINER_SEEDS:
        feat, w, bbox, label = trainer_inputs(seed)
        load_trainer_weights(trainer, w)
        trainer.feat_extra = Fixed(feat)
        img = torch.zeros(3, 320, 320)
        losses, anchors_pred, classes_pred, classes_score_pred, gt_b, gt_l = trainer([img], [bbox], [label])
        arrs[f"s{seed}_losses"] = np.array([float(l) for l in losses], dtype=np.float64)
        arrs[f"s{seed}_anchors_pred"] = npy(anchors_pred)
        arrs[f"s{seed}_classes_pred"] = npy(classes_pred)
        arrs[f"s{seed}_gt_label"] = npy(gt_l)
        print(f"  trainer seed {seed}: losses {[round(float(l), 5) for l in losses]}")
    save("trainer", **arrs)


def gen_detections():
    """Post-head step: class-specific loc gather + loc2bbox + torch.max over the class scores
    exactly as nets/frcnn_training.py:311-320 does them, then the evaluator's per-class NMS
    (`where(classes_pred == c)` -> torchvision nms, :441-454) and multi_inference.py:84's class-agnostic
    nms(iou_threshold=0.1)."""
    g = torch.Generator().manual_seed(909)
    arrs = {}
    cases = [(128, 21, 320.0, "rand"), (300, 21, 600.0, "ties"), (600, 5, 600.0, "dense"), (1, 3, 100.0, "rand"),
             (1024, 21, 800.0, "rand"), (800, 2, 300.0, "dense")]
    for i, (R, C, size, kind) in enumerate(cases):
        roi = rand_boxes(g, R, size, 16, size / 3 if kind != "dense" else size / 1.5)
        cls_loc = torch.randn(R, C * 4, generator=g) * 0.3
        score = torch.randn(R, C, generator=g)
        if kind == "ties":
            score = torch.round(score * 2) / 2          # exact ties between classes and between RoIs
        label = torch.randint(0, C, (R,), generator=g)
        # nets/frcnn_training.py:311-320
        n_sample = R
        roi_cls_loc = cls_loc.view(n_sample, -1, 4)
        roi_loc = roi_cls_loc[torch.arange(0, n_sample).type_as(label), label]
        boxes = ref_box.loc2bbox(roi, roi_loc)
        cls_score_pred, cls_index_pred = torch.max(score, dim=1)
        # the same with the predicted class instead of the ground-truth one (plain inference)
        roi_loc_p = roi_cls_loc[torch.arange(0, n_sample), cls_index_pred]
        boxes_p = ref_box.loc2bbox(roi, roi_loc_p)
        arrs[f"roi{i}"] = npy(roi)
        arrs[f"cls_loc{i}"] = npy(cls_loc)
        arrs[f"score{i}"] = npy(score)
        arrs[f"label{i}"] = npy(label)
        arrs[f"boxes{i}"] = npy(boxes)
        arrs[f"boxes_pred_class{i}"] = npy(boxes_p)
        arrs[f"cls_score{i}"] = npy(cls_score_pred)
        arrs[f"cls_index{i}"] = npy(cls_index_pred)
        # evaluator: per class, nets/frcnn_training.py:441-454 (nms_iou_threshold default 0.7)
        keep_all = []
        for c in range(C):
            idx = torch.where(cls_index_pred == c)[0]
            k = tv_nms(boxes[idx], cls_score_pred[idx], 0.7)
            keep_all.append(idx[k])
            arrs[f"keep{i}_c{c}"] = npy(idx[k]).astype(np.int32)
            arrs[f"keep03_{i}_c{c}"] = npy(idx[tv_nms(boxes[idx], cls_score_pred[idx], 0.3)]).astype(np.int32)
        # multi_inference.py:84
        arrs[f"keep_agnostic{i}"] = npy(tv_nms(boxes, cls_score_pred, 0.1)).astype(np.int32)
        print(f"  detections case {i} ({kind}): R={R} C={C} kept {sum(len(k) for k in keep_all)} per class, "
              f"{len(arrs[f'keep_agnostic{i}'])} agnostic")
    arrs["n_cases"] = len(cases)
    arrs["n_class"] = np.array([c[1] for c in cases])
    save("detections", **arrs)


if __name__ == "__main__":
    only = sys.argv[1:]
    for fn in (gen_anchors, gen_boxmath, gen_proposals, gen_proposals_large, gen_nms, gen_anchor_targets,
               gen_proposal_targets, gen_roi, gen_roi_large, gen_roi_backward, gen_rpn_forward, gen_trainer,
               gen_detections):
        if only and fn.__name__ not in only:
            continue
        print(fn.__name__)
        fn()
