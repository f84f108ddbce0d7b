"""Pins the oracle (oracle/ref_port.py + oracle/frcnn_oracle.c) against fixtures produced by the real
reference + torchvision 0.26.0 CPU kernels (tests/golden/make_golden.py).  CPU only.

Bit-exact for every integer/index output and for all exp/log-free fp32 arithmetic; decode/encode (numpy
exp/log vs the reference's SLEEF) within ULP_BOUND units in the last place of the largest operand behind
each coordinate (measured here: decode 3, encode 2), which implies the 1e-5 relative bound of north_star."""

ULP_BOUND = 4
import numpy as np
import pytest

from conftest import box_close, decode_operands, load_golden, max_ulp_error
from divergence import first_divergence
from oracle import ref_port as O


def test_base_and_shifted_anchors_exact():
    g = load_golden("anchors_kat")
    base = O.base_anchors()
    assert np.array_equal(base, g["base"])
    base2 = O.base_anchors(base_size=16, ratios=[0.5, 1, 2, 3], anchor_scales=[4, 8])
    assert np.array_equal(base2, g["base2"])
    assert np.array_equal(O.shifted_anchors(base, 16, 5, 7), g["shifted_5x7"])
    assert np.array_equal(O.shifted_anchors(base, 16, 38, 38), g["shifted_38x38"])
    assert np.array_equal(O.shifted_anchors(base2, 8, 9, 4), g["shifted_rect"])


def test_reference_known_answers():
    """utils/loc_bbox_iou.py:99-103 -- the only known-answer material in the reference tree."""
    g = load_golden("anchors_kat")
    a = np.array([[100, 100, 200, 200]], np.float32)
    b = np.array([[150, 150, 250, 250]], np.float32)
    v = O.iou(a, b)
    assert np.array_equal(v, g["kat_iou"])
    assert abs(float(v[0, 0]) - 1.0 / 7.0) < 1e-7
    assert box_close(O.decode(a, O.encode(a, b)), b, 250.0)
    assert box_close(O.encode(a, b), g["kat_loc"], 1.0)


def test_boxmath():
    g = load_golden("boxmath")
    assert np.array_equal(O.iou(g["src"], g["gts"]), g["iou"])
    assert np.array_equal(O.iou(g["src"][:64], g["src_d"][:64]), g["iou_self"])
    assert box_close(O.decode(g["src"], g["loc"]), g["decode"], 600.0)
    assert box_close(O.decode(g["src"], g["loc8"]), g["decode8"], 600.0)
    assert box_close(O.encode(g["src"], g["dst"]), g["encode"], 1.0)
    assert box_close(O.encode(g["src_d"], g["dst_d"]), g["encode_d"], 1.0)
    # the same four, as a measured bound in units in the last place
    errs = {"decode": max_ulp_error(O.decode(g["src"], g["loc"]), g["decode"], decode_operands(g["src"], g["loc"])),
            "decode8": max_ulp_error(O.decode(g["src"], g["loc8"]), g["decode8"], decode_operands(g["src"], g["loc8"])),
            "encode": max_ulp_error(O.encode(g["src"], g["dst"]), g["encode"]),
            "encode_d": max_ulp_error(O.encode(g["src_d"], g["dst_d"]), g["encode_d"])}
    print("oracle vs reference, max ulp error:", errs)
    assert max(errs.values()) <= ULP_BOUND, errs
    with pytest.raises(IndexError):
        O.iou(np.zeros((3, 5), np.float32), np.zeros((2, 4), np.float32))
    assert O.decode(np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32)).shape == (0, 4)


def test_nms_exact():
    g = load_golden("nms")
    for i in range(int(g["n_cases"])):
        keep = O.nms(g[f"boxes{i}"], g[f"scores{i}"], float(g[f"thr{i}"]))
        assert np.array_equal(keep, g[f"keep{i}"].astype(np.int64)), i
    assert O.nms(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.5).shape == (0,)


PROPOSAL_CASES = ["proposal_small_train", "proposal_small_overlap", "proposal_small_ties",
                  "proposal_small_pad", "proposal_small_error", "proposal_small_scale",
                  "proposal_600_test", "proposal_600_train",
                  "proposal_800_test",      # BASELINE config 4: N = 22 500, 3 000 -> 300
                  "proposal_1024_stress"]   # BASELINE config 5: N = 36 864, 30 000 -> 2 000


@pytest.mark.parametrize("name", PROPOSAL_CASES)
def test_proposal_layer_stage_isolated(name):
    """Fed the reference's own decoded boxes, every later stage is bit-exact."""
    g = load_golden(name)
    img = tuple(int(v) for v in g["img_size"])
    scale, min_size = float(g["scale"]), float(g["min_size"])
    clipped, valid = O.clip_filter(g["decoded"], img, np.float32(min_size * scale))
    assert np.array_equal(valid, g["valid_idx"].astype(np.int64))
    order = O.argsort_desc_stable(g["score"][valid])
    n_pre = int(g["n_pre"])
    order = order[:n_pre] if n_pre > 0 else order
    assert np.array_equal(order, g["order"].astype(np.int64))
    sorted_boxes = clipped[valid][order]
    keep = O.nms(sorted_boxes, g["score"][valid][order], float(g["nms_iou"]))
    assert np.array_equal(keep, g["nms_keep"].astype(np.int64))
    kw = dict(scale=scale, nms_iou=float(g["nms_iou"]), n_pre_nms=n_pre, n_post_nms=int(g["n_post"]),
              min_size=min_size)
    if int(g["expect_error"]):
        with pytest.raises(IndexError):
            O.proposal_layer_from_boxes(g["decoded"], g["score"], img, **kw)
        return
    roi, src, nk = O.proposal_layer_from_boxes(g["decoded"], g["score"], img, return_extra=True, **kw)
    assert nk == keep.shape[0]
    assert np.array_equal(roi, g["roi"])
    assert np.array_equal(clipped[src], roi)


@pytest.mark.parametrize("name", ["proposal_small_train", "proposal_small_scale", "proposal_600_test",
                                  "proposal_600_train", "proposal_800_test", "proposal_1024_stress"])
def test_proposal_layer_end_to_end_tolerance(name):
    """From (loc, score): decode uses numpy exp, so boxes agree to a few ulp, and the output rows agree up to
    the first decision that sat on a threshold (tests/divergence.py proves that it did)."""
    g = load_golden(name)
    img = tuple(int(v) for v in g["img_size"])
    base = O.base_anchors()
    anchor = O.shifted_anchors(base, 16, int(g["H"]), int(g["W"]))
    dec = O.decode(anchor, g["loc"])
    assert box_close(dec, g["decoded"], float(max(img)))
    mode = str(g["mode"])
    lim = ({"n_train_pre_nms": int(g["n_pre"]), "n_train_post_nms": int(g["n_post"])} if mode == "train"
           else {"n_test_pre_nms": int(g["n_pre"]), "n_test_post_nms": int(g["n_post"])})
    roi = O.proposal_layer(g["loc"], g["score"], anchor, img, scale=float(g["scale"]), mode=mode,
                           nms_iou=float(g["nms_iou"]), min_size=float(g["min_size"]), **lim)
    assert roi.shape == g["roi"].shape
    assert max_ulp_error(dec, g["decoded"], decode_operands(anchor, g["loc"])) <= ULP_BOUND
    why = first_divergence(g["decoded"], dec, g["score"], img, float(g["min_size"]) * float(g["scale"]),
                           float(g["nms_iou"]), int(g["n_pre"]), int(g["n_post"]))
    print(name, why)
    rows = why["rows_equal"]
    assert box_close(roi[:rows], g["roi"][:rows], float(max(img)))


def test_anchor_targets():
    g = load_golden("anchor_targets")
    base = O.base_anchors()
    for name in [str(n) for n in g["names"]] + ["custom"]:
        anchor = O.shifted_anchors(base, 16, int(g[f"{name}_H"]), int(g[f"{name}_W"]))
        kw = {}
        if name == "custom":
            p = g["custom_params"]
            kw = dict(n_sample=int(p[0]), pos_iou_thresh=float(p[1]), neg_iou_thresh=float(p[2]),
                      pos_ratio=float(p[3]))
        loc, label = O.anchor_targets(g[f"{name}_bbox"], anchor, **kw)
        assert np.array_equal(label, g[f"{name}_label"].astype(np.int64)), name
        assert box_close(loc, g[f"{name}_loc"], 1.0), name


def test_proposal_targets():
    g = load_golden("proposal_targets")
    for name in [str(n) for n in g["names"]]:
        p = g[f"{name}_params"]
        kw = dict(n_sample=int(p[0]), pos_ratio=float(p[1]), pos_iou_thresh=float(p[2]),
                  neg_iou_thresh_high=float(p[3]), neg_iou_thresh_low=float(p[4]))
        args = (g[f"{name}_roi"], g[f"{name}_bbox"], g[f"{name}_label"])
        if int(g[f"{name}_error"]):
            with pytest.raises(IndexError):
                O.proposal_targets(*args, **kw)
            continue
        s, l, y = O.proposal_targets(*args, **kw)
        assert np.array_equal(s, g[f"{name}_sample_roi"]), name
        assert np.array_equal(y, g[f"{name}_gt_label"]), name
        assert box_close(l, g[f"{name}_gt_loc"], 1.0), name


def test_roi_pool_and_align_exact():
    g = load_golden("roi_ops")
    feat, rois = g["feat"], g["rois"]
    for key in g.files:
        if key.startswith("pool_P"):
            _, P, s = key.split("_")
            out = O.roi_pool(feat, rois, int(P[1:]), float(s[1:]))
            assert np.array_equal(out, g[key]), key
        elif key.startswith("align_P"):
            _, P, sr, al, s = key.split("_")
            out = O.roi_align(feat, rois, int(P[1:]), float(s[1:]), int(sr[2:]), bool(int(al[2:])))
            assert np.array_equal(out, g[key]), key
    out, am = O.roi_pool(feat, rois, 7, 1.0, return_argmax=True)
    assert np.array_equal(am, g["pool_argmax_P7_s1.0"])


def test_roi_ops_on_config_sized_maps_exact():
    """torchvision roi_pool 7x7 / 14x14 (+ argmax) and roi_align 7x7 on 38x38 / 50x50 / 64x64 maps."""
    g = load_golden("roi_large")
    for H in (38, 50, 64):
        feat, rois = g[f"feat{H}"], g[f"rois{H}"]
        for P in (7, 14):
            out, am = O.roi_pool(feat, rois, P, 1.0, return_argmax=True)
            assert np.array_equal(out, g[f"pool{H}_P{P}"]), (H, P)
            assert np.array_equal(am, g[f"argmax{H}_P{P}"]), (H, P)
        assert np.array_equal(O.roi_align(feat, rois, 7, 1.0, 2, False), g[f"align{H}_P7_sr2"]), H
        assert np.array_equal(O.roi_align(feat, rois, 7, 1.0, 2, True), g[f"align{H}_P7_sr2_al"]), H


def test_roi_backward_matches_torchvision_autograd():
    """Gradients w.r.t. the features: the oracle repeats the CPU kernels' sequential accumulation order, so
    it reproduces torch.autograd.grad through torchvision.ops.roi_pool / roi_align bit for bit."""
    g = load_golden("roi_backward")
    src = load_golden("roi_ops")
    feat, rois = src["feat"], src["rois"]
    for P in (7, 14):
        _, am = O.roi_pool(feat, rois, P, 1.0, return_argmax=True)
        gi = O.roi_pool_backward(g[f"pool_P{P}_go"], am, rois, feat.shape)
        assert np.array_equal(gi, g[f"pool_P{P}_gi"]), P
    for key in g.files:
        if key.startswith("align_") and key.endswith("_go"):
            _, sr, al, sc, _ = key.split("_")
            gi = O.roi_align_backward(g[key], rois, feat.shape, float(sc[1:]), int(sr[2:]), bool(int(al[2:])))
            want = g[key[:-3] + "_gi"]
            assert np.allclose(gi, want, rtol=0, atol=1e-6 * float(np.abs(want).max())), key
            assert np.array_equal(gi, want), key


def test_roi_head_coordinate_map_and_gather():
    g = load_golden("roi_head")
    x, rois = g["x"], g["rois"]
    for tag in ("chw", "hw"):
        img = tuple(int(v) for v in g[f"{tag}_img"])
        fm = O.roi_to_feature_coords(rois.reshape(-1, 4), img, x.shape[2], x.shape[3])
        assert np.array_equal(fm, g[f"{tag}_rois5"][:, 1:]), tag
        pool = O.roi_head_gather(x, rois, np.zeros(1), img, roi_size=7, spatial_scale=1.0)
        assert np.array_equal(pool[:16, :32], g[f"{tag}_pool_head"]), tag
        assert np.allclose(pool.astype(np.float64).sum((2, 3)), g[f"{tag}_pool_sum"], rtol=0, atol=1e-9)


def test_rpn_forward_glue():
    g = load_golden("rpn_forward")
    fg = O.fg_scores(g["rpn_scores"])
    assert np.allclose(fg, g["fg"], rtol=1e-5, atol=1e-7)
    H, W = g["x"].shape[2:]
    anchor = O.shifted_anchors(O.base_anchors(), 16, H, W)
    assert np.array_equal(anchor[None], g["anchor"])
    img = tuple(int(v) for v in g["img_size"])
    roi = O.proposal_layer(g["rpn_locs"][0], g["fg"][0], anchor, img, mode="test",
                           n_test_pre_nms=500, n_test_post_nms=40)
    assert box_close(roi, g["rois"][0], float(max(img)))


def test_detections_decode_and_per_class_nms():
    """post-head decode (frcnn_training.py:311-320), the evaluator's per-class NMS (:441-454) and
    multi_inference.py:84's class-agnostic NMS, against fixtures produced by the reference's own loc2bbox,
    torch.max and torchvision.ops.nms."""
    g = load_golden("detections")
    for i in range(int(g["n_cases"])):
        C = int(g["n_class"][i])
        boxes, sc, ci = O.detection_decode(g[f"roi{i}"], g[f"cls_loc{i}"], g[f"score{i}"], g[f"label{i}"])
        assert np.array_equal(ci, g[f"cls_index{i}"]) and np.array_equal(sc, g[f"cls_score{i}"])
        assert box_close(boxes, g[f"boxes{i}"], 800.0)
        boxes_p, _, _ = O.detection_decode(g[f"roi{i}"], g[f"cls_loc{i}"], g[f"score{i}"])
        assert box_close(boxes_p, g[f"boxes_pred_class{i}"], 800.0)
        for thr, tag in ((0.7, "keep"), (0.3, "keep03_")):
            kept = O.nms_by_class(g[f"boxes{i}"], g[f"cls_score{i}"], g[f"cls_index{i}"], thr)
            for c in range(C):
                want = g[f"{tag}{i}_c{c}"]
                assert np.array_equal(kept[g[f"cls_index{i}"][kept] == c], want), (i, c, thr)
        assert np.array_equal(O.nms_by_class(g[f"boxes{i}"], g[f"cls_score{i}"], None, 0.1), g[f"keep_agnostic{i}"])
