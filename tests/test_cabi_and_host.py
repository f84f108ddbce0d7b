"""CPU-only checks: the C-ABI library loads and exports every symbol include/frcnn_b200.h declares,
the ctypes signature table covers the header, host-side argument logic, loud failure without CUDA,
and the multi-process sharding / all-gather plumbing over gloo (world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "frcnn_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(frcnn_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from two_stage_object_detection_b200 import build, _lib
    path = build.build()
    lib = ctypes.CDLL(path)
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    loaded = _lib.load()
    assert loaded.frcnn_abi_version() == _lib.ABI_VERSION
    assert loaded.frcnn_last_error() is not None


def test_ctypes_structs_match_header_layout():
    from two_stage_object_detection_b200 import _lib
    # sizes implied by the C declarations (natural alignment, x86-64)
    assert ctypes.sizeof(_lib.AnchorSpec) == 32
    assert ctypes.sizeof(_lib.ProposalParams) == 56
    assert _lib.ProposalParams.nms_thresh.offset == 32
    assert ctypes.sizeof(_lib.AnchorTargetParams) == 28
    assert ctypes.sizeof(_lib.ProposalTargetParams) == 32


def test_host_side_argument_checks_need_no_gpu():
    """Calls that fail validation return before any CUDA work."""
    from two_stage_object_detection_b200 import _lib
    lib = _lib.load()
    p = _lib.ProposalParams()
    assert lib.frcnn_proposals_workspace_bytes(ctypes.byref(p)) == 0
    p.batch, p.num_anchors, p.n_pre_nms, p.n_post_nms = 16, 12996, 3000, 300
    n = lib.frcnn_proposals_workspace_bytes(ctypes.byref(p))
    assert n > 16 * 12996 * 20 and n % 256 == 0
    assert lib.frcnn_topk_workspace_bytes(2, 1000) >= 2 * 1000 * 16
    rc = lib.frcnn_bbox_iou(None, None, -1, 4, None, None)
    assert rc == -1 and b"bad shape" in lib.frcnn_last_error()
    rc = lib.frcnn_proposals(ctypes.byref(p), None, None, None, None, None, None, None, None, 0, None)
    assert rc == -1
    rc = lib.frcnn_roi_pool_forward(None, 0, 1, 1, 1, None, 0, 0, 7, 7, 1.0, None, None, None, 0, None)
    assert rc == -1
    # the entry points added for SURVEY 8f-3 / 8f-4 validate the same way
    assert lib.frcnn_roi_pool_mean_forward(None, 0, 1, 1, 1, None, 0, 0, 7, 7, 1.0, None, None, 0, None) == -1
    assert lib.frcnn_roi_align_mean_forward(None, 0, 1, 1, 1, None, 0, 0, 7, 7, 1.0, 2, 0, None, None, 0, None) == -1
    assert lib.frcnn_roi_align_mean_workspace_bytes(4, 1200) >= 1200 * 528 and \
        lib.frcnn_roi_align_mean_workspace_bytes(4, 1200) % 256 == 0
    assert lib.frcnn_detection_decode(None, None, None, None, -1, 3, None, None, None, None, None) == -1
    assert lib.frcnn_detection_decode(None, None, None, None, 0, 3, None, None, None, None, None) == 0  # nothing to do
    assert lib.frcnn_nms_by_class(None, None, None, None, 1, 5, 0.5, None, None, None) == -1
    assert lib.frcnn_nms_by_class(None, None, None, None, 0, 5, 0.5, None, None, None) == 0


def test_no_cpu_fallback():
    from two_stage_object_detection_b200 import functional as F, _lib
    from two_stage_object_detection_b200.nets import AnchorTargetCreator, ProposalCreator
    with pytest.raises(_lib.FrcnnError):
        F.bbox_iou(torch.zeros(2, 4), torch.zeros(3, 4))
    with pytest.raises(_lib.FrcnnError):
        F.loc2bbox(torch.zeros(2, 4), torch.zeros(2, 4))
    with pytest.raises(_lib.FrcnnError):
        ProposalCreator("test")(torch.zeros(9, 4), torch.zeros(9), torch.zeros(9, 4), (3, 16, 16))
    with pytest.raises(_lib.FrcnnError):
        AnchorTargetCreator()(torch.zeros(1, 4), torch.zeros(9, 4))
    with pytest.raises(_lib.FrcnnError):
        F.roi_pool(torch.zeros(1, 1, 4, 4), torch.zeros(1, 5), 2)


def test_product_package_never_imports_the_oracle():
    """The product path may not route through the oracle, torchvision ops or any CPU fallback."""
    pkg = os.path.join(ROOT, "two_stage_object_detection_b200")
    bad = re.compile(r"^\s*(import|from)\s+(oracle|torchvision|triton)\b", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not bad.search(text), f
                assert "ref_port" not in text, f


def test_dropin_signatures_match_reference_table():
    """names, argument order and defaults."""
    import inspect
    from two_stage_object_detection_b200 import nets, utils

    def params(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values() if p.name != "self"]

    E = inspect.Parameter.empty
    assert params(utils.generate_basic_anchor) == [("base_size", 8), ("ratios", [0.5, 1, 2]), ("anchor_scales", [8, 16, 32])]
    assert params(utils.enumerate_shifted_anchor) == [("anchor_base", E), ("feat_stride", E), ("height", E), ("width", E)]
    assert [n for n, _ in params(utils.bbox_iou)] == ["bbox_a", "bbox_b"]
    assert [n for n, _ in params(utils.loc2bbox)] == ["src_bbox", "loc"]
    assert [n for n, _ in params(utils.bbox2loc)] == ["src_bbox", "dst_bbox"]
    assert params(nets.ProposalCreator.__init__) == [("mode", E), ("nms_iou", 0.7), ("n_train_pre_nms", 12000),
                                                     ("n_train_post_nms", 600), ("n_test_pre_nms", 3000),
                                                     ("n_test_post_nms", 300), ("min_size", 16)]
    assert params(nets.ProposalCreator.__call__) == [("loc", E), ("score", E), ("anchor", E), ("img_size", E), ("scale", 1.)]
    assert params(nets.RegionProposalNetwork.__init__) == [("in_channels", 512), ("ratios", [0.5, 1, 2]),
                                                           ("anchor_scales", [8, 16, 32]), ("feat_stride", 16),
                                                           ("mode", "training")]
    assert params(nets.RegionProposalNetwork.forward) == [("x", E), ("img_size", E), ("scale", 1.)]
    assert params(nets.AnchorTargetCreator.__init__) == [("n_sample", 256), ("pos_iou_thresh", 0.7),
                                                         ("neg_iou_thresh", 0.3), ("pos_ratio", 0.5)]
    assert params(nets.AnchorTargetCreator.__call__) == [("bbox", E), ("anchor", E)]
    assert params(nets.ProposalTargetCreator.__init__) == [("n_sample", 128), ("pos_ratio", 0.5), ("pos_iou_thresh", 0.5),
                                                           ("neg_iou_thresh_high", 0.5), ("neg_iou_thresh_low", 0)]
    assert params(nets.ProposalTargetCreator.__call__) == [("roi", E), ("bbox", E), ("label", E),
                                                           ("loc_normalize_std", (0.1, 0.1, 0.2, 0.2))]
    assert params(nets.HarNetRoIHead.__init__)[:4] == [("n_class", E), ("roi_size", E), ("spatial_scale", E), ("classifier", E)]
    assert params(nets.HarNetRoIHead.forward) == [("x", E), ("rois", E), ("roi_indices", E), ("img_size", E)]
    assert params(nets.FasterRCNN.__init__)[:5] == [("num_classes", E), ("mode", "training"), ("feat_stride", 16),
                                                    ("anchor_scales", [8, 16, 32]), ("ratios", [0.5, 1, 2])]
    assert params(nets.FasterRCNN.forward) == [("x", E), ("scale", 1.), ("mode", "forward")]
    pc = nets.ProposalCreator("train")
    assert pc.limits() == (12000, 600) and nets.ProposalCreator("training").limits() == (3000, 300)


def test_shard_bounds():
    from two_stage_object_detection_b200.distributed import shard_bounds
    for n in (0, 1, 7, 16, 64):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from two_stage_object_detection_b200.distributed import shard_bounds, all_gather_detections
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
B, n_post = 5, 7            # uneven split: 3 + 2
full = torch.arange(B * n_post * 4, dtype=torch.float32).view(B, n_post, 4)
keep = torch.arange(B, dtype=torch.int32) + 10
lo, hi = shard_bounds(B, rank, 2)
rois, n_keep = all_gather_detections(full[lo:hi].clone(), keep[lo:hi].clone())
assert torch.equal(rois, full), rank
assert torch.equal(n_keep, keep), rank
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


@pytest.mark.timeout(120)
def test_all_gather_detections_gloo_world2(tmp_path):
    """Sharded result == single-process result, bit-identical, over a real 2-process group."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=100)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"ok {r}" in o


def test_divergence_proof_accepts_threshold_flips_only():
    """tests/divergence.py (the 'why may rows differ' proof used by the drop-in tests): a decision that flips
    because it sat on its threshold is reported, a flip with a real margin is an assertion failure."""
    from divergence import first_divergence
    a = np.array([[0, 0, 100, 100]], np.float32)
    # IoU of [0,0,100,100] with [0,0,100,h]: h/100 for h < 100 -> 0.7 at h = 70
    ref = np.concatenate([a, [[0, 0, 100, 70.0001]], [[200, 200, 300, 300]]]).astype(np.float32)
    oth = np.concatenate([a, [[0, 0, 100, 69.9999]], [[200, 200, 300, 300]]]).astype(np.float32)
    score = np.array([0.9, 0.8, 0.7], np.float32)
    kw = dict(img_size=(3, 400, 400), min_size=16.0, nms_iou=0.7, n_pre=10, n_post=3)
    same = first_divergence(ref, ref, score, **kw)
    assert same["kind"] is None and same["rows_equal"] == 3
    flip = first_divergence(ref, oth, score, **kw)
    assert flip["kind"] == "nms" and flip["rows_equal"] == 1
    far = oth.copy()
    far[1, 3] = 60.0  # IoU 0.6 against the reference's 0.700001: this side is nowhere near the threshold
    with pytest.raises(AssertionError):
        first_divergence(ref, far, score, **kw)
    ref_clear = ref.copy()
    ref_clear[1, 3] = 90.0  # IoU 0.9 in the reference, 0.6 in the other: not a threshold effect
    with pytest.raises(AssertionError):
        first_divergence(ref_clear, far, score, **kw)
    small_r = np.array([[0, 0, 16.00001, 50]], np.float32)
    small_o = np.array([[0, 0, 15.99999, 50]], np.float32)
    ms = first_divergence(np.concatenate([small_r, a]), np.concatenate([small_o, a]), np.array([0.9, 0.8], np.float32), **kw)
    assert ms["kind"] == "min_size" and ms["rows_equal"] == 0
    with pytest.raises(AssertionError):
        first_divergence(np.concatenate([small_r, a]), np.concatenate([[[0, 0, 12, 50]], a]).astype(np.float32),
                         np.array([0.9, 0.8], np.float32), **kw)


def test_lookup_list_cover_is_exact():
    """The rule roi_pool_desc_kernel uses to cover a RoIPool bin with square windows (csrc/roi_ops.cu, pd_anchor):
    windows of side s = min(smax, height, width) anchored at lo, lo + s, ... and hi - s for the remainder.  The union of
    the windows must be exactly the bin -- max over a superset would read pixels torchvision's bin does not contain, a
    subset would miss some -- with ceil(L / s) windows per axis, so every bin up to 8 x 8 fits a 16-entry list with the
    2 x 2 table alone and a 3 x 3 bin is one lookup with the 3 x 3 table."""
    def anchors(lo, hi, s):
        full = (hi - lo) // s
        return [lo + k * s if k < full else hi - s for k in range(-(-(hi - lo) // s))]

    for smax in (2, 3):
        for hh in range(1, 20):
            for ww in range(1, 20):
                s = min(smax, hh, ww)
                ry, cx = anchors(5, 5 + hh, s), anchors(9, 9 + ww, s)
                assert len(ry) == -(-hh // s) and len(cx) == -(-ww // s)
                covered = {(y + dy, x + dx) for y in ry for x in cx for dy in range(s) for dx in range(s)}
                assert covered == {(y, x) for y in range(5, 5 + hh) for x in range(9, 9 + ww)}, (smax, hh, ww)
                if hh <= 8 and ww <= 8 and min(hh, ww) >= 2:
                    assert len(ry) * len(cx) <= 16
    assert anchors(0, 3, 3) == [0] and anchors(0, 5, 3) == [0, 2] and anchors(4, 11, 2) == [4, 6, 8, 9]
