import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a box without a CUDA device."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device: the hot path has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def box_close(a, b, extent, rtol=1e-5):
    """fp32 box tolerance used everywhere: |a-b| <= 1e-5 * max(|b|, extent) (1e-5 relative to the
    box/image scale; coordinates are differences of ~extent-sized terms)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False
    tol = rtol * np.maximum(np.abs(b), extent)
    both_nan = np.isnan(a) & np.isnan(b)
    same_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    with np.errstate(invalid="ignore"):
        ok = (np.abs(a - b) <= tol) | both_nan | same_inf
    return bool(ok.all())


def max_ulp_error(a, b, operand=None):
    """Largest |a-b| in units of the fp32 spacing of the LARGEST OPERAND behind each element: `operand` (same
    shape or broadcastable; magnitudes of the terms the element was summed from) or, when absent, |b| itself.
    NaN == NaN and equal infinities count as 0; a NaN / inf on one side only gives inf."""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = np.abs(b) if operand is None else np.maximum(np.abs(b), np.abs(np.asarray(operand, dtype=np.float32)))
    with np.errstate(invalid="ignore", over="ignore"):
        ulp = np.spacing(np.where(np.isfinite(scale), scale, np.float32(1)).astype(np.float32)).astype(np.float64)
        err = np.abs(a.astype(np.float64) - b.astype(np.float64)) / ulp
    same = (np.isnan(a) & np.isnan(b)) | (np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b)))
    err = np.where(same, 0.0, err)
    err = np.where(np.isnan(err), np.inf, err)
    return float(err.max()) if err.size else 0.0


def decode_operands(src, loc):
    """Magnitude of the largest term each decoded coordinate is built from (utils/loc_bbox_iou.py:36-59):
    max(|centre|, |half size|) per box and axis, tiled to (x1, y1, x2, y2) per group of four."""
    src = np.asarray(src, dtype=np.float64)
    loc = np.asarray(loc, dtype=np.float64)
    w, h = src[:, 2] - src[:, 0], src[:, 3] - src[:, 1]
    cx, cy = src[:, 0] + 0.5 * w, src[:, 1] + 0.5 * h
    out = np.empty_like(loc)
    with np.errstate(over="ignore", invalid="ignore"):
        for k in range(loc.shape[1] // 4):
            dx, dy, dw, dh = (loc[:, 4 * k + i] for i in range(4))
            ox = np.maximum(np.maximum(np.abs(dx * w), np.abs(cx)), 0.5 * np.exp(dw) * w)
            oy = np.maximum(np.maximum(np.abs(dy * h), np.abs(cy)), 0.5 * np.exp(dh) * h)
            out[:, 4 * k + 0] = out[:, 4 * k + 2] = ox
            out[:, 4 * k + 1] = out[:, 4 * k + 3] = oy
    return out
