"""CPU: the pre-processing oracle against (1) digests of the unmodified reference pipeline's outputs
(tests/golden/preprocess_digests.json, written by tests/golden/make_golden.py from dataset/custom_dataset.py +
torchvision) and (2) cv2.resize itself when opencv is importable (it is a third-party dependency of the reference)."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from cases import PRE_GEOMS, preprocess_case_image  # noqa: E402

from oracle import preprocess_numpy as P  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("ci", range(len(PRE_GEOMS)))
def test_oracle_reproduces_reference_pipeline_digest(ci):
    g = json.load(open(os.path.join(GOLD, "preprocess_digests.json")))["c%d" % ci]
    img = preprocess_case_image(ci)
    assert [img.shape[1], img.shape[0]] == g["geom"]
    u8, _ = P.aug_croppad(img)
    assert hashlib.sha256(np.ascontiguousarray(u8).tobytes()).hexdigest() == g["u8_sha256"]
    t, sc = P.preprocess(img)
    assert t.shape == (3, 512, 832) and t.dtype == np.float32
    assert hashlib.sha256(t.tobytes()).hexdigest() == g["sha256"]
    for k, v in g["scale"].items():
        assert float(sc[k]) == v


def test_value_level_fixture():
    f = np.load(os.path.join(GOLD, "preprocess_small.npz"))
    t, _ = P.preprocess(preprocess_case_image(int(f["ci"])))
    assert np.array_equal(t[:, ::4, ::4], f["tensor"])


def test_resize_against_cv2_when_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for (W, H) in [(1920, 1080), (700, 933), (1664, 1024), (64, 40), (832, 512), (1234, 777)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        s = min(832 / W, 512 / H)
        assert np.array_equal(cv2.resize(img, (0, 0), fx=s, fy=s), P.resize_linear_u8(img, s))


def test_library_resize_plan_equals_oracle_tables_over_many_geometries():
    """The host-side table builder inside libsmap_b200.so (make_resize_plan) against the oracle's tables - no GPU needed.
    Sweeps 400 source geometries incl. up-scaling, exact 1/2 and 1/1 scales and extreme aspect ratios."""
    import ctypes

    from smap_b200 import _lib

    lib = _lib.load()
    fn = lib.smapb_debug_resize_plan
    fn.argtypes = [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double), ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    fn.restype = ctypes.c_int
    rng = np.random.default_rng(77)
    geoms = [(1920, 1080), (1664, 1024), (832, 512), (416, 256), (3328, 2048), (2, 2), (16384, 16384), (5000, 40), (40, 5000)]
    geoms += [(int(rng.integers(8, 4200)), int(rng.integers(8, 3200))) for _ in range(391)]
    for (W, H) in geoms:
        dims = (ctypes.c_int * 6)()
        sc = ctypes.c_double()
        xo = np.zeros(832, np.int32); xc = np.zeros(832 * 2, np.int16); yo = np.zeros(512 * 2, np.int32); yc = np.zeros(512 * 2, np.int16)
        assert fn(W, H, 832, 512, dims, ctypes.byref(sc), xo.ctypes.data, xc.ctypes.data, yo.ctypes.data, yc.ctypes.data) == 0
        s = min(832 / W, 512 / H)
        assert sc.value == s
        dw, dh = P.cv_round(W * s), P.cv_round(H * s)
        assert (dims[0], dims[1]) == (dw, dh), (W, H)
        inv = 1.0 / s
        mode = 2 if (dw, dh) == (W, H) else (1 if int(inv) == 2 and abs(2 - inv) < np.finfo(np.float64).eps else 0)
        assert dims[4] == mode, (W, H)
        pad_l = (832 - dw) // 2 if dw < 832 else 0
        pad_t = (512 - dh) // 2 if (dw >= 832 and dh < 512) else 0
        assert (dims[2], dims[3]) == (pad_l, pad_t), (W, H)
        if mode != 0:
            continue
        oxo, oxa = P.linear_tables(W, dw, s)
        assert np.array_equal(xo[:dw], oxo) and np.array_equal(xc[:2 * dw].reshape(dw, 2), oxa), (W, H)
        # vertical taps: clamped rows, unsnapped weights (oracle/preprocess_numpy.py resize_linear_u8)
        for d in range(dh):
            f = np.float32((d + 0.5) * inv - 0.5)
            sy = int(np.floor(f))
            f = np.float32(f - np.float32(sy))
            assert yo[2 * d] == min(max(sy, 0), H - 1) and yo[2 * d + 1] == min(max(sy + 1, 0), H - 1), (W, H, d)
            assert yc[2 * d] == P.cv_round(float(np.float32((np.float32(1.0) - f) * np.float32(2048)))), (W, H, d)
            assert yc[2 * d + 1] == P.cv_round(float(np.float32(f * np.float32(2048)))), (W, H, d)
