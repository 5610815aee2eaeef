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
