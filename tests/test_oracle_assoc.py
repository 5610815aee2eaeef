"""CPU: known-answer tests of the association oracle (oracle/assoc_oracle.cpp), covering every branch the
reference has (nmsBase.cu:24-49,84-133; bodyPartConnectorBase.cu:23-62; association.cpp:133-136,185,190,202-228).
The reference ships no vectors for this path; on the GPU box the oracle itself is pinned against the unmodified
reference extension (tests/test_assoc_gpu.py)."""
import numpy as np
import pytest

from oracle import assoc
from smap_b200.synth import make_scene

H, W = 128, 208


def blank():
    return np.zeros((43, H, W), np.float32)


def put_gauss(plane, x, y, amp=1.0, sigma=1.5):
    yy, xx = np.mgrid[0:H, 0:W]
    plane += (amp * np.exp(-((xx - x) ** 2 + (yy - y) ** 2) / (2 * sigma * sigma))).astype(np.float32)


def test_single_peak_position_and_score():
    hms = blank()
    put_gauss(hms[3], 50, 40)
    peaks, _ = assoc.extract(hms)
    assert peaks[3, 0, 0] == 1
    assert np.all(peaks[[c for c in range(15) if c != 3], 0, 0] == 0)
    np.testing.assert_allclose(peaks[3, 1], [50.5, 40.5, 1.0], atol=1e-4)


def test_plateau_border_threshold_produce_no_peak():
    hms = blank()
    hms[0, 10:12, 10:12] = 0.9  # 2x2 plateau: strict > fails everywhere
    hms[1, 0, 5] = 0.9  # border row
    hms[1, 5, W - 1] = 0.9  # border column
    hms[2, 20, 20] = 0.2  # == threshold, not >
    hms[4, 30, 30] = np.float32(0.2) + np.float32(1e-6)
    peaks, _ = assoc.extract(hms)
    assert peaks[0, 0, 0] == 0 and peaks[1, 0, 0] == 0 and peaks[2, 0, 0] == 0
    assert peaks[4, 0, 0] == 1


def test_raster_order_and_127_truncation():
    hms = blank()
    ys, xs = np.meshgrid(np.arange(2, 126, 4), np.arange(2, 206, 4), indexing="ij")  # 31 x 51 = 1581 peaks
    hms[5, ys, xs] = 0.5 + 0.001 * (xs % 7)
    peaks, _ = assoc.extract(hms)
    assert peaks[5, 0, 0] == 127
    got = peaks[5, 1:128, :2]
    exp = np.stack([xs.ravel()[:127] + 0.5, ys.ravel()[:127] + 0.5], 1)  # isolated pixels: centroid == pixel
    np.testing.assert_allclose(got, exp, atol=1e-5)


def test_refinement_window_clipped_at_border_and_ignores_negatives():
    hms = blank()
    hms[6, 1, 1] = 1.0
    hms[6, 1, 3] = 0.125  # below the NMS threshold, still inside the 7x7 centroid window
    hms[6, 3, 1] = -4.0  # negative scores are skipped by the centroid
    peaks, _ = assoc.extract(hms)
    assert peaks[6, 0, 0] == 1
    x = (1 * 1.0 + 3 * 0.125) / 1.125 + 0.5
    np.testing.assert_allclose(peaks[6, 1], [x, 1.5, 1.0], rtol=1e-6)


def test_paf_scores_straight_limb_and_far_pairs():
    hms = blank()
    # limb 0 = joints (0,1), PAF planes 15,16.  Two necks, two heads.
    put_gauss(hms[0], 40, 60)
    put_gauss(hms[0], 120, 60)
    put_gauss(hms[1], 40, 30)
    put_gauss(hms[1], 120, 30)
    hms[16, 28:62, 39:42] = -1.0  # unit vector (0,-1) along the first limb only
    peaks, scores = assoc.extract(hms)
    assert peaks[0, 0, 0] == 2 and peaks[1, 0, 0] == 2
    s = scores[0, :2, :2]
    assert abs(s[0, 0] - 1.0) < 1e-6  # full support
    assert s[1, 1] == -1.0  # no PAF under the second limb, far apart
    assert s[0, 1] == -1.0 and s[1, 0] == -1.0
    assert np.all(scores[0, 2:, :] == -1.0) and np.all(scores[0, :, 2:] == -1.0)


def test_paf_near_coincident_pairs_get_min_score():
    hms = blank()
    hms[0, 50, 50] = 1.0
    hms[1, 50, 51] = 1.0  # 1 px apart < sqrt(128*208)/150 = 1.0878, zero PAF
    hms[0, 90, 90] = 1.0
    hms[1, 90, 90] = 1.0  # coincident -> -1
    _, scores = assoc.extract(hms)
    assert scores[0, 0, 0] == np.float32(np.float32(0.1) + 1e-6)
    assert scores[0, 1, 1] == -1.0


def test_connect_empty_root_returns_zero_persons():
    hms = blank()
    put_gauss(hms[0], 40, 60)
    b = assoc.connect(hms, np.ones((H, W), np.float32))
    assert b.shape == (0, 15, 4)


def test_connect_depth_order_and_limb_assignment():
    s = make_scene(3, persons=4)
    b, peaks, _ = assoc.connect(s["hms"], s["root_d"], return_all=True)
    assert len(b) == 4
    # persons come out in ascending root depth
    order = np.argsort(s["depth"])
    np.testing.assert_allclose(b[:, 2, :2] - 0.5, s["joints"][order][:, 2], atol=1.5)  # overlapping blobs shift centroids
    assert np.all(b[:, :, 2] == 0)
    # a well separated scene recovers every joint of every person
    found = b[:, :, 3] > 0
    assert found.mean() > 0.8
    err = np.abs(b[:, :, :2] - 0.5 - s["joints"][order])[found]
    assert np.median(err) < 0.25


def test_connect_dist_flag_changes_only_scores_not_shapes():
    s = make_scene(5, persons=15)
    b1 = assoc.connect(s["hms"], s["root_d"], dist_flag=True)
    b0 = assoc.connect(s["hms"], s["root_d"], dist_flag=False)
    assert b1.shape == b0.shape == (15, 15, 4)


def test_connect_zero_source_score_is_skipped():
    # a person whose neck is missing cannot grow arms/head: association.cpp:190
    s = make_scene(9, persons=1)
    hms = s["hms"].copy()
    hms[0] = 0  # remove every neck peak
    b = assoc.connect(hms, s["root_d"])
    assert len(b) == 1
    assert b[0, 0, 3] == 0 and b[0, 1, 3] == 0 and np.all(b[0, [3, 4, 5, 9, 10, 11], 3] == 0)
    assert b[0, 2, 3] > 0 and b[0, 6, 3] > 0 and b[0, 12, 3] > 0  # pelvis + hips still found


@pytest.mark.parametrize("seed", range(4))
def test_connect_deterministic(seed):
    s = make_scene(seed, persons=15)
    a = assoc.connect(s["hms"], s["root_d"])
    b = assoc.connect(s["hms"], s["root_d"])
    assert np.array_equal(a, b)


def test_depth_sort_matches_torch_unstable_sort():
    """association.cpp:144 is an UNSTABLE at::sort; the oracle replays it with std::sort (same algorithm)."""
    import torch

    rng = np.random.default_rng(1)
    for t in range(300):
        n = int(rng.integers(1, 128))
        k = rng.integers(0, max(2, n // 3), n).astype(np.float32)  # many ties
        if t % 7 == 0:
            k[rng.integers(0, n)] = np.nan
        _, idx = torch.from_numpy(k.copy()).sort(0, False)
        assert np.array_equal(assoc.depth_order(k), idx.numpy())
