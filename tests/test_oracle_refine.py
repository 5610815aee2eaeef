"""CPU: the RefineNet oracle against the goldens produced by the unmodified reference (tests/golden/make_golden.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from cases import N_LIFT_CASES, refine_state_dict  # noqa: E402

from oracle import refine_torch  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_refine_oracle_reproduces_reference_goldens_bit_for_bit():
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in refine_state_dict().items()}
    lift = np.load(os.path.join(GOLD, "lift_cases.npz"))
    gold = np.load(os.path.join(GOLD, "refine_cases.npz"))
    nonempty = 0
    for ci in range(N_LIFT_CASES):
        got = refine_torch.refine(lift["c%d_pred2d" % ci], lift["c%d_pred3d" % ci], sd)
        want = gold["c%d_refined" % ci]
        assert got.shape == want.shape and got.dtype == np.float64
        assert np.array_equal(got, want)
        nonempty += len(want) > 0
    assert nonempty >= 15


def test_state_dict_schema():
    keys = refine_torch.refine_keys()
    sd = refine_state_dict()
    assert [k for k, _ in keys] == list(sd.keys())
    for k, shp in keys:
        assert tuple(np.asarray(sd[k]).shape) == tuple(shp)
