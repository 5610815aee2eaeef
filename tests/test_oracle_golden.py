"""CPU: the oracle restatements against golden vectors produced by the REFERENCE code itself
(tests/golden/make_golden.py imports /root/reference in the build container)."""
import os

import numpy as np
import torch

from cases import N_LIFT_CASES, lift_case_inputs
from oracle import lift_numpy, smap_torch

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_backbone_oracle_matches_reference_module():
    g = np.load(os.path.join(G, "backbone_64x96.npz"))
    for bn, seed in (("identity", 11), ("random", 12)):
        sd = smap_torch.make_state_dict(seed, bn)
        x = smap_torch.make_input(2, 64, 96, seed=seed + 100)
        o2d, dd, rd = smap_torch.smap_forward(sd, x)
        for name, t in (("hm2d", o2d), ("detd", dd), ("rootd", rd)):
            ref = g["%s_%s" % (bn, name)]
            # same ops in the same order on the same CPU: bit-identical here; allow 1e-6 relative for other hosts
            assert np.abs(t.numpy() - ref).max() <= 1e-6 * np.abs(ref).max(), (bn, name)
        f2d, _, _ = smap_torch.smap_forward(sd, torch.flip(x, [-1]))
        merged = smap_torch.flip_merge(o2d.clone(), f2d)
        ref = g["%s_hm2d_flipmerged" % bn]
        assert np.abs(merged.numpy() - ref).max() <= 1e-6 * np.abs(ref).max()


def test_state_dict_schema_size():
    sd = smap_torch.make_state_dict(0)
    assert len(sd) == 1876 and len(smap_torch.unit_specs()) == 268


def test_lift_oracle_matches_reference_functions():
    g = np.load(os.path.join(G, "lift_cases.npz"))
    for ci in range(N_LIFT_CASES):
        b, det_d, root_d, (iw, ih) = lift_case_inputs(ci)
        p2, p3, rdep = lift_numpy.lift(b, det_d, root_d, lift_numpy.default_scale(iw, ih))
        assert np.array_equal(p2, g["c%d_pred2d" % ci]), ci
        assert np.array_equal(rdep, g["c%d_rootdepth" % ci]), ci
        assert np.array_equal(p3, g["c%d_pred3d" % ci]), ci


def test_lift_oracle_gt_branch_matches_reference_functions():
    """register_pred with ground truth (test_util.py:21-39) + the float64 lift it implies."""
    from cases import N_GT_CASES, lift_gt_case_inputs

    g = np.load(os.path.join(G, "lift_gt_cases.npz"))
    for ci in range(N_GT_CASES):
        b, det_d, root_d, (iw, ih), gt = lift_gt_case_inputs(ci)
        sc = lift_numpy.default_scale(iw, ih)
        sc.update(f_x=gt[0, 0, 7], f_y=gt[0, 0, 8], cx=gt[0, 0, 9], cy=gt[0, 0, 10])  # test.py:91-95
        p2, p3, rdep = lift_numpy.lift(b, det_d, root_d, sc, gt_bodys=gt)
        assert p2.dtype == np.float64
        assert np.array_equal(p2, g["c%d_pred2d" % ci]), ci
        assert np.array_equal(rdep, g["c%d_rootdepth" % ci]), ci
        assert np.array_equal(p3, g["c%d_pred3d" % ci]), ci
