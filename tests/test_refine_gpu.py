"""GPU: RefineNet post-processing through the C ABI against the oracle and the reference goldens."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from cases import N_LIFT_CASES, refine_state_dict  # noqa: E402

from oracle import lift_numpy, refine_torch  # noqa: E402
from smap_b200 import schema  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-3  # north_star tolerance for floating-point results, relative to the tensor's max magnitude


@pytest.fixture(scope="module")
def eng():
    from smap_b200.engine import Engine

    e = Engine(0, max_batch=3, in_h=512, in_w=832)
    e.load_refine_state_dict(refine_state_dict())
    yield e
    e.close()


def _sd_torch():
    return {k: torch.from_numpy(np.asarray(v)) for k, v in refine_state_dict().items()}


def test_mlp_matches_oracle(eng):
    g = torch.Generator().manual_seed(3)
    for n in (1, 3, 4, 5, 127, 300):
        x = (torch.randn(n, 75, generator=g) * 50).float()
        want = refine_torch.mlp(_sd_torch(), x).numpy()
        got = eng.refine_mlp(x.cuda()).cpu().numpy()
        assert got.shape == (n, 45)
        assert np.abs(got - want).max() <= RTOL * np.abs(want).max()
    assert eng.refine_mlp(torch.zeros(0, 75, device="cuda")).shape == (0, 45)


def test_refine_matches_reference_golden(eng):
    """inputs: the reference lift goldens; expected: unmodified lift_and_refine_3d_pose + RefineNet (tests/golden)."""
    lift = np.load(os.path.join(GOLD, "lift_cases.npz"))
    gold = np.load(os.path.join(GOLD, "refine_cases.npz"))
    B = N_LIFT_CASES
    p2 = torch.zeros(B, 127, 15, 4)
    p3 = torch.zeros(B, 127, 15, 4, dtype=torch.float64)
    cnt = torch.zeros(B, dtype=torch.int32)
    for ci in range(B):
        a2, a3 = lift["c%d_pred2d" % ci], lift["c%d_pred3d" % ci]
        n = len(a3)
        cnt[ci] = n
        if n:
            p2[ci, :n] = torch.from_numpy(a2)
            p3[ci, :n] = torch.from_numpy(a3)
    out = eng.refine(p2.cuda(), p3.cuda(), cnt.cuda()).cpu().numpy()
    for ci in range(B):
        want = gold["c%d_refined" % ci]
        n = len(want)
        got = out[ci, :n]
        assert (out[ci, n:] == 0).all()
        if n == 0:
            continue
        assert np.array_equal(got[:, :, 3], want[:, :, 3])          # score column exact
        assert np.array_equal(got[:, 2, :3], want[:, 2, :3])        # root row = the lifted root, exact
        assert np.abs(got[:, :, :3] - want[:, :, :3]).max() <= RTOL * np.abs(want[:, :, :3]).max()
        # values are float32-representable, as in the reference (float32 array concatenated into float64)
        assert np.array_equal(got[:, :, :3], got[:, :, :3].astype(np.float32).astype(np.float64))


def test_infer_device_with_refine(eng):
    """whole path with set_refine(True): pred3d in the records = oracle refine of the un-refined records."""
    from smap_b200.engine import records_to_numpy, scale_row

    eng.load_state_dict(schema.make_state_dict(0, "identity"))
    x = schema.make_input(2, 512, 832, seed=21).cuda()
    sc = lift_numpy.default_scale(1920, 1080)
    scales = torch.from_numpy(np.stack([scale_row(sc)] * 2)).cuda()
    eng.set_refine(False)
    plain = [records_to_numpy(eng.infer_device(x, scales)) for _ in range(3)][-1]   # 3 calls: the third replays a graph
    eng.set_refine(True)
    refined = [records_to_numpy(eng.infer_device(x, scales)) for _ in range(3)][-1]
    eng.set_refine(False)
    again = records_to_numpy(eng.infer_device(x, scales))
    assert again["pred3d"].tobytes() == plain["pred3d"].tobytes()
    sd = _sd_torch()
    for i in range(2):
        n = int(plain["count"][i])
        assert refined["count"][i] == n
        assert np.array_equal(refined["pred2d"][i], plain["pred2d"][i])
        assert np.array_equal(refined["root_depth"][i], plain["root_depth"][i])
        if n == 0:
            continue
        want = refine_torch.refine(plain["pred2d"][i, :n], plain["pred3d"][i, :n], sd)
        got = refined["pred3d"][i, :n]
        assert np.array_equal(got[:, :, 3], want[:, :, 3])
        assert np.abs(got[:, :, :3] - want[:, :, :3]).max() <= RTOL * max(1e-6, np.abs(want[:, :, :3]).max())


def test_refinenet_shim_is_a_drop_in(eng):
    shim_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "smap_b200", "shims")
    sys.path.insert(0, shim_dir)
    try:
        for m in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[m]
        from model.refinenet import RefineNet

        net = RefineNet()
        assert [k for k, _ in refine_torch.refine_keys()] == list(net.state_dict().keys())
        net.load_state_dict(_sd_torch())
        net.to("cuda").eval()
        x = (torch.randn(9, 75, generator=torch.Generator().manual_seed(5)) * 30).float()
        with torch.no_grad():
            got = net(x.cuda())
        want = refine_torch.mlp(_sd_torch(), x)
        assert got.is_cuda and got.dtype == torch.float32 and got.shape == (9, 45)
        assert (got.cpu() - want).abs().max() <= RTOL * want.abs().max()
    finally:
        sys.path.remove(shim_dir)
