"""GPU: full backbone (stem + 3 stages + heads) through the C ABI against
 (a) the committed golden vectors produced by the reference module (tests/golden/backbone_64x96.npz), and
 (b) the fp32 PyTorch oracle (oracle/smap_torch.py, TF32 off) at the BASELINE resolution 832x512.
Tolerance (BASELINE.json north_star): max|a-b| / max|ref| <= 1e-3 per tensor, bf16x3 mode."""
import os

import numpy as np
import pytest
import torch

from oracle import smap_torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-3


def rel(a, b):
    return (a - b).abs().max().item() / b.abs().max().item()


@pytest.fixture(scope="module", autouse=True)
def no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("bn,seed", [("identity", 11), ("random", 12)])
def test_backbone_golden_64x96(bn, seed):
    from smap_b200.engine import Engine

    g = np.load(os.path.join(G, "backbone_64x96.npz"))
    eng = Engine(0, max_batch=2, in_h=64, in_w=96)
    eng.load_state_dict(smap_torch.make_state_dict(seed, bn))
    x = smap_torch.make_input(2, 64, 96, seed=seed + 100).cuda()
    hm, dd, rd = eng.forward(x)
    torch.cuda.synchronize()
    for name, t in (("hm2d", hm), ("detd", dd), ("rootd", rd)):
        ref = torch.from_numpy(g["%s_%s" % (bn, name)]).cuda()
        assert rel(t, ref) < TOL, (bn, name, rel(t, ref))
    # flip-TTA merge + rescale (exps/stage3_root2/test.py:55-70,111-112)
    hm_f, _, _ = eng.forward(torch.flip(x, [-1]))
    ref_raw = torch.from_numpy(g["%s_hm2d_flipraw" % bn]).cuda()
    assert rel(hm_f, ref_raw) < TOL
    merged = eng.merge_scale(hm.clone(), hm_f, do_scale=False)
    ref = torch.from_numpy(g["%s_hm2d_flipmerged" % bn]).cuda()
    assert rel(merged, ref) < TOL
    # the merge arithmetic itself is bit-exact given identical inputs
    a = torch.from_numpy(g["%s_hm2d" % bn]).cuda()
    m2 = eng.merge_scale(a.clone(), ref_raw, do_scale=True)
    exp = ref.clone()
    exp[:, :15] /= 255
    exp[:, 15:] /= 127
    assert torch.equal(m2, exp)
    eng.close()


@pytest.mark.parametrize("bn", ["identity", "random"])
def test_backbone_832x512_vs_fp32_oracle(bn):
    from smap_b200.engine import Engine

    sd = smap_torch.make_state_dict(0, bn)
    x = smap_torch.make_input(2, 512, 832, seed=1).cuda()
    ref = smap_torch.smap_forward({k: v.cuda() for k, v in sd.items()}, x)
    eng = Engine(0, max_batch=2, in_h=512, in_w=832)
    eng.load_state_dict(sd)
    out = eng.forward(x)
    torch.cuda.synchronize()
    for name, a, b in zip(("hm2d", "detd", "rootd"), out, ref):
        assert a.shape == b.shape
        assert torch.isfinite(a).all()
        assert rel(a, b) < TOL, (name, rel(a, b))
    # batch invariance, bit for bit: image 1 alone gives the same tensors (different batch = different tile boundaries
    # and possibly different tile shapes; every output element still accumulates the same products in the same order)
    o1 = eng.forward(x[1:2])
    torch.cuda.synchronize()
    for a, b in zip(o1, out):
        assert torch.equal(a[0], b[1])
    eng.close()


def test_backbone_1024x1024_config5():
    """BASELINE.json configs[4]: backbone-only at 1024x1024 (the association is defined at 128x208 maps only)."""
    from smap_b200.engine import Engine

    sd = smap_torch.make_state_dict(0, "identity")
    x = smap_torch.make_input(1, 1024, 1024, seed=2).cuda()
    ref = smap_torch.smap_forward({k: v.cuda() for k, v in sd.items()}, x)
    eng = Engine(0, max_batch=1, in_h=1024, in_w=1024)
    eng.load_state_dict(sd)
    out = eng.forward(x)
    torch.cuda.synchronize()
    for name, a, b in zip(("hm2d", "detd", "rootd"), out, ref):
        assert a.shape == b.shape == (1, {"hm2d": 43, "detd": 14, "rootd": 1}[name], 256, 256)
        assert rel(a, b) < TOL, (name, rel(a, b))
    n_conv, flops = eng.plan_info(1)
    assert abs(flops * 1e-9 - 740.0) / 740.0 < 0.06  # 740 GFLOP live graph (we commute the up_conv 1x1s)
    eng.close()
