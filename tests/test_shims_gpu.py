"""GPU: the reference-facing Python entry points (drop-in `dapalib` module and `model.smap.SMAP` class)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import assoc, smap_torch
from smap_b200 import schema
from smap_b200.synth import make_scene

pytestmark = pytest.mark.gpu
SHIMS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "smap_b200", "shims")


@pytest.fixture(scope="module")
def shims():
    sys.path.insert(0, SHIMS)
    for m in [m for m in sys.modules if m == "model" or m.startswith("model.") or m == "dapalib"]:
        del sys.modules[m]
    import dapalib
    from model.smap import SMAP

    yield dapalib, SMAP
    sys.path.remove(SHIMS)


def test_dapalib_connect_and_extract_signatures_and_values(shims):
    dapalib, _ = shims
    s = make_scene(77, persons=6)
    hms = torch.from_numpy(s["hms"]).cuda()
    rd = torch.from_numpy(s["root_d"])  # CPU tensor, as exps/stage3_root2/test.py:113 passes it
    out = dapalib.connect(hms, rd, 2, distFlag=True)
    assert out.device.type == "cpu" and out.dtype == torch.float32 and out.shape == (6, 15, 4)
    assert np.array_equal(out.numpy(), assoc.connect(s["hms"], s["root_d"]))
    pose, paf = dapalib.extract(hms)
    assert len(pose) == 15 and len(paf) == 14
    op, os_ = assoc.extract(s["hms"])
    for j in range(15):
        n = int(op[j, 0, 0])
        assert pose[j].shape == (n, 3) and np.array_equal(pose[j].numpy(), op[j, 1:n + 1])
    assert paf[0].shape == (pose[0].shape[0], pose[1].shape[0])
    assert np.array_equal(paf[3].numpy(), os_[3, :paf[3].shape[0], :paf[3].shape[1]])
    # no root peak -> 1-D empty tensor (association.cpp:133-136); the caller tests len(...) > 0
    empty = dapalib.connect(torch.zeros(43, 128, 208).cuda(), rd)
    assert empty.dim() == 1 and len(empty) == 0
    with pytest.raises(RuntimeError):
        dapalib.connect(torch.zeros(43, 64, 104).cuda(), rd)


def test_smap_module_drop_in(shims):
    _, SMAP = shims

    class NS(types.SimpleNamespace):
        pass

    cfg = NS(MODEL=NS(STAGE_NUM=3, UPSAMPLE_CHANNEL_NUM=256), DATASET=NS(KEYPOINT=NS(NUM=15), PAF=NS(NUM=14)),
             OUTPUT_SHAPE=(128, 208), LOSS=NS(OHKM=True, TOPK=8, COARSE_TO_FINE=True))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = SMAP(cfg, run_efficient=False)
    assert len(model.state_dict()) == 1876
    sd = schema.make_state_dict(4, "random")
    model.load_state_dict(sd)  # strict, reference schema
    model.to("cuda").eval()
    x = schema.make_input(1, 512, 832, seed=9).cuda()
    with torch.no_grad():
        o2d, o3d, ord_ = model(x)
    ref = smap_torch.smap_forward({k: v.cuda() for k, v in sd.items()}, x)
    for a, b in zip((o2d, o3d, ord_), ref):
        assert a.shape == b.shape and a.device == x.device and a.dtype == torch.float32
        assert (a - b).abs().max().item() / b.abs().max().item() < 1e-3
    # the caller mutates outputs_2d in place (exps/stage3_root2/test.py:111-112): ordinary writable tensors
    o2d[0, :15] /= 255
    # weight updates are picked up
    with torch.no_grad():
        model.top.conv.conv.bias.add_(0.5)
        o2 = model(x)[0]
    assert (o2 - ref[0]).abs().max().item() > 1e-3
    model.train()
    with pytest.raises(NotImplementedError):
        model(x)
