"""CPU: the drop-in modules expose exactly the reference's state-dict schema (keys, order, shapes) and random init.

When /root/reference is mounted (the build container) the comparison is made against the reference modules themselves;
everywhere else against the key lists the oracle restates (oracle/smap_torch.py, oracle/refine_torch.py), which
tests/test_oracle_golden.py pins to the reference.  No forward pass: the shims need a B200 for that."""
import os
import sys
import types

import pytest
import torch

from oracle import refine_torch
from smap_b200 import schema

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "smap_b200", "shims")
REF = "/root/reference"


def _cfg():
    NS = types.SimpleNamespace
    return NS(MODEL=NS(STAGE_NUM=3, UPSAMPLE_CHANNEL_NUM=256), DATASET=NS(KEYPOINT=NS(NUM=15), PAF=NS(NUM=14)),
              OUTPUT_SHAPE=(128, 208), LOSS=NS(OHKM=True, TOPK=8, COARSE_TO_FINE=True))


def _import_from(path, names):
    """import `model.smap` / `model.refinenet` with `path` first on sys.path, isolated from other `model` packages"""
    saved = {k: v for k, v in sys.modules.items() if k == "model" or k.startswith("model.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, path)
    try:
        mods = [__import__(n, fromlist=["x"]) for n in names]
    finally:
        sys.path.remove(path)
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return mods


def test_shim_schemas_match_the_restated_key_lists():
    smap_mod, refine_mod = _import_from(SHIMS, ["model.smap", "model.refinenet"])
    m = smap_mod.SMAP(_cfg())
    sd = m.state_dict()
    want = schema.make_state_dict(0, "identity")
    assert list(sd.keys()) == list(want.keys()) and len(sd) == 1876
    assert all(tuple(sd[k].shape) == tuple(want[k].shape) for k in sd)
    r = refine_mod.RefineNet()
    assert [(k, tuple(v.shape)) for k, v in r.state_dict().items()] == [(k, tuple(s)) for k, s in refine_torch.refine_keys()]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "model")), reason="reference not mounted")
def test_shims_match_the_reference_modules_key_for_key_and_init_for_init():
    ref_smap, ref_refine = _import_from(REF, ["model.smap", "model.refinenet"])
    shim_smap, shim_refine = _import_from(SHIMS, ["model.smap", "model.refinenet"])
    torch.manual_seed(0)
    a = ref_smap.SMAP(_cfg()).state_dict()
    torch.manual_seed(0)
    b = shim_smap.SMAP(_cfg()).state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
        assert torch.equal(a[k], b[k]), "random init differs at " + k   # same construction order -> same RNG stream
    torch.manual_seed(3)
    ra = ref_refine.RefineNet().state_dict()
    torch.manual_seed(3)
    rb = shim_refine.RefineNet().state_dict()
    assert list(ra.keys()) == list(rb.keys())
    for k in ra:
        assert ra[k].shape == rb[k].shape and torch.equal(ra[k], rb[k]), k
    # strict loading in both directions
    shim_refine.RefineNet().load_state_dict(ra)
    ref_refine.RefineNet().load_state_dict(rb)


def test_oracle_and_product_generators_agree_bit_for_bit():
    """oracle/schema_ref.py (the checker's own generators) and smap_b200/schema.py (the product's) are separate files on
    purpose; the synthetic weights and frames they make must be the same bytes."""
    from oracle import schema_ref

    assert schema_ref.unit_specs() == schema.unit_specs()
    for bn in ("identity", "random"):
        a, b = schema_ref.make_state_dict(3, bn), schema.make_state_dict(3, bn)
        assert list(a.keys()) == list(b.keys())
        assert all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(schema_ref.make_input(2, 64, 96, seed=5), schema.make_input(2, 64, 96, seed=5))
