"""CPU: the C-ABI library builds, loads and exports every symbol include/smap_b200.h declares.
No compute call is made (no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "smap_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smapb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from smap_b200 import _lib, build

    build.build()
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    assert set(_lib.EXPORTS) == set(syms)
    assert lib.smapb_version() >= 100


def test_debug_header_symbols_are_exported_too():
    from smap_b200 import _lib

    src = open(os.path.join(ROOT, "include", "smap_b200_debug.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    syms = sorted(set(re.findall(r"\b(smapb_debug_[a-z0-9_]+)\s*\(", src)))
    assert syms == ["smapb_debug_checksums", "smapb_debug_dump", "smapb_debug_resize_plan"]
    lib = _lib.load()
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s


def test_record_layout_matches_header():
    from smap_b200 import _lib, engine

    assert engine.RECORD_DTYPE.itemsize == _lib.RECORD_BYTES == 127 * 15 * 4 * 8 + 127 * 8 + 127 * 15 * 4 * 4 + 8
    assert engine.RECORD_DTYPE.fields["root_depth"][1] == 127 * 15 * 4 * 8
    assert engine.RECORD_DTYPE.fields["pred2d"][1] == 127 * 15 * 4 * 8 + 127 * 8


def test_engine_fails_loudly_without_gpu():
    import torch

    from smap_b200 import engine

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine.SmapB200Error):
        engine.Engine()


def test_create_reports_error_without_device():
    import ctypes

    import torch

    from smap_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.smapb_create(ctypes.byref(h), 0, 1, 512, 832)
    assert rc != 0 and not h.value
    assert len(lib.smapb_last_error(None)) > 0
