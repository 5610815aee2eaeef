"""CPU: smap_b200/mat_export.py against the unmodified reference converter (lib/eval/convert.py) - the committed fixture
tests/golden/mat_cases.npz holds a synthetic result JSON and the bytes of the .mat files the reference wrote for it."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden():
    g = np.load(os.path.join(HERE, "golden", "mat_cases.npz"))
    return bytes(g["json"]).decode(), bytes(g["pose3d"]), bytes(g["pose2d"])


def test_mat_files_are_byte_identical_behind_the_timestamped_header(golden, tmp_path):
    from smap_b200 import mat_export

    text, m3, m2 = golden
    p = tmp_path / "in.json"
    p.write_text(text)
    mat_export.convert(str(p), str(tmp_path))
    a3 = (tmp_path / "pose3d.mat").read_bytes()
    a2 = (tmp_path / "pose2d.mat").read_bytes()
    assert a3[:19] == b"MATLAB 5.0 MAT-file" and a3[124:128] == b"\x00\x01IM"
    assert a3[128:] == m3
    assert a2[128:] == m2


def test_save_result_spelling_is_accepted_too(golden, tmp_path):
    """exps/stage3_root2/test_util.py:146-158 writes 'pred_3d' / 'gt_3d'; the reference converter reads 'pred' / 'gt'."""
    from smap_b200 import mat_export

    text, m3, _ = golden
    d = json.loads(text)
    for e in d["3d_pairs"]:
        e["pred_3d"] = e.pop("pred")
        e["gt_3d"] = e.pop("gt")
    p = tmp_path / "in2.json"
    p.write_text(json.dumps(d))
    mat_export.convert(str(p), str(tmp_path))
    assert (tmp_path / "pose3d.mat").read_bytes()[128:] == m3


def test_sequence_geometry_and_unletterbox():
    from smap_b200 import mat_export

    assert mat_export.sequence_geometry("a/b/TS3/img_1.jpg") == ("TS3/img_1.jpg", 3, 2048, 2048)
    assert mat_export.sequence_geometry("TS20/x.jpg")[2:] == (1920, 1080)
    with pytest.raises(NotImplementedError):
        mat_export.sequence_geometry("TS21/x.jpg")
    p2 = np.zeros((1, 15, 4))
    p2[0, 0] = (416.0, 256.0, 0.0, 1.0)
    out = mat_export.unletterbox(p2, 1920, 1080)      # scale 832/1920, vertical pad (512 - 468) // 2 = 22
    assert np.allclose(out[0, 0, :2], [960.0, (256.0 - 22.0) / (832 / 1920)])
    out = mat_export.unletterbox(p2, 2048, 2048)      # scale 0.25, horizontal pad (832 - 512) // 2 = 160
    assert np.allclose(out[0, 0, :2], [(416.0 - 160.0) / 0.25, 1024.0])
