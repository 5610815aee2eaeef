"""GPU: images on disk -> result JSON (the reference's `test.py -t run_inference` flow, smap_b200/run_inference.py).

The pieces (pre-processing, whole path, RefineNet, JSON text) are each checked against the oracle / the reference goldens
in their own tests; this one checks the plumbing: file discovery and naming, batching with a ragged last batch, flip,
refine and the result file, by rebuilding the expected file from the library's primitives."""
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from cases import preprocess_case_image, refine_state_dict  # noqa: E402

from smap_b200 import schema  # noqa: E402

pytestmark = pytest.mark.gpu


def test_cli_writes_the_reference_result_file(tmp_path, monkeypatch):
    cv2 = pytest.importorskip("cv2")
    from smap_b200.engine import Engine, records_to_numpy
    from smap_b200.results import ResultWriter
    from smap_b200.run_inference import main

    monkeypatch.setenv("SMAPB_NO_AUTOTUNE", "1")   # two handles must choose the same tile shapes for a byte comparison
    data = tmp_path / "imgs"
    (data / "sub").mkdir(parents=True)
    frames = {"a.png": preprocess_case_image(1), "sub/b.jpg": preprocess_case_image(8), "sub/c.png": preprocess_case_image(13)}
    for k, im in frames.items():
        assert cv2.imwrite(str(data / k), im)
    (data / "notes.txt").write_text("not an image")
    sd = schema.make_state_dict(0, "identity")
    rsd = {k: torch.from_numpy(np.asarray(v)) for k, v in refine_state_dict().items()}
    torch.save({"model": sd}, tmp_path / "smap.pth")
    torch.save(rsd, tmp_path / "refine.pth")
    out_dir = tmp_path / "out"
    assert main(["-p", str(tmp_path / "missing.pth"), "--dataset_path", str(data)]) == 1
    rc = main(["-p", str(tmp_path / "smap.pth"), "-rp", str(tmp_path / "refine.pth"), "--dataset_path", str(data),
               "--batch_size", "2", "--do_flip", "1", "--json_name", "t", "--output_dir", str(out_dir)])
    assert rc == 0
    got_text = open(out_dir / "stage3_root2_run_inference_test_t.json").read()

    # expected: same order (sorted full paths), same batches, library primitives
    order = sorted(frames, key=lambda k: str(data / k))
    eng = Engine(0, max_batch=2, in_h=512, in_w=832)
    eng.load_state_dict(sd)
    eng.load_refine_state_dict(rsd)
    eng.set_refine(True)
    exp = tmp_path / "expected.json"
    total = 0
    with ResultWriter(str(exp), "CMU") as w:
        for lo in range(0, 3, 2):
            names = order[lo:lo + 2]
            ims = [torch.from_numpy(cv2.imread(str(data / n), cv2.IMREAD_COLOR)) for n in names]
            x, sc = eng.preprocess(ims)
            rec = eng.infer_device(x, sc.cuda(), do_flip=True).cpu()
            total += int(records_to_numpy(rec)["count"].sum())
            w.append(rec, names)
    eng.close()
    assert got_text == open(exp).read()
    res = json.loads(got_text)
    assert res["model_pattern"] == "CMU"
    assert [p["image_path"] for p in res["3d_pairs"]] == [n for n in order if any(q["image_path"] == n for q in res["3d_pairs"])]
    assert sum(len(p["pred_3d"]) for p in res["3d_pairs"]) == total
    for p in res["3d_pairs"]:
        assert p["gt_3d"] == [] and p["gt_2d"] == [] and len(p["root_d"]) == len(p["pred_3d"]) == len(p["pred_2d"])
        assert all(len(b) == 15 and all(len(j) == 4 for j in b) for b in p["pred_3d"])
