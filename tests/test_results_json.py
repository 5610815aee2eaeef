"""CPU: the native result writer (smapb_json_*) against Python's json.dump of the reference's own result dict.

The expected text is produced the way exps/stage3_root2/test.py does it: save_result (test_util.py:146-158, restated
below because the reference is absent on the GPU box; tests/golden/make_golden.py pins the restatement against the
real function in results_json.txt) followed by json.dump."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

from smap_b200.engine import RECORD_DTYPE  # noqa: E402
from smap_b200.results import ResultWriter, result_file_name  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def save_result(pred_bodys_2d, pred_bodys_3d, gt_bodys, pred_rdepths, img_path, result):
    pair = {"pred_2d": pred_bodys_2d.tolist(), "pred_3d": pred_bodys_3d.tolist(), "root_d": pred_rdepths.tolist(),
            "image_path": img_path, "gt_3d": [], "gt_2d": []}
    result["3d_pairs"].append(pair)


def python_json(records, paths, name):
    result = {"model_pattern": name, "3d_pairs": []}
    for r, p in zip(records, paths):
        n = int(r["count"])
        if n == 0:
            continue
        save_result(r["pred2d"][:n], r["pred3d"][:n], None, r["root_depth"][:n], p, result)
    return json.dumps(result)


def make_records(seed, B):
    rng = np.random.default_rng(seed)
    rec = np.zeros(B, RECORD_DTYPE)
    for b in range(B):
        n = int(rng.integers(0, 6)) if b != 1 else 0
        rec["count"][b] = n
        rec["pred2d"][b, :n] = rng.uniform(-50, 900, (n, 15, 4)).astype(np.float32)
        rec["pred3d"][b, :n] = rng.normal(0, 300, (n, 15, 4))
        rec["root_depth"][b, :n] = rng.uniform(100, 900, n)
    return rec


SPECIAL = [0.0, -0.0, 1.0, -1.5, 1e16, 1e15, 123456789012345680.0, 9999999999999998.0, 1e-4, 9.999e-5, 1e-5, 5e-324,
           1.7976931348623157e308, 0.1, 1 / 3, 2 ** 53, 1e22, 1e23, 123456.789, float(np.float32(0.3)), 1e-7, 100.0,
           float("inf"), float("-inf"), float("nan"), 4.35, 0.000123456, 12345678.9, 1.5e300, -2.5e-300]


def test_float_repr_and_layout_match_python(tmp_path):
    rec = make_records(1, 7)
    # plant the special values (shortest-repr corner cases, exponent thresholds, non-finite) into person 0 of image 0
    rec["count"][0] = max(1, rec["count"][0])
    flat3 = rec["pred3d"][0, 0].reshape(-1)
    flat3[:len(SPECIAL)] = SPECIAL
    flat2 = rec["pred2d"][0, 0].reshape(-1)
    flat2[:8] = np.array([0.0, -0.0, 1e-5, 3.4028235e38, 1.17549435e-38, 16777216.0, 0.1, 1e16], np.float32)
    paths = ["a/b.jpg", 'q"uo\\te.png', "", "café/東京.jpg", "tab\tnl\n\x01.jpg", "\U0001F600.jpeg", "z.jpg"]
    out = os.path.join(tmp_path, "r.json")
    with ResultWriter(out, "CMU") as w:
        w.append(rec[:3], paths[:3])
        w.append(rec[3:].view(np.uint8).reshape(4, -1), paths[3:])   # raw byte view of the records works too
    got = open(out, "rb").read().decode("ascii")
    assert got == python_json(rec, paths, "CMU")


def test_empty_result_and_all_empty_images(tmp_path):
    out = os.path.join(tmp_path, "e.json")
    with ResultWriter(out, "MIX") as w:
        w.append(np.zeros(3, RECORD_DTYPE), ["a", "b", "c"])
        w.append(np.zeros(0, RECORD_DTYPE), [])
    assert open(out).read() == json.dumps({"model_pattern": "MIX", "3d_pairs": []})


def test_many_batches_round_trip(tmp_path):
    out = os.path.join(tmp_path, "m.json")
    recs, paths = [], []
    with ResultWriter(out, "CMU") as w:
        for s in range(40):   # > 1 MB of text: exercises the buffered flushes
            r = make_records(100 + s, 8)
            p = ["dir%d/img_%04d.jpg" % (s, i) for i in range(8)]
            w.append(r, p)
            recs.append(r)
            paths += p
    text = open(out).read()
    assert text == python_json(np.concatenate(recs), paths, "CMU")
    back = json.loads(text)
    assert back["model_pattern"] == "CMU" and len(back["3d_pairs"]) == sum(int((r["count"] > 0).sum()) for r in recs)


def test_restated_save_result_is_pinned_to_the_reference():
    """results_json.txt was written by the unmodified reference save_result + json.dump (tests/golden/make_golden.py)."""
    rec = make_records(1, 7)
    paths = ["img_%d.jpg" % i for i in range(7)]
    assert python_json(rec, paths, "CMU") == open(os.path.join(GOLD, "results_json.txt")).read()


def test_result_file_name():
    assert result_file_name("/o", suffix="x") == "/o/stage3_root2_run_inference_test_x.json"


def test_random_bit_patterns_print_like_python(tmp_path):
    """Shortest round-trip digits + CPython's layout rule over the whole double / float range (incl. subnormals, huge and
    tiny exponents, NaN and infinities): 60 k random bit patterns."""
    rng = np.random.default_rng(12345)
    rec = np.zeros(4, RECORD_DTYPE)
    rec["count"] = 127
    rec["pred3d"] = rng.integers(0, 2 ** 64, rec["pred3d"].shape, dtype=np.uint64).view(np.float64)
    rec["pred2d"] = rng.integers(0, 2 ** 32, rec["pred2d"].shape, dtype=np.uint64).astype(np.uint32).view(np.float32)
    rec["root_depth"] = rng.integers(0, 2 ** 64, rec["root_depth"].shape, dtype=np.uint64).view(np.float64)
    # values near the exponent-format thresholds and round numbers
    edge = np.array([10.0 ** k for k in range(-8, 24)] + [9.5 * 10.0 ** k for k in range(-8, 20)], np.float64)
    rec["pred3d"][0, 0].reshape(-1)[:] = np.resize(edge, 60)
    paths = ["p%d.jpg" % i for i in range(4)]
    out = os.path.join(tmp_path, "rnd.json")
    with ResultWriter(out, "CMU") as w:
        w.append(rec, paths)
    assert open(out).read() == python_json(rec, paths, "CMU")
