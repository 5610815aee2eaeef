"""Golden fixture for smap_b200/mat_export.py: runs the UNMODIFIED reference converter (/root/reference/lib/eval/convert.py)
on a seeded synthetic result JSON (MuPoTS-style image paths, GT intrinsics in gt[...,3:7]) and commits the JSON together with
the bytes of the two .mat files it wrote (everything behind the 128-byte MAT-5 header, whose text holds a timestamp).
    python tests/golden/make_golden_mat.py"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/lib/eval")


def synthetic_result(seed=0):
    rng = np.random.default_rng(seed)
    pairs = []
    cases = [("data/MultiPersonTestSet/TS1/img_000000.jpg", 3, 3), ("x/TS7/img_000010.jpg", 2, 3), ("TS20/img_000003.jpg", 1, 1),
             ("TS5/img_000001.jpg", 4, 2), ("TS6/img_000002.jpg", 2, 2)]
    for name, P, G in cases:
        pred = rng.normal(0, 100, (P, 15, 4))
        pred[..., 3] = rng.uniform(0, 1, (P, 15))
        pred[..., 2] = np.abs(pred[..., 2]) + 200
        p2 = rng.uniform(0, 500, (P, 15, 4))
        p2[..., 3] = (rng.uniform(0, 1, (P, 15)) > 0.25).astype(np.float64)
        gt = rng.normal(0, 100, (G, 15, 7))
        gt[..., 3], gt[..., 4], gt[..., 5], gt[..., 6] = 1495.5, 1500.25, 962.0, 538.5
        pairs.append(dict(image_path=name, pred=pred.tolist(), gt=gt.tolist(), pred_2d=p2.tolist()))
    return {"model_pattern": "MIX", "3d_pairs": pairs}


def main():
    import convert as ref_convert  # the reference module

    text = json.dumps(synthetic_result())
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            open("in.json", "w").write(text)
            ref_convert.convert("in.json")
            m3 = open("pose3d.mat", "rb").read()
            m2 = open("pose2d.mat", "rb").read()
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "mat_cases.npz"), json=np.frombuffer(text.encode(), np.uint8),
                        pose3d=np.frombuffer(m3[128:], np.uint8), pose2d=np.frombuffer(m2[128:], np.uint8))
    print("mat golden: json %d B, pose3d.mat %d B, pose2d.mat %d B" % (len(text), len(m3), len(m2)))


if __name__ == "__main__":
    main()
