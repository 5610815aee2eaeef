"""Result serialisation, second half: result JSON -> the .mat files the reference's MATLAB evaluation
reads (lib/eval/convert.py -> ./pose3d.mat {'preds_3d_kpt'}, ./pose2d.mat {'preds_2d_kpt'}; lib/eval/mupots_smap.m).

What the reference does per JSON entry (lib/eval/convert.py:13-84), restated array-wise:
  * sequence number from the 'TS<n>/' part of image_path -> image size 2048x2048 (TS1-5) or 1920x1080 (TS6-20);
  * 2D joints back from the 832x512 letterbox to image pixels: subtract the pad on the letterboxed axis
    ((crop - size*scale) // 2), divide by scale = min(832/w, 512/h)                                          (:43-60)
  * 3D joints re-projected through the GT intrinsics K = [[f,0,cx],[0,f,cy],[0,0,1]] (f = gt[0,0,4], cx, cy = gt[0,0,5:7]):
    (X,Y,Z) = Z * K^-1 (x, y, 1)^T for joints with a 2D score, predicted values kept otherwise               (:62-76)
  * cm -> mm (x10) on X, Y, Z, score untouched                                                               (:81-82)
The reference's converter reads the keys 'pred' / 'gt' / 'pred_2d'; its own save_result (exps/stage3_root2/test_util.py:146-158)
writes 'pred_3d' / 'gt_3d' / 'pred_2d' - both spellings are accepted here.  Host-side, offline (numpy + scipy.io.savemat, the
reference's own writer): this is file conversion for the MATLAB scripts, not part of the GPU path."""
import json
import os

import numpy as np

CROP_W, CROP_H = 832, 512  # lib/eval/convert.py:43-44


def sequence_geometry(image_path):
    """-> (name from 'TS' on, ts number, image width, image height)   (lib/eval/convert.py:15-23)"""
    name = image_path[image_path.index("TS"):]
    ts = int(name[2:name.index("/")])
    if ts < 6:
        return name, ts, 2048, 2048
    if ts <= 20:
        return name, ts, 1920, 1080
    raise NotImplementedError("MuPoTS sequence TS%d" % ts)


def unletterbox(pred_2d, width, height):
    """[P,15,4] network-input pixels -> image pixels (in place on a float64 copy)   (lib/eval/convert.py:43-60)"""
    p2 = np.array(pred_2d, dtype=np.float64)
    scale = min(CROP_W / float(width), CROP_H / float(height))
    adj = np.array([0, 0])
    if height * scale < CROP_H:
        adj = np.array([0, (CROP_H - height * scale) // 2])
    if width * scale < CROP_W:
        adj = np.array([(CROP_W - width * scale) // 2, 0])
    if p2.size:
        xy = p2[:, :, :2] - adj[None, None, :]
        p2[:, :, :2] = xy / scale
    return p2


def reproject(pred_3d, p2_img, gt):
    """Z-preserving re-projection through the GT intrinsics   (lib/eval/convert.py:36-41,62-76)"""
    intri = gt[0, 0, 3:7]
    K = np.array([[intri[1], 0, intri[2]], [0, intri[1], intri[3]], [0, 0, 1]])
    iK = np.linalg.inv(K)
    out = pred_3d.copy()
    n = min(pred_3d.shape[0], len(p2_img))
    for ih in range(n):
        for ij in range(pred_3d.shape[1]):
            if p2_img[ih, ij][3] == 0:
                continue  # no 2D evidence: the predicted joint stays
            ray = np.array([p2_img[ih, ij][0], p2_img[ih, ij][1], 1]).reshape([3, 1])
            out[ih, ij, :3] = (out[ih, ij, 2] * iK @ ray).squeeze()  # same expression (BLAS gemv): bit-identical values
    return out


def convert_entries(pairs):
    """-> (pose3d, pose2d, counts): dicts keyed by the 'TS...' image name, in JSON order"""
    pose3d, pose2d, counts = {}, {}, {}
    for entry in pairs:
        name, ts, width, height = sequence_geometry(entry["image_path"])
        counts[ts] = counts.get(ts, 0) + 1
        pred_3d = np.array(entry["pred"] if "pred" in entry else entry["pred_3d"])
        gt = np.array(entry["gt"] if "gt" in entry else entry["gt_3d"])
        p2 = unletterbox(entry["pred_2d"], width, height)
        p3 = reproject(pred_3d, p2, gt)
        p3 = p3 * 10
        p3[:, :, 3] /= 10
        pose3d[name] = p3
        pose2d[name] = p2
    return pose3d, pose2d, counts


def convert(path, out_dir="."):
    """lib/eval/convert.py:convert - writes <out_dir>/pose3d.mat and <out_dir>/pose2d.mat; returns the two dicts."""
    import scipy.io as scio

    with open(path, "r") as f:
        data = json.load(f)
    pose3d, pose2d, _ = convert_entries(data["3d_pairs"])
    scio.savemat(os.path.join(out_dir, "pose3d.mat"), {"preds_3d_kpt": pose3d})
    scio.savemat(os.path.join(out_dir, "pose2d.mat"), {"preds_2d_kpt": pose2d})
    return pose3d, pose2d


if __name__ == "__main__":
    import sys

    convert(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ".")
