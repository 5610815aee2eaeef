"""ORACLE - TEST INFRASTRUCTURE ONLY.

numpy restatement of the reference 3D lift, rows C1-C4 of SURVEY.md section 8:
  stride + nearest upsample   exps/stage3_root2/test.py:117,120-126
  register_pred (no GT)       exps/stage3_root2/test_util.py:18-42 (branch :41)
  generate_relZ / chain_bones exps/stage3_root2/test_util.py:45-86
  gen_3d_pose                 exps/stage3_root2/test_util.py:89-99
  get_3d_points/back_projection  lib/utils/post_3d.py:4-27

The restatement never materialises the 4x nearest-neighbour upsampled maps:
cv2.resize(..., INTER_NEAREST) by exactly 4 is index//4.  dtype
behaviour (float32 arrays, float64 islands) follows what the reference code does
under numpy>=2 (NEP 50) - the only numpy the reference can execute under in this
image; tests/golden/make_golden.py pins it against the real reference functions.
"""
import numpy as np

LIMBS = [[0, 1], [0, 2], [0, 9], [9, 10], [10, 11], [0, 3], [3, 4], [4, 5],
         [2, 12], [12, 13], [13, 14], [2, 6], [6, 7], [7, 8]]  # dataset/data_settings.py:27-31
STRIDE = 4


def default_scale(img_w=1920, img_h=1080, net_w=832, net_h=512):
    """Scale dict as built by dataset/custom_dataset.py:45-54 + test.py:99-103."""
    s = min(net_w / img_w, net_h / img_h)
    return dict(scale=s, img_width=img_w, img_height=img_h, net_width=net_w, net_height=net_h,
                f_x=float(img_w), f_y=float(img_w), cx=img_w / 2, cy=img_h / 2)


def register_pred_gt(pred, gt_bodys, root_n=2):
    """register_pred with ground truth (exps/stage3_root2/test_util.py:21-39): greedy one-to-one matching of GT roots to
    predicted roots by ascending pixel distance below 30; entries are visited in ascending (distance, row-major index)
    order - the reference finds the minimum, visits every entry equal to it in np.where order, overwrites them with 50 and
    repeats.  Returns float64 [G,15,4]: row g = the matched prediction or zeros."""
    root_gt = np.asarray(gt_bodys, np.float64)[:, root_n, :2]
    root_pd = pred[:, root_n, :2]
    diff = root_gt[:, None, :] - root_pd[None, :, :]          # float64 - float32 -> float64
    dist = np.sqrt(diff[:, :, 0] * diff[:, :, 0] + diff[:, :, 1] * diff[:, :, 1])  # np.linalg.norm(axis=2)
    G, P = dist.shape
    order = sorted((dist[g, p], g * P + p) for g in range(G) for p in range(P) if dist[g, p] < 30)
    corres = -np.ones(G, np.int64)
    occupied = np.zeros(P, bool)
    for _d, idx in order:
        g, p = divmod(idx, P)
        if corres[g] >= 0 or occupied[p]:
            continue
        corres[g] = p
        occupied[p] = True
    out = np.zeros((G, pred.shape[1], 4), np.float64)
    for g in range(G):
        if corres[g] >= 0:
            out[g] = pred[corres[g]]
    return out


def lift(bodies_hm, det_d, root_d, scale, root_n=2, gt_bodys=None):
    """bodies_hm: float32 [P,15,4] from dapalib.connect (heat-map pixels).
    det_d: float32 [14,h,w]; root_d: float32 [h,w]; scale: dict.
    Returns (pred_2d float32 [P',15,4] with z filled in, pred_3d float64 [P',15,4],
             root_depth float64 [P']).
    gt_bodys (float64 [G,15,>=4], network-input pixels): the GT-matching branch of register_pred - rows follow the GT
    order, everything downstream is float64 (np.zeros(..., np.float), test_util.py:35), pred_2d is returned as float64."""
    if len(bodies_hm) == 0:
        dt = np.float32 if gt_bodys is None else np.float64
        return (np.zeros((0, 15, 4), dt), np.zeros((0, 15, 4), np.float64), np.zeros((0,), np.float64))
    pred = np.array(bodies_hm, dtype=np.float32, copy=True)
    pred[:, :, :2] *= np.float32(STRIDE)  # test.py:117 (float32 tensor op)
    if gt_bodys is None:
        pred = pred[pred[:, root_n, 3] != 0]  # test_util.py:41
    else:
        pred = register_pred_gt(pred, gt_bodys, root_n)
    P = len(pred)
    sc = np.float64(scale["scale"])
    fx = np.float64(scale["f_x"])
    root_depth = np.zeros(P, np.float64)
    for i in range(P):
        body = pred[i]
        if body[root_n][3] > 0:
            ry, rx = int(body[root_n][1]), int(body[root_n][0])
            root_depth[i] = root_d[ry // STRIDE, rx // STRIDE] * sc * fx  # test_util.py:66
            dz = np.zeros(len(LIMBS), np.float64)
            for k, (a, b) in enumerate(LIMBS):
                src, dst = body[a], body[b]
                if dst[3] > 0 and src[3] > 0:
                    xs = np.round(np.linspace(src[0], dst[0], num=10)).astype(np.intp)  # float32 linspace
                    ys = np.round(np.linspace(src[1], dst[1], num=10)).astype(np.intp)
                    v = det_d[k, ys // STRIDE, xs // STRIDE].astype(np.float32)
                    lo, hi = np.percentile(v, [10, 90])  # float64, linear interpolation
                    v[v < lo] = lo
                    v[v > hi] = hi
                    dz[k] = np.mean(v)  # float32 pairwise sum
            # chain_bones, test_util.py:45-57 (writes float32 column 2 in place)
            body[2][2] = 0
            body[0][2] = body[2][2] - dz[1]
            body[1][2] = body[0][2] + dz[0]
            for k in range(2, len(LIMBS)):
                a, b = LIMBS[k]
                body[b][2] = body[a][2] + dz[k]
    # gen_3d_pose, test_util.py:89-99
    bodys = pred.copy()
    bodys[:, :, 0] = bodys[:, :, 0] / sc - (scale["net_width"] / sc - scale["img_width"]) / 2
    bodys[:, :, 1] = bodys[:, :, 1] / sc - (scale["net_height"] / sc - scale["img_height"]) / 2
    out = np.zeros(bodys.shape, np.float64)
    out[:, :, 3] = bodys[:, :, 3]
    fxx, fyy = np.float64(scale["f_x"]), np.float64(scale["f_y"])
    cx, cy = np.float64(scale["cx"]), np.float64(scale["cy"])
    for i in range(P):
        if bodys[i][root_n][3] == 0:
            continue
        bodys[i][:, 2] += root_depth[i]  # post_3d.py:25 (float32 in place)
        d = bodys[i][:, 2]
        out[i][:, 0] = (bodys[i][:, 0] - cx) * d / fxx  # post_3d.py:13-15
        out[i][:, 1] = (bodys[i][:, 1] - cy) * d / fyy
        out[i][:, 2] = d
    out[out[:, :, 3] == 0] = 0  # test_util.py:95-98
    return pred, out, root_depth
