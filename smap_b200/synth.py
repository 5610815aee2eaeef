"""Synthetic association inputs (BASELINE.json config 4).

Builds the tensors the association step consumes - keypoint heat-maps + 2D PAFs
(`hms` [43,h,w], already in the /255,/127 scale that dapalib.connect expects),
the relative-depth PAFs (`det_d` [14,h,w]) and the root-depth map (`root_d` [h,w])
- for scenes of N skeletons, mirroring how the reference synthesises its training
labels (dataset/representation.py:5-21 Gaussian key-point maps,
dataset/representation.py:55-112 PAFs of unit limb vectors within
LINE_WIDTH_THRE=1 px of the segment, averaged where limbs overlap).

This is an input generator, not a checker: it contains no association logic.
"""
import numpy as np

LIMBS = [[0, 1], [0, 2], [0, 9], [9, 10], [10, 11], [0, 3], [3, 4], [4, 5],
         [2, 12], [12, 13], [13, 14], [2, 6], [6, 7], [7, 8]]

# canonical MPI-15 template in (net-input px x depth) units, y down, pelvis at origin; bone lengths follow
# extensions/association.cpp:27-31 (the table is in input pixels: the grouping divides by dsScale = 4).
_T = np.zeros((15, 2), np.float64)
_T[2] = (0, 0)            # pelvis
_T[0] = (0, -48.37)       # neck
_T[1] = (0, -48.37 - 26.42)  # head
_T[3] = (14.9, -48.37)    # l shoulder
_T[4] = (17.0, -48.37 + 31.2)
_T[5] = (18.0, -48.37 + 31.2 + 23.9)
_T[9] = (-14.9, -48.37)
_T[10] = (-17.0, -48.37 + 31.2)
_T[11] = (-18.0, -48.37 + 31.2 + 23.9)
_T[6] = (12.46, 0)        # l hip
_T[7] = (13.0, 48.2)
_T[8] = (13.5, 48.2 + 39.0)
_T[12] = (-12.46, 0)
_T[13] = (-13.0, 48.2)
_T[14] = (-13.5, 48.2 + 39.0)
_Z = np.array([0.0, 0.02, 0.0, 0.05, 0.12, 0.2, 0.03, 0.06, 0.1, -0.05, -0.1, -0.15, -0.03, 0.0, 0.05])


def make_scene(seed, persons=15, h=128, w=208, noise=0.01, sigma=1.5):
    """Returns dict(hms float32 [43,h,w], root_d float32 [h,w], det_d float32 [14,h,w],
    joints float64 [P,15,2], depth float64 [P])."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    kp = np.zeros((15, h, w), np.float64)
    paf = np.zeros((14, 2, h, w), np.float64)
    pafz = np.zeros((14, h, w), np.float64)
    cnt = np.zeros((14, h, w), np.float64)
    root_d = np.zeros((h, w), np.float64)
    depth = np.sort(rng.uniform(0.5, 2.0, persons)) + np.arange(persons) * 1e-3
    rng.shuffle(depth)
    joints = np.zeros((persons, 15, 2))
    for p in range(persons):
        d = depth[p]
        ang = rng.normal(0, 0.15)
        rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        tpl = (_T + rng.normal(0, 1.5, _T.shape)) @ rot.T / (4.0 * d)
        ext = np.abs(tpl).max(0)
        cx = rng.uniform(4 + ext[0], w - 5 - ext[0]) if w - 5 - ext[0] > 4 + ext[0] else w / 2
        lo, hi = 4 + (-tpl[:, 1].min()), h - 5 - tpl[:, 1].max()
        cy = rng.uniform(lo, hi) if hi > lo else h / 2
        j = tpl + (cx, cy)
        j[:, 0] = np.clip(j[:, 0], 2, w - 3)
        j[:, 1] = np.clip(j[:, 1], 2, h - 3)
        joints[p] = j
        zj = _Z * 10.0 + rng.normal(0, 0.5, 15)
        for k in range(15):
            kp[k] = np.maximum(kp[k], np.exp(-((xx - j[k, 0]) ** 2 + (yy - j[k, 1]) ** 2) / (2 * sigma * sigma)))
        for l, (a, b) in enumerate(LIMBS):
            v = j[b] - j[a]
            n = np.linalg.norm(v)
            if n < 1e-6:
                continue
            u = v / n
            rx, ry = xx - j[a, 0], yy - j[a, 1]
            along = rx * u[0] + ry * u[1]
            perp = np.abs(rx * u[1] - ry * u[0])
            # support is 2 px wide: the reference samples the PAF at (int)(x + 0.5) of the +0.5-offset peak
            # coordinates, i.e. up to one pixel off the segment (bodyPartConnectorBase.cu:38-39)
            m = (along >= -2) & (along <= n + 2) & (perp <= 2.0)
            paf[l, 0][m] += u[0]
            paf[l, 1][m] += u[1]
            pafz[l][m] += zj[b] - zj[a]
            cnt[l][m] += 1
        disc = (xx - j[2, 0]) ** 2 + (yy - j[2, 1]) ** 2 <= 9.0
        root_d[disc] = d
    nz = cnt > 0
    paf[:, 0][nz] /= cnt[nz]
    paf[:, 1][nz] /= cnt[nz]
    pafz[nz] /= cnt[nz]
    hms = np.concatenate([kp, paf.reshape(28, h, w)], 0)
    hms += rng.normal(0, noise, hms.shape)
    det_d = pafz + rng.normal(0, noise, pafz.shape)
    root_d = root_d + rng.normal(0, noise * 0.1, root_d.shape)
    return dict(hms=hms.astype(np.float32), root_d=root_d.astype(np.float32), det_d=det_d.astype(np.float32),
                joints=joints, depth=depth)


def make_batch(seed, batch, persons=15, h=128, w=208):
    scenes = [make_scene(seed * 1000 + i, persons, h, w) for i in range(batch)]
    return (np.stack([s["hms"] for s in scenes]), np.stack([s["root_d"] for s in scenes]),
            np.stack([s["det_d"] for s in scenes]))
