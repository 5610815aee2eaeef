"""Seeded input generators shared by make_golden.py (which feeds them to the reference) and the tests
(which feed them to the oracle / the CUDA path).  Pure numpy, no reference and no oracle code."""
import numpy as np

GEOMS = [(1920, 1080), (640, 480), (1000, 1500), (832, 512)]
N_LIFT_CASES = 24


def lift_case_inputs(ci):
    """-> bodies float32 [P,15,4] (heat-map px, as dapalib.connect returns), det_d [14,128,208],
    root_d [128,208], (img_w, img_h)."""
    rng = np.random.default_rng(100 + ci)
    P = int(rng.integers(0, 9))
    b = np.zeros((P, 15, 4), np.float32)
    b[:, :, 0] = rng.uniform(0.5, 207.4, (P, 15))
    b[:, :, 1] = rng.uniform(0.5, 127.4, (P, 15))
    b[:, :, 3] = rng.uniform(0.2, 1, (P, 15)) * (rng.uniform(size=(P, 15)) > 0.25)
    if P and ci % 5 == 0:
        b[0, 2, 3] = 0  # root missing -> dropped by register_pred
    b[b[:, :, 3] == 0] = 0
    det_d = rng.normal(0, 20, (14, 128, 208)).astype(np.float32)
    root_d = rng.uniform(1, 9, (128, 208)).astype(np.float32)
    return b, det_d, root_d, GEOMS[ci % 4]


def refine_state_dict(seed=7):
    """Seeded RefineNet weights with non-trivial BatchNorm statistics (pure generator: keys/shapes follow
    model/refinenet.py:8-17).  -> {key: float32 ndarray}"""
    rng = np.random.default_rng(seed)
    layers = [(75, 160), (160, 256), (256, 256), (256, 128)]
    sd = {}
    for i, (k, n) in enumerate(layers, start=1):
        p = "block.layer%d." % i
        sd[p + "0.weight"] = (rng.uniform(-1, 1, (n, k)) / np.sqrt(k)).astype(np.float32)
        sd[p + "0.bias"] = rng.uniform(-0.1, 0.1, n).astype(np.float32)
        sd[p + "1.weight"] = rng.uniform(0.5, 1.5, n).astype(np.float32)
        sd[p + "1.bias"] = rng.normal(0, 0.2, n).astype(np.float32)
        sd[p + "1.running_mean"] = rng.normal(0, 0.3, n).astype(np.float32)
        sd[p + "1.running_var"] = rng.uniform(0.5, 2.0, n).astype(np.float32)
        sd[p + "1.num_batches_tracked"] = np.asarray(100, np.int64)
    sd["block.layer5.weight"] = (rng.uniform(-1, 1, (45, 128)) / np.sqrt(128)).astype(np.float32)
    sd["block.layer5.bias"] = rng.uniform(-0.1, 0.1, 45).astype(np.float32)
    return sd


PRE_GEOMS = [(1920, 1080), (640, 480), (1000, 1500), (832, 512), (1664, 1024), (1280, 720), (333, 517), (2592, 1944),
             (500, 300), (831, 511), (1665, 1025), (100, 60), (2048, 1024), (416, 256), (3840, 2160), (517, 333)]


def preprocess_case_image(ci):
    """Seeded uint8 BGR test image [H,W,3] for geometry PRE_GEOMS[ci]: smooth low-frequency structure (so that the
    bilinear weights matter) plus full-range noise (so that every rounding case occurs)."""
    W, H = PRE_GEOMS[ci]
    rng = np.random.default_rng(500 + ci)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    base = 127 + 90 * np.sin(xx / (7 + ci))[:, :, None] * np.cos(yy / (11 + ci))[:, :, None] * np.array([1, 0.7, -0.8], np.float32)
    img = base + rng.normal(0, 40, (H, W, 3))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


N_GT_CASES = 12


def lift_gt_case_inputs(ci):
    """Inputs of the GT-matching branch (exps/stage3_root2/test.py:73-95, test_util.py:21-39): the lift case `ci` plus
    ground-truth bodies float64 [G,15,11] = (x, y, Z, vis, X, Y, Z, f_x, f_y, cx, cy) in network-input pixels.
    GT roots are placed near predicted roots (inside and outside the 30 px gate), with exact ties, duplicates competing for
    one prediction, and unmatched persons."""
    b, det_d, root_d, (iw, ih) = lift_case_inputs(ci)
    rng = np.random.default_rng(900 + ci)
    P = len(b)
    G = int(rng.integers(1, 7))
    gt = np.zeros((G, 15, 11), np.float64)
    gt[:, :, 0] = rng.uniform(0, 832, (G, 15))
    gt[:, :, 1] = rng.uniform(0, 512, (G, 15))
    gt[:, :, 2] = rng.uniform(100, 800, (G, 15))
    gt[:, :, 3] = 2
    gt[:, :, 4:7] = rng.normal(0, 100, (G, 15, 3))
    gt[:, :, 7], gt[:, :, 8], gt[:, :, 9], gt[:, :, 10] = 1100.0 + ci, 1105.0 + ci, iw / 2 + 3.5, ih / 2 - 2.25
    for g in range(G):
        if P and rng.uniform() < 0.8:
            p = int(rng.integers(0, P))
            r = float(rng.choice([0.0, 3.0, 12.5, 29.0, 31.0, 45.0]))
            ang = rng.uniform(0, 2 * np.pi)
            gt[g, 2, 0] = np.float64(b[p, 2, 0]) * 4 + r * np.cos(ang)
            gt[g, 2, 1] = np.float64(b[p, 2, 1]) * 4 + r * np.sin(ang)
    if G >= 2 and P and ci % 3 == 0:  # two GT persons at EXACTLY the same distance from one prediction (tie order)
        p = 0
        gt[0, 2, :2] = (np.float64(b[p, 2, 0]) * 4 + 6.0, np.float64(b[p, 2, 1]) * 4)
        gt[1, 2, :2] = (np.float64(b[p, 2, 0]) * 4 - 6.0, np.float64(b[p, 2, 1]) * 4)
    return b, det_d, root_d, (iw, ih), gt
