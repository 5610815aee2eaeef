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
