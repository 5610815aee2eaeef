"""GPU: smapb_preprocess / smapb_preprocess_host against the oracle and the reference digests: bit-exact."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from cases import PRE_GEOMS, preprocess_case_image  # noqa: E402

from oracle import preprocess_numpy as P  # noqa: E402
from smap_b200.engine import scale_row  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def eng():
    from smap_b200.engine import Engine

    e = Engine(0, max_batch=4, in_h=512, in_w=832)
    yield e
    e.close()


def test_all_reference_geometries_bit_exact(eng):
    gold = json.load(open(os.path.join(GOLD, "preprocess_digests.json")))
    imgs = [preprocess_case_image(ci) for ci in range(len(PRE_GEOMS))]
    for lo in range(0, len(imgs), 4):
        chunk = imgs[lo:lo + 4]
        # alternate device-resident and host images: both entry points
        feed = [torch.from_numpy(im).cuda() if (lo + i) % 2 == 0 else im for i, im in enumerate(chunk)]
        out, scales = eng.preprocess(feed)
        out = out.cpu().numpy()
        for i, im in enumerate(chunk):
            ci = lo + i
            want, sc = P.preprocess(im)
            assert np.array_equal(out[i], want), "geometry %s differs from the oracle" % (PRE_GEOMS[ci],)
            assert hashlib.sha256(np.ascontiguousarray(out[i]).tobytes()).hexdigest() == gold["c%d" % ci]["sha256"]
            assert np.array_equal(scales[i].numpy(), scale_row(sc))


def test_same_geometry_reuses_tables_and_random_noise(eng):
    rng = np.random.default_rng(9)
    for _ in range(3):
        im = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
        out, _ = eng.preprocess([torch.from_numpy(im).cuda()])
        assert np.array_equal(out[0].cpu().numpy(), P.preprocess(im)[0])


def test_preprocess_feeds_the_whole_path(eng):
    """uint8 frames -> preprocess -> infer_device equals feeding the oracle-preprocessed tensor."""
    from smap_b200 import schema
    from smap_b200.engine import records_to_numpy

    eng.load_state_dict(schema.make_state_dict(0, "identity"))
    ims = [preprocess_case_image(0), preprocess_case_image(5)]
    x, scales = eng.preprocess([torch.from_numpy(i).cuda() for i in ims])
    rec = records_to_numpy(eng.infer_device(x, scales.cuda()))
    xo = torch.from_numpy(np.stack([P.preprocess(i)[0] for i in ims])).cuda()
    so = torch.from_numpy(np.stack([scale_row(P.preprocess(i)[1]) for i in ims])).cuda()
    ref = records_to_numpy(eng.infer_device(xo, so))
    assert rec.tobytes() == ref.tobytes()
