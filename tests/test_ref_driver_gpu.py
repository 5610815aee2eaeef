"""GPU: the reference driver's call sequence over the drop-in modules.

exps/stage3_root2/test.py imports cleanly only with the reference tree on disk (and three modules this image lacks), and
/root/reference does not exist on the GPU box, so the unmodified file cannot be executed where the shims can run.  This test
therefore walks the body of generate_3d_point_pairs (exps/stage3_root2/test.py:25-152, run_inference mode) statement by
statement - every step cites the line it stands for - with `from model.smap import SMAP` and `import dapalib` resolved to
smap_b200/shims exactly as INTEGRATION.md section 2 prescribes, a batch in the format the reference's DataLoader yields
(imgs, img_path, scales-dict of tensors; dataset/custom_dataset.py:36-56), the reference's in-place tensor idioms
(`outputs_2d[:, i] += ...`, `hmsIn[:15] /= 255`, `.cpu()`), and the reference's own host post-processing restated in oracle/
(register_pred / generate_relZ / gen_3d_pose / lift_and_refine_3d_pose / save_result, each pinned to the reference by
tests/golden).  The resulting dict must equal what the fused path (smap_b200.run_inference's engine calls) writes."""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from cases import preprocess_case_image, refine_state_dict  # noqa: E402

from oracle import lift_numpy, preprocess_numpy, refine_torch  # noqa: E402
from smap_b200 import schema  # noqa: E402

pytestmark = pytest.mark.gpu
SHIMS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "smap_b200", "shims")
NS = types.SimpleNamespace
FLIP_ORDER = [0, 1, 2, 9, 10, 11, 12, 13, 14, 3, 4, 5, 6, 7, 8]                      # dataset/data_settings.py:22
FLIP_CHANNEL = [0, 1, 2, 3, 10, 11, 12, 13, 14, 15, 4, 5, 6, 7, 8, 9, 22, 23, 24, 25, 26, 27, 16, 17, 18, 19, 20, 21]  # :33-34


def make_cfg(do_flip):
    return NS(MODEL=NS(STAGE_NUM=3, UPSAMPLE_CHANNEL_NUM=256), OUTPUT_SHAPE=(128, 208), INPUT_SHAPE=(512, 832),
              LOSS=NS(OHKM=True, TOPK=8, COARSE_TO_FINE=True), TEST_MODE="run_inference", DO_FLIP=do_flip,
              DATASET=NS(NAME="CMU", ROOT_IDX=2, KEYPOINT=NS(NUM=15, FLIP_ORDER=FLIP_ORDER), PAF=NS(NUM=14, FLIP_CHANNEL=FLIP_CHANNEL)),
              dataset=NS(STRIDE=4))


def generate_3d_point_pairs(model, refine_sd, data_loader, cfg, device, dapalib):
    """exps/stage3_root2/test.py:25-152, TEST_MODE == 'run_inference'."""
    result = dict()
    result["model_pattern"] = cfg.DATASET.NAME                                            # :32-34
    result["3d_pairs"] = []
    kpt_num = cfg.DATASET.KEYPOINT.NUM                                                    # :40
    for idx, batch in enumerate(data_loader):                                             # :42
        imgs, img_path, scales = batch                                                    # :44
        meta_data = None
        imgs = imgs.to(device)                                                            # :48
        with torch.no_grad():
            outputs_2d, outputs_3d, outputs_rd = model(imgs)                              # :50
            outputs_3d = outputs_3d.cpu()                                                 # :52-53
            outputs_rd = outputs_rd.cpu()
            if cfg.DO_FLIP:                                                               # :55-70
                imgs_flip = torch.flip(imgs, [-1])
                outputs_2d_flip, outputs_3d_flip, outputs_rd_flip = model(imgs_flip)
                outputs_2d_flip = torch.flip(outputs_2d_flip, dims=[-1])
                pair = cfg.DATASET.KEYPOINT.FLIP_ORDER + [x + kpt_num for x in cfg.DATASET.PAF.FLIP_CHANNEL]
                for i in range(len(pair)):
                    if i >= kpt_num and (i - kpt_num) % 2 == 0:
                        outputs_2d[:, i] += outputs_2d_flip[:, pair[i]] * -1
                    else:
                        outputs_2d[:, i] += outputs_2d_flip[:, pair[i]]
                outputs_2d[:, kpt_num:] *= 0.5
            for i in range(len(imgs)):                                                    # :72
                assert meta_data is None
                scale = {k: scales[k][i].numpy() for k in scales}                         # :98-103
                scale["f_x"] = scale["img_width"]
                scale["f_y"] = scale["img_width"]
                scale["cx"] = scale["img_width"] / 2
                scale["cy"] = scale["img_height"] / 2
                hmsIn = outputs_2d[i]                                                     # :105
                hmsIn[:cfg.DATASET.KEYPOINT.NUM] /= 255                                   # :111-112
                hmsIn[cfg.DATASET.KEYPOINT.NUM:] /= 127
                rDepth = outputs_rd[i][0]                                                 # :113
                pred_bodys_2d = dapalib.connect(hmsIn, rDepth, cfg.DATASET.ROOT_IDX, distFlag=True)  # :115
                if len(pred_bodys_2d) > 0:                                                # :116-118
                    bodies_hm = pred_bodys_2d.numpy()  # the x STRIDE of :117 is applied inside the restated lift (float32)
                else:
                    bodies_hm = np.zeros((0, 15, 4), np.float32)
                # :120-134  nearest up-sampling (== index // 4), register_pred, generate_relZ, gen_3d_pose
                p2, p3, rdep = lift_numpy.lift(bodies_hm, outputs_3d[i].numpy(), outputs_rd[i][0].numpy(), scale)
                if len(p2) == 0:                                                          # :130-131
                    continue
                if refine_sd is not None:                                                 # :136-140
                    p3 = refine_torch.refine(p2, p3, refine_sd, root_n=cfg.DATASET.ROOT_IDX)
                pair = dict()                                                             # :145 -> test_util.py:146-158
                pair["pred_2d"] = p2.tolist()
                pair["pred_3d"] = p3.tolist()
                pair["root_d"] = rdep.tolist()
                pair["image_path"] = img_path[i]
                pair["gt_3d"] = list()
                pair["gt_2d"] = list()
                result["3d_pairs"].append(pair)
    return result


def reference_loader(frames, names, batch_size):
    """What DataLoader(CustomDataset) yields (dataset/custom_dataset.py:27-56 + default collate): normalised fp32 images,
    names, and a dict of per-key tensors."""
    for lo in range(0, len(frames), batch_size):
        ims, scs = [], []
        for f in frames[lo:lo + batch_size]:
            t, sc = preprocess_numpy.preprocess(f)
            ims.append(torch.from_numpy(t))
            scs.append(sc)
        scales = {k: torch.tensor([float(s[k]) for s in scs], dtype=torch.float64) for k in scs[0]}
        yield torch.stack(ims), names[lo:lo + batch_size], scales


@pytest.mark.parametrize("do_flip,with_refine", [(0, False), (1, True)])
def test_reference_driver_body_over_the_shims_equals_the_fused_path(do_flip, with_refine):
    from smap_b200.engine import Engine, records_to_numpy

    sys.path.insert(0, SHIMS)
    for m in [m for m in sys.modules if m == "model" or m.startswith("model.") or m == "dapalib"]:
        del sys.modules[m]
    try:
        import dapalib
        from model.smap import SMAP
    finally:
        sys.path.remove(SHIMS)
    torch.backends.cudnn.allow_tf32 = False
    cfg = make_cfg(do_flip)
    sd = schema.make_state_dict(0, "identity")
    rsd = {k: torch.from_numpy(np.asarray(v)) for k, v in refine_state_dict().items()} if with_refine else None
    frames = [preprocess_case_image(ci) for ci in (1, 8, 13)]
    names = ["a.png", "sub/b.jpg", "sub/c.png"]
    device = torch.device("cuda")
    model = SMAP(cfg)                                                                     # test.py:190-192
    model.to(device)
    model.load_state_dict(sd)                                                             # test.py:210-212 (strict)
    model.eval()
    got = generate_3d_point_pairs(model, rsd, reference_loader(frames, names, 2), cfg, device, dapalib)

    # the fused path on the same frames (uint8 upload + GPU pre-processing + smapb_infer_device)
    eng = Engine(0, max_batch=2, in_h=512, in_w=832)
    eng.load_state_dict(sd)
    if with_refine:
        eng.load_refine_state_dict(rsd)
        eng.set_refine(True)
    fused = []
    for lo in range(0, 3, 2):
        x, sc = eng.preprocess([torch.from_numpy(f) for f in frames[lo:lo + 2]])
        rec = records_to_numpy(eng.infer_device(x, sc.cuda(), do_flip=bool(do_flip)))
        for b in range(len(rec)):
            n = int(rec["count"][b])
            if n:
                fused.append((names[lo + b], rec["pred2d"][b, :n], rec["pred3d"][b, :n], rec["root_depth"][b, :n]))
    eng.close()
    assert [p["image_path"] for p in got["3d_pairs"]] == [f[0] for f in fused]
    assert got["model_pattern"] == "CMU" and len(fused) >= 2
    for p, (name, p2, p3, rdep) in zip(got["3d_pairs"], fused):
        assert np.array_equal(np.asarray(p["pred_2d"], np.float32), p2), name           # bit-exact 2D + relative depth
        assert np.array_equal(np.asarray(p["root_d"]), rdep), name
        # 3D: float64 lift within 1e-12; with RefineNet the MLP is fp32 on both sides (fixed fmaf order vs ATen): 1e-5
        a3 = np.asarray(p["pred_3d"])
        if with_refine:
            assert np.abs(a3 - p3).max() <= 1e-5 * max(1e-6, np.abs(p3).max()), name
        else:
            np.testing.assert_allclose(a3, p3, rtol=1e-12, atol=1e-12)
        assert p["gt_3d"] == [] and p["gt_2d"] == []
    json.dumps(got)  # serialisable exactly as test.py:150-151 does
