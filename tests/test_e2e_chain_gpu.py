"""GPU: the quantity north_star names - final 3D joint positions - compared END TO END: the full oracle chain
(oracle fp32 backbone -> reference-order rescale -> oracle association -> oracle lift; exps/stage3_root2/test.py:50-134)
against the fused path (smapb_infer_device), both starting from the same frames and the same state dict.

Unlike tests/test_pipeline_gpu.py (which feeds OUR backbone tensors to the oracle association and therefore demands
bit-exactness), the two chains here differ by the backbone's 1.5e-4 .. 2e-4 (of the tensor max): a heat-map value that
close to the 0.2 threshold or to a neighbour (strict `>` NMS, nmsBase.cu:24-49) can flip a candidate.  With random-init
heads the maps are noise (0 .. 178 peaks per channel), the worst case for that; the test MEASURES the flip
rate, bounds it, and checks that wherever the candidate sets agree the final 3D joints agree to 1e-3.  The second test
repeats the comparison in the trained-like regime (config-4 scenes: isolated Gaussian peaks) by injecting the MEASURED
backbone error field into the association input.  Flip statistics are printed (-s) and summarised in DESIGN.md."""
import numpy as np
import pytest
import torch

from oracle import assoc, lift_numpy, smap_torch
from smap_b200 import schema
from smap_b200.synth import make_scene

pytestmark = pytest.mark.gpu
TOL_3D = 1e-3  # north_star: final 3D joint positions within 1e-3 relative (here: of the largest |coordinate| of the frame)


@pytest.fixture(scope="module")
def setup():
    from smap_b200.engine import Engine

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = schema.make_state_dict(0, "identity")
    e = Engine(0, max_batch=8, in_h=512, in_w=832)
    e.load_state_dict(sd)
    sd_dev = {k: v.cuda() for k, v in smap_torch.make_state_dict(0, "identity").items()}
    yield e, sd_dev
    e.close()


def peak_sets(peaks):
    """peaks [15,128,3] -> set of (channel, x_q, y_q): sub-pixel positions quantised to 1/8 px identify a candidate."""
    out = set()
    for c in range(15):
        n = int(peaks[c, 0, 0])
        for k in range(1, n + 1):
            out.add((c, int(round(float(peaks[c, k, 0]) * 8)), int(round(float(peaks[c, k, 1]) * 8))))
    return out


def compare_frame(p3_a, p3_b):  # noqa: D103
    """max |a-b| / max|b| over persons x joints of two [P,15,4] float64 arrays of equal shape"""
    den = max(np.abs(p3_b[..., :3]).max(), 1e-12)
    return np.abs(p3_a[..., :3] - p3_b[..., :3]).max() / den


def test_full_oracle_chain_vs_fused_path_random_init(setup, capsys):
    from smap_b200.engine import records_to_numpy, scale_row

    eng, sd_dev = setup
    B = 8
    x = schema.make_input(B, 512, 832, seed=1).cuda()  # configs[1]: the bench workload's first batch
    sc = lift_numpy.default_scale(1920, 1080)
    scales = torch.from_numpy(np.stack([scale_row(sc)] * B)).cuda()
    rec = records_to_numpy(eng.infer_device(x, scales))
    hm_o, dd_o, rd_o = smap_torch.smap_forward(sd_dev, x)  # oracle backbone, fp32 (TF32 off)
    hm_f, dd_f, rd_f = eng.forward(x)
    torch.cuda.synchronize()
    for name, a, b in (("hm2d", hm_f, hm_o), ("detd", dd_f, dd_o), ("rootd", rd_f, rd_o)):
        assert ((a - b).abs().max() / b.abs().max()).item() < 1e-3, name
    hms_o = smap_torch.rescale_reference_cuda(hm_o.clone())
    hms_f = eng.merge_scale(hm_f.clone(), None, True)
    torch.cuda.synchronize()
    tot = flips = same_count = 0
    persons_o = persons_matched = 0
    worst = 0.0
    for i in range(B):
        bod_o, pk_o, _ = assoc.connect(hms_o[i].cpu().numpy(), rd_o[i, 0].cpu().numpy(), return_all=True)
        _, pk_f, _ = assoc.connect(hms_f[i].cpu().numpy(), rd_f[i, 0].cpu().numpy(), return_all=True)
        so, sf = peak_sets(pk_o), peak_sets(pk_f)
        tot += len(so | sf)
        flips += len(so ^ sf)
        p2, p3, rdep = lift_numpy.lift(bod_o, dd_o[i].cpu().numpy(), rd_o[i, 0].cpu().numpy(), sc)
        n = int(rec["count"][i])
        same_count += int(n == len(p2))
        # person level: a person of the oracle chain is "the same person" in the fused output when every joint was built
        # from the same candidates (same visibility pattern, 2D positions within 0.05 input px); its 3D joints must then
        # agree to 1e-3 of the frame's largest coordinate
        ours2, ours3 = rec["pred2d"][i, :n], rec["pred3d"][i, :n]
        den = max(np.abs(p3[..., :3]).max(), 1e-12) if len(p3) else 1.0
        persons_o += len(p2)
        for k in range(len(p2)):
            d = np.abs(ours2[:, :, :2] - p2[k][None, :, :2]).max(axis=(1, 2)) if n else np.zeros(0)
            vis = (ours2[:, :, 3] != 0) == (p2[k][None, :, 3] != 0) if n else np.zeros((0, 15), bool)
            cand = [j for j in range(n) if d[j] < 0.05 and vis[j].all()]
            if cand:
                persons_matched += 1
                worst = max(worst, np.abs(ours3[cand[0], :, :3] - p3[k, :, :3]).max() / den)
    rate = flips / max(tot, 1)
    with capsys.disabled():
        print("\n[e2e random-init] frames=%d candidates=%d flipped=%d (%.3f %%) person-count-equal=%d/%d "
              "persons-built-from-identical-candidates=%d/%d worst-3D-rel-among-them=%.2e" %
              (B, tot, flips, 100 * rate, same_count, B, persons_matched, persons_o, worst))
    assert rate < 0.02, "candidate flip rate %.4f" % rate  # noise maps: ~1 % of the candidates sit on a decision boundary
    assert persons_matched > 0.3 * persons_o
    assert worst < TOL_3D


def test_trained_like_scenes_with_measured_backbone_error_injected(setup, capsys):
    """Config-4 scenes (15 persons, Gaussian key-point maps, unit-vector PAFs) are what a TRAINED head emits.  The backbone
    cannot be trained here, so the measured difference field (fused backbone - oracle backbone, same frames) is added to
    the scene tensors at the same relative magnitude it has on the real outputs, and the fused association + lift must
    return the same persons, the same limb assignments and 3D joints within 1e-3 of the oracle chain on the clean scene."""
    from smap_b200.engine import records_to_numpy, scale_row  # noqa: F401

    eng, sd_dev = setup
    B = 8
    x = schema.make_input(B, 512, 832, seed=2).cuda()
    hm_o, dd_o, rd_o = smap_torch.smap_forward(sd_dev, x)
    hm_f, dd_f, rd_f = eng.forward(x)
    torch.cuda.synchronize()
    e_hm = ((hm_f - hm_o) / hm_o.abs().max()).cpu().numpy()   # relative error fields, ~1.5e-4 peak
    e_dd = ((dd_f - dd_o) / dd_o.abs().max()).cpu().numpy()
    e_rd = ((rd_f - rd_o) / rd_o.abs().max()).cpu().numpy()
    sc = lift_numpy.default_scale(1920, 1080)
    worst, persons = 0.0, 0
    scenes = [make_scene(100 + i, 15) for i in range(B)]
    hms = np.stack([s["hms"] for s in scenes])
    dd = np.stack([s["det_d"] for s in scenes])
    rd = np.stack([s["root_d"] for s in scenes])
    hms_p = (hms + e_hm * np.abs(hms).max()).astype(np.float32)
    dd_p = (dd + e_dd * np.abs(dd).max()).astype(np.float32)
    rd_p = (rd + e_rd[:, 0] * np.abs(rd).max()).astype(np.float32)
    bodies, counts = eng.connect(torch.from_numpy(hms_p).cuda(), torch.from_numpy(rd_p).cuda())
    scales = torch.from_numpy(np.stack([scale_row(sc)] * B)).cuda()
    p2, p3, rdp, co = eng.lift(bodies, counts, torch.from_numpy(dd_p).cuda(), torch.from_numpy(rd_p).cuda(), scales)
    torch.cuda.synchronize()
    p3, co, bodies = p3.cpu().numpy(), co.cpu().numpy(), bodies.cpu().numpy()
    for i in range(B):
        ref_b = assoc.connect(hms[i], rd[i])
        r2, r3, rr = lift_numpy.lift(ref_b, dd[i], rd[i], sc)
        assert int(co[i]) == len(r3), "frame %d: %d persons vs %d" % (i, int(co[i]), len(r3))
        n = len(r3)
        persons += n
        assert np.array_equal(p3[i, :n, :, 3] != 0, r3[..., 3] != 0), "frame %d: limb assignment differs" % i
        # same candidates: sub-pixel positions move by at most the centroid's sensitivity to the injected error
        assert np.abs(bodies[i, :n, :, :2] - ref_b[..., :2]).max() < 1e-2
        worst = max(worst, compare_frame(p3[i, :n], r3))
    with capsys.disabled():
        print("\n[e2e trained-like] frames=%d persons=%d worst-3D-rel=%.2e (injected error: hm %.1e, det_d %.1e, root_d %.1e)" %
              (B, persons, worst, np.abs(e_hm).max(), np.abs(e_dd).max(), np.abs(e_rd).max()))
    assert persons >= 8 * 10
    assert worst < TOL_3D
