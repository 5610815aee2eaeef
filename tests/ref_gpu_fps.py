"""Reference single-GPU end-to-end FPS on this box (SURVEY.md 8(d) CPU-baseline row (iii): the denominator of north_star's
">= 10x the reference's single-GPU end-to-end FPS").  Not a pytest file: `python tests/ref_gpu_fps.py [frames]`.

The reference path as exps/stage3_root2/test.py runs it: eager PyTorch fp32 backbone on the GPU (cuDNN, PyTorch's default
TF32 setting for convolutions), the UNMODIFIED reference association extension (oracle/_ref/dapalib_ref.so, compiled from
/root/reference by oracle/build_ref.py) called per image, numpy lift on the host.  The backbone module is the functional
restatement oracle/smap_torch.py (same ATen ops as model/smap.py: conv2d, batch_norm, relu, interpolate, add), because
/root/reference does not travel to the GPU box.  Lives under tests/ because it imports oracle/."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import build_ref, lift_numpy, smap_torch


def main():
    B = 8
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    ref = build_ref.load_ref()
    if ref is None:
        print("oracle/_ref/dapalib_ref.so is not built")
        return 1
    sd = {k: v.cuda() for k, v in smap_torch.make_state_dict(0, "identity").items()}
    scale = lift_numpy.default_scale(1920, 1080)
    xs = [smap_torch.make_input(B, 512, 832, seed=1 + s) for s in range(steps + 1)]

    def step(x):
        persons = 0
        with torch.no_grad():
            imgs = x.cuda()                                            # test.py:48
            o2d, o3d, ord_ = smap_torch.smap_forward(sd, imgs)          # test.py:50
            o3d, ord_ = o3d.cpu(), ord_.cpu()                          # test.py:52-53
            for i in range(B):                                          # test.py:72-134
                hms = o2d[i]
                hms[:15] /= 255
                hms[15:] /= 127
                rdepth = ord_[i][0]
                bodies = ref.connect(hms, rdepth, 2, True)              # test.py:115 (unmodified extension)
                if len(bodies) > 0:
                    p2, p3, rd = lift_numpy.lift(bodies.numpy(), o3d[i].numpy(), rdepth.numpy(), scale)
                    persons += len(p2)
        return persons

    step(xs[0])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    for s in range(steps):
        n += step(xs[1 + s])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("reference GPU path: %d frames in %.2f s -> %.2f frames/s (batch %d, %d persons, cudnn tf32=%s, %s)" %
          (steps * B, dt, steps * B / dt, B, n, torch.backends.cudnn.allow_tf32, torch.cuda.get_device_name(0)))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
