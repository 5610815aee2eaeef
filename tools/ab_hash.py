"""Digest of the whole path's outputs on fixed inputs (tools only): run once per library build (SMAPB_LIB=...) and diff the
lines - a kernel change that claims to keep the arithmetic must reproduce every digest."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from smap_b200 import schema
from smap_b200.engine import Engine, scale_row


def dig(t):
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]


def main():
    B, H, W = 2, 512, 832
    for bn in ("random",):
        eng = Engine(0, max_batch=B, in_h=H, in_w=W)
        eng.load_state_dict(schema.make_state_dict(0, bn))
        x = schema.make_input(B, H, W, seed=5).cuda()
        hm, dd, rd = eng.forward(x)
        torch.cuda.synchronize()
        print("forward %-8s hm2d %s detd %s rootd %s" % (bn, dig(hm), dig(dd), dig(rd)))
        sc = dict(scale=W / 1920, img_width=1920, img_height=1080, net_width=W, net_height=H, f_x=1920.0, f_y=1920.0,
                  cx=960.0, cy=540.0)
        scales = torch.from_numpy(np.stack([scale_row(sc)] * B)).cuda()
        for flip in (False, True):
            rec = eng.infer_device(x, scales, do_flip=flip)
            torch.cuda.synchronize()
            print("records %-8s flip=%d %s" % (bn, int(flip), dig(rec)))
        eng.close()
    # single layers through smapb_conv_test: residual, post adds, strided, bf16 mode
    eng = Engine(0, max_batch=1, in_h=64, in_w=96)
    g = torch.Generator(device="cuda").manual_seed(7)
    for (Bc, Hc, Wc, Cin, Cout, k, s, res, posts) in [(2, 64, 104, 64, 256, 1, 1, True, 0), (2, 64, 104, 256, 256, 1, 1, True, 2),
                                                       (2, 32, 52, 256, 256, 3, 1, False, 0), (2, 32, 52, 128, 128, 3, 2, False, 0),
                                                       (1, 16, 26, 512, 64, 1, 1, False, 1)]:
        xx = torch.randn(Bc, Hc, Wc, Cin, device="cuda", generator=g)
        ww = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5
        bb = torch.randn(Cout, device="cuda", generator=g)
        Ho, Wo = (Hc + 2 * (k // 2) - k) // s + 1, (Wc + 2 * (k // 2) - k) // s + 1
        rr = torch.randn(Bc, Ho, Wo, Cout, device="cuda", generator=g) if res else None
        pp = [torch.randn(Bc, Ho, Wo, Cout, device="cuda", generator=g) for _ in range(posts)]
        for prec in ("bf16x3", "bf16"):
            y = eng.conv_test(xx, ww, bb, res=rr, stride=s, relu=True, precision=prec, post1=pp[0] if posts > 0 else None,
                              post2=pp[1] if posts > 1 else None)
            y = y[0] if isinstance(y, tuple) else y
            print("conv %dx%dx%d %d->%d k%d s%d res%d post%d %s %s" % (Bc, Hc, Wc, Cin, Cout, k, s, int(res), posts, prec, dig(y)))
    eng.close()


if __name__ == "__main__":
    main()
