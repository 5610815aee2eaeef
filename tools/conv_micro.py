"""GPU microbench of single conv layers through smapb_conv_test (tools only; not a bench value)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smap_b200.engine import Engine

SHAPES = {
    # name: (B, H, W, Cin, Cout, k, stride, relu, res)
    "l1_c3": (8, 128, 208, 64, 256, 1, 1, True, True),
    "l1_c3_nores": (8, 128, 208, 64, 256, 1, 1, True, False),
    "l1_c1": (8, 128, 208, 256, 64, 1, 1, True, False),
    "l1_c2": (8, 128, 208, 64, 64, 3, 1, True, False),
    "up4_1x1": (8, 128, 208, 256, 256, 1, 1, True, False),
    "l2_c3": (8, 64, 104, 128, 512, 1, 1, True, True),
    "l3_c3": (8, 32, 52, 256, 1024, 1, 1, True, True),
    "l3_c2": (8, 32, 52, 256, 256, 3, 1, True, False),
    "l3_c1": (8, 32, 52, 1024, 256, 1, 1, True, False),
    "l4_c2": (8, 16, 26, 512, 512, 3, 1, True, False),
}
# CONV_MICRO_BATCH=n overrides the batch of every shape (fewer tiles -> fewer SMs busy: separates per-SM from chip-wide limits)
names = [a for a in sys.argv[1:] if not a.startswith("-")] or list(SHAPES)
if os.environ.get("CONV_MICRO_BATCH"):
    SHAPES = {k: (int(os.environ["CONV_MICRO_BATCH"]),) + v[1:] for k, v in SHAPES.items()}
eng = Engine(0, max_batch=1, in_h=64, in_w=96)
for n in names:
    B, H, W, Cin, Cout, k, s, relu, res = SHAPES[n]
    x = torch.randn(B, H, W, Cin, device="cuda")
    w = torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, device="cuda")
    Ho, Wo = H // s, W // s
    r = torch.randn(B, Ho, Wo, Cout, device="cuda") if res else None
    y, ms = eng.conv_test(x, w, b, res=r, stride=s, relu=relu, time_it=True)
    flops = 2.0 * B * Ho * Wo * Cout * Cin * k * k
    byts = 4.0 * B * (H * W * Cin * (1 if k == 1 and s == 1 else 1) + Ho * Wo * Cout * (2 if res else 1))
    print("%-12s %7.3f ms  %7.1f TF/s(algo)  %6.0f GB/s(min traffic)" % (n, ms, flops / ms * 1e-9, byts / ms * 1e-6))
