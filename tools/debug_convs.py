"""Run every distinct conv shape of the model (B=8, 832x512) through conv_test several times: determinism + accuracy."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from smap_b200 import schema
from smap_b200.engine import Engine
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
# spatial size per unit
def level_of(name):
    if "downsample.layer" in name:
        li = int(name.split("downsample.layer")[1][0]); blk = int(name.split(".")[3])
        return li, blk
    return None
shapes = set()
for (name, cin, cout, k, s, p, relu, enc) in schema.unit_specs():
    if name == "top.conv": continue
    if "downsample.layer" in name:
        li, blk = level_of(name)
        hin = 128 >> (li - 1); win = 208 >> (li - 1)
        if blk == 0 and li > 1 and ("conv_bn_relu1" in name or name.endswith(".downsample")):
            hin, win = hin * 2, win * 2   # input of the strided block is the previous level
        if blk == 0 and li > 1 and "conv_bn_relu2" in name:
            hin, win = hin * 2, win * 2
        res = "conv_bn_relu3" in name
    else:
        ind = int(name.split("upsample.up")[1][0]) - 1
        hin, win = 16 << ind, 26 << ind
        if name.endswith("up_conv"): hin, win = hin // 2, win // 2
        res = False
    shapes.add((hin, win, cin, cout, k, s, relu, res))
eng = Engine(0, max_batch=1, in_h=64, in_w=96)
bad = 0
LOAD = os.environ.get("DEBUG_LOAD")
side = torch.cuda.Stream()
big = torch.randn(64 * 1024 * 1024, device="cuda")
ONLY = os.environ.get("DEBUG_ONLY")
for (H, W, cin, cout, k, s, relu, res) in sorted(shapes):
    if ONLY and ("%dx%d_%d_%d" % (H, W, cin, cout)) != ONLY: continue
    if os.environ.get("DEBUG_RELU"): relu = True
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(B, H, W, cin, generator=g).cuda()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    b = torch.randn(cout, generator=g).cuda()
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, b, stride=s, padding=k // 2)
    r = None
    if res and cout % 32 == 0:
        r = torch.randn(B, ref.shape[2], ref.shape[3], cout, generator=g).cuda()
        ref = ref + r.permute(0, 3, 1, 2)
    if relu: ref = F.relu(ref)
    ref = ref.permute(0, 2, 3, 1)
    quiet = eng.conv_test(x, w, b, res=r, stride=s, relu=relu)
    torch.cuda.synchronize()
    if LOAD:
        with torch.cuda.stream(side):
            for _ in range(40):
                c = big * 1.0001 + 1.0
    ys = [eng.conv_test(x, w, b, res=r, stride=s, relu=relu) for _ in range(4)]
    torch.cuda.synchronize()
    det = all(torch.equal(ys[0], y) for y in ys[1:]) and torch.equal(ys[0], quiet)
    dq = (ys[0] - quiet).abs().max().item() / ref.abs().max().item()
    err = max(((y - ref).abs().max() / ref.abs().max()).item() for y in ys)
    flag = "" if (det and err < 2e-5) else "  <<<<<< BAD"
    if flag: bad += 1
    print("%3dx%-3d cin%-4d cout%-4d k%d s%d relu%d res%d  err %.1e det %s dq %.1e%s" % (H, W, cin, cout, k, s, relu, res, err, det, dq, flag))
print("bad:", bad, "of", len(shapes))
