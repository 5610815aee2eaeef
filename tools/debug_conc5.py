import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smap_b200 import schema
from smap_b200.engine import Engine
sd = schema.make_state_dict(0, "identity")
E = Engine(0, max_batch=2, in_h=512, in_w=832); E.load_state_dict(sd)
x = schema.make_input(2, 512, 832, seed=50).cuda()
ref = [t.clone() for t in E.forward(x)]
torch.cuda.synchronize()
side = torch.cuda.Stream()
big = torch.randn(64 * 1024 * 1024, device="cuda")
for rnd in range(3):
    with torch.cuda.stream(side):
        for _ in range(30):
            c = big * 1.0001 + 1.0
    o = E.forward(x)
    torch.cuda.synchronize()
    for name, a, b in zip(("hm", "dd", "rd"), o, ref):
        d = (a - b).abs()
        idx = torch.nonzero(d > 0)
        if idx.shape[0]:
            print(rnd, name, "ndiff", idx.shape[0], "of", d.numel(), "max %.3g rel %.2g" % (d.max().item(), (d.max() / b.abs().max()).item()),
                  "n", idx[:, 0].unique().tolist(), "y", idx[:, 2].min().item(), idx[:, 2].max().item(), "x", idx[:, 3].min().item(), idx[:, 3].max().item())
        else:
            print(rnd, name, "equal")
