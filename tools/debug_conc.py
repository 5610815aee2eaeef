import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smap_b200 import schema
from smap_b200.engine import Engine
sd = schema.make_state_dict(0, "identity")
E = [Engine(0, max_batch=2, in_h=512, in_w=832) for _ in range(2)]
for e in E: e.load_state_dict(sd)
xs = [schema.make_input(2, 512, 832, seed=50 + i).cuda() for i in range(2)]
ref = [[t.clone() for t in E[k].forward(xs[k])] for k in range(2)]
torch.cuda.synchronize()
bad = [0, 0, 0]
for rnd in range(6):
    outs = [E[k].forward(xs[k]) for k in range(2)]   # both in flight
    outs2 = [E[k].forward(xs[k]) for k in range(2)]
    torch.cuda.synchronize()
    for k in range(2):
        for j in range(3):
            if not torch.equal(outs[k][j], ref[k][j]) or not torch.equal(outs2[k][j], ref[k][j]):
                bad[j] += 1
print("mismatching concurrent forwards (hm, detd, rootd):", bad)
for rnd in range(4):
    outs = [E[k].forward(xs[k]) for k in range(2)]
    outs2 = [E[k].forward(xs[k]) for k in range(2)]
    torch.cuda.synchronize()
    for name, o in (("a", outs), ("b", outs2)):
        for k in range(2):
            d = (o[k][0] - ref[k][0]).abs()
            if d.max().item() > 0:
                idx = torch.nonzero(d > 0)
                print("  rnd", rnd, name, "eng", k, "hm ndiff", idx.shape[0], "max %.3g" % d.max().item(), "rel %.2g" % (d.max() / ref[k][0].abs().max()).item(),
                      "n", idx[:, 0].unique().tolist(), "c", idx[:, 1].unique().numel(), "y", idx[:, 2].min().item(), idx[:, 2].max().item(), "x", idx[:, 3].min().item(), idx[:, 3].max().item())
