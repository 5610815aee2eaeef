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
def cmp(o, k): return all(torch.equal(a, b) for a, b in zip(o, ref[k]))
# A: single engine, 4 forwards back-to-back, no sync
res = []
for rnd in range(6):
    o = [E[0].forward(xs[0]) for _ in range(4)]
    torch.cuda.synchronize()
    res.append([cmp(x, 0) for x in o])
print("A single engine x4 back-to-back:", res)
# B: two engines, one forward each
res = []
for rnd in range(8):
    o = [E[k].forward(xs[k]) for k in range(2)]
    torch.cuda.synchronize()
    res.append([cmp(o[k], k) for k in range(2)])
print("B two engines, one forward each:", res)
# C: two engines, 2 forwards each interleaved
res = []
for rnd in range(6):
    o = [(k, E[k].forward(xs[k])) for _ in range(2) for k in range(2)]
    torch.cuda.synchronize()
    res.append([cmp(x, k) for k, x in o])
print("C interleaved 2x2:", res)
