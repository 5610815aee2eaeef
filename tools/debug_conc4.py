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
def cmp(o): return all(torch.equal(a, b) for a, b in zip(o, ref))
side = torch.cuda.Stream()
a = torch.randn(4096, 4096, device="cuda"); b = torch.randn(4096, 4096, device="cuda")
big = torch.randn(64 * 1024 * 1024, device="cuda")
for mode in ("matmul", "elementwise", "memcpy"):
    res = []
    for rnd in range(6):
        with torch.cuda.stream(side):
            for _ in range(30):
                if mode == "matmul": c = a @ b
                elif mode == "elementwise": c = big * 1.0001 + 1.0
                else: c = big.clone()
        o = [E.forward(x) for _ in range(3)]
        torch.cuda.synchronize()
        res.append([cmp(t) for t in o])
    print(mode, res)
