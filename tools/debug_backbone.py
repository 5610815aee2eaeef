import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import smap_torch
from smap_b200 import schema
from smap_b200.engine import Engine
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
sd = schema.make_state_dict(0, "identity")
sdc = {k: v.cuda() for k, v in sd.items()}
for (H, W, B) in [(512, 832, 1)]:
    t0 = time.time()
    eng = Engine(0, max_batch=B, in_h=H, in_w=W)
    eng.load_state_dict(sd)
    x = schema.make_input(B, H, W, seed=1).cuda()
    t1 = time.time()
    ref = smap_torch.smap_forward(sdc, x)
    torch.cuda.synchronize(); t2 = time.time()
    outs = [eng.forward(x) for _ in range(3)]
    torch.cuda.synchronize(); t3 = time.time()
    det = all(torch.equal(a, b) for a, b in zip(outs[0], outs[1])) and all(torch.equal(a, b) for a, b in zip(outs[0], outs[2]))
    errs = [((a - b).abs().max() / b.abs().max()).item() for a, b in zip(outs[0], ref)]
    # where is the error?
    d = (outs[0][0] - ref[0]).abs()
    idx = torch.nonzero(d > 1e-3 * ref[0].abs().max())
    print(H, W, B, "errs", ["%.2e" % e for e in errs], "deterministic", det, "bad px", idx.shape[0],
          "setup %.1fs oracle %.1fs ours %.1fs" % (t1 - t0, t2 - t1, t3 - t2))
    if idx.shape[0]:
        print("   bad idx sample", idx[:5].tolist(), "y-range", idx[:, 2].min().item(), idx[:, 2].max().item(), "x-range", idx[:, 3].min().item(), idx[:, 3].max().item(), "n", idx[:,0].unique().tolist())
    eng.close()
