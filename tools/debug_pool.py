import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import lift_numpy
from smap_b200 import schema
from smap_b200.engine import RECORD_BYTES, EnginePool, Engine, scale_row, records_to_numpy
sd = schema.make_state_dict(0, "identity")
pool = EnginePool(2, 0, max_batch=2, in_h=512, in_w=832); pool.load_state_dict(sd)
sc = lift_numpy.default_scale(1920, 1080)
scales = torch.from_numpy(np.stack([scale_row(sc)] * 2)).pin_memory()
xs = [schema.make_input(2, 512, 832, seed=40 + i).pin_memory() for i in range(6)]
# reference: each engine alone, synchronous, twice (determinism of a lone engine)
refs = []
for k, e in enumerate(pool.engines):
    a = [e.infer_host(x, scales).tobytes() for x in xs]
    b = [e.infer_host(x, scales).tobytes() for x in xs]
    print("engine", k, "alone deterministic:", a == b)
    refs.append(a)
print("engines agree with each other:", refs[0] == refs[1])
# backbone tensors equality between engines
x = xs[0].cuda()
o0 = pool.engines[0].forward(x); o1 = pool.engines[1].forward(x)
print("backbone bit-equal across engines:", [bool(torch.equal(a, b)) for a, b in zip(o0, o1)])
for rnd in range(3):
    outs = [torch.zeros(2, RECORD_BYTES, dtype=torch.uint8).pin_memory() for _ in xs]
    tickets = [pool.submit(x, scales, o) for x, o in zip(xs, outs)]
    for t in tickets:
        if t in pool._tickets: pool.result(t)
    ok = [outs[t].numpy().tobytes() == refs[t % 2][t] for t in range(len(xs))]
    print("round", rnd, "pipelined == sync on same engine:", ok)
