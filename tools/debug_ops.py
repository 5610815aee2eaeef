"""Which plan op first produces different bits when the GPU is shared with other work? (debug)"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smap_b200 import schema
from smap_b200.engine import Engine
sd = schema.make_state_dict(0, "identity")
E = Engine(0, max_batch=2, in_h=512, in_w=832); E.load_state_dict(sd)
x = schema.make_input(2, 512, 832, seed=50).cuda()
lib = E.lib
lib.smapb_debug_checksums.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
def sums():
    arr = (ctypes.c_ulonglong * 400)(); desc = ctypes.create_string_buffer(400 * 160)
    n = lib.smapb_debug_checksums(E._h, 2, arr, 400, desc, 160)
    return [arr[i] for i in range(n)], [desc.raw[i * 160:(i + 1) * 160].split(b"\0")[0].decode() for i in range(n)]
E.forward(x); torch.cuda.synchronize()
q1, d = sums()
E.forward(x); torch.cuda.synchronize()
q2, _ = sums()
print("quiet runs identical:", q1 == q2, "ops", len(q1))
side = torch.cuda.Stream(); big = torch.randn(64 * 1024 * 1024, device="cuda")
for rnd in range(3):
    with torch.cuda.stream(side):
        for _ in range(30): c = big * 1.0001 + 1.0
    E.forward(x); torch.cuda.synchronize()
    l, _ = sums()
    diff = [i for i in range(len(l)) if l[i] != q1[i]]
    print("round", rnd, "differing ops:", len(diff), "first:", diff[:6])
    for i in diff[:4]: print("    ", i, d[i])
