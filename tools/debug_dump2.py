"""What do the wrong values of op 17 look like? (debug)"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smap_b200 import schema
from smap_b200.engine import Engine
OP = 17; RES = int(os.environ.get("DUMP_RES", "14"))
sd = schema.make_state_dict(0, "identity")
E = Engine(0, max_batch=2, in_h=512, in_w=832); E.load_state_dict(sd)
x = schema.make_input(2, 512, 832, seed=50).cuda()
lib = E.lib
lib.smapb_debug_dump.restype = ctypes.c_longlong
lib.smapb_debug_dump.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int]
C = 512; M = 2 * 64 * 104
def dump(i):
    buf = np.zeros(64 << 20, np.uint16)
    n = lib.smapb_debug_dump(E._h, 2, i, buf.ctypes.data, buf.nbytes, 0)
    assert n > 0, n
    return buf[: n // 2].copy()
def val(w):
    t = torch.from_numpy(w.view(np.int16).reshape(2, M, C)).view(torch.bfloat16).float()
    return (t[0] + t[1]).numpy()
E.forward(x); torch.cuda.synchronize()
q = dump(OP); res = dump(RES); print("res words", res.size, "out words", q.size)
vq = val(q); vres = val(res) if res.size == q.size else None
side = torch.cuda.Stream(); big = torch.randn(64 * 1024 * 1024, device="cuda")
np.set_printoptions(precision=3, linewidth=220, suppress=True)
for rnd in range(6):
    with torch.cuda.stream(side):
        for _ in range(30): c = big * 1.0001 + 1.0
    E.forward(x); torch.cuda.synchronize()
    l = dump(OP); ne = (l != q).reshape(2, M, C).any(0)
    if not ne.any(): print("round", rnd, "clean"); continue
    vl = val(l)
    rows = np.nonzero(ne.any(1))[0]; cols = np.nonzero(ne.any(0))[0]
    print("round", rnd, "rows", rows.min(), rows.max(), "n", len(rows), "cols", cols.min(), cols.max())
    for tile in np.unique(rows // 128):
        rr = rows[rows // 128 == tile]; ch = np.unique(np.nonzero(ne[rr].any(0))[0] // 32)
        print("  tile", tile, "rows in tile", (rr % 128).min(), (rr % 128).max(), "chunks", ch)
        for cc in ch:
            r = rr[0]; sl = slice(cc * 32, cc * 32 + 32)
            print("   row", r, "chunk", cc); print("    q  ", vq[r, sl]); print("    l  ", vl[r, sl])
            if vres is not None: print("    res", vres[r, sl]); print("    l-q", (vl - vq)[r, sl])
            # stale-staging hypothesis: wrong row equals the quiet output somewhere else (same row-in-tile)?
            target = l.reshape(2, M, C)[0][r, sl]
            cand = q.reshape(2, M, C)[0].reshape(M // 128, 128, C // 32, 32)[:, r % 128]  # [tiles, chunks, 32]
            hit = np.argwhere((cand == target).all(-1))
            print("    equals quiet output at (tile,chunk):", hit[:6].tolist())
            if vres is not None:
                # residual-mixup hypothesis: pre = q - res (where q > 0); l == relu(pre + res_other)?
                pre = vq[r, sl] - vres[r, sl]; pos = (vq[r, sl] > 0) & (vl[r, sl] > 0)
                want = (vl[r, sl] - pre)
                rc = vres.reshape(M // 128, 128, C // 32, 32)[:, r % 128]
                err = np.abs(rc - want)[..., pos].max(-1) if pos.any() else None
                if err is not None:
                    best = np.unravel_index(np.argmin(err), err.shape); print("    residual best match (tile,chunk)", best, "err", err[best], "npos", int(pos.sum()))
    break
