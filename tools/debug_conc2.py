import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import lift_numpy
from smap_b200 import schema
from smap_b200.engine import Engine, scale_row, records_to_numpy
sd = schema.make_state_dict(0, "identity")
E = [Engine(0, max_batch=2, in_h=512, in_w=832) for _ in range(2)]
for e in E: e.load_state_dict(sd)
xs = [schema.make_input(2, 512, 832, seed=50 + i).cuda() for i in range(2)]
sc = lift_numpy.default_scale(1920, 1080)
scales = torch.from_numpy(np.stack([scale_row(sc)] * 2)).cuda()
# stage-wise references (sequential)
hm = []; ref_conn = []; ref_ext = []; ref_rec = []
for k in range(2):
    h, dd, rd = E[k].forward(xs[k]); h = E[k].merge_scale(h.clone(), None, True)
    hm.append((h, dd, rd))
    b, c = E[k].connect(h, rd[:, 0]); ref_conn.append((b.clone(), c.clone()))
    p, s = E[k].extract(h); ref_ext.append((p.clone(), s.clone()))
    ref_rec.append(E[k].infer_device(xs[k], scales).clone())
torch.cuda.synchronize()
bad = dict(extract_peaks=0, extract_scores=0, connect=0, infer=0)
for rnd in range(8):
    ext = [E[k].extract(hm[k][0]) for k in range(2)]
    con = [E[k].connect(hm[k][0], hm[k][2][:, 0]) for k in range(2)]
    rec = [E[k].infer_device(xs[k], scales) for k in range(2)]
    torch.cuda.synchronize()
    for k in range(2):
        bad["extract_peaks"] += int(not torch.equal(ext[k][0], ref_ext[k][0]))
        bad["extract_scores"] += int(not torch.equal(ext[k][1], ref_ext[k][1]))
        bad["connect"] += int(not (torch.equal(con[k][0], ref_conn[k][0]) and torch.equal(con[k][1], ref_conn[k][1])))
        bad["infer"] += int(not torch.equal(rec[k], ref_rec[k]))
print(bad)
