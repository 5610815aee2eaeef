"""Association kernels at BASELINE config 4 scale (B=64 crowded scenes, 15 persons): time and DRAM bytes per kernel.
  python tools/assoc_bw.py                      # CUDA-event timing of extract (nms+paf) and connect (nms+paf+group)
  ncu --profile-from-start off -k regex:'nms|paf|group' --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      python tools/assoc_bw.py --ncu            # one bracketed connect call
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from smap_b200 import synth
from smap_b200.engine import Engine

B = int(os.environ.get("ASSOC_B", "64"))
hms, rd, dd = synth.make_batch(7, B, persons=15)
eng = Engine(0, max_batch=B, in_h=512, in_w=832)
hms_d, rd_d = torch.from_numpy(hms).cuda(), torch.from_numpy(rd).cuda()
for _ in range(3):
    bodies, counts = eng.connect(hms_d, rd_d)
torch.cuda.synchronize()
if "--ncu" in sys.argv:
    torch.cuda.profiler.start()
    eng.connect(hms_d, rd_d)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    sys.exit(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.zero_()  # > L2: the heat-maps come from HBM, as they do after a backbone forward of this batch size
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


t_ext = timed(lambda: eng.extract(hms_d))
t_con = timed(lambda: eng.connect(hms_d, rd_d))
hm_bytes = B * 43 * 128 * 208 * 4
print("B=%d persons/frame=%s" % (B, counts[:4].tolist()))
print("extract (nms+paf, dense -1 fill of the score tensor): %.3f ms" % t_ext)
print("connect (nms+paf+group): %.3f ms  -> %.0f frames/s; heat-map bytes %.1f MB -> %.0f GB/s algorithmic" %
      (t_con, B / t_con * 1e3, hm_bytes / 1e6, hm_bytes / t_con / 1e6))
eng.close()
