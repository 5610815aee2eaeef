"""torchrun tool: latency of the path's one exchange step in isolation - smapb_allgather_records (handle-owned NCCL communicator)
vs torch.distributed.all_gather_into_tensor - for B records per rank, and what transport NCCL picked (run with NCCL_DEBUG=INFO).
    python -m torch.distributed.run --nproc-per-node 2 tools/allgather_latency.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from smap_b200.engine import RECORD_BYTES, Engine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = 8
st = torch.cuda.Stream(dev)
eng = Engine(local, max_batch=B, in_h=512, in_w=832, stream=st)
eng.init_comm()
rec = torch.zeros(B, RECORD_BYTES, dtype=torch.uint8, device=dev)
out = torch.zeros(world * B, RECORD_BYTES, dtype=torch.uint8, device=dev)
for name, fn in (("smapb_allgather_records", lambda: eng.allgather(rec, out)),
                 ("torch all_gather_into_tensor", lambda: dist.all_gather_into_tensor(out, rec))):
    s = st if name.startswith("smapb") else torch.cuda.current_stream()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(200):
        fn()
    e1.record(s)
    torch.cuda.synchronize()
    if rank == 0:
        print("%-32s world=%d payload=%d B/rank: %.1f us per call" % (name, world, B * RECORD_BYTES, e0.elapsed_time(e1) * 1e3 / 200), flush=True)
p2p = torch.cuda.can_device_access_peer(local, (local + 1) % world) if world > 1 else None
if rank == 0:
    print("can_device_access_peer:", p2p, flush=True)
dist.barrier()
eng.close()
dist.destroy_process_group()
