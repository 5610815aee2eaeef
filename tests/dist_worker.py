"""torchrun worker for tests/test_multigpu_gpu.py (BASELINE.json config 3): N frames are sharded over
the ranks in contiguous blocks (smap_b200.dist.shard_range = lib/utils/dataloader.py:80-85), every rank runs the fused path
on its block and ONE ncclAllGather (smapb_infer_device_gather: on the compute stream, inside the CUDA graph) exchanges the
skeleton records.  Every rank then recomputes ALL N frames by itself - the 1-GPU result - and the gathered bytes must be
identical.  Also covered: the stand-alone smapb_allgather_records over torch.distributed's own communicator
(ProcessGroupNCCL._comm_ptr), the host variant (smapb_submit_host_gather) and graph replay (several rounds).
    python -m torch.distributed.run --nproc-per-node 2 tests/dist_worker.py [frames_total]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


def main():
    from smap_b200 import dist as sdist
    from smap_b200 import schema
    from smap_b200.engine import RECORD_BYTES, Engine, scale_row

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 8 * world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    assert total % world == 0
    B = total // world
    lo, hi = sdist.shard_range(total, rank, world)
    assert hi - lo == B
    sd = schema.make_state_dict(0, "identity")
    st = torch.cuda.Stream(dev)
    eng = Engine(local, max_batch=B, in_h=512, in_w=832, stream=st)
    eng.load_state_dict(sd)
    eng.init_comm()
    sdist.sync_tile_table()
    # the same `total` frames on every rank (seeded), each rank takes its block
    frames = torch.cat([schema.make_input(1, 512, 832, seed=1000 + i) for i in range(total)], 0)
    sc = dict(scale=832 / 1920, img_width=1920, img_height=1080, net_width=832, net_height=512, f_x=1920.0, f_y=1920.0,
              cx=960.0, cy=540.0)
    scales = torch.from_numpy(np.stack([scale_row(sc)] * B)).to(dev)
    mine = frames[lo:hi].to(dev)
    torch.cuda.synchronize()
    gathered = []
    for rnd in range(4):  # 2 eager runs, graph capture, graph replay
        out = torch.zeros(total, RECORD_BYTES, dtype=torch.uint8, device=dev)
        eng.infer_device(mine, scales, out=out, gather=True)
        st.synchronize()
        gathered.append(out.cpu())
    for g in gathered[1:]:
        assert torch.equal(g, gathered[0]), "gathered records changed between eager run and graph replay"
    # decoupled form: exchange on the gather stream, valid after gather_sync; several calls in flight (double-buffered records)
    outs = [torch.zeros(total, RECORD_BYTES, dtype=torch.uint8, device=dev) for _ in range(5)]
    for o in outs:
        eng.infer_device(mine, scales, out=o, gather=True, defer=True)
    eng.gather_sync()
    st.synchronize()
    for o in outs:
        assert torch.equal(o.cpu(), gathered[0]), "deferred exchange differs from the stream-ordered one"
    # host variant: H2D -> path -> device all-gather -> ONE D2H of the gathered records
    host_out = torch.zeros(total, RECORD_BYTES, dtype=torch.uint8).pin_memory()
    eng.submit_host(0, frames[lo:hi].contiguous().pin_memory(), scales.cpu().pin_memory(), host_out, gather=True)
    eng.wait(0)
    assert torch.equal(host_out, gathered[0]), "host gather differs from device gather"
    # 1-GPU result for all frames, computed locally in blocks of B with a second handle (no communicator)
    solo = Engine(local, max_batch=B, in_h=512, in_w=832)
    solo.load_state_dict(sd)
    ref = torch.cat([solo.infer_device(frames[k:k + B].to(dev), scales).cpu() for k in range(0, total, B)], 0)
    assert torch.equal(gathered[0], ref), "rank %d: gathered records differ from the 1-GPU result" % rank
    # and with a different batch split on the 1-GPU side (tile boundaries move, bits must not)
    if B % 2 == 0:
        half = Engine(local, max_batch=B // 2, in_h=512, in_w=832)
        half.load_state_dict(sd)
        ref2 = torch.cat([half.infer_device(frames[k:k + B // 2].to(dev), scales[:B // 2]).cpu()
                          for k in range(0, total, B // 2)], 0)
        assert torch.equal(ref2, ref), "batch split changed the records"
        half.close()
    # the exchange step alone over torch.distributed's own communicator
    e2 = Engine(local, max_batch=B, in_h=512, in_w=832)
    e2.attach_torch_comm()
    own = ref[lo:hi].to(dev)
    allr = e2.allgather(own)
    torch.cuda.synchronize()
    assert torch.equal(allr.cpu(), ref), "smapb_allgather_records over torch's communicator"
    # all ranks hold the same bytes
    digest = torch.tensor([int.from_bytes(__import__("hashlib").sha256(gathered[0].numpy().tobytes()).digest()[:7], "little")],
                          device=dev)
    alld = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(alld, digest)
    assert all(int(d) == int(digest) for d in alld)
    persons = int(gathered[0].numpy().view(np.uint8).reshape(total, RECORD_BYTES)[:, -8:-4].copy().view(np.int32).sum())
    if rank == 0:
        print("MULTIGPU OK world=%d frames=%d persons=%d bytes=%d sha=%x" % (world, total, persons, gathered[0].numel(), int(digest)))
    dist.barrier()
    for e in (eng, solo, e2):
        e.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
