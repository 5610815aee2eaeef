"""Multi-GPU plumbing: frames are independent end-to-end, so a batch is sharded in contiguous blocks of frames per
rank (the reference's own split rule, lib/utils/dataloader.py:80-85) and the only exchange is ONE all-gather of the
fixed-stride per-image skeleton records (smapb_record, 92,464 B/frame) per batch.  NCCL over NVLink for CUDA tensors;
the same code path runs on gloo/CPU tensors for the host-logic tests."""
import torch
import torch.distributed as dist

from ._lib import RECORD_BYTES


def shard_range(n_frames, rank, world):
    """Contiguous block [lo, hi) of frames owned by `rank`; the first (n_frames % world) ranks get one extra."""
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_frames, world):
    return [shard_range(n_frames, r, world)[1] - shard_range(n_frames, r, world)[0] for r in range(world)]


def allgather_records(rec, n_frames=None, group=None):
    """rec: uint8 [B_local, RECORD_BYTES] (cuda -> NCCL, cpu -> gloo).  Returns uint8 [n_frames, RECORD_BYTES] in
    global frame order on every rank.  With equal shards this is a single all_gather_into_tensor; ragged shards are
    padded to the largest shard for the collective and compacted afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rec
    world = dist.get_world_size(group)
    assert rec.dtype == torch.uint8 and rec.shape[1] == RECORD_BYTES
    if n_frames is None:
        n_frames = rec.shape[0] * world
    sizes = shard_sizes(n_frames, world)
    bmax = max(sizes)
    send = rec
    if rec.shape[0] != bmax:
        send = torch.zeros(bmax, RECORD_BYTES, dtype=torch.uint8, device=rec.device)
        send[:rec.shape[0]] = rec
    out = torch.empty(world * bmax, RECORD_BYTES, dtype=torch.uint8, device=rec.device)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    if all(s == bmax for s in sizes):
        return out
    return torch.cat([out[r * bmax:r * bmax + sizes[r]] for r in range(world)], 0)


def sync_tile_table(group=None, src=0):
    """Make every rank use rank `src`'s tile shapes for geometries the committed table does not cover: call after the
    engines of rank `src` have built their plans (first forward) and before the other ranks build theirs - or simply
    before any forward on all ranks when the committed table covers the workload (then this is a no-op in effect)."""
    from . import engine

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    box = [engine.get_tile_table() if dist.get_rank(group) == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    engine._lib.load().smapb_set_tile_table(box[0].encode())
