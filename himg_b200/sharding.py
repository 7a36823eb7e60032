"""Multi-GPU sharding of an image batch: images are independent, so each rank codes a
contiguous range of the batch with NO collective on the coding path.  The single exchange step of
the job is an all-gather of the per-image bitstream sizes, from which every rank derives the
global offset table (exclusive scan) -- a few KiB, latency bound.  NCCL on GPUs, gloo on CPU."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int):
    """Contiguous range [first, first+count) of rank `rank`; ranges differ by at most one image."""
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def gather_sizes(sizes: torch.Tensor, world: int | None = None, pad_to: int = 16):
    """sizes: this rank's per-image byte sizes (int32/int64 tensor on the rank's device).

    Returns (all_sizes int64 [total], offsets int64 [total+1]) on the same device; offsets are the
    positions of the images inside the concatenated stream of the whole job, each stream padded
    to `pad_to` bytes.  Ranks may hold different image counts."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    local = sizes.to(torch.int64)
    if world == 1:
        allsz = local
    else:
        count = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
        counts = [torch.zeros_like(count) for _ in range(world)]
        dist.all_gather(counts, count)
        counts = [int(c.item()) for c in counts]
        m = max(counts)
        padded = torch.zeros(m, dtype=torch.int64, device=local.device)
        padded[: local.numel()] = local
        gathered = [torch.zeros_like(padded) for _ in range(world)]
        dist.all_gather(gathered, padded)
        allsz = torch.cat([g[:c] for g, c in zip(gathered, counts)])
    padded_sizes = (allsz + (pad_to - 1)) // pad_to * pad_to
    offsets = torch.zeros(allsz.numel() + 1, dtype=torch.int64, device=allsz.device)
    torch.cumsum(padded_sizes, 0, out=offsets[1:])
    return allsz, offsets
