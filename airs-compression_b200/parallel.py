"""Multi-GPU plumbing: contiguous sharding of independent jobs over ranks and the
end-of-run gather of per-frame sizes and concatenated streams (torch.distributed).

The hot path has no collective: every job (context) is independent (ref cmp.c:228-236),
so rank r simply owns jobs [r*J/W, (r+1)*J/W).  Only sizes and, if asked for, the
streams travel at the end.
"""
import torch
import torch.distributed as dist


def shard_range(n_units, rank, world):
    """Contiguous, balanced [begin, end) of n_units for this rank."""
    base, extra = divmod(n_units, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allgather_sizes(sizes, group=None):
    """All ranks' per-frame result arrays, concatenated in rank order (equal lengths not required)."""
    world = dist.get_world_size(group)
    n = torch.tensor([sizes.numel()], dtype=torch.int64, device=sizes.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    pad = max(counts)
    buf = torch.zeros(pad, dtype=sizes.dtype, device=sizes.device)
    buf[:sizes.numel()] = sizes
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)]), counts


def allgather_streams(stream, group=None):
    """Variable-length byte streams of all ranks laid out back to back in rank order.

    Returns (gathered uint8 tensor, byte offset of every rank's part).  NCCL moves the bytes over
    NVLink; with gloo (CPU tensors) the same code is what the world_size-2 tests run."""
    world = dist.get_world_size(group)
    n = torch.tensor([stream.numel()], dtype=torch.int64, device=stream.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    pad = max(max(counts), 1)
    buf = torch.zeros(pad, dtype=torch.uint8, device=stream.device)
    buf[:stream.numel()] = stream
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    offsets = [0]
    for c in counts[:-1]:
        offsets.append(offsets[-1] + c)
    return torch.cat([o[:c] for o, c in zip(out, counts)]), offsets
