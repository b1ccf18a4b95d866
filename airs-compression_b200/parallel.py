"""Multi-GPU plumbing: contiguous sharding of independent jobs over ranks and the
end-of-run gather of per-frame sizes and concatenated streams (torch.distributed).

The hot path has no collective: every job (context) is independent (ref cmp.c:228-236),
so rank r simply owns jobs [r*J/W, (r+1)*J/W).  Only sizes and, if asked for, the
streams travel at the end.
"""
import os

import torch
import torch.distributed as dist


def shard_range(n_units, rank, world):
    """Contiguous, balanced [begin, end) of n_units for this rank."""
    base, extra = divmod(n_units, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _counts(n, device, group):
    world = dist.get_world_size(group)
    mine = torch.tensor([n], dtype=torch.int64, device=device)
    out = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return [int(c) for c in out.tolist()]


def _gather_ragged(part, group, out=None):
    """Every rank's 1-D tensor laid out back to back in rank order in ONE preallocated buffer: every rank sends its
    part to every peer and receives every peer's part straight into its final place, all transfers of a rank in one
    group (ncclSend / ncclRecv between ncclGroupStart / End: they run at the same time, so a rank's NVLink ports
    are busy in both directions - no padding to the longest part, no list of temporaries, no concatenation
    afterwards, no serial chain of broadcasts).  `out`: a buffer to reuse (at least the gathered size)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = _counts(part.numel(), part.device, group)
    offsets = [0]
    for c in counts[:-1]:
        offsets.append(offsets[-1] + c)
    total = sum(counts)
    if out is None or out.numel() < total or out.dtype != part.dtype:
        out = torch.empty(total, dtype=part.dtype, device=part.device)
    out = out[:total]
    if os.environ.get("AIRS_GATHER", "sendrecv") == "allgather":
        # The alternative, kept for measurements: ONE all-gather of parts padded to the longest, then every part moves
        # to its place behind the part in front of it (device copies).  On 4 B200s 3.4 GB took 11.8 ms this way and
        # 10.8 ms with the grouped transfers below (290 against 316 GB/s): the default stays the grouped transfers
        longest = max(counts)
        mine = part
        if counts[rank] != longest:
            mine = torch.empty(longest, dtype=part.dtype, device=part.device)
            mine[:counts[rank]].copy_(part)
        padded = torch.empty(world * longest, dtype=part.dtype, device=part.device)
        dist.all_gather_into_tensor(padded, mine, group=group)
        for r in range(world):
            out[offsets[r]:offsets[r] + counts[r]].copy_(padded[r * longest:r * longest + counts[r]])
        return out, counts, offsets
    out[offsets[rank]:offsets[rank] + counts[rank]].copy_(part)
    ops = []
    for step in range(1, world):  # (peers in a rotated order: no rank is everybody's first target)
        to, frm = (rank + step) % world, (rank - step) % world
        if counts[rank]:
            ops.append(dist.P2POp(dist.isend, part, dist.get_global_rank(group, to) if group is not None else to, group))
        if counts[frm]:
            ops.append(dist.P2POp(dist.irecv, out[offsets[frm]:offsets[frm] + counts[frm]],
                                  dist.get_global_rank(group, frm) if group is not None else frm, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out, counts, offsets


def allgather_sizes(sizes, group=None):
    """All ranks' per-frame result arrays, concatenated in rank order (equal lengths not required)."""
    out, counts, _ = _gather_ragged(sizes, group)
    return out, counts


def allgather_streams(stream, group=None, out=None):
    """Variable-length byte streams of all ranks laid out back to back in rank order.

    Returns (gathered uint8 tensor, byte offset of every rank's part).  NCCL moves the bytes over
    NVLink; with gloo (CPU tensors) the same code is what the world_size-2 tests run."""
    out, _, offsets = _gather_ragged(stream, group, out)
    return out, offsets
