"""Shard-by-proof partitioning and the final all-gather of proof bytes.

Proofs are independent (zkpAttestList.ts:104-145 touches no shared state), so a batch of B
proofs is cut into contiguous index ranges, one per rank; params, tables and the ring are
replicated.  The only collective on the path is ONE all-gather of the serialized proofs at the
end (BASELINE.json north_star).  Backend-agnostic: NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of proof indices owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_proofs(local_proofs, local_len, world: int, rank: int, per_rank: int, trim: bool = False):
    """One all-gather of the proof rows plus (a tiny one of) their lengths.

    local_proofs: [per_rank, stride] uint8, local_len: [per_rank] int32 (rows beyond this rank's
    share are padding and have length 0).  Returns ([world*per_rank, width], [world*per_rank]).
    width = stride, or with `trim` the longest proof of the whole job rounded up to 16 bytes: rows
    are stride-padded for the worst case (all 80 challenge bits zero), real proofs are ~40 % shorter,
    so trimming takes that padding off the wire.
    """
    import torch
    import torch.distributed as dist
    stride = local_proofs.shape[1]
    lens = torch.empty(world * per_rank, dtype=local_len.dtype, device=local_len.device)
    if world == 1:
        lens.copy_(local_len)
    else:
        dist.all_gather_into_tensor(lens, local_len.contiguous())
    width = stride
    if trim:
        width = min(stride, (int(lens.max().item()) + 15) & ~15)
    send = local_proofs if width == stride else local_proofs[:, :width]
    out = torch.empty((world * per_rank, width), dtype=local_proofs.dtype, device=local_proofs.device)
    if world == 1:
        out.copy_(send)
        return out, lens
    dist.all_gather_into_tensor(out, send.contiguous())
    return out, lens
