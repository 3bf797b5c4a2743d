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


def all_gather_proofs(local_proofs, local_len, world: int, rank: int, per_rank: int):
    """One all-gather of the stride-padded proof rows plus their lengths.

    local_proofs: [per_rank, stride] uint8, local_len: [per_rank] int32 (rows beyond this rank's
    share are padding and have length 0).  Returns ([world*per_rank, stride], [world*per_rank]).
    """
    import torch
    import torch.distributed as dist
    stride = local_proofs.shape[1]
    out = torch.empty((world * per_rank, stride), dtype=local_proofs.dtype, device=local_proofs.device)
    lens = torch.empty(world * per_rank, dtype=local_len.dtype, device=local_len.device)
    if world == 1:
        out.copy_(local_proofs)
        lens.copy_(local_len)
        return out, lens
    dist.all_gather_into_tensor(out, local_proofs.contiguous())
    dist.all_gather_into_tensor(lens, local_len.contiguous())
    return out, lens
