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


# ----------------------------------------------------------------------------------------------------
# Overlapped, compacted all-gather (round 2).  A rank's batch is proved in `groups` sub-batches; as soon as
# one is finished its proofs are packed to their true lengths (zka_proofs_pack, 16-byte aligned starts) and
# their all-gather is queued on a communication stream, so the transfer of group g overlaps the proving of
# group g + 1.  No host synchronisation sizes anything: every group's send block has a fixed capacity
# (rows x the length of a proof with `Z_BOUND` zero challenge bits — the mean is SecLevel / 2 — or the worst
# case for small groups); the true block lengths travel in offsets[] and are checked after the step.
Z_BOUND_SIGMAS = 12


def _proof_len(z: int, n: int, reps: int) -> int:
    gk = 1 + 4 * n * 67 + (3 * n + 1) * 33
    return 264 + z * 3596 + (reps - z) * 330 + gk


def group_capacity(rows: int, ring_size: int, sec_level: int) -> int:
    """Bytes reserved for the packed proofs of `rows` proofs: mean + 12 sigma of the sum of their lengths
    (zero bits are Binomial(sec_level, 1/2) per proof), never more than the worst case; 256-byte aligned."""
    n = max(1, (ring_size - 1).bit_length())
    worst = rows * ((_proof_len(sec_level, n, sec_level) + 15) & ~15)
    mean = rows * (_proof_len(0, n, sec_level) + 15) + rows * (sec_level / 2) * (3596 - 330)
    sigma = (rows * sec_level / 4) ** 0.5 * (3596 - 330)
    return (min(worst, int(mean + Z_BOUND_SIGMAS * sigma) + 4096) + 255) & ~255


def group_ranges(B: int, groups: int):
    """Row ranges of the sub-batches: every group about half the size of the one before (2/3 + 1/3, 4/7 + 2/7 + 1/7,
    ...), multiples of 32 rows.  Only the LAST group's all-gather is exposed after the last kernel, so it should be
    small; few large groups keep the prover's chunks large (every zka_prove_batch call ends with a full
    synchronisation of its lanes: four equal groups cost 25 % of the 2-GPU throughput, profiles/README.md)."""
    groups = max(1, min(groups, B))
    weights = [2 ** (groups - 1 - i) for i in range(groups)]
    tot = sum(weights)
    unit = 32 if B >= 64 * groups else 1
    out, b0 = [], 0
    for i, w in enumerate(weights):
        n = B - b0 if i == groups - 1 else min(B - b0 - (groups - 1 - i), max(unit, (B * w // tot + unit - 1) // unit * unit))
        if n <= 0:
            break
        out.append((b0, b0 + n))
        b0 += n
    return out


class ProofGather:
    def __init__(self, lib, world: int, rank: int, B: int, stride: int, ring_size: int, sec_level: int, device, groups: int = 2,
                 ranges=None):
        import torch
        self.lib, self.world, self.rank, self.B, self.stride = lib, world, rank, B, stride
        # `ranges`: the chunk schedule of ONE prove call (lib.chunk_schedule) for prove_overlapped(); else sub-batches
        # proved by separate calls (submit() after each)
        self.ranges = [tuple(r) for r in ranges] if ranges is not None else group_ranges(B, groups)
        G = len(self.ranges)
        self.nrows = [b1 - b0 for (b0, b1) in self.ranges]
        self.caps = [group_capacity(n, ring_size, sec_level) for n in self.nrows]
        self.dev = device
        self.send = [torch.zeros(self.caps[g], dtype=torch.uint8, device=device) for g in range(G)]
        self.recv = [torch.zeros(world * self.caps[g], dtype=torch.uint8, device=device) for g in range(G)]
        self.lens = [torch.zeros(self.nrows[g], dtype=torch.int32, device=device) for g in range(G)]
        self.lens_all = [torch.zeros(world * self.nrows[g], dtype=torch.int32, device=device) for g in range(G)]
        self.offs = [torch.zeros(self.nrows[g] + 1, dtype=torch.int64, device=device) for g in range(G)]
        self.cuda = device.type == 'cuda'
        self.comm = torch.cuda.Stream(device=device) if self.cuda else None
        self.works = []
        self.g = 0
        self.exposed_ms = []

    def describe(self) -> str:
        return (f'{len(self.ranges)} groups per rank of {self.nrows} proofs; per group one NCCL all-gather of the proofs packed to '
                f'their true lengths (capacities {self.caps} B per rank) + one of their lengths, queued on a communication stream '
                'while the next group is proved; no host synchronisation sizes the transfer')

    def begin(self):
        self.works = []
        self.g = 0

    def submit(self, proofs, plen, b0: int, b1: int):
        """Queue the all-gather of rows [b0, b1) (already complete on the device)."""
        import torch
        import torch.distributed as dist
        g = self.g
        self.g += 1
        rows = b1 - b0
        assert (b0, b1) == self.ranges[g] and rows == self.nrows[g]
        ctx = torch.cuda.stream(self.comm) if self.cuda else _Null()
        with ctx:
            self.lens[g].copy_(plen[b0:b1])
            self.lib.proofs_pack(rows, proofs[b0:].data_ptr(), self.stride, self.lens[g].data_ptr(), self.send[g].data_ptr(),
                                 self.caps[g], self.offs[g].data_ptr(), self.comm.cuda_stream if self.cuda else 0)
            if self.world > 1:
                self.works.append(dist.all_gather_into_tensor(self.lens_all[g], self.lens[g], async_op=True))
                self.works.append(dist.all_gather_into_tensor(self.recv[g], self.send[g], async_op=True))
            else:
                self.lens_all[g].copy_(self.lens[g])
                self.recv[g].copy_(self.send[g])

    def prove_overlapped(self, prove_fn, proofs, plen, poll_s: float = 20e-6):
        """ONE prove call over the whole batch (prove_fn blocks; it runs on a helper thread, ctypes releases the GIL) while
        this thread watches the library's progress flags and queues the all-gather of every finished chunk — in chunk
        order, which is the same on all ranks, so the collectives match.  Only the chunks that finish last are exposed."""
        import threading
        import time
        import numpy as np
        flags = np.zeros(len(self.ranges), np.uint32)
        self.lib.set_progress(flags)
        err = []

        def work():
            try:
                prove_fn()
            except Exception as e:      # noqa: BLE001
                err.append(e)
        th = threading.Thread(target=work)
        self.begin()
        th.start()
        try:
            for k, (b0, b1) in enumerate(self.ranges):
                while flags[k] == 0:
                    if err or not th.is_alive() and flags[k] == 0:
                        break
                    time.sleep(poll_s)
                if err or flags[k] == 0:
                    break
                self.submit(proofs, plen, b0, b1)
        finally:
            th.join()
            self.lib.set_progress(None)
        if err:
            raise err[0]
        assert self.g == len(self.ranges), 'the prove call ended before every chunk reported completion'
        self.finish()

    def finish(self):
        import time
        import torch
        t0 = time.perf_counter()
        for w in self.works:
            w.wait()
        if self.cuda:
            self.comm.synchronize()
            torch.cuda.current_stream().synchronize()
        self.exposed_ms.append((time.perf_counter() - t0) * 1e3)

    def totals(self, g: int):
        """Packed block length of every rank for group g (from the gathered lengths)."""
        import torch
        l = self.lens_all[g].view(self.world, self.nrows[g]).to(torch.int64)
        return ((l + 15) & ~15).sum(dim=1)

    def unpack(self, g: int, r: int):
        """Rows [rows, stride] of rank r's group g from the gathered block (zka_proofs_unpack)."""
        import torch
        n, cap = self.nrows[g], self.caps[g]
        out = torch.zeros((n, self.stride), dtype=torch.uint8, device=self.dev)
        offs = torch.zeros(n + 1, dtype=torch.int64, device=self.dev)
        lens = self.lens_all[g][r * n:(r + 1) * n].contiguous()
        if self.cuda:
            # `out` / `offs` were zero-filled on torch's current stream; the library runs on its own non-blocking stream
            torch.cuda.current_stream().synchronize()
        self.lib.proofs_unpack(n, self.recv[g][r * cap:].data_ptr(), cap, lens.data_ptr(), out.data_ptr(),
                               self.stride, offs.data_ptr(), 0)
        return out, lens

    def check(self, proofs, plen):
        """After a step: no block overflowed its capacity, every rank's block arrived intact on this rank
        (checksum exchange), and this rank's own rows survive pack -> gather -> unpack bit for bit."""
        import torch
        import torch.distributed as dist
        G = len(self.ranges)
        info = {'groups': G, 'group_rows': self.nrows, 'capacity_bytes': self.caps,
                'exposed_ms_per_step': sum(self.exposed_ms) / max(1, len(self.exposed_ms))}
        local = torch.zeros(G, dtype=torch.int64, device=self.dev)
        seen = torch.zeros((G, self.world), dtype=torch.int64, device=self.dev)
        maxfill = 0.0
        for g in range(G):
            tot = self.totals(g)
            cap = self.caps[g]
            assert int(tot.max().item()) <= cap, 'packed block exceeded its capacity'
            maxfill = max(maxfill, float(tot.max().item()) / cap)
            for r in range(self.world):
                n = int(tot[r].item())
                blk = self.recv[g][r * cap:r * cap + n]
                seen[g, r] = torch.sum(blk.view(torch.int32), dtype=torch.int64) if n else 0
            n = int(tot[self.rank].item())
            local[g] = torch.sum(self.send[g][:n].view(torch.int32), dtype=torch.int64) if n else 0
        allc = torch.zeros(self.world * G, dtype=torch.int64, device=self.dev)
        if self.world > 1:
            dist.all_gather_into_tensor(allc, local)
        else:
            allc.copy_(local)
        ok = bool(torch.equal(allc.view(self.world, G).t().contiguous(), seen))
        # own rows back from the gathered block
        same = True
        for g, (b0, b1) in enumerate(self.ranges):
            rows, lens = self.unpack(g, self.rank)
            k = b1 - b0
            col = torch.arange(self.stride, device=self.dev).unsqueeze(0)
            valid = col < plen[b0:b1].unsqueeze(1)
            same = same and bool(torch.equal(lens[:k], plen[b0:b1])) and bool(((rows[:k] == proofs[b0:b1]) | ~valid).all().item())
            del col, valid, rows
        info.update({'checksums_match_all_ranks': ok, 'own_rows_roundtrip': same, 'max_fill': maxfill,
                     'bytes_received_per_step': int(sum(int(self.totals(g).sum().item()) for g in range(G)))})
        assert ok and same, info
        return info


class _Null:
    def __enter__(self): return self
    def __exit__(self, *a): return False
