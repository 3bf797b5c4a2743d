"""World-size-2 gloo test of the N>1 host logic: shard-by-proof + one all-gather of proof bytes.

Each rank proves its shard with the host-simulator build (CPU), the gathered buffer must equal
the single-process result for the whole batch byte for byte (proofs are independent, so
sharding must not change any proof).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, N, SEC = 4, 5, 8


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _prove(lib, lo, hi):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import common
    from zkp_ecdsa_b200 import synth
    P, _ = common.make_params(lib, seed=9, sec_level=SEC)
    wl = synth.Workload(B=B, N=N, seed=9)
    tape = synth.random_tape(B, lib.prove_tape_len(N, SEC), seed=10)
    for name in ('msg_hash', 'sig', 'pk', 'which'):
        setattr(wl, name, np.ascontiguousarray(getattr(wl, name)[lo:hi]))
    wl.B = hi - lo
    proofs, plen, status = common.run_prove(lib, P, wl, np.ascontiguousarray(tape[lo:hi]), SEC)
    assert (status == 0).all()
    return proofs, plen


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import __graft_entry__ as g
    from zkp_ecdsa_b200 import sharding
    from zkp_ecdsa_b200.capi import ZkaLib
    os.environ['ZKA_TOM_W'] = '10'
    os.environ['ZKA_P256_HW'] = '8'   # small tables: the host simulator builds them on one CPU core
    lib = ZkaLib(g.HOSTSIM)
    lo, hi = sharding.shard_range(B, rank, world)
    per = (B + world - 1) // world
    proofs, plen = _prove(lib, lo, hi)
    stride = proofs.shape[1]
    lp = torch.zeros((per, stride), dtype=torch.uint8)
    ll = torch.zeros(per, dtype=torch.int32)
    lp[:hi - lo] = torch.from_numpy(proofs)
    ll[:hi - lo] = torch.from_numpy(plen.astype(np.int32))
    allp, alll = sharding.all_gather_proofs(lp, ll, world, rank, per)
    # trimmed variant: same rows, cut at the longest proof of the whole job (rounded up to 16 bytes)
    trimmed, tlen = sharding.all_gather_proofs(lp, ll, world, rank, per, trim=True)
    w = trimmed.shape[1]
    assert w <= stride and w >= int(alll.max()) and w - int(alll.max()) < 16
    assert torch.equal(trimmed, allp[:, :w]) and torch.equal(tlen, alll)
    # overlapped, compacted gather (bench.py's N > 1 path): two groups per rank, pack -> all-gather -> unpack
    stride16 = (stride + 15) & ~15
    lp16 = torch.zeros((per, stride16), dtype=torch.uint8)
    lp16[:, :stride] = lp
    pg = sharding.ProofGather(lib, world, rank, per, stride16, N, SEC, torch.device('cpu'), groups=2)
    pg.begin()
    for (b0, b1) in pg.ranges:
        pg.submit(lp16, ll, b0, b1)
    pg.finish()
    info = pg.check(lp16, ll)
    assert info['checksums_match_all_ranks'] and info['own_rows_roundtrip'] and info['max_fill'] <= 1.0
    for gi, (b0, b1) in enumerate(pg.ranges):
        for r in range(world):
            rows, lens = pg.unpack(gi, r)
            for k in range(b1 - b0):
                j = r * per + b0 + k
                assert int(lens[k]) == int(alll[j]) and torch.equal(rows[k, :int(lens[k])], allp[j, :int(lens[k])])
    # one prove call per rank, finished chunks gathered while later ones are proved (zka_set_progress; bench.py's default
    # N > 1 path): one-proof chunks here, so that there is more than one
    import common
    from zkp_ecdsa_b200 import synth
    lib.set_option('chunk', 1)
    lib.set_option('host_chunk', 1)
    off = lib.chunk_schedule(hi - lo, host_buffers=True)
    assert off[0] == 0 and off[-1] == hi - lo and len(off) == hi - lo + 1
    P, _ = common.make_params(lib, seed=9, sec_level=SEC)
    wl = synth.Workload(B=B, N=N, seed=9)
    tape = np.ascontiguousarray(synth.random_tape(B, lib.prove_tape_len(N, SEC), seed=10)[lo:hi])
    out = np.zeros((hi - lo, stride16), np.uint8)
    olen = np.zeros(hi - lo, np.uint32)
    ost = np.zeros(hi - lo, np.int32)
    tp, tl = torch.from_numpy(out), torch.from_numpy(olen.view(np.int32))

    def prove_again():
        lib.prove_batch(P, hi - lo, np.ascontiguousarray(wl.msg_hash[lo:hi]), np.ascontiguousarray(wl.sig[lo:hi]),
                        np.ascontiguousarray(wl.pk[lo:hi]), np.ascontiguousarray(wl.which[lo:hi]), wl.ring, N, tape, tape.shape[1],
                        out, stride16, olen, ost)
    pg2 = sharding.ProofGather(lib, world, rank, hi - lo, stride16, N, SEC, torch.device('cpu'), ranges=list(zip(off[:-1], off[1:])))
    pg2.prove_overlapped(prove_again, tp, tl)
    assert not ost.any()
    info2 = pg2.check(tp, tl)
    assert info2['checksums_match_all_ranks'] and info2['own_rows_roundtrip']
    for k in range(hi - lo):
        assert int(olen[k]) == int(plen[k]) and bytes(out[k, :olen[k]]) == bytes(proofs[k, :plen[k]])
    if rank == 0:
        q.put((allp.numpy(), alll.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_batch():
    from zkp_ecdsa_b200 import sharding
    for total in (1, 7, 8, 65536):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_two_rank_gather_equals_single_process():
    import __graft_entry__ as g
    g.build_hostsim()
    from zkp_ecdsa_b200.capi import ZkaLib
    os.environ['ZKA_TOM_W'] = '10'
    os.environ['ZKA_P256_HW'] = '8'
    ref_proofs, ref_len = _prove(ZkaLib(g.HOSTSIM), 0, B)
    os.environ.pop('ZKA_TOM_W', None)
    os.environ.pop('ZKA_P256_HW', None)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    allp, alll = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert list(alll[:B]) == list(ref_len)
    for b in range(B):
        assert allp[b, :alll[b]].tobytes() == ref_proofs[b, :ref_len[b]].tobytes()
