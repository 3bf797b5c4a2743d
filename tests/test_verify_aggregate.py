"""Chunk-wide aggregate check of the verifier (zk_verify_agg.cuh): a valid chunk is accepted by ONE wide-window MSM per
group; a chunk holding a wrong proof falls back to the per-proof evaluation and gives exactly its verdicts."""
import numpy as np
import pytest

import common
from zkp_ecdsa_b200 import synth
from zkp_ecdsa_b200 import verify_tape as VT


def _batch(L, B, N, seed, sec_level=80):
    P, po = common.make_params(L, seed, sec_level)
    wl = synth.Workload(B=B, N=N, seed=seed)
    tape = synth.random_tape(B, L.prove_tape_len(N, sec_level), seed=seed + 100)
    proofs, plen, status = common.run_prove(L, P, wl, tape, sec_level)
    assert (status == 0).all()
    vt = VT.random_verify_tape(B, L.verify_tape_len(N, sec_level), N, sec_level, seed=seed + 7)
    return P, wl, proofs, plen, vt


def check_aggregate(L, B=5, N=6, seed=31, cs=(0, 4, 7, 11)):
    P, wl, proofs, plen, vt = _batch(L, B, N, seed)
    try:
        for c in cs:
            if c:
                L.set_option('agg_c', c)
            p0, f0 = L.stat('agg_pass'), L.stat('agg_fail')
            ok, st = common.run_verify(L, P, wl.msg_hash, wl.ring, proofs, plen, vt)
            assert list(ok) == [1] * B and not st.any(), (c, ok, st)
            assert L.stat('agg_pass') > p0 and L.stat('agg_fail') == f0, c      # decided by the aggregate
        # one wrong proof (a flipped bit in the last GK response scalar, which every verification reads): the chunk
        # goes to the per-proof path
        bad = proofs.copy()
        bad[2, plen[2] - 1] ^= 1
        p0, f0 = L.stat('agg_pass'), L.stat('agg_fail')
        ok, st = common.run_verify(L, P, wl.msg_hash, wl.ring, bad, plen, vt)
        assert L.stat('agg_fail') > f0 and L.stat('agg_pass') == p0
        L.set_option('agg', 1)                                                   # aggregate off: the reference verdicts
        ok2, st2 = common.run_verify(L, P, wl.msg_hash, wl.ring, bad, plen, vt)
        okv, stv = common.run_verify(L, P, wl.msg_hash, wl.ring, proofs, plen, vt)
        assert L.stat('agg_fail') == f0 + (L.stat('agg_fail') - f0) and list(okv) == [1] * B and not stv.any()
        L.set_option('agg', 2)
        assert (ok == ok2).all() and (st == st2).all()
        assert ok[2] == 0 and list(np.delete(ok, 2)) == [1] * (B - 1), ok
        # a malformed row (truncated) keeps its own status and sends the chunk to the per-proof path too
        plen2 = plen.copy()
        plen2[1] -= 3
        ok3, st3 = common.run_verify(L, P, wl.msg_hash, wl.ring, proofs, plen2, vt)
        assert ok3[1] == 0 and st3[1] != 0 and list(np.delete(ok3, 1)) == [1] * (B - 1)
    finally:
        L.set_option('agg', 2)
        L.set_option('agg_c', 6)
        L.params_destroy(P)


def test_aggregate_check_hostsim(hostsim):
    check_aggregate(hostsim)


@pytest.mark.gpu
def test_aggregate_check_on_gpu(gpu_engine):
    check_aggregate(gpu_engine.lib, B=40, N=17, seed=33, cs=(0, 9, 13))
