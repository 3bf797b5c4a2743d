"""Parity tests proper: the sm_100a library through the C ABI vs the oracle (run on the B200 box)."""
import numpy as np
import pytest

import common
from zkp_ecdsa_b200 import synth

pytestmark = pytest.mark.gpu


def test_field_ops(gpu_engine):
    common.check_field_ops(gpu_engine.lib, count=200)


def test_hash80(gpu_engine):
    common.check_hash(gpu_engine.lib)


def test_p256_mul(gpu_engine):
    common.check_p256_mul(gpu_engine.lib, count=12)


def test_params_and_commit(gpu_engine):
    P, po = common.make_params(gpu_engine.lib, seed=5)
    common.check_tom_commit(gpu_engine.lib, P, po, count=12)
    gpu_engine.lib.params_destroy(P)


def test_prove_bit_exact_small_ring(gpu_engine):
    common.check_prove_parity(gpu_engine.lib, B=2, N=6, seed=3)


def test_prove_bit_exact_ring_256(gpu_engine):
    common.check_prove_parity(gpu_engine.lib, B=2, N=256, seed=8)


def test_prove_bit_exact_sec_level_16(gpu_engine):
    common.check_prove_parity(gpu_engine.lib, B=5, N=17, seed=4, sec_level=16)


def test_prove_batch_properties_config2(gpu_engine):
    """BASELINE config 2 shape (batch 1024, ring 8): size-independent properties."""
    L = gpu_engine.lib
    P, po = common.make_params(L, seed=21)
    B, N = 1024, 8
    wl = synth.Workload(B=B, N=N, seed=21)
    tape = synth.random_tape(B, L.prove_tape_len(N), seed=22)
    proofs, plen, status = common.run_prove(L, P, wl, tape)
    assert (status == 0).all()
    from oracle import flat
    # lengths follow the layout formula; zero bits are binomial(80, 1/2)
    zs = (plen.astype(np.int64) - flat.proof_len(0, 3)) // (flat.REP0_LEN - flat.REP1_LEN)
    assert ((plen == [flat.proof_len(int(z), 3) for z in zs])).all()
    assert 30 < zs.mean() < 50
    # determinism: same inputs -> same bytes; different tape -> different proof
    proofs2, plen2, _ = common.run_prove(L, P, wl, tape)
    assert (plen == plen2).all() and (proofs == proofs2).all()
    # spot-check three proofs bit-exactly against the oracle
    for b in (0, 511, 1023):
        pr, _ = common.oracle_proof(po, wl, tape, b)
        assert proofs[b, :plen[b]].tobytes() == flat.ser_proof(pr)
    L.params_destroy(P)


def test_error_statuses(gpu_engine):
    L = gpu_engine.lib
    P, po = common.make_params(L, seed=31)
    wl = synth.Workload(B=4, N=8, seed=31)
    tape = synth.random_tape(4, L.prove_tape_len(8), seed=32)
    wl.pk[1, 40] ^= 1                      # not on the curve -> 'invalid public key'
    wl.which[2] = 8                        # outside the ring
    tape[3, 32 * 3:32 * 3 + 4] = 255       # alpha_0 >= n -> tape range
    _, _, status = common.run_prove(L, P, wl, tape)
    assert list(status) == [0, 1, 6, 5]
    L.params_destroy(P)


def test_verify_decisions_match_oracle(gpu_engine):
    common.check_verify_parity(gpu_engine.lib, N=6, seed=3, tampers=32)


def test_verify_ring_256(gpu_engine):
    common.check_verify_parity(gpu_engine.lib, N=256, seed=12, tampers=8)


def test_prove_verify_roundtrip_batch(gpu_engine):
    """encode -> verify round trip at batch 512, ring 1024 (size-independent property)."""
    from zkp_ecdsa_b200 import verify_tape as VT
    L = gpu_engine.lib
    P, po = common.make_params(L, seed=41)
    B, N = 512, 1024
    wl = synth.Workload(B=B, N=N, seed=41)
    tape = synth.random_tape(B, L.prove_tape_len(N), seed=42)
    proofs, plen, status = common.run_prove(L, P, wl, tape)
    assert (status == 0).all()
    vt = VT.random_verify_tape(B, L.verify_tape_len(N), N, seed=43)
    ok, st = common.run_verify(L, P, wl.msg_hash, wl.ring, proofs, plen, vt)
    assert (st == 0).all() and (ok == 1).all()
    # swap two messages: exactly those two verifications must fail
    msg = wl.msg_hash.copy()
    msg[[3, 4]] = msg[[4, 3]]
    ok, st = common.run_verify(L, P, msg, wl.ring, proofs, plen, vt)
    assert (st == 0).all() and list(np.nonzero(ok == 0)[0]) == [3, 4]
    L.params_destroy(P)


def test_edge_cases_on_gpu(gpu_engine):
    import test_edge_cases as E
    L = gpu_engine.lib
    E.test_ring_of_two(L)
    E.test_ragged_ring_signer_last(L)
    E.test_zero_message_hash(L)
    E.test_zero_r_and_zero_s(L)
    E.test_key_to_int(L)


def test_alternate_code_paths_on_gpu():
    """Paths the default engine does not take: ragged / small table windows, and a ring above 1024
    entries (Groth-Kohlweiss block sums)."""
    import os
    from zkp_ecdsa_b200 import api
    os.environ.update(ZKA_TOM_W='14', ZKA_P256_HW='11')
    try:
        eng = api.Engine(device=0)
    finally:
        for k in ('ZKA_TOM_W', 'ZKA_P256_HW'):
            os.environ.pop(k, None)
    try:
        common.check_prove_parity(eng.lib, B=3, N=6, sec_level=80, seed=41)
        common.check_prove_parity(eng.lib, B=1, N=2100, sec_level=20, seed=42)
        common.check_verify_parity(eng.lib, N=2100, sec_level=20, seed=43, tampers=4)
    finally:
        eng.close()
