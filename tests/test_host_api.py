"""Host mirror of the reference API (zkp_ecdsa_b200/api.py) on the host simulator: default randomness,
security levels other than 80, error rows."""
import numpy as np
import pytest

import common
from zkp_ecdsa_b200 import api, synth


def _engine(lib):
    eng = api.Engine.__new__(api.Engine)
    eng.lib = lib
    return eng


def _one(seed, N=4):
    wl = synth.Workload(B=1, N=N, seed=seed)
    return wl, (wl.msg_hash[0].tobytes(), wl.sig[0].tobytes(), wl.pk[0].tobytes(), int(wl.which[0]), wl.ring_ints())


def test_default_tape_is_os_csprng(hostsim, monkeypatch):
    """tape=None must draw from the OS CSPRNG like crypto.getRandomValues (big.ts:175): numpy's generators
    are made unusable for the duration of the calls, two proofs of one statement differ, both verify."""
    eng = _engine(hostsim)
    params = eng.generate_params_list(20, rnd=synth.params_rnd(3))
    wl, (msg, sig, pk, which, keys) = _one(61)

    def boom(*a, **k):
        raise AssertionError('numpy.random reached from the default-randomness path')
    for name in ('Generator', 'PCG64', 'default_rng', 'RandomState', 'seed', 'randint', 'bytes'):
        monkeypatch.setattr(np.random, name, boom)
    p1 = eng.prove_signature_list(params, msg, sig, pk, which, keys)
    p2 = eng.prove_signature_list(params, msg, sig, pk, which, keys)
    assert p1.data != p2.data
    assert eng.verify_signature_list(params, msg, keys, p1) is True
    assert eng.verify_signature_list(params, msg, keys, p2) is True
    monkeypatch.undo()
    params.close()


def test_os_tape_respects_draw_moduli():
    t = api.synth_os_tape(3, 32 * 400, 80).reshape(3, 400, 32)
    for r in range(3):
        for k in range(400):
            assert int.from_bytes(t[r, k].tobytes(), 'big') < synth.draw_modulus(k, 80)
    # a forced out-of-range candidate is redrawn
    import os
    real = os.urandom
    calls = {'n': 0}

    def fake(n):
        calls['n'] += 1
        return b'\xff' * n if calls['n'] == 1 else real(n)
    os.urandom = fake
    try:
        t = api.synth_os_tape(1, 64, 80).reshape(2, 32)
    finally:
        os.urandom = real
    assert int.from_bytes(t[0].tobytes(), 'big') < synth.P256_N and int.from_bytes(t[1].tobytes(), 'big') < synth.P256_P
    assert calls['n'] >= 3


@pytest.mark.parametrize('sec', [20, 33])
def test_verify_default_tape_other_sec_levels(hostsim, sec):
    """SecLevel in [20, 80): the generateIndices bytes must be drawn for `sec` repetitions."""
    eng = _engine(hostsim)
    params = eng.generate_params_list(sec, rnd=synth.params_rnd(4))
    wl, (msg, sig, pk, which, keys) = _one(62)
    proof = eng.prove_signature_list(params, msg, sig, pk, which, keys)
    for _ in range(3):
        assert eng.verify_signature_list(params, msg, keys, proof) is True
    bad = bytearray(proof.data)
    bad[300] ^= 1
    with pytest.raises(api.ZkaProofError):
        eng.verify_signature_list(params, msg, keys, api.SignatureProofList(bytes(bad)))
    params.close()


def test_wild_index_and_error_rows(hostsim):
    """`which` far outside the ring: status BAD_INDEX, no out-of-bounds access, and a failed proof leaves
    the library as an all-zero row of length 0."""
    L = hostsim
    P, po = common.make_params(L, 71, 16)
    wl = synth.Workload(B=3, N=8, seed=71)
    wl.which[1] = 0xFFFFFFFF
    wl.pk[2, 10] ^= 4
    tape = synth.random_tape(3, L.prove_tape_len(8, 16), seed=72)
    proofs, plen, status = common.run_prove(L, P, wl, tape, 16)
    assert list(status) == [0, 6, 1]
    assert plen[0] > 0 and plen[1] == 0 and plen[2] == 0
    assert not proofs[1].any() and not proofs[2].any()
    pr, _ = common.oracle_proof(po, wl, tape, 0)
    from oracle import flat
    assert proofs[0, :plen[0]].tobytes() == flat.ser_proof(pr)
    L.params_destroy(P)


def test_engine_proof_group_selection(hostsim_war):
    """api.Engine picks the library by ProofGroup name (instances.ts:58-69: only the groups the reference ships are
    valid); the war256 engine runs the reference-shaped host functions end to end."""
    with pytest.raises(ValueError):
        api.Engine(proof_group='curve25519')
    eng = api.Engine.__new__(api.Engine)
    eng.lib = hostsim_war
    eng.proof_group = hostsim_war.group
    params = eng.generate_params_list(20, rnd=synth.params_rnd(6))     # verifySignatureList samples 20 repetitions
    assert params.proof_group == 'war256' and len(params.h_proof) == 65
    wl, (msg, sig, pk, which, keys) = _one(63)
    proof = eng.prove_signature_list(params, msg, sig, pk, which, keys)
    assert eng.verify_signature_list(params, msg, keys, proof) is True
    bad = bytearray(proof.data)
    bad[-1] ^= 1
    assert eng.verify_signature_list(params, msg, keys, api.SignatureProofList(bytes(bad))) is False
    params.close()
