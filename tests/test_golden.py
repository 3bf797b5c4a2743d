"""Committed golden fixtures (tests/golden/zkattest_v{1,2}.npz, made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import common
from oracle import flat
from oracle import zkattest as OZ
from oracle.big import Tape
from zkp_ecdsa_b200 import verify_tape as VT

G = {}
for _f in ('zkattest_v1.npz', 'zkattest_v2.npz'):
    _z = np.load(os.path.join(os.path.dirname(__file__), 'golden', _f))
    G.update({k: _z[k] for k in _z.files})


def _case(tag):
    B, N, sec, seed = (int(v) for v in G[f'{tag}_meta'])
    return B, N, sec, {k[len(tag) + 1:]: G[k] for k in G if k.startswith(tag + '_')}


def _check_lib(L, tag):
    B, N, sec, d = _case(tag)
    hn, hp = L.params_generate(d['params_rnd'].tobytes())
    assert hn == d['h_nist'].tobytes() and hp == d['h_proof'].tobytes()
    P = L.params_create(hn, hp, sec)
    ps = L.proof_max_len(N, sec)
    proofs = np.zeros((B, ps), np.uint8)
    plen = np.zeros(B, np.uint32)
    status = np.zeros(B, np.int32)
    L.prove_batch(P, B, d['msg_hash'], d['sig'], d['pk'], d['which'], d['ring'], N, d['tape'], d['tape'].shape[1],
                  proofs, ps, plen, status)
    assert (status == 0).all() and (plen == d['proof_len']).all()
    for b in range(B):
        assert proofs[b, :plen[b]].tobytes() == d['proofs'][b, :plen[b]].tobytes(), (tag, b)
    ok, st = common.run_verify(L, P, d['msg_hash'], d['ring'], np.ascontiguousarray(d['proofs']), d['proof_len'], d['vtape'])
    assert (st == 0).all() and list(ok) == list(d['verdict'])
    L.params_destroy(P)


def test_oracle_reproduces_golden_a():
    B, N, sec, d = _case('a')
    po = OZ.generate_params_list(Tape(d['params_rnd'].tobytes()), sec)
    ring = [int.from_bytes(d['ring'][i].tobytes(), 'big') for i in range(N)]
    pr = OZ.prove_signature_list(po, d['msg_hash'][0].tobytes(), d['sig'][0].tobytes(), d['pk'][0].tobytes(),
                                 int(d['which'][0]), ring, Tape(d['tape'][0].tobytes()))
    assert flat.ser_proof(pr) == d['proofs'][0, :d['proof_len'][0]].tobytes()
    assert OZ.verify_signature_list(po, d['msg_hash'][0].tobytes(), ring, pr,
                                    Tape(VT.oracle_stream(d['vtape'][0].tobytes(), N, sec))) == bool(d['verdict'][0])


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd'])
def test_hostsim_matches_golden(hostsim, tag):
    _check_lib(hostsim, tag)


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd'])
def test_gpu_matches_golden(gpu_engine, tag):
    _check_lib(gpu_engine.lib, tag)
