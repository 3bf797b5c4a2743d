"""ProofGroup = war256 (/root/reference/src/curves/instances.ts:34-41): the war256 build of the library
(libzkattest_war256.so, same sources with -DZKA_PG_WAR256, same C ABI with 65-byte points / 32-byte scalars) against the
oracle run with that proof group — field, commitments, whole proofs byte for byte, verifier verdicts."""
import numpy as np
import pytest

import common


def test_war_field_ops(hostsim_war):
    common.check_field_ops(hostsim_war, count=12)


def test_war_params_and_commit(hostsim_war):
    P, po = common.make_params(hostsim_war, seed=5, sec_level=16)
    assert po.ProofGroup.g.group.name == 'war256'
    common.check_tom_commit(hostsim_war, P, po)
    hostsim_war.params_destroy(P)


def test_war_prove_bit_exact(hostsim_war):
    common.check_prove_parity(hostsim_war, B=2, N=5, seed=51, sec_level=16)


def test_war_prove_bit_exact_sec_level_80(hostsim_war):
    common.check_prove_parity(hostsim_war, B=1, N=8, seed=52, sec_level=80)


def test_war_verify_decisions_match_oracle(hostsim_war):
    common.check_verify_parity(hostsim_war, N=6, seed=53, tampers=16, sec_level=20)


def test_war_prove_few_keys_and_aggregate(hostsim_war):
    common.check_prove_few_keys(hostsim_war, B=8, N=5, signers=1, seed=54, sec_level=16, spots=(0, 7))
    import test_verify_aggregate as tva
    tva.check_aggregate(hostsim_war, B=4, N=6, seed=55, cs=(0, 7))


@pytest.mark.gpu
def test_war_on_gpu(gpu_engine_war):
    """The sm_100a war256 library through the C ABI: field, commitments, whole proofs (SecLevel 80 and 16, a ragged
    ring, ring 256), verdicts incl. tampered proofs, the aggregate check and a prove -> verify round trip of 256 proofs."""
    L = gpu_engine_war.lib
    assert (L.group, L.wp, L.ws) == ('war256', 65, 32)
    common.check_field_ops(L, count=50)
    P, po = common.make_params(L, seed=5)
    common.check_tom_commit(L, P, po)
    L.params_destroy(P)
    common.check_prove_parity(L, B=2, N=6, seed=61)
    common.check_prove_parity(L, B=3, N=17, seed=62, sec_level=16)
    common.check_prove_parity(L, B=1, N=256, seed=63)
    common.check_verify_parity(L, N=6, seed=64, tampers=24)
    common.check_prove_few_keys(L, B=256, N=64, signers=2, seed=65, sec_level=16, spots=(0, 255))
    import test_verify_aggregate as tva
    tva.check_aggregate(L, B=40, N=17, seed=66, cs=(0, 9, 13))
    import test_subproofs as ts
    ts.check_verify_exp(L, sec=16, K=7, with_q=True, seed=67)
    ts.check_verify_membership(L, list(range(100, 100 + 37)), 20, seed=68)
    for kind in ('equality', 'mult', 'pointadd'):
        ts.check_verify_small(L, kind, seed=69, tampers=4)
        ts.check_prove_small(L, kind, seed=70)
    ts.check_prove_exp(L, sec=12, with_q=False, seed=77, B=3)


def test_war_subproofs(hostsim_war):
    """The stand-alone sub-proof entry points of the war256 build: verifyExp (with / without Q), verifyMembership,
    verify / prove Equality, Mult, PointAdd, proveExp, proveMembership — against the oracle run on war256."""
    import test_subproofs as ts
    L = hostsim_war
    ts.check_verify_exp(L, sec=12, K=12, with_q=False, seed=71, tampers=2)
    ts.check_verify_exp(L, sec=10, K=6, with_q=True, seed=72, tampers=2)
    ts.check_verify_membership(L, [3, 5, 7, 11, 13], 3, seed=73)
    for kind in ('equality', 'mult', 'pointadd'):
        ts.check_verify_small(L, kind, seed=74, tampers=2)
        ts.check_prove_small(L, kind)
    ts.check_prove_exp(L, sec=10, with_q=True, seed=75, B=2)
    ts.check_prove_membership(L, [3, 5, 7, 11, 13, 17], [0, 5], seed=76)
