"""The oracle against every fixed-answer / property test the reference has."""
import pytest

from oracle import commit as OC
from oracle import exp as OE
from oracle import flat, gk
from oracle import zkattest as OZ
from oracle.big import OsTape, Tape, inv_euclid, rnd, rnd_range
from oracle.curves import ALL_GROUPS, p256, tomEdwards256
from oracle.multimult import MultiMult, Relation
from zkp_ecdsa_b200 import synth


def test_inv_euclid_kat():
    # test/bignum/big.test.ts:19-21
    assert inv_euclid(3, 5) == 2 and inv_euclid(7, 41) == 6


def test_interpolate_kat():
    # test/proofGK/interpolate.test.ts:19-26
    assert gk.interpolate([1, 2, 3], [1, 2, 3], 401) == [0, 1, 0]


def test_rnd_rejection_and_range():
    # big.ts:171-185: rejection on byteLen(n) bytes
    t = Tape(bytes([200, 7]))
    assert rnd(80, t) == 7 and t.calls == 2
    t = Tape(bytes([3]))
    assert rnd_range(5, 79, t) == 8


@pytest.mark.parametrize('group', ALL_GROUPS, ids=lambda g: g.name)
def test_ec_properties(group):
    # test/curves/ec.test.ts:21-90 (10 iterations instead of 100 to keep the CPU suite short)
    tape = OsTape()
    P1 = group.generator().mul(group.new_scalar(group.order))
    assert group.is_on_group(P1) and P1.is_identity()
    P = group.generator().mul(group.random_scalar(tape))
    for _ in range(10):
        P = P.mul(group.random_scalar(tape))
        assert group.is_on_group(P)
    Q = P.mul(group.new_scalar(group.order - 1))
    assert P.add(Q).is_identity()
    k1, k2 = group.random_scalar(tape), group.random_scalar(tape)
    assert P.dblmul(k1, Q, k2).eq(P.mul(k1).add(Q.mul(k2)))
    ident = group.identity()
    assert group.deserialize_point(ident.to_bytes()).eq(ident)
    for _ in range(10):
        pt = group.generator().mul(group.random_scalar(tape))
        assert group.deserialize_point(pt.to_bytes()).eq(pt)


def test_multimult():
    # test/curves/multimult.test.ts:21-55
    tape = OsTape()
    g = tomEdwards256
    mm = MultiMult(g)
    mm2 = MultiMult(g)
    for _ in range(7):
        pt = g.generator().mul(g.random_scalar(tape))
        s = g.random_scalar(tape)
        mm.insert(pt, s)
        mm2.insert(pt, s)
    assert mm.evaluate().eq(mm2.evaluate_naive())
    rel = Relation(g, tape)
    s = g.random_scalar(tape)
    P = g.generator()
    rel.insert(P, s)
    rel.insert(P.mul(s).neg(), g.new_scalar(1))
    m3 = MultiMult(g)
    rel.drain(m3)
    assert m3.evaluate().is_identity()


def test_equality_mult_pointadd_roundtrip():
    # test/commit/{equality,mult}.test.ts, test/exp/pointAdd.test.ts
    tape = OsTape()
    params = OC.generate_pedersen_params(tomEdwards256, tape)
    x = rnd(tomEdwards256.order, tape)
    C1, C2 = params.commit(x, tape), params.commit(x, tape)
    assert OC.verify_equality(params, C1.p, C2.p, OC.prove_equality(params, x, C1, C2, tape), tape)
    q = tomEdwards256.order
    a, b = rnd(q, tape), rnd(q, tape)
    Ca, Cb, Cc = params.commit(a, tape), params.commit(b, tape), params.commit(a * b % q, tape)
    pi = OC.prove_mult(params, a, b, a * b % q, Ca, Cb, Cc, tape)
    assert OC.verify_mult(params, Ca.p, Cb.p, Cc.p, pi, tape)
    bad = OC.prove_mult(params, a, b, (a * b + 1) % q, Ca, Cb, Cc, tape)
    assert not OC.verify_mult(params, Ca.p, Cb.p, Cc.p, bad, tape)       # negative test the reference lacks
    P = p256.generator().mul(p256.random_scalar(tape))
    Q = p256.generator().mul(p256.random_scalar(tape))
    R = P.add(Q)
    cs = [params.commit(v, tape) for v in (*P.to_affine(), *Q.to_affine(), *R.to_affine())]
    pa = OE.prove_point_add(params, P, Q, R, cs[0], cs[1], cs[2], cs[3], cs[4], cs[5], tape)
    assert OE.verify_point_add(params, cs[0].p, cs[1].p, cs[2].p, cs[3].p, cs[4].p, cs[5].p, pa, tape)


def test_gk_roundtrip():
    # test/proofGK/gk.test.ts:22-30 (ring of 5 exercises the padding 5 -> 8)
    tape = OsTape()
    params = OC.generate_pedersen_params(tomEdwards256, tape)
    vec = [3, 5, 7, 11, 13]
    com = params.commit(11, tape)
    proof = gk.prove_membership(params, com, 3, vec, tape)
    assert gk.verify_membership(params, com.p, vec, proof, tape)
    assert not gk.verify_membership(params, params.commit(12, tape).p, vec, proof, tape)


def test_zkattest_roundtrip_and_flat_layout():
    # test/zkpAttestList.test.ts:28-54 with the flat layout standing in for JSON serde
    wl = synth.Workload(B=1, N=6, seed=11)
    tape = OsTape()
    params = OZ.generate_params_list(tape)
    ring = wl.ring_ints()
    assert OZ.key_to_int(wl.pk[0].tobytes()) == ring[int(wl.which[0])]
    ptape = OsTape()
    proof = OZ.prove_signature_list(params, wl.msg_hash[0].tobytes(), wl.sig[0].tobytes(), wl.pk[0].tobytes(),
                                    int(wl.which[0]), ring, ptape)
    z = sum(1 for e in proof.expProof if e.alpha is None)
    assert ptape.calls == 323 + 40 * z + 5 * 3
    data = flat.ser_proof(proof)
    assert len(data) == flat.proof_len(z, 3)
    back = flat.de_proof(data)
    assert back.eq(proof) and flat.ser_proof(back) == data
    assert OZ.verify_signature_list(params, wl.msg_hash[0].tobytes(), ring, back, OsTape())
    # tampering one response byte must be rejected (negative test the reference lacks)
    bad = bytearray(data)
    bad[-1] ^= 1
    assert not OZ.verify_signature_list(params, wl.msg_hash[0].tobytes(), ring, flat.de_proof(bytes(bad)), OsTape())
    other = bytes(32)
    assert not OZ.verify_signature_list(params, other, ring, back, OsTape())
