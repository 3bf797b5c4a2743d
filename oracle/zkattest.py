"""Restatement of src/zkpAttestList.ts (top-level ZKAttest protocol).

TEST INFRASTRUCTURE (oracle) — see oracle/__init__.py.
WebCrypto `exportKey('raw')` (zkpAttestList.ts:95,113) yields the 65-byte SEC1
uncompressed key; here the caller passes those bytes directly.
"""
from __future__ import annotations

from dataclasses import dataclass

from .big import bit_len, from_bytes, inv_mod, pos_mod
from .commit import PedersenParams, generate_pedersen_params
from .curves import p256, tomEdwards256
from .exp import prove_exp, verify_exp
from .gk import GKProof, prove_membership, verify_membership


@dataclass
class SignatureProofList:
    # zkpAttestList.ts:29-61
    R: object
    comS1: object
    keyXcom: object
    keyYcom: object
    expProof: list
    membershipProof: GKProof

    def eq(self, o):
        return (self.R.eq(o.R) and self.comS1.eq(o.comS1) and self.keyXcom.eq(o.keyXcom)
                and self.keyYcom.eq(o.keyYcom) and len(self.expProof) == len(o.expProof)
                and all(a.eq(b) for a, b in zip(self.expProof, o.expProof))
                and self.membershipProof.eq(o.membershipProof))


@dataclass
class SystemParametersList:
    # zkpAttestList.ts:65-78
    NistGroup: PedersenParams
    ProofGroup: PedersenParams
    SecLevel: int

    def eq(self, o):
        return self.NistGroup.eq(o.NistGroup) and self.ProofGroup.eq(o.ProofGroup) and self.SecLevel == o.SecLevel


def truncate_to_n(msg: int, n: int) -> int:
    # zkpAttestList.ts:80-86
    delta = bit_len(msg) - bit_len(n)
    if delta > 0:
        msg >>= delta
    return msg


def generate_params_list(tape, sec_level: int = 80, proof_group=tomEdwards256) -> SystemParametersList:
    # zkpAttestList.ts:88-92 (draw order: p256 scalar, then tomEdwards256 scalar).  `proof_group`: the reference
    # hard-codes tomEdwards256 here, but SystemParametersList.ProofGroup (zkpAttestList.ts:70) is any group whose order
    # is p256.p — war256 (instances.ts:34-41) is the other one the JSON initialiser accepts (instances.ts:58-69).
    nist = generate_pedersen_params(p256, tape)
    proof = generate_pedersen_params(proof_group, tape)
    return SystemParametersList(nist, proof, sec_level)


def key_to_int(pk_bytes: bytes) -> int:
    # zkpAttestList.ts:94-102
    pt = p256.deserialize_point(pk_bytes)
    c = pt.to_affine()
    if not c:
        raise ValueError('invalid public key')
    return c[0]


def prove_signature_list(params, msg_hash: bytes, sig_bytes: bytes, pk_bytes: bytes, which: int, keys, tape):
    # zkpAttestList.ts:104-145
    ec = p256
    pk_point = p256.deserialize_point(pk_bytes)
    pk_coords = pk_point.to_affine()
    if not pk_coords:
        raise ValueError('invalid public key')
    ln = len(sig_bytes)
    n = ec.order
    z = truncate_to_n(from_bytes(msg_hash), n)
    r = from_bytes(sig_bytes[:ln // 2])
    s = from_bytes(sig_bytes[ln // 2:])
    sinv = inv_mod(s, n)
    u1 = pos_mod(sinv * z, n)
    u2 = pos_mod(sinv * r, n)
    R = ec.generator().mul(ec.new_scalar(u1)).add(pk_point.mul(ec.new_scalar(u2)))
    rinv = inv_mod(r, n)
    s1 = pos_mod(rinv * s, n)
    z1 = pos_mod(rinv * z, n)
    Q = ec.generator().mul(ec.new_scalar(z1))
    params_sig_exp = PedersenParams(p256, R, params.NistGroup.h)
    comS1 = params_sig_exp.commit(s1, tape)
    pkX = params.ProofGroup.commit(pk_coords[0], tape)
    pkY = params.ProofGroup.commit(pk_coords[1], tape)
    sig_proof = prove_exp(params_sig_exp, params.ProofGroup, s1, comS1, pk_point, pkX, pkY,
                          params.SecLevel, tape, Q)
    membership = prove_membership(params.ProofGroup, pkX, which, keys, tape)
    return SignatureProofList(R, comS1.p, pkX.p, pkY.p, sig_proof, membership)


def verify_signature_list(params, msg_hash: bytes, keys, proof: SignatureProofList, tape, secparam: int = 20) -> bool:
    # zkpAttestList.ts:147-184 (the reference passes the literal secparam = 20 at :177; other values exercise
    # verifyExp's own parameter, exp.ts:233-245, as test/exp/exp.test.ts does with 80)
    ec = p256
    n = ec.order
    z = truncate_to_n(from_bytes(msg_hash), n)
    R = proof.R
    coordR = R.to_affine()
    if not coordR:
        raise ValueError('R is at infinity')
    rinv = inv_mod(coordR[0], n)
    params_sig_exp = PedersenParams(p256, R, params.NistGroup.h)
    z1 = pos_mod(rinv * z, n)
    Q = ec.generator().mul(ec.new_scalar(z1))
    if not verify_membership(params.ProofGroup, proof.keyXcom, keys, proof.membershipProof, tape):
        return False
    if not verify_exp(params_sig_exp, params.ProofGroup, proof.comS1, proof.keyXcom, proof.keyYcom,
                      proof.expProof, secparam, tape, Q):
        return False
    return True
