"""Restatement of src/exp/{pointAdd,exp}.ts.

TEST INFRASTRUCTURE (oracle) — see oracle/__init__.py.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

from .big import inv_mod, is_odd, pos_mod, rnd_range
from .commit import (Commitment, EqualityProof, MultProof, aggregate_equality, aggregate_mult,
                     prove_equality, prove_mult)
from .curves import hash_points
from .multimult import MultiMult, Relation


@dataclass
class PointAddProof:
    # pointAdd.ts:28-76
    C_8: object
    C_10: object
    C_11: object
    C_13: object
    pi_8: MultProof
    pi_10: MultProof
    pi_11: MultProof
    pi_13: MultProof
    pi_x: EqualityProof
    pi_y: EqualityProof

    def eq(self, o):
        return all(getattr(self, f).eq(getattr(o, f)) for f in self.__dataclass_fields__)


def prove_point_add(params, P, Q, R, PX, PY, QX, QY, RX, RY, tape) -> PointAddProof:
    # pointAdd.ts:92-163
    if not P.add(Q).eq(R):
        raise ValueError("Points don't add up!")
    prime = params.c.order
    C1, C2, C3, C4, C5, C6 = PX, QX, RX, PY, QY, RY
    coordP = P.to_affine()
    coordQ = Q.to_affine()
    coordR = R.to_affine()
    if not coordP:
        raise ValueError('P is at infinity')
    if not coordQ:
        raise ValueError('Q is at infinity')
    if not coordR:
        raise ValueError('R is at infinity')
    x1, y1 = coordP
    x2, y2 = coordQ
    x3, _ = coordR
    i7 = pos_mod(x2 - x1, prime)
    i8 = inv_mod(i7, prime)
    i9 = pos_mod(y2 - y1, prime)
    i10 = pos_mod(i8 * i9, prime)
    i11 = pos_mod(i10 * i10, prime)
    i12 = pos_mod(x1 - x3, prime)
    i13 = pos_mod(i10 * i12, prime)
    C7 = C2.sub(C1)
    C8 = params.commit(i8, tape)
    C9 = C5.sub(C4)
    C10 = params.commit(i10, tape)
    C11 = params.commit(i11, tape)
    C12 = C1.sub(C3)
    C13 = params.commit(i13, tape)
    C14 = Commitment(params.g, params.c.new_scalar(0))
    pi8 = prove_mult(params, i7, i8, 1, C7, C8, C14, tape)
    pi10 = prove_mult(params, i8, i9, i10, C8, C9, C10, tape)
    pi11 = prove_mult(params, i10, i10, i11, C10, C10, C11, tape)
    Cint = Commitment(C3.p.add(C1.p).add(C2.p), C3.r.add(C1.r).add(C2.r))
    pix = prove_equality(params, i11, C11, Cint, tape)
    pi13 = prove_mult(params, i10, i12, i13, C10, C12, C13, tape)
    Cint = Commitment(C6.p.add(C4.p), C6.r.add(C4.r))
    piy = prove_equality(params, i13, C13, Cint, tape)
    return PointAddProof(C8.p, C10.p, C11.p, C13.p, pi8, pi10, pi11, pi13, pix, piy)


def aggregate_point_add(params, PX, PY, QX, QY, RX, RY, pi, multi, tape) -> bool:
    # pointAdd.ts:199-259
    C1, C2, C3, C4, C5, C6 = PX, QX, RX, PY, QY, RY
    C7 = C2.sub(C1)
    C9 = C5.sub(C4)
    C12 = C1.sub(C3)
    C_14 = params.g
    if not aggregate_mult(params, C7, pi.C_8, C_14, pi.pi_8, multi, tape):
        return False
    if not aggregate_mult(params, pi.C_8, C9, pi.C_10, pi.pi_10, multi, tape):
        return False
    if not aggregate_mult(params, pi.C_10, pi.C_10, pi.C_11, pi.pi_11, multi, tape):
        return False
    Cint = C3.add(C1).add(C2)
    if not aggregate_equality(params, pi.C_11, Cint, pi.pi_x, multi, tape):
        return False
    if not aggregate_mult(params, pi.C_10, C12, pi.C_13, pi.pi_13, multi, tape):
        return False
    Cint = C4.add(C6)
    if not aggregate_equality(params, pi.C_13, Cint, pi.pi_y, multi, tape):
        return False
    return True


def verify_point_add(params, PX, PY, QX, QY, RX, RY, pi, tape) -> bool:
    # pointAdd.ts:181-197
    multi = MultiMult(params.c)
    if not aggregate_point_add(params, PX, PY, QX, QY, RX, RY, pi, multi, tape):
        return False
    return multi.evaluate().is_identity()


# --------------------------------------------------------------------- exp.ts
@dataclass
class ExpProof:
    # exp.ts:26-84
    A: object
    Tx: object
    Ty: object
    alpha: Optional[object] = None
    beta1: Optional[object] = None
    beta2: Optional[object] = None
    beta3: Optional[object] = None
    z: Optional[object] = None
    z2: Optional[object] = None
    proof: Optional[PointAddProof] = None
    r1: Optional[object] = None
    r2: Optional[object] = None

    def eq(self, o):
        c0 = self.A.eq(o.A) and self.Tx.eq(o.Tx) and self.Ty.eq(o.Ty)

        def both(a, b):
            return a.eq(b) if (a is not None and b is not None) else False
        r0 = (both(self.alpha, o.alpha) and both(self.beta1, o.beta1)
              and both(self.beta2, o.beta2) and both(self.beta3, o.beta3))
        r1 = (both(self.z, o.z) and both(self.z2, o.z2) and both(self.proof, o.proof)
              and both(self.r1, o.r1) and both(self.r2, o.r2))
        return c0 and (r0 or r1)


def padded_bits(val: int, length: int):
    # exp.ts:86-93
    ret = []
    for _ in range(length):
        ret.append(val % 2 == 1)
        val >>= 1
    return ret


def generate_indices(indnum: int, limit: int, tape):
    # exp.ts:95-109 (Knuth Algorithm P; `ret.slice(indnum)` is a no-op)
    ret = list(range(limit))
    for i in range(limit - 2):
        j = rnd_range(i, limit - 1, tape)
        ret[i], ret[j] = ret[j], ret[i]
    return ret


def prove_exp(paramsNIST, paramsWario, s, Cs, P, Px, Py, secparam, tape, Q=None):
    # exp.ts:126-231
    alpha, r, T, A, Tx, Ty = [], [], [], [], [], []
    for i in range(secparam):
        alpha.append(paramsNIST.c.random_scalar(tape))
        r.append(paramsNIST.c.random_scalar(tape))
        T.append(paramsNIST.g.mul(alpha[i]))
        A.append(T[i].add(paramsNIST.h.mul(r[i])))
        coordT = T[i].to_affine()
        if not coordT:
            raise ValueError('T[i] is at infinity')
        x, y = coordT
        Tx.append(paramsWario.commit(x, tape))
        Ty.append(paramsWario.commit(y, tape))
    arr = [Px.p, Py.p]
    for i in range(secparam):
        arr += [A[i], Tx[i].p, Ty[i].p]
    challenge = hash_points(arr)
    all_proofs = []
    for i in range(secparam):
        if is_odd(challenge):
            proof = ExpProof(A[i], Tx[i].p, Ty[i].p, alpha[i], r[i], Tx[i].r, Ty[i].r)
        else:
            z = alpha[i].sub(paramsNIST.c.new_scalar(s))
            T1 = paramsNIST.g.mul(z)
            if Q is not None:
                T1 = T1.add(Q)
            coordT1 = T1.to_affine()
            if not coordT1:
                raise ValueError('T1 is at infinity')
            x, y = coordT1
            T1x = paramsWario.commit(x, tape)
            T1y = paramsWario.commit(y, tape)
            pap = prove_point_add(paramsWario, T1, P, T[i], T1x, T1y, Px, Py, Tx[i], Ty[i], tape)
            proof = ExpProof(A[i], Tx[i].p, Ty[i].p, None, None, None, None,
                             z, r[i].sub(Cs.r), pap, T1x.r, T1y.r)
        all_proofs.append(proof)
        challenge >>= 1
    return all_proofs


def verify_exp(paramsNIST, paramsWario, Clambda, Px, Py, pi, secparam, tape, Q=None) -> bool:
    # exp.ts:233-349
    if secparam > len(pi):
        raise ValueError('security level not achieved')
    multiW = MultiMult(paramsWario.c)
    multiN = MultiMult(paramsNIST.c)
    multiW.add_known(paramsWario.g)
    multiW.add_known(paramsWario.h)
    multiN.add_known(paramsNIST.g)
    multiN.add_known(paramsNIST.h)
    multiN.add_known(Clambda)
    arr = [Px, Py]
    for e in pi:
        arr += [e.A, e.Tx, e.Ty]
    challenge = hash_points(arr)
    indices = generate_indices(secparam, len(pi), tape)
    bits = padded_bits(challenge, len(pi))
    cN, cW = paramsNIST.c, paramsWario.c
    for j in range(secparam):
        i = indices[j]
        e = pi[i]
        if bits[i]:
            if any(v is None for v in (e.alpha, e.beta1, e.beta2, e.beta3)):
                raise ValueError('params not found')
            T = paramsNIST.g.mul(e.alpha)
            relA = Relation(cN, tape)
            relA.insert_m([T, paramsNIST.h, e.A.neg()], [cN.new_scalar(1), e.beta1, cN.new_scalar(1)])
            relA.drain(multiN)
            coordT = T.to_affine()
            if not coordT:
                raise ValueError('T is at infinity')
            sx, sy = cW.new_scalar(coordT[0]), cW.new_scalar(coordT[1])
            relTx = Relation(cW, tape)
            relTy = Relation(cW, tape)
            relTx.insert_m([paramsWario.g, paramsWario.h, e.Tx.neg()], [sx, e.beta2, cW.new_scalar(1)])
            relTy.insert_m([paramsWario.g, paramsWario.h, e.Ty.neg()], [sy, e.beta3, cW.new_scalar(1)])
            relTx.drain(multiW)
            relTy.drain(multiW)
        else:
            if any(v is None for v in (e.z, e.z2, e.proof, e.r1, e.r2)):
                raise ValueError('params not found')
            T1 = paramsNIST.g.mul(e.z)
            relA = Relation(cN, tape)
            relA.insert_m([T1, Clambda, e.A.neg(), paramsNIST.h],
                          [cN.new_scalar(1), cN.new_scalar(1), cN.new_scalar(1), e.z2])
            relA.drain(multiN)
            if Q is not None:
                T1 = T1.add(Q)
            coordT1 = T1.to_affine()
            if not coordT1:
                raise ValueError('T1 is at infinity')
            sx, sy = cW.new_scalar(coordT1[0]), cW.new_scalar(coordT1[1])
            T1x = paramsWario.g.dblmul(sx, paramsWario.h, e.r1)
            T1y = paramsWario.g.dblmul(sy, paramsWario.h, e.r2)
            if not aggregate_point_add(paramsWario, T1x, T1y, Px, Py, e.Tx, e.Ty, e.proof, multiW, tape):
                return False
    return multiW.evaluate().is_identity() and multiN.evaluate().is_identity()
