"""Restatement of src/commit/{pedersen,equality,mult}.ts.

TEST INFRASTRUCTURE (oracle) — see oracle/__init__.py.
"""
from __future__ import annotations

from dataclasses import dataclass

from .big import rnd
from .curves import hash_points
from .multimult import MultiMult, Relation


class Commitment:
    # pedersen.ts:21-36
    __slots__ = ('p', 'r')

    def __init__(self, p, r):
        self.p, self.r = p, r

    def add(self, c): return Commitment(self.p.add(c.p), self.r.add(c.r))
    def sub(self, c): return Commitment(self.p.sub(c.p), self.r.sub(c.r))

    def mul(self, k: int):
        sk = self.p.group.new_scalar(k)
        return Commitment(self.p.mul(sk), self.r.mul(sk))


class PedersenParams:
    # pedersen.ts:40-59
    def __init__(self, c, g, h):
        self.c, self.g, self.h = c, g, h

    def eq(self, o): return self.c.eq(o.c) and self.g.eq(o.g) and self.h.eq(o.h)

    def commit(self, value: int, tape) -> Commitment:
        # pedersen.ts:53-58:  r <- random; p = h.dblmul(r, g, v)
        r = self.c.random_scalar(tape)
        v = self.c.new_scalar(value)
        return Commitment(self.h.dblmul(r, self.g, v), r)


def generate_pedersen_params(c, tape, g=None) -> PedersenParams:
    # pedersen.ts:61-69
    if g is None:
        g = c.generator()
    r = c.random_scalar(tape)
    return PedersenParams(c, g, g.mul(r))


# ---------------------------------------------------------------- equality.ts
@dataclass
class EqualityProof:
    A_1: object
    A_2: object
    t_x: object
    t_r1: object
    t_r2: object

    def eq(self, o):
        return (self.A_1.eq(o.A_1) and self.A_2.eq(o.A_2) and self.t_x.eq(o.t_x)
                and self.t_r1.eq(o.t_r1) and self.t_r2.eq(o.t_r2))


def prove_equality(params, x, C1, C2, tape) -> EqualityProof:
    # equality.ts:60-78
    k = rnd(params.c.order, tape)
    A1 = params.commit(k, tape)
    A2 = params.commit(k, tape)
    c = hash_points([C1.p, C2.p, A1.p, A2.p])
    cc = params.c.new_scalar(c)
    xx = params.c.new_scalar(x)
    kk = params.c.new_scalar(k)
    tx = kk.sub(cc.mul(xx))
    tr1 = A1.r.sub(cc.mul(C1.r))
    tr2 = A2.r.sub(cc.mul(C2.r))
    return EqualityProof(A1.p, A2.p, tx, tr1, tr2)


def aggregate_equality(params, C1, C2, pi, multi, tape) -> bool:
    # equality.ts:94-116
    challenge = hash_points([C1, C2, pi.A_1, pi.A_2])
    cc = params.c.new_scalar(challenge)
    one = params.c.new_scalar(1)
    A1rel = Relation(params.c, tape)
    A1rel.insert(params.g, pi.t_x)
    A1rel.insert(params.h, pi.t_r1)
    A1rel.insert(C1, cc)
    A1rel.insert(pi.A_1.neg(), one)
    A2rel = Relation(params.c, tape)
    A2rel.insert(params.g, pi.t_x)
    A2rel.insert(params.h, pi.t_r2)
    A2rel.insert(C2, cc)
    A2rel.insert(pi.A_2.neg(), one)
    A1rel.drain(multi)
    A2rel.drain(multi)
    return True


def verify_equality(params, C1, C2, pi, tape) -> bool:
    # equality.ts:80-92
    multi = MultiMult(params.c)
    if not aggregate_equality(params, C1, C2, pi, multi, tape):
        return False
    return multi.evaluate().is_identity()


# -------------------------------------------------------------------- mult.ts
@dataclass
class MultProof:
    C_4: object
    A_x: object
    A_y: object
    A_z: object
    A_4_1: object
    A_4_2: object
    t_x: object
    t_y: object
    t_z: object
    t_rx: object
    t_ry: object
    t_rz: object
    t_r4: object

    def eq(self, o):
        return all(getattr(self, f).eq(getattr(o, f)) for f in self.__dataclass_fields__)


def prove_mult(params, x, y, z, Cx, Cy, Cz, tape) -> MultProof:
    # mult.ts:93-131
    c_ = params.c
    xx = c_.new_scalar(x)
    C4 = Cy.p.mul(xx)
    r4 = Cy.r.mul(xx)
    k_x = rnd(c_.order, tape)
    k_y = rnd(c_.order, tape)
    k_z = rnd(c_.order, tape)
    kx = c_.new_scalar(k_x)
    Ax = params.commit(k_x, tape)
    Ay = params.commit(k_y, tape)
    Az = params.commit(k_z, tape)
    A4_1 = params.commit(k_z, tape)
    A4_2 = Cy.p.mul(kx)
    c = hash_points([Cx.p, Cy.p, Cz.p, C4, Ax.p, Ay.p, Az.p, A4_1.p, A4_2])
    cc = c_.new_scalar(c)
    ky, kz = c_.new_scalar(k_y), c_.new_scalar(k_z)
    yy, zz = c_.new_scalar(y), c_.new_scalar(z)
    t_x = kx.sub(cc.mul(xx))
    t_y = ky.sub(cc.mul(yy))
    t_z = kz.sub(cc.mul(zz))
    t_rx = Ax.r.sub(cc.mul(Cx.r))
    t_ry = Ay.r.sub(cc.mul(Cy.r))
    t_rz = Az.r.sub(cc.mul(Cz.r))
    t_r4 = A4_1.r.sub(cc.mul(r4))
    return MultProof(C4, Ax.p, Ay.p, Az.p, A4_1.p, A4_2, t_x, t_y, t_z, t_rx, t_ry, t_rz, t_r4)


def aggregate_mult(params, Cx, Cy, Cz, pi, multi, tape) -> bool:
    # mult.ts:148-175
    c_ = params.c
    challenge = hash_points([Cx, Cy, Cz, pi.C_4, pi.A_x, pi.A_y, pi.A_z, pi.A_4_1, pi.A_4_2])
    cc = c_.new_scalar(challenge)
    one = c_.new_scalar(1)
    A_xrel = Relation(c_, tape)
    A_xrel.insert_m([params.g, params.h, Cx, pi.A_x.neg()], [pi.t_x, pi.t_rx, cc, one])
    A_yrel = Relation(c_, tape)
    A_yrel.insert_m([params.g, params.h, Cy, pi.A_y.neg()], [pi.t_y, pi.t_ry, cc, one])
    A_zrel = Relation(c_, tape)
    A_zrel.insert_m([params.g, params.h, Cz, pi.A_z.neg()], [pi.t_z, pi.t_rz, cc, one])
    A_4_1rel = Relation(c_, tape)
    A_4_1rel.insert_m([params.g, params.h, pi.C_4, pi.A_4_1.neg()], [pi.t_z, pi.t_r4, cc, one])
    A_4_2rel = Relation(c_, tape)
    A_4_2rel.insert_m([Cy, pi.C_4, pi.A_4_2.neg()], [pi.t_x, cc, one])
    A_xrel.drain(multi)
    A_yrel.drain(multi)
    A_zrel.drain(multi)
    A_4_1rel.drain(multi)
    A_4_2rel.drain(multi)
    return True


def verify_mult(params, Cx, Cy, Cz, pi, tape) -> bool:
    # mult.ts:133-146
    multi = MultiMult(params.c)
    if not aggregate_mult(params, Cx, Cy, Cz, pi, multi, tape):
        return False
    return multi.evaluate().is_identity()
