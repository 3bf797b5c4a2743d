"""Restatement of src/proofGK/{gk,interpolate}.ts (Groth-Kohlweiss 1-of-N, scalar variant).

TEST INFRASTRUCTURE (oracle) — see oracle/__init__.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

from .big import exp_mod, inv_mod, pos_mod, rnd
from .curves import hash_points
from .multimult import MultiMult, Relation


def _eval_poly(coeff, x, m):
    # interpolate.ts:19-25
    ret = 0
    for c in reversed(coeff):
        ret = pos_mod(c + x * ret, m)
    return ret


def interpolate(x, y, m):
    # interpolate.ts:27-70.  The reference keeps s[] signed/unreduced (JS %
    # keeps the sign); all observable values pass through posMod, so eager
    # canonical reduction gives identical results.
    if len(x) != len(y):
        raise ValueError('inconsistent args')
    n = len(x)
    if n == 0:
        return []
    s = [0] * (n + 1)
    coeff = [0] * n
    s[n] = 1
    s[n - 1] = -x[0] % m
    for i in range(1, n):
        for j in range(n - i - 1, n - 1):
            s[j] = (s[j] - x[i] * s[j + 1]) % m
        s[n - 1] = (s[n - 1] - x[i]) % m
    for i in range(n):
        phi = 0
        for j in range(n, 0, -1):
            phi = j * s[j] + x[i] * phi
        phi = pos_mod(phi, m)
        ff = inv_mod(phi, m) % m
        b = 1
        for j in range(n - 1, -1, -1):
            coeff[j] = pos_mod(coeff[j] + b * ff * y[i], m)
            b = s[j] + x[i] * b
    for i in range(n):
        if y[i] != _eval_poly(coeff, x[i], m):
            raise ValueError('incorrect interpolation')
    return coeff


@dataclass
class GKProof:
    # gk.ts:31-73
    cl: list
    ca: list
    cb: list
    cd: list
    f: list
    za: list
    zb: list
    zd: object

    def eq(self, o):
        def cmp(a, b):
            return len(a) == len(b) and all(u.eq(v) for u, v in zip(a, b))
        return (cmp(self.cl, o.cl) and cmp(self.ca, o.ca) and cmp(self.cb, o.cb) and cmp(self.cd, o.cd)
                and cmp(self.f, o.f) and cmp(self.za, o.za) and cmp(self.zb, o.zb) and self.zd.eq(o.zd))


def _ceil_log2(v: int) -> int:
    # Math.ceil(Math.log2(v)) (gk.ts:80,103,207); float log2 like the reference
    return math.ceil(math.log2(v))


def pad(vals, c):
    # gk.ts:75-86
    ret = [c.new_scalar(v) for v in vals]
    pad_len = 2 ** _ceil_log2(len(vals))
    for _ in range(len(vals), pad_len):
        ret.append(ret[0])
    return ret


def gk_commit(params, val, blinder):
    # gk.ts:88-92
    o = params.c.order
    return params.g.dblmul(params.c.new_scalar(pos_mod(val, o)), params.h, params.c.new_scalar(pos_mod(blinder, o)))


def prove_membership(params, com, index, initial_values, tape) -> GKProof:
    # gk.ts:94-195
    values = pad(initial_values, params.c)
    c = params.c
    n = _ceil_log2(len(values))
    eli = []
    l_tmp = index
    for _ in range(n):
        eli.append(l_tmp % 2)
        l_tmp //= 2
    ri, ai, si, ti, rho = [], [], [], [], []
    for _ in range(n):
        ri.append(rnd(c.order, tape))
        ai.append(rnd(c.order, tape))
        si.append(rnd(c.order, tape))
        ti.append(rnd(c.order, tape))
        rho.append(rnd(c.order, tape))
    cl, ca, cb, cd = [], [], [], []
    for i in range(n):
        cl.append(gk_commit(params, eli[i], ri[i]))
        ca.append(gk_commit(params, ai[i], si[i]))
        cb.append(gk_commit(params, eli[i] * ai[i], ti[i]))
    omegas = list(range(n))
    dv = []
    for w in omegas:
        f0j, f1j, ratio = [], [], []
        for j in range(n):
            f0j.append(pos_mod((1 - eli[j]) * w - ai[j], c.order))
            f1j.append(pos_mod(eli[j] * w + ai[j], c.order))
            ratio.append(pos_mod(f1j[j] * inv_mod(f0j[j], c.order), c.order))
        prod = 1
        for v in f0j:
            prod = pos_mod(prod * v, c.order)
        p = [prod]
        for i in range(n):
            oldlen = len(p)
            for j in range(oldlen):
                p.append(pos_mod(ratio[i] * p[j], c.order))
        dval = 0
        for i in range(len(values)):
            dval = pos_mod(dval + (values[index].k - values[i].k) * p[i], c.order)
        dv.append(dval)
    di = interpolate(omegas, dv, c.order)
    for i in range(n):
        cd.append(gk_commit(params, di[i], rho[i]))
    x = hash_points(cl + ca + cb + cd)
    f, za, zb = [], [], []
    zd = (com.r.k * exp_mod(x, n, c.order)) % c.order
    for i in range(n):
        f.append(c.new_scalar(pos_mod(eli[i] * x + ai[i], c.order)))
        za.append(c.new_scalar(pos_mod(ri[i] * x + si[i], c.order)))
        zb.append(c.new_scalar(pos_mod(ri[i] * (x - f[i].k) + ti[i], c.order)))
    for i in range(n):
        zd = pos_mod(zd - rho[i] * exp_mod(x, i, c.order), c.order)
    return GKProof(cl, ca, cb, cd, f, za, zb, c.new_scalar(zd))


def verify_membership(params, com, init_vec, proof, tape) -> bool:
    # gk.ts:197-262
    c = params.c
    multi = MultiMult(c)
    vec = pad(init_vec, c)
    n = _ceil_log2(len(vec))
    if any(n != len(a) for a in (proof.cl, proof.ca, proof.cb, proof.cd, proof.f, proof.za, proof.zb)):
        return False
    f = proof.f
    x = hash_points(proof.cl + proof.ca + proof.cb + proof.cd)
    multi.add_known(params.g)
    multi.add_known(params.h)
    for i in range(n):
        rel0 = Relation(c, tape)
        rel0.insert_m([proof.cl[i], proof.ca[i], params.g, params.h],
                      [c.new_scalar(x), c.new_scalar(1), proof.f[i].neg(), proof.za[i].neg()])
        rel0.drain(multi)
        rel1 = Relation(c, tape)
        rel1.insert_m([proof.cl[i], proof.cb[i], params.h],
                      [c.new_scalar(pos_mod(x - f[i].k, c.order)), c.new_scalar(1), proof.zb[i].neg()])
        rel1.drain(multi)
    total = 0
    for i in range(len(vec)):
        pix = 1
        for j in range(n):
            if i & (1 << j):
                pix = pos_mod(pix * f[j].k, c.order)
            else:
                pix = pos_mod(pix * (x - f[j].k), c.order)
        total = pos_mod(total + vec[i].k * pix, c.order)
    rel_final = Relation(c, tape)
    for i in range(n):
        rel_final.insert(proof.cd[i], c.new_scalar(pos_mod(-exp_mod(x, i, c.order), c.order)))
    rel_final.insert(com, c.new_scalar(exp_mod(x, n, c.order)))
    rel_final.insert_m([params.g, params.h], [c.new_scalar(pos_mod(-total, c.order)), proof.zd.neg()])
    rel_final.drain(multi)
    return multi.evaluate().is_identity()
