"""Flat binary proof layout used at the C-ABI (include/zkattest.h).

TEST INFRASTRUCTURE (oracle) — see oracle/__init__.py.

The reference has no binary whole-proof format (only typedjson JSON,
src/serde.ts:21-36).  The layout is the concatenation, in class-field order,
of the reference's own primitives `Point.toBytes()` (weier.ts:244-255: 65 B;
edwards.ts:195-203: 67 B) and `Scalar.toBytes()` (group.ts:196-199: 32 B
p256 / 33 B tomEdwards256):

  proof   := R(65) comS1(65) keyXcom(67) keyYcom(67) rep[80] GK
  rep     := tag(1) A(65) Tx(67) Ty(67) body
  body    := tag==1: alpha(32) beta1(32) beta2(33) beta3(33)            (exp.ts:31-34)
             tag==0: z(32) z2(32) PointAddProof(3266) r1(33) r2(33)     (exp.ts:36-40)
  PointAddProof := C_8 C_10 C_11 C_13 (4x67) pi_8 pi_10 pi_11 pi_13 (4x633) pi_x pi_y (2x233)
  MultProof     := C_4 A_x A_y A_z A_4_1 A_4_2 (6x67) t_x t_y t_z t_rx t_ry t_rz t_r4 (7x33)
  EqualityProof := A_1 A_2 (2x67) t_x t_r1 t_r2 (3x33)
  GK      := n(1) cl[n] ca[n] cb[n] cd[n] (67 each) f[n] za[n] zb[n] (33 each) zd(33)

The P-256 identity (weier.ts:247 serialises it as ONE 0x00 byte) cannot occur
in a valid proof slot; the fixed 65-byte slot encodes it as 65 zero bytes.
"""
from __future__ import annotations

from .commit import EqualityProof, MultProof
from .curves import p256, tomEdwards256
from .exp import ExpProof, PointAddProof
from .gk import GKProof
from .zkattest import SignatureProofList

NP, WP, NS, WS = 65, 67, 32, 33
EQ_LEN = 2 * WP + 3 * WS            # 233
MULT_LEN = 6 * WP + 7 * WS          # 633
PA_LEN = 4 * WP + 4 * MULT_LEN + 2 * EQ_LEN   # 3266
REP1_LEN = 1 + NP + 2 * WP + 2 * NS + 2 * WS  # 330
REP0_LEN = 1 + NP + 2 * WP + 2 * NS + PA_LEN + 2 * WS  # 3596
HEAD_LEN = 2 * NP + 2 * WP          # 264
PROOF_GROUP = tomEdwards256


def set_proof_group(group):
    """Select the ProofGroup of the flat layout: tomEdwards256 (default; 67-byte points, 33-byte scalars) or war256
    (65 / 32).  The grammar is the same, only the primitive sizes change (group.ts:49-52 sizeFieldBytes)."""
    global WP, WS, EQ_LEN, MULT_LEN, PA_LEN, REP1_LEN, REP0_LEN, HEAD_LEN, PROOF_GROUP
    PROOF_GROUP = group
    WS = group.size_field_bytes()
    WP = 1 + 2 * WS
    EQ_LEN = 2 * WP + 3 * WS
    MULT_LEN = 6 * WP + 7 * WS
    PA_LEN = 4 * WP + 4 * MULT_LEN + 2 * EQ_LEN
    REP1_LEN = 1 + NP + 2 * WP + 2 * NS + 2 * WS
    REP0_LEN = 1 + NP + 2 * WP + 2 * NS + PA_LEN + 2 * WS
    HEAD_LEN = 2 * NP + 2 * WP


def gk_len(n: int) -> int:
    return 1 + 4 * n * WP + (3 * n + 1) * WS


def proof_len(zero_bits: int, n: int, reps: int = 80) -> int:
    return HEAD_LEN + zero_bits * REP0_LEN + (reps - zero_bits) * REP1_LEN + gk_len(n)


def max_proof_len(n: int, reps: int = 80) -> int:
    return proof_len(reps, n, reps)


def _pt(p, size):
    b = p.to_bytes()
    if len(b) == 1:  # P-256 identity
        return bytes(size)
    assert len(b) == size
    return b


def ser_equality(pi: EqualityProof) -> bytes:
    return _pt(pi.A_1, WP) + _pt(pi.A_2, WP) + pi.t_x.to_bytes() + pi.t_r1.to_bytes() + pi.t_r2.to_bytes()


def ser_mult(pi: MultProof) -> bytes:
    out = b''.join(_pt(getattr(pi, f), WP) for f in ('C_4', 'A_x', 'A_y', 'A_z', 'A_4_1', 'A_4_2'))
    out += b''.join(getattr(pi, f).to_bytes() for f in ('t_x', 't_y', 't_z', 't_rx', 't_ry', 't_rz', 't_r4'))
    return out


def ser_point_add(pi: PointAddProof) -> bytes:
    return (_pt(pi.C_8, WP) + _pt(pi.C_10, WP) + _pt(pi.C_11, WP) + _pt(pi.C_13, WP)
            + ser_mult(pi.pi_8) + ser_mult(pi.pi_10) + ser_mult(pi.pi_11) + ser_mult(pi.pi_13)
            + ser_equality(pi.pi_x) + ser_equality(pi.pi_y))


def ser_exp(e: ExpProof) -> bytes:
    head = _pt(e.A, NP) + _pt(e.Tx, WP) + _pt(e.Ty, WP)
    if e.alpha is not None:
        return b'\x01' + head + e.alpha.to_bytes() + e.beta1.to_bytes() + e.beta2.to_bytes() + e.beta3.to_bytes()
    return (b'\x00' + head + e.z.to_bytes() + e.z2.to_bytes() + ser_point_add(e.proof)
            + e.r1.to_bytes() + e.r2.to_bytes())


def ser_gk(g: GKProof) -> bytes:
    n = len(g.cl)
    out = bytes([n])
    for arr in (g.cl, g.ca, g.cb, g.cd):
        out += b''.join(_pt(p, WP) for p in arr)
    for arr in (g.f, g.za, g.zb):
        out += b''.join(s.to_bytes() for s in arr)
    return out + g.zd.to_bytes()


def ser_proof(pr: SignatureProofList) -> bytes:
    out = _pt(pr.R, NP) + _pt(pr.comS1, NP) + _pt(pr.keyXcom, WP) + _pt(pr.keyYcom, WP)
    out += b''.join(ser_exp(e) for e in pr.expProof)
    return out + ser_gk(pr.membershipProof)


# ------------------------------------------------------------------ parsing
class _Rd:
    def __init__(self, b):
        self.b, self.o = bytes(b), 0

    def take(self, n):
        if self.o + n > len(self.b):
            raise ValueError('truncated proof')
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def npt(self):
        v = self.take(NP)
        return p256.identity() if v == bytes(NP) else p256.deserialize_point(v)

    def wpt(self): return PROOF_GROUP.deserialize_point(self.take(WP))
    def nsc(self): return p256.deserialize_scalar(self.take(NS))
    def wsc(self): return PROOF_GROUP.deserialize_scalar(self.take(WS))


def _de_mult(r):
    pts = [r.wpt() for _ in range(6)]
    scs = [r.wsc() for _ in range(7)]
    return MultProof(*pts, *scs)


def _de_eq(r):
    return EqualityProof(r.wpt(), r.wpt(), r.wsc(), r.wsc(), r.wsc())


def _de_pa(r):
    cs = [r.wpt() for _ in range(4)]
    ms = [_de_mult(r) for _ in range(4)]
    return PointAddProof(*cs, *ms, _de_eq(r), _de_eq(r))


def de_proof(b: bytes, reps: int = 80) -> SignatureProofList:
    r = _Rd(b)
    R, comS1, kx, ky = r.npt(), r.npt(), r.wpt(), r.wpt()
    exps = []
    for _ in range(reps):
        tag = r.take(1)[0]
        A, Tx, Ty = r.npt(), r.wpt(), r.wpt()
        if tag == 1:
            exps.append(ExpProof(A, Tx, Ty, r.nsc(), r.nsc(), r.wsc(), r.wsc()))
        elif tag == 0:
            z, z2 = r.nsc(), r.nsc()
            pa = _de_pa(r)
            exps.append(ExpProof(A, Tx, Ty, None, None, None, None, z, z2, pa, r.wsc(), r.wsc()))
        else:
            raise ValueError('bad repetition tag')
    n = r.take(1)[0]
    cl = [r.wpt() for _ in range(n)]
    ca = [r.wpt() for _ in range(n)]
    cb = [r.wpt() for _ in range(n)]
    cd = [r.wpt() for _ in range(n)]
    f = [r.wsc() for _ in range(n)]
    za = [r.wsc() for _ in range(n)]
    zb = [r.wsc() for _ in range(n)]
    zd = r.wsc()
    if r.o != len(r.b):
        raise ValueError('trailing bytes in proof')
    return SignatureProofList(R, comS1, kx, ky, exps, GKProof(cl, ca, cb, cd, f, za, zb, zd))
