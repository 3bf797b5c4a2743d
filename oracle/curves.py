"""Restatement of src/curves/{group,weier,edwards,instances}.ts.

TEST INFRASTRUCTURE (oracle) — see oracle/__init__.py.
Counters in `STATS` tally field multiplications / inversions so bench.py can
report reference-equivalent modmul counts.
"""
from __future__ import annotations

import hashlib

from .big import from_bytes, inv_mod, pos_mod, rnd, to_bytes, verify_pos_range

STATS = {'inv': 0, 'mul_calls': 0, 'dblmul_calls': 0, 'hash_bytes': 0, 'hash_calls': 0}


class Group:
    # group.ts:20-67
    def __init__(self, name: str, p: int, order: int):
        self.name = name
        self.p = p
        self.order = order

    def size_field_bytes(self) -> int:
        # group.ts:49-52 -- bytes of the FIELD prime (33 for tomEdwards256)
        return -(-self.p.bit_length() // 8)

    def size_point_bytes(self) -> int:
        return 1 + 2 * self.size_field_bytes()

    def eq(self, g: 'Group') -> bool:
        return self.name == g.name

    def is_compat_point(self, pt) -> bool:
        if not self.eq(pt.group):
            raise ValueError('points not compatible')
        return True

    def is_compat_scalar(self, s) -> bool:
        if not self.eq(s.group):
            raise ValueError('scalar not compatible')
        return True

    def new_scalar(self, s: int) -> 'Scalar':
        return Scalar(self, s)

    def random_scalar(self, tape) -> 'Scalar':
        # group.ts:59-61
        return self.new_scalar(rnd(self.order, tape))

    def deserialize_scalar(self, a: bytes) -> 'Scalar':
        # group.ts:62-66
        s = from_bytes(a)
        verify_pos_range(s, self.order)
        return self.new_scalar(s)


class Scalar:
    # group.ts:159-218
    __slots__ = ('group', 'k')

    def __init__(self, group: Group, s: int):
        self.group = group
        self.k = pos_mod(s, group.order) if s else 0

    def base16(self) -> str:
        return format(self.k, 'x')  # no leading zeros; "0" for zero

    def eq(self, s: 'Scalar') -> bool:
        return self.group.eq(s.group) and self.k % self.group.order == s.k % s.group.order

    def add(self, s): return Scalar(self.group, self.k + s.k)
    def sub(self, s): return Scalar(self.group, self.k - s.k)
    def mul(self, s): return Scalar(self.group, self.k * s.k)
    def neg(self): return Scalar(self.group, -self.k)

    def to_bytes(self) -> bytes:
        # group.ts:196-199: length = sizeFieldBytes(), NOT bytes of the order
        return to_bytes(self.k % self.group.order, self.group.size_field_bytes())

    def is_zero(self): return self.k == 0
    def is_one(self): return self.k == 1

    def cmp(self, s) -> int:
        return -1 if self.k < s.k else (1 if self.k > s.k else 0)


_DIGITS = '0123456789abcdef'


class Point:
    """group.ts:71-153: generic 4-bit windowed mul / dblmul."""
    group: Group

    def sub(self, pt):
        return self.add(pt.neg())

    def dblmul(self, s1: Scalar, p2: 'Point', s2: Scalar):
        # group.ts:97-132
        self.group.is_compat_scalar(s1)
        self.group.is_compat_scalar(s2)
        self.group.is_compat_point(p2)
        STATS['dblmul_calls'] += 1
        mult1, mult2 = {}, {}
        curr1, curr2 = self.group.identity(), p2.group.identity()
        for d in _DIGITS:
            mult1[d] = curr1
            mult2[d] = curr2
            curr1 = curr1.add(self)
            curr2 = curr2.add(p2)
        k1, k2 = s1.base16(), s2.base16()
        if len(k1) < len(k2):
            k1 = k1.rjust(len(k2), '0')
        if len(k2) < len(k1):
            k2 = k2.rjust(len(k1), '0')
        q = self.group.identity()
        for i in range(len(k1)):
            q = q.dbl().dbl().dbl().dbl()
            q = q.add(mult1[k1[i]])
            q = q.add(mult2[k2[i]])
        return q

    def mul(self, s: Scalar):
        # group.ts:133-152
        self.group.is_compat_scalar(s)
        STATS['mul_calls'] += 1
        k = s.base16()
        q = self.group.identity()
        mults = {}
        curr = self.group.identity()
        for d in _DIGITS:
            mults[d] = curr
            curr = curr.add(self)
        for ki in k:
            q = q.dbl().dbl().dbl().dbl()
            q = q.add(mults[ki])
        return q


def hash_points(points) -> int:
    # group.ts:221-233: SHA-256 over concatenated toBytes(); first 10 bytes
    data = b''.join(p.to_bytes() for p in points)
    STATS['hash_bytes'] += len(data)
    STATS['hash_calls'] += 1
    return from_bytes(hashlib.sha256(data).digest()[:10])


# ----------------------------------------------------------------------------
# Weierstrass a = -3 (src/curves/weier.ts)
# ----------------------------------------------------------------------------
class WeierstrassGroup(Group):
    def __init__(self, name, p, a, b, order, gen):
        super().__init__(name, p, order)
        verify_pos_range(a, p)
        verify_pos_range(b, p)
        verify_pos_range(gen[0], p)
        verify_pos_range(gen[1], p)
        if pos_mod(a, p) != p - 3:  # weier.ts:41
            raise ValueError('only supports a=-3')
        self.a, self.b, self.gen = a, b, gen
        if not self.is_on_group(self.generator()):  # weier.ts:45-48
            raise ValueError('generator not on group')

    def identity(self):
        return WeierstrassPoint(self, 0, 1, 0)

    def generator(self):
        return WeierstrassPoint(self, self.gen[0], self.gen[1], 1)

    def is_on_group(self, pt) -> bool:
        # weier.ts:56-70
        p, a, b = self.p, self.a, self.b
        x, y, z = pt.x, pt.y, pt.z
        y2z = (y * y % p) * z % p
        x3 = (x * x * x) % p
        z2 = (z * z) % p
        axz2 = ((a * x) % p) * z2 % p
        bz3 = b * ((z2 * z) % p) % p
        return self.eq(pt.group) and pos_mod(y2z - (x3 + axz2 + bz3), p) == 0

    def deserialize_point(self, a: bytes):
        # weier.ts:74-89
        a = bytes(a)
        if len(a) == 1 and a[0] == 0:
            return self.identity()
        elif len(a) == self.size_point_bytes() and a[0] == 0x04:
            cs = self.size_field_bytes()
            x = from_bytes(a[1:1 + cs])
            y = from_bytes(a[1 + cs:])
            pt = WeierstrassPoint(self, x, y)
            if not self.is_on_group(pt):
                raise ValueError('point not in group')
            return pt
        raise ValueError('error deserializing Point')


class WeierstrassPoint(Point):
    __slots__ = ('group', 'x', 'y', 'z')

    def __init__(self, g, x, y, z=1):
        self.group, self.x, self.y, self.z = g, x, y, z

    def is_identity(self):
        return self.x == 0 and self.y != 0 and self.z == 0

    def eq(self, pt):
        p = self.group.p
        return (self.group.eq(pt.group)
                and (self.x * pt.z) % p == (pt.x * self.z) % p
                and (self.y * pt.z) % p == (pt.y * self.z) % p)

    def neg(self):
        return WeierstrassPoint(self.group, self.x, pos_mod(-self.y, self.group.p), self.z)

    def dbl(self):
        # weier.ts:133-175 (Renes-Costello-Batina 2015, Alg. 6, a = -3)
        x, y, z = self.x, self.y, self.z
        p, b = self.group.p, self.group.b
        t0 = x * x % p
        t1 = y * y % p
        t2 = z * z % p
        t3 = x * y % p
        t3 = (t3 + t3) % p
        z3 = x * z % p
        z3 = (z3 + z3) % p
        y3 = b * t2 % p
        y3 = (y3 - z3) % p
        x3 = (y3 + y3) % p
        y3 = (x3 + y3) % p
        x3 = (t1 - y3) % p
        y3 = (t1 + y3) % p
        y3 = x3 * y3 % p
        x3 = x3 * t3 % p
        t3 = (t2 + t2) % p
        t2 = (t2 + t3) % p
        z3 = b * z3 % p
        z3 = (z3 - t2) % p
        z3 = (z3 - t0) % p
        t3 = (z3 + z3) % p
        z3 = (z3 + t3) % p
        t3 = (t0 + t0) % p
        t0 = (t3 + t0) % p
        t0 = (t0 - t2) % p
        t0 = t0 * z3 % p
        y3 = (y3 + t0) % p
        t0 = y * z % p
        t0 = (t0 + t0) % p
        z3 = t0 * z3 % p
        x3 = (x3 - z3) % p
        z3 = t0 * t1 % p
        z3 = (z3 + z3) % p
        z3 = (z3 + z3) % p
        return WeierstrassPoint(self.group, x3, y3, z3)

    def add(self, pt):
        # weier.ts:176-230 (RCB15 Alg. 4, a = -3)
        self.group.is_compat_point(pt)
        x1, y1, z1 = self.x, self.y, self.z
        x2, y2, z2 = pt.x, pt.y, pt.z
        p, b = self.group.p, self.group.b
        t0 = x1 * x2 % p
        t1 = y1 * y2 % p
        t2 = z1 * z2 % p
        t3 = (x1 + y1) % p
        t4 = (x2 + y2) % p
        t3 = t3 * t4 % p
        t4 = (t0 + t1) % p
        t3 = (t3 - t4) % p
        t4 = (y1 + z1) % p
        x3 = (y2 + z2) % p
        t4 = t4 * x3 % p
        x3 = (t1 + t2) % p
        t4 = (t4 - x3) % p
        x3 = (x1 + z1) % p
        y3 = (x2 + z2) % p
        x3 = x3 * y3 % p
        y3 = (t0 + t2) % p
        y3 = (x3 - y3) % p
        z3 = b * t2 % p
        x3 = (y3 - z3) % p
        z3 = (x3 + x3) % p
        x3 = (x3 + z3) % p
        z3 = (t1 - x3) % p
        x3 = (t1 + x3) % p
        y3 = b * y3 % p
        t1 = (t2 + t2) % p
        t2 = (t1 + t2) % p
        y3 = (y3 - t2) % p
        y3 = (y3 - t0) % p
        t1 = (y3 + y3) % p
        y3 = (t1 + y3) % p
        t1 = (t0 + t0) % p
        t0 = (t1 + t0) % p
        t0 = (t0 - t2) % p
        t1 = t4 * y3 % p
        t2 = t0 * y3 % p
        y3 = x3 * z3 % p
        y3 = (y3 + t2) % p
        x3 = t3 * x3 % p
        x3 = (x3 - t1) % p
        z3 = t4 * z3 % p
        t1 = t3 * t0 % p
        z3 = (z3 + t1) % p
        return WeierstrassPoint(self.group, x3, y3, z3)

    def to_affine(self):
        # weier.ts:231-243 (mutates in place; returns False for the identity)
        if self.is_identity():
            self.y = 1
            return False
        STATS['inv'] += 1
        p = self.group.p
        zinv = inv_mod(self.z, p)
        x = pos_mod(self.x * zinv, p)
        y = pos_mod(self.y * zinv, p)
        self.x, self.y, self.z = x, y, 1
        return (x, y)

    def to_bytes(self) -> bytes:
        # weier.ts:244-255: identity -> single 0x00 byte
        c = self.to_affine()
        if not c:
            return b'\x00'
        cs = self.group.size_field_bytes()
        return b'\x04' + to_bytes(c[0], cs) + to_bytes(c[1], cs)


# ----------------------------------------------------------------------------
# Twisted Edwards, extended coordinates (src/curves/edwards.ts)
# ----------------------------------------------------------------------------
class TEdwards(Group):
    def __init__(self, name, p, a, d, order, gen):
        super().__init__(name, p, order)
        verify_pos_range(a, p)
        verify_pos_range(d, p)
        verify_pos_range(gen[0], p)
        verify_pos_range(gen[1], p)
        self.a, self.d, self.gen = a, d, gen
        if not self.is_on_group(self.generator()):
            raise ValueError('generator not on group')

    def identity(self):
        # edwards.ts:46-48: t = x*y = 0, z = 1
        return TEdwardsPoint(self, 0, 1)

    def generator(self):
        return TEdwardsPoint(self, self.gen[0], self.gen[1], pos_mod(self.gen[0] * self.gen[1], self.p), 1)

    def is_on_group(self, pt) -> bool:
        # edwards.ts:52-65
        p, a, d = self.p, self.a, self.d
        x, y, t, z = pt.x, pt.y, pt.t, pt.z
        x2, y2, t2, z2 = x * x % p, y * y % p, t * t % p, z * z % p
        l0 = (a * x2 + y2) % p
        r0 = (z2 + d * t2) % p
        l1 = x * y % p
        r1 = z * t % p
        return self.eq(pt.group) and pos_mod(l0 - r0, p) == 0 and pos_mod(l1 - r1, p) == 0

    def deserialize_point(self, b: bytes):
        # edwards.ts:70-86
        b = bytes(b)
        if len(b) == self.size_point_bytes() and b[0] == 0x04:
            cs = self.size_field_bytes()
            x = from_bytes(b[1:1 + cs])
            y = from_bytes(b[1 + cs:])
            verify_pos_range(x, self.p)
            verify_pos_range(y, self.p)
            t = pos_mod(x * y, self.p)
            pt = TEdwardsPoint(self, x, y, t, 1)
            if not self.is_on_group(pt):
                raise ValueError(f'point not on TEdwards group: {self.name} ')
            return pt
        raise ValueError('error deserializing TEdwardsPoint')


class TEdwardsPoint(Point):
    __slots__ = ('group', 'x', 'y', 't', 'z')

    def __init__(self, g, x, y, t=None, z=None):
        self.group, self.x, self.y = g, x, y
        self.t = t if t is not None else x * y
        self.z = z if z is not None else 1

    def is_identity(self):
        # edwards.ts:117-125
        return self.x == 0 and self.y != 0 and self.t == 0 and self.z != 0 and self.y == self.z

    def eq(self, pt):
        p = self.group.p
        return (self.group.eq(pt.group)
                and pos_mod(self.x * pt.z, p) == pos_mod(pt.x * self.z, p)
                and pos_mod(self.y * pt.z, p) == pos_mod(pt.y * self.z, p))

    def neg(self):
        p = self.group.p
        return TEdwardsPoint(self.group, pos_mod(-self.x, p), self.y, pos_mod(-self.t, p), self.z)

    def dbl(self):
        # edwards.ts:141-160 (Hisil et al. 2008, section 3.3)
        x, y, z = self.x, self.y, self.z
        p, a = self.group.p, self.group.a
        A = x * x % p
        B = y * y % p
        C = (z * z * 2) % p
        D = a * A % p
        EE = x + y
        E = (EE * EE - A - B) % p
        G = (D + B) % p
        F = (G - C) % p
        H = (D - B) % p
        return TEdwardsPoint(self.group, pos_mod(E * F, p), pos_mod(G * H, p), pos_mod(E * H, p), pos_mod(F * G, p))

    def add(self, pt):
        # edwards.ts:161-183 (Hisil et al. 2008, section 3.1, unified)
        self.group.is_compat_point(pt)
        x1, y1, t1, z1 = self.x, self.y, self.t, self.z
        x2, y2, t2, z2 = pt.x, pt.y, pt.t, pt.z
        p, a, d = self.group.p, self.group.a, self.group.d
        A = x1 * x2 % p
        B = y1 * y2 % p
        C = d * t1 * t2 % p
        D = z1 * z2 % p
        E = (((x1 + y1) % p) * ((x2 + y2) % p) - A - B) % p
        F = (D - C) % p
        G = (D + C) % p
        H = (B - a * A) % p
        return TEdwardsPoint(self.group, pos_mod(E * F, p), pos_mod(G * H, p), pos_mod(E * H, p), pos_mod(F * G, p))

    def to_affine(self):
        # edwards.ts:184-193 (mutates in place)
        STATS['inv'] += 1
        p = self.group.p
        zinv = inv_mod(self.z, p)
        x = pos_mod(self.x * zinv, p)
        y = pos_mod(self.y * zinv, p)
        self.x, self.y, self.t, self.z = x, y, pos_mod(x * y, p), 1
        return (x, y)

    def to_bytes(self) -> bytes:
        # edwards.ts:195-203: 0x04 || x || y, 33-byte coordinates
        x, y = self.to_affine()
        cs = self.group.size_field_bytes()
        return b'\x04' + to_bytes(x, cs) + to_bytes(y, cs)


# ----------------------------------------------------------------------------
# src/curves/instances.ts:22-56
# ----------------------------------------------------------------------------
p256 = WeierstrassGroup(
    'p256',
    0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff,
    0xffffffff00000001000000000000000000000000fffffffffffffffffffffffc,
    0x5ac635d8aa3a93e7b3ebbd55769886bc651d06b0cc53b0f63bce3c3e27d2604b,
    0xffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551,
    (0x6b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296,
     0x4fe342e2fe1a7f9b8ee7eb4a7c0f9e162bce33576b315ececbb6406837bf51f5),
)

war256 = WeierstrassGroup(
    'war256',
    0xffffffff0000000100000000000000017e72b42b30e7317793135661b1c4b117,
    0xffffffff0000000100000000000000017e72b42b30e7317793135661b1c4b114,
    0xb441071b12f4a0366fb552f8e21ed4ac36b06aceeb354224863e60f20219fc56,
    0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff,
    (0x3, 0x5a6dd32df58708e64e97345cbe66600decd9d538a351bb3c30b4954925b1f02d),
)

tomEdwards256 = TEdwards(
    'tomEdwards256',
    0x3fffffffc000000040000000000000002ae382c7957cc4ff9713c3d82bc47d3af,
    0x1abce3fd8e1d7a21252515332a512e09d4249bd5b1ec35e316c02254fe8cedf5d,
    0x051781d9823abde00ec99295ba542c8b1401874bcbeb9e9c861174c7bca6a02aa,
    0x0ffffffff00000001000000000000000000000000ffffffffffffffffffffffff,
    (0x7907055d0a7d4abc3eafdc25d431d9659fbe007ee2d8ddc4e906206ea9ba4fdb,
     0xbe231cb9f9bf18319c9f081141559b0a33dddccd2221f0464a9cd57081b01a01),
)

ALL_GROUPS = [p256, war256, tomEdwards256]
