"""Restatement of src/bignum/big.ts (reference lines cited per function).

TEST INFRASTRUCTURE (oracle) — see oracle/__init__.py.
"""
from __future__ import annotations


class TapeExhausted(Exception):
    pass


class Tape:
    """Stand-in for WebCrypto `crypto.getRandomValues` (big.ts:175).

    A byte stream consumed front to back; `fill(n)` returns the next n bytes.
    `log` records (offset, nbytes) of every call so tests can check the draw
    order contract (323 + 40*Z + 5n draws of 32 bytes).
    """

    def __init__(self, data: bytes):
        self.data = bytes(data)
        self.pos = 0
        self.calls = 0

    def fill(self, n: int) -> bytes:
        if self.pos + n > len(self.data):
            raise TapeExhausted(f"tape exhausted at {self.pos}+{n} > {len(self.data)}")
        out = self.data[self.pos:self.pos + n]
        self.pos += n
        self.calls += 1
        return out


class OsTape:
    """Tape backed by os.urandom (for randomised property tests)."""

    def __init__(self):
        import os
        self._os = os
        self.pos = 0
        self.calls = 0
        self.record = bytearray()

    def fill(self, n: int) -> bytes:
        b = self._os.urandom(n)
        self.record += b
        self.pos += n
        self.calls += 1
        return b


def verify_pos_range(a: int, n: int) -> bool:
    # big.ts:17-22
    if not (0 <= a < n):
        raise ValueError('a not in range')
    return True


def bit_len(n: int) -> int:
    # big.ts:23-25  n.toString(2).length  (0 -> "0" -> 1; negative counts the '-')
    if n == 0:
        return 1
    if n < 0:
        return (-n).bit_length() + 1
    return n.bit_length()


def byte_len(n: int) -> int:
    # big.ts:26-28
    return -(-bit_len(n) // 8)


def is_odd(n: int) -> bool:
    # big.ts:29-31  (JS % keeps sign: -3 % 2 === -1 !== 1; only used on n >= 0)
    return n >= 0 and (n & 1) == 1


def pos_mod(n: int, p: int) -> int:
    # big.ts:36-42
    return n % p  # Python % with p > 0 is already in [0, p)


def exp_mod(n: int, e: int, p: int) -> int:
    # big.ts:44-59 (square-and-multiply; result identical to pow for n >= 0)
    if e < 0:
        raise ValueError('neg expo')
    r, q, k = 1, n, e
    while k > 0:
        if k & 1:
            r = (r * q) % p
        q = (q * q) % p
        k >>= 1
    return r


def _extended_euclid(X: int, Y: int):
    # big.ts:80-110.  JS BigInt `/` truncates toward zero; inputs are >= 0 here
    # so floor == trunc.
    a, b, c, d, x, y = 1, 0, 0, 1, X, Y
    while y != 0:
        q = x // y
        a = a - c * q
        b = b - d * q
        x = x - q * y
        x, y = y, x
        a, c = c, a
        b, d = d, b
    return x, a, b


def inv_euclid(t: int, N: int) -> int:
    # big.ts:112-119
    _, inv, _ = _extended_euclid(t, N)
    if inv < 0:
        inv += N
    return inv


def inv_mod(n: int, p: int) -> int:
    # big.ts:76-78
    return inv_euclid(n, p)


def to_bytes(n: int, length: int) -> bytes:
    # big.ts:121-134 (big-endian, must fit)
    if not (length > 0 and 0 <= n < (1 << (8 * length))):
        raise ValueError("number doesn't fit in array")
    return n.to_bytes(length, 'big')


def from_bytes(a: bytes) -> int:
    # big.ts:161-168
    return int.from_bytes(bytes(a), 'big')


def rnd(n: int, tape) -> int:
    # big.ts:171-181: rejection sampling on byteLen(n) fresh random bytes
    ln = byte_len(n)
    while True:
        ret = from_bytes(tape.fill(ln))
        if ret < n:
            return ret


def rnd_range(lo: int, hi: int, tape) -> int:
    # big.ts:183-185
    return rnd(hi - lo + 1, tape) + lo
