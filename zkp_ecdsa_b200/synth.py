"""Deterministic synthetic workloads for tests and bench.py.

Everything derives from a SHA-256 counter DRBG keyed by (seed, label).  P-256 key generation
and ECDSA signing use the `cryptography` package (OpenSSL) — inputs only, nothing measured.
"""
from __future__ import annotations

import hashlib

import numpy as np

P256_N = 0xffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551
P256_P = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff  # == tomEdwards256 order


class Drbg:
    def __init__(self, seed: int, label: str):
        self.key = hashlib.sha256(f'zkattest-synth|{seed}|{label}'.encode()).digest()
        self.ctr = 0

    def bytes(self, n: int) -> bytes:
        out = bytearray()
        while len(out) < n:
            out += hashlib.sha256(self.key + self.ctr.to_bytes(8, 'big')).digest()
            self.ctr += 1
        return bytes(out[:n])

    def below(self, m: int) -> int:
        while True:
            v = int.from_bytes(self.bytes(32), 'big')
            if v < m:
                return v


def draw_modulus(k: int, sec_level: int = 80) -> int:
    """Modulus of prover tape draw k (depends only on k; include/zkattest.h)."""
    if k == 0:
        return P256_N
    if k < 3:
        return P256_P
    if k < 3 + 4 * sec_level:
        return P256_N if ((k - 3) % 4) < 2 else P256_P
    return P256_P


def filter_prove_tape(raw: bytes, ndraws: int, sec_level: int = 80) -> bytes:
    """Host-side rnd() rejection loop (big.ts:171-181): consume 32-byte candidates from `raw`
    in order, dropping those >= the modulus of the draw they would feed."""
    out = bytearray()
    pos = 0
    for k in range(ndraws):
        m = draw_modulus(k, sec_level)
        while True:
            cand = raw[pos:pos + 32]
            if len(cand) < 32:
                raise ValueError('raw tape exhausted')
            pos += 32
            if int.from_bytes(cand, 'big') < m:
                out += cand
                break
    return bytes(out)


def random_tape(rows: int, stride: int, seed: int) -> np.ndarray:
    """`rows` tapes of `stride` bytes (multiple of 32) of uniform 32-byte draws, each forced below
    0xffffffff00000000... (< p256.n < p256.p) by resampling the top word (probability 2^-32)."""
    assert stride % 32 == 0
    rng = np.random.Generator(np.random.PCG64(seed))
    t = rng.integers(0, 256, size=(rows, stride), dtype=np.uint8)
    top = t.reshape(rows, stride // 32, 32)[:, :, :4]
    bad = (top == 255).all(axis=2)
    if bad.any():
        t.reshape(rows, stride // 32, 32)[bad, 3] = 0xfe
    return t


class Workload:
    """B signing instances sharing one ring of N key x-coordinates."""

    def __init__(self, B: int, N: int, seed: int = 0, distinct_signers: int | None = None):
        from cryptography.hazmat.primitives import serialization
        from cryptography.hazmat.primitives.asymmetric import ec
        self.B, self.N, self.seed = B, N, seed
        ns = min(B, N) if distinct_signers is None else min(distinct_signers, B, N)
        d = Drbg(seed, 'signers')
        dn = Drbg(seed, 'nonces')
        sk_ints = [d.below(P256_N - 1) + 1 for _ in range(ns)]

        def pub(k):   # k*G as 65 raw bytes (OpenSSL does the scalar multiplication)
            return ec.derive_private_key(k, ec.SECP256R1()).public_key().public_bytes(
                serialization.Encoding.X962, serialization.PublicFormat.UncompressedPoint)
        pks = [pub(k) for k in sk_ints]
        # ring: signer j sits at slot slot[j]; filler entries are arbitrary 256-bit values
        dr = Drbg(seed, 'ring')
        ring = [dr.bytes(32) for _ in range(N)]
        perm = np.random.Generator(np.random.PCG64(seed + 7)).permutation(N)[:ns]
        for j in range(ns):
            ring[int(perm[j])] = pks[j][1:33]
        self.ring = np.frombuffer(b''.join(ring), np.uint8).reshape(N, 32).copy()
        self.msg_hash = np.zeros((B, 32), np.uint8)
        self.sig = np.zeros((B, 64), np.uint8)
        self.pk = np.zeros((B, 65), np.uint8)
        self.which = np.zeros(B, np.uint32)
        for b in range(B):
            j = b % ns
            msg = b'zkattest-bench-%d' % b
            digest = hashlib.sha256(msg).digest()
            # deterministic ECDSA: nonce from the DRBG; r = (kG).x mod n, s = (z + r d)/k
            while True:
                k = dn.below(P256_N - 1) + 1
                r = int.from_bytes(pub(k)[1:33], 'big') % P256_N
                s = pow(k, -1, P256_N) * (int.from_bytes(digest, 'big') + r * sk_ints[j]) % P256_N
                if r and s:
                    break
            self.msg_hash[b] = np.frombuffer(digest, np.uint8)
            self.sig[b] = np.frombuffer(r.to_bytes(32, 'big') + s.to_bytes(32, 'big'), np.uint8)
            self.pk[b] = np.frombuffer(pks[j], np.uint8)
            self.which[b] = int(perm[j])

    def ring_ints(self):
        return [int.from_bytes(self.ring[i].tobytes(), 'big') for i in range(self.N)]


def params_rnd(seed: int = 0) -> bytes:
    d = Drbg(seed, 'params')
    return d.below(P256_N).to_bytes(32, 'big') + d.below(P256_P).to_bytes(32, 'big')
