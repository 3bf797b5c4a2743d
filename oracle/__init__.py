"""CPU oracle for the ZKAttest hot path (TEST INFRASTRUCTURE — NOT PRODUCT CODE).

A line-faithful Python-int restatement of cloudflare/zkp-ecdsa v0.2.6
(`/root/reference/src/**`), with `crypto.getRandomValues` replaced by an
injectable randomness tape (oracle/big.py::Tape).  Python `int` has the same
semantics as JS `BigInt` for + - *; every reference `%`/`/` on possibly
negative values is wrapped by posMod in the reference, and we keep that.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product path
(`zkp_ecdsa_b200`) never does and fails loudly if the CUDA library is absent.

PARITY PINNING: the reference ships no proof-byte golden vectors and
cannot be executed here (no node/tsc).  The oracle is pinned against the only
fixed-answer tests the reference has (invEuclid, interpolate, curve
self-checks, `order*G == O`, dblmul == mul+mul, prove->verify round trips;
see tests/test_oracle_*.py).  Proof BYTES are therefore "parity unpinned"
w.r.t. the TypeScript implementation itself: they are pinned only by this
restatement reviewed against the cited lines.
"""
