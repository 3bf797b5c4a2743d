#!/usr/bin/env python3
"""Export the golden fixtures (zkattest_v{1,2}.npz) as JSON for oracle/gen_ts_vectors.mjs (run from the repo root).

The TypeScript reference cannot run in this image (no node).  Where node >= 24 exists, a maintainer runs
  node oracle/gen_ts_vectors.mjs /path/to/zkp-ecdsa tests/golden/ts_inputs.json tests/golden
which feeds these exact inputs and randomness tapes to the REAL proveSignatureList / verifySignatureList
(crypto.getRandomValues mocked with the tape) and writes tests/golden/ts_<tag>.bin; tests/test_ts_vectors.py then
compares them with the oracle-made fixtures — the pin of the oracle to the TypeScript proof bytes."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from zkp_ecdsa_b200 import verify_tape as VT  # noqa: E402


def main():
    G = {}
    for f in ('zkattest_v1.npz', 'zkattest_v2.npz'):
        z = np.load(os.path.join(ROOT, 'tests', 'golden', f))
        G.update({k: z[k] for k in z.files})
    out = {}
    for tag in ('a', 'b', 'c', 'd'):
        B, N, sec, seed = (int(v) for v in G[f'{tag}_meta'])
        hx = lambda a: bytes(np.ascontiguousarray(a)).hex()   # noqa: E731
        out[tag] = {
            'B': B, 'N': N, 'sec_level': sec,
            'params_rnd': hx(G[f'{tag}_params_rnd']), 'h_nist': hx(G[f'{tag}_h_nist']), 'h_proof': hx(G[f'{tag}_h_proof']),
            'msg_hash': [hx(G[f'{tag}_msg_hash'][b]) for b in range(B)], 'sig': [hx(G[f'{tag}_sig'][b]) for b in range(B)],
            'pk': [hx(G[f'{tag}_pk'][b]) for b in range(B)], 'which': [int(w) for w in G[f'{tag}_which']],
            'ring': [hx(G[f'{tag}_ring'][i]) for i in range(N)],
            # prover tape: 32-byte draws in rnd() call order, already below their moduli (no rejection fires)
            'tape': [hx(G[f'{tag}_tape'][b]) for b in range(B)],
            # verifier tape as the byte stream rnd() consumes: GK drains, 1 byte per generateIndices draw, exp drains
            'vtape': [VT.oracle_stream(G[f'{tag}_vtape'][b].tobytes(), N, sec).hex() for b in range(B)],
            'proof_len': [int(v) for v in G[f'{tag}_proof_len']], 'verdict': [int(v) for v in G[f'{tag}_verdict']],
        }
    p = os.path.join(ROOT, 'tests', 'golden', 'ts_inputs.json')
    json.dump(out, open(p, 'w'))
    print(p, os.path.getsize(p), 'bytes')


if __name__ == '__main__':
    main()
