"""Pin of the oracle to the TypeScript reference.  oracle/gen_ts_vectors.mjs (run where node >= 24 exists) drives the
REAL proveSignatureList / verifySignatureList with the committed inputs and tapes (tests/golden/ts_inputs.json) and
writes tests/golden/ts_<tag>.bin.  When those files are present they must equal the oracle-made fixtures byte for
byte; until then the proof bytes stay "parity unpinned" w.r.t. TypeScript and these tests skip."""
import json
import os
import struct

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(__file__), 'golden')


def test_ts_inputs_match_the_npz_fixtures():
    """The JSON handed to node is exactly the committed fixtures (export_ts_inputs.py is reproducible)."""
    from zkp_ecdsa_b200 import verify_tape as VT
    J = json.load(open(os.path.join(HERE, 'ts_inputs.json')))
    G = {}
    for f in ('zkattest_v1.npz', 'zkattest_v2.npz'):
        z = np.load(os.path.join(HERE, f))
        G.update({k: z[k] for k in z.files})
    for tag, c in J.items():
        B, N, sec, _ = (int(v) for v in G[f'{tag}_meta'])
        assert (c['B'], c['N'], c['sec_level']) == (B, N, sec)
        for b in range(B):
            assert bytes.fromhex(c['tape'][b]) == G[f'{tag}_tape'][b].tobytes()
            assert bytes.fromhex(c['vtape'][b]) == VT.oracle_stream(G[f'{tag}_vtape'][b].tobytes(), N, sec)
            assert bytes.fromhex(c['pk'][b]) == G[f'{tag}_pk'][b].tobytes()
        assert [bytes.fromhex(x) for x in c['ring']] == [G[f'{tag}_ring'][i].tobytes() for i in range(N)]


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd'])
def test_typescript_proof_bytes_equal_oracle(tag):
    p = os.path.join(HERE, f'ts_{tag}.bin')
    if not os.path.exists(p):
        pytest.skip('tests/golden/ts_%s.bin not generated yet (needs node: oracle/gen_ts_vectors.mjs)' % tag)
    z = np.load(os.path.join(HERE, 'zkattest_v1.npz' if tag in 'ab' else 'zkattest_v2.npz'))
    data = open(p, 'rb').read()
    o = 0
    for b in range(int(z[f'{tag}_meta'][0])):
        (ln,) = struct.unpack_from('<I', data, o)
        o += 4
        assert ln == int(z[f'{tag}_proof_len'][b])
        assert data[o:o + ln] == z[f'{tag}_proofs'][b, :ln].tobytes(), f'TypeScript proof {tag}[{b}] differs from the oracle'
        o += ln
    v = json.load(open(os.path.join(HERE, f'ts_{tag}.verdict.json')))
    assert v['params_ok'] and v['verdicts'] == [int(x) for x in z[f'{tag}_verdict']]
