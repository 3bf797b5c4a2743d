#!/usr/bin/env python3
"""Generate tests/golden/zkattest_v1.npz and zkattest_v2.npz from the ORACLE (run from the repo root).

The reference has no proof-byte vectors and cannot run here, so the fixtures
pin (a) the oracle against its own regressions and (b) the CUDA path against the oracle on
fixed inputs.  Inputs are deterministic (zkp_ecdsa_b200/synth.py); re-running reproduces the file.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import flat  # noqa: E402
from oracle import zkattest as OZ  # noqa: E402
from oracle.big import Tape  # noqa: E402
from zkp_ecdsa_b200 import synth, verify_tape as VT  # noqa: E402


def case(tag, B, N, sec, seed):
    rnd = synth.params_rnd(seed)
    po = OZ.generate_params_list(Tape(rnd), sec)
    wl = synth.Workload(B=B, N=N, seed=seed)
    n = VT.ceil_log2(N)
    tape = synth.random_tape(B, 32 * (3 + 4 * sec + 40 * sec + 5 * n), seed=seed + 1)
    vts = VT.verify_tape_len(N)
    vtape = VT.random_verify_tape(B, vts, N, sec, seed=seed + 2)
    proofs, verdicts = [], []
    for b in range(B):
        pr = OZ.prove_signature_list(po, wl.msg_hash[b].tobytes(), wl.sig[b].tobytes(), wl.pk[b].tobytes(),
                                     int(wl.which[b]), wl.ring_ints(), Tape(tape[b].tobytes()))
        data = flat.ser_proof(pr)
        proofs.append(data)
        verdicts.append(OZ.verify_signature_list(po, wl.msg_hash[b].tobytes(), wl.ring_ints(), flat.de_proof(data, sec),
                                                 Tape(VT.oracle_stream(vtape[b].tobytes(), N, sec))))
    stride = max(len(p) for p in proofs)
    arr = np.zeros((B, stride), np.uint8)
    for b, p in enumerate(proofs):
        arr[b, :len(p)] = np.frombuffer(p, np.uint8)
    return {f'{tag}_params_rnd': np.frombuffer(rnd, np.uint8), f'{tag}_h_nist': np.frombuffer(po.NistGroup.h.to_bytes(), np.uint8),
            f'{tag}_h_proof': np.frombuffer(po.ProofGroup.h.to_bytes(), np.uint8),
            f'{tag}_meta': np.array([B, N, sec, seed], np.int64), f'{tag}_msg_hash': wl.msg_hash, f'{tag}_sig': wl.sig,
            f'{tag}_pk': wl.pk, f'{tag}_which': wl.which, f'{tag}_ring': wl.ring, f'{tag}_tape': tape,
            f'{tag}_vtape': vtape, f'{tag}_proofs': arr, f'{tag}_proof_len': np.array([len(p) for p in proofs], np.uint32),
            f'{tag}_verdict': np.array(verdicts, np.uint8)}


def main():
    out = {}
    out.update(case('a', B=1, N=6, sec=80, seed=1001))    # the reference's own test shape (ring of 6, test/zkpAttestList.test.ts:36)
    out.update(case('b', B=3, N=17, sec=20, seed=1002))   # ragged ring (padding 17 -> 32), smaller SecLevel
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'zkattest_v1.npz'), **out)
    print({k: v.shape for k, v in out.items()})
    out2 = {}
    out2.update(case('c', B=1, N=2100, sec=20, seed=1003))  # ring above 1024 entries: block-parallel ring sums
    out2.update(case('d', B=4, N=3, sec=20, seed=1004))     # repeated signers: shared per-key tables
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'zkattest_v2.npz'), **out2)
    print({k: v.shape for k, v in out2.items()})


if __name__ == '__main__':
    main()
