#!/usr/bin/env python3
"""One prove pass + one verify pass of a bench workload between cudaProfilerStart/Stop — the command the
ncu captures under profiles/ are taken from (`ncu --profile-from-start off ... python tools/profile_step.py`).

  python tools/profile_step.py [--workload config2] [--batch B] [--ring N] [--what prove,verify]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    from bench import SEC_LEVEL, WORKLOADS
    from zkp_ecdsa_b200 import api, synth
    from zkp_ecdsa_b200 import verify_tape as VT
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='config2')
    ap.add_argument('--batch', type=int, default=0)
    ap.add_argument('--ring', type=int, default=0)
    ap.add_argument('--what', default='prove,verify')
    ap.add_argument('--warm', type=int, default=2)
    a = ap.parse_args()
    B, N = WORKLOADS[a.workload]
    B, N = a.batch or B, a.ring or N
    dev = torch.device('cuda', 0)
    eng = api.Engine(device=0)
    L = eng.lib
    params = eng.generate_params_list(SEC_LEVEL, rnd=synth.params_rnd(0))
    wl = synth.Workload(B, N, seed=100)
    ts, ps, vts = L.prove_tape_len(N), (L.proof_max_len(N) + 15) & ~15, L.verify_tape_len(N)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)   # noqa: E731
    d = {'msg': t(wl.msg_hash), 'sig': t(wl.sig), 'pk': t(wl.pk), 'which': t(wl.which.view(np.uint8)), 'ring': t(wl.ring)}
    tape = t(synth.random_tape(B, ts, seed=200))
    vt = t(VT.random_verify_tape(B, vts, N, SEC_LEVEL, seed=300))
    proofs = torch.zeros((B, ps), dtype=torch.uint8, device=dev)
    plen = torch.zeros(B, dtype=torch.int32, device=dev)
    st = torch.zeros(B, dtype=torch.int32, device=dev)
    ok = torch.zeros(B, dtype=torch.uint8, device=dev)

    def prove():
        L.prove_batch(params.handle, B, d['msg'].data_ptr(), d['sig'].data_ptr(), d['pk'].data_ptr(), d['which'].data_ptr(),
                      d['ring'].data_ptr(), N, tape.data_ptr(), ts, proofs.data_ptr(), ps, plen.data_ptr(), st.data_ptr())

    def verify():
        L.verify_batch(params.handle, B, d['msg'].data_ptr(), d['ring'].data_ptr(), N, proofs.data_ptr(), ps, plen.data_ptr(),
                       vt.data_ptr(), vts, ok.data_ptr(), st.data_ptr())
    for _ in range(a.warm):
        prove()
        verify()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    if 'prove' in a.what:
        prove()
    if 'verify' in a.what:
        verify()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    assert bool((ok == 1).all().item())
    print('profile_step ok', B, N, 'launches', L.launch_count())


if __name__ == '__main__':
    main()
