"""Throughput of the war256 build (libzkattest_war256.so: ProofGroup = war256, instances.ts:34-41) on config2's shape,
device-resident buffers, CUDA events on the library's stream.  bench.py is the contract benchmark (tomEdwards256, the
group generateParamsList builds); this prints one JSON line for the other legal ProofGroup.
usage: python tools/bench_war256.py [batch] [ring] [steps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from zkp_ecdsa_b200 import api, synth
from zkp_ecdsa_b200 import verify_tape as VT

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
group = os.environ.get('ZKA_BENCH_GROUP', 'war256')
dev = torch.device('cuda:0')
eng = api.Engine(0, proof_group=group)
L = eng.lib
params = eng.generate_params_list(80, rnd=synth.params_rnd(0))
wl = synth.Workload(B, N, seed=100)
ts, ps = L.prove_tape_len(N, 80), (L.proof_max_len(N, 80) + 15) & ~15
d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
     {'msg': wl.msg_hash, 'sig': wl.sig, 'pk': wl.pk, 'which': wl.which.view(np.uint8), 'ring': wl.ring}.items()}
tape_d = torch.from_numpy(synth.random_tape(B, ts, seed=200)).to(dev)
proofs_d = torch.zeros((B, ps), dtype=torch.uint8, device=dev)
plen_d = torch.zeros(B, dtype=torch.int32, device=dev)
stat_d = torch.zeros(B, dtype=torch.int32, device=dev)
vts = L.verify_tape_len(N, 80)
vt_d = torch.from_numpy(VT.random_verify_tape(B, vts, N, 80, seed=300)).to(dev)
ok_d = torch.zeros(B, dtype=torch.uint8, device=dev)
vst_d = torch.zeros(B, dtype=torch.int32, device=dev)
stream = torch.cuda.ExternalStream(L.stream_ptr(), device=dev)


def prove():
    L.prove_batch(params.handle, B, d['msg'].data_ptr(), d['sig'].data_ptr(), d['pk'].data_ptr(), d['which'].data_ptr(),
                  d['ring'].data_ptr(), N, tape_d.data_ptr(), ts, proofs_d.data_ptr(), ps, plen_d.data_ptr(), stat_d.data_ptr())


def verify():
    L.verify_batch(params.handle, B, d['msg'].data_ptr(), d['ring'].data_ptr(), N, proofs_d.data_ptr(), ps, plen_d.data_ptr(),
                   vt_d.data_ptr(), vts, ok_d.data_ptr(), vst_d.data_ptr())


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / steps


p_ms = timed(prove)
assert int((stat_d != 0).sum().item()) == 0
v_ms = timed(verify)
assert bool((ok_d == 1).all().item()) and bool((vst_d == 0).all().item())
L.set_option('lanes', 1)
L.profile_reset()
L.set_profiling(True)
prove()
L.set_profiling(False)
prof = L.profile()
top = sorted(((k.replace('zk::', ''), round(v['ms'], 3)) for k, v in prof.items()), key=lambda kv: -kv[1])[:8]
print(json.dumps({'proof_group': L.group, 'point_bytes': L.wp, 'scalar_bytes': L.ws, 'batch': B, 'ring': N, 'sec_level': 80,
                  'proofs_per_s': B / (p_ms * 1e-3), 'prove_ms_per_step': p_ms, 'verifies_per_s': B / (v_ms * 1e-3),
                  'verify_ms_per_step': v_ms, 'mean_proof_bytes': float(plen_d.float().mean().item()),
                  'all_accepted': True, 'agg_chunks_accepted': L.stat('agg_pass'), 'top_prove_kernels_ms': top,
                  'config': L.config(), 'data': 'synthetic', 'timing': 'CUDA events on the library stream, device-resident buffers'}))
