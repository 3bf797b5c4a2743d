"""Quick prove-throughput probe (development aid; bench.py is the contract benchmark)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from zkp_ecdsa_b200 import api, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
t0 = time.time()
eng = api.Engine(0)
params = eng.generate_params_list(rnd=synth.params_rnd(0))
t1 = time.time()
wl = synth.Workload(B, N, seed=0)
tape = synth.random_tape(B, eng.lib.prove_tape_len(N), seed=1)
t2 = time.time()
print(json.dumps({'init_s': t1 - t0, 'workload_s': t2 - t1}))
for r in range(reps):
    t = time.time()
    res = eng.prove_batch(params, wl.msg_hash, wl.sig, wl.pk, wl.which, wl.ring, tape)
    dt = time.time() - t
    print(json.dumps({'B': B, 'N': N, 'prove_s': dt, 'proofs_per_s': B / dt, 'bad': int((res.status != 0).sum()),
                      'launches': eng.lib.launch_count()}))
