"""Pinned host <-> device copy bandwidth of this box (what bounds the end-to-end legs of bench.py)."""
import json
import torch

n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
d2 = torch.empty(n, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    s1.synchronize()
    s2.synchronize()


h2d = t(lambda: d.copy_(h, non_blocking=True))
d2h = t(lambda: h2.copy_(d2, non_blocking=True))
import time
both()
t0 = time.perf_counter()
for _ in range(5):
    both()
bi = (time.perf_counter() - t0) / 5 * 1e3
print(json.dumps({'bytes': n, 'h2d_GBps': n / h2d / 1e6, 'd2h_GBps': n / d2h / 1e6, 'bidirectional_each_GBps': n / bi / 1e6}))
