#!/usr/bin/env python
"""Host <-> device copy rates of this box (pinned H2D / D2H, both at once, pageable) and host memcpy rates -- the
denominators of the gpu_transport_photons drop-in (bench.py e2e_aos_dropin)."""
import threading
import time

import numpy as np
import torch

N = 1 << 30  # bytes
dev = torch.device("cuda:0")
d_a = torch.empty(N, dtype=torch.uint8, device=dev)
d_b = torch.empty(N, dtype=torch.uint8, device=dev)
h_pin_a = torch.empty(N, dtype=torch.uint8).pin_memory()
h_pin_b = torch.empty(N, dtype=torch.uint8).pin_memory()
h_page = torch.empty(N, dtype=torch.uint8)
h_page.fill_(1)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return N / best / 1e9


print(f"pinned   H2D           {timed(lambda: d_a.copy_(h_pin_a, non_blocking=True)):7.1f} GB/s")
print(f"pinned   D2H           {timed(lambda: h_pin_b.copy_(d_b, non_blocking=True)):7.1f} GB/s")


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_pin_a, non_blocking=True)
    with torch.cuda.stream(s2):
        h_pin_b.copy_(d_b, non_blocking=True)


print(f"pinned   H2D + D2H     {timed(both):7.1f} GB/s per direction")
print(f"pageable H2D           {timed(lambda: d_a.copy_(h_page)):7.1f} GB/s")
print(f"pageable D2H           {timed(lambda: h_page.copy_(d_b)):7.1f} GB/s")

src = np.ones(N, dtype=np.uint8)
dst = np.empty(N, dtype=np.uint8)
dst[:] = 0
for nt in (1, 2, 4, 8):
    def work(k):
        lo, hi = k * N // nt, (k + 1) * N // nt
        np.copyto(dst[lo:hi], src[lo:hi])
    best = 1e9
    for _ in range(3):
        th = [threading.Thread(target=work, args=(k,)) for k in range(nt)]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        best = min(best, time.perf_counter() - t0)
    print(f"host memcpy, {nt} threads {N / best / 1e9:7.1f} GB/s")
