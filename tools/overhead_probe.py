#!/usr/bin/env python
"""Host-side overhead of a query: wall clock per call through pss_b200.search.search and through the raw ctypes call against the
device span (first launch -> end of the last kernel) and the kernels' own time, for tiny and small ranges (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "prime-square-sum_b200"))
from pss_b200 import _lib, search
h = _lib.default_handle()
T = 37005443752611483714216385166550857181329086284892731078593232926279977894581784762614450464857290
for n in (1000, 10**6):
    for _ in range(20): search.search(T, n, handle=h)
    t0 = time.perf_counter()
    N = 300
    for _ in range(N): o = search.search(T, n, handle=h)
    dt = (time.perf_counter() - t0) / N
    h.stats_reset(); search.search(T, n, handle=h); s = h.stats()
    print(f"n_max={n}: wall per call {dt*1e6:.1f} us, device span {s['ms_span']*1e3:.1f} us, kernels {s['ms_total_device']*1e3:.1f} us")
    # raw C call without outcome conversion
    t0 = time.perf_counter()
    for _ in range(N): r = h.search(1, n, 2, T)
    print(f"   raw handle.search: {(time.perf_counter()-t0)/N*1e6:.1f} us")
for n in (1000, 10**6, 10**7):
    h.stats_reset(); search.search(T, n, handle=h); s = h.stats()
    print(n, {k: round(v * 1e3, 1) for k, v in s.items() if k.startswith("ms_")})
