#!/usr/bin/env python
"""ms_bitmap_compare of the triangular-predicate query (head pass + scan kernel) for several n_max / m_max;
run with PSS_TRI_HEAD=0 / 1 to compare."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prime-square-sum_b200"))
from pss_b200 import _lib

h = _lib.Handle(-1)
h.set_kernel_events(2)
print("PSS_TRI_HEAD", os.environ.get("PSS_TRI_HEAD", "1"))
for n_max, m_max in [(10 ** 8, 10 ** 7), (10 ** 6, 10 ** 7), (10 ** 8, 10 ** 3), (10 ** 8, 2 ** 28), (10 ** 4, 10 ** 7)]:
    best = None
    for _ in range(4):
        h.stats_reset()
        pairs, _ = h.search_tri(1, n_max, 2, 1, m_max)
        s = h.stats()
        best = s["ms_bitmap_compare"] if best is None else min(best, s["ms_bitmap_compare"])
    print(f"n_max={n_max:.0e} m_max={m_max:.1e} pairs={sorted(pairs)} ms_compare={best:.4f} ms_sieve={s['ms_sieve']:.4f}", flush=True)
h.close()
