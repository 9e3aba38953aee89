#!/usr/bin/env python
"""Device span of a query whose target IS a partial sum (the crossing block has to be walked) against one that crosses nothing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prime-square-sum_b200"))
from pss_b200 import _lib, search

h = _lib.Handle(-1)
h.set_kernel_events(2)
BIG = 37005443752611483714216385166550857181329086284892731078593232926279977894581784762614450464857290
for n_max, n_hit in [(10 ** 6, 7), (10 ** 6, 300), (10 ** 6, 123456), (10 ** 6, 999999), (10 ** 8, 54321987)]:
    t = search.search(1, n_hit, power=2, handle=h).powers[0].final_sum
    for name, target in (("hit", t), ("miss", BIG)):
        best = None
        for _ in range(5):
            h.stats_reset()
            o = search.search(target, n_max, power=2, handle=h)
            s = h.stats()
            best = (s["ms_span"], s["ms_bitmap_compare"]) if best is None or s["ms_span"] < best[0] else best
        print(f"n_max={n_max:.0e} n_hit={n_hit} {name}: first={o.first()} span={best[0]:.4f} compare={best[1]:.4f}", flush=True)
h.close()
