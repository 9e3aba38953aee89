#!/usr/bin/env python
"""ms_compact (k_compact_lut) of the materialised C3 query; run with PSS_COMPACT_PACK=0 / 1 to compare the staging forms."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prime-square-sum_b200"))
from pss_b200 import _lib, search

T = 37005443752611483714216385166550857181329086284892731078593232926279977894581784762614450464857290
S2 = 168072405102068540986037048787
h = _lib.Handle(-1)
h.set_kernel_events(2)
h.set_search_path(True)
best = None
for _ in range(4):
    h.stats_reset()
    o = search.search(T, 10 ** 9, power=2, handle=h)
    assert o.first() is None and o.powers[0].final_sum == S2
    s = h.stats()
    best = s["ms_compact"] if best is None else min(best, s["ms_compact"])
print("PSS_COMPACT_PACK", os.environ.get("PSS_COMPACT_PACK", "1"), "ms_compact", round(best, 4), "ms_power_scan", round(s["ms_power_scan"], 4))
h.close()
