#!/usr/bin/env python
"""One launch of every kernel of the three query paths at full size (n = 1e9), for `ncu --set full -k regex:...`:
   C3 fused (k_sieve, k_bitmap_moments2, k_scan_columns, k_bitmap_scan<2,compare>), C5 fused all powers
   (k_bitmap_moments_tab<5,true>, k_bitmap_scan<5,true,compare>) and C3 materialised (k_compact_lut, k_power_tiles,
   k_power_walk).  Results are checked so that a profile is never taken of a wrong run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prime-square-sum_b200"))
from pss_b200 import _lib, search

T = 37005443752611483714216385166550857181329086284892731078593232926279977894581784762614450464857290
S2 = 168072405102068540986037048787
h = _lib.Handle(-1)
n = int(os.environ.get("PSS_PROFILE_N", 10 ** 9))
for rep in range(int(os.environ.get("PSS_PROFILE_REPS", 2))):      # first pass = warm-up (allocations, pattern build)
    o = search.search(T, n, power=2, handle=h)
    assert o.first() is None and (n != 10 ** 9 or o.powers[0].final_sum == S2)
    o = search.search(T, n, power=5, n_min=n // 20, all_powers=True, handle=h)
    assert o.first() is None and (n != 10 ** 9 or o.powers[1].final_sum == S2)
    h.set_search_path(True)
    o = search.search(T, n, power=2, handle=h)
    assert o.first() is None and (n != 10 ** 9 or o.powers[0].final_sum == S2)
    h.set_search_path(False)
h.close()
print("profile driver ok")
