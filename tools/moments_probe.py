#!/usr/bin/env python
"""Block-sum pass (ms_bitmap_reduce: moments kernel + column scan) for every power configuration, at several range sizes.
Run twice (PSS_MOMENTS_WIDE=0 / 1, PSS_MOMENTS2_WIDE=0 / 1) to compare the two table layouts on one box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prime-square-sum_b200"))
from pss_b200 import _lib

h = _lib.Handle(-1)
h.set_kernel_events(2)
print("PSS_MOMENTS_WIDE", os.environ.get("PSS_MOMENTS_WIDE", "1"), "PSS_MOMENTS2_WIDE", os.environ.get("PSS_MOMENTS2_WIDE", "1"))
for hi in (86028121, 179424673, 982451653, 22801763490):
    row = []
    for multi in (False, True):
        for k in range(1, 6):
            if multi and k == 1:
                continue
            best = None
            for _ in range(4):
                h.stats_reset()
                h.slice_sieve_sums(0, hi, k, multi)
                s = h.stats()
                best = s["ms_bitmap_reduce"] if best is None else min(best, s["ms_bitmap_reduce"])
            row.append(f"{'m' if multi else 's'}{k}={best * 1e3:.1f}us")
    print(f"hi={hi:.3e}", " ".join(row), flush=True)
h.close()
