#!/usr/bin/env python
"""Sweep the sieve tunables on the C3 range and print per-kernel device times (run on the GPU box)."""
import itertools, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prime-square-sum_b200"))
from pss_b200 import _lib

HI = int(os.environ.get("TUNE_HI", 22801763490))
rows = []
grid = list(itertools.product([32, 48, 64, 96, 112, 160, 216], [256, 512, 1024], [4], [256], [8]))
extra = [(96, 512, 2, 256, 8), (96, 512, 3, 256, 8), (96, 512, 0, 256, 8), (96, 512, 4, 128, 8), (96, 512, 4, 512, 8),
         (96, 512, 4, 1 << 30, 8), (96, 512, 4, 256, 4), (96, 512, 4, 256, 16), (96, 512, 4, 256, 32), (96, 512, 4, 256, 2)]
for tile_kb, thr, npat, cdiv, wdiv in grid + extra:
    os.environ.update(PSS_TILE_BYTES=str(tile_kb * 1024), PSS_SIEVE_THREADS=str(thr), PSS_NPAT=str(npat),
                      PSS_CTA_WIDE_DIV=str(cdiv), PSS_WARP_WIDE_DIV=str(wdiv))
    try:
        h = _lib.Handle(-1)
        best = None
        for _ in range(3):
            h.stats_reset()
            c = h.count_primes(0, HI)
            s = h.stats()
            best = s["ms_sieve"] if best is None else min(best, s["ms_sieve"])
        rows.append(dict(tile_kb=tile_kb, threads=thr, npat=npat, cta_div=cdiv, warp_div=wdiv, ms_sieve=round(best, 3), count=c))
        print(rows[-1], flush=True)
        h.close()
    except Exception as e:  # noqa: BLE001
        print("FAIL", tile_kb, thr, npat, cdiv, wdiv, e, flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "tune_sieve.json"), "w"), indent=1)
