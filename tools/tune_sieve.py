#!/usr/bin/env python
"""Sweep the sieve tunables on the C3 range and print per-kernel device times (run on the GPU box)."""
import itertools, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prime-square-sum_b200"))
from pss_b200 import _lib

HI = int(os.environ.get("TUNE_HI", 22801763490))
rows = []
grid = list(itertools.product([int(x) for x in os.environ.get("TUNE_TILES", "216,224").split(",")], [int(x) for x in os.environ.get("TUNE_THREADS", "1024").split(",")], [int(x) for x in os.environ.get("TUNE_NPAT", "4").split(",")], [int(x) for x in os.environ.get("TUNE_DIVS", "104").split(",")], [int(x) for x in os.environ.get("TUNE_LEADS", "4294967295").split(",")]))
for tile_kb, thr, npat, wdiv, lead in grid:
    os.environ.update(PSS_TILE_BYTES=str(tile_kb * 1024), PSS_SIEVE_THREADS=str(thr), PSS_NPAT=str(npat), PSS_WARP_WIDE_DIV=str(wdiv), PSS_SPARSE_LEAD=str(lead))
    try:
        h = _lib.Handle(-1)
        best = None
        for _ in range(3):
            h.stats_reset()
            c = h.count_primes(0, HI)
            s = h.stats()
            best = s["ms_sieve"] if best is None else min(best, s["ms_sieve"])
        rows.append(dict(tile_kb=tile_kb, threads=thr, npat=npat, warp_div=wdiv, sparse_lead=lead, ms_sieve=round(best, 3), count=c))
        print(rows[-1], flush=True)
        h.close()
    except Exception as e:  # noqa: BLE001
        print("FAIL", tile_kb, thr, npat, wdiv, lead, e, flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "tune_sieve.json"), "w"), indent=1)
