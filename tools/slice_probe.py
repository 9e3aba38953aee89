import os, sys
sys.path.insert(0, "/root/repo/prime-square-sum_b200")
from pss_b200 import _lib
N = 22801763490
h = _lib.Handle(-1)
for lo, hi in [(N * 7 // 8, N), (0, N // 8), (N * 3 // 8, N * 4 // 8), (0, N)]:
    best = None
    for _ in range(4):
        h.stats_reset(); c = h.count_primes(lo, hi); s = h.stats()
        best = s["ms_sieve"] if best is None else min(best, s["ms_sieve"])
    print(f"[{lo:.3e}, {hi:.3e}) count={c} ms_sieve={best:.4f} -> x8 = {best*8:.3f}", flush=True)
h.close()
