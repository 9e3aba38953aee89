#!/usr/bin/env python
"""Randomised soak of the block-sum kernels: byte-table moments (default) against the per-prime multi-limb power chain
(PSS_REDUCE_MOMENTS=0) on random value windows up to 2^40, every power / all-powers mode; a sample of the windows is also
checked against the oracle.  Usage: tools/soak_moments.py [seconds] [seed]"""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prime-square-sum_b200"))
sys.path.insert(0, ROOT)
from pss_b200 import _lib

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
os.environ["PSS_REDUCE_MOMENTS"] = "0"
h_chain = _lib.Handle(-1)
os.environ["PSS_REDUCE_MOMENTS"] = "1"
h_walk = _lib.Handle(-1)
del os.environ["PSS_REDUCE_MOMENTS"]
h_tab = _lib.Handle(-1)
try:
    from oracle import oracle as orc
    orc.build()
except Exception:
    orc = None
modes = [(1, False), (2, False), (3, False), (4, False), (5, False), (2, True), (3, True), (4, True), (5, True)]
t0 = time.time()
n = n_orc = 0
while time.time() - t0 < budget:
    top = rng.choice([10 ** 4, 10 ** 6, 10 ** 8, 2 ** 32, 10 ** 11, 2 ** 40])
    width = rng.choice([1, 30, 2880, 10 ** 5, 3 * 10 ** 6, 4 * 10 ** 7])
    wide = rng.random() < 0.04
    if wide:                     # wide enough for the position-specific tables of k != 2 (two rounds of warp tiles per SM)
        width = rng.choice([7 * 10 ** 8, 15 * 10 ** 8]) + rng.randrange(0, 10 ** 6)
        top = max(top, 2 * width)
    lo = rng.randrange(0, max(1, top - width))
    hi = min(2 ** 40, lo + (width if wide else rng.randrange(1, width + 1)))
    power, multi = rng.choice(modes)
    a = h_tab.slice_sieve_sums(lo, hi, power, multi)
    b = h_chain.slice_sieve_sums(lo, hi, power, multi)
    c = h_walk.slice_sieve_sums(lo, hi, power, multi)
    assert a == b == c, (lo, hi, power, multi, a, b, c)
    if orc is not None and hi - lo <= 3 * 10 ** 6 and n % 5 == 0:
        p = orc.c_primes_range(lo, hi)
        powers = list(range(1, power + 1)) if multi else [power]
        assert a == (len(p), orc.c_power_sums(p, powers)), (lo, hi, power, multi)
        n_orc += 1
    n += 1
print(f"soak ok: {n} windows ({n_orc} also against the oracle) in {time.time() - t0:.1f} s")
