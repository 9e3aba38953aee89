#!/usr/bin/env python
"""Sieve-only driver for profiling: count primes below TUNE_HI a few times, print the sieve kernel time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prime-square-sum_b200"))
from pss_b200 import _lib

HI = int(os.environ.get("TUNE_HI", 22801763490))
h = _lib.Handle(-1)
for _ in range(int(os.environ.get("REPS", 3))):
    h.stats_reset()
    c = h.count_primes(0, HI)
    print(c, h.stats()["ms_sieve"], flush=True)
h.close()
