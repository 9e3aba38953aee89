#!/usr/bin/env python
"""Randomised soak of the query paths against each other on the device: the fused path with interval pruning (default), the
fused path with the exhaustive per-prime compare, and the materialised path (u64 primes, tile sums + pruned walk) must
return identical records — match count / first / last, crossing index and prime, the two bracket sums, S_k(n_max), p_(n_max) —
for random (n_max, power, all-powers, operator, window, target) where the target is a partial sum produced by an independent
pass (so hits land on arbitrary block / word / tile positions, including n_max itself and 2, 3, 5).
Usage: tools/soak_search.py [seconds] [seed]"""
import os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "prime-square-sum_b200"))
from pss_b200 import _lib, search

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 20261020)
h = _lib.Handle(-1)


def record(o):
    return (o.last_prime, [(p.power, p.match_count, p.first_match_n, p.last_match_n, p.cross_n, p.cross_prime, p.sum_at_cross,
                            p.sum_before_cross, p.final_sum) for p in o.powers])


def run(mode, *a, **kw):
    h.set_search_path(mode == "materialised")
    h.set_compare_mode(mode == "exhaustive")
    try:
        return search.search(*a, handle=h, **kw)
    finally:
        h.set_search_path(False)
        h.set_compare_mode(False)


t0, n_cases, bad = time.time(), 0, 0
while time.time() - t0 < budget:
    n_max = rng.choice([rng.randint(1, 40), rng.randint(1, 5000), rng.randint(1, 10 ** 6), rng.randint(1, 2 * 10 ** 7)] +
                       ([rng.randint(10 ** 8, 10 ** 9)] if rng.random() < 0.02 else []))
    power = rng.randint(1, 8)
    allp = power <= 6 and power > 1 and rng.random() < 0.4
    op = rng.choice(["==", "==", ">=", "<", "!=", "<=", ">"])
    n_min = 1 if rng.random() < 0.6 else rng.randint(1, n_max)
    hit = rng.choice([n_max, 1, rng.randint(1, n_max), rng.randint(max(1, n_max - 3000), n_max)])
    k_t = rng.randint(1, power) if allp else power
    tgt = run("fused", 0, hit, power=k_t, op="<").powers[0].final_sum + rng.choice([0, 0, 0, 1, -1])
    tgt = max(tgt, 0)
    ref = record(run("fused", tgt, n_max, power=power, n_min=n_min, op=op, all_powers=allp))
    for mode in ("exhaustive", "materialised"):
        got = record(run(mode, tgt, n_max, power=power, n_min=n_min, op=op, all_powers=allp))
        if got != ref:
            bad += 1
            print(f"MISMATCH mode={mode} n_max={n_max} power={power} all={allp} op={op} n_min={n_min} hit={hit} k_t={k_t}\n  fused {ref}\n  {mode} {got}", flush=True)
    n_cases += 1
print(f"soak_search: {n_cases} cases x 3 paths, mismatches={bad}, {time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
