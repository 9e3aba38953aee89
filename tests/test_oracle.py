"""The oracle (oracle/oracle.py + oracle/oracle.c) against the reference's golden vectors
(tests/golden/reference_vectors.json, produced by oracle/gen_golden.py from the reference itself)
and the reference's own KATs."""
import hashlib

import numpy as np
import pytest


def sha_u64(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<u8").tobytes()).hexdigest()


def test_reference_kats(orc):
    # tests/test_sieve.py:29-44, tests/test_gpu.py:20-23, tests/test_sum_cache.py:235-248, test_number_theory.py:386-388
    for k, v in zip(range(1, 7), [4, 25, 168, 1229, 9592, 78498]):
        assert orc.prime_count(10 ** k) == v
    assert [orc.nth_prime(n) for n in (1, 7, 100, 1000)] == [2, 17, 541, 7919]
    first7 = orc.first_n_primes(7)
    assert [orc.power_sum(first7, k) for k in (1, 2, 3)] == [58, 666, 8944]
    assert orc.power_sum(orc.first_n_primes(15), 2) == 10466
    assert [orc.primesum(n, 2) for n in range(1, 8)] == [4, 13, 38, 87, 208, 377, 666]   # proofs/PrimesumMod6.lean:86-94
    assert (orc.trisum(3), orc.trisum(6), orc.trisum(10)) == (3, 35, 666)
    assert orc.primesum(0, 2) == 0 and orc.primesum(5, 0) == 5
    with pytest.raises(ValueError):
        orc.primesum(-1, 2)
    with pytest.raises(ValueError):
        orc.nth_prime(0)


def test_generate_primes_golden(orc, golden):
    for row in golden["generate_primes"]:
        p = orc.primes_below(row["limit"])
        assert len(p) == row["count"] and sha_u64(p) == row["sha256"]
        c = orc.c_primes_range(0, max(row["limit"], 0))
        assert len(c) == row["count"] and sha_u64(c) == row["sha256"]
        if row["primes"] is not None:
            assert [int(x) for x in p] == row["primes"]


def test_ranges_golden(orc, golden):
    for row in golden["generate_primes_range"]:
        c = orc.c_primes_range(row["start"], row["stop"])
        assert len(c) == row["count"] and sha_u64(c) == row["sha256"], row
        for seg in (1000, 4096, 99991):
            assert orc.c_count_range(row["start"], row["stop"], seg) == row["count"]


def test_first_n_and_counts_golden(orc, golden):
    for row in golden["generate_n_primes"]:
        c = orc.c_first_n_primes(row["n"])
        assert len(c) == row["count"] and sha_u64(c) == row["sha256"]
    for row in golden["nth_prime"]:
        assert orc.nth_prime(row["n"]) == row["p"]
    for row in golden["prime_count"]:
        assert orc.c_count_range(0, row["limit"] + 1) == row["count"]


def test_primesum_golden(orc, golden):
    cache = {}
    for row in golden["primesum"]:
        n, k = row["n"], row["k"]
        if n > 10 ** 6:
            continue   # the 5e6 / 1e7 rows are covered by the GPU parity tests and gen_golden itself
        if n not in cache:
            cache[n] = orc.c_first_n_primes(n)
        got = orc.c_power_sums(cache[n], [k])[0] if n else 0
        assert str(got) == row["value"], (n, k)


def test_prefix_sums_golden(orc, golden):
    primes = orc.c_first_n_primes(5000)
    for row in golden["prefix_sums"]:
        sums = orc.c_prefix_sums(primes, row["k"])
        h = hashlib.sha256()
        for s in sums:
            h.update(str(s).encode() + b"\n")
        assert h.hexdigest() == row["sha256_decimal_lines"]
        for i, v in row["samples"].items():
            assert str(sums[int(i) - 1]) == v


def test_find_matches_golden(orc, golden):
    for case in golden["find_matches"]:
        expr, bounds = case["expr"], case["bounds"]
        if "tri(m)" in expr:
            got = []
            for m in range(1, bounds["m"] + 1):   # m outer, n inner (alphabetical)
                t = orc.tri(m)
                for n, _ in orc.find_all(1, bounds["n"], [2], t, "=="):
                    got.append({"m": m, "n": n})
            assert got == case["matches"]
            continue
        quant, rest = expr.split(" ", 1)
        lhs, op, rhs = rest.rsplit(" ", 2)
        target = int(rhs)
        if "(n,p)" in lhs:
            powers = list(range(1, bounds["p"] + 1))
            res = [{"n": n, "p": p} for n, p in orc.find_all(1, bounds["n"], powers, target, op)]
        else:
            k = int(lhs.split(",")[1].rstrip(")"))
            res = [{"n": n} for n, _ in orc.find_all(1, bounds["n"], [k], target, op)]
        if quant == "does_exist":
            res = res[:1]
        assert res == case["matches"], expr


def test_scan_compare_matches_python(orc):
    primes = orc.c_first_n_primes(3000)
    for k in (1, 2, 5, 8):
        sums = orc.c_prefix_sums(primes, k)
        for hit in (1, 7, 1500, 3000):
            r = orc.c_scan_compare(primes, 1, k, 0, sums[hit - 1], "==", 1, 3000)
            assert (r["match_count"], r["first_match_n"], r["cross_n"]) == (1, hit, hit)
            assert r["sum_at_cross"] == sums[hit - 1] and r["final_sum"] == sums[-1]
        r = orc.c_scan_compare(primes, 1, k, 0, sums[99] + 1, "<", 1, 3000)
        assert (r["match_count"], r["first_match_n"], r["last_match_n"], r["cross_n"]) == (100, 1, 100, 101)
    # carry-in / n_first
    half = orc.c_power_sums(primes[:1000], [3])[0]
    r = orc.c_scan_compare(primes[1000:], 1001, 3, half, orc.c_prefix_sums(primes, 3)[2221], "==", 1, 3000)
    assert r["first_match_n"] == 2222


def test_block_sieve_sum_equals_two_step(orc):
    for lo, hi in [(0, 100000), (10 ** 6, 10 ** 6 + 50000), (2 ** 32 - 1000, 2 ** 32 + 1000)]:
        n, sums, last = orc.c_block_sieve_sum(lo, hi, [1, 2, 5])
        p = orc.c_primes_range(lo, hi)
        assert n == len(p) and last == int(p[-1]) and sums == orc.c_power_sums(p, [1, 2, 5])


def test_primesieve_standin_equals_the_port(orc, golden):
    """bench.py's second CPU number (bit-packed wheel-30 segmented sieve + 256-bit sums, a stand-in for the primesieve
    library the reference prefers) gives exactly the port's counts, sums and last primes: range edges inside a wheel byte,
    below 7, a segment boundary, around 2^32, near 2^40, and the reference's KATs (sum of p^k over the first 7 primes)."""
    import random
    rng = random.Random(5)
    cases = [(0, 2), (0, 3), (2, 3), (0, 8), (5, 8), (0, 18), (7, 8), (49, 50), (0, 30), (29, 32), (0, 1000), (0, 3_000_000),
             (30 * 262144 - 40, 30 * 262144 + 40), (10 ** 9, 10 ** 9 + 300_000), (2 ** 32 - 10 ** 5, 2 ** 32 + 10 ** 5),
             (2 ** 40 - 200_000, 2 ** 40)]
    for _ in range(25):
        lo = rng.randrange(0, 10 ** 7)
        cases.append((lo, lo + rng.randrange(1, 200_000)))
    for lo, hi in cases:
        for powers in ([2], [1, 2, 3, 4, 5]):
            assert orc.c_fast_block_sieve_sum(lo, hi, powers) == orc.c_block_sieve_sum(lo, hi, powers), (lo, hi, powers)
    assert orc.c_fast_block_sieve_sum(0, 18, [1, 2, 3]) == (7, [58, 666, 8944], 17)      # tests/test_gpu.py:20-23 of the reference


def test_at_scale_consistency(golden):
    # rows recorded from SURVEY.md Appendix B: internal consistency checks that need no sieve
    rows = golden["at_scale"]
    for r in rows:
        n = r["n"]
        assert int(r["sums"]["2"]) % 6 == (n + 5) % 6     # proofs/PrimesumMod6.lean:228-250
    for a, b in zip(rows, rows[1:]):
        for k in "12345":
            assert int(a["sums"][k]) < int(b["sums"][k])
