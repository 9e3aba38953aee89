"""GPU parity tests: the CUDA path through the C ABI (ctypes) against the oracle and the reference's
golden vectors.  Bit-exact everywhere (integer work): prime arrays, counts, every partial sum, match index,
bracket.  Run on the GPU box with `pytest -m gpu`."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def sha_u64(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<u8").tobytes()).hexdigest()


# ------------------------------------------------------------------------------------- K1 sieve
def test_native_library_loaded(handle):
    from pss_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    info = handle.device_info()
    assert "sm_10" in info, info


def test_generate_primes_golden(handle, golden):
    from pss_b200 import sieve
    for row in golden["generate_primes"]:
        p = sieve.generate_primes(row["limit"])
        assert p.dtype == np.int64
        assert len(p) == row["count"], row
        assert sha_u64(p) == row["sha256"], row
        if row["primes"] is not None:
            assert [int(x) for x in p] == row["primes"]


def test_ranges_golden(handle, golden):
    from pss_b200 import sieve
    for row in golden["generate_primes_range"]:
        p = sieve.generate_primes_range(row["start"], row["stop"])
        assert len(p) == row["count"] and sha_u64(p) == row["sha256"], row
        assert handle.count_primes(row["start"], row["stop"]) == row["count"]


def test_first_n_nth_count_golden(handle, golden):
    from pss_b200 import sieve
    for row in golden["generate_n_primes"]:
        p = sieve.generate_n_primes(row["n"])
        assert len(p) == row["count"] and sha_u64(p) == row["sha256"], row
    for row in golden["nth_prime"]:
        assert sieve.nth_prime(row["n"]) == row["p"]
    for row in golden["prime_count"]:
        assert sieve.prime_count(row["limit"]) == row["count"]
    # reference KATs: tests/test_sieve.py:29-44
    assert [sieve.prime_count(10 ** k) for k in range(1, 7)] == [4, 25, 168, 1229, 9592, 78498]
    with pytest.raises(ValueError):
        sieve.nth_prime(0)


@pytest.mark.parametrize("tile,threads,npat", [(8192, 256, 4), (98304, 512, 4), (65536, 1024, 2), (16384, 512, 0),
                                               (216 * 1024, 1024, 4), (32768, 256, 1), (49152, 512, 3), (224 * 1024, 1024, 4),
                                               (112 * 1024, 768, 4), (224 * 1024, 1024, 5), (96 * 1024, 512, 6)])
def test_sieve_tunings_vs_oracle(orc, tile, threads, npat):
    """Every tile size / block size / pattern count gives the same primes (segment-size independence,
    the analogue of tests/test_sieve.py:80-120)."""
    from pss_b200 import _lib
    os.environ["PSS_NPAT"] = str(npat)
    try:
        h = _lib.Handle(-1)
    finally:
        del os.environ["PSS_NPAT"]
    h.set_tuning(tile, threads, 0)
    for lo, hi in [(0, 3_000_000), (0, 61), (59, 62), (3600, 3800), (10 ** 9, 10 ** 9 + 2_000_000),
                   (2 ** 32 - 10 ** 6, 2 ** 32 + 10 ** 6), (22_801_000_000, 22_801_763_490), (2 ** 39, 2 ** 39 + 500_000)]:
        got = h.generate_primes(lo, hi).astype(np.uint64)
        exp = orc.c_primes_range(lo, hi)
        assert np.array_equal(got, exp), (lo, hi, len(got), len(exp))
    h.close()


@pytest.mark.parametrize("lead", [0, 3, 40])
def test_sieve_queue_orders_vs_oracle(orc, lead):
    """phase B's work queue: the sparse batches before the dense primes (default), after them (0), or split (3, 40) -- every
    slot -> item mapping crosses off the same composites"""
    from pss_b200 import _lib
    os.environ["PSS_SPARSE_LEAD"] = str(lead)
    try:
        h = _lib.Handle(-1)
    finally:
        del os.environ["PSS_SPARSE_LEAD"]
    for lo, hi in [(0, 8_000_000), (10 ** 9, 10 ** 9 + 8_000_000), (22_790_000_000, 22_801_763_490)]:
        assert np.array_equal(h.generate_primes(lo, hi).astype(np.uint64), orc.c_primes_range(lo, hi)), (lead, lo, hi)
    h.close()


@pytest.mark.parametrize("env", [None, "PSS_COMPACT_WALK", "PSS_COMPACT_BLOCK"])
def test_compaction_kernels_vs_oracle(orc, env):
    """The three stream-compaction kernels (byte-table form = default, per-prime walk, block-synchronous) write the same
    sorted u64 buffer: ranges below 120 (bytes with more than four primes), ragged edges, first-n truncation, 2^32, 2^39."""
    from pss_b200 import _lib
    if env:
        os.environ[env] = "1"
    try:
        h = _lib.Handle(-1)
        for lo, hi in [(0, 10), (0, 121), (0, 30_721), (7, 3_000_000), (30_719, 30_722), (10 ** 9, 10 ** 9 + 3_000_000),
                       (2 ** 32 - 10 ** 6, 2 ** 32 + 10 ** 6), (2 ** 39, 2 ** 39 + 500_000)]:
            got = h.generate_primes(lo, hi).astype(np.uint64)
            assert np.array_equal(got, orc.c_primes_range(lo, hi)), (env, lo, hi)
        for n in (1, 3, 4, 10, 3316, 3317, 100_000, 1_000_003):
            got = h.generate_n_primes(n).astype(np.uint64)
            assert np.array_equal(got, orc.c_first_n_primes(n)), (env, n)
        h.close()
    finally:
        if env:
            del os.environ[env]


def test_individual_variant_matches_the_sieve(handle, orc):
    """--algorithm sieve:individual (reference utils/sieve.py:330-387): per-candidate trial division on the device gives
    the same primes, counts and sums as the segmented sieve and the oracle (the reference's own cross-check of its
    variants, tests/test_sieve.py:80-120)"""
    from pss_b200 import sieve
    try:
        for limit in (2, 3, 8, 100, 10007, 300_000):
            got = sieve.generate_primes(limit, algorithm="individual")
            assert np.array_equal(got.astype(np.uint64), orc.c_primes_range(0, limit)), limit
        got = sieve.generate_primes_range(10 ** 9, 10 ** 9 + 20_000, algorithm="individual")
        assert np.array_equal(got.astype(np.uint64), orc.c_primes_range(10 ** 9, 10 ** 9 + 20_000))
        assert [int(x) for x in sieve.generate_n_primes(25, algorithm="individual")] == [int(x) for x in orc.c_first_n_primes(25)]
        assert sieve.nth_prime(1000, algorithm="individual") == 7919 and sieve.prime_count(10 ** 5, algorithm="individual") == 9592
        sieve.configure(algorithm="individual")
        assert sieve.prime_count(10 ** 4) == 1229
        cnt, sums = handle.slice_sieve_sums(0, 200_000, 3, True)          # the fused passes run on its bitmap unchanged
        assert (cnt, sums) == (17984, orc.c_power_sums(orc.c_primes_range(0, 200_000), [1, 2, 3]))
    finally:
        sieve.configure(algorithm="auto")
        assert sieve.prime_count(10 ** 6) == 78498                         # back on the segmented sieve
    for name in ("segmented", "basic", "b200"):
        assert sieve.prime_count(10 ** 5, algorithm=name) == 9592


def test_sieve_medium_vs_oracle(handle, orc):
    exp = orc.c_primes_range(0, 2 * 10 ** 8)
    got = handle.generate_primes(0, 2 * 10 ** 8).astype(np.uint64)
    assert np.array_equal(got, exp)
    assert handle.count_primes(0, 10 ** 9) == 50847534          # pi(1e9)  


# ------------------------------------------------------------------------------- K2 power sums
def test_power_sum_kats(handle):
    from pss_b200 import gpu
    small = np.array([2, 3, 5, 7, 11, 13, 17], dtype=np.int64)
    assert [gpu.power_sum(small, k) for k in (1, 2, 3)] == [58, 666, 8944]      # tests/test_gpu.py:20-23
    assert gpu.power_sum(np.array([], dtype=np.int64), 2) == 0                   # tests/test_gpu.py:67-69
    for k in range(0, 9):
        assert gpu.power_sum(small, k) == sum(int(p) ** k for p in small)
    big = np.array([1_000_000_007, 2 ** 40 - 87, 22_801_763_489], dtype=np.int64)
    for k in range(1, 9):
        assert gpu.power_sum(big, k) == sum(int(p) ** k for p in big)
    with pytest.raises(ValueError):
        gpu.power_sum(np.array([2 ** 41], dtype=np.int64), 2)


def test_primesum_golden(handle, golden):
    from pss_b200 import sequences
    for row in golden["primesum"]:
        got = sequences.primesum_direct(row["n"], row["k"])
        assert str(got) == row["value"], (row["n"], row["k"])


def test_primesum_operator_semantics(handle):
    from pss_b200 import primesum
    assert [primesum(n, 2) for n in range(1, 8)] == [4, 13, 38, 87, 208, 377, 666]
    assert primesum(3, 1) == 10 and primesum(4, 1) == 17 and primesum(1, 3) == 8 and primesum(2, 3) == 35
    assert primesum(0, 2) == 0 and primesum(9, 0) == 9 and primesum(7.0, 2) == 666
    with pytest.raises(ValueError, match="non-negative"):
        primesum(-1, 2)
    with pytest.raises(ValueError, match="integral"):
        primesum(7.5, 2)
    with pytest.raises(TypeError):
        primesum("7", 2)


def test_every_partial_sum(handle, orc, golden):
    """every S_k(n), n <= 5000 and across several scan tiles, bit-exact"""
    n = 70000
    handle.slice_sieve(0, handle.nth_prime_upper(n) + 1)
    primes = orc.c_first_n_primes(n)
    for row in golden["prefix_sums"]:
        k = row["k"]
        sums = handle.slice_prefix_sums(0, 5000, k)
        h = hashlib.sha256()
        for s in sums:
            h.update(str(s).encode() + b"\n")
        assert h.hexdigest() == row["sha256_decimal_lines"], k
    for k in (1, 2, 3, 5, 8):
        got = handle.slice_prefix_sums(0, n, k)
        assert got == orc.c_prefix_sums(primes, k), k
        # window with a carry, not tile aligned
        carry = got[12344]
        assert handle.slice_prefix_sums(12345, 20000, k, carry) == got[12345:32345]


# ------------------------------------------------------------------------------- K3 scan + compare
@pytest.fixture(params=["fused", "fused-exhaustive", "materialised"])
def search_path(request, handle):
    """every query path: fused (K2/K3 read the sieve bitmap) with interval pruning (default) and with the exhaustive
    per-prime compare, and materialised (K1c u64 primes + look-back scan)"""
    handle.set_search_path(request.param == "materialised")
    handle.set_compare_mode(request.param == "fused-exhaustive")
    yield request.param
    handle.set_search_path(False)
    handle.set_compare_mode(False)


def test_find_matches_golden(handle, golden, search_path):
    """the reference evaluator's own answers (utils/grammar.py find_matches) for constant targets"""
    from pss_b200 import search
    for case in golden["find_matches"]:
        expr, bounds = case["expr"], case["bounds"]
        if "tri(m)" in expr:
            continue
        quant, rest = expr.split(" ", 1)
        lhs, op, rhs = rest.rsplit(" ", 2)
        target = int(rhs)
        if "(n,p)" in lhs:
            it = search.for_any(target, bounds["n"], power=bounds["p"], op=op, all_powers=True)
        else:
            k = int(lhs.split(",")[1].rstrip(")"))
            it = search.for_any(target, bounds["n"], power=k, op=op)
        res = list(it)
        if quant == "does_exist":
            res = res[:1]
        assert res == case["matches"], expr


def test_fastpath_reproduces_reference_cli_output(handle, golden):
    """exact stdout lines / rc of the reference CLI for the hot-path query shapes
    (tests/test_cli_integration.py:156-182,257-267; golden find_matches cases)"""
    from pss_b200.fastpath import try_fast_query
    assert try_fast_query("does_exist primesum(n,2) == 666") == (["Found: n=7"], 0)
    assert try_fast_query("primesum(n,2) == 666") == (["Found: n=7"], 0)                  # default quantifier
    assert try_fast_query("does_exist primesum(n,2) == 667", {"n": 1000}) == (["No match found within bounds."], 1)
    assert try_fast_query("for_any primesum(n,2) == tri(m)", {"n": 100, "m": 3200}) == (["Found: m=36, n=7", "Found: m=3169, n=86"], 0)
    assert try_fast_query("does_exist primesum(n,2) == 666", fmt="json") == (['{"found": true, "variables": {"n": 7}}'], 0)
    assert try_fast_query("does_exist primesum(n,2) == 667", {"n": 50}, fmt="json") == (['{"found": false, "variables": null}'], 1)
    assert try_fast_query("does_exist primesum(n,2) == 667", {"n": 50}, fmt="csv") == (["not_found"], 1)
    assert try_fast_query("does_exist primesum(n,2) == trisum(10)") == (["Found: n=7"], 0)
    assert try_fast_query("does_exist primesum(n,2) == trisum(666)", {"n": 200000}) == (["No match found within bounds."], 1)
    assert try_fast_query("does_exist primesum(n,2) == 666", {"n": 100}, {"n": 8}) == (["No match found within bounds."], 1)
    assert try_fast_query("primesum(n,2) + 1 == 667") is None and try_fast_query("fibonacci(n) == 8") is None
    # shapes a regex would miss: parentheses, qualified name, reversed comparison
    assert try_fast_query("does_exist (primesum(n, 2)) == (666)") == (["Found: n=7"], 0)
    assert try_fast_query("does_exist pss.primesum(n,2) == 666") == (["Found: n=7"], 0)
    assert try_fast_query("does_exist trisum(10) == primesum(n,2)") == (["Found: n=7"], 0)
    assert try_fast_query("for_any 100 < primesum(n,2)", {"n": 8}) == (["Found: n=5", "Found: n=6", "Found: n=7", "Found: n=8"], 0)
    assert try_fast_query("for_any primesum(n,p) == 8944", {"n": 12, "p": 3}) == (["Found: n=7, p=3"], 0)
    assert try_fast_query("for_any primesum(n,p) == 8944", {"n": 12, "p": 7}) is None       # p beyond the all-powers pass
    for case in golden["find_matches"]:
        res = try_fast_query(case["expr"], case["bounds"])
        exp = ["Found: " + ", ".join(f"{k}={v}" for k, v in sorted(mm.items())) for mm in case["matches"]]
        assert res == ((exp, 0) if exp else (["No match found within bounds."], 1)), case["expr"]


def test_tri_predicate(handle, orc, golden):
    """config C2's RHS: for_any primesum(n,2) == tri(m) (reference answer from find_matches: m=36,n=7; m=3169,n=86)"""
    from pss_b200 import search
    case = next(c for c in golden["find_matches"] if "tri(m)" in c["expr"])
    assert search.for_any_tri(case["bounds"]["n"], case["bounds"]["m"]) == case["matches"]
    assert search.for_any_tri(100, 3168) == [{"m": 36, "n": 7}]            # m bound is inclusive
    assert search.for_any_tri(100, 3200, n_min=8) == [{"m": 3169, "n": 86}]
    assert search.for_any_tri(85, 10 ** 7) == [{"m": 36, "n": 7}]
    # every power against the oracle's qtri on every partial sum
    primes = orc.c_first_n_primes(200000)
    for k in (1, 2, 3):
        sums = orc.c_prefix_sums(primes, k)
        exp = [{"m": orc.qtri(s), "n": i + 1} for i, s in enumerate(sums) if orc.qtri(s) is not None and 1 <= orc.qtri(s) <= 2 ** 28]
        assert search.for_any_tri(200000, 2 ** 28, power=k) == sorted(exp, key=lambda d: (d["m"], d["n"])), k
    # windows in n and m (k = 2: word-parallel head pass + scan kernel; the window edges fall inside the head, at its end
    # and beyond it)
    sums2 = orc.c_prefix_sums(primes, 2)
    allhits = [(i + 1, orc.qtri(s)) for i, s in enumerate(sums2) if orc.qtri(s) is not None]
    assert len(allhits) >= 2
    for n_lo, n_hi, m_lo, m_hi in [(1, 6, 1, 10 ** 7), (7, 7, 36, 36), (8, 85, 1, 10 ** 7), (86, 86, 3169, 3169), (1, 200000, 37, 2 ** 29 - 1),
                                   (5000, 150000, 1, 2 ** 28), (1, 11263, 1, 10 ** 7), (11264, 200000, 1, 10 ** 7), (87, 200000, 1, 2 ** 28)]:
        exp = sorted(({"m": m, "n": n} for n, m in allhits if n_lo <= n <= n_hi and m_lo <= m <= m_hi), key=lambda d: (d["m"], d["n"]))
        assert search.for_any_tri(n_hi, m_hi, n_min=n_lo, m_min=m_lo) == exp, (n_lo, n_hi, m_lo, m_hi)
    # within tri(1e7) the only hits of primesum(n,2) are (7,36) and (86,3169)
    assert search.for_any_tri(10 ** 6, 10 ** 7) == [{"m": 36, "n": 7}, {"m": 3169, "n": 86}]


@pytest.mark.parametrize("k,multi", [(1, False), (2, False), (3, False), (4, False), (5, False), (6, False), (7, False),
                                     (8, False), (2, True), (3, True), (5, True), (6, True)])
def test_synthetic_hits_and_brackets(handle, orc, k, multi, search_path):
    """No trisum(666) config can match: validate match index / bracket with synthetic targets
    at thread, warp, tile and range boundaries."""
    from pss_b200 import search
    n_max = 40000
    primes = orc.c_first_n_primes(n_max)
    powers = list(range(1, k + 1)) if multi else [k]
    sums = {p: orc.c_prefix_sums(primes, p) for p in powers}
    items = 16 if (k <= 3 and not multi) else 8
    tile = 256 * items
    # thread / warp / tile boundaries of the materialised scan; the fused scan's tile boundaries fall at data
    # dependent indices (one tile = 737 280 integers), 3, 4 and the range ends exercise its 2-3-5 head
    hits = sorted({1, 2, 3, 4, 7, items, items + 1, 32 * items, 32 * items + 1, tile - 1, tile, tile + 1, 3 * tile,
                   3 * tile + 1, n_max - 1, n_max})
    for hit in hits:
        p = powers[hit % len(powers)]
        target = sums[p][hit - 1]
        out = search.search(target, n_max, power=k, all_powers=multi)
        assert out.first() == (hit, p) or (multi and out.first()[0] <= hit), (hit, p, out.first())
        po = next(x for x in out.powers if x.power == p)
        assert (po.match_count, po.first_match_n, po.cross_n) == (1, hit, hit)
        assert po.sum_at_cross == target and po.cross_prime == int(primes[hit - 1])
        assert po.sum_before_cross == (sums[p][hit - 2] if hit > 1 else 0)
        assert po.final_sum == sums[p][-1]
        # near miss: bracket only
        out = search.search(target + 1, n_max, power=k, all_powers=multi)
        po = next(x for x in out.powers if x.power == p)
        if hit < n_max:
            assert po.match_count == 0 and po.cross_n == hit + 1 and po.sum_before_cross == target
            assert out.bracket(p) == (hit, hit + 1)
        else:
            assert po.match_count == 0 and po.cross_n == 0 and out.bracket(p) == (n_max, n_max + 1)
    # window limits: a hit below n_min is not a match
    target = sums[powers[0]][99]
    out = search.search(target, n_max, power=k, n_min=101, all_powers=multi)
    assert out.powers[0].match_count == 0 and out.powers[0].cross_n == 100
    # other operators against the oracle
    for op in ("!=", "<", "<=", ">", ">="):
        o = orc.c_scan_compare(primes, 1, powers[0], 0, target, op, 50, 20000)
        out = search.search(target, 20000, power=k, n_min=50, op=op, all_powers=multi)
        po = out.powers[0]
        assert (po.match_count, po.first_match_n, po.last_match_n) == (o["match_count"], o["first_match_n"], o["last_match_n"]), op
        got = [n for n, pp in out.matches() if pp == powers[0]]
        exp = [n for n in range(50, 20001) if {"!=": sums[powers[0]][n - 1] != target, "<": sums[powers[0]][n - 1] < target,
                                                "<=": sums[powers[0]][n - 1] <= target, ">": sums[powers[0]][n - 1] > target,
                                                ">=": sums[powers[0]][n - 1] >= target}[op]]
        assert got == exp, op


def test_window_edges_at_every_level(handle, orc, search_path):
    """[n_min, n_max] windows whose edges fall inside scan tiles, warp tiles and thread blocks (several tiles deep): with
    interval pruning a window edge does not force a walk, so the match counts come from the emission arithmetic of all
    three levels; checked in closed form ('<' of a huge target / '>' of 0 match the whole window) and against the oracle's
    prefix sums for a crossing inside the window"""
    from pss_b200 import search
    n_top = 420_000                      # ~6.1e6 integers: 8 scan tiles of the fused path, 100+ tiles of the materialised one
    primes = orc.c_first_n_primes(n_top)
    for k, multi in ((2, False), (3, True)):
        sums = orc.c_prefix_sums(primes, k)
        for n_min, n_max in [(1, n_top), (2, 3), (4, 4), (777, 100_003), (61_234, 61_235), (59_999, 300_001), (200_000, n_top - 1),
                             (n_top, n_top), (123_456, 123_456 + 4096), (4097, 8 * 4096)]:
            w = n_max - n_min + 1
            for op, target in (("<", 10 ** 80), (">", 0), ("!=", 1)):
                o = search.search(target, n_max, power=k, n_min=n_min, op=op, all_powers=multi)
                for po in o.powers:
                    assert (po.match_count, po.first_match_n, po.last_match_n) == (w, n_min, n_max), (k, multi, n_min, n_max, op, po.power)
                assert o.last_prime == int(primes[n_max - 1]) and o.powers[-1].final_sum == sums[n_max - 1]
            hit = (n_min + n_max) // 2
            o = search.search(sums[hit - 1], n_max, power=k, n_min=n_min, op=">=", all_powers=multi)
            po = o.powers[-1]
            assert (po.match_count, po.first_match_n, po.last_match_n, po.cross_n) == (n_max - hit + 1, hit, n_max, hit)
            o = search.search(sums[hit - 1], n_max, power=k, n_min=n_min, op="<", all_powers=multi)
            po = o.powers[-1]
            assert po.match_count == hit - n_min and (po.match_count == 0 or (po.first_match_n, po.last_match_n) == (n_min, hit - 1))


def test_query_graph_survives_interleaving(orc):
    """the single-GPU query chain becomes a CUDA graph from the second query with the same geometry on:
    replays with other targets / windows / operators, other geometries and other API calls in between (which overwrite the
    staged tables and buffers the graph relies on), a tuning change and the exhaustive mode must all give the kernel-by-kernel
    answers (PSS_GRAPH=0 handle)"""
    from pss_b200 import _lib, search
    os.environ["PSS_GRAPH"] = "0"
    try:
        h_ref = _lib.Handle(-1)
    finally:
        del os.environ["PSS_GRAPH"]
    h = _lib.Handle(-1)
    primes = orc.c_first_n_primes(300_000)
    sums = {k: orc.c_prefix_sums(primes, k) for k in (1, 2, 3)}

    def same(n_max, **kw):
        a = search.search(handle=h, n_max=n_max, **kw)
        b = search.search(handle=h_ref, n_max=n_max, **kw)
        rec = lambda o: (o.first(), o.last_prime, [(p.power, p.match_count, p.first_match_n, p.last_match_n, p.cross_n, p.cross_prime,
                                                     p.sum_at_cross, p.sum_before_cross, p.final_sum) for p in o.powers])
        assert rec(a) == rec(b), (n_max, kw)
        return a

    G, G2 = 250_000, 77_777
    for hit in (5, 123_456, G, 9, 200_000):                       # capture on the 2nd, replays with other targets after
        assert same(G, target=sums[2][hit - 1], power=2).first() == (hit, 2)
    assert h.count_primes(10 ** 9, 10 ** 9 + 50_000_000) == h_ref.count_primes(10 ** 9, 10 ** 9 + 50_000_000)   # other tables
    assert same(G, target=sums[2][41], power=2, n_min=40, op=">=").powers[0].first_match_n == 42
    assert same(G2, target=sums[2][776], power=2).first() == (777, 2)                                          # other geometry
    assert same(G, target=sums[2][99_998], power=2).first() == (99_999, 2)                                     # back: no stale replay
    assert same(G, target=sums[2][99_998] + 1, power=2).first() is None
    assert int(h.generate_n_primes(1000)[-1]) == 7919                                                        # compaction in between
    for _ in range(3):
        assert same(G, target=sums[3][4_999], power=3, all_powers=True).first() == (5_000, 3)                 # other kernel instantiation
    h.set_compare_mode(True); h_ref.set_compare_mode(True)
    for hit in (17, 150_000):
        assert same(G, target=sums[2][hit - 1], power=2).first() == (hit, 2)                                   # exhaustive through the same graph
    h.set_compare_mode(False); h_ref.set_compare_mode(False)
    h.set_tuning(96 * 1024, 512, 0); h_ref.set_tuning(96 * 1024, 512, 0)
    for hit in (3, 222_222, 222_223):
        assert same(G, target=sums[1][hit - 1], power=1).first() == (hit, 1)                                   # new tuning: new geometry key
    h.set_kernel_events(2)
    for hit in (1, 2):
        assert same(G, target=sums[2][hit - 1], power=2).first() == (hit, 2)
    h.close(); h_ref.close()


def test_trisum666_configs_small(handle, orc, search_path):
    """config C1 and the C3/C4 shape at a size the oracle finishes in seconds: no match, bracket = n_max"""
    from pss_b200 import search
    T = orc.trisum(666)
    assert search.does_exist(666, 1_000_000) == {"n": 7}                      # C1
    n_max = 10 ** 6
    primes = orc.c_first_n_primes(n_max)
    for k in (2, 3, 4, 5):
        out = search.search(T, n_max, power=k)
        assert out.first() is None and out.bracket() == (n_max, n_max + 1)
        assert out.powers[0].final_sum == orc.c_power_sums(primes, [k])[0]
        assert out.last_prime == 15485863
    out = search.search(T, n_max, power=5, n_min=5 * 10 ** 5, all_powers=True)   # C5 shape
    assert out.first() is None
    assert [p.final_sum for p in out.powers] == orc.c_power_sums(primes, [1, 2, 3, 4, 5])


def test_fused_tile_boundaries(handle, orc):
    """fused path: hits on the first / last prime of several 737 280-integer scan tiles and sieve tiles"""
    from pss_b200 import search
    n_max = 400000
    primes = orc.c_first_n_primes(n_max)
    sums = orc.c_prefix_sums(primes, 2)
    hits = set()
    for boundary in (737280, 2 * 737280, 98304 * 30, 5 * 737280):
        idx = int(np.searchsorted(primes, boundary))   # primes[idx] is the first prime >= boundary
        hits.update({idx, idx + 1, idx + 2})
    for hit in sorted(hits):
        out = search.search(sums[hit - 1], n_max, power=2)
        assert out.first() == (hit, 2), hit
        po = out.powers[0]
        assert po.cross_prime == int(primes[hit - 1]) and po.sum_before_cross == sums[hit - 2]
        assert po.final_sum == sums[-1] and out.last_prime == int(primes[-1])


def test_moment_reduce_matches_walk_and_oracle(orc):
    """block sums from moments (default: per-byte tables for every k <= 5; PSS_REDUCE_MOMENTS=1: per-prime small moments +
    one binomial expansion per thread) against the per-prime multi-limb power chain (PSS_REDUCE_MOMENTS=0) and the oracle:
    slice counts and slice sums on ragged ranges (range edges inside a byte, below 7, around 2^32, near 2^40)"""
    from pss_b200 import _lib

    def handle_with(mode):
        os.environ["PSS_REDUCE_MOMENTS"] = mode
        try:
            return _lib.Handle(-1)
        finally:
            del os.environ["PSS_REDUCE_MOMENTS"]

    h_chain, h_walk = handle_with("0"), handle_with("1")
    h_tab = _lib.Handle(-1)
    ranges = [(0, 2), (0, 8), (0, 100), (3, 2881), (0, 3_000_000), (2879, 2 * 737280 + 17), (10 ** 9 + 7, 10 ** 9 + 4_000_001),
              (2 ** 32 - 10 ** 6, 2 ** 32 + 10 ** 6), (22_800_000_000, 22_801_763_490), (2 ** 39, 2 ** 39 + 3_000_000),
              (2 ** 40 - 2_000_000, 2 ** 40)]
    for power, multi in [(2, False), (1, False), (3, False), (4, False), (5, False), (2, True), (3, True), (4, True), (5, True),
                         (6, False), (6, True)]:
        powers = list(range(1, power + 1)) if multi else [power]
        for lo, hi in ranges:
            exp_p = orc.c_primes_range(lo, hi)
            exp = orc.c_power_sums(exp_p, powers)
            a = h_tab.slice_sieve_sums(lo, hi, power, multi)
            b = h_walk.slice_sieve_sums(lo, hi, power, multi)
            c = h_chain.slice_sieve_sums(lo, hi, power, multi)
            assert a == b == c == (len(exp_p), exp), (power, multi, lo, hi)
    for h in (h_chain, h_walk, h_tab):
        h.close()


def test_wide_table_block_sums_match_chain_at_scale():
    """the position-specific byte tables (k_bitmap_moments2w for k = 2, k_bitmap_moments_tabw for every other k <= 5 once every
    SM has two rounds of warp tiles: ranges of 6.6e8 integers and more) against the per-prime power chain (PSS_REDUCE_MOMENTS=0) on the
    same ranges, below and above 2^32, and through a search whose hit sits inside such a range (thread / warp records)"""
    from pss_b200 import _lib, search
    os.environ["PSS_REDUCE_MOMENTS"] = "0"
    try:
        h_chain = _lib.Handle(-1)
    finally:
        del os.environ["PSS_REDUCE_MOMENTS"]
    h_tab = _lib.Handle(-1)
    for lo, hi in [(0, 700_000_000), (2 ** 36 + 12345, 2 ** 36 + 700_000_000), (3, 1_400_000_011)]:
        for power, multi in [(2, False), (4, False), (5, False), (4, True), (5, True), (3, False), (1, False), (2, True), (3, True)]:
            assert h_tab.slice_sieve_sums(lo, hi, power, multi) == h_chain.slice_sieve_sums(lo, hi, power, multi), (lo, hi, power, multi)
    n_hit, n_max = 20_000_123, 36_000_000
    for power, multi in [(4, False), (5, False), (5, True), (2, False), (3, False), (3, True), (1, False)]:
        ref = search.search(1, n_hit, power=power, all_powers=multi, handle=h_chain)
        target = ref.powers[-1].final_sum
        a = search.search(target, n_max, power=power, all_powers=multi, handle=h_tab)
        b = search.search(target, n_max, power=power, all_powers=multi, handle=h_chain)
        assert a.first() == (n_hit, power) == b.first()
        assert a.last_prime == b.last_prime
        for pa, pb in zip(a.powers, b.powers):
            assert (pa.final_sum, pa.cross_n, pa.cross_prime, pa.sum_at_cross, pa.sum_before_cross) == \
                   (pb.final_sum, pb.cross_n, pb.cross_prime, pb.sum_at_cross, pb.sum_before_cross)
    h_chain.close()
    h_tab.close()


def test_fused_slices_compose(handle, orc):
    """fused multi-GPU decomposition on one device: slice_sieve_sums / slice_scan_bitmap with carries"""
    from pss_b200 import search
    n_max = 300000
    hi = handle.nth_prime_upper(n_max) + 1
    primes = orc.c_first_n_primes(n_max)
    for power, multi in ((3, False), (4, True)):
        powers = list(range(1, power + 1)) if multi else [power]
        sums = orc.c_prefix_sums(primes[:200001], powers[-1])
        target = sums[200000]
        slices = search.partition_range(hi, 4)
        info = [handle.slice_sieve_sums(lo, hi_g, power, multi) for lo, hi_g in slices]
        counts = [c for c, _ in info]
        ssums = [s for _, s in info]
        assert sum(counts) >= n_max
        # slice totals against the oracle
        off = 0
        for (c, sm) in info[:3]:
            assert sm == orc.c_power_sums(primes[off:off + c], powers)
            off += c
        parts = []
        for g, (lo, hi_g) in enumerate(slices):
            n_before, carry = search.combine_slices(counts, ssums, g)
            take = min(counts[g], max(0, n_max - n_before))
            if take == 0:
                parts.append(None)
                continue
            handle.slice_sieve_sums(lo, hi_g, power, multi)
            res = handle.slice_scan_bitmap(n_before + 1, carry, 1, n_max, power, target, "==", multi)
            parts.append(search._outcome_from_result(res, 1, n_max, "==", target))
        merged = search.merge_outcomes(parts, 1, n_max, "==", target)
        assert merged.first() == (200001, powers[-1])
        assert [p.final_sum for p in merged.powers] == orc.c_power_sums(primes, powers)
        assert merged.powers[-1].sum_before_cross == sums[199999]
        assert merged.last_prime == int(primes[-1])


def test_slices_compose(handle, orc):
    """the multi-GPU decomposition on one device: slices scanned with carries == one scan"""
    from pss_b200 import search
    n_max = 300000
    hi = handle.nth_prime_upper(n_max) + 1
    primes = orc.c_first_n_primes(n_max)
    sums = orc.c_prefix_sums(primes[:200001], 3)
    target = sums[200000]   # S_3(200001)
    whole = search.search(target, n_max, power=3)
    parts, counts, ssums = [], [], []
    slices = search.partition_range(hi, 4)
    for lo, hi_g in slices:
        c, _ = handle.slice_sieve(lo, hi_g)
        counts.append(c)
        ssums.append(handle.slice_reduce(0, c, 3) if c else [0])
    assert sum(counts) >= n_max
    for g, (lo, hi_g) in enumerate(slices):
        n_before, carry = search.combine_slices(counts, ssums, g)
        take = min(counts[g], max(0, n_max - n_before))
        if take == 0:
            parts.append(None)
            continue
        handle.slice_sieve(lo, hi_g)
        res = handle.slice_scan(0, take, n_before + 1, carry, 1, n_max, 3, target, "==")
        parts.append(search._outcome_from_result(res, 1, n_max, "==", target))
    merged = search.merge_outcomes(parts, 1, n_max, "==", target)
    assert merged.first() == whole.first() == (200001, 3)
    assert merged.powers[0].final_sum == whole.powers[0].final_sum == orc.c_power_sums(primes, [3])[0]
    assert merged.powers[0].sum_before_cross == sums[199999]
    assert merged.last_prime == whole.last_prime == int(primes[-1])


def test_sum_cache_device_advance(handle, orc, tmp_path):
    from pss_b200.sum_cache import IncrementalSumCache
    c = IncrementalSumCache(powers=[2, 3, 4])
    for p in orc.first_n_primes(100):   # tests/test_cache_integration.py:30-50 — 666 at exactly n = 7
        c.add_prime(int(p))
        if c.get_prime_count() == 7:
            assert c.get_sum(2) == 666
    c.advance_to(50000)
    primes = orc.c_first_n_primes(50000)
    assert c.get_prime_count() == 50000 and c.get_last_prime() == int(primes[-1])
    assert [c.get_sum(k) for k in (2, 3, 4)] == orc.c_power_sums(primes, [2, 3, 4])
    c.save_checkpoint(tmp_path / "s.npz")
    d = IncrementalSumCache.load_checkpoint(tmp_path / "s.npz")
    d.advance_to(120000)
    primes = orc.c_first_n_primes(120000)
    assert [d.get_sum(k) for k in (2, 3, 4)] == orc.c_power_sums(primes, [2, 3, 4])


def test_primesum_window_walk_and_jumps(handle, orc):
    """the plugin's primesum: sequential n (the evaluator's loop) crosses several device windows, each continued from the
    end of the previous one; jumps to arbitrary n take their carry from one fused pass"""
    from pss_b200 import sequences
    sequences.clear_cache()
    old = sequences.WINDOW
    sequences.WINDOW = 1000
    try:
        primes = orc.c_first_n_primes(60000)
        for k in (1, 2, 5):
            sums = orc.c_prefix_sums(primes, k)
            sequences.clear_cache()
            for n in list(range(1, 7001)) + [7, 6999, 20000, 20001, 59999, 60000, 12345, 1]:
                assert sequences.primesum(n, k) == sums[n - 1], (n, k)
    finally:
        sequences.WINDOW = old
        sequences.clear_cache()
    assert sequences.primesum(10 ** 7, 2) == 103158861357874372432083
    assert sequences.primesum(10 ** 7 + 1, 2) == 103158861357874372432083 + 179424691 ** 2


def test_search_resumes_from_checkpoint(handle, orc, tmp_path):
    """a search continued from an IncrementalSumCache checkpoint (reference ROADMAP #26 --start-n) gives the answer of
    the search from n = 1, for hits after the checkpoint, and never looks below it"""
    from pss_b200 import search
    from pss_b200.sum_cache import IncrementalSumCache
    n_max = 250000
    primes = orc.c_first_n_primes(n_max)
    c = IncrementalSumCache(powers=[1, 2, 3])
    c.advance_to(100000)
    c.save_checkpoint(tmp_path / "ck.npz")
    st = IncrementalSumCache.load_checkpoint(tmp_path / "ck.npz")
    for k, allp in ((2, False), (3, True)):
        sums = orc.c_prefix_sums(primes, k)
        for hit in (100001, 100002, 180000, n_max):
            full = search.search(sums[hit - 1], n_max, power=k, all_powers=allp)
            res = search.search(sums[hit - 1], n_max, power=k, all_powers=allp, resume=st)
            assert res.first() == full.first() == (hit, k)
            assert [(p.cross_n, p.cross_prime, p.sum_at_cross, p.sum_before_cross, p.final_sum) for p in res.powers] == \
                   [(p.cross_n, p.cross_prime, p.sum_at_cross, p.sum_before_cross, p.final_sum) for p in full.powers]
            assert res.last_prime == full.last_prime == int(primes[-1]) and res.stats["resumed_from_n"] == 100000
        assert search.search(sums[49999], n_max, power=k, all_powers=allp, resume=st).first() is None    # hit below the checkpoint
        o = search.search(10 ** 80, n_max, power=k, all_powers=allp, resume=st)
        assert o.first() is None and o.bracket() == (n_max, n_max + 1) and o.powers[-1].final_sum == sums[-1]
    assert st.resume_search(orc.c_prefix_sums(primes, 2)[199999], n_max).first() == (200000, 2)
    with pytest.raises(ValueError, match="does not track"):
        search.search(1, n_max, power=5, resume=st)


# ------------------------------------------------------------------------------- at scale
def test_at_scale_1e8(handle, golden):
    """n = 1e8 (config C2's stream; reference _basic_sieve + cpu_power_sum row)"""
    from pss_b200 import search
    row = next(r for r in golden["at_scale"] if r["n"] == 10 ** 8)
    T = int(golden["trisum"][-1]["value"])
    out = search.search(T, 10 ** 8, power=5, all_powers=True)
    assert out.last_prime == row["p_n"]
    assert [str(p.final_sum) for p in out.powers] == [row["sums"][str(k)] for k in range(1, 6)]
    assert out.first() is None
    got = handle.generate_n_primes(10 ** 8)
    ptr, cnt = handle.slice_device_buffer()
    assert cnt == 10 ** 8
    assert sha_u64(got) == row["sha256_primes_le_u64"]


def test_at_scale_c2(handle, golden):
    """config C2 at full size: for_any primesum(n,2) == tri(m), n <= 1e8, m <= 1e7"""
    from pss_b200 import search
    assert search.for_any_tri(10 ** 8, 10 ** 7) == [{"m": 36, "n": 7}, {"m": 3169, "n": 86}]


def test_at_scale_1e9_c3(handle, golden, search_path):
    """config C3: does_exist primesum(n,2) == trisum(666), n <= 1e9 — no match, bracket n = 1e9, exact S_2"""
    from pss_b200 import search
    row = next(r for r in golden["at_scale"] if r["n"] == 10 ** 9)
    T = int(golden["trisum"][-1]["value"])
    out = search.search(T, 10 ** 9, power=2)
    assert out.first() is None and out.bracket() == (10 ** 9, 10 ** 9 + 1)
    assert out.last_prime == row["p_n"] == 22801763489
    assert str(out.powers[0].final_sum) == row["sums"]["2"]
    assert out.powers[0].final_sum % 6 == (10 ** 9 + 5) % 6
    # every Appendix-B checkpoint (n = 1e8 .. 1e9) through a synthetic hit target: match index, p_n, S(n-1) < T
    for r in golden["at_scale"]:
        tgt = int(r["sums"]["2"])
        o = search.search(tgt, 10 ** 9, power=2)
        assert o.first() == (r["n"], 2), r["n"]
        po = o.powers[0]
        assert po.cross_prime == r["p_n"] and po.sum_at_cross == tgt
        assert po.sum_before_cross == tgt - r["p_n"] ** 2
        assert str(po.final_sum) == row["sums"]["2"]


def test_at_scale_checkpoints_all_powers(handle, golden):
    """C5 shape at every Appendix-B checkpoint: S_1..S_5(n) and p_n for n = 1e8, 2e8, .., 1e9 (all powers in one pass),
    and a synthetic hit of the all-powers search on each checkpoint (power rotating 1..5)"""
    from pss_b200 import search
    T = int(golden["trisum"][-1]["value"])
    rows = sorted(golden["at_scale"], key=lambda r: r["n"])
    for i, r in enumerate(rows):
        out = search.search(T, r["n"], power=5, n_min=min(5 * 10 ** 7, r["n"]), all_powers=True)
        assert out.last_prime == r["p_n"], r["n"]
        assert [str(p.final_sum) for p in out.powers] == [r["sums"][str(k)] for k in range(1, 6)], r["n"]
        assert out.first() is None
        k = 1 + i % 5
        o = search.search(int(r["sums"][str(k)]), 10 ** 9, power=5, all_powers=True)
        po = next(p for p in o.powers if p.power == k)
        assert (po.match_count, po.first_match_n, po.cross_n, po.cross_prime) == (1, r["n"], r["n"], r["p_n"]), (r["n"], k)
        assert [str(p.final_sum) for p in o.powers] == [rows[-1]["sums"][str(j)] for j in range(1, 6)]


def test_at_scale_prime_buffer_sha256(handle, golden):
    """the u64 prime buffer of the first 1e9 primes against the SHA-256 rows of SURVEY Appendix B (n = 1e8 .. 1e9),
    hashed window by window (value windows of 2^31 integers through pss_generate_primes: sieve -> k_compact_lut -> D2H)"""
    rows = sorted(golden["at_scale"], key=lambda r: r["n"])
    hsh = hashlib.sha256()
    done, lo, idx = 0, 0, 0
    end = rows[-1]["p_n"] + 1
    while idx < len(rows):
        assert lo < end, "ran out of range before the last checkpoint"
        hi = min(lo + (1 << 31), end)
        p = np.ascontiguousarray(handle.generate_primes(lo, hi), dtype="<u8")
        pos = 0
        while idx < len(rows) and rows[idx]["n"] <= done + len(p):
            take = rows[idx]["n"] - done
            hsh.update(p[pos:take].tobytes())
            pos = take
            assert int(p[take - 1]) == rows[idx]["p_n"], rows[idx]["n"]
            assert hsh.copy().hexdigest() == rows[idx]["sha256_primes_le_u64"], rows[idx]["n"]
            idx += 1
        hsh.update(p[pos:].tobytes())
        done += len(p)
        lo = hi
    assert done == 10 ** 9


def test_at_scale_c4_c5(handle, golden):
    """C4: k=3 n<=5e7, k=4 n<=1e7, k=5 n<=5e6;  C5 shape: all powers 1..5 at n = 1e9"""
    from pss_b200 import search
    T = int(golden["trisum"][-1]["value"])
    ex = golden["at_scale_extra"]
    out = search.search(T, 5 * 10 ** 7, power=3)
    assert str(out.powers[0].final_sum) == ex["S_5e7"][2] and out.first() is None and out.last_prime == ex["p_5e7"]
    out = search.search(T, 5 * 10 ** 6, power=5)
    assert str(out.powers[0].final_sum) == ex["S_5e6"][4] and out.first() is None
    p1e7 = {r["k"]: r["value"] for r in golden["primesum"] if r["n"] == 10 ** 7}
    out = search.search(T, 10 ** 7, power=4)
    assert str(out.powers[0].final_sum) == p1e7[4] and out.first() is None
    row = next(r for r in golden["at_scale"] if r["n"] == 10 ** 9)
    out = search.search(T, 10 ** 9, power=5, n_min=5 * 10 ** 7, all_powers=True)
    assert [str(p.final_sum) for p in out.powers] == [row["sums"][str(k)] for k in range(1, 6)]
    assert out.first() is None


def test_prime_io_export_and_file_power_sum(handle, orc, tmp_path):
    """.npy files either side of the path (reference: utils/prime_io.py:74-136): device export in value windows ==
    oracle primes; power sums over a file in index windows == oracle sums"""
    from pss_b200 import prime_io
    f = tmp_path / "primes.npy"
    n = prime_io.export_primes(f, n=300_000, window_values=1_000_000)
    exp = orc.c_first_n_primes(300_000)
    got = np.load(f)
    assert n == 300_000 and got.dtype == np.int64 and np.array_equal(got.astype(np.uint64), exp)
    assert np.array_equal(prime_io.load_primes_range(f, 1000, 1010).astype(np.uint64), exp[1000:1010])
    g = tmp_path / "window.npy"
    m = prime_io.export_primes(g, limit=10 ** 9 + 2_000_000, lo=10 ** 9, window_values=700_000)
    expw = orc.c_primes_range(10 ** 9, 10 ** 9 + 2_000_000)
    assert m == len(expw) and np.array_equal(np.load(g).astype(np.uint64), expw)
    for k in (1, 2, 5):
        assert prime_io.power_sum_file(f, k, window=70_000) == orc.c_power_sums(exp, [k])[0]
    assert prime_io.power_sum_file(f, 3, window=50_000, start_idx=12_345, end_idx=222_222) == orc.c_power_sums(exp[12_345:222_222], [3])[0]


def test_multi_target_and_sequence_rhs(handle, orc):
    """pss_search_targets: several targets in one device pass == one pss_search per target; sequence right-hand sides
    (fibonacci / factorial / catalan / tri over an m range) against a brute-force oracle scan"""
    from pss_b200 import fastpath, search
    n_max = 200_000
    primes = orc.c_first_n_primes(n_max)
    for power, allp in ((1, False), (2, False), (3, True)):
        sums = orc.c_prefix_sums(primes, power)
        targets = [sums[0], sums[6], sums[6] + 1, sums[99_999], sums[-1], sums[-1] + 1, 0, 10 ** 40]
        for op in ("==", "<=", ">"):
            many = search.search_many(targets, n_max, power=power, op=op, all_powers=allp)
            for t, o in zip(targets, many):
                one = search.search(t, n_max, power=power, op=op, all_powers=allp)
                assert o.first() == one.first() and o.last_prime == one.last_prime, (power, op, t)
                assert [(p.match_count, p.first_match_n, p.last_match_n, p.cross_n, p.final_sum) for p in o.powers] == \
                       [(p.match_count, p.first_match_n, p.last_match_n, p.cross_n, p.final_sum) for p in one.powers]
    # brute force: which primesum(n,k) values are members of the sequence?
    for fn, power in (("fibonacci", 1), ("fibonacci", 2), ("factorial", 1), ("catalan", 1), ("tri", 2)):
        sums = orc.c_prefix_sums(primes[:5000], power)
        f = search.RHS_SEQUENCES[fn]
        exp = []
        for m in range(0, 400):
            v = f(m)
            if v > sums[-1]:
                break
            lo = int(np.searchsorted(np.array(sums, dtype=object), v))
            if lo < len(sums) and sums[lo] == v:
                exp.append({"m": m, "n": lo + 1})
        got = search.for_any_sequence(fn, 5000, 399, power=power, m_min=0)
        assert got == exp, (fn, power, got, exp)
    # primesum(n,1): 2 = F(3), 5 = F(5); through the recogniser with the reference's output format
    lines, rc = fastpath.try_fast_query("for_any primesum(n,1) == fibonacci(m)", {"n": 1000, "m": 60})
    assert rc == 0 and lines[:2] == ["Found: m=3, n=1", "Found: m=5, n=2"]
    lines, rc = fastpath.try_fast_query("does_exist primesum(n,2) == factorial(m)", {"n": 1000, "m": 30})
    assert (lines, rc) == (["No match found within bounds."], 1) or rc == 0
