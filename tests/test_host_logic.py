"""Host-side logic of the drop-in layer, no GPU: API mirror semantics and errors, slice partitioning, carry
composition, outcome merging / enumeration order, checkpoint format, and the N>1 path over gloo (world_size 2)
with a handle stand-in backed by the oracle."""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# --------------------------------------------------------------------------- API mirror (no device needed)
def test_sieve_facade_semantics_without_device():
    from pss_b200 import sieve
    # edge cases never reach the device (utils/sieve.py:473-474, 558-559, 605-606)
    for limit in (-5, 0, 1):
        p = sieve.generate_primes(limit)
        assert p.dtype == np.int64 and len(p) == 0
    assert len(sieve.generate_n_primes(0)) == 0 and len(sieve.generate_n_primes(-3)) == 0
    assert len(sieve.generate_primes_range(10, 10)) == 0 and len(sieve.generate_primes_range(0, 1)) == 0
    assert sieve.prime_count(1) == 0
    with pytest.raises(ValueError, match=r"n must be positive \(1-indexed\)"):
        sieve.nth_prime(0)
    with pytest.raises(ValueError, match="Invalid algorithm: nope"):
        sieve.configure(algorithm="nope")
    with pytest.raises(ValueError, match="Unknown algorithm"):
        sieve.generate_primes(100, algorithm="nope")
    with pytest.raises(ValueError, match="primesieve requested but not installed"):   # utils/sieve.py:492-498 wording
        sieve.generate_primes(100, algorithm="primesieve")
    sieve.configure(algorithm="b200", max_memory_mb=123, prefer="gpu")
    assert sieve.get_config() == {"algorithm": "b200", "max_memory_mb": 123, "prefer": "gpu"}
    sieve.configure(algorithm="auto")
    info = sieve.get_algorithm_info()
    assert info["b200"]["available"] and info["auto"]["available"] and info["individual"]["available"]
    assert not info["primesieve"]["available"] and set(info) == set(sieve.VALID_ALGORITHMS)
    assert sieve.get_available_algorithms() == ["auto", "b200", "basic", "segmented", "individual"]


def test_primesum_argument_errors_without_device():
    from pss_b200 import primesum
    assert primesum(0, 2) == 0 and primesum(0, 5) == 0 and primesum(9, 0) == 9 and primesum(9.0, 0) == 9
    with pytest.raises(ValueError, match=r"primesum\(\) requires non-negative n, got -1"):
        primesum(-1, 2)
    with pytest.raises(ValueError, match=r"primesum\(\) requires integral value, got 7.5"):
        primesum(7.5, 2)
    with pytest.raises(TypeError, match=r"primesum\(\) requires int, got str"):
        primesum("7", 2)
    with pytest.raises(ValueError, match="power"):
        primesum(5, 9)
    import inspect
    sig = inspect.signature(primesum)
    required = [p for p in sig.parameters.values() if p.default is inspect.Parameter.empty and p.kind == p.POSITIONAL_OR_KEYWORD]
    assert len(required) == 1            # registry arg_count == 1 (tests/test_sequences.py:279-283)


def test_plugin_file_exposes_user_functions():
    import importlib.util
    path = os.path.join(ROOT, "prime-square-sum_b200", "pss_b200", "plugin.py")
    spec = importlib.util.spec_from_file_location("pss_plugin_under_test", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    public = [n for n in dir(mod) if not n.startswith("_") and callable(getattr(mod, n)) and not isinstance(getattr(mod, n), type)]
    assert sorted(public) == ["nth_prime", "prime_count", "primesum"]
    assert mod.primesum(0, 2) == 0


@pytest.mark.skipif(not os.path.exists("/root/reference/utils/function_registry.py"), reason="reference tree not present")
def test_plugin_loads_into_the_reference_registry():
    """the reference's own loader accepts the plugin and `user.primesum` wins unqualified lookup
    (utils/function_registry.py:134,499-551)"""
    sys.path.insert(0, "/root/reference")
    try:
        from utils.function_registry import FunctionRegistry
        reg = FunctionRegistry()
        n = reg.load_from_file(os.path.join(ROOT, "prime-square-sum_b200", "pss_b200", "plugin.py"))
        assert n >= 3
        f = reg.get("primesum")
        assert f(0, 2) == 0
        assert "pss_b200" in sys.modules[f.__module__].__file__ or f.__module__ != "utils.sequences"
    finally:
        sys.path.remove("/root/reference")


def test_limb_conversion_roundtrip():
    from pss_b200 import _lib
    T = 37005443752611483714216385166550857181329086284892731078593232926279977894581784762614450464857290
    limbs = _lib.int_to_limbs(T)
    assert len(limbs) == 11 and _lib.limbs_to_int(limbs) == T         # 325 bits = 11 limbs
    assert _lib.limbs_to_int(_lib.int_to_limbs(0)) == 0
    assert len(_lib.int_to_limbs(5, 12)) == 12
    with pytest.raises(OverflowError):
        _lib.int_to_limbs(1 << 400, 12)
    with pytest.raises(ValueError):
        _lib.int_to_limbs(-1)


# --------------------------------------------------------------------------- partitioning / composition
def test_partition_range_is_contiguous_aligned_and_balanced(orc):
    from pss_b200 import search
    hi = 22801763490
    for parts in (1, 2, 4, 8):
        sl = search.partition_range(hi, parts)
        assert len(sl) == parts and sl[0][0] == 0 and sl[-1][1] == hi
        for (a, b), (c, d) in zip(sl, sl[1:]):
            assert b == c and b % 480 == 0
    # prime-count balance (sieve_weight = 0): li(x) estimate vs the true pi(x) at 5e8 primes (p_{5e8} = 11 037 271 757)
    two = search.partition_range(hi, 2, sieve_weight=0.0)
    assert abs(two[0][1] - 11037271757) / 11037271757 < 0.002
    # default = estimated device time (sieve and block sums both go with the slice WIDTH; the upper slice costs a little
    # more per integer): a cut above the equal-count one and just above the middle
    two_t = search.partition_range(hi, 2)
    assert 11037271757 < 0.5 * hi < two_t[0][1] < 0.54 * hi
    eight = search.partition_range(hi, 8)
    widths = [b - a for a, b in eight]
    assert all(x >= y for x, y in zip(widths, widths[1:])) and widths[0] / widths[-1] < 1.35
    # small ranges collapse to one slice
    assert search.partition_range(1000, 4) == [(0, 1000)] + [(1000, 1000)] * 3
    # exact balance check on a range the oracle can count
    sl = search.partition_range(2 * 10 ** 7, 4, sieve_weight=0.0)
    counts = [orc.c_count_range(a, b) for a, b in sl]
    assert max(counts) / min(counts) < 1.02


def test_combine_slices_is_an_exclusive_prefix():
    from pss_b200 import search
    counts = [5, 0, 7, 3]
    sums = [[10, 100], [0, 0], [20, 200], [5, 50]]
    assert search.combine_slices(counts, sums, 0) == (0, [0, 0])
    assert search.combine_slices(counts, sums, 2) == (5, [10, 100])
    assert search.combine_slices(counts, sums, 3) == (12, [30, 300])


def _po(power, cnt=0, first=0, last=0, cross=0, prime=0, at=0, before=0, final=0):
    from pss_b200.search import PowerOutcome
    return PowerOutcome(power, cnt, first, last, cross, prime, at, before, final)


def test_merge_and_enumeration_order():
    from pss_b200.search import SearchOutcome, merge_outcomes
    a = SearchOutcome(1, 20, "==", 8944, 29, [_po(1, final=129), _po(2, final=3000), _po(3, 1, 7, 7, 7, 17, 8944, 4031, 9000)], {"ms_sieve": 1.0, "primes": 10})
    b = SearchOutcome(1, 20, "==", 8944, 71, [_po(1, final=639), _po(2, final=30000), _po(3, final=90000)], {"ms_sieve": 2.0, "primes": 10})
    m = merge_outcomes([a, None, b], 1, 20, "==", 8944)
    assert m.first() == (7, 3) and m.last_prime == 71 and m.powers[2].final_sum == 90000
    assert m.stats["ms_sieve"] == 2.0 and m.stats["primes"] == 20
    assert m.bracket(3) == (6, 7) and m.bracket(1) == (20, 21)
    # n outer / p inner (utils/grammar.py:939,977): ranges for "<" enumerate (n, p) in that order
    c = SearchOutcome(1, 5, "<", 40, 11, [_po(1, 5, 1, 5, 0, 0, 0, 0, 28), _po(2, 3, 1, 3, 4, 7, 87, 38, 208)], {})
    assert list(c.matches()) == [(1, 1), (1, 2), (2, 1), (2, 2), (3, 1), (3, 2), (4, 1), (5, 1)]
    assert list(c.matches(limit=3)) == [(1, 1), (1, 2), (2, 1)]
    # "!=" with the single equal index carved out
    d = SearchOutcome(1, 6, "!=", 10, 13, [_po(1, 5, 1, 6, 3, 5, 10, 5, 41)], {})
    assert [n for n, _ in d.matches()] == [1, 2, 4, 5, 6]


def test_li_estimate():
    from pss_b200.search import li
    assert abs(li(1e9) - 50847534) / 50847534 < 2e-4
    assert abs(li(2.28e10) - 1e9) / 1e9 < 2e-4
    assert li(2) == 0.0


# --------------------------------------------------------------------------- checkpoint format
def test_sum_cache_checkpoint_matches_reference_format(tmp_path, orc):
    from pss_b200.sum_cache import IncrementalSumCache
    c = IncrementalSumCache(powers=[2, 3])
    for p in orc.first_n_primes(15):
        c.add_prime(np.int64(p))        # np.int64 input must not wrap
    assert c.get_sum(2) == 10466        # tests/test_sum_cache.py:235-248
    c.add_prime(1_000_000_007)
    assert c.get_sum(3) == orc.power_sum(list(orc.first_n_primes(15)) + [1_000_000_007], 3)
    with pytest.raises(ValueError, match="Power 5 not tracked"):
        c.get_sum(5)
    path = tmp_path / "sub" / "sums.npz"
    c.save_checkpoint(path)
    data = np.load(path, allow_pickle=True)
    assert set(data.files) == {"metadata", "initial_values", "recent_values", "sum_p2", "sum_p3"}
    import json
    md = json.loads(str(data["metadata"]))
    assert md["version"] == "1.0" and md["powers"] == [2, 3] and md["prime_count"] == 16 and md["last_prime"] == 1_000_000_007
    d = IncrementalSumCache.load_checkpoint(path)
    assert d.sums == c.sums and d.get_prime_count() == 16 and d.get_last_prime() == 1_000_000_007
    # adaptive cadence (utils/sum_cache.py:143-167)
    assert not d.should_checkpoint()
    with pytest.raises(FileNotFoundError):
        IncrementalSumCache.load_checkpoint(tmp_path / "missing.npz")
    if os.path.exists("/root/reference/utils/sum_cache.py"):
        sys.path.insert(0, "/root/reference")
        try:
            from utils.sum_cache import IncrementalSumCache as Ref
            r = Ref.load_checkpoint(path)       # the reference reads our file
            assert r.sums == c.sums and r.get_prime_count() == 16
            r.save_checkpoint(tmp_path / "ref.npz")
            back = IncrementalSumCache.load_checkpoint(tmp_path / "ref.npz")   # and we read the reference's
            assert back.sums == c.sums
        finally:
            sys.path.remove("/root/reference")


def test_sum_cache_advance_equals_per_prime_path(tmp_path, orc):
    """sum_cache.advance (one fused device pass; here through the oracle stand-in handle) leaves exactly the state the
    per-prime path leaves: counts, last prime, sums, head and tail windows — on this package's class and, when the
    reference tree is present, on an instance of the reference's own class (state in / state out)."""
    from pss_b200 import sum_cache
    h = OracleHandle()
    primes = [int(p) for p in orc.c_first_n_primes(3000)]
    classes = [sum_cache.IncrementalSumCache]
    if os.path.exists("/root/reference/utils/sum_cache.py"):
        sys.path.insert(0, "/root/reference")
        try:
            from utils.sum_cache import IncrementalSumCache as Ref
            classes.append(Ref)
        finally:
            sys.path.remove("/root/reference")
    for cls in classes:
        for powers in ([2, 3, 4], [2], [5, 1], [0, 3], [7, 2]):
            for start, stops in ((0, (1, 2, 3, 700, 3000)), (1, (2, 50)), (2, (3, 4)), (40, (41, 42, 45, 2999))):
                ref, dev = cls(powers=list(powers)), cls(powers=list(powers))
                for p in primes[:start]:
                    ref.add_prime(p)
                    dev.add_prime(p)
                done = start
                for stop in stops:
                    for p in primes[done:stop]:
                        ref.add_prime(p)
                    done = stop
                    sum_cache.advance(dev, stop, handle=h)
                    assert dev.sums == ref.sums, (cls.__name__, powers, start, stop)
                    assert (dev.metadata.prime_count, dev.metadata.last_prime) == (stop, primes[stop - 1])
                    assert [tuple(r) for r in dev.initial_values] == [tuple(r) for r in ref.initial_values]
                    assert [tuple(r) for r in dev.recent_values] == [tuple(r) for r in ref.recent_values], (cls.__name__, powers, start, stop)
                sum_cache.advance(dev, 1, handle=h)           # going backwards is a no-op
                assert dev.metadata.prime_count == done
    # numpy powers must serialise (the reference converts them)
    c = sum_cache.IncrementalSumCache(powers=list(np.arange(2, 4)))
    c.add_prime(np.int64(2))
    c.save_checkpoint(tmp_path / "np.npz")
    assert sum_cache.IncrementalSumCache.load_checkpoint(tmp_path / "np.npz").sums == {2: 4, 3: 8}
    # old JSON checkpoints (utils/sum_cache.py:324-368)
    import json
    (tmp_path / "old.json").write_text(json.dumps({"primes_processed": 7, "last_prime": 17, "current_sum": "666"}))
    sum_cache.migrate_old_checkpoint(tmp_path / "old.json", tmp_path / "new.npz")
    m = sum_cache.IncrementalSumCache.load_checkpoint(tmp_path / "new.npz")
    assert (m.get_sum(2), m.get_prime_count(), m.get_last_prime()) == (666, 7, 17)
    m.report()


# --------------------------------------------------------------------------- N > 1 over gloo
class OracleHandle:
    """Stand-in for _lib.Handle implementing the fused slice API with the oracle (host logic tests only)."""

    def __init__(self):
        from oracle import oracle
        self.orc = oracle
        self.primes = None
        self.cfg = None

    def nth_prime_upper(self, n):
        if n < 6:
            return 15
        b = n * (math.log(n) + math.log(math.log(n)) - (0.9484 if n >= 39017 else 0.0))
        return int(b) + 3

    def generate_primes(self, lo, hi):
        return self.orc.c_primes_range(lo, hi).astype(np.int64)

    def nth_prime(self, n):
        return int(self.orc.c_first_n_primes(n)[-1])

    def slice_sieve_sums(self, lo, hi, power_max, all_powers=False):
        self.primes = self.orc.c_primes_range(lo, hi)
        powers = list(range(1, power_max + 1)) if (all_powers and power_max > 1) else [power_max]
        self.cfg = powers
        return len(self.primes), self.orc.c_power_sums(self.primes, powers)

    def search(self, n_min, n_max, power_max, target, op="==", all_powers=False):
        self.primes = self.orc.c_first_n_primes(n_max)
        self.cfg = list(range(1, power_max + 1)) if (all_powers and power_max > 1) else [power_max]
        return self.slice_scan_bitmap(1, [0] * len(self.cfg), n_min, n_max, power_max, target, op, all_powers)

    def search_targets(self, n_min, n_max, power_max, targets, op="==", all_powers=False):
        return [self.search(n_min, n_max, power_max, t, op, all_powers) for t in targets]

    def search_tri(self, n_min, n_max, power, m_min, m_max, cap=4096):
        sums = self.orc.c_prefix_sums(self.orc.c_first_n_primes(n_max), power)
        pairs = []
        for i, s_ in enumerate(sums):
            m = self.orc.qtri(s_)
            if i + 1 >= n_min and m is not None and max(m_min, 1) <= m <= m_max:
                pairs.append((i + 1, m))
        return pairs, {}

    def slice_scan_bitmap(self, n_first, carry, n_min, n_max, power_max, target, op="==", all_powers=False):
        from pss_b200 import _lib
        take = min(len(self.primes), n_max - n_first + 1)
        res = _lib.SearchResult()
        res.n_powers = len(self.cfg)
        res.primes_scanned = take
        res.last_prime = int(self.primes[take - 1]) if n_first + take - 1 == n_max else 0
        for j, k in enumerate(self.cfg):
            r = self.orc.c_scan_compare(self.primes[:take], n_first, k, carry[j], target, op, n_min, n_max)
            pr = res.powers[j]
            pr.power = k
            pr.match_count, pr.first_match_n, pr.last_match_n = r["match_count"], r["first_match_n"], r["last_match_n"]
            pr.cross_n, pr.cross_prime = r["cross_n"], r["cross_prime"]
            for name in ("sum_at_cross", "sum_before_cross", "final_sum"):
                limbs = _lib.int_to_limbs(r[name], _lib.PSS_LIMBS)
                for i in range(_lib.PSS_LIMBS):
                    getattr(pr, name)[i] = int(limbs[i])
        return res


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(ROOT, "prime-square-sum_b200"))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from pss_b200 import search
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        h = OracleHandle()
        from oracle import oracle as orc
        n_max = 60000
        primes = orc.c_first_n_primes(n_max)
        out = {}
        # a hit in rank 1's slice, one in rank 0's, a miss, and an all-powers query
        for name, (k, multi, hit) in {"hi": (2, False, 45000), "lo": (3, False, 7), "multi": (4, True, 31000)}.items():
            target = orc.c_prefix_sums(primes[:hit], k)[-1]
            o = search.distributed_search(target, n_max, power=k, all_powers=multi, handle=h)
            out[name] = (o.first(), [p.final_sum for p in o.powers], o.last_prime, o.powers[-1].sum_before_cross)
        T = orc.trisum(666)
        o = search.distributed_search(T, n_max, power=2, handle=h)
        out["miss"] = (o.first(), o.bracket(), o.powers[0].final_sum, o.last_prime)
        o = search.distributed_search(10 ** 12, n_max, power=2, n_min=50, op="<", handle=h)
        out["lt"] = (o.powers[0].match_count, o.powers[0].first_match_n, o.powers[0].last_match_n)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_distributed_search_world_size_2_gloo(orc):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_max = 60000
    primes = orc.c_first_n_primes(n_max)
    S = {k: orc.c_prefix_sums(primes, k) for k in (1, 2, 3, 4)}
    for rank in (0, 1):
        out = results[rank]                      # every rank returns the merged outcome
        assert out["hi"] == ((45000, 2), [S[2][-1]], int(primes[-1]), S[2][44998])
        assert out["lo"] == ((7, 3), [S[3][-1]], int(primes[-1]), S[3][5])
        assert out["multi"][0] == (31000, 4) and out["multi"][1] == [S[k][-1] for k in (1, 2, 3, 4)]
        assert out["miss"] == (None, (n_max, n_max + 1), S[2][-1], int(primes[-1]))
        below = sum(1 for n in range(50, n_max + 1) if S[2][n - 1] < 10 ** 12)
        assert out["lt"] == (below, 50, 49 + below)


# --------------------------------------------------------------------------- fast-path recogniser (host part)
def test_fastpath_recogniser_host_side(orc):
    from pss_b200 import fastpath
    for b in (1, 2, 3, 6, 10, 15, 28, 100, 666):
        assert fastpath._trisum(b) == orc.trisum(b)
    assert fastpath._trisum(666).bit_length() == 325
    # shapes that are NOT the hot path stay with the generic evaluator
    for expr in ("fibonacci(n) == 8", "primesum(n,2) + 1 == 667", "does_exist primesum(n,2) == fibonacci(m) + 1",
                 "does_exist primesum(n,2) < fibonacci(m)", "does_exist primesum(n,2) == fibonacci(n)",
                 "for_any primesum(n,2) < tri(m)", "verify primesum(7,2) == 666", "primesum(n, n) == 4",
                 "primesum(n,2) == 666 == 666", "does_exist primesum(n,9) == 4", "primesum(n,p) == 4 and n > 2"):
        assert fastpath.try_fast_query(expr, {"n": 10, "m": 10}) is None, expr
    from pss_b200 import search
    assert [search.RHS_SEQUENCES["fibonacci"](i) for i in range(11)] == [0, 1, 1, 2, 3, 5, 8, 13, 21, 34, 55]     # utils/sequences.py:83-117
    assert [search.RHS_SEQUENCES["catalan"](i) for i in range(10)] == [1, 1, 2, 5, 14, 42, 132, 429, 1430, 4862]  # :160-200
    assert [search.RHS_SEQUENCES["factorial"](i) for i in (0, 1, 5, 10)] == [1, 1, 120, 3628800]                  # :120-157
    Q = fastpath.Query
    table = {
        "for_any primesum(n,p) >= 8944": Q("for_any", "n", ">=", pvar="p", target=8944),
        "primesum(n,2) == 666": Q("does_exist", "n", "==", power=2, target=666),
        "does_exist primesum(n) == 666": Q("does_exist", "n", "==", power=2, target=666),
        "does_exist (primesum(n, 2)) == (666)": Q("does_exist", "n", "==", power=2, target=666),
        "does_exist (primesum(n,2) == 666)": Q("does_exist", "n", "==", power=2, target=666),
        "does_exist pss.primesum(k,3) != trisum(10)": Q("does_exist", "k", "!=", power=3, target=666),
        "does_exist trisum(666) == primesum(n,2)": Q("does_exist", "n", "==", power=2, target=orc.trisum(666)),
        "for_any 100 < primesum(n,2)": Q("for_any", "n", ">", power=2, target=100),
        "for_any 100 >= primesum(n,1)": Q("for_any", "n", "<=", power=1, target=100),
        "for_any tri(m) == primesum(n,2)": Q("for_any", "n", "==", power=2, seq_fn="tri", mvar="m"),
        "for_any primesum(n,1) == pss.fibonacci(j)": Q("for_any", "n", "==", power=1, seq_fn="fibonacci", mvar="j"),
        "primesum(n,2) == factorial(5)": Q("does_exist", "n", "==", power=2, target=120),
    }
    for text, want in table.items():
        tree = fastpath.parse(text)
        got = fastpath.recognise(tree) if tree is not None else None
        assert got == want, (text, got)
    assert fastpath.format_match({"n": 7, "m": 36}) == "Found: m=36, n=7"                      # utils/cli.py:1253,1271
    assert fastpath.format_match({"n": 7}, "json") == '{"found": true, "variables": {"n": 7}}'  # tests/test_cli_integration.py:257-267
    assert fastpath.format_match({"n": 7}, "csv") == "n=7"
    assert fastpath.format_no_match() == "No match found within bounds."
    assert fastpath.format_no_match("json") == '{"found": false, "variables": null}'
    with pytest.raises(ValueError, match=r"Missing bounds for variable\(s\): q. Use --max q:VALUE to specify."):
        fastpath.try_fast_query("primesum(q,2) == 4")
    # free power variable beyond the device's all-powers pass: generic path, not an error
    h = OracleHandle()
    q = fastpath.recognise(fastpath.parse("for_any primesum(n,p) == 8944"))
    assert fastpath.matches(q, {"n": (1, 20), "p": (1, 7)}, handle=h) is None
    assert fastpath.matches(q, {"n": (1, 20), "p": (1, 3)}, handle=h) == [{"n": 7, "p": 3}]
    assert fastpath.matches(q, {"n": (8, 20), "p": (1, 3)}, handle=h) == []


def test_prime_io_window_plan():
    """value windows of the .npy export: contiguous, cover [lo, limit), interior cuts on multiples of 480"""
    from pss_b200 import prime_io
    for lo, limit, w in [(0, 10_000, 960), (7, 100_003, 4_801), (10 ** 9, 10 ** 9 + 12_345, 5_000), (0, 0, 1000), (100, 50, 1000)]:
        wins = list(prime_io.plan_value_windows(limit, w, lo))
        if limit <= lo:
            assert wins == []
            continue
        assert wins[0][0] == lo and wins[-1][1] == limit
        for (a, b), (c, d) in zip(wins, wins[1:]):
            assert b == c and a < b and b % 480 == 0
        assert all(b - a <= w for a, b in wins)
    with pytest.raises(ValueError):
        list(prime_io.plan_value_windows(10, 100))
