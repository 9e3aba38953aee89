"""
search.py — the fused query path: `does_exist / for_any  primesum(n, p) <op> <constant>`.

Replaces the reference's O(n^2) host loop (utils/grammar.py:937-999 find_matches calling
utils/sequences.py:30-75 once per n) by ONE device pass: sieve -> compaction -> power + scan + compare.
Results are reported exactly as the reference would enumerate them: variables iterate alphabetically,
first variable outermost (utils/grammar.py:939,977), so `primesum(n,p)` walks n outer / p inner and
`does_exist` stops at the first hit (:987-988).

Multi-GPU: the value range [2, N) is cut into G contiguous slices balanced by the
li(x) prime-count estimate; every rank sieves and reduces its slice independently, ONE tiny allgather
of (count, per-power sums) over NCCL gives each rank its index offset and limb carry, then the rank scans
its slice with that carry.  Nothing else crosses NVLink.
"""
from __future__ import annotations

import functools
import math
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

from . import _lib

OPS = ("==", "!=", "<", "<=", ">", ">=")


@dataclass
class PowerOutcome:
    power: int
    match_count: int
    first_match_n: int          # 0 = none
    last_match_n: int
    cross_n: int                # first n with S(n) >= target (0: every scanned S(n) < target)
    cross_prime: int
    sum_at_cross: int
    sum_before_cross: int
    final_sum: int              # S(n_max)


@dataclass
class SearchOutcome:
    n_min: int
    n_max: int
    op: str
    target: int
    last_prime: int
    powers: List[PowerOutcome]
    stats: Dict[str, float] = field(default_factory=dict)

    def matches(self, limit: Optional[int] = None) -> Iterator[Tuple[int, int]]:
        """(n, power) pairs in the reference's iteration order (n outer, p inner)."""
        yield from _enumerate_matches(self, limit)

    def first(self) -> Optional[Tuple[int, int]]:
        for m in self.matches(limit=1):
            return m
        return None

    def bracket(self, power: Optional[int] = None) -> Tuple[int, int]:
        """(n_below, n_at_or_above): largest n with S(n) < target and the following index (n_max+1 if none)."""
        po = self.powers[0] if power is None else next(p for p in self.powers if p.power == power)
        if po.cross_n:
            return po.cross_n - 1, po.cross_n
        if po.final_sum >= self.target:
            # no crossing inside the scan although S(n_max) >= target: the running sum was already >= target before the
            # first scanned prime (target = 0, or a resumed / sliced scan whose carry-in is past the target)
            first = int(self.stats.get("resumed_from_n", 0)) + 1
            return first - 1, first
        return self.n_max, self.n_max + 1


def _cmp(a: int, b: int) -> int:
    return (a > b) - (a < b)


_HOLDS = {"==": lambda c: c == 0, "!=": lambda c: c != 0, "<": lambda c: c < 0, "<=": lambda c: c <= 0,
          ">": lambda c: c > 0, ">=": lambda c: c >= 0}


def _match_ranges(o: SearchOutcome, po: PowerOutcome) -> List[Tuple[int, int]]:
    """Inclusive index ranges of n where S(n) <op> target.  S is strictly increasing, so the device's
    (count, first, last, crossing) describe them completely."""
    if po.match_count == 0:
        return []
    if o.op == "!=" and po.match_count != po.last_match_n - po.first_match_n + 1:
        # everything except the single n with S(n) == target
        eq = po.cross_n
        out = []
        if po.first_match_n <= eq - 1:
            out.append((po.first_match_n, eq - 1))
        if eq + 1 <= po.last_match_n:
            out.append((eq + 1, po.last_match_n))
        return out
    return [(po.first_match_n, po.last_match_n)]


def _enumerate_matches(o: SearchOutcome, limit: Optional[int]) -> Iterator[Tuple[int, int]]:
    ranges = {po.power: _match_ranges(o, po) for po in o.powers}
    if not any(ranges.values()):
        return
    emitted = 0
    if len(o.powers) == 1:
        p = o.powers[0].power
        for lo, hi in ranges[p]:
            for n in range(lo, hi + 1):
                yield (n, p)
                emitted += 1
                if limit is not None and emitted >= limit:
                    return
        return
    # several powers: merge by n (outer) then p (inner)
    lo_all = min(r[0][0] for r in ranges.values() if r)
    hi_all = max(r[-1][1] for r in ranges.values() if r)
    n = lo_all
    while n <= hi_all:
        hit = False
        for po in o.powers:
            if any(lo <= n <= hi for lo, hi in ranges[po.power]):
                hit = True
                yield (n, po.power)
                emitted += 1
                if limit is not None and emitted >= limit:
                    return
        if not hit:
            # jump to the next range start
            nxt = [lo for r in ranges.values() for lo, _ in r if lo > n]
            if not nxt:
                return
            n = min(nxt)
        else:
            n += 1


def _outcome_from_result(res: "_lib.SearchResult", n_min, n_max, op, target) -> SearchOutcome:
    powers = []
    for j in range(res.n_powers):
        r = res.powers[j]
        powers.append(PowerOutcome(
            power=int(r.power), match_count=int(r.match_count), first_match_n=int(r.first_match_n),
            last_match_n=int(r.last_match_n), cross_n=int(r.cross_n), cross_prime=int(r.cross_prime),
            sum_at_cross=_lib.limbs_to_int(r.sum_at_cross), sum_before_cross=_lib.limbs_to_int(r.sum_before_cross),
            final_sum=_lib.limbs_to_int(r.final_sum)))
    stats = _lib.stats_to_dict(res.stats)
    return SearchOutcome(n_min, n_max, op, target, int(res.last_prime), powers, stats)


def search(target: int, n_max: int, power: int = 2, n_min: int = 1, op: str = "==",
           all_powers: bool = False, handle: Optional["_lib.Handle"] = None, resume=None) -> SearchOutcome:
    """Single-GPU fused search.  `all_powers=True` evaluates every p in 1..power in the same pass.

    `resume`: an IncrementalSumCache-like state (this package's or the reference's class: .metadata.prime_count,
    .metadata.last_prime, .sums) that already holds S_k(c) of the first c primes.  The search then covers
    n in (c, n_max] only: the sieve starts above last_prime and the running sums start from the checkpointed
    values instead of being re-accumulated from n = 1 (reference ROADMAP #26 `--start-n`,
    utils/sum_cache.py:169-261 is the file format)."""
    if op not in OPS:
        raise ValueError(f"unknown comparison operator {op!r}")
    if target < 0:
        raise ValueError("target must be non-negative")
    h = handle or _lib.default_handle()
    if resume is not None and int(resume.metadata.prime_count) > 0:
        return _resume_search(h, resume, target, n_max, power, n_min, op, all_powers)
    res = h.search(n_min, n_max, power, target, op, all_powers)
    return _outcome_from_result(res, n_min, n_max, op, target)


def _resume_search(h, state, target, n_max, power, n_min, op, all_powers) -> SearchOutcome:
    c0, p0 = int(state.metadata.prime_count), int(state.metadata.last_prime)
    ks = list(range(1, power + 1)) if (all_powers and power > 1) else [power]
    missing = [k for k in ks if k not in state.sums]
    if missing:
        raise ValueError(f"the checkpoint does not track power(s) {missing}; tracked: {sorted(state.sums)}")
    if n_max <= c0:
        raise ValueError(f"nothing to search: the checkpoint already covers n <= {c0} (n_max = {n_max})")
    lo_n = max(int(n_min), c0 + 1)
    hi = h.nth_prime_upper(n_max) + 1
    cnt, _ = h.slice_sieve_sums(p0 + 1, hi, power, all_powers)
    if c0 + cnt < n_max:
        raise RuntimeError("prime count bound violated")
    res = h.slice_scan_bitmap(c0 + 1, [int(state.sums[k]) for k in ks], lo_n, n_max, power, target, op, all_powers)
    out = _outcome_from_result(res, lo_n, n_max, op, target)
    out.stats["resumed_from_n"] = c0
    return out


def does_exist(target: int, n_max: int, power: int = 2, n_min: int = 1, op: str = "==",
               all_powers: bool = False) -> Optional[Dict[str, int]]:
    """`does_exist primesum(n,p) <op> target`: the first match as the reference's variable dict, or None."""
    m = search(target, n_max, power, n_min, op, all_powers).first()
    if m is None:
        return None
    return {"n": m[0], "p": m[1]} if all_powers else {"n": m[0]}


def for_any(target: int, n_max: int, power: int = 2, n_min: int = 1, op: str = "==",
            all_powers: bool = False, limit: Optional[int] = None) -> Iterator[Dict[str, int]]:
    for n, p in search(target, n_max, power, n_min, op, all_powers).matches(limit):
        yield {"n": n, "p": p} if all_powers else {"n": n}


def for_any_tri(n_max: int, m_max: int, power: int = 2, n_min: int = 1, m_min: int = 1,
                handle: Optional["_lib.Handle"] = None) -> List[Dict[str, int]]:
    """`for_any primesum(n,power) == tri(m)`: the reference iterates m outer / n inner (alphabetical,
    utils/grammar.py:939,977) and prints `Found: m=36, n=7`; S(n) is strictly increasing, so ordering by m is
    ordering by n.  Every partial sum is tested for triangularity on the device."""
    h = handle or _lib.default_handle()
    pairs, _ = h.search_tri(n_min, n_max, power, m_min, m_max)
    return [{"m": m, "n": n} for n, m in sorted(pairs, key=lambda t: (t[1], t[0]))]


def search_many(targets: Sequence[int], n_max: int, power: int = 2, n_min: int = 1, op: str = "==",
                all_powers: bool = False, handle: Optional["_lib.Handle"] = None) -> List[SearchOutcome]:
    """Several targets over the same n range in ONE device pass over the primes (one sieve, one block-sum pass)
    followed by one pruned compare launch per target.  outcomes[i] belongs to targets[i]."""
    if op not in OPS:
        raise ValueError(f"unknown comparison operator {op!r}")
    if any(t < 0 for t in targets):
        raise ValueError("targets must be non-negative")
    h = handle or _lib.default_handle()
    res = h.search_targets(n_min, n_max, power, list(targets), op, all_powers)
    return [_outcome_from_result(r, n_min, n_max, op, t) for r, t in zip(res, targets)]


# monotone integer sequences of the reference's registry usable as a right-hand side (utils/sequences.py:83-200,
# utils/number_theory.py:250-289): name -> f(m)
def _fibonacci(m: int) -> int:
    if m < 0:
        raise ValueError(f"fibonacci() requires non-negative n, got {m}")
    a, b = 0, 1
    for _ in range(m):
        a, b = b, a + b
    return a


def _catalan(m: int) -> int:
    if m < 0:
        raise ValueError(f"catalan() requires non-negative n, got {m}")
    return math.comb(2 * m, m) // (m + 1)


def _factorial(m: int) -> int:
    if m < 0:
        raise ValueError(f"factorial() requires non-negative n, got {m}")
    return math.factorial(m)


RHS_SEQUENCES = {"fibonacci": _fibonacci, "factorial": _factorial, "catalan": _catalan, "tri": lambda m: m * (m + 1) // 2}


def for_any_sequence(fn: str, n_max: int, m_max: int, power: int = 2, n_min: int = 1, m_min: int = 1,
                     handle: Optional["_lib.Handle"] = None) -> List[Dict[str, int]]:
    """`for_any primesum(n,power) == fn(m)`, m in [m_min, m_max], for a non-decreasing sequence fn of the reference's
    registry: every (m, n) with primesum(n, power) == fn(m), in the reference's iteration order (m outer, n inner:
    variables sorted by name, utils/grammar.py:939,977).  The distinct values fn(m) that can be reached at all
    (fn(m) <= n_max * p_max^power) are searched in one device pass (`search_many`)."""
    if fn not in RHS_SEQUENCES:
        raise ValueError(f"unknown sequence {fn!r}")
    h = handle or _lib.default_handle()
    f = RHS_SEQUENCES[fn]
    cap = n_max * (h.nth_prime_upper(n_max) ** power)        # S_power(n_max) < cap: larger values cannot match
    values: Dict[int, List[int]] = {}
    for m in range(max(m_min, 0), m_max + 1):
        v = f(m)
        if v > cap:
            break                                            # all four sequences are non-decreasing
        values.setdefault(v, []).append(m)
    if not values:
        return []
    targets = sorted(values)
    outs = search_many(targets, n_max, power=power, n_min=n_min, op="==", handle=h)
    hits = []
    for t, o in zip(targets, outs):
        first = o.first()
        if first is not None:
            for m in values[t]:
                hits.append({"m": m, "n": first[0]})         # S strictly increasing: at most one n per value
    hits.sort(key=lambda d: (d["m"], d["n"]))
    return hits


# ------------------------------------------------------------------------------- multi-GPU
def li(x: float) -> float:
    """Offset logarithmic integral estimate of pi(x) (asymptotic series; only used to balance slices)."""
    if x < 3:
        return 0.0
    lx = math.log(x)
    s, term = 0.0, 1.0
    for k in range(1, 12):
        term *= k / lx
        if term < 1e-6:
            break
        s += term
    return x / lx * (1.0 + s)


# Cost model behind the slice cuts (measured on B200, profiles/r01 and profiles/r02): the sieve's time is proportional to the
# slice's width times a slowly growing per-integer factor (more base primes are active in higher tiles: ~ (x/N)^0.12).  The
# block-sum pass works per bitmap WORD (byte-table moments), so it is proportional to the width too — not to the prime count:
# the 8-GPU timeline of round 2 showed rank 0 (the densest primes, hence the widest slice under a prime-count model) with the
# longest block-sum pass.  Fit to that timeline (profiles/r02/bench_c3_v18_n8_timeline.json): sieve time per unit width
# = 2.565 (x/N)^0.0925 ms, block sums (k = 2) 0.51 ms per unit width, i.e. cost(f) = f + 4.6 f^1.0925 up to a constant factor;
# the all-powers block sums cost 1.28 / 0.36 times the k = 2 ones.
SIEVE_WEIGHT = 4.6
SIEVE_EXPONENT = 0.0925


def sieve_weight_for(power: int, all_powers: bool) -> float:
    """sieve : block-sum cost ratio of a query configuration (see the fit above)"""
    if all_powers and power > 1:
        return SIEVE_WEIGHT * 0.36 / (0.36 + (1.28 - 0.36) * (power - 1) / 4.0)
    return SIEVE_WEIGHT if power == 2 else 2.7


def partition_range(hi_excl: int, parts: int, lo: int = 0, sieve_weight: float = SIEVE_WEIGHT,
                    sieve_exponent: float = SIEVE_EXPONENT) -> List[Tuple[int, int]]:
    return list(_partition_range(int(hi_excl), int(parts), int(lo), float(sieve_weight), float(sieve_exponent)))


@functools.lru_cache(maxsize=64)
def _partition_range(hi_excl: int, parts: int, lo: int, sieve_weight: float, sieve_exponent: float) -> Tuple[Tuple[int, int], ...]:
    """Cut [lo, hi_excl) into `parts` contiguous slices of ~equal estimated device time; sieve_weight = 0 gives equal
    (li(x)-estimated) prime counts instead.  Interior cuts are multiples of 480 (16 bitmap bytes) so every slice starts
    on a 16-byte aligned bitmap byte."""
    if parts <= 1 or hi_excl - lo < 480 * parts * 4:
        return tuple([(lo, hi_excl)] + [(hi_excl, hi_excl)] * (parts - 1))

    if sieve_weight == 0.0:
        cost = li
    else:
        def cost(x: float) -> float:
            f = max(x, 0.0) / hi_excl
            return f + sieve_weight * f ** (1.0 + sieve_exponent)

    total = cost(hi_excl) - cost(lo)
    cuts = [lo]
    for g in range(1, parts):
        want = cost(lo) + total * g / parts
        a, b = cuts[-1], hi_excl
        for _ in range(64):
            mid = (a + b) / 2
            if cost(mid) < want:
                a = mid
            else:
                b = mid
        c = int(b) // 480 * 480
        cuts.append(min(max(c, cuts[-1]), hi_excl))
    cuts.append(hi_excl)
    return tuple((cuts[i], cuts[i + 1]) for i in range(parts))


def combine_slices(counts: Sequence[int], sums: Sequence[Sequence[int]], rank: int) -> Tuple[int, List[int]]:
    """Exclusive prefix over ranks of (count, per-power sums): the rank's index offset and limb carry."""
    n_before = sum(counts[:rank])
    np_ = len(sums[0]) if sums else 0
    carry = [sum(sums[r][j] for r in range(rank)) for j in range(np_)]
    return n_before, carry


def merge_outcomes(parts: Sequence[Optional[SearchOutcome]], n_min: int, n_max: int, op: str, target: int) -> SearchOutcome:
    """Merge per-rank scan outcomes (ranks in value order; None for ranks that had nothing to scan)."""
    parts = [p for p in parts if p is not None]
    if not parts:
        raise ValueError("no slice scanned anything")
    powers = []
    for j, p0 in enumerate(parts[0].powers):
        recs = [p.powers[j] for p in parts]
        cross = next((r for r in recs if r.cross_n), None)
        matched = [r for r in recs if r.match_count]
        powers.append(PowerOutcome(
            power=p0.power, match_count=sum(r.match_count for r in recs),
            first_match_n=min((r.first_match_n for r in matched), default=0),
            last_match_n=max((r.last_match_n for r in matched), default=0),
            cross_n=cross.cross_n if cross else 0, cross_prime=cross.cross_prime if cross else 0,
            sum_at_cross=cross.sum_at_cross if cross else 0, sum_before_cross=cross.sum_before_cross if cross else 0,
            final_sum=recs[-1].final_sum))
    stats = {}
    for k in parts[0].stats:
        vals = [p.stats.get(k, 0) for p in parts]
        stats[k] = max(vals) if k.startswith("ms_") else (sum(vals) if k in ("primes", "launches") else vals[0])
    return SearchOutcome(n_min, n_max, op, target, parts[-1].last_prime, powers, stats)


def _np(power: int, all_powers: bool) -> int:
    return power if (all_powers and power > 1) else 1


def distributed_search(target: int, n_max: int, power: int = 2, n_min: int = 1, op: str = "==",
                       all_powers: bool = False, group=None, handle: Optional["_lib.Handle"] = None,
                       gather=None) -> SearchOutcome:
    """One process per GPU (torch.distributed, NCCL).  Every rank returns the merged outcome.

    `gather(obj) -> list` may be injected (tests drive the host logic over gloo); by default the two
    exchanges go through torch.distributed: the (count, sums) allgather as a uint32 CUDA tensor over NCCL,
    the final merge as an object gather.
    """
    import time
    import torch
    import torch.distributed as dist

    t0 = time.perf_counter()
    h = handle or _lib.default_handle()
    # world / rank / transport of this (handle, group) are looked up once: the query path of a 0.4 ms multi-GPU step should not
    # pay for torch.distributed introspection every call
    import os
    exchange = os.environ.get("PSS_EXCHANGE", "peer")
    ctx = getattr(h, "_dist_ctx", None)
    if ctx is None or ctx[0] is not group or ctx[4] != exchange or gather is not None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        peer = world > 1 and gather is None and _peer_exchange_enabled(group)
        ctx = (group, world, rank, peer, exchange)
        if gather is None:
            h._dist_ctx = ctx
    _, world, rank, peer, _ = ctx
    if world == 1:
        return search(target, n_max, power, n_min, op, all_powers, handle=h)

    hi = h.nth_prime_upper(n_max) + 1
    lo_g, hi_g = partition_range(hi, world, sieve_weight=sieve_weight_for(power, all_powers))[rank]
    np_ = _np(power, all_powers)
    if peer:
        return _peer_search(h, group, rank, world, lo_g, hi_g, target, n_max, power, n_min, op, all_powers, np_)
    # fused phase A: sieve + per-tile sums (no u64 prime buffer); leaves the slice's bitmap resident
    count_g, sums_g = h.slice_sieve_sums(lo_g, hi_g, power, all_powers)
    t1 = time.perf_counter()

    # ---- the one exchange: 8 B count + np_ * 48 B of limbs per rank
    payload = [count_g & 0xFFFFFFFF, count_g >> 32]
    for s in sums_g:
        payload.extend(int(x) for x in _lib.int_to_limbs(s, _lib.PSS_LIMBS))
    dev = None
    if gather is None:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    if gather is not None:
        rows = gather(payload)
    else:
        mine = torch.tensor(payload, dtype=torch.int64, device=dev)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)       # NCCL over NVLink on the GPU box, gloo in CPU tests
        rows = [t.cpu().tolist() for t in parts]
    counts = [r[0] | (r[1] << 32) for r in rows]
    sums = [[_lib.limbs_to_int(r[2 + j * _lib.PSS_LIMBS: 2 + (j + 1) * _lib.PSS_LIMBS]) for j in range(np_)] for r in rows]
    if sum(counts) < n_max:
        raise RuntimeError("prime count bound violated")
    n_before, carry = combine_slices(counts, sums, rank)
    t2 = time.perf_counter()

    # ---- local scan with the carry; the rank holding n_max truncates, later ranks skip
    local = None
    take = min(count_g, max(0, n_max - n_before))
    if take > 0:
        res = h.slice_scan_bitmap(n_before + 1, carry, n_min, n_max, power, target, op, all_powers)
        local = _outcome_from_result(res, n_min, n_max, op, target)
    t3 = time.perf_counter()
    # ---- merge: one more fixed-size exchange (per power: 5 u64 + 3 x 12 limbs, + last_prime / stats)
    rec = _pack_outcome(local, np_)
    if gather is not None:
        rows = gather(rec)
    else:
        mine = torch.tensor(rec, dtype=torch.int64, device=dev)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        rows = [t.cpu().tolist() for t in parts]
    outs = [_unpack_outcome(r, np_, power, all_powers, n_min, n_max, op, target) for r in rows]
    merged = merge_outcomes(outs, n_min, n_max, op, target)
    t4 = time.perf_counter()
    merged.stats["host_ms"] = {"sieve_sums": (t1 - t0) * 1e3, "exchange": (t2 - t1) * 1e3, "scan": (t3 - t2) * 1e3,
                               "merge": (t4 - t3) * 1e3}
    return merged


def _peer_exchange_enabled(group) -> bool:
    """the device-side exchange needs one CUDA process per GPU on one node (CUDA IPC over NVLink); PSS_EXCHANGE=nccl
    selects the host-driven NCCL all-gathers instead"""
    import os
    import torch.distributed as dist
    if os.environ.get("PSS_EXCHANGE", "peer").lower() != "peer":
        return False
    try:
        return dist.get_backend(group) == "nccl" and dist.get_world_size(group) <= _lib.PSS_MAX_PEERS
    except Exception:
        return False


def _peer_setup(h: "_lib.Handle", group, rank: int, world: int):
    """once per handle: all-gather the mailbox IPC handles (torch.distributed plumbing), map them, barrier"""
    import torch
    import torch.distributed as dist
    if getattr(h, "_peer_group_id", None) == (id(group), world, rank):
        return
    mine = torch.frombuffer(bytearray(h.peer_export()), dtype=torch.uint8).to(torch.device("cuda", torch.cuda.current_device()))
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    h.peer_connect(rank, world, [bytes(t.cpu().numpy().tobytes()) for t in parts])
    dist.barrier(group=group)      # nobody publishes before every mailbox is mapped and zeroed
    torch.cuda.synchronize()
    h._peer_group_id = (id(group), world, rank)


def _peer_search(h, group, rank, world, lo_g, hi_g, target, n_max, power, n_min, op, all_powers, np_) -> SearchOutcome:
    """multi-GPU query with the exchange on the device: one C call, one host synchronisation per rank"""
    _peer_setup(h, group, rank, world)
    recs, stats = h.peer_search(lo_g, hi_g, n_min, n_max, power, target, op, all_powers)
    if sum(int(r.count) for r in recs) < n_max:
        raise RuntimeError("prime count bound violated")
    # merge the ranks' records field by field (ranks are in value order); big integers are converted only where the merged
    # outcome needs them: the crossing rank's two sums and the last scanning rank's final sum
    valid = [r for r in recs if r.valid]
    if not valid:
        raise ValueError("no slice scanned anything")
    powers = []
    for j in range(np_):
        qs = [r.powers[j] for r in valid]
        cross = next((q for q in qs if q.cross_n), None)
        matched = [q for q in qs if q.match_count]
        powers.append(PowerOutcome(
            power=int(qs[0].power), match_count=sum(int(q.match_count) for q in matched),
            first_match_n=min((int(q.first_match_n) for q in matched), default=0),
            last_match_n=max((int(q.last_match_n) for q in matched), default=0),
            cross_n=int(cross.cross_n) if cross else 0, cross_prime=int(cross.cross_prime) if cross else 0,
            sum_at_cross=_lib.limbs_to_int(cross.sum_at_cross) if cross else 0,
            sum_before_cross=_lib.limbs_to_int(cross.sum_before_cross) if cross else 0,
            final_sum=_lib.limbs_to_int(qs[-1].final_sum)))
    stats["exchange"] = "peer memory (NVLink, device-side)"
    return SearchOutcome(n_min, n_max, op, target, int(valid[-1].last_prime), powers, stats)


_STAT_KEYS = ("ms_sieve", "ms_count_scan", "ms_compact", "ms_power_scan", "ms_bitmap_reduce", "ms_bitmap_compare",
              "ms_total_device", "primes", "sieve_lo", "sieve_hi", "launches")


def _pack_outcome(o: Optional[SearchOutcome], np_: int) -> List[int]:
    """fixed-size int64 record of a rank's outcome (big integers as 32-bit limbs; times in nanoseconds)"""
    L = _lib.PSS_LIMBS
    if o is None:
        return [0] * (2 + len(_STAT_KEYS) + np_ * (5 + 3 * L))
    rec = [1, o.last_prime]
    for k in _STAT_KEYS:
        v = o.stats.get(k, 0)
        rec.append(int(round(v * 1e6)) if k.startswith("ms_") else int(v))
    for p in o.powers:
        rec += [p.match_count, p.first_match_n, p.last_match_n, p.cross_n, p.cross_prime]
        for big in (p.sum_at_cross, p.sum_before_cross, p.final_sum):
            rec += [int(x) for x in _lib.int_to_limbs(big, L)]
    return rec


def _unpack_outcome(r: Sequence[int], np_: int, power: int, all_powers: bool, n_min, n_max, op, target) -> Optional[SearchOutcome]:
    L = _lib.PSS_LIMBS
    if not r[0]:
        return None
    stats = {}
    for i, k in enumerate(_STAT_KEYS):
        v = r[2 + i]
        stats[k] = v / 1e6 if k.startswith("ms_") else v
    pos = 2 + len(_STAT_KEYS)
    powers = []
    for j in range(np_):
        mc, fm, lm, cn, cp = r[pos:pos + 5]
        pos += 5
        bigs = []
        for _ in range(3):
            bigs.append(_lib.limbs_to_int(r[pos:pos + L]))
            pos += L
        powers.append(PowerOutcome(power=(j + 1) if (all_powers and power > 1) else power, match_count=mc, first_match_n=fm,
                                   last_match_n=lm, cross_n=cn, cross_prime=cp, sum_at_cross=bigs[0],
                                   sum_before_cross=bigs[1], final_sum=bigs[2]))
    return SearchOutcome(n_min, n_max, op, target, r[1], powers, stats)
