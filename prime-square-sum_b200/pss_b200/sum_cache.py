"""
sum_cache.py — running prime power sums whose state moves forward on the B200, file-compatible with the
reference's `IncrementalSumCache` (utils/sum_cache.py:82-321).

What is shared with the reference is the ON-DISK SCHEMA only (utils/sum_cache.py:169-261):
    <file>.npz   "metadata"        JSON object {version, powers, prime_count, last_prime, created_timestamp,
                                   updated_timestamp, oeis_sequence_ids}
                 "sum_p<k>"        0-d object array holding the Python int S_k
                 "initial_values"  object array of (prime, prime**k, S_k, index) rows of the first two primes
                 "recent_values"   the last three such rows
so a checkpoint written here loads in the reference and the other way round (tests/test_host_logic.py).

The state (prime_count, last_prime, {k: S_k}) is exactly the carry a device scan slice takes and returns
(SURVEY §5), so moving it forward is ONE fused device pass: sieve (last_prime, bound) -> block sums of every
tracked power -> truncating scan to the n-th prime (`pss_slice_sieve_sums` + `pss_slice_scan_bitmap`).

    advance(cache, n)        works on ANY object with the reference's attribute surface — in particular on an
                             instance of the reference's own class (state in, state out)
    IncrementalSumCache      this package's class with the same public methods, plus .advance_to(n) and
                             .resume_search(...) (continue a does_exist search from the checkpoint, ROADMAP #26)
"""
from __future__ import annotations

import bisect
import json
import time
from pathlib import Path
from types import SimpleNamespace
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from . import _lib

SCHEMA_VERSION = "1.0"
_META_FIELDS = ("version", "powers", "prime_count", "last_prime", "created_timestamp", "updated_timestamp", "oeis_sequence_ids")
# checkpoint cadence of the reference (utils/sum_cache.py:143-167): primes processed so far -> primes between checkpoints
_CADENCE_LIMITS = (10_000, 100_000, 1_000_000)
_CADENCE_EVERY = (1_000, 100, 10, 1)
_HEAD_ROWS, _TAIL_ROWS = 2, 3           # diagnostics: rows of the first two primes, the last three rows


def _plain(x):
    """numpy scalars / containers -> plain Python (JSON cannot encode np.int64)"""
    if isinstance(x, np.generic):
        return x.item()
    if isinstance(x, dict):
        return {_plain(k): _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_plain(v) for v in x]
    return x


def _new_metadata(powers: Sequence[int], **kw) -> SimpleNamespace:
    now = time.time()
    md = SimpleNamespace(version=SCHEMA_VERSION, powers=[int(p) for p in powers], prime_count=0, last_prime=None,
                         created_timestamp=now, updated_timestamp=now, oeis_sequence_ids=None)
    for k, v in kw.items():
        if k not in _META_FIELDS:
            raise TypeError(f"unknown metadata field {k!r}")
        setattr(md, k, v)
    if not md.created_timestamp:
        md.created_timestamp = now
    if not md.updated_timestamp:
        md.updated_timestamp = md.created_timestamp
    return md


def _note_row(cache, prime: int, power: int, total: int, index: int) -> None:
    """diagnostic windows, filled the way the reference's add_prime fills them: one row per (prime, power)"""
    row = (prime, prime ** power, total, index)
    if index < _HEAD_ROWS:
        cache.initial_values.append(row)
    cache.recent_values.append(row)
    del cache.recent_values[:-_TAIL_ROWS]


def _host_add(cache, prime: int) -> None:
    md = cache.metadata
    prime = int(prime)                   # np.int64 would wrap at p^k >= 2^63; Python ints are exact
    for k in md.powers:
        cache.sums[k] += prime ** k
        _note_row(cache, prime, k, cache.sums[k], md.prime_count)
    md.prime_count += 1
    md.last_prime = prime
    cache._checkpoint_counter = getattr(cache, "_checkpoint_counter", 0) + 1


def _pass_groups(powers: Sequence[int]):
    """device passes that cover `powers`: [(power_max, all_powers, [k carried by the pass])].  Several powers <= 6 share one
    all-powers pass (which carries every k in 1..max); a single one, and every k in 7..8, gets a single-power pass."""
    pos = sorted({int(k) for k in powers if k > 0})
    low = [k for k in pos if k <= 6]
    groups = []
    if len(low) == 1:
        groups.append((low[0], False, low))
    elif low:
        groups.append((low[-1], True, list(range(1, low[-1] + 1))))
    groups += [(k, False, [k]) for k in pos if k > 6]
    return groups


def _device_sums_to(h: "_lib.Handle", have: int, last_prime: Optional[int], n: int, powers: Sequence[int],
                    carry: Dict[int, int]):
    """S_k(n) for every k in `powers` and p_n, continuing from (have, last_prime, carry): per group of powers one sieve of
    (last_prime, bound) + one block-sum pass + one truncating scan"""
    lo = (int(last_prime) + 1) if have else 0
    hi = h.nth_prime_upper(n) + 1
    out: Dict[int, int] = {}
    p_n = None
    for top, allp, ks in _pass_groups(powers):
        cnt, _ = h.slice_sieve_sums(lo, hi, top, allp)
        if cnt < n - have:
            raise RuntimeError("prime count bound violated")
        res = h.slice_scan_bitmap(have + 1, [carry[k] if have else 0 for k in ks], have + 1, n, top, 0, "<", allp)
        p_n = int(res.last_prime)
        for j, k in enumerate(ks):
            out[k] = _lib.limbs_to_int(res.powers[j].final_sum)
    if p_n is None:                      # only power 0 tracked: p_n is still needed for the state
        p_n = h.nth_prime(n)
    if 0 in powers:
        out[0] = carry.get(0, 0) + (n - have)
    return out, p_n


def advance(cache, n: int, handle: Optional["_lib.Handle"] = None):
    """Move `cache` (this package's class or the reference's own IncrementalSumCache instance) forward to the first n
    primes on the GPU.  No-op when it is already there.  Returns the cache."""
    md = cache.metadata
    n = int(n)
    have = int(md.prime_count)
    if n <= have:
        return cache
    for k in md.powers:
        if not 0 <= int(k) <= _lib.PSS_MAX_POWER:
            raise ValueError(f"power {k} outside 0..{_lib.PSS_MAX_POWER} cannot be advanced on the device")
    # the first two primes go through the host path so that the head window is filled exactly like add_prime fills it
    for p in (2, 3)[have:min(n, _HEAD_ROWS)]:
        _host_add(cache, p)
    have = int(md.prime_count)
    if n <= have:
        return cache
    h = handle or _lib.default_handle()
    want = sorted({int(k) for k in md.powers})
    carry = {int(k): int(cache.sums[k]) for k in md.powers}
    # an all-powers pass scans 1..max together: lower powers the cache does not track need their own carry S_k(have)
    lacking = sorted({k for _, _, ks in _pass_groups(want) for k in ks} - set(carry))
    if lacking:
        prev, _ = _device_sums_to(h, 0, None, have, lacking, dict.fromkeys(lacking, 0))
        carry.update(prev)
    sums, p_n = _device_sums_to(h, have, md.last_prime, n, want, carry)
    # diagnostics tail: the last rows the per-prime path would have left (needs the last few primes)
    tail_primes = [int(x) for x in h.generate_primes(max(0, p_n - 4096), p_n + 1)[-_TAIL_ROWS:]]
    rows = []
    running = {k: sums[k] for k in want}
    for back, p in enumerate(reversed(tail_primes)):
        idx = n - 1 - back
        if idx < have:
            break
        for k in reversed(md.powers):
            rows.append((p, p ** k, running[k], idx))
            running[k] -= p ** k
    for k in md.powers:
        cache.sums[k] = sums[int(k)]
    cache.recent_values = (list(cache.recent_values) + rows[::-1])[-_TAIL_ROWS:]
    cache._checkpoint_counter = getattr(cache, "_checkpoint_counter", 0) + (n - have)
    md.prime_count = n
    md.last_prime = p_n
    return cache


def write_checkpoint(cache, path) -> None:
    """serialise `cache` in the reference's .npz schema (see the module docstring)"""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    cache.metadata.updated_timestamp = time.time()
    meta = {f: _plain(getattr(cache.metadata, f)) for f in _META_FIELDS}
    arrays = {f"sum_p{_plain(k)}": np.array(int(v), dtype=object) for k, v in cache.sums.items()}
    arrays["metadata"] = json.dumps(meta)
    for name in ("initial_values", "recent_values"):
        arrays[name] = np.array([tuple(_plain(list(r))) for r in getattr(cache, name)], dtype=object)
    np.savez_compressed(path, **arrays)
    cache._checkpoint_counter = 0


def read_checkpoint(path, into) -> None:
    """fill `into` (a fresh cache object) from a .npz checkpoint written by this package or by the reference"""
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(f"Checkpoint not found: {path}")
    try:
        npz = np.load(path, allow_pickle=True)
    except Exception as e:  # noqa: BLE001
        raise ValueError(f"Failed to load checkpoint: {e}")
    try:
        fields = json.loads(str(npz["metadata"]))
        into.metadata = _new_metadata(fields.get("powers", []), **{k: v for k, v in fields.items() if k != "powers"})
    except Exception as e:  # noqa: BLE001
        raise ValueError(f"Invalid metadata in checkpoint: {e}")
    into.sums = {}
    for k in into.metadata.powers:
        if f"sum_p{k}" not in npz:
            raise ValueError(f"Missing sum for power {k} in checkpoint")
        into.sums[k] = int(npz[f"sum_p{k}"].item())
    for name in ("initial_values", "recent_values"):      # diagnostics only: tolerate their absence
        try:
            setattr(into, name, [tuple(r) for r in npz[name].tolist()])
        except Exception:  # noqa: BLE001
            setattr(into, name, [])


class IncrementalSumCache:
    """Same public surface as the reference class (add_prime, should_checkpoint, save_checkpoint, load_checkpoint,
    get_sum, get_prime_count, get_last_prime, report), state advanced on the device by `advance_to`."""

    def __init__(self, powers: Optional[Iterable[int]] = None):
        powers = [2, 3, 4] if powers is None else list(powers)
        self.metadata = _new_metadata(powers)
        self.sums: Dict[int, int] = dict.fromkeys(self.metadata.powers, 0)
        self.initial_values: List[tuple] = []
        self.recent_values: List[tuple] = []
        self._checkpoint_counter = 0

    # ---- per-prime host path (bookkeeping only; the reference's add_prime, exact for any prime size)
    def add_prime(self, prime: int) -> None:
        _host_add(self, prime)

    # ---- device path
    def advance_to(self, n: int, handle: Optional["_lib.Handle"] = None) -> None:
        """first n primes, one fused device pass from the current state"""
        advance(self, n, handle)

    def resume_search(self, target: int, n_max: int, power: int = 2, op: str = "==", handle: Optional["_lib.Handle"] = None):
        """`does_exist primesum(n, power) <op> target` for n in (prime_count, n_max], continuing from this state instead
        of re-accumulating from n = 1 (reference ROADMAP #26 `--start-n`); see search.search(resume=...)"""
        from . import search as _search
        return _search.search(target, n_max, power=power, op=op, resume=self, handle=handle)

    def should_checkpoint(self) -> bool:
        every = _CADENCE_EVERY[bisect.bisect_right(_CADENCE_LIMITS, self.metadata.prime_count)]
        return self._checkpoint_counter >= every

    def save_checkpoint(self, filepath) -> None:
        write_checkpoint(self, filepath)

    @classmethod
    def load_checkpoint(cls, filepath) -> "IncrementalSumCache":
        obj = cls(powers=[])
        read_checkpoint(filepath, obj)
        return obj

    def get_sum(self, power: int) -> int:
        try:
            return self.sums[power]
        except KeyError:
            raise ValueError(f"Power {power} not tracked in cache. Available: {list(self.sums.keys())}") from None

    def get_prime_count(self) -> int:
        return self.metadata.prime_count

    def get_last_prime(self) -> Optional[int]:
        return self.metadata.last_prime

    def report(self) -> None:
        """diagnostic print-out: counters, the head / tail windows and the current sums"""
        md = self.metadata
        lines = ["", "IncrementalSumCache Report", "=" * 70,
                 f"Primes processed: {md.prime_count}", f"Last prime: {md.last_prime}", f"Powers tracked: {md.powers}",
                 f"Created: {time.ctime(md.created_timestamp)}", f"Updated: {time.ctime(md.updated_timestamp)}", ""]
        for title, rows in (("Initial values (first 2 primes):", self.initial_values),
                            ("Recent values (last 3 primes):", self.recent_values)):
            if rows:
                lines.append(title)
                lines += [f"  [{i}] p={p:>6d}, p^n={v:>12d}, sum={s:>15d}, count={c}" for i, (p, v, s, c) in enumerate(rows)]
                lines.append("")
        lines.append("Current sums:")
        lines += [f"  Σp^{k} = {self.sums[k]}" for k in sorted(md.powers)]
        print("\n".join(lines) + "\n")


def migrate_old_checkpoint(old_checkpoint_path, new_cache_path) -> None:
    """old JSON checkpoint {primes_processed, last_prime, current_sum} (power 2 only) -> .npz in the schema above
    (reference: utils/sum_cache.py:324-368)"""
    src, dst = Path(old_checkpoint_path), Path(new_cache_path)
    if not src.exists():
        raise FileNotFoundError(f"Old checkpoint not found: {src}")
    old = json.loads(src.read_text())
    cache = IncrementalSumCache(powers=[2])
    try:
        cache.sums[2] = int(old["current_sum"])
    except (KeyError, ValueError) as e:
        raise ValueError(f"Invalid current_sum in old checkpoint: {e}")
    cache.metadata.prime_count = old.get("primes_processed", 0)
    cache.metadata.last_prime = old.get("last_prime", 0)
    cache.save_checkpoint(dst)
    print(f"Migration complete:\n  Source: {src}\n  Target: {dst}\n  Primes: {cache.metadata.prime_count}\n"
          "  Powers: Only power=2 recovered (power=3, 4, etc. not in old format)")
