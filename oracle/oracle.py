"""
oracle.py — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker or the timed CPU baseline.  The product package
(prime-square-sum_b200/pss_b200) never imports it.

Two layers, both restating the reference's semantics (nothing copied; licence CC BY-NC-ND):

  * pure Python / numpy (small cases; the form closest to the reference)
      primes_below(limit)        utils/sieve.py:224-254 _basic_sieve            (exclusive bound)
      first_n_primes(n)          utils/sieve.py:541-585 generate_n_primes       (n <= 0 -> empty)
      nth_prime(n)               utils/sieve.py:588-616 (1-indexed, ValueError for n <= 0)
      prime_count(limit)         utils/sieve.py:619-641 (inclusive)
      power_sum(primes, k)       utils/gpu.py:244-247 sum(int(p) ** k)
      primesum(n, k)             utils/sequences.py:30-75
      trisum(b), tri(n), qtri(x) utils/number_theory.py:250-479 (targets)
      find_first(...) / find_all utils/grammar.py:937-999 does_exist / for_any over n (and p)
  * ctypes wrapper of oracle.c (liboracle.so) for sizes where Python loops are too slow.

Parity is PINNED by oracle/gen_golden.py (runs the reference itself, imported read-only from
/root/reference, and commits tests/golden/*.json) and tests/test_oracle.py.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORC_LIMBS = 16


# ---------------------------------------------------------------------------- pure Python layer
def primes_below(limit: int) -> np.ndarray:
    """All primes < limit as int64 (utils/sieve.py:224-254; limit < 2 -> empty, :240-241)."""
    if limit < 2:
        return np.array([], dtype=np.int64)
    flags = np.ones(limit, dtype=bool)
    flags[:2] = False
    for i in range(2, math.isqrt(limit - 1) + 1):
        if flags[i]:
            flags[i * i::i] = False
    return np.nonzero(flags)[0].astype(np.int64)


def nth_prime_bound(n: int) -> int:
    """Value bound that certainly contains the first n primes (Rosser; reference uses 1.3x PNT, :574-578)."""
    if n < 6:
        return 15
    return int(n * (math.log(n) + math.log(math.log(n)))) + 3


def first_n_primes(n: int) -> np.ndarray:
    if n <= 0:
        return np.array([], dtype=np.int64)
    return primes_below(nth_prime_bound(n) + 1)[:n]


def nth_prime(n: int) -> int:
    if n <= 0:
        raise ValueError("n must be positive (1-indexed)")
    return int(first_n_primes(n)[n - 1])


def prime_count(limit: int) -> int:
    return len(primes_below(limit + 1))


def power_sum(primes: Iterable[int], power: int = 2) -> int:
    return sum(int(p) ** power for p in primes)


def primesum(n: int, power: int = 2) -> int:
    if n < 0:
        raise ValueError(f"primesum() requires non-negative n, got {n}")
    if n == 0:
        return 0
    return power_sum(first_n_primes(n), power)


def tri(n: int) -> int:
    return n * (n + 1) // 2


def qtri(x: int) -> Optional[int]:
    """Inverse triangular number: m with tri(m) == x, else None (utils/number_theory.py:292-354)."""
    if x < 0:
        return None
    d = 1 + 8 * x
    s = math.isqrt(d)
    if s * s != d or (s - 1) % 2:
        return None
    return (s - 1) // 2


def trisum(b: int) -> int:
    """Row-sum of the triangular digit arrangement in base b, restated from
    utils/number_theory.py:397-479: r = largest r with tri(r) <= b rows; the digits 0..b-1 are dealt out
    to rows of length r, r-1, ..., 1 (longest row first: 0123 / 456 / 78 / 9 for b = 10), each row is
    read as a base-b number, and the rows are added.  trisum(10) = 666; trisum(666) is the 98-digit target."""
    if b < 1:
        raise ValueError(f"trisum() requires b >= 1, got {b}")
    if b == 1:
        return 0
    rows = 1
    while tri(rows + 1) <= b:
        rows += 1
    total, digit = 0, 0
    for length in range(rows, 0, -1):
        value = 0
        for d in range(digit, min(digit + length, b)):
            value = value * b + d
        digit += length
        total += value
        if digit >= b:
            break
    return total


_CMP = {
    "==": lambda a, b: a == b, "!=": lambda a, b: a != b, "<": lambda a, b: a < b,
    "<=": lambda a, b: a <= b, ">": lambda a, b: a > b, ">=": lambda a, b: a >= b,
}


def find_all(n_min: int, n_max: int, powers: Sequence[int], target: int, op: str = "==") -> List[Tuple[int, int]]:
    """for_any primesum(n,p) <op> target, n outer / p inner (utils/grammar.py:939,977): list of (n, p)."""
    primes = first_n_primes(n_max)
    sums = {p: 0 for p in powers}
    out = []
    f = _CMP[op]
    for i, q in enumerate(primes, start=1):
        q = int(q)
        for p in powers:
            sums[p] += q ** p
        if i >= n_min:
            for p in powers:
                if f(sums[p], target):
                    out.append((i, p))
    return out


def find_first(n_min: int, n_max: int, powers: Sequence[int], target: int, op: str = "==") -> Optional[Tuple[int, int]]:
    """does_exist: first match in iteration order (utils/grammar.py:987-988)."""
    r = find_all(n_min, n_max, powers, target, op)
    return r[0] if r else None


# ---------------------------------------------------------------------------- C layer (oracle.c)
class ScanResult(ctypes.Structure):
    _fields_ = [
        ("match_count", ctypes.c_uint64), ("first_match_n", ctypes.c_uint64), ("last_match_n", ctypes.c_uint64),
        ("cross_n", ctypes.c_uint64), ("cross_prime", ctypes.c_uint64),
        ("sum_at_cross", ctypes.c_uint32 * ORC_LIMBS), ("sum_before_cross", ctypes.c_uint32 * ORC_LIMBS),
        ("final_sum", ctypes.c_uint32 * ORC_LIMBS),
    ]


_lib = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-Wall", "-fPIC", "-shared", "-o", so, src, "-lm"])
    return so


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        u64, u32 = ctypes.c_uint64, ctypes.c_uint32
        pu64, pu32 = ctypes.POINTER(u64), ctypes.POINTER(u32)
        L.orc_sieve_range.restype = u64
        L.orc_sieve_range.argtypes = [u64, u64, u64, pu64, u64]
        L.orc_power_sums.restype = None
        L.orc_power_sums.argtypes = [pu64, u64, pu32, u32, pu32, pu32]
        L.orc_scan_compare.restype = None
        L.orc_scan_compare.argtypes = [pu64, u64, u64, u32, pu32, pu32, u32, u64, u64, ctypes.POINTER(ScanResult)]
        L.orc_prefix_sums.restype = None
        L.orc_prefix_sums.argtypes = [pu64, u64, u32, pu32, pu32]
        L.orc_block_sieve_sum.restype = u64
        L.orc_block_sieve_sum.argtypes = [u64, u64, u64, pu32, u32, pu32, pu64]
        L.orc_fast_block_sieve_sum.restype = u64
        L.orc_fast_block_sieve_sum.argtypes = [u64, u64, pu32, u32, pu32, pu64]
        _lib = L
    return _lib


def to_limbs(x: int, n: int = ORC_LIMBS) -> np.ndarray:
    if x < 0 or x >> (32 * n):
        raise OverflowError("value does not fit the limb array")
    return np.frombuffer(x.to_bytes(4 * n, "little"), dtype=np.uint32).copy()


def from_limbs(a) -> int:
    return int.from_bytes(np.asarray(a, dtype=np.uint32).tobytes(), "little")


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def c_primes_range(lo: int, hi: int, segment_size: int = 0) -> np.ndarray:
    """primes in [lo, hi) via oracle.c (two calls: count, then fill)."""
    L = lib()
    n = L.orc_sieve_range(lo, hi, segment_size, None, 0)
    out = np.empty(n, dtype=np.uint64)
    if n:
        L.orc_sieve_range(lo, hi, segment_size, _p(out, ctypes.c_uint64), n)
    return out


def c_count_range(lo: int, hi: int, segment_size: int = 0) -> int:
    return int(lib().orc_sieve_range(lo, hi, segment_size, None, 0))


def c_first_n_primes(n: int) -> np.ndarray:
    if n <= 0:
        return np.empty(0, dtype=np.uint64)
    return c_primes_range(0, nth_prime_bound(n) + 1)[:n]


def c_power_sums(values: np.ndarray, powers: Sequence[int], carry: Optional[Sequence[int]] = None) -> List[int]:
    values = np.ascontiguousarray(values, dtype=np.uint64)
    pw = np.asarray(list(powers), dtype=np.uint32)
    out = np.zeros(len(pw) * ORC_LIMBS, dtype=np.uint32)
    cin = None
    if carry is not None:
        cin = np.concatenate([to_limbs(int(c)) for c in carry])
    lib().orc_power_sums(_p(values, ctypes.c_uint64), len(values), _p(pw, ctypes.c_uint32), len(pw),
                         _p(cin, ctypes.c_uint32) if cin is not None else None, _p(out, ctypes.c_uint32))
    return [from_limbs(out[j * ORC_LIMBS:(j + 1) * ORC_LIMBS]) for j in range(len(pw))]


_OPS = {"==": 0, "!=": 1, "<": 2, "<=": 3, ">": 4, ">=": 5}


def c_scan_compare(values: np.ndarray, n_first: int, power: int, carry: int, target: int, op: str,
                   n_min: int, n_max: int) -> dict:
    values = np.ascontiguousarray(values, dtype=np.uint64)
    res = ScanResult()
    cin, tg = to_limbs(carry), to_limbs(target)
    lib().orc_scan_compare(_p(values, ctypes.c_uint64), len(values), n_first, power, _p(cin, ctypes.c_uint32),
                           _p(tg, ctypes.c_uint32), _OPS[op], n_min, n_max, ctypes.byref(res))
    return {
        "match_count": int(res.match_count), "first_match_n": int(res.first_match_n),
        "last_match_n": int(res.last_match_n), "cross_n": int(res.cross_n), "cross_prime": int(res.cross_prime),
        "sum_at_cross": from_limbs(np.ctypeslib.as_array(res.sum_at_cross)),
        "sum_before_cross": from_limbs(np.ctypeslib.as_array(res.sum_before_cross)),
        "final_sum": from_limbs(np.ctypeslib.as_array(res.final_sum)),
    }


def c_prefix_sums(values: np.ndarray, power: int, carry: int = 0) -> List[int]:
    values = np.ascontiguousarray(values, dtype=np.uint64)
    out = np.zeros(len(values) * ORC_LIMBS, dtype=np.uint32)
    cin = to_limbs(carry)
    lib().orc_prefix_sums(_p(values, ctypes.c_uint64), len(values), power, _p(cin, ctypes.c_uint32), _p(out, ctypes.c_uint32))
    return [from_limbs(out[i * ORC_LIMBS:(i + 1) * ORC_LIMBS]) for i in range(len(values))]


def c_block_sieve_sum(lo: int, hi: int, powers: Sequence[int], segment_size: int = 0) -> Tuple[int, List[int], int]:
    """(count, [sum p^k for k in powers], last_prime) for the primes in [lo, hi) — one CPU-baseline work unit."""
    pw = np.asarray(list(powers), dtype=np.uint32)
    out = np.zeros(len(pw) * ORC_LIMBS, dtype=np.uint32)
    last = ctypes.c_uint64(0)
    n = lib().orc_block_sieve_sum(lo, hi, segment_size, _p(pw, ctypes.c_uint32), len(pw), _p(out, ctypes.c_uint32),
                                  ctypes.byref(last))
    return int(n), [from_limbs(out[j * ORC_LIMBS:(j + 1) * ORC_LIMBS]) for j in range(len(pw))], int(last.value)


def c_fast_block_sieve_sum(lo: int, hi: int, powers: Sequence[int]) -> Tuple[int, List[int], int]:
    """Same result as c_block_sieve_sum from the primesieve stand-in (bit-packed wheel-30 segmented sieve, 256-bit sums):
    the stronger CPU number bench.py reports beside the port of the reference's own formulation."""
    pw = np.asarray(list(powers), dtype=np.uint32)
    out = np.zeros(len(pw) * ORC_LIMBS, dtype=np.uint32)
    last = ctypes.c_uint64(0)
    n = lib().orc_fast_block_sieve_sum(lo, hi, _p(pw, ctypes.c_uint32), len(pw), _p(out, ctypes.c_uint32), ctypes.byref(last))
    return int(n), [from_limbs(out[j * ORC_LIMBS:(j + 1) * ORC_LIMBS]) for j in range(len(pw))], int(last.value)
