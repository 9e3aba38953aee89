"""
sieve.py — host-side mirror of the reference's sieve facade (utils/sieve.py:440-693) for the new
`b200` variant.  Same names, argument meaning and error behaviour; every prime comes from the sm_100a
segmented wheel-30 sieve (csrc/pss_sieve.cuh) — or, for `individual`, the per-candidate trial-division kernel — through
the C ABI.  No CPU fallback.

Reference semantics kept:
  generate_primes(limit)        exclusive bound, int64 array, limit < 2 -> empty      (utils/sieve.py:444,473-474)
  generate_primes_range(a, b)   [a, b)                                               (:512-538)
  generate_n_primes(n)          exactly n primes, n <= 0 -> empty                    (:541-585)
  nth_prime(n)                  1-indexed, ValueError("n must be positive (1-indexed)") (:588-616)
  prime_count(limit)            inclusive                                            (:619-641)
  configure / get_config / VALID_ALGORITHMS / get_available_algorithms / get_algorithm_info (:56-105,644-693)
Unlike the reference the configured algorithm really is the one that runs.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import _lib

ALGORITHM_AUTO = "auto"
ALGORITHM_B200 = "b200"
ALGORITHM_PRIMESIEVE = "primesieve"
ALGORITHM_BASIC = "basic"
ALGORITHM_SEGMENTED = "segmented"
ALGORITHM_INDIVIDUAL = "individual"
# What each of the reference's names runs HERE (every variant computes on the device; there is no CPU path):
#   auto, b200, segmented, basic -> the segmented bit-packed wheel-30 sieve (the only sieve in this build is segmented;
#                                   "basic" — one segment spanning the range — is served by the same kernel)
#   individual                   -> per-candidate trial division on the device (k_individual), the reference's
#                                   _individual_sieve (utils/sieve.py:330-387): same results, O(n sqrt(n) / log n)
#   primesieve                   -> not installed (third-party CPU library): ValueError like utils/sieve.py:492-498
VALID_ALGORITHMS = [ALGORITHM_AUTO, ALGORITHM_B200, ALGORITHM_PRIMESIEVE, ALGORITHM_BASIC, ALGORITHM_SEGMENTED, ALGORITHM_INDIVIDUAL]
_DEVICE_VARIANT = {ALGORITHM_AUTO: False, ALGORITHM_B200: False, ALGORITHM_BASIC: False, ALGORITHM_SEGMENTED: False,
                   ALGORITHM_INDIVIDUAL: True}

_config = {"algorithm": ALGORITHM_AUTO, "max_memory_mb": None, "prefer": None}


def configure(algorithm: str = None, max_memory_mb: int = None, prefer: str = None) -> None:
    """Same contract as utils/sieve.py:81-100: only the arguments that are given change; an unknown algorithm name raises
    ValueError("Invalid algorithm: ...") and changes nothing."""
    if algorithm is not None and algorithm not in VALID_ALGORITHMS:
        raise ValueError(f"Invalid algorithm: {algorithm}")
    for key, value in (("algorithm", algorithm), ("max_memory_mb", max_memory_mb), ("prefer", prefer)):
        if value is not None:
            _config[key] = value


def get_config() -> dict:
    return dict(_config)


def _resolve(algorithm: Optional[str]) -> bool:
    """-> True when the 'individual' device variant is selected, False for the segmented sieve"""
    if algorithm is None or algorithm == ALGORITHM_AUTO:
        algorithm = _config.get("algorithm", ALGORITHM_AUTO)
    if algorithm not in VALID_ALGORITHMS:
        raise ValueError(f"Unknown algorithm: {algorithm}. Valid options: {', '.join(VALID_ALGORITHMS)}")
    if algorithm == ALGORITHM_PRIMESIEVE:
        raise ValueError("primesieve requested but not installed (pss_b200 computes on the GPU: use 'b200', 'segmented', "
                         "'individual' or 'auto')")
    return _DEVICE_VARIANT[algorithm]


def _handle(algorithm: Optional[str]) -> "_lib.Handle":
    individual = _resolve(algorithm)
    h = _lib.default_handle()
    if getattr(h, "_individual", False) != individual:
        h.set_sieve_variant(individual)
        h._individual = individual
    return h


def _as_int(x, what: str) -> int:
    if isinstance(x, (bool,)):
        return int(x)
    if isinstance(x, (int, np.integer)):
        return int(x)
    if isinstance(x, float) and x == int(x):
        return int(x)
    raise TypeError(f"{what} must be an integer, got {type(x).__name__}")


def generate_primes(limit: int, algorithm: str = None, max_memory_mb: Optional[int] = None,
                    prefer: Optional[str] = None) -> np.ndarray:
    """All primes < limit (exclusive) as an int64 array."""
    limit = _as_int(limit, "limit")
    if limit < 2:
        return np.array([], dtype=np.int64)
    return _handle(algorithm).generate_primes(0, limit)


def generate_primes_range(start: int, stop: int, algorithm: str = ALGORITHM_AUTO,
                          max_memory_mb: Optional[int] = None) -> np.ndarray:
    """All primes in [start, stop).  The device sieves exactly that window (no [0, stop) pre-sieve)."""
    start, stop = _as_int(start, "start"), _as_int(stop, "stop")
    _resolve(algorithm)
    if stop < 2 or start >= stop:
        return np.array([], dtype=np.int64)
    return _handle(algorithm).generate_primes(max(start, 0), stop)


def generate_n_primes(n: int, algorithm: str = ALGORITHM_AUTO, max_memory_mb: Optional[int] = None) -> np.ndarray:
    n = _as_int(n, "n")
    if n <= 0:
        return np.array([], dtype=np.int64)
    return _handle(algorithm).generate_n_primes(n)


def nth_prime(n: int, algorithm: str = ALGORITHM_AUTO) -> int:
    n = _as_int(n, "n")
    if n <= 0:
        raise ValueError("n must be positive (1-indexed)")
    return _handle(algorithm).nth_prime(n)


def prime_count(limit: int, algorithm: str = ALGORITHM_AUTO) -> int:
    """pi(limit), inclusive."""
    limit = _as_int(limit, "limit")
    _resolve(algorithm)
    if limit < 2:
        return 0
    return _handle(algorithm).count_primes(0, limit + 1)


def get_available_algorithms() -> List[str]:
    return [ALGORITHM_AUTO, ALGORITHM_B200, ALGORITHM_BASIC, ALGORITHM_SEGMENTED, ALGORITHM_INDIVIDUAL]


def get_algorithm_info() -> dict:
    """Same shape as utils/sieve.py:657-693 so `--list algorithms` (utils/list_commands.py:60-97) can print it."""
    seg = {"available": True, "complexity_time": "O(n log log n)",
           "complexity_space": "O(sqrt(n)) base primes + n/30 bytes bitmap in HBM"}
    return {
        ALGORITHM_AUTO: {"description": "Auto-select best available algorithm", "available": True},
        ALGORITHM_B200: dict(seg, description="B200 (sm_100a) segmented bit-packed wheel-30 sieve, shared-memory tiles"),
        ALGORITHM_PRIMESIEVE: {"description": "C++ primesieve library (third-party CPU code, not part of pss_b200)", "available": False},
        ALGORITHM_BASIC: dict(seg, description="served by the B200 segmented sieve (one kernel for every range)"),
        ALGORITHM_SEGMENTED: dict(seg, description="B200 segmented sieve (same kernel as 'b200')"),
        ALGORITHM_INDIVIDUAL: {"description": "Individual testing on the device (trial division per candidate)", "available": True,
                               "complexity_time": "O(n * sqrt(n) / log(n))", "complexity_space": "n/30 bytes bitmap in HBM"},
    }
