"""
sieve.py — host-side mirror of the reference's sieve facade (utils/sieve.py:440-693) for the new
`b200` variant.  Same names, argument meaning and error behaviour; every prime comes from the sm_100a
segmented wheel-30 sieve (csrc/pss_sieve.cuh) through the C ABI.  No CPU fallback.

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
# the reference's CPU variants: recognised names, not provided by this package
_REFERENCE_CPU_VARIANTS = ["primesieve", "basic", "segmented", "individual"]
VALID_ALGORITHMS = [ALGORITHM_AUTO, ALGORITHM_B200] + _REFERENCE_CPU_VARIANTS

_config = {"algorithm": ALGORITHM_AUTO, "max_memory_mb": None, "prefer": None}


def configure(algorithm: str = None, max_memory_mb: int = None, prefer: str = None) -> None:
    """Same contract as utils/sieve.py:81-100 (ValueError("Invalid algorithm: ...") for unknown names)."""
    if algorithm is not None:
        if algorithm not in VALID_ALGORITHMS:
            raise ValueError(f"Invalid algorithm: {algorithm}")
        _config["algorithm"] = algorithm
    if max_memory_mb is not None:
        _config["max_memory_mb"] = max_memory_mb
    if prefer is not None:
        _config["prefer"] = prefer


def get_config() -> dict:
    return _config.copy()


def _resolve(algorithm: Optional[str]) -> str:
    if algorithm is None:
        algorithm = _config.get("algorithm", ALGORITHM_AUTO)
    if algorithm not in VALID_ALGORITHMS:
        raise ValueError(f"Unknown algorithm: {algorithm}. Valid options: {', '.join(VALID_ALGORITHMS)}")
    if algorithm in _REFERENCE_CPU_VARIANTS:
        # mirrors "primesieve requested but not installed" (utils/sieve.py:492-498): a recognised variant
        # this build does not carry.  pss_b200 has no CPU path by design.
        raise ValueError(f"{algorithm} requested but not provided by pss_b200 (GPU-only build; use 'b200' or 'auto')")
    return ALGORITHM_B200


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
    _resolve(algorithm)
    return _lib.default_handle().generate_primes(0, limit)


def generate_primes_range(start: int, stop: int, algorithm: str = ALGORITHM_AUTO,
                          max_memory_mb: Optional[int] = None) -> np.ndarray:
    """All primes in [start, stop).  The device sieves exactly that window (no [0, stop) pre-sieve)."""
    start, stop = _as_int(start, "start"), _as_int(stop, "stop")
    _resolve(algorithm)
    if stop < 2 or start >= stop:
        return np.array([], dtype=np.int64)
    return _lib.default_handle().generate_primes(max(start, 0), stop)


def generate_n_primes(n: int, algorithm: str = ALGORITHM_AUTO, max_memory_mb: Optional[int] = None) -> np.ndarray:
    n = _as_int(n, "n")
    if n <= 0:
        return np.array([], dtype=np.int64)
    _resolve(algorithm)
    return _lib.default_handle().generate_n_primes(n)


def nth_prime(n: int, algorithm: str = ALGORITHM_AUTO) -> int:
    n = _as_int(n, "n")
    if n <= 0:
        raise ValueError("n must be positive (1-indexed)")
    _resolve(algorithm)
    return _lib.default_handle().nth_prime(n)


def prime_count(limit: int, algorithm: str = ALGORITHM_AUTO) -> int:
    """pi(limit), inclusive."""
    limit = _as_int(limit, "limit")
    _resolve(algorithm)
    if limit < 2:
        return 0
    return _lib.default_handle().count_primes(0, limit + 1)


def get_available_algorithms() -> List[str]:
    return [ALGORITHM_AUTO, ALGORITHM_B200]


def get_algorithm_info() -> dict:
    """Same shape as utils/sieve.py:657-693 so `--list algorithms` (utils/list_commands.py:60-97) can print it."""
    info = {
        ALGORITHM_AUTO: {"description": "Auto-select best available algorithm", "available": True},
        ALGORITHM_B200: {
            "description": "B200 (sm_100a) segmented bit-packed wheel-30 sieve, shared-memory tiles",
            "available": True,
            "complexity_time": "O(n log log n)",
            "complexity_space": "O(sqrt(n)) base primes + n/30 bytes bitmap in HBM",
        },
    }
    for name in _REFERENCE_CPU_VARIANTS:
        info[name] = {"description": f"reference CPU variant '{name}' (not part of pss_b200)", "available": False}
    return info
