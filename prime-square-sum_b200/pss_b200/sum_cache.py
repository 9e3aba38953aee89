"""
sum_cache.py — IncrementalSumCache-compatible running sums whose state is advanced on the B200.

The reference's cache (utils/sum_cache.py:82-321) keeps (prime_count, last_prime, {k: S_k}) and adds one
prime at a time in Python (~1.8 M primes/s).  That state is exactly the carry a device scan slice needs /
produces, so `advance_to(n)` moves it forward by whole slices on the GPU.  The `.npz`
checkpoint layout is the reference's (:169-261): key `metadata` = JSON of the CacheMetadata fields,
`sum_p{k}` = 0-d object array holding a Python int, plus the two diagnostic windows — the reference's
`IncrementalSumCache.load_checkpoint` reads these files and vice versa.
"""
from __future__ import annotations

import json
import time
from dataclasses import asdict, dataclass, field
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np

from . import _lib


@dataclass
class CacheMetadata:
    version: str = "1.0"
    powers: List[int] = field(default_factory=lambda: [2, 3, 4])
    prime_count: int = 0
    last_prime: Optional[int] = None
    created_timestamp: float = 0.0
    updated_timestamp: float = 0.0
    oeis_sequence_ids: Optional[Dict[int, str]] = None

    def __post_init__(self):
        if self.created_timestamp == 0.0:
            self.created_timestamp = time.time()
        if self.updated_timestamp == 0.0:
            self.updated_timestamp = self.created_timestamp


class IncrementalSumCache:
    """Same public surface as the reference class; `advance_to` is the device-accelerated addition."""

    def __init__(self, powers: List[int] = None):
        if powers is None:
            powers = [2, 3, 4]
        self.metadata = CacheMetadata(powers=list(powers))
        self.sums: Dict[int, int] = {p: 0 for p in powers}
        self.initial_values: list = []
        self.recent_values: list = []
        self._checkpoint_counter = 0

    # ---- reference-compatible per-prime path (host bookkeeping only)
    def add_prime(self, prime: int) -> None:
        prime = int(prime)   # the reference wraps silently when fed np.int64; exact here
        for power in self.metadata.powers:
            value = prime ** power
            self.sums[power] += value
            if self.metadata.prime_count < 2:
                self.initial_values.append((prime, value, self.sums[power], self.metadata.prime_count))
            self.recent_values.append((prime, value, self.sums[power], self.metadata.prime_count))
            if len(self.recent_values) > 3:
                self.recent_values.pop(0)
        self.metadata.prime_count += 1
        self.metadata.last_prime = prime
        self._checkpoint_counter += 1

    # ---- device path
    def advance_to(self, n: int, handle: Optional["_lib.Handle"] = None) -> None:
        """Advance the state to the first n primes on the GPU (no-op if already there)."""
        have = self.metadata.prime_count
        if n <= have:
            return
        h = handle or _lib.default_handle()
        lo = (self.metadata.last_prime + 1) if have else 0
        hi = h.nth_prime_upper(n) + 1
        cnt, _ = h.slice_sieve(lo, hi)
        need = n - have
        if cnt < need:
            raise RuntimeError("prime count bound violated")
        for power in self.metadata.powers:
            if power == 0:
                self.sums[power] += need
            else:
                self.sums[power] += h.slice_reduce(0, need, power)[0]
        tail = h.slice_primes(max(0, need - 3), min(3, need))
        self.metadata.prime_count = n
        self.metadata.last_prime = int(tail[-1])
        self._checkpoint_counter += need

    def should_checkpoint(self) -> bool:
        count = self.metadata.prime_count
        interval = 1_000 if count < 10_000 else 100 if count < 100_000 else 10 if count < 1_000_000 else 1
        return self._checkpoint_counter >= interval

    def save_checkpoint(self, filepath) -> None:
        filepath = Path(filepath)
        filepath.parent.mkdir(parents=True, exist_ok=True)
        self.metadata.updated_timestamp = time.time()
        md = asdict(self.metadata)
        md["last_prime"] = None if md["last_prime"] is None else int(md["last_prime"])
        data = {
            "metadata": json.dumps(md),
            "initial_values": np.array(self.initial_values, dtype=object),
            "recent_values": np.array(self.recent_values, dtype=object),
        }
        for power, s in self.sums.items():
            data[f"sum_p{power}"] = np.array(s, dtype=object)
        np.savez_compressed(filepath, **data)
        self._checkpoint_counter = 0

    @classmethod
    def load_checkpoint(cls, filepath) -> "IncrementalSumCache":
        filepath = Path(filepath)
        if not filepath.exists():
            raise FileNotFoundError(f"Checkpoint not found: {filepath}")
        try:
            data = np.load(filepath, allow_pickle=True)
        except Exception as e:  # noqa: BLE001
            raise ValueError(f"Failed to load checkpoint: {e}")
        try:
            metadata = CacheMetadata(**json.loads(str(data["metadata"])))
        except Exception as e:  # noqa: BLE001
            raise ValueError(f"Invalid metadata in checkpoint: {e}")
        cache = cls(powers=metadata.powers)
        cache.metadata = metadata
        for power in metadata.powers:
            key = f"sum_p{power}"
            if key not in data:
                raise ValueError(f"Missing sum for power {power} in checkpoint")
            cache.sums[power] = int(data[key].item())
        try:
            cache.initial_values = data["initial_values"].tolist()
            cache.recent_values = data["recent_values"].tolist()
        except Exception:  # noqa: BLE001
            cache.initial_values, cache.recent_values = [], []
        return cache

    def get_sum(self, power: int) -> int:
        if power not in self.sums:
            raise ValueError(f"Power {power} not tracked in cache. Available: {list(self.sums.keys())}")
        return self.sums[power]

    def get_prime_count(self) -> int:
        return self.metadata.prime_count

    def get_last_prime(self) -> int:
        return self.metadata.last_prime
