"""
prime_io.py — the on-disk prime formats either side of the hot path, filled / consumed by the
B200 kernels.

The reference stores primes as an int64 `.npy` that `utils/prime_io.py:74-136` loads whole (`load_primes`) or
memory-mapped by index range (`load_primes_range`), and feeds such arrays to `utils/gpu.py:285 power_sum`.

  export_primes(path, n=... | limit=...)   device-generated primes streamed slice by slice into a `.npy` the
                                           reference's loaders read unchanged (same dtype, C order, v1/v2 header)
  power_sum_file(path, power)              exact sum of p^power over an externally supplied `.npy`, streamed
                                           through the device in index windows (memmap, like load_primes_range)
  load_primes / save_primes / load_primes_range   thin numpy mirrors of the reference functions for `.npy`

Only `.npy` is handled here: pickle / text are the reference's legacy formats and stay with its own module.
"""
from __future__ import annotations

from pathlib import Path
from typing import Iterator, Optional, Tuple, Union

import numpy as np

from . import _lib

PathLike = Union[str, Path]
DEFAULT_WINDOW = 1 << 26      # primes per device window (512 MiB of u64)


def plan_value_windows(limit: int, window_values: int, lo: int = 0) -> Iterator[Tuple[int, int]]:
    """[lo, limit) cut into half-open value windows; interior cuts are multiples of 480 (16 bitmap bytes), the
    alignment every device slice starts on"""
    if window_values < 480:
        raise ValueError("window must span at least 480 values")
    step = window_values // 480 * 480
    a = lo
    while a < limit:
        b = min(limit, (a // 480) * 480 + step)
        if b <= a:
            b = min(limit, a + step)
        yield a, b
        a = b


def export_primes(path: PathLike, n: Optional[int] = None, limit: Optional[int] = None, lo: int = 0,
                  window_values: int = 1 << 31, handle: Optional["_lib.Handle"] = None) -> int:
    """Write primes to `path` as int64 .npy: the first `n` primes, or all primes in [lo, limit).  Returns how many.
    The file is created with its final size and filled window by window (never more than one window in host RAM)."""
    if (n is None) == (limit is None):
        raise ValueError("give exactly one of n= and limit=")
    h = handle or _lib.default_handle()
    path = Path(path)
    if n is not None:
        if n < 0:
            raise ValueError("n must be non-negative")
        total = int(n)
        bound = h.nth_prime_upper(total) + 1 if total else 0
        lo = 0
    else:
        bound = max(int(limit), 0)
        total = h.count_primes(lo, bound) if bound > lo else 0
    out = np.lib.format.open_memmap(path, mode="w+", dtype=np.int64, shape=(total,))
    done = 0
    for a, b in plan_value_windows(bound, window_values, lo):
        if done >= total:
            break
        block = h.generate_primes(a, b)
        take = min(len(block), total - done)
        out[done:done + take] = block[:take]
        done += take
    out.flush()
    del out
    if done != total:
        raise RuntimeError(f"exported {done} primes, expected {total}")
    return total


def power_sum_file(path: PathLike, power: int = 2, window: int = DEFAULT_WINDOW, start_idx: int = 0,
                   end_idx: Optional[int] = None, handle: Optional["_lib.Handle"] = None) -> int:
    """Exact sum of p^power over primes[start_idx:end_idx] of an int64 .npy (memory-mapped: only the window in
    flight is resident), each window in one device pass"""
    h = handle or _lib.default_handle()
    arr = np.load(Path(path), mmap_mode="r")
    if arr.ndim != 1 or arr.dtype.kind not in "iu":
        raise TypeError("expected a one-dimensional integer .npy")
    end_idx = len(arr) if end_idx is None else min(int(end_idx), len(arr))
    total = 0
    for i in range(int(start_idx), end_idx, int(window)):
        chunk = np.ascontiguousarray(arr[i:min(i + window, end_idx)], dtype=np.uint64)
        total += h.power_sum(chunk, int(power))
    return total


# ---- thin mirrors of the reference's .npy functions (utils/prime_io.py:28-136) -------------------------------------
def load_primes(filepath: PathLike) -> np.ndarray:
    filepath = Path(filepath)
    if not filepath.exists():
        raise FileNotFoundError(f"Prime file not found: {filepath}")
    return np.asarray(np.load(filepath), dtype=np.int64)


def save_primes(primes, filepath: PathLike) -> None:
    np.save(Path(filepath), np.asarray(primes, dtype=np.int64))


def load_primes_range(filepath: PathLike, start_idx: int, end_idx: int) -> np.ndarray:
    primes = np.load(Path(filepath), mmap_mode="r")
    return np.array(primes[start_idx:end_idx], dtype=np.int64)
