"""
gpu.py — host-side mirror of the reference's array-level API (utils/gpu.py:27-334) on the B200 kernels.

power_sum(primes, power) is exact for every power 0..8 and every value < 2^40 in ONE device pass
(multi-limb IMAD carry chains), so the reference's int64 cut-off / chunking / hybrid CPU split
(utils/gpu.py:100-137,285-334) has no equivalent here: nothing overflows, nothing runs on the CPU.
Errors follow the reference: no usable GPU -> RuntimeError("GPU not available ...") (:160-161,207-208).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _lib

GPU_AVAILABLE = False
GPU_INFO: Optional[dict] = None


def init_gpu() -> bool:
    """Probe the device (reference: utils/gpu.py:27-73).  Never raises; records the reason on failure."""
    global GPU_AVAILABLE, GPU_INFO
    try:
        h = _lib.default_handle()
        GPU_INFO = {"name": h.device_info(), "backend": "libpss_b200 (sm_100a)"}
        GPU_AVAILABLE = True
    except Exception as e:  # noqa: BLE001 - mirror of the reference's broad probe
        GPU_INFO = {"error": str(e)}
        GPU_AVAILABLE = False
    return GPU_AVAILABLE


def _as_array(primes) -> np.ndarray:
    a = np.asarray(primes)
    if a.size and a.dtype.kind not in "iu":
        raise TypeError("power_sum() requires an integer array")
    if a.size and a.min() < 0:
        raise ValueError("power_sum() requires non-negative values")
    return np.ascontiguousarray(a, dtype=np.uint64)


def gpu_power_sum(primes, power: int = 2, chunk_size: int = None) -> int:
    """Sum of p^power, exact (reference: utils/gpu.py:190-241; chunk_size accepted and ignored)."""
    a = _as_array(primes)
    if a.size == 0:
        return 0
    return _lib.default_handle().power_sum(a, int(power))


def gpu_power_values(primes, power: int = 2, batch_size: int = 100000) -> int:
    """Reference: utils/gpu.py:140-187 (GPU powers + CPU big-int sum).  Here the sum is exact on device."""
    return gpu_power_sum(primes, power)


def power_sum(primes, power: int = 2, use_gpu: bool = True, **kwargs) -> int:
    """Main entry (reference: utils/gpu.py:285-334).  Empty -> 0 (:303).  use_gpu=False is refused:
    this build has no CPU path."""
    if not use_gpu:
        raise RuntimeError("pss_b200 has no CPU path; use the reference's cpu_power_sum for CPU sums")
    return gpu_power_sum(primes, power)
