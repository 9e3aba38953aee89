"""
pss_b200 — B200-native (sm_100a) implementation of Prime-Square-Sum's one data-parallel hot path:
prime enumeration -> exact power prefix sums primesum(n, p) -> compare every partial sum with a target.

Host modules mirror the reference's operator / plugin API for this path (same names, arguments, errors):
  sieve       utils/sieve.py     generate_primes, generate_primes_range, generate_n_primes, nth_prime,
                                 prime_count, configure, get_algorithm_info      (variant "b200")
  sequences   utils/sequences.py primesum(n, power=2)
  gpu         utils/gpu.py       init_gpu, power_sum, gpu_power_sum, gpu_power_values
  sum_cache   utils/sum_cache.py IncrementalSumCache (+ device advance_to), same .npz checkpoints
  search      utils/grammar.py   fused does_exist / for_any over primesum(n,p) <op> constant, 1..8 GPUs
  prime_io    utils/prime_io.py  .npy export of device-generated primes, power sums over .npy files (memmap windows)
  fastpath    recogniser: hot-path expression strings -> one fused device pass, reference stdout / rc
  plugin      --functions file for the reference CLI
Everything computes in libpss_b200.so (C ABI in include/pss_b200.h); there is no CPU fallback.
"""
from . import _lib, fastpath, gpu, prime_io, search, sequences, sieve, sum_cache  # noqa: F401
from ._lib import Handle, default_handle  # noqa: F401
from .sequences import primesum  # noqa: F401
from .sieve import (generate_n_primes, generate_primes, generate_primes_range, nth_prime,  # noqa: F401
                    prime_count)
from .gpu import power_sum  # noqa: F401

__version__ = "0.1.0"
