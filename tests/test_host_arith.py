"""The kernels' shared __host__ __device__ arithmetic (pss_common.cuh), compiled for the host and replayed
serially against the oracle: wheel geometry, first offsets, edge masks, pattern bytes, limb multiply chain.
(tests/host_harness.cpp is test infrastructure, not a CPU fallback.)"""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("harness") / "libhost_harness.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "host_harness.cpp")])
    L = ctypes.CDLL(so)
    u64, u32 = ctypes.c_uint64, ctypes.c_uint32
    L.emul_sieve.restype = u64
    L.emul_sieve.argtypes = [u64, u64, u32, u32, ctypes.POINTER(u64), u64]
    L.emul_power_chain.argtypes = [u64, ctypes.POINTER(u32)]
    L.emul_limb_widths.restype = ctypes.c_int
    return L


def emul(L, lo, hi, tile, npat):
    n = L.emul_sieve(lo, hi, tile, npat, None, 0)
    out = np.empty(n, dtype=np.uint64)
    if n:
        L.emul_sieve(lo, hi, tile, npat, out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), n)
    return out


def test_tile_replay_matches_oracle(harness, orc):
    cases = [(0, 100), (0, 2), (0, 3), (2, 3), (3, 6), (0, 7), (0, 8), (5, 8), (7, 8), (0, 30), (0, 31), (0, 61), (0, 62),
             (0, 3722), (29, 32), (59, 62), (0, 100000), (480, 10000), (999000, 1002000), (10 ** 9, 10 ** 9 + 100000),
             (2 ** 32 - 30000, 2 ** 32 + 30000), (2401, 2900), (3 * 10 ** 10, 3 * 10 ** 10 + 50000), (2 ** 39, 2 ** 39 + 50000)]
    for lo, hi in cases:
        for tile, npat in [(8192, 4), (1024, 4), (16384, 2), (8192, 0), (2048, 1), (3072, 3)]:
            assert np.array_equal(emul(harness, lo, hi, tile, npat), orc.c_primes_range(lo, hi)), (lo, hi, tile, npat)
    rng = random.Random(7)
    for _ in range(25):
        lo = rng.randrange(0, 10 ** rng.randrange(1, 11))
        hi = lo + rng.randrange(1, 60000)
        tile, npat = rng.choice([1024, 2048, 8192, 98304]), rng.randrange(0, 5)
        assert np.array_equal(emul(harness, lo, hi, tile, npat), orc.c_primes_range(lo, hi)), (lo, hi, tile, npat)


def test_power_chain_limbs(harness, orc):
    buf = np.zeros(8 * 16, dtype=np.uint32)
    rng = random.Random(3)
    vals = [2, 3, 2 ** 32 - 5, 2 ** 32 + 15, 22801763489, 2 ** 40 - 87, 2 ** 35 + 1] + [rng.randrange(2, 2 ** 40) for _ in range(300)]
    for p in vals:
        harness.emul_power_chain(p, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
        for k in range(1, 9):
            assert orc.from_limbs(buf[(k - 1) * 16:k * 16]) == p ** k, (p, k)


def test_limb_widths_hold_the_documented_bounds(harness):
    """pow_limbs / sum_limbs (pss_common.cuh) cover p < 2^40, n < 2^36 (SURVEY §8a limb table)"""
    for k in range(1, 9):
        assert harness.emul_limb_widths(k, 0) * 32 >= 40 * k
        assert harness.emul_limb_widths(k, 1) * 32 >= 40 * k + 36
        assert harness.emul_limb_widths(k, 1) <= 12
    # the at-scale sums of SURVEY §8a fit with room to spare
    assert harness.emul_limb_widths(2, 1) == 4 and harness.emul_limb_widths(5, 1) == 8


def test_mulhi_modulo_identity():
    """tile_mod() in pss_sieve.cuh: r = b - mulhi(b, floor(2^32/p)) * p lies in [0, 2p) for every base prime"""
    from oracle import oracle as o
    ps = o.primes_below(2 ** 20)[3:].astype(np.uint64)
    inv = (np.uint64(1) << np.uint64(32)) // ps
    rng = np.random.default_rng(1)
    for trial in range(12):
        b = rng.integers(0, 2 ** 32, size=len(ps), dtype=np.uint64)
        if trial == 0:
            b[:] = 2 ** 32 - 1
        q = (b * inv) >> np.uint64(32)
        r = (b - q * ps) & np.uint64(0xFFFFFFFF)
        assert (r < 2 * ps).all()
        assert (np.where(r >= ps, r - ps, r) == b % ps).all()


def test_msb_offset_table():
    """msb_value_offset() in pss_bscan.cuh: PRMT byte table for bit-reversed words"""
    R = [1, 7, 11, 13, 17, 19, 23, 29]
    lo, hi = 0x6B6D7177, 0x5B616567
    tab = [(lo >> (8 * i)) & 0xFF for i in range(4)] + [(hi >> (8 * i)) & 0xFF for i in range(4)]
    assert all(t < 128 for t in tab)            # sign replication yields zero bytes
    for b in range(32):
        m = 31 - b
        assert tab[m & 7] - 30 * (m >> 3) == 30 * (b >> 3) + R[b & 7]
