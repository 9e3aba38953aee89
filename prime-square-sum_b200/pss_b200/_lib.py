"""
_lib.py — ctypes binding of libpss_b200.so (C ABI: include/pss_b200.h).

There is NO CPU fallback here: if the CUDA library is missing or no B200 is visible every compute call
raises (RuntimeError("GPU not available ...") like utils/gpu.py:160-161 of the reference).  The oracle
under /oracle is never imported from this package.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence

import numpy as np

PSS_LIMBS = 12
PSS_MAX_POWER = 8
ABI_VERSION = 3   # PSS_ABI_VERSION of include/pss_b200.h this binding was written against
PSS_MAX_PRIME_BITS = 40
PSS_OK, PSS_EINVAL, PSS_ENODEV, PSS_ENOMEM, PSS_EINTERNAL = 0, 1, 2, 3, 4

CMP_CODES = {"==": 0, "!=": 1, "<": 2, "<=": 3, ">": 4, ">=": 5}

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpss_b200.so")

u32, u64 = ctypes.c_uint32, ctypes.c_uint64
Limbs = u32 * PSS_LIMBS


class PowerResult(ctypes.Structure):
    _fields_ = [
        ("power", u32), ("reserved", u32),
        ("match_count", u64), ("first_match_n", u64), ("last_match_n", u64),
        ("cross_n", u64), ("cross_prime", u64),
        ("sum_at_cross", Limbs), ("sum_before_cross", Limbs), ("final_sum", Limbs),
    ]


class Stats(ctypes.Structure):
    _fields_ = [
        ("ms_sieve", ctypes.c_double), ("ms_count_scan", ctypes.c_double), ("ms_compact", ctypes.c_double),
        ("ms_power_scan", ctypes.c_double), ("ms_total_device", ctypes.c_double),
        ("primes", u64), ("sieve_lo", u64), ("sieve_hi", u64), ("launches", u64),
        ("ms_bitmap_reduce", ctypes.c_double), ("ms_bitmap_compare", ctypes.c_double), ("ms_span", ctypes.c_double),
        ("ms_peer_pre_publish", ctypes.c_double), ("ms_peer_wait_totals", ctypes.c_double),
        ("ms_peer_compare_region", ctypes.c_double), ("ms_peer_wait_results", ctypes.c_double),
    ]


PSS_IPC_HANDLE_BYTES = 64
PSS_MAX_PEERS = 16


class PeerRecord(ctypes.Structure):
    """pss_peer_record: the outcome of one rank of a peer-memory query"""
    _fields_ = [
        ("valid", u64), ("n_first", u64), ("scanned", u64), ("last_prime", u64), ("count", u64), ("error", u64),
        ("powers", PowerResult * PSS_MAX_POWER),
    ]


class SearchArgs(ctypes.Structure):
    _fields_ = [
        ("n_min", u64), ("n_max", u64), ("power_max", u32), ("all_powers", u32), ("cmp", u32),
        ("target_nlimbs", u32), ("target", ctypes.POINTER(u32)),
    ]


class SearchResult(ctypes.Structure):
    _fields_ = [
        ("n_powers", u32), ("reserved", u32), ("primes_scanned", u64), ("last_prime", u64),
        ("powers", PowerResult * PSS_MAX_POWER), ("stats", Stats),
    ]


EXPORTS = [
    "pss_abi_version", "pss_open", "pss_close", "pss_last_error", "pss_device_info", "pss_set_tuning",
    "pss_count_primes", "pss_generate_primes", "pss_generate_n_primes", "pss_nth_prime",
    "pss_power_sum", "pss_primesum", "pss_search",
    "pss_slice_sieve", "pss_slice_reduce", "pss_slice_scan", "pss_slice_prefix_sums", "pss_slice_primes",
    "pss_slice_device_buffer", "pss_nth_prime_upper", "pss_stats_reset", "pss_stats_get",
    "pss_set_search_path", "pss_slice_sieve_sums", "pss_slice_scan_bitmap", "pss_search_tri",
    "pss_set_compare_mode", "pss_peer_export", "pss_peer_connect", "pss_peer_search", "pss_search_targets",
    "pss_set_sieve_variant", "pss_set_kernel_events",
]

_cdll = None


def load_library() -> ctypes.CDLL:
    """dlopen the in-tree CUDA library and declare every prototype.  Loud failure if it was not built."""
    global _cdll
    if _cdll is not None:
        return _cdll
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"GPU not available: {LIB_PATH} is missing — build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    H = ctypes.c_void_p
    pu32, pu64 = ctypes.POINTER(u32), ctypes.POINTER(u64)
    L.pss_abi_version.restype = ctypes.c_int
    L.pss_open.argtypes = [ctypes.c_int, ctypes.POINTER(H)]
    L.pss_close.argtypes = [H]
    L.pss_last_error.argtypes = [H]
    L.pss_last_error.restype = ctypes.c_char_p
    L.pss_device_info.argtypes = [H, ctypes.c_char_p, ctypes.c_size_t]
    L.pss_set_tuning.argtypes = [H, u32, u32, u32]
    L.pss_count_primes.argtypes = [H, u64, u64, pu64]
    L.pss_generate_primes.argtypes = [H, u64, u64, pu64, u64, pu64]
    L.pss_generate_n_primes.argtypes = [H, u64, pu64]
    L.pss_nth_prime.argtypes = [H, u64, pu64]
    L.pss_power_sum.argtypes = [H, pu64, u64, u32, pu32]
    L.pss_primesum.argtypes = [H, u64, u32, pu32]
    L.pss_search.argtypes = [H, ctypes.POINTER(SearchArgs), ctypes.POINTER(SearchResult)]
    L.pss_slice_sieve.argtypes = [H, u64, u64, pu64, pu64]
    L.pss_slice_reduce.argtypes = [H, u64, u64, u32, u32, pu32]
    L.pss_slice_scan.argtypes = [H, u64, u64, u64, pu32, ctypes.POINTER(SearchArgs), ctypes.POINTER(SearchResult)]
    L.pss_slice_prefix_sums.argtypes = [H, u64, u64, u32, pu32, pu32]
    L.pss_slice_primes.argtypes = [H, u64, u64, pu64]
    L.pss_slice_device_buffer.argtypes = [H, ctypes.POINTER(ctypes.c_void_p), pu64]
    L.pss_search_tri.argtypes = [H, u64, u64, u32, u64, u64, pu64, u32, pu32, ctypes.POINTER(Stats)]
    L.pss_set_search_path.argtypes = [H, ctypes.c_int]
    L.pss_set_compare_mode.argtypes = [H, ctypes.c_int]
    L.pss_set_sieve_variant.argtypes = [H, ctypes.c_int]
    L.pss_set_kernel_events.argtypes = [H, ctypes.c_int]
    L.pss_search_targets.argtypes = [H, ctypes.POINTER(SearchArgs), u32, ctypes.POINTER(SearchResult)]
    L.pss_peer_export.argtypes = [H, ctypes.c_void_p]
    L.pss_peer_connect.argtypes = [H, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.pss_peer_search.argtypes = [H, u64, u64, ctypes.POINTER(SearchArgs), ctypes.POINTER(PeerRecord), ctypes.POINTER(Stats)]
    L.pss_slice_sieve_sums.argtypes = [H, u64, u64, u32, u32, pu64, pu32]
    L.pss_slice_scan_bitmap.argtypes = [H, u64, pu32, ctypes.POINTER(SearchArgs), ctypes.POINTER(SearchResult)]
    L.pss_stats_reset.argtypes = [H]
    L.pss_stats_get.argtypes = [H, ctypes.POINTER(Stats)]
    L.pss_nth_prime_upper.argtypes = [u64]
    L.pss_nth_prime_upper.restype = u64
    for name in EXPORTS:
        if name not in ("pss_last_error", "pss_nth_prime_upper"):
            getattr(L, name).restype = ctypes.c_int
    if L.pss_abi_version() != ABI_VERSION:
        raise RuntimeError("libpss_b200.so ABI version mismatch")
    _cdll = L
    return L


def stats_to_dict(s: "Stats") -> dict:
    return {"ms_sieve": s.ms_sieve, "ms_count_scan": s.ms_count_scan, "ms_compact": s.ms_compact,
            "ms_power_scan": s.ms_power_scan, "ms_bitmap_reduce": s.ms_bitmap_reduce,
            "ms_bitmap_compare": s.ms_bitmap_compare, "ms_total_device": s.ms_total_device, "ms_span": s.ms_span, "primes": int(s.primes),
            "sieve_lo": int(s.sieve_lo), "sieve_hi": int(s.sieve_hi), "launches": int(s.launches),
            "ms_peer_pre_publish": s.ms_peer_pre_publish, "ms_peer_wait_totals": s.ms_peer_wait_totals,
            "ms_peer_compare_region": s.ms_peer_compare_region, "ms_peer_wait_results": s.ms_peer_wait_results}


def int_to_limbs(x: int, n: Optional[int] = None) -> np.ndarray:
    """Python int -> little-endian uint32 limbs (at least one limb; exactly n if given)."""
    if x < 0:
        raise ValueError("negative values have no limb representation")
    need = max(1, (x.bit_length() + 31) // 32)
    if n is None:
        n = need
    elif need > n:
        raise OverflowError(f"value needs {need} limbs, only {n} available")
    return np.frombuffer(x.to_bytes(4 * n, "little"), dtype=np.uint32).copy()


def limbs_to_int(a) -> int:
    if isinstance(a, ctypes.Array):          # a limb field of a result struct: no numpy round trip on the query path
        return int.from_bytes(bytes(a), "little")
    return int.from_bytes(np.asarray(a, dtype=np.uint32).tobytes(), "little")


def _pu32(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(u32))


def _pu64(a: np.ndarray):
    return a.ctypes.data_as(ctypes.POINTER(u64))


class Handle:
    """One device context (one per process / rank).  Not thread-safe, like the reference."""

    def __init__(self, device: int = -1):
        self._L = load_library()
        self._h = ctypes.c_void_p()
        rc = self._L.pss_open(device, ctypes.byref(self._h))
        if rc != PSS_OK:
            msg = (self._L.pss_last_error(None) or b"").decode()
            self._h = ctypes.c_void_p()
            self._raise(rc, msg)

    # -- error mapping (ValueError / RuntimeError conventions of the reference)
    @staticmethod
    def _raise(rc: int, msg: str):
        if rc == PSS_EINVAL:
            raise ValueError(msg)
        if rc == PSS_ENOMEM:
            raise MemoryError(msg)
        raise RuntimeError(msg)

    def _check(self, rc: int):
        if rc != PSS_OK:
            self._raise(rc, (self._L.pss_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.pss_close(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_info(self) -> str:
        buf = ctypes.create_string_buffer(256)
        self._check(self._L.pss_device_info(self._h, buf, 256))
        return buf.value.decode()

    def set_tuning(self, tile_bytes: int = 0, sieve_threads: int = 0, n_patterns: int = 0):
        self._check(self._L.pss_set_tuning(self._h, tile_bytes, sieve_threads, n_patterns))

    # -- sieve
    def count_primes(self, lo: int, hi_excl: int) -> int:
        out = u64(0)
        self._check(self._L.pss_count_primes(self._h, max(lo, 0), max(hi_excl, 0), ctypes.byref(out)))
        return int(out.value)

    def generate_primes(self, lo: int, hi_excl: int) -> np.ndarray:
        lo, hi_excl = max(int(lo), 0), max(int(hi_excl), 0)
        n = self.count_primes(lo, hi_excl)
        out = np.empty(n, dtype=np.uint64)
        cnt = u64(0)
        if n:
            self._check(self._L.pss_generate_primes(self._h, lo, hi_excl, _pu64(out), n, ctypes.byref(cnt)))
        return out.view(np.int64)

    def generate_n_primes(self, n: int) -> np.ndarray:
        out = np.empty(max(n, 0), dtype=np.uint64)
        if n > 0:
            self._check(self._L.pss_generate_n_primes(self._h, n, _pu64(out)))
        return out.view(np.int64)

    def nth_prime(self, n: int) -> int:
        out = u64(0)
        self._check(self._L.pss_nth_prime(self._h, n, ctypes.byref(out)))
        return int(out.value)

    # -- power sums
    def power_sum(self, values: np.ndarray, power: int) -> int:
        vals = np.ascontiguousarray(values, dtype=np.uint64)
        out = np.zeros(PSS_LIMBS, dtype=np.uint32)
        self._check(self._L.pss_power_sum(self._h, _pu64(vals), len(vals), power, _pu32(out)))
        return limbs_to_int(out)

    def primesum(self, n: int, power: int) -> int:
        out = np.zeros(PSS_LIMBS, dtype=np.uint32)
        self._check(self._L.pss_primesum(self._h, n, power, _pu32(out)))
        return limbs_to_int(out)

    # -- fused search
    def _make_args(self, n_min, n_max, power_max, all_powers, op, target):
        tl = int_to_limbs(int(target))
        a = SearchArgs()
        a.n_min, a.n_max = int(n_min), int(n_max)
        a.power_max, a.all_powers = int(power_max), 1 if all_powers else 0
        a.cmp = CMP_CODES[op]
        a.target_nlimbs = len(tl)
        a.target = _pu32(tl)
        return a, tl   # keep tl alive

    def search(self, n_min: int, n_max: int, power_max: int, target: int, op: str = "==",
               all_powers: bool = False) -> SearchResult:
        a, keep = self._make_args(n_min, n_max, power_max, all_powers, op, target)
        res = SearchResult()
        self._check(self._L.pss_search(self._h, ctypes.byref(a), ctypes.byref(res)))
        del keep
        return res

    def search_targets(self, n_min: int, n_max: int, power_max: int, targets: Sequence[int], op: str = "==",
                       all_powers: bool = False) -> List["SearchResult"]:
        """several targets over the same n range: one sieve + one block-sum pass, one compare launch per target"""
        n = len(targets)
        if n == 0:
            return []
        arr = (SearchArgs * n)()
        keep = []
        for i, t in enumerate(targets):
            a, k = self._make_args(n_min, n_max, power_max, all_powers, op, int(t))
            arr[i] = a
            keep.append(k)
        res = (SearchResult * n)()
        self._check(self._L.pss_search_targets(self._h, arr, n, res))
        del keep
        return list(res)

    # -- slice API (multi-GPU decomposition)
    def slice_sieve(self, lo: int, hi_excl: int):
        cnt, last = u64(0), u64(0)
        self._check(self._L.pss_slice_sieve(self._h, lo, hi_excl, ctypes.byref(cnt), ctypes.byref(last)))
        return int(cnt.value), int(last.value)

    def slice_reduce(self, first: int, count: int, power_max: int, all_powers: bool = False) -> List[int]:
        np_ = power_max if (all_powers and power_max > 1) else 1
        out = np.zeros(np_ * PSS_LIMBS, dtype=np.uint32)
        self._check(self._L.pss_slice_reduce(self._h, first, count, power_max, 1 if all_powers else 0, _pu32(out)))
        return [limbs_to_int(out[j * PSS_LIMBS:(j + 1) * PSS_LIMBS]) for j in range(np_)]

    def slice_scan(self, first: int, count: int, n_first: int, carry: Sequence[int], n_min: int, n_max: int,
                   power_max: int, target: int, op: str = "==", all_powers: bool = False) -> SearchResult:
        a, keep = self._make_args(n_min, n_max, power_max, all_powers, op, target)
        np_ = power_max if (all_powers and power_max > 1) else 1
        cin = np.zeros(np_ * PSS_LIMBS, dtype=np.uint32)
        for j in range(np_):
            cin[j * PSS_LIMBS:(j + 1) * PSS_LIMBS] = int_to_limbs(int(carry[j]) if carry else 0, PSS_LIMBS)
        res = SearchResult()
        self._check(self._L.pss_slice_scan(self._h, first, count, n_first, _pu32(cin), ctypes.byref(a), ctypes.byref(res)))
        del keep
        return res

    def slice_prefix_sums(self, first: int, count: int, power: int, carry: int = 0) -> List[int]:
        out = np.zeros(max(count, 0) * PSS_LIMBS, dtype=np.uint32)
        cin = int_to_limbs(int(carry), PSS_LIMBS)
        self._check(self._L.pss_slice_prefix_sums(self._h, first, count, power, _pu32(cin), _pu32(out)))
        raw = out.tobytes()
        step = 4 * PSS_LIMBS
        return [int.from_bytes(raw[i * step:(i + 1) * step], "little") for i in range(count)]

    def slice_primes(self, first: int, count: int) -> np.ndarray:
        out = np.empty(max(count, 0), dtype=np.uint64)
        if count > 0:
            self._check(self._L.pss_slice_primes(self._h, first, count, _pu64(out)))
        return out.view(np.int64)

    def slice_device_buffer(self):
        ptr, cnt = ctypes.c_void_p(), u64(0)
        self._check(self._L.pss_slice_device_buffer(self._h, ctypes.byref(ptr), ctypes.byref(cnt)))
        return ptr.value, int(cnt.value)

    def stats_reset(self):
        self._check(self._L.pss_stats_reset(self._h))

    def stats(self) -> dict:
        s = Stats()
        self._check(self._L.pss_stats_get(self._h, ctypes.byref(s)))
        return stats_to_dict(s)

    def search_tri(self, n_min: int, n_max: int, power: int, m_min: int, m_max: int, cap: int = 4096):
        """[(n, m)] with primesum(n, power) == tri(m), sorted by n; plus the per-kernel stats"""
        out = np.zeros(2 * max(cap, 1), dtype=np.uint64)
        cnt = u32(0)
        st = Stats()
        self._check(self._L.pss_search_tri(self._h, n_min, n_max, power, m_min, m_max, _pu64(out), cap, ctypes.byref(cnt),
                                           ctypes.byref(st)))
        if cnt.value > cap:
            raise MemoryError(f"{cnt.value} matches exceed the buffer of {cap}")
        pairs = sorted((int(out[2 * i]), int(out[2 * i + 1])) for i in range(cnt.value))
        return pairs, stats_to_dict(st)

    def set_search_path(self, materialised: bool):
        self._check(self._L.pss_set_search_path(self._h, 1 if materialised else 0))

    def set_kernel_events(self, level: int):
        """per-kernel CUDA timing events: 0 none, 1 (default) the sieve only, 2 every kernel (see pss_b200.h)"""
        self._check(self._L.pss_set_kernel_events(self._h, int(level)))

    def set_sieve_variant(self, individual: bool):
        """False (default): the segmented wheel-30 sieve; True: per-candidate trial division (reference 'individual')"""
        self._check(self._L.pss_set_sieve_variant(self._h, 1 if individual else 0))

    def set_compare_mode(self, exhaustive: bool):
        """fused compare pass: False (default) = interval pruning, True = walk every prime / form every S_k(n)"""
        self._check(self._L.pss_set_compare_mode(self._h, 1 if exhaustive else 0))

    # -- peer-memory exchange (multi-GPU query without host / NCCL on the data path)
    def peer_export(self) -> bytes:
        """allocate this rank's mailbox and return its CUDA IPC handle (64 bytes) for the all-gather at set-up"""
        buf = ctypes.create_string_buffer(PSS_IPC_HANDLE_BYTES)
        self._check(self._L.pss_peer_export(self._h, buf))
        return buf.raw

    def peer_connect(self, rank: int, world: int, handles: Sequence[bytes]):
        """map every rank's mailbox (handles in rank order); call a barrier afterwards, before the first query"""
        if len(handles) != world or any(len(b) != PSS_IPC_HANDLE_BYTES for b in handles):
            raise ValueError("need one 64-byte IPC handle per rank")
        blob = ctypes.create_string_buffer(b"".join(handles), PSS_IPC_HANDLE_BYTES * world)
        self._check(self._L.pss_peer_connect(self._h, rank, world, blob))
        self.peer_world, self.peer_rank = world, rank

    def peer_search(self, lo: int, hi_excl: int, n_min: int, n_max: int, power_max: int, target: int, op: str = "==",
                    all_powers: bool = False):
        """this rank's part of a multi-GPU query over its value slice [lo, hi); returns (records of all ranks, stats)"""
        a, keep = self._make_args(n_min, n_max, power_max, all_powers, op, target)
        recs = getattr(self, "_peer_recs", None)          # reused across queries: the caller consumes the records at once
        if recs is None or len(recs) != self.peer_world:
            recs = self._peer_recs = (PeerRecord * self.peer_world)()
        st = Stats()
        self._check(self._L.pss_peer_search(self._h, lo, hi_excl, ctypes.byref(a), recs, ctypes.byref(st)))
        del keep
        return recs, stats_to_dict(st)

    def slice_sieve_sums(self, lo: int, hi_excl: int, power_max: int, all_powers: bool = False):
        """fused phase A: (count, [sum p^k per power record]) of the primes in [lo, hi)"""
        np_ = power_max if (all_powers and power_max > 1) else 1
        cnt = u64(0)
        out = np.zeros(np_ * PSS_LIMBS, dtype=np.uint32)
        self._check(self._L.pss_slice_sieve_sums(self._h, lo, hi_excl, power_max, 1 if all_powers else 0, ctypes.byref(cnt), _pu32(out)))
        return int(cnt.value), [limbs_to_int(out[j * PSS_LIMBS:(j + 1) * PSS_LIMBS]) for j in range(np_)]

    def slice_scan_bitmap(self, n_first: int, carry: Sequence[int], n_min: int, n_max: int, power_max: int, target: int,
                          op: str = "==", all_powers: bool = False) -> SearchResult:
        """fused phase B over the slice left resident by slice_sieve_sums"""
        a, keep = self._make_args(n_min, n_max, power_max, all_powers, op, target)
        np_ = power_max if (all_powers and power_max > 1) else 1
        cin = np.zeros(np_ * PSS_LIMBS, dtype=np.uint32)
        for j in range(np_):
            cin[j * PSS_LIMBS:(j + 1) * PSS_LIMBS] = int_to_limbs(int(carry[j]) if carry else 0, PSS_LIMBS)
        res = SearchResult()
        self._check(self._L.pss_slice_scan_bitmap(self._h, n_first, _pu32(cin), ctypes.byref(a), ctypes.byref(res)))
        del keep
        return res

    def nth_prime_upper(self, n: int) -> int:
        return int(self._L.pss_nth_prime_upper(n))


_default: Optional[Handle] = None


def default_handle() -> Handle:
    """Lazily created per-process handle (CUDA is not touched before the first compute call: fork safety)."""
    global _default
    if _default is None:
        dev = int(os.environ.get("PSS_DEVICE", os.environ.get("LOCAL_RANK", "-1")))
        _default = Handle(dev)
    return _default


def reset_default_handle():
    global _default
    if _default is not None:
        _default.close()
        _default = None
