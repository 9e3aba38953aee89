"""The C-ABI library loads and exports every symbol include/pss_b200.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pss_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pss_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ("pss_open", "pss_close", "pss_search", "pss_generate_primes", "pss_generate_n_primes", "pss_nth_prime",
                 "pss_count_primes", "pss_power_sum", "pss_primesum", "pss_slice_sieve_sums", "pss_slice_scan_bitmap"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from pss_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the CUDA library first: python -c 'import __graft_entry__ as g; g.build()'"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(L, name), f"{name} declared in include/pss_b200.h but not exported"
    assert L.pss_abi_version() == _lib.ABI_VERSION == 3
    # the ctypes binding covers the whole header too
    assert sorted(set(_lib.EXPORTS)) == declared_functions()


def test_struct_layouts_match_the_header():
    from pss_b200 import _lib
    # pss_power_result: 2*u32 + 5*u64 + 3*12*u32
    assert ctypes.sizeof(_lib.PowerResult) == 8 + 40 + 144
    assert ctypes.sizeof(_lib.Stats) == 5 * 8 + 4 * 8 + 3 * 8 + 4 * 8
    assert ctypes.sizeof(_lib.SearchArgs) == 8 + 8 + 4 * 4 + 8
    assert ctypes.sizeof(_lib.SearchResult) == 8 + 8 + 8 + 8 * ctypes.sizeof(_lib.PowerResult) + ctypes.sizeof(_lib.Stats)


def test_no_gpu_is_a_loud_error():
    """No CPU fallback: without a device the product raises RuntimeError("GPU not available ...")
    (reference convention: utils/gpu.py:160-161)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pss_b200 import _lib
    with pytest.raises(RuntimeError, match="GPU not available"):
        _lib.Handle(0)
    import pss_b200
    with pytest.raises(RuntimeError, match="GPU not available"):
        pss_b200.primesum(7, 2)
    assert pss_b200.gpu.init_gpu() is False and "error" in pss_b200.gpu.GPU_INFO


def test_nth_prime_upper_bound_is_an_upper_bound(golden):
    from pss_b200 import _lib
    L = _lib.load_library()
    known = {1: 2, 2: 3, 5: 11, 6: 13, 7: 17, 100: 541, 1000: 7919, 10 ** 4: 104729, 39016: 467473, 39017: 467479,
             10 ** 5: 1299709, 10 ** 6: 15485863, 5 * 10 ** 6: 86028121, 10 ** 7: 179424673, 5 * 10 ** 7: 982451653}
    for row in golden["at_scale"]:
        known[row["n"]] = row["p_n"]
    for n, p in known.items():
        b = L.pss_nth_prime_upper(n)
        assert b >= p, (n, p, b)
        if n >= 10 ** 6:
            assert b < p * 1.01      # tight: at most 1 % over-sieve (the reference over-sieves by 30-35 %)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "prime-square-sum_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
