"""The runnable front-end `python -m pss_b200.cli` against the UNMODIFIED reference CLI.

CPU tests (they need the reference tree, which exists in the build container only): the reference is run as a
subprocess; the front-end runs in-process with the device handle replaced by the oracle-backed stand-in of
test_host_logic.py, so what is checked here is the host logic — hooks, AST-shape recognition on the reference's own
parser output, bounds / iterator handling, output formatting and return codes.  The device side of the same query
shapes (fastpath.try_fast_query -> search.* -> pss_search) is covered on the GPU by
tests/test_gpu_parity.py::test_fastpath_reproduces_reference_cli_output.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PSS_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "prime-square-sum.py")),
                                reason="reference tree not present (it never travels to the GPU box)")

# command lines in the style of the reference's tests/test_cli_integration.py:156-182,257-267 plus the shapes VERDICT r1
# listed as silently falling off the fast path (extra parentheses, qualified name, reversed comparison)
CASES = [
    ["--expr", "does_exist primesum(n,2) == 666"],
    ["--expr", "primesum(n,2) == 666"],
    ["--expr", "does_exist primesum(n,2) == 667", "--max", "n:60"],
    ["--expr", "does_exist primesum(n,2) == 666", "--format", "json"],
    ["--expr", "does_exist primesum(n,2) == 667", "--max", "n:50", "--format", "json"],
    ["--expr", "does_exist primesum(n,2) == 667", "--max", "n:50", "--format", "csv"],
    ["--expr", "does_exist primesum(n,2) == 666", "--format", "csv"],
    ["--expr", "does_exist primesum(n,2) == trisum(10)", "--max", "n:100"],
    ["--expr", "does_exist (primesum(n, 2)) == (666)", "--max", "n:100"],
    ["--expr", "does_exist 666 == primesum(n,2)", "--max", "n:100"],
    ["--expr", "does_exist pss.primesum(n,2) == 666", "--max", "n:100"],
    ["--expr", "for_any 100 < primesum(n,2)", "--max", "n:12"],
    ["--expr", "for_any primesum(n,2) > 100", "--max", "n:40", "--limit", "3"],
    ["--expr", "does_exist primesum(n,2) == 666", "--max", "n:100", "--min", "n:8"],
    ["--expr", "for_any primesum(n,p) == 8944", "--max", "n:12", "--max", "p:3"],
    ["--expr", "for_any primesum(n,2) == tri(m)", "--max", "n:9", "--max", "m:40"],
    ["--expr", "does_exist primesum(n,2) == 600 + 66", "--max", "n:30"],
    ["--expr", "for_any primesum(n,1) == fibonacci(m)", "--max", "n:12", "--max", "m:12"],
    ["--target", "666"],
    ["--lhs", "primesum(n,3)", "--target", "8944", "--max", "n:20"],
]
# not hot-path shapes: must still give the reference's answer (generic evaluator behind the hooks)
GENERIC = [
    ["--expr", "does_exist primesum(n,2) + 1 == 667", "--max", "n:20"],
    ["--expr", "verify primesum(7,2) == 666"],
    ["--expr", "for_any tri(n)", "--max", "n:5"],
    ["--expr", "does_exist n**2 == 25", "--max", "n:10"],
]


def run_reference(argv):
    env = dict(os.environ, PRIME_SQUARE_SUM_QUIET="1")
    p = subprocess.run([sys.executable, os.path.join(REF, "prime-square-sum.py")] + argv, capture_output=True, text=True, cwd=REF,
                       env=env, timeout=300)
    return p.stdout, p.returncode


@pytest.fixture()
def device_stub(monkeypatch):
    from pss_b200 import _lib, sequences
    from test_host_logic import OracleHandle
    h = OracleHandle()
    h.slice_sieve = lambda lo, hi: (_ for _ in ()).throw(AssertionError("the generic primesum window is not under test here"))
    monkeypatch.setattr(_lib, "_default", h)
    monkeypatch.setenv("PSS_REFERENCE_ROOT", REF)
    monkeypatch.setenv("PRIME_SQUARE_SUM_QUIET", "1")
    monkeypatch.chdir(REF)          # the reference finds equations.json / config.json relative to the working directory
    sequences.clear_cache()
    yield h


def run_frontend(argv, capsys):
    from pss_b200 import cli
    capsys.readouterr()
    try:
        rc = cli.main(list(argv))
    except SystemExit as e:
        rc = e.code
    return capsys.readouterr().out, rc


@pytest.mark.parametrize("argv", CASES, ids=lambda a: " ".join(a))
def test_hot_path_shapes_take_the_device_and_print_like_the_reference(argv, device_stub, capsys):
    from pss_b200 import cli
    want = run_reference(argv)
    before = cli._HOOKS.fast_queries if cli._HOOKS else 0
    got = run_frontend(argv + ["--algorithm", "sieve:b200"], capsys)
    assert got == want, argv
    assert cli._HOOKS.fast_queries == before + 1, "the expression was not answered by the fused search"
    # --prefer gpu selects the variant as well
    before = cli._HOOKS.fast_queries
    assert run_frontend(argv + ["--algorithm", "sieve:auto", "--prefer", "gpu"], capsys) == want
    assert cli._HOOKS.fast_queries == before + 1


@pytest.mark.parametrize("argv", GENERIC, ids=lambda a: " ".join(a))
def test_other_shapes_stay_with_the_generic_evaluator(argv, device_stub, capsys):
    from pss_b200 import cli
    want = run_reference(argv)
    before = cli._HOOKS.fast_queries if cli._HOOKS else 0
    device_stub.slice_sieve = None     # generic primesum may use the plugin window: not available in the stub -> use reference primesum
    got = run_frontend(argv + ["--no-gpu", "--algorithm", "sieve:b200"], capsys)
    assert got == want and (cli._HOOKS.fast_queries == before)


def test_without_selection_the_frontend_is_the_reference(capsys, monkeypatch):
    """no --algorithm sieve:b200 / --prefer gpu: every hook passes through (fresh process, no device needed)"""
    env = dict(os.environ, PSS_REFERENCE_ROOT=REF, PRIME_SQUARE_SUM_QUIET="1",
               PYTHONPATH=os.path.join(ROOT, "prime-square-sum_b200") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    for argv in (["--expr", "does_exist primesum(n,2) == 666"], ["--expr", "for_any primesum(n,2) == 38", "--max", "n:5", "--format", "json"]):
        p = subprocess.run([sys.executable, "-m", "pss_b200.cli"] + argv, capture_output=True, text=True, cwd=REF, env=env, timeout=300)
        assert (p.stdout, p.returncode) == run_reference(argv)


def test_list_algorithms_and_config_selection(tmp_path):
    env = dict(os.environ, PSS_REFERENCE_ROOT=REF, PRIME_SQUARE_SUM_QUIET="1", CUDA_VISIBLE_DEVICES="",
               PYTHONPATH=os.path.join(ROOT, "prime-square-sum_b200") + os.pathsep + os.environ.get("PYTHONPATH", ""))
    p = subprocess.run([sys.executable, "-m", "pss_b200.cli", "--list", "algorithms"], capture_output=True, text=True, cwd=REF, env=env, timeout=300)
    ref_out, ref_rc = run_reference(["--list", "algorithms"])
    assert p.returncode == ref_rc == 0
    assert "  b200         [" in p.stdout and "--algorithm sieve:b200" in p.stdout
    # everything the reference prints is still there, in order
    it = iter(p.stdout.splitlines())
    assert all(any(line == got for got in it) for line in ref_out.splitlines())
    # the reference itself rejects the variant; the front-end accepts it (and fails loudly without a device: no CPU fallback)
    assert run_reference(["--algorithm", "sieve:b200", "--expr", "primesum(n,2) == 666"])[1] == 1
    p = subprocess.run([sys.executable, "-m", "pss_b200.cli", "--algorithm", "sieve:b200", "--expr", "primesum(n,2) == 666"],
                       capture_output=True, text=True, cwd=REF, env=env, timeout=300)
    assert p.returncode == 1 and "GPU not available" in p.stderr and "Found" not in p.stdout
    # config.json algorithms.sieve = "b200" reaches the hooks (reference: prime-square-sum.py:546-550)
    (tmp_path / "config.json").write_text('{"algorithms": {"sieve": "b200"}}')
    p = subprocess.run([sys.executable, "-m", "pss_b200.cli", "--expr", "primesum(n,2) == 666"], capture_output=True, text=True,
                       cwd=tmp_path, env=env, timeout=300)
    assert p.returncode == 1 and "GPU not available" in p.stderr
