"""
cli.py — a runnable front-end:  python -m pss_b200.cli <the reference's own command line>

It imports the UNMODIFIED reference (prime-square-sum.py + utils/, located by $PSS_REFERENCE_ROOT; nothing of it is
copied into this package), installs a handful of run-time hooks, and then hands control to the reference's own
`main()`.  What the hooks add (the reference lines they attach to):

  --algorithm sieve:b200       "b200" is appended to utils.sieve.VALID_ALGORITHMS (:56-62), so the CLI accepts it
                               (prime-square-sum.py:562-573) and config.json `algorithms.sieve: "b200"` (:546-550) works
  --prefer gpu                 with the algorithm left on "auto", selects the b200 variant (the reference stores the
                               preference and never reads it: SURVEY §0.3); with --algorithm sieve:individual it selects
                               the trial-division variant as a device kernel
  generate_primes / generate_primes_range / generate_n_primes / nth_prime / prime_count   (utils/sieve.py:440-641)
                               dispatch to pss_b200.sieve when the b200 variant is selected — also through the names
                               other reference modules imported (`from utils.sieve import generate_n_primes`)
  primesum                     (utils/sequences.py:30-75) served by pss_b200.sequences.primesum (device window)
  --list algorithms            (utils/list_commands.py:60-97) also lists the b200 variant
  handle_expression            (prime-square-sum.py:339-472): the PARSED expression goes through
                               pss_b200.fastpath.recognise; a hot-path shape
                               `primesum(n[,p]) <op> <closed term>` (either operand order, qualified names,
                               --min / --max / --limit / --format honoured) is answered by ONE fused device search and
                               printed with the reference's own format_match / format_no_match; everything else runs
                               the untouched generic evaluator.

Without `--algorithm sieve:b200` / `--prefer gpu` / a config.json that selects b200 — or with `--no-gpu` — every hook is
a pass-through and the command behaves exactly like the reference.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys
import time
from typing import Optional

ALGORITHM_B200 = "b200"
ENTRY_SCRIPT = "prime-square-sum.py"


def find_reference_root() -> str:
    for cand in (os.environ.get("PSS_REFERENCE_ROOT"), os.getcwd()):
        if cand and os.path.isfile(os.path.join(cand, ENTRY_SCRIPT)) and os.path.isdir(os.path.join(cand, "utils")):
            return os.path.abspath(cand)
    raise SystemExit("pss_b200.cli: set PSS_REFERENCE_ROOT to a checkout of djdarcy/Prime-Square-Sum "
                     f"(the directory holding {ENTRY_SCRIPT} and utils/), or run from inside it")


def load_reference(root: str):
    """import the reference's entry script as a module (its file name is not an identifier) and return it"""
    if root not in sys.path:
        sys.path.insert(0, root)
    spec = importlib.util.spec_from_file_location("pss_reference_entry", os.path.join(root, ENTRY_SCRIPT))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["pss_reference_entry"] = mod
    spec.loader.exec_module(mod)
    return mod


class Hooks:
    """the run-time hooks; `selected()` decides at CALL time whether the b200 variant is in force"""

    def __init__(self, entry, no_gpu: bool):
        self.entry = entry
        self.no_gpu = no_gpu
        self.rs = sys.modules["utils.sieve"]
        self.fast_queries = 0            # expressions answered by the fused device search (for tests / -v output)

    def selected(self, algorithm: Optional[str] = None) -> Optional[str]:
        """the pss_b200.sieve variant in force ("b200" / "individual"), or None when the call belongs to the reference"""
        if self.no_gpu:
            return None
        cfg = self.rs.get_config()
        algo = algorithm if algorithm not in (None, "auto") else cfg.get("algorithm", "auto")
        gpu = cfg.get("prefer") == "gpu"
        if algo == ALGORITHM_B200 or (algo == "auto" and gpu):
            return ALGORITHM_B200
        if algo == "individual" and gpu:
            return "individual"          # --algorithm sieve:individual --prefer gpu: the trial-division variant as a device kernel
        return None

    # ---- rebinding: a reference module that did `from utils.sieve import f` holds its own reference to f
    @staticmethod
    def _rebind(old, new):
        for name, mod in list(sys.modules.items()):
            if mod is None or not (name == "pss_reference_entry" or name == "utils" or name.startswith("utils.")):
                continue
            for attr, val in list(vars(mod).items()):
                if val is old:
                    setattr(mod, attr, new)

    def install(self):
        from . import sequences as b_seq
        from . import sieve as b_sieve
        rs, entry = self.rs, self.entry
        if ALGORITHM_B200 not in rs.VALID_ALGORITHMS:
            rs.VALID_ALGORITHMS.append(ALGORITHM_B200)

        def route(name, pick_algorithm):
            ref_fn, dev_fn = getattr(rs, name), getattr(b_sieve, name)

            def dispatch(*args, **kwargs):
                variant = self.selected(pick_algorithm(args, kwargs))
                if variant:
                    return dev_fn(*args[:pick_algorithm.n_positional], algorithm=variant)
                return ref_fn(*args, **kwargs)

            dispatch.__name__, dispatch.__doc__ = ref_fn.__name__, ref_fn.__doc__
            dispatch.__wrapped__ = ref_fn
            self._rebind(ref_fn, dispatch)

        def picker(n_positional, keep=()):
            def pick(args, kwargs):
                return kwargs.get("algorithm", args[n_positional] if len(args) > n_positional else None)
            pick.n_positional, pick.keep = n_positional, keep
            return pick

        route("generate_primes", picker(1))
        route("generate_primes_range", picker(2))
        route("generate_n_primes", picker(1))
        route("nth_prime", picker(1))
        route("prime_count", picker(1))

        # primesum: the evaluator calls it once per n; the device version serves windows of exact prefix sums
        ref_seq = sys.modules.get("utils.sequences")
        if ref_seq is not None:
            ref_primesum = ref_seq.primesum

            def primesum(n, power=2, *, _cache=None):
                if self.selected() and isinstance(power, int) and 0 <= power <= 8:
                    return b_seq.primesum(n, power)
                return ref_primesum(n, power) if _cache is None else ref_primesum(n, power, _cache=_cache)

            primesum.__doc__, primesum.__wrapped__ = ref_primesum.__doc__, ref_primesum
            self._rebind(ref_primesum, primesum)

        # algorithm listings
        ref_info, ref_avail = rs.get_algorithm_info, rs.get_available_algorithms

        def get_algorithm_info():
            info = dict(ref_info())
            info[ALGORITHM_B200] = b_sieve.get_algorithm_info()[ALGORITHM_B200]
            info[ALGORITHM_B200]["available"] = self._device_available()
            return info

        def get_available_algorithms():
            return list(ref_avail()) + ([ALGORITHM_B200] if self._device_available() else [])

        self._rebind(ref_info, get_algorithm_info)
        self._rebind(ref_avail, get_available_algorithms)

        lc = sys.modules.get("utils.list_commands")
        if lc is not None:
            ref_list = lc.handle_list_algorithms

            def handle_list_algorithms():
                buf = io.StringIO()
                with contextlib.redirect_stdout(buf):
                    rc = ref_list()
                text = buf.getvalue()
                info = get_algorithm_info()[ALGORITHM_B200]
                block = (f"  {ALGORITHM_B200:12} [{'OK' if info['available'] else 'NO DEVICE'}]\n"
                         f"               {info['description']}\n"
                         f"               Time: {info['complexity_time']}, Space: {info['complexity_space']}\n\n")
                usage = "  --algorithm sieve:b200        B200 GPU sieve + fused device search (also: --prefer gpu)\n"
                if "Usage:\n" in text:
                    head, tail = text.split("Usage:\n", 1)
                    first_blank = tail.find("\n\n")
                    tail = tail[:first_blank + 1] + usage + tail[first_blank + 1:] if first_blank >= 0 else tail + usage
                    text = head + block + "Usage:\n" + tail
                else:
                    text += block + usage
                sys.stdout.write(text)
                return rc

            self._rebind(ref_list, handle_list_algorithms)

        ref_handle = entry.handle_expression

        def handle_expression(args, registry, config=None):
            rc = self._fast_expression(args, registry, config)
            return ref_handle(args, registry, config) if rc is None else rc

        entry.handle_expression = handle_expression

    _dev_ok: Optional[bool] = None

    def _device_available(self) -> bool:
        if self._dev_ok is None:
            try:
                from . import _lib
                _lib.default_handle()
                Hooks._dev_ok = True
            except Exception:  # noqa: BLE001
                Hooks._dev_ok = False
        return bool(self._dev_ok)

    # ---- the expression hook
    def _fast_expression(self, args, registry, config) -> Optional[int]:
        """answer a hot-path expression with one fused device search; None = not ours, run the reference's handler"""
        if not self.selected():
            return None
        from . import fastpath
        entry = self.entry
        out = entry.get_output()
        if out.channel_active("vals"):
            return None                      # --show vals prints the compared values of every match: generic path
        t0 = time.time()
        try:
            expr_str = entry.build_expression_from_args(args)
            imag = {"ii": "[iI][iI]", "i": "[iI]", "j": "[jJ]", "both": "[iIjJ]", "none": ""}
            unit = config.imaginary_unit if config else "none"
            ast = entry.ExpressionParser(imaginary_suffix_pattern=imag.get(unit, "[iI][iI]")).parse(expr_str)
            if entry.validate_expression(ast, registry):
                return None                  # the reference reports validation errors itself
        except Exception:  # noqa: BLE001 — parse / argument errors: the reference prints them in its own words
            return None
        evaluator = entry.ExpressionEvaluator(registry)

        def is_builtin(name: str) -> bool:
            base = name.split(".", 1)[1] if "." in name else name
            try:
                return registry.get(name) is registry.get("pss." + base)
            except Exception:  # noqa: BLE001
                return False

        q = fastpath.recognise(ast, eval_closed=lambda node: evaluator.evaluate(node, {}), is_builtin=is_builtin)
        if q is None:
            return None
        try:
            bounds = entry.build_bounds_from_args(args, expr_str)
            factories = entry.build_iterator_factories_from_args(args, bounds)
        except Exception:  # noqa: BLE001
            return None
        ranges = {}
        for v in fastpath.variables_of(q):
            if v in factories:
                it = factories[v]()
                if type(it).__name__ != "IntIterator" or getattr(it, "_step", None) != 1:
                    return None              # float / stepped / custom iterators: generic path
                ranges[v] = (int(it._start), int(it._stop))
            elif v in bounds:
                ranges[v] = (1, int(bounds[v]))
            else:
                return None                  # the reference raises its "Missing bounds" error
        limit = getattr(args, "limit", None)
        out.emit(1, "Expression: {expr}", channel="parse", expr=expr_str)
        out.emit(1, "[b200] fused device search over {r}", channel="algorithm", r=", ".join(f"{k}={a}..{b}" for k, (a, b) in sorted(ranges.items())))
        try:
            found = fastpath.matches(q, ranges, limit)
        except (ValueError, RuntimeError) as e:      # bad arguments / "GPU not available": the reference's error style
            out.error(f"Error: {e}")
            return 1
        if found is None:
            return None
        self.fast_queries += 1
        for m in found:
            print(entry.format_match(m, args.format))
        if not found:
            print(entry.format_no_match(args.format))
        out.emit(1, "", channel="timing")
        out.emit(1, "Matches: {count}", channel="timing", count=len(found))
        out.emit(1, "Time: {elapsed:.2f}s", channel="timing", elapsed=time.time() - t0)
        # the reference's post-result hints (prime-square-sum.py:460-470)
        does_exist = q.quantifier == "does_exist"
        two_vars = len(fastpath.variables_of(q)) >= 2
        if ((does_exist and found and two_vars) or (does_exist and limit is not None)) and not out.quiet:
            out.emit(0, "", channel="hint")
        if does_exist and found and two_vars:
            out.hint("quantifier.use_for_any", "result")
        if does_exist and limit is not None:
            out.hint("quantifier.limit_with_does_exist", "result")
        return 0 if found else 1


_HOOKS: Optional[Hooks] = None


def main(argv=None) -> int:
    global _HOOKS
    argv = list(sys.argv[1:] if argv is None else argv)
    root = find_reference_root()
    entry = load_reference(root)
    if _HOOKS is None:
        _HOOKS = Hooks(entry, no_gpu="--no-gpu" in argv)
        _HOOKS.install()
    else:
        # a second main() in the same process (tests): the utils.* hooks are in place, the fresh entry module has
        # imported the hooked names; only its expression handler needs hooking again
        _HOOKS.entry, _HOOKS.no_gpu = entry, "--no-gpu" in argv
        ref_handle = entry.handle_expression
        hooks = _HOOKS

        def handle_expression(args, registry, config=None):
            rc = hooks._fast_expression(args, registry, config)
            return ref_handle(args, registry, config) if rc is None else rc

        entry.handle_expression = handle_expression
    old_argv = sys.argv
    sys.argv = [os.path.join(root, ENTRY_SCRIPT)] + argv
    try:
        rc = entry.main()
    finally:
        sys.argv = old_argv
    return 0 if rc is None else int(rc)


if __name__ == "__main__":
    sys.exit(main())
