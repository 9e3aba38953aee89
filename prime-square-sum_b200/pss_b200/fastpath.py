"""
fastpath.py — AST-shape recogniser for the query shapes of the hot path, so that the reference front-end can reach
the fused device search at scale.

The reference evaluates `does_exist primesum(n,2) == trisum(666) --max n:1000000000` as 1e9 independent
`primesum` calls after materialising the iterator range as a Python list (utils/grammar.py:951,977-980).
`recognise()` looks at the PARSED expression (the reference's own AST from utils/grammar.py:42-106 when called from
pss_b200.cli, or the equivalent tree of the small parser below when only a string is at hand) and accepts

    [does_exist | for_any]   primesum(<var> [, <int> | <var>])  <op>  <closed term>          op in == != < <= > >=
    ... or the same comparison written the other way round:       <closed term>  <op'>  primesum(...)
    ... or   primesum(<var> [, <int>]) == f(<var>)   (either order)   f in tri, fibonacci, factorial, catalan

where `primesum` may be qualified (`pss.primesum`), operands may be wrapped in parentheses, and a closed term is any
sub-expression without free variables that evaluates to a non-negative integer: an integer literal, trisum(666),
tri(36), 2**10 + 1, ... (closed terms beyond literals and the five known functions are evaluated by the caller's
evaluator — pss_b200.cli passes the reference's).  Everything else is NOT a hot-path shape: `recognise` returns None
and the query stays with the generic evaluator.

`run()` answers a recognised query with ONE device pass and produces exactly the reference's stdout lines and return
code (format_match / format_no_match, utils/cli.py:1201-1297; rc 0 = match, 1 = none, prime-square-sum.py:452-472).
Iteration order, inclusive bounds and defaults follow the reference: variables sorted by name with the first
outermost (utils/grammar.py:939), `--min`/`--max` inclusive (utils/iterators.py:99), default quantifier does_exist
for a comparison with free variables (utils/grammar.py:862-865), default bounds n: 1 000 000, m: 10 000
(utils/cli.py:57-60).
"""
from __future__ import annotations

import json
import re
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

from . import _lib
from . import search as _search

DEFAULT_BOUNDS = {"n": 1_000_000, "m": 10_000}
_FLIP = {"==": "==", "!=": "!=", "<": ">", ">": "<", "<=": ">=", ">=": "<="}
_SEQ_FUNCS = ("tri", "fibonacci", "factorial", "catalan")
MAX_ALL_POWERS = 6            # the device's all-powers pass (pss_search with all_powers) carries at most S_1..S_6


# ---------------------------------------------------------------------------------------- a small AST + parser
# Same class names and field names as the reference's nodes (utils/grammar.py:42-106): the shape matcher below works
# on either by duck typing.  The parser covers exactly the sub-grammar the hot-path shapes need; anything else
# (arithmetic, boolean operators, context blocks, chained comparisons) is a parse failure = "not a hot-path shape".
@dataclass
class Literal:
    value: int


@dataclass
class Variable:
    name: str


@dataclass
class FunctionCall:
    name: str
    args: List[object] = field(default_factory=list)


@dataclass
class Comparison:
    operands: List[object]
    operators: List[str]


@dataclass
class Expression:
    quantifier: Optional[str]
    body: object


_TOKEN = re.compile(r"\s*(?:(?P<int>\d+)|(?P<name>[A-Za-z_]\w*(?:\.[A-Za-z_]\w*)?)|(?P<op>==|!=|<=|>=|<|>)|(?P<punct>[(),]))")


class _NotHotPath(Exception):
    pass


def parse(text: str) -> Optional[Expression]:
    """the sub-grammar  [quantifier] term OP term  (terms: integer, name, call, parenthesised term); None otherwise"""
    toks: List[Tuple[str, str]] = []
    pos = 0
    text = text.rstrip()
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            return None
        toks.append((m.lastgroup, m.group(m.lastgroup)))
        pos = m.end()
    i = 0

    def peek():
        return toks[i] if i < len(toks) else (None, None)

    def take(kind=None, val=None):
        nonlocal i
        k, v = peek()
        if k is None or (kind and k != kind) or (val and v != val):
            raise _NotHotPath
        i += 1
        return v

    def term():
        k, v = peek()
        if k == "int":
            take()
            return Literal(int(v))
        if k == "punct" and v == "(":
            take()
            t = term()
            take("punct", ")")
            return t
        if k == "name":
            take()
            if peek() == ("punct", "("):
                take()
                args = []
                if peek() != ("punct", ")"):
                    args.append(term())
                    while peek() == ("punct", ","):
                        take()
                        args.append(term())
                take("punct", ")")
                return FunctionCall(v, args)
            return Variable(v)
        raise _NotHotPath

    def comparison():
        nonlocal i
        start = i
        try:
            left = term()
            op = take("op")
            right = term()
            return Comparison([left, right], [op])
        except _NotHotPath:
            i = start
            take("punct", "(")
            c = comparison()
            take("punct", ")")
            return c

    try:
        quant = None
        if peek()[0] == "name" and peek()[1] in ("does_exist", "for_any", "verify", "solve"):
            quant = take()
        body = comparison()
        if i != len(toks):
            return None
        return Expression(quant, body)
    except _NotHotPath:
        return None


# ---------------------------------------------------------------------------------------------- shape matcher
@dataclass
class Query:
    quantifier: str                  # "does_exist" | "for_any"
    nvar: str
    op: str                          # as seen with primesum on the LEFT
    power: Optional[int] = None      # literal power (default 2) ...
    pvar: Optional[str] = None       # ... or the free power variable
    target: Optional[int] = None     # constant right-hand side ...
    seq_fn: Optional[str] = None     # ... or f(mvar) with f a non-decreasing sequence
    mvar: Optional[str] = None


def _kind(node) -> str:
    return type(node).__name__


def _base_name(name: str) -> str:
    return name.split(".", 1)[1] if name.startswith(("pss.", "math.")) else name


def _tri(n: int) -> int:
    return n * (n + 1) // 2


def _trisum(b: int) -> int:
    """utils/number_theory.py:397-479 (rows of length r, r-1, .., 1 read as base-b numbers)"""
    if b < 1:
        raise ValueError(f"trisum() requires b >= 1, got {b}")
    if b == 1:
        return 0
    rows = 1
    while _tri(rows + 1) <= b:
        rows += 1
    total, digit = 0, 0
    for length in range(rows, 0, -1):
        value = 0
        for d in range(digit, min(digit + length, b)):
            value = value * b + d
        digit += length
        total += value
        if digit >= b:
            break
    return total


_CLOSED_FUNCS: Dict[str, Callable[[int], int]] = {"trisum": _trisum, **_search.RHS_SEQUENCES}


def _has_free_variable(node) -> bool:
    k = _kind(node)
    if k == "Variable":
        return True
    if k == "Literal":
        return False
    for attr in ("args", "operands"):
        if hasattr(node, attr):
            return any(_has_free_variable(a) for a in getattr(node, attr))
    return any(_has_free_variable(getattr(node, a)) for a in ("left", "right", "operand", "body") if hasattr(node, a))


def _int_literal(node) -> Optional[int]:
    if _kind(node) == "Literal" and isinstance(node.value, int) and not isinstance(node.value, bool):
        return int(node.value)
    return None


def _closed_value(node, eval_closed) -> Optional[int]:
    """non-negative integer value of a sub-expression without free variables, or None"""
    v = _int_literal(node)
    if v is not None:
        return v if v >= 0 else None
    if _has_free_variable(node):
        return None
    if _kind(node) == "FunctionCall" and len(node.args) == 1 and _base_name(node.name) in _CLOSED_FUNCS and eval_closed is None:
        a = _int_literal(node.args[0])
        if a is None:
            return None
        try:
            return _CLOSED_FUNCS[_base_name(node.name)](a)
        except ValueError:
            return None
    if eval_closed is not None:
        try:
            v = eval_closed(node)
        except Exception:  # noqa: BLE001 — the generic evaluator will report it
            return None
        if isinstance(v, bool):
            return None
        if isinstance(v, float) and v == int(v) and abs(v) < 2 ** 53:
            v = int(v)
        if isinstance(v, int) and v >= 0:
            return v
    return None


def _primesum_call(node, is_builtin) -> Optional[Tuple[str, Optional[int], Optional[str]]]:
    """(nvar, literal power | None, power variable | None) when node is primesum(<var> [, <int> | <var>])"""
    if _kind(node) != "FunctionCall" or _base_name(node.name) != "primesum" or not 1 <= len(node.args) <= 2:
        return None
    if not is_builtin(node.name):
        return None
    if _kind(node.args[0]) != "Variable":
        return None
    nvar = node.args[0].name
    if len(node.args) == 1:
        return nvar, 2, None
    second = node.args[1]
    lit = _int_literal(second)
    if lit is not None:
        return nvar, lit, None
    if _kind(second) == "Variable" and second.name != nvar:
        return nvar, None, second.name
    return None


def recognise(expr, eval_closed: Optional[Callable] = None, is_builtin: Callable[[str], bool] = lambda name: True) -> Optional[Query]:
    """Query description of a parsed expression, or None when it is not a hot-path shape.

    eval_closed(node) -> value    evaluator for closed sub-expressions (the reference's, when called from the CLI)
    is_builtin(name) -> bool      False when `name` resolves to a user plugin function instead of the built-in"""
    if _kind(expr) != "Expression" or expr.quantifier not in (None, "does_exist", "for_any"):
        return None
    body = expr.body
    if _kind(body) != "Comparison" or len(body.operands) != 2 or len(body.operators) != 1:
        return None
    op = body.operators[0]
    if op not in _FLIP:
        return None
    left, right = body.operands
    ps = _primesum_call(left, is_builtin)
    other = right
    if ps is None:
        ps = _primesum_call(right, is_builtin)
        other, op = left, _FLIP[op]
    if ps is None:
        return None
    nvar, power, pvar = ps
    quant = expr.quantifier or "does_exist"          # a comparison with free variables defaults to a search
    if power is not None and not 1 <= power <= _lib.PSS_MAX_POWER:
        return None
    if pvar is not None and pvar < nvar:
        return None                                   # p outer / n inner: not the order the fused pass enumerates
    target = _closed_value(other, eval_closed)
    if target is not None:
        return Query(quant, nvar, op, power=power, pvar=pvar, target=target)
    # primesum(n,k) == f(m): one free variable m as the only argument of a non-decreasing sequence
    if (_kind(other) == "FunctionCall" and _base_name(other.name) in _SEQ_FUNCS and len(other.args) == 1 and
            _kind(other.args[0]) == "Variable" and op == "==" and pvar is None and is_builtin(other.name)):
        mvar = other.args[0].name
        if mvar != nvar:
            return Query(quant, nvar, op, power=power, seq_fn=_base_name(other.name), mvar=mvar)
    return None


# ---------------------------------------------------------------------------------------------- execution
def format_match(match: Dict[str, int], fmt: str = "text") -> str:
    items = sorted(match.items())
    if fmt == "json":
        return json.dumps({"found": True, "variables": {k: v for k, v in items}})
    if fmt == "csv":
        return ",".join(f"{k}={v}" for k, v in items)
    return "Found: " + ", ".join(f"{k}={v}" for k, v in items)


def format_no_match(fmt: str = "text") -> str:
    if fmt == "json":
        return json.dumps({"found": False, "variables": None})
    if fmt == "csv":
        return "not_found"
    return "No match found within bounds."


def variables_of(q: Query) -> List[str]:
    return sorted(v for v in (q.nvar, q.pvar, q.mvar) if v)


def matches(q: Query, ranges: Dict[str, Tuple[int, int]], limit: Optional[int] = None,
            handle: Optional["_lib.Handle"] = None) -> Optional[List[Dict[str, int]]]:
    """The matches of a recognised query in the reference's enumeration order, from one device pass.
    ranges: inclusive (lo, hi) per variable.  None when the ranges put the query outside what the device pass
    enumerates (power variable beyond 1..6, n below 1): the caller falls back to the generic evaluator."""
    n_lo, n_hi = ranges[q.nvar]
    if n_lo < 1:
        return None
    if n_hi < n_lo:
        return []
    cap = 1 if q.quantifier == "does_exist" else limit
    if q.seq_fn is not None:
        m_lo, m_hi = ranges[q.mvar]
        power = q.power
        if q.seq_fn == "tri":
            if m_hi >= 2 ** 29:
                return None
            hits = _search.for_any_tri(n_hi, m_hi, power=power, n_min=n_lo, m_min=m_lo, handle=handle)
        else:
            hits = _search.for_any_sequence(q.seq_fn, n_hi, m_hi, power=power, n_min=n_lo, m_min=m_lo, handle=handle)
        out = [{q.mvar: h["m"], q.nvar: h["n"]} for h in hits]
        out.sort(key=lambda d: tuple(d[v] for v in sorted(d)))
        return out if cap is None else out[:cap]
    if q.pvar is not None:
        p_lo, p_hi = ranges[q.pvar]
        if p_lo < 1 or p_hi > MAX_ALL_POWERS:
            return None
        if p_hi < p_lo:
            return []
        power, allp = p_hi, True
    else:
        p_lo, power, allp = 1, q.power, False
    res = _search.search(q.target, n_hi, power=power, n_min=n_lo, op=q.op, all_powers=allp, handle=handle)
    out = []
    for n, p in res.matches(None):
        if q.pvar is not None and p < p_lo:
            continue
        out.append({q.nvar: n, q.pvar: p} if q.pvar is not None else {q.nvar: n})
        if cap is not None and len(out) >= cap:
            break
    return out


def try_fast_query(expr: str, max_bounds: Optional[Dict[str, int]] = None, min_bounds: Optional[Dict[str, int]] = None,
                   fmt: str = "text", limit: Optional[int] = None) -> Optional[Tuple[List[str], int]]:
    """String front door: (stdout lines, return code) of the reference CLI for a hot-path expression, or None when the
    expression is not a hot-path shape (the caller then runs the generic evaluator)."""
    tree = parse(expr)
    if tree is None:
        return None
    q = recognise(tree)
    if q is None:
        return None
    maxb = dict(DEFAULT_BOUNDS)
    maxb.update(max_bounds or {})
    minb = min_bounds or {}
    missing = [v for v in variables_of(q) if v not in maxb]
    if missing:        # the reference's own wording (utils/grammar.py:929-935)
        raise ValueError(f"Missing bounds for variable(s): {', '.join(missing)}. "
                         f"Use {', '.join(f'--max {v}:VALUE' for v in missing)} to specify.")
    ranges = {v: (int(minb.get(v, 1)), int(maxb[v])) for v in variables_of(q)}
    found = matches(q, ranges, limit)
    if found is None:
        return None
    if not found:
        return [format_no_match(fmt)], 1
    return [format_match(m, fmt) for m in found], 0
