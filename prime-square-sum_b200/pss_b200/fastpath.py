"""
fastpath.py — recogniser for the query shapes of the hot path, so that the reference CLI can reach the
fused device search at scale.

The reference evaluates `does_exist primesum(n,2) == trisum(666) --max n:1000000000` as 1e9 independent
`primesum` calls after materialising the iterator range as a Python list (utils/grammar.py:951,977-980).
`try_fast_query` recognises

    [does_exist|for_any] primesum(<var>[, <int>|<var>]) <op> <rhs>       op in == != < <= > >=
    rhs:  <int literal> | trisum(<int>) | tri(<int>) | fibonacci(<int>) | factorial(<int>) | catalan(<int>)
          | tri(<var>) | fibonacci(<var>) | factorial(<var>) | catalan(<var>)     (only == for f(<var>))

and answers it with ONE device pass, producing exactly the reference's stdout lines and return code
(format_match / format_no_match, utils/cli.py:1201-1297; rc 0 = match, 1 = none, prime-square-sum.py:452-472).
Anything else returns None and stays with the generic evaluator (which can still use the B200 through the
`--functions` plugin).  Iteration order, inclusive bounds and defaults follow the reference: variables sorted
by name with the first outermost (utils/grammar.py:939), `--min`/`--max` inclusive (utils/iterators.py:99),
default quantifier does_exist, default bounds n: 1 000 000, m: 10 000 (utils/cli.py:57-60).
"""
from __future__ import annotations

import json
import re
from typing import Dict, List, Optional, Tuple

from . import search as _search

DEFAULT_BOUNDS = {"n": 1_000_000, "m": 10_000}

_EXPR = re.compile(
    r"^\s*(?:(?P<quant>does_exist|for_any)\s+)?primesum\(\s*(?P<nvar>[A-Za-z_]\w*)\s*(?:,\s*(?P<parg>[A-Za-z_]\w*|\d+)\s*)?\)"
    r"\s*(?P<op>==|!=|<=|>=|<|>)\s*(?P<rhs>.+?)\s*$")
_RHS_INT = re.compile(r"^\d+$")
_RHS_CALL = re.compile(r"^(?P<fn>trisum|tri|fibonacci|factorial|catalan)\(\s*(?P<arg>[A-Za-z_]\w*|\d+)\s*\)$")


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


def try_fast_query(expr: str, max_bounds: Optional[Dict[str, int]] = None, min_bounds: Optional[Dict[str, int]] = None,
                   fmt: str = "text", limit: Optional[int] = None) -> Optional[Tuple[List[str], int]]:
    """Returns (stdout lines, return code) or None when the expression is not a hot-path shape."""
    m = _EXPR.match(expr)
    if not m:
        return None
    quant = m.group("quant") or "does_exist"
    nvar, parg, op, rhs = m.group("nvar"), m.group("parg"), m.group("op"), m.group("rhs")
    maxb = dict(DEFAULT_BOUNDS)
    maxb.update(max_bounds or {})
    minb = min_bounds or {}

    def bound(var):
        if var not in maxb:
            raise ValueError(f"Missing bound for variable '{var}'")      # utils/grammar.py:493-503 wording family
        return int(minb.get(var, 1)), int(maxb[var])

    n_lo, n_hi = bound(nvar)
    # power: default 2, literal, or a free variable iterated 1..max (inner loop: p sorts after n)
    p_var = None
    if parg is None:
        power = 2
    elif parg.isdigit():
        power = int(parg)
    else:
        p_var = parg
        p_lo, power = bound(p_var)
    if power < 1 or power > 8 or n_lo < 1 or n_hi < n_lo:
        return None

    mt = _RHS_CALL.match(rhs)
    if _RHS_INT.match(rhs):
        target = int(rhs)
    elif mt and mt.group("arg").isdigit():
        fn, arg = mt.group("fn"), int(mt.group("arg"))
        target = _trisum(arg) if fn == "trisum" else _search.RHS_SEQUENCES[fn](arg)
    elif mt and mt.group("fn") in ("fibonacci", "factorial", "catalan") and op == "==" and p_var is None:
        # primesum(n,k) == f(m) for a non-decreasing sequence f: the reachable values f(m) are searched in one device pass
        mvar = mt.group("arg")
        if mvar == nvar:
            return None
        m_lo, m_hi = bound(mvar)
        hits = _search.for_any_sequence(mt.group("fn"), n_hi, m_hi, power=power, n_min=n_lo, m_min=m_lo)
        hits = [{mvar: h["m"], nvar: h["n"]} for h in hits]
        hits.sort(key=lambda d: tuple(d[v] for v in sorted(d)))
        if quant == "does_exist":
            hits = hits[:1]
        elif limit is not None:
            hits = hits[:limit]
        if not hits:
            return [format_no_match(fmt)], 1
        return [format_match(h, fmt) for h in hits], 0
    elif mt and mt.group("fn") == "tri" and op == "==" and p_var is None:
        # primesum(n,k) == tri(m): the device tests every partial sum for triangularity
        mvar = mt.group("arg")
        if mvar == nvar:
            return None
        m_lo, m_hi = bound(mvar)
        hits = _search.for_any_tri(n_hi, m_hi, power=power, n_min=n_lo, m_min=m_lo)
        hits = [{mvar: h["m"], nvar: h["n"]} for h in hits]
        # iteration order: variables sorted by name, first outermost
        hits.sort(key=lambda d: tuple(d[v] for v in sorted(d)))
        if quant == "does_exist":
            hits = hits[:1]
        elif limit is not None:
            hits = hits[:limit]
        if not hits:
            return [format_no_match(fmt)], 1
        return [format_match(h, fmt) for h in hits], 0
    else:
        return None

    if p_var is not None and p_var < nvar:
        return None          # p outer / n inner: not the order the fused pass enumerates
    out = _search.search(target, n_hi, power=power, n_min=n_lo, op=op, all_powers=p_var is not None)
    lines = []
    cap = 1 if quant == "does_exist" else limit
    for n, p in out.matches(None):
        if p_var is not None and p < p_lo:
            continue
        lines.append(format_match({nvar: n, p_var: p} if p_var is not None else {nvar: n}, fmt))
        if cap is not None and len(lines) >= cap:
            break
    if not lines:
        return [format_no_match(fmt)], 1
    return lines, 0
