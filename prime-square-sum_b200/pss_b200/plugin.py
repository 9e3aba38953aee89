"""
plugin.py — pass this file to the reference CLI:  --functions /path/to/pss_b200/plugin.py

The reference loads every public, non-class callable of the file into the `user.` namespace, which wins
unqualified lookup over `pss.` (utils/function_registry.py:134,221-267,499-551), so
`does_exist primesum(n,2) == 666` then evaluates primesum on the B200.  Signatures follow the reference:
primesum(n, power=2) — one required positional argument (tests/test_sequences.py:279-283).
"""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))

from pss_b200 import sequences as _sequences  # noqa: E402
from pss_b200 import sieve as _sieve  # noqa: E402


def primesum(n, power=2):
    """Sum of first n primes raised to power (B200)."""
    return _sequences.primesum(n, power)


def nth_prime(n):
    """The n-th prime, 1-indexed (B200)."""
    return _sieve.nth_prime(int(n))


def prime_count(limit):
    """pi(limit), inclusive (B200)."""
    return _sieve.prime_count(int(limit))
