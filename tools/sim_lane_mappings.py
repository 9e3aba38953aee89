#!/usr/bin/env python
"""numpy model of two lane mappings for the sieve's sparse primes (bank = word & 31, one wavefront per maximum bank
multiplicity of a warp instruction), by prime range:
    lane <-> prime       the kernel's sparse batches: 32 consecutive primes, strand by strand, hit by hit (lanes that have
                         left the tile park on their own bank)
    lane <-> hit         one warp per prime, lanes take 32 consecutive hits of its strand-major hit list
Prints wavefronts per instruction and per 32 cross-offs.  Used for profiles/r02/experiments/README.md."""
import numpy as np

T = 229376
R = [1, 7, 11, 13, 17, 19, 23, 29]


def primes_upto(n):
    s = np.ones(n + 1, bool)
    s[:2] = False
    for i in range(2, int(n ** .5) + 1):
        if s[i]:
            s[i * i::i] = False
    return np.nonzero(s)[0]


def maxload(banks):
    return np.bincount(banks, minlength=32).max()


def lane_per_hit(p, B0):
    r = B0 % p
    hits = np.concatenate([np.arange((p * R[j] // 30 - r) % p, T, p) for j in range(8)])
    wf = sum(maxload((hits[i:i + 32] >> 2) & 31) for i in range(0, len(hits), 32))
    return wf, (len(hits) + 31) // 32, len(hits)


def lane_per_prime(ps, B0):
    wf = n = tot = 0
    for j in range(8):
        xs = [np.arange((p * R[j] // 30 - B0 % p) % p, T, p) for p in ps]
        for k in range(max(len(a) for a in xs)):
            b = [(a[k] >> 2) & 31 if k < len(a) else lane for lane, a in enumerate(xs)]
            wf += maxload(np.array(b))
            n += 1
            tot += sum(1 for a in xs if k < len(a))
    return wf, n, tot


if __name__ == "__main__":
    P = primes_upto(160000)
    rng = np.random.default_rng(1)
    for lo, hi in [(2205, 3600), (3600, 7168), (7168, 14336), (14336, 28672), (28672, 60000), (60000, 151000)]:
        ps = P[(P > lo) & (P <= hi)]
        sel = ps[::max(1, len(ps) // 64)][:64]
        B0 = int(rng.integers(1 << 20, 1 << 29)) & ~15
        a = [lane_per_hit(int(p), B0) for p in sel]
        wf, n, h = (sum(x[i] for x in a) for i in range(3))
        idx = np.searchsorted(P, sel[len(sel) // 2])
        w2, n2, h2 = lane_per_prime([int(x) for x in P[idx:idx + 32]], B0)
        print(f"p in ({lo},{hi}]: lane<->hit {wf / n:.2f} wavefronts/instr, {wf / h * 32:.2f} per 32 cross-offs | "
              f"lane<->prime {w2 / n2:.2f}/instr, {w2 / h2 * 32:.2f} per 32 cross-offs")
