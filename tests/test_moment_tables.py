"""CPU checks of the arithmetic behind the byte-table block-sum kernels (csrc/pss_bscan.cuh: k_bitmap_moments_tab and the
compaction table of k_compact_lut): the packed table fields, their head-room under every weight the kernel applies, the
32/64-bit widths of the deferred accumulators, and the whole chain  bytes -> word moments -> thread moments -> S_k  against
direct sums over the primes' offsets.  Pure numpy / Python ints (uint32 arithmetic wraps exactly like the device's)."""
import random

import numpy as np

R = [1, 7, 11, 13, 17, 19, 23, 29]
M32 = 0xFFFFFFFF


def mu(v, i):
    return sum(R[j] ** i for j in range(8) if (v >> j) & 1)


TABLE = [((mu(v, 0) | (mu(v, 3) << 12)) & M32, (mu(v, 1) | (mu(v, 2) << 14)) & M32, mu(v, 4), mu(v, 5)) for v in range(256)]


def word_moments(w):
    """omega_0..5 of one 32-bit bitmap word about the word origin, computed exactly like the kernel (32-bit wrap)."""
    e = [TABLE[(w >> (8 * q)) & 0xFF] for q in range(4)]

    def X(f, p):   # sum_q q^p * field word f  (mod 2^32)
        return sum((q ** p) * e[q][f] for q in range(4)) & M32

    x = [X(0, p) for p in range(6)]
    y = [X(1, p) for p in range(5)]
    z = [X(2, p) for p in range(2)]
    t0 = X(3, 0)
    A = {}
    for p in range(6):
        A[p, 0] = x[p] & 0xFFF
    for p in range(3):
        A[p, 3] = x[p] >> 12
    for p in range(5):
        A[p, 1] = y[p] & 0x3FFF
    for p in range(4):
        A[p, 2] = y[p] >> 14
    A[0, 4], A[1, 4], A[0, 5] = z[0], z[1], t0
    w0 = A[0, 0]
    w1 = A[0, 1] + 30 * A[1, 0]
    w2 = A[0, 2] + 60 * A[1, 1] + 900 * A[2, 0]
    w3 = A[0, 3] + 90 * A[1, 2] + 2700 * A[2, 1] + 27000 * A[3, 0]
    w4 = A[0, 4] + 120 * A[1, 3] + 5400 * A[2, 2] + 108000 * A[3, 1] + 810000 * A[4, 0]
    inner = A[1, 4] + 60 * A[2, 3] + 1800 * A[3, 2] + 27000 * A[4, 1] + 162000 * A[5, 0]
    assert max(w0, w1, w2, w3, w4, inner) <= M32          # every 32-bit intermediate of the kernel really fits
    w5 = A[0, 5] + 150 * inner
    return [w0, w1, w2, w3, w4, w5]


def offsets_of_word(w):
    return [30 * (b >> 3) + R[b & 7] for b in range(32) if (w >> b) & 1]


def test_table_fields_have_headroom_for_every_weight():
    worst = [sum(R[j] ** i for j in range(8)) for i in range(6)]        # byte 0xFF
    sq = [sum(q ** p for q in range(4)) for p in range(6)]
    assert sq[5] * worst[0] < 2 ** 12            # mu0 under q^5
    assert sq[4] * worst[1] < 2 ** 14            # mu1 under q^4
    assert (sq[2] * worst[3]) << 12 < 2 ** 32    # mu3 in the high field, needed up to q^2
    assert (sq[3] * worst[2]) << 14 < 2 ** 32    # mu2 in the high field, needed up to q^3
    assert sq[1] * worst[4] < 2 ** 32 and sq[0] * worst[5] < 2 ** 32


def test_word_moments_match_direct_sums():
    rng = random.Random(7)
    words = [0, M32, 0x80000001, 0x000000FF, 0xFF000000, 0x7EEFDFFE] + [rng.getrandbits(32) for _ in range(4000)]
    words += [rng.getrandbits(32) & rng.getrandbits(32) & rng.getrandbits(32) for _ in range(2000)]    # sparse, like a real bitmap
    for w in words:
        offs = offsets_of_word(w)
        assert word_moments(w) == [sum(o ** j for o in offs) for j in range(6)], hex(w)


def test_deferred_accumulator_widths_and_thread_moments():
    # widths: all-bits-set thread block (24 words), the bound the kernel's 32-bit accumulators rely on
    full = word_moments(M32)
    wide = {(0, 4), (0, 5), (1, 4), (2, 3), (3, 2)}
    for e in range(6):
        for i in range(6 - e):
            bound = sum(w ** e for w in range(24)) * full[i]
            assert (bound < 2 ** 32) != ((e, i) in wide), (e, i, bound)
            assert bound < 2 ** 64
    # thread moments from D(e,i) = sum_w w^e omega_i(w) and S_k by the binomial expansion in B, against direct sums
    rng = random.Random(11)
    binom = lambda n, k: __import__("math").comb(n, k)
    for trial in range(60):
        dens = rng.choice([1, 2, 3])
        words = [M32 if trial == 0 else int.from_bytes(bytes(rng.getrandbits(8) & (rng.getrandbits(8) if dens > 1 else 0xFF) &
                                                                 (rng.getrandbits(8) if dens > 2 else 0xFF) for _ in range(4)), "little")
                 for _ in range(24)]
        om = [word_moments(w) for w in words]
        D = {(e, i): sum((wi ** e) * om[wi][i] for wi in range(24)) for e in range(6) for i in range(6 - e)}
        M = [sum(binom(j, i) * 120 ** (j - i) * D[j - i, i] for i in range(j + 1)) for j in range(6)]
        offs = [120 * wi + o for wi in range(24) for o in offsets_of_word(words[wi])]
        assert M == [sum(o ** j for o in offs) for j in range(6)]
        B = rng.choice([0, 2880 * 7, 30 * (2 ** 33), (2 ** 40 // 2880 - 1) * 2880])
        for k in range(1, 6):
            assert sum(binom(k, j) * B ** (k - j) * M[j] for j in range(k + 1)) == sum((B + o) ** k for o in offs)


def test_compaction_byte_table_lists_residues_in_order():
    for v in range(256):
        res = [R[j] for j in range(8) if (v >> j) & 1]
        lo = hi = 0
        for n, r in enumerate(res):
            if n < 6:
                lo |= r << (5 * n)
            else:
                hi |= r << (5 * (n - 6))
        hi |= len(res) << 16
        assert lo < 2 ** 32 and hi < 2 ** 32
        dec = [(lo >> (5 * k)) & 31 for k in range(min(len(res), 6))] + [(hi >> (5 * (k - 6))) & 31 for k in range(6, len(res))]
        assert dec == res and (hi >> 16) == len(res)
    # (bytes with five to seven primes are common near the origin -- 2.5 % of the first 1e5 bytes -- and vanish quickly as
    # the density falls; the kernel stores the first four with predicated instructions and takes a branch for the rest)


# ---- position-specific tables (k_bitmap_moments2w, k_bitmap_moments_tabw) ------------------------------------------------
def sigma(v, q, i):
    """moment i of byte value v AT byte position q of a word, about the word origin"""
    return sum((R[j] + 30 * q) ** i for j in range(8) if (v >> j) & 1)


def entry128(v, q):   # .x = sigma_4, (.y, .z) = sigma_5 | sigma_3 << 38, .w = sigma_1 | sigma_2 << 12
    yz = sigma(v, q, 5) | (sigma(v, q, 3) << 38)
    return (sigma(v, q, 4), yz & M32, yz >> 32, sigma(v, q, 1) | (sigma(v, q, 2) << 12))


def entry64(v, q):    # sigma_1 | sigma_2 << 12 | sigma_3 << 30
    return sigma(v, q, 1) | (sigma(v, q, 2) << 12) | (sigma(v, q, 3) << 30)


def test_position_table_fields_cannot_overflow():
    """the four entries of a word are ADDED field by field: every field must hold the sum for the all-ones word"""
    tot = [sum(sigma(0xFF, q, i) for q in range(4)) for i in range(6)]
    assert tot[1] < 2 ** 12 and tot[2] < 2 ** 18 and tot[3] < 2 ** 24 and tot[4] < 2 ** 31 and tot[5] < 2 ** 38
    for q in range(4):
        for v in range(256):
            e = entry128(v, q)
            assert all(0 <= c <= M32 for c in e) and (e[2] << 32 | e[1]) < 2 ** 63
            assert entry64(v, q) < 2 ** 54
            assert sigma(v, q, 1) | (sigma(v, q, 2) << 12) < 2 ** 31          # the 32-bit entry of the k = 2 kernel


def test_position_tables_give_the_word_moments():
    rng = random.Random(7)
    words = [0, M32, 1, 1 << 31, 0x80000001] + [rng.getrandbits(32) for _ in range(400)]
    for w in words:
        direct = [sum(o ** i for o in offsets_of_word(w)) for i in range(6)]
        es = [entry128((w >> (8 * q)) & 0xFF, q) for q in range(4)]
        sx = sum(e[0] for e in es) & M32
        syz = sum((e[2] << 32) | e[1] for e in es) & (2 ** 64 - 1)
        sw = sum(e[3] for e in es) & M32
        got = [bin(w).count("1"), sw & 0xFFF, sw >> 12, syz >> 38, sx, syz & (2 ** 38 - 1)]
        assert got == direct, hex(w)
        sn = sum(entry64((w >> (8 * q)) & 0xFF, q) for q in range(4)) & (2 ** 64 - 1)
        assert [sn & 0xFFF, (sn >> 12) & 0x3FFFF, sn >> 30] == direct[1:4], hex(w)
        s2 = sum(sigma((w >> (8 * q)) & 0xFF, q, 1) | (sigma((w >> (8 * q)) & 0xFF, q, 2) << 12) for q in range(4)) & M32
        assert [s2 & 0xFFF, s2 >> 12] == direct[1:3], hex(w)


def test_horner_step_widths_hold_the_running_value():
    """expand_moments: after step j the running value  sum_{i<=j} C(k,i) B^(j-i) M_i  must fit ceil((14 + 40 j) / 32) limbs for
    the largest base (B + o < 2^40) and a thread full of primes at the top of its block"""
    from math import comb
    B = 2 ** 40 - 2880
    offs = list(range(2880 - 768, 2880))          # 768 offsets, all as large as they can be
    M = [sum(o ** i for o in offs) for i in range(6)]
    for k in range(1, 6):
        a = M[0]
        for j in range(1, k + 1):
            a = a * B + comb(k, j) * M[j]
            limbs = min((14 + 40 * j + 31) // 32, (40 * k + 36 + 31) // 32)
            assert a < 2 ** (32 * limbs), (k, j)
        assert a == sum((B + o) ** k for o in offs)
