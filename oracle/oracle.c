/*
 * oracle.c — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker or the timed CPU baseline — never on the product path
 * (prime-square-sum_b200/ never imports it; the product fails loudly without its CUDA library).
 *
 * The reference (djdarcy/Prime-Square-Sum v0.8.3) is pure Python; this file restates, in plain C,
 * the algorithm of these reference functions (semantics only — licence CC BY-NC-ND, nothing copied):
 *   orc_sieve_range      utils/sieve.py:257-327  _segmented_sieve: base primes <= sqrt(limit) by a basic
 *                        sieve (:224-254), then fixed-size byte segments, first multiple
 *                        max(ceil(low/p)*p, p*p) (:310-312), stop at p*p >= high (:305-307);
 *                        generalised to [lo, hi) as generate_primes_range does by filtering (:535-538).
 *   orc_power_sum        utils/gpu.py:244-247 / utils/sequences.py:75   sum(int(p) ** power), exact.
 *   orc_scan_compare     utils/sum_cache.py:116-141 add_prime running sums, compared per n as
 *                        find_matches does: utils/grammar.py:686-701 (Comparison), :770-787 (_compare),
 *                        does_exist first-match :987-988; n is the 1-based prime index
 *                        (utils/sequences.py:30-75).
 *   orc_prefix_sums      every partial sum primesum(n,k), n = n_first .. n_first+count-1.
 *
 * Parity is PINNED: oracle/gen_golden.py runs the reference's own code (imported read-only from
 * /root/reference) and this restatement on the same inputs and commits the vectors under tests/golden/;
 * tests/test_oracle.py re-checks this file against those vectors and the reference's own KATs.
 *
 * Big integers: little-endian arrays of ORC_LIMBS 32-bit limbs (512 bit).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_LIMBS 16

typedef struct orc_big { uint32_t v[ORC_LIMBS]; } orc_big;

static void big_zero(orc_big* a) { memset(a->v, 0, sizeof a->v); }
static void big_from_u64(orc_big* a, uint64_t x) { big_zero(a); a->v[0] = (uint32_t)x; a->v[1] = (uint32_t)(x >> 32); }
static void big_add(orc_big* a, const orc_big* b) {
    uint64_t c = 0;
    for (int i = 0; i < ORC_LIMBS; ++i) { uint64_t t = (uint64_t)a->v[i] + b->v[i] + c; a->v[i] = (uint32_t)t; c = t >> 32; }
}
static void big_sub(orc_big* r, const orc_big* a, const orc_big* b) {
    uint64_t borrow = 0;
    for (int i = 0; i < ORC_LIMBS; ++i) {
        uint64_t bi = (uint64_t)b->v[i] + borrow, ai = a->v[i];
        r->v[i] = (uint32_t)(ai - bi);
        borrow = ai < bi;
    }
}
/* a *= m (m < 2^64) */
static void big_mul_u64(orc_big* a, uint64_t m) {
    uint32_t m0 = (uint32_t)m, m1 = (uint32_t)(m >> 32);
    uint32_t r[ORC_LIMBS + 2];
    memset(r, 0, sizeof r);
    uint64_t c = 0;
    for (int i = 0; i < ORC_LIMBS; ++i) { uint64_t t = (uint64_t)a->v[i] * m0 + c; r[i] = (uint32_t)t; c = t >> 32; }
    c = 0;
    for (int i = 0; i + 1 < ORC_LIMBS; ++i) { uint64_t t = (uint64_t)a->v[i] * m1 + r[i + 1] + c; r[i + 1] = (uint32_t)t; c = t >> 32; }
    memcpy(a->v, r, sizeof a->v);
}
static int big_cmp(const orc_big* a, const orc_big* b) {
    for (int i = ORC_LIMBS - 1; i >= 0; --i)
        if (a->v[i] != b->v[i]) return a->v[i] < b->v[i] ? -1 : 1;
    return 0;
}
/* int(p) ** power */
static void big_pow_u64(orc_big* r, uint64_t p, unsigned power) {
    big_from_u64(r, 1);
    for (unsigned i = 0; i < power; ++i) big_mul_u64(r, p);
}

/* ---- sieve ------------------------------------------------------------------------------------- */

/* _basic_sieve(limit): primes < limit, one byte per integer (utils/sieve.py:224-254) */
static uint64_t basic_sieve(uint64_t limit, uint64_t** out) {
    *out = NULL;
    if (limit < 2) return 0;
    uint8_t* is_prime = (uint8_t*)malloc(limit);
    if (!is_prime) return 0;
    memset(is_prime, 1, limit);
    is_prime[0] = 0; is_prime[1] = 0;
    for (uint64_t i = 2; i * i < limit; ++i)
        if (is_prime[i])
            for (uint64_t m = i * i; m < limit; m += i) is_prime[m] = 0;
    uint64_t n = 0;
    for (uint64_t i = 0; i < limit; ++i) n += is_prime[i];
    uint64_t* p = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
    uint64_t k = 0;
    for (uint64_t i = 0; i < limit; ++i) if (is_prime[i]) p[k++] = i;
    free(is_prime);
    *out = p;
    return n;
}

/* primes in [lo, hi) by the segmented algorithm; writes at most cap primes to out (out may be NULL to
 * count only); returns the number of primes in the range. */
uint64_t orc_sieve_range(uint64_t lo, uint64_t hi, uint64_t segment_size, uint64_t* out, uint64_t cap) {
    if (hi < 2 || lo >= hi) return 0;
    if (lo < 2) lo = 2;
    if (segment_size == 0) segment_size = 1u << 18;   /* reference default is 10 M elements; any size gives the same primes */
    uint64_t sqrt_limit = (uint64_t)sqrt((double)hi) + 1;
    while (sqrt_limit * sqrt_limit < hi) ++sqrt_limit;
    uint64_t* small = NULL;
    uint64_t n_small = basic_sieve(sqrt_limit + 1, &small);
    uint8_t* seg = (uint8_t*)malloc(segment_size);
    uint64_t count = 0;
    for (uint64_t low = lo; low < hi; low += segment_size) {
        uint64_t high = low + segment_size < hi ? low + segment_size : hi;
        uint64_t len = high - low;
        memset(seg, 1, len);
        for (uint64_t s = 0; s < n_small; ++s) {
            uint64_t p = small[s];
            if (p * p >= high) break;
            uint64_t start = ((low + p - 1) / p) * p;
            if (start < p * p) start = p * p;
            for (uint64_t m = start - low; m < len; m += p) seg[m] = 0;
        }
        for (uint64_t i = 0; i < len; ++i)
            if (seg[i]) {
                if (out && count < cap) out[count] = low + i;
                ++count;
            }
    }
    free(seg);
    free(small);
    return count;
}

/* ---- exact power sums ---------------------------------------------------------------------------- */

/* sums_out[j] (+)= sum_i values[i]^(power_j); powers[] lists n_powers exponents; carry_in may be NULL */
void orc_power_sums(const uint64_t* values, uint64_t count, const uint32_t* powers, uint32_t n_powers,
                    const uint32_t* carry_in, uint32_t* sums_out) {
    for (uint32_t j = 0; j < n_powers; ++j) {
        orc_big s, t;
        if (carry_in) memcpy(s.v, carry_in + (size_t)j * ORC_LIMBS, sizeof s.v); else big_zero(&s);
        for (uint64_t i = 0; i < count; ++i) {
            big_pow_u64(&t, values[i], powers[j]);
            big_add(&s, &t);
        }
        memcpy(sums_out + (size_t)j * ORC_LIMBS, s.v, sizeof s.v);
    }
}

typedef struct orc_scan_result {
    uint64_t match_count;       /* #n in [n_min,n_max] with S(n) <cmp> target */
    uint64_t first_match_n;     /* 0 if none */
    uint64_t last_match_n;      /* 0 if none */
    uint64_t cross_n;           /* first scanned n with S(n) >= target, 0 if none */
    uint64_t cross_prime;
    uint32_t sum_at_cross[ORC_LIMBS];
    uint32_t sum_before_cross[ORC_LIMBS];
    uint32_t final_sum[ORC_LIMBS];
} orc_scan_result;

/* cmp: 0 ==, 1 !=, 2 <, 3 <=, 4 >, 5 >=   (utils/grammar.py:770-787) */
static int cmp_holds(int c, uint32_t op) {
    switch (op) {
        case 0: return c == 0;
        case 1: return c != 0;
        case 2: return c < 0;
        case 3: return c <= 0;
        case 4: return c > 0;
        default: return c >= 0;
    }
}

/* running sum over values[0..count) whose 1-based prime indices are n_first.., seeded with carry_in,
 * every partial sum compared with target */
void orc_scan_compare(const uint64_t* values, uint64_t count, uint64_t n_first, uint32_t power,
                      const uint32_t* carry_in, const uint32_t* target, uint32_t cmp,
                      uint64_t n_min, uint64_t n_max, orc_scan_result* res) {
    orc_big s, t, tgt, before;
    memset(res, 0, sizeof *res);
    if (carry_in) memcpy(s.v, carry_in, sizeof s.v); else big_zero(&s);
    memcpy(tgt.v, target, sizeof tgt.v);
    int prevc = big_cmp(&s, &tgt);
    for (uint64_t i = 0; i < count; ++i) {
        const uint64_t n = n_first + i;
        big_pow_u64(&t, values[i], power);
        before = s;
        big_add(&s, &t);
        int c = big_cmp(&s, &tgt);
        if (n >= n_min && n <= n_max && cmp_holds(c, cmp)) {
            if (!res->match_count) res->first_match_n = n;
            res->last_match_n = n;
            res->match_count++;
        }
        if (prevc < 0 && c >= 0) {
            res->cross_n = n;
            res->cross_prime = values[i];
            memcpy(res->sum_at_cross, s.v, sizeof s.v);
            memcpy(res->sum_before_cross, before.v, sizeof before.v);
        }
        prevc = c;
    }
    memcpy(res->final_sum, s.v, sizeof s.v);
}

/* out[i] = carry + sum_{j<=i} values[j]^power  (ORC_LIMBS limbs each) */
void orc_prefix_sums(const uint64_t* values, uint64_t count, uint32_t power, const uint32_t* carry_in, uint32_t* out) {
    orc_big s, t;
    if (carry_in) memcpy(s.v, carry_in, sizeof s.v); else big_zero(&s);
    for (uint64_t i = 0; i < count; ++i) {
        big_pow_u64(&t, values[i], power);
        big_add(&s, &t);
        memcpy(out + i * ORC_LIMBS, s.v, sizeof s.v);
    }
}

/* streaming form of the whole path for one value block (the unit the CPU baseline parallelises over):
 * sieve [lo,hi), accumulate sum p^power for each requested power, return the count.  Equivalent to
 * orc_sieve_range followed by orc_power_sums without materialising the primes. */
uint64_t orc_block_sieve_sum(uint64_t lo, uint64_t hi, uint64_t segment_size, const uint32_t* powers, uint32_t n_powers,
                             uint32_t* sums_out, uint64_t* last_prime) {
    orc_big s[8], t;
    if (n_powers > 8) n_powers = 8;
    for (uint32_t j = 0; j < n_powers; ++j) big_zero(&s[j]);
    uint64_t count = 0, last = 0;
    if (!(hi < 2 || lo >= hi)) {
        if (lo < 2) lo = 2;
        if (segment_size == 0) segment_size = 1u << 18;
        uint64_t sqrt_limit = (uint64_t)sqrt((double)hi) + 1;
        while (sqrt_limit * sqrt_limit < hi) ++sqrt_limit;
        uint64_t* small = NULL;
        uint64_t n_small = basic_sieve(sqrt_limit + 1, &small);
        uint8_t* seg = (uint8_t*)malloc(segment_size);
        for (uint64_t low = lo; low < hi; low += segment_size) {
            uint64_t high = low + segment_size < hi ? low + segment_size : hi;
            uint64_t len = high - low;
            memset(seg, 1, len);
            for (uint64_t k = 0; k < n_small; ++k) {
                uint64_t p = small[k];
                if (p * p >= high) break;
                uint64_t start = ((low + p - 1) / p) * p;
                if (start < p * p) start = p * p;
                for (uint64_t m = start - low; m < len; m += p) seg[m] = 0;
            }
            for (uint64_t i = 0; i < len; ++i)
                if (seg[i]) {
                    const uint64_t p = low + i;
                    for (uint32_t j = 0; j < n_powers; ++j) {
                        big_pow_u64(&t, p, powers[j]);
                        big_add(&s[j], &t);
                    }
                    last = p;
                    ++count;
                }
        }
        free(seg);
        free(small);
    }
    for (uint32_t j = 0; j < n_powers; ++j) memcpy(sums_out + (size_t)j * ORC_LIMBS, s[j].v, sizeof s[j].v);
    if (last_prime) *last_prime = last;
    return count;
}

/* ------------------------------------------------------------------------------------------------------------------
 * CPU stand-in for the reference's preferred third-party sieve (primesieve: PyPI, version unpinned by the reference, not
 * installed and not installable offline -- SURVEY 8(c)).  NOT a restatement of reference code: a conventional cache-blocked,
 * bit-packed wheel-30 segmented sieve (the published primesieve design without its bucket lists) plus 256-bit accumulators,
 * so that bench.py can report a CPU number that is not handicapped by the reference's byte-per-number numpy formulation.
 * Checked against orc_block_sieve_sum in tests/test_oracle.py.  Same unit of work: one value block [lo, hi).
 * ------------------------------------------------------------------------------------------------------------------ */
static const uint8_t W30_RES[8] = {1, 7, 11, 13, 17, 19, 23, 29};
static int w30_bit_of(uint32_t r) {
    for (int j = 0; j < 8; ++j) if (W30_RES[j] == r) return j;
    return -1;
}
typedef struct { uint64_t v[4]; } acc256;
static inline void acc256_add_pow(acc256* a, uint64_t p, unsigned power) {
    /* a += p^power, p < 2^40, power <= 5: p^power < 2^200 */
    uint64_t w[4] = {1, 0, 0, 0};
    if (power <= 3 && p < (1ull << 40)) {                         /* p^power < 2^120: one 128-bit product chain */
        unsigned __int128 x = 1;
        for (unsigned e = 0; e < power; ++e) x *= p;
        unsigned __int128 c = (unsigned __int128)a->v[0] + (uint64_t)x;
        a->v[0] = (uint64_t)c;
        c = (c >> 64) + a->v[1] + (uint64_t)(x >> 64);
        a->v[1] = (uint64_t)c;
        c = (c >> 64) + a->v[2];
        a->v[2] = (uint64_t)c;
        a->v[3] += (uint64_t)(c >> 64);
        return;
    }
    for (unsigned e = 0; e < power; ++e) {
        unsigned __int128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (unsigned __int128)w[i] * p;
            w[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (unsigned __int128)a->v[i] + w[i];
        a->v[i] = (uint64_t)c;
        c >>= 64;
    }
}
uint64_t orc_fast_block_sieve_sum(uint64_t lo, uint64_t hi, const uint32_t* powers, uint32_t n_powers, uint32_t* sums_out,
                                  uint64_t* last_prime) {
    acc256 s[8];
    memset(s, 0, sizeof s);
    if (n_powers > 8) n_powers = 8;
    uint64_t count = 0, last = 0;
    if (hi >= 2 && lo < hi) {
        for (uint64_t q = 2; q <= 5; q += (q == 2) ? 1 : 2)       /* 2, 3, 5 are not on the wheel */
            if (q >= lo && q < hi) {
                for (uint32_t j = 0; j < n_powers; ++j) acc256_add_pow(&s[j], q, powers[j]);
                ++count; last = q;
            }
        uint64_t sqrt_limit = (uint64_t)sqrt((double)hi) + 1;
        while (sqrt_limit * sqrt_limit < hi) ++sqrt_limit;
        uint64_t* base = NULL;
        const uint64_t n_base = basic_sieve(sqrt_limit + 1, &base);
        enum { SEG_BYTES = 256 * 1024 };                           /* 7.9 M integers per segment: L2 resident */
        uint8_t* seg = (uint8_t*)malloc(SEG_BYTES);
        const uint64_t B_first = lo / 30, B_end = (hi + 29) / 30;  /* bitmap bytes [B_first, B_end) */
        for (uint64_t B0 = B_first; B0 < B_end; B0 += SEG_BYTES) {
            const uint64_t nb = (B_end - B0 < SEG_BYTES) ? (B_end - B0) : SEG_BYTES;
            memset(seg, 0xFF, nb);
            const uint64_t seg_hi = (B0 + nb) * 30;
            for (uint64_t k = 0; k < n_base; ++k) {
                const uint64_t p = base[k];
                if (p < 7) continue;
                if (p * p >= seg_hi) break;
                for (int j = 0; j < 8; ++j) {                       /* cofactor residue R[j]: bytes c + p*q, one fixed bit */
                    const uint64_t prod = p * W30_RES[j];
                    const uint64_t c = prod / 30;
                    const uint8_t mask = (uint8_t)~(1u << w30_bit_of((uint32_t)(prod % 30)));
                    /* first byte index >= B0 on the strand, cofactor >= p (multiples below p*p are already crossed) */
                    uint64_t q = (B0 > c) ? (B0 - c + p - 1) / p : 0;
                    const uint64_t q_min = (p - W30_RES[j] + 29) / 30;   /* 30 q + R[j] >= p */
                    if (q < q_min) q = q_min;
                    if (j == 0 && q == 0) q = 1;                    /* cofactor 1 is the prime itself */
                    for (uint64_t b = c + p * q; b < B0 + nb; b += p) seg[b - B0] &= mask;
                }
            }
            for (uint64_t i = 0; i < nb; ++i) {
                unsigned bits = seg[i];
                while (bits) {
                    const int j = __builtin_ctz(bits);
                    bits &= bits - 1;
                    const uint64_t v = (B0 + i) * 30 + W30_RES[j];
                    if (v < lo || v >= hi || v < 7) continue;
                    for (uint32_t jj = 0; jj < n_powers; ++jj) acc256_add_pow(&s[jj], v, powers[jj]);
                    ++count; last = v;
                }
            }
        }
        free(seg);
        free(base);
    }
    for (uint32_t j = 0; j < n_powers; ++j) {
        uint32_t* o = sums_out + (size_t)j * ORC_LIMBS;
        memset(o, 0, ORC_LIMBS * sizeof(uint32_t));
        for (int i = 0; i < 4 && 2 * i + 1 < ORC_LIMBS; ++i) { o[2 * i] = (uint32_t)s[j].v[i]; o[2 * i + 1] = (uint32_t)(s[j].v[i] >> 32); }
    }
    if (last_prime) *last_prime = last;
    return count;
}

int orc_limbs(void) { return ORC_LIMBS; }

/* unused-function guard for big_sub when compiled with -Wall */
void orc_big_sub(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    orc_big ra, aa, bb;
    memcpy(aa.v, a, sizeof aa.v); memcpy(bb.v, b, sizeof bb.v);
    big_sub(&ra, &aa, &bb);
    memcpy(r, ra.v, sizeof ra.v);
}
