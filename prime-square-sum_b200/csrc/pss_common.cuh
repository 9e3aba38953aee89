// pss_common.cuh — wheel-30 geometry and fixed-width multi-limb arithmetic shared by the kernels
// (device) and by the host-side unit harness tests/host_arith_test.cu (same code, compiled for the host,
// so the index arithmetic can be checked against the oracle without a GPU).
//
// Bitmap layout (the HBM-resident form of a sieved range): byte b covers the integers [30b, 30b+30);
// bit j of that byte stands for 30b + R[j], R = {1,7,11,13,17,19,23,29}.  A 32-bit word covers 120
// integers, a 1 KiB "chunk" (256 words) 30 720 integers.  2, 3 and 5 are not representable and are
// handled by the callers.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PSS_HD __host__ __device__ __forceinline__
#else
#define PSS_HD inline
#endif

typedef uint32_t u32;
typedef uint64_t u64;

namespace pss {

constexpr int kChunkWords = 256;                 // words per count chunk
constexpr int kChunkBytes = kChunkWords * 4;     // 1 KiB = 30 720 integers
constexpr int kMaxChunkPrimes = 4096;            // pi(30720) = 3316 is the densest chunk

// wheel residues and the inverse map residue -> bit index (0xFF when not coprime to 30)
PSS_HD u32 wheel_residue(u32 j) {
    // packed 5-bit fields would also do; a switchless table via constant shifts keeps this HD-safe
    const u64 packed = (1ull) | (7ull << 5) | (11ull << 10) | (13ull << 15) | (17ull << 20) | (19ull << 25) |
                       (23ull << 30) | (29ull << 35);
    return (u32)((packed >> (5 * j)) & 31u);
}
// bit index of residue r (r in {1,7,11,13,17,19,23,29}); for other r the result is meaningless
PSS_HD u32 wheel_bit_of_residue(u32 r) {
    // r:   1 7 11 13 17 19 23 29  ->  0..7 ; (r*... ) closed form: bit = (r * 17) >> 6 works? use table in 4-bit fields
    // index by r/2 (0..14): r/2 = 0,3,5,6,8,9,11,14
    const u64 packed = (0ull << 0) | (1ull << 12) | (2ull << 20) | (3ull << 24) | (4ull << 32) | (5ull << 36) |
                       (6ull << 44) | (7ull << 56);
    return (u32)((packed >> (4 * (r >> 1))) & 15u);
}

// For a base prime p (coprime to 30) and strand j (cofactor m == R[j] mod 30) the multiples p*m live in
// bytes c + p*q (q = 0,1,...) at one fixed bit.  c = floor(p*R[j]/30), bit = bit_of(p*R[j] mod 30).
PSS_HD void strand_geometry(u32 p, u32 j, u32& c, u32& bit) {
    u32 prod = p * wheel_residue(j);          // p < 2^26 -> prod < 2^31
    c = prod / 30u;
    bit = wheel_bit_of_residue(prod - c * 30u);
}

// First byte offset (relative to tile start byte B0) that strand j of prime p hits inside [B0, ...).
// r = B0 mod p.  The cofactor m = 1 (p itself) is never crossed off.
PSS_HD u32 strand_first_offset(u32 p, u32 j, u32 c, u64 B0, u32 r) {
    u32 x = (c >= r) ? (c - r) : (c + p - r);
    if (j == 0 && B0 <= (u64)c) x += p;       // q = 0 would be p*1
    return x;
}

// keep-mask for the byte at absolute index B: bits whose integer lies in [lo, hi)
PSS_HD u32 edge_byte_mask(u64 B, u64 lo, u64 hi) {
    u64 base = B * 30ull;
    if (base >= lo && base + 30ull <= hi) return 0xFFu;
    if (base + 30ull <= lo || base >= hi) return 0u;
    u32 m = 0;
    for (u32 j = 0; j < 8; ++j) {
        u64 v = base + wheel_residue(j);
        if (v >= lo && v < hi) m |= (1u << j);
    }
    return m;
}

// integer represented by bit `bit` (0..31) of the word at absolute byte index Bw (multiple of 4)
PSS_HD u64 word_bit_value(u64 Bw, u32 bit) {
    return (Bw + (bit >> 3)) * 30ull + wheel_residue(bit & 7u);
}

// ------------------------------------------------------------------------------------------------
// Fixed-width little-endian multi-limb unsigned integers (32-bit limbs).  Widths are compile-time so
// every loop unrolls into IMAD.WIDE.U32 / IADD3(.X) carry chains.
// ------------------------------------------------------------------------------------------------

// limbs of p^k and of S_k for p < 2^40, n < 2^36
PSS_HD constexpr int pow_limbs(int k) { return k == 0 ? 1 : (40 * k + 31) / 32; }
PSS_HD constexpr int sum_limbs(int k) { return (40 * k + 36 + 31) / 32; }

template <int L>
struct Big {
    u32 v[L];
};

template <int L>
PSS_HD void big_zero(Big<L>& a) {
#pragma unroll
    for (int i = 0; i < L; ++i) a.v[i] = 0;
}

// a += b (LB <= LA), wraps silently beyond LA limbs (widths are chosen so it cannot)
template <int LA, int LB>
PSS_HD void big_add(Big<LA>& a, const Big<LB>& b) {
    static_assert(LB <= LA, "addend wider than accumulator");
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < LA; ++i) {
        u64 t = (u64)a.v[i] + (i < LB ? (u64)b.v[i] : 0ull) + c;
        a.v[i] = (u32)t;
        c = t >> 32;
    }
}

// r = a * p, p = plo + phi*2^32 (phi < 2^8); result truncated to LR limbs (callers guarantee it fits)
template <int LR, int LA>
PSS_HD void big_mul_p(Big<LR>& r, const Big<LA>& a, u32 plo, u32 phi) {
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < LA; ++i) {
        if (i < LR) {
            u64 t = (u64)a.v[i] * plo + c;
            r.v[i] = (u32)t;
            c = t >> 32;
        }
    }
#pragma unroll
    for (int i = LA; i < LR; ++i) {
        r.v[i] = (u32)c;
        c = 0;
    }
    c = 0;
#pragma unroll
    for (int i = 0; i < LA; ++i) {
        if (i + 1 < LR) {
            u64 t = (u64)a.v[i] * phi + r.v[i + 1] + c;
            r.v[i + 1] = (u32)t;
            c = t >> 32;
        }
    }
    if (LA + 1 < LR) {
        u64 t = (u64)r.v[LA + 1] + c;
        r.v[LA + 1] = (u32)t;
    }
}

// three-way compare of a (L limbs) with target t (L limbs): -1, 0, +1
template <int L>
PSS_HD int big_cmp(const Big<L>& a, const u32* t) {
    int res = 0;
#pragma unroll
    for (int i = 0; i < L; ++i) {  // least significant first; later (more significant) limbs override
        if (a.v[i] != t[i]) res = (a.v[i] < t[i]) ? -1 : 1;
    }
    return res;
}

}  // namespace pss
