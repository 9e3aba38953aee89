// pss_trihead.cuh -- word-parallel head pass of the triangular-number predicate for k = 2
//
// `primesum(n,2) == tri(m)` (reference: utils/grammar.py find_matches over the two iterators; utils/number_theory.py:292-354
// qtri) can only hold while S_2(n) <= tri(m_max) < 2^59, i.e. for primes below ~3e6: a few dozen of the densest
// 2880-integer blocks at the very start of the range.  k_bitmap_scan<2, false, kBsTri> walks such a block with ONE lane
// (up to 418 primes, ~270 cycles each: the whole kernel waits for that chain).  Here the first kTriHeadWords bitmap words
// (32 warp tiles = 2 949 120 integers, more than any tri(m_max) < 2^59 can reach) are handled one WORD per thread:
//   k_tri_head_sums   per word: popcount and the sum of the squares of its primes (u64, saturating)
//   k_tri_head_scan   one CTA: exclusive prefixes (index of the word's first prime, S_2 before it) seeded with 2, 3, 5
//   k_tri_head_test   per word: walk its <= 32 candidates from the exact prefix, test every in-window partial sum
// The scan kernel then skips every block that ends inside the head and does not test primes the head has covered
// (BScanParams::tri_head_n), so hits are reported exactly once and S_2(n_max), p_(n_max) still come from the scan kernel.
#pragma once
#include "pss_common.cuh"

namespace pss {

constexpr u32 kTriHeadWords = 32u * 32u * 24u;      // 32 warp tiles of 32 threads x 24 words
constexpr u64 kTriSat = ~0ull;                      // "beyond 2^64": larger than any tri(m_max)

struct TriHeadParams {
    const u32* bitmap;
    u64 n_words;
    u64 B_start;
    u32 n_head_words;        // <= kTriHeadWords, a multiple of the 768 words of a warp tile
    u32 n_small;
    u64 small_primes[3];
    u32* cnt;                // [n_head_words] popcount -> exclusive prefix (index of the prime before the word's first one)
    u64* sum;                // [n_head_words] sum of squares -> exclusive prefix
    u64* head_n;             // out: number of primes the head covers (2, 3, 5 included)
    u64 n_min, n_max;
    u64 tri_m_min, tri_m_max, tri_max_value;
    u64* tri_out;
    u32* tri_count;
    u32 tri_cap;
};

__device__ __forceinline__ u64 tri_sat_add(u64 a, u64 b) {
    const u64 s = a + b;
    return (a == kTriSat || b == kTriSat || s < a) ? kTriSat : s;
}

// value of bit `bit` (ascending value order = ascending bit index) of head word i
__device__ __forceinline__ u64 tri_bit_value(u64 word_base, u32 bit) { return word_base + 30u * (bit >> 3) + wheel_residue(bit & 7u); }

__device__ __forceinline__ u64 tri_square_or_sat(u64 p) { return (p >> 32) ? kTriSat : p * p; }

// S == tri(m) with m in [m_min, m_max]?  8 S + 1 < 2^62 must be a perfect square r^2 (then m = (r - 1) / 2); see
// BsCmpFn::post in pss_bscan.cuh for the argument behind the three candidates
__device__ __forceinline__ void tri_test(const TriHeadParams& P, u64 S, u64 n) {
    const u64 d = 8ull * S + 1ull;
    const u32 r0 = (u32)(u64)sqrt((double)d);
    const u64 e = d - (u64)r0 * r0, two_r = 2ull * r0;
    const bool at = (e == 0ull), above = (e == two_r + 1ull), below_ = (e + two_r == 1ull);
    if (at || above || below_) {
        const u64 r = (u64)r0 + (above ? 1ull : 0ull) - (below_ ? 1ull : 0ull);
        const u64 m = (r - 1) >> 1;
        if (m >= P.tri_m_min && m <= P.tri_m_max) {
            const u32 slot = atomicAdd(P.tri_count, 1u);
            if (slot < P.tri_cap) {
                P.tri_out[2 * slot] = n;
                P.tri_out[2 * slot + 1] = m;
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_tri_head_sums(const TriHeadParams P) {
    const u32 i = blockIdx.x * 256u + threadIdx.x;
    if (i >= P.n_head_words) return;
    u32 w = (i < P.n_words) ? __ldg(P.bitmap + i) : 0u;
    const u64 base = (P.B_start + 4ull * i) * 30ull;
    const u32 c = __popc(w);
    u64 s = 0;
    while (w) {
        const u32 bit = (u32)__ffs((int)w) - 1u;
        w &= w - 1u;
        s = tri_sat_add(s, tri_square_or_sat(tri_bit_value(base, bit)));
    }
    P.cnt[i] = c;
    P.sum[i] = s;
}

// one CTA of 1024 threads, every thread owns a contiguous run of words: run totals -> block-wide exclusive scan -> in-place
// exclusive prefixes, seeded with the primes below 7 that the bitmap cannot represent
__global__ void __launch_bounds__(1024) k_tri_head_scan(const TriHeadParams P) {
    __shared__ u32 s_c[32];
    __shared__ u64 s_s[32];
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u32 per = (P.n_head_words + 1023u) / 1024u;
    const u32 i0 = min(P.n_head_words, tid * per), i1 = min(P.n_head_words, i0 + per);
    u32 c = 0;
    u64 s = 0;
#pragma unroll 8
    for (u32 i = i0; i < i1; ++i) {       // (independent loads: unrolled so that they are in flight together)
        c += P.cnt[i];
        s = tri_sat_add(s, P.sum[i]);
    }
    // inclusive warp scan, then the warp totals by warp 0
    u32 ci = c;
    u64 si = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 cu = __shfl_up_sync(0xFFFFFFFFu, ci, d);
        const u64 su = __shfl_up_sync(0xFFFFFFFFu, si, d);
        if (lane >= (u32)d) {
            ci += cu;
            si = tri_sat_add(si, su);
        }
    }
    // exclusive inside the warp: the inclusive value of the lane below
    const u32 ce_l = __shfl_up_sync(0xFFFFFFFFu, ci, 1);
    const u64 se_l = __shfl_up_sync(0xFFFFFFFFu, si, 1);
    const u32 ex_c = lane ? ce_l : 0u;
    const u64 ex_s = lane ? se_l : 0ull;
    if (lane == 31) {
        s_c[warp] = ci;
        s_s[warp] = si;
    }
    __syncthreads();
    if (warp == 0) {
        u32 wci = s_c[lane];
        u64 wsi = s_s[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 cu = __shfl_up_sync(0xFFFFFFFFu, wci, d);
            const u64 su = __shfl_up_sync(0xFFFFFFFFu, wsi, d);
            if (lane >= (u32)d) {
                wci += cu;
                wsi = tri_sat_add(wsi, su);
            }
        }
        // exclusive over warps: shift down by one
        const u32 ce = __shfl_up_sync(0xFFFFFFFFu, wci, 1);
        const u64 se = __shfl_up_sync(0xFFFFFFFFu, wsi, 1);
        __syncwarp();
        s_c[lane] = lane ? ce : 0u;
        s_s[lane] = lane ? se : 0ull;
        if (lane == 31) *P.head_n = (u64)wci + P.n_small;
    }
    __syncthreads();
    u64 s0 = 0;
    for (u32 k = 0; k < P.n_small; ++k) s0 += P.small_primes[k] * P.small_primes[k];
    // exclusive prefix of this thread's run = seed (2, 3, 5) + earlier warps + earlier lanes of this warp
    u32 run_c = P.n_small + s_c[warp] + ex_c;
    u64 run_s = tri_sat_add(tri_sat_add(s0, s_s[warp]), ex_s);
#pragma unroll 8
    for (u32 i = i0; i < i1; ++i) {
        const u32 ci_ = P.cnt[i];
        const u64 si_ = P.sum[i];
        P.cnt[i] = run_c;
        P.sum[i] = run_s;
        run_c += ci_;
        run_s = tri_sat_add(run_s, si_);
    }
}

__global__ void __launch_bounds__(256) k_tri_head_test(const TriHeadParams P) {
    const u32 i = blockIdx.x * 256u + threadIdx.x;
    if (i >= P.n_head_words) return;
    if (i == 0) {                       // 2, 3, 5 come first
        u64 S = 0;
        for (u32 k = 0; k < P.n_small; ++k) {
            S += P.small_primes[k] * P.small_primes[k];
            const u64 n = (u64)k + 1;
            if (n >= P.n_min && n <= P.n_max && S <= P.tri_max_value) tri_test(P, S, n);
        }
    }
    u64 n = P.cnt[i];                   // primes before this word
    u64 S = P.sum[i];
    if (S == kTriSat || S > P.tri_max_value || n >= P.n_max) return;
    u32 w = (i < P.n_words) ? __ldg(P.bitmap + i) : 0u;
    const u64 base = (P.B_start + 4ull * i) * 30ull;
    while (w) {
        const u32 bit = (u32)__ffs((int)w) - 1u;
        w &= w - 1u;
        S = tri_sat_add(S, tri_square_or_sat(tri_bit_value(base, bit)));
        n += 1;
        if (n > P.n_max || S == kTriSat || S > P.tri_max_value) break;
        if (n >= P.n_min) tri_test(P, S, n);
    }
}

}  // namespace pss
