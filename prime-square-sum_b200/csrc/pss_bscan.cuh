// pss_bscan.cuh — the fused search path: K2 + K3 reading the sieve's bit-packed output directly.
//
// The wheel-30 bitmap IS the compact prime buffer (0.76 B / prime at 2.28e10 instead of 8 B / prime), so the
// query path never materialises u64 primes: each thread owns kWordsPerThread consecutive bitmap words
// (blocked order, staged through padded shared memory from coalesced uint4 loads), decodes its primes with
// ffs, raises them to the k-th power with the same IMAD carry chains as pss_scan.cuh and
//   kBsReduce   per-thread and per-warp (count, S_k limbs) records + per-tile deferred-carry columns (u64 atomics)
//   k_scan_columns  exclusive u64 prefix of every column (count, each 32-bit limb as a deferred-carry
//                   64-bit lane) across tiles — the grid level of the warp -> block -> grid scan
//   kBsCompare  re-walks the tile from its exact prefix (carry-in + normalised column prefixes + block
//               scan), forms EVERY partial sum S_k(n) and compares it with the target limbs.
// Both passes are warp-autonomous (warp-local staging, no __syncthreads) and no CTA ever waits for another
// CTA (three launches instead of a look-back chain), so the multi-GPU
// "phase A" (slice count + slice sums) is simply the column totals of the first two launches.
//
// Replaces the same reference code as pss_scan.cuh (utils/sequences.py:75, utils/gpu.py:169-247,
// utils/sum_cache.py:116-141, utils/grammar.py:686-701,770-787).
#pragma once
#include "pss_common.cuh"
#include "pss_scan.cuh"
#include "pss_sieve.cuh"

namespace pss {

constexpr int kBsThreads = 256;
constexpr int kWordsPerThread = 24;                          // 96 B of bitmap = 2880 integers per thread
constexpr int kBsTileWords = kBsThreads * kWordsPerThread;   // 6144 words = 24 KiB = 737 280 integers
enum BsMode { kBsReduce = 0, kBsCompare = 1, kBsTri = 2 };

struct BScanParams {
    const u32* bitmap;       // word 0 = absolute byte index B_start
    u64 n_words;
    u64 B_start;
    u32 n_tiles;
    u32 n_small;             // 2, 3, 5 inside [lo, hi): handled by thread 0 of tile 0 (not representable)
    u64 small_primes[3];
    // tile records, column major: col c of tile t at cols[c * col_stride + t]; col 0 = count, 1.. = limbs
    u64* cols;               // deferred-carry 64-bit lanes, accumulated with atomics (zeroed before the reduce pass)
    const u64* col_prefix;   // exclusive prefixes, same layout, (n_tiles + 1) entries per column
    u64 col_stride;
    u32* warp_rec;           // exact per-warp (count, limbs), col c of global warp g at warp_rec[c * warp_stride + g]
    u64 warp_stride;
    // per-thread records of the reduce pass (count, sums), column major over global thread ids, so that
    // the compare pass needs no second power pass to rebuild its block scan
    u32* thread_rec;         // col c of global thread g at thread_rec[c * thread_stride + g]
    u64 thread_stride;
    // compare mode
    u64 n_first;             // global 1-based index of the slice's first prime
    const u64* n_first_dev;  // if set: read it from device memory instead (written by k_peer_gather of the same stream)
    u64 n_min, n_max;
    u32 cmp_mask, target_over_mask;
    u32 reduce_chain;        // reduce pass: 1 = multi-limb power chain per prime, 0 = small moments + one expansion per thread (k <= 5)
    u32 exhaustive;          // 1: walk every prime and form every S_k(n); 0: prune warp tiles / threads whose sum interval
                             // (S_before, S_after] cannot contain the target (identical results: S is strictly increasing)
    u32 target[PSS_LIMBS];
    const u32* carry_in;     // NP * PSS_LIMBS limbs (device)
    pss_power_result* results;
    u64* last_prime_out;     // written by the thread that owns n_max
    u32* prefix_out;         // unused here (Pass2Fn interface)
    u32* error_flag;
    // triangular-number predicate (kBsTri): S_k(n) == tri(m), m in [tri_m_min, tri_m_max]
    u64 tri_m_min, tri_m_max, tri_max_value;
    u64* tri_out;            // (n, m) pairs
    u32* tri_count;
    u32 tri_cap;
    const u64* tri_head_n;   // null, or: the first *tri_head_n primes were tested by the word-parallel head pass (pss_trihead.cuh)
};

// Words are staged BIT-REVERSED: the lowest prime of a word is its most significant set bit m (one FLO).
// Original bit b = 31 - m, so the integer offset inside the word is 30*(b/8) + R[b%8]
//   = (R[7 - m%8] + 90) - 30*(m/8);  the bracketed byte table sits in two registers and is read with one
// PRMT (selector nibbles 1..3 = sign replication of a byte < 128, i.e. zero: no masking needed).
__device__ __forceinline__ u32 msb_value_offset(u32 m) {
    // R reversed + 90: j = 0..7 -> 29,23,19,17,13,11,7,1 (+90) = 119,113,109,107,103,101,97,91
    // (inline PTX: the __byte_perm intrinsic masks the selector to 3 bits per nibble and would drop the
    // sign-replication flags)
    u32 r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(0x6B6D7177u), "r"(0x5B616567u), "r"((m & 7u) | 0x8880u));
    return r - 30u * (m >> 3);
}

__device__ __forceinline__ u32 msb_index(u32 w) {   // FLO: index of the most significant set bit (w != 0)
    u32 r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(w));
    return r;
}

// Decode loop: one prime per iteration for every lane that still has one (lanes own equal word counts, so
// the trip counts differ only by the local density fluctuation).  Refill of an exhausted word is branch-free:
// an unconditional shared-memory load + select; the thread's padded row ends with a zero sentinel word, so no
// bounds select is needed.  The body is a plain if-region: the warp reconverges every iteration.
// f(plo, phi) -> false stops this lane.  Returns the last prime handed to f.
template <class F>
__device__ __forceinline__ u64 walk_words(const u32* my_words, u64 value_base, F& f) {
    // 32-bit bookkeeping throughout: word index instead of a generic pointer, and the 64-bit value base split
    // into two opaque registers (otherwise ptxas rematerialises "byte index * 30" inside the loop)
    u32 base_lo = (u32)value_base, base_hi = (u32)(value_base >> 32);
    asm volatile("" : "+r"(base_lo), "+r"(base_hi));
    // refill of an exhausted word as four predicated instructions (the C++ select form compiles to thirteen)
    u32 addr = (u32)__cvta_generic_to_shared(my_words), w = 0, woff = 0u - 120u;
    const u32 end = addr + 4u * (u32)kWordsPerThread;      // the sentinel slot of this thread's padded row
    u32 plo = 0, phi = 0;
    bool alive = true;
    while (alive) {
        asm volatile("{\n\t.reg .pred q;\n\tsetp.eq.u32 q, %0, 0;\n\t@q ld.shared.u32 %0, [%1];\n\t@q add.u32 %1, %1, 4;\n\t@q add.u32 %2, %2, 120;\n\t}"
                     : "+r"(w), "+r"(addr), "+r"(woff) :: "memory");
        if (w != 0) {
            const u32 m = msb_index(w);
            w ^= (1u << m);
            const u32 off = woff + msb_value_offset(m);
            asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, 0;" : "=r"(plo), "=r"(phi) : "r"(base_lo), "r"(off), "r"(base_hi));
            alive = f(plo, phi);
        } else {
            alive = (addr <= end);   // empty word (rare): try the next one; past the sentinel: done
        }
    }
    return ((u64)phi << 32) | plo;
}

template <int K, bool MULTI>
struct BsAccFn {
    u32* acc;
    __device__ __forceinline__ bool operator()(u32 plo, u32 phi) {
        AccFn<K, MULTI> fn{acc};
        power_chain2<K, MULTI>(plo, phi, fn);
        return true;
    }
};

// ---- block sums for k <= 5 from small moments -----------------------------------------------------------------
// A thread's primes lie in [B, B + 2880).  With o = p - B (12 bits) the thread accumulates M_j = sum o^j, j <= K, in
// machine words (M_1: 32 bit, M_2..M_4: 64 bit, M_5: 96 bit; 768 set bits * 2880^5 < 2^68) -- about 12 integer
// instructions per prime for K = 5 instead of the ~100 of the multi-limb chain p, p^2, .., p^5 -- and expands once:
//     S_k = sum_p (B + o)^k = sum_j C(k,j) B^(k-j) M_j      evaluated by Horner in B with multi-limb arithmetic.
template <int K>
struct BsMomFn {
    u32 base_lo;
    u32 m1;
    u64 m2, m3, m4, m5lo;
    u32 m5hi;
    __device__ __forceinline__ bool operator()(u32 plo, u32 /*phi*/) {
        const u32 o = plo - base_lo;              // < 2880 (exact also across a 2^32 boundary)
        m1 += o;
        if constexpr (K >= 2) {
            const u32 o2 = o * o;                 // < 2^23
            m2 += (u64)o2;
            if constexpr (K >= 3) m3 += (u64)o2 * o;
            if constexpr (K >= 4) {
                const u64 o4 = (u64)o2 * o2;      // < 2^46
                m4 += o4;
                if constexpr (K >= 5) {
                    const u64 t = o4 * o;         // < 2^58
                    const u64 s = m5lo + t;
                    m5hi += (s < t) ? 1u : 0u;
                    m5lo = s;
                }
            }
        }
        return true;
    }
};

PSS_HD constexpr u32 binom(int n, int k) {
    u32 r = 1;
    for (int i = 1; i <= k; ++i) r = r * (u32)(n - k + i) / (u32)i;
    return r;
}

// out[0..W) = sum_j C(k,j) B^(k-j) M_j  (Horner in B; W = sum_limbs(k) limbs, nothing can overflow them).
// The running value after step j, a_j = sum_{i<=j} C(k,i) B^(j-i) M_i, is at most 10 * 768 * 2^(40 j) (every binomial ratio
// C(k,i) / C(j,i) is <= 10 for k <= 5, a thread holds <= 768 primes, all below 2^40), so step j only multiplies the
// ceil((14 + 40 (j-1)) / 32) limbs that can be non-zero and produces ceil((14 + 40 j) / 32): 2 + 4 + 6 + 10 + 12 limb
// products for k = 5 instead of 2 + 8 + 14 + 16 + 16.
template <int KK>
PSS_HD constexpr int horner_limbs(int j) {
    return ((14 + 40 * j + 31) / 32 < sum_limbs(KK)) ? (14 + 40 * j + 31) / 32 : sum_limbs(KK);
}

template <int J, int KK, int K>
__device__ __forceinline__ void horner_steps(Big<sum_limbs(KK)>& a, const BsMomFn<K>& f, u32 blo, u32 bhi) {
    if constexpr (J <= KK) {
        constexpr int LI = horner_limbs<KK>(J - 1), LO = horner_limbs<KK>(J);
        Big<LI> in;
#pragma unroll
        for (int i = 0; i < LI; ++i) in.v[i] = a.v[i];
        Big<LO> nx;
        big_mul_p<LO, LI>(nx, in, blo, bhi);
        // t = C(KK, J) * M_J  (at most 72 bits)
        u64 lo = 0;
        u32 hi = 0;
        if (J == 1) lo = f.m1;
        if (J == 2) lo = f.m2;
        if (J == 3) lo = f.m3;
        if (J == 4) lo = f.m4;
        if (J == 5) { lo = f.m5lo; hi = f.m5hi; }
        const u32 c = binom(KK, J);
        const u64 p0 = (lo & 0xFFFFFFFFull) * c;
        const u64 p1 = (lo >> 32) * c + (p0 >> 32);
        const u64 p2 = (u64)hi * c + (p1 >> 32);
        const u32 t[3] = {(u32)p0, (u32)p1, (u32)p2};
        u64 cy = 0;
#pragma unroll
        for (int i = 0; i < LO; ++i) {
            const u64 v = (u64)nx.v[i] + (i < 3 ? (u64)t[i] : 0ull) + cy;
            a.v[i] = (u32)v;
            cy = v >> 32;
        }
        horner_steps<J + 1, KK, K>(a, f, blo, bhi);
    }
}

template <int KK, int K>
__device__ __forceinline__ void expand_moments(u32* out, const BsMomFn<K>& f, u32 m0, u32 blo, u32 bhi) {
    constexpr int W = sum_limbs(KK);
    Big<W> a;
    big_zero(a);
    a.v[0] = m0;
    horner_steps<1, KK, K>(a, f, blo, bhi);
#pragma unroll
    for (int i = 0; i < W; ++i) out[i] = a.v[i];
}

template <int J, int K, bool MULTI>
__device__ __forceinline__ void expand_all(u32* acc, const BsMomFn<K>& f, u32 m0, u32 blo, u32 bhi) {
    using Lay = Layout<K, MULTI>;
    if constexpr (J < Lay::NP) {
        constexpr int KK = MULTI ? J + 1 : K;
        expand_moments<KK, K>(acc + Lay::offset(J), f, m0, blo, bhi);
        expand_all<J + 1, K, MULTI>(acc, f, m0, blo, bhi);
    }
}

// most-significant-limb-first three-way compare (the outcome is almost always decided by the top limb)
template <int W>
__device__ __forceinline__ int limbs_cmp_top(const u32* a, const u32* t) {
#pragma unroll
    for (int i = W - 1; i >= 0; --i)
        if (a[i] != t[i]) return (a[i] < t[i]) ? -1 : 1;
    return 0;
}

// Compare-pass functor.  Every partial sum S_k(n) is formed in `run` and compared with the target; the
// per-thread bookkeeping uses that S is strictly increasing: inside one thread the outcomes are
// "<" ... "<" ["=="] ">" ... ">", so (number below, index of the equal one, number processed) describe every
// operator's matches exactly.  quota = primes left up to n_max, skip = primes still below n_min.
template <int K, bool MULTI, int PMODE, bool HAS_SKIP>
struct BsCmpFn {
    using Lay = Layout<K, MULTI>;
    static constexpr int NP = Lay::NP;
    u32* run;
    const BScanParams* P;
    u64 tile_n0;       // global index of the tile's first prime
    u32 li;            // tile-local index of the next prime
    u32 quota, skip;
    u32 n_in;          // in-window primes processed
    u32 n_lt[NP];      // ... of which S < target
    u32 eq_li[NP];     // tile-local index with S == target (in window), 0xFFFFFFFF if none
    bool below[NP];    // previous partial sum < target
    u32 cur_lo, cur_hi;
    u64 tri_head_n = 0;   // triangular predicate: primes with index <= tri_head_n were tested by the head pass

    template <int E>
    __device__ __forceinline__ void apply(const Big<pow_limbs(E)>& pe) {
        constexpr int J = MULTI ? E - 1 : 0;
        acc_add<Lay::offset(J), Lay::width(J), pow_limbs(E)>(run, pe.v);
        post<E>(&pe);
    }
    __device__ __forceinline__ void apply_square(u32 plo, u32 phi) {
        square_acc4(run, plo, phi);
        post<2>(nullptr);
    }
    template <int E>
    __device__ __forceinline__ void post(const Big<pow_limbs(E)>* pe) {
        constexpr int J = MULTI ? E - 1 : 0;
        constexpr int OFF = Lay::offset(J), W = Lay::width(J);
        const bool inw = HAS_SKIP ? (skip == 0) : true;
        if constexpr (PMODE == kModeTri) {
            // S == tri(m), m in [tri_m_min, tri_m_max]  <=>  8S+1 = s^2 and m = (s-1)/2 in range
            // (utils/number_theory.py:292-354 qtri).  tri(m_max) < 2^59: only sums that fit 64 bits can match.
            bool small = true;
#pragma unroll
            for (int i = 2; i < W; ++i) small = small && (run[OFF + i] == 0u);
            const u64 S = ((u64)run[OFF + 1] << 32) | run[OFF];
            if (inw && small && S <= P->tri_max_value && tile_n0 + li > tri_head_n) {
                // d < 2^62, so r0 = trunc(sqrt((double) d)) < 2^32 is floor(sqrt d) - 1, floor or floor + 1 (the conversion of d
                // moves the root by < 2^-22, the rounding of the root by less): d is a square iff it is r0^2, (r0 + 1)^2 or
                // (r0 - 1)^2 -- one 32 x 32 multiply and three compares instead of correction loops of 64-bit multiplies
                const u64 d = 8ull * S + 1ull;
                const u32 r0 = (u32)(u64)sqrt((double)d);
                const u64 e = d - (u64)r0 * r0, two_r = 2ull * r0;
                const bool at = (e == 0ull), above = (e == two_r + 1ull), below_ = (e + two_r == 1ull);
                if (at || above || below_) {
                    const u64 r = (u64)r0 + (above ? 1ull : 0ull) - (below_ ? 1ull : 0ull);
                    const u64 m = (r - 1) >> 1;
                    if (m >= P->tri_m_min && m <= P->tri_m_max) {
                        const u32 slot = atomicAdd(P->tri_count, 1u);
                        if (slot < P->tri_cap) {
                            P->tri_out[2 * slot] = tile_n0 + li;
                            P->tri_out[2 * slot + 1] = m;
                        }
                    }
                }
            }
        } else {
            const int c = ((P->target_over_mask >> J) & 1u) ? -1 : limbs_cmp_top<W>(run + OFF, P->target);
            if (inw) {
                n_lt[J] += (c < 0) ? 1u : 0u;
                if (c == 0) eq_li[J] = li;
            }
            if (below[J] && c >= 0) {   // the unique crossing of this power: S(n-1) < T <= S(n)
                pss_power_result* r = P->results + J;
                r->cross_n = tile_n0 + li;
                r->cross_prime = ((u64)cur_hi << 32) | cur_lo;
                Big<pow_limbs(E)> pw;
                if (pe) {
                    pw = *pe;
                } else {   // fused-square path: rebuild p^2 for the (once per query) bracket record
                    Big<2> p1;
                    p1.v[0] = cur_lo;
                    p1.v[1] = cur_hi;
                    big_mul_p<pow_limbs(2), 2>(reinterpret_cast<Big<pow_limbs(2)>&>(pw), p1, cur_lo, cur_hi);
                }
                u32 before[W];
                limbs_sub<W, pow_limbs(E)>(before, run + OFF, pw.v);
#pragma unroll
                for (int i = 0; i < PSS_LIMBS; ++i) {
                    r->sum_at_cross[i] = (i < W) ? run[OFF + i] : 0u;
                    r->sum_before_cross[i] = (i < W) ? before[i] : 0u;
                }
            }
            below[J] = (c < 0);
        }
    }
    __device__ __forceinline__ bool operator()(u32 plo, u32 phi) {
        cur_lo = plo;
        cur_hi = phi;
        power_chain2<K, MULTI>(plo, phi, *this);
        if (HAS_SKIP) {
            if (skip == 0) ++n_in; else --skip;
        } else {
            ++n_in;
        }
        ++li;
        return --quota != 0;
    }
};

// Tail of the reduce pass, shared by the walking kernel and the moment kernel: per-thread record, exact warp totals
// (butterfly of carry-propagating adds) -> warp record, deferred-carry u64 atomics -> scan-tile columns.
template <int K, bool MULTI>
__device__ __forceinline__ void bs_reduce_tail(const BScanParams& P, u32* acc, u32 cnt, u64 wt, u32 tile, u32 lane) {
    using Lay = Layout<K, MULTI>;
    constexpr int TOTAL = Lay::TOTAL;
    const u64 gthread = wt * 32 + lane;
    P.thread_rec[gthread] = cnt;
#pragma unroll
    for (int i = 0; i < TOTAL; ++i) P.thread_rec[(u64)(1 + i) * P.thread_stride + gthread] = acc[i];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        OpShflXorAdd op{acc, d};
        for_each_power<0, Lay>(op);
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if (lane == 0) {
        P.warp_rec[wt] = cnt;
        atomicAdd(reinterpret_cast<unsigned long long*>(P.cols + tile), (unsigned long long)cnt);
#pragma unroll
        for (int i = 0; i < TOTAL; ++i) {
            P.warp_rec[(u64)(1 + i) * P.warp_stride + wt] = acc[i];
            atomicAdd(reinterpret_cast<unsigned long long*>(P.cols + (u64)(1 + i) * P.col_stride + tile), (unsigned long long)acc[i]);
        }
    }
}

// Warp-autonomous kernel: a warp owns one "warp tile" = 32 threads x 24 words (768 words, 92 160 integers);
// 8 consecutive warp tiles form a scan tile (the granularity of the column prefix).  No __syncthreads.
// resident CTAs per SM the register allocation is bounded for (issue-bound kernel: occupancy hides the
// dependent IMAD/carry chains)
template <int K, bool MULTI>
constexpr int bs_min_blocks() {
    // (k <= 2, one power: three CTAs as before the warp-wide localisation code was added -- it needs 91 registers if left alone,
    //  which costs the exhaustive compare its third CTA: 4.81 -> 5.13 ms)
    return MULTI ? (K >= 4 ? 2 : 3) : (K >= 6 ? 3 : (K <= 2 ? 3 : 1));
}

template <int K, bool MULTI, int MODE>
__global__ void __launch_bounds__(kBsThreads, bs_min_blocks<K, MULTI>()) k_bitmap_scan(const __grid_constant__ BScanParams P) {
    using Lay = Layout<K, MULTI>;
    constexpr int TOTAL = Lay::TOTAL, NP = Lay::NP, T = kBsThreads, NWARP = T / 32;
    constexpr int WPT = kWordsPerThread, PAD = WPT + 1, WARP_WORDS = 32 * WPT;

    __shared__ u32 s_words[T * PAD];

    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    u32* my_warp_words = s_words + warp * 32 * PAD;
    const u32* my_words = my_warp_words + lane * PAD;
    my_warp_words[lane * PAD + WPT] = 0u;        // sentinel word of this thread's row (never overwritten)
    const u64 n_warp_tiles = (u64)P.n_tiles * NWARP;
    const u64 warp_stride = (u64)gridDim.x * NWARP;
    const u64 n_first = (MODE != kBsReduce && P.n_first_dev) ? *P.n_first_dev : P.n_first;
    const u64 tri_head_n = (MODE == kBsTri && P.tri_head_n) ? *P.tri_head_n : 0ull;

    // normalised prefix of scan tile t (carry-in + deferred-carry column prefixes), uniform across the warp
    auto tile_prefix = [&](u32 t, u32* out) {
#pragma unroll
        for (int jj = 0; jj < NP; ++jj) {
            u64 c = 0;
#pragma unroll
            for (int i = 0; i < PSS_LIMBS; ++i) {
                if (i < Lay::width(jj)) {
                    const int col = 1 + Lay::offset(jj) + i;
                    const u64 v = P.col_prefix[(u64)col * P.col_stride + t];
                    const u64 lo = (v & 0xFFFFFFFFull) + (c & 0xFFFFFFFFull) + P.carry_in[jj * PSS_LIMBS + i];
                    out[Lay::offset(jj) + i] = (u32)lo;
                    c = (v >> 32) + (c >> 32) + (lo >> 32);
                }
            }
        }
    };
    // matches of a run of consecutive primes: in-window tile-local indices [a, a + n_in), of which n_lt below the
    // target and eq_li the index of an exact hit; outcomes are  < ... < [==] > ... >  (S strictly increasing)
    auto emit = [&](u32 a, u32 n_in, const u32* n_lt, const u32* eq_li, u64 tile_n0) {
#pragma unroll
        for (int jj = 0; jj < NP; ++jj) {
            const u32 eq = (eq_li[jj] != 0xFFFFFFFFu) ? 1u : 0u;
            const u32 lt = n_lt[jj], gt = n_in - lt - eq;
            u32 mc = 0, mf = 0xFFFFFFFFu, ml = 0;   // count, first (tile-local), last + 1
            if (n_in) {
                const u32 m = P.cmp_mask;
                // contiguous runs: [a, a+lt) below, [a+lt, a+lt+eq) equal, [a+lt+eq, a+n_in) above
                if ((m & 1u) && lt) { mc += lt; mf = min(mf, a); ml = max(ml, a + lt); }
                if ((m & 2u) && eq) { mc += 1; mf = min(mf, a + lt); ml = max(ml, a + lt + 1); }
                if ((m & 4u) && gt) { mc += gt; mf = min(mf, a + lt + eq); ml = max(ml, a + n_in); }
            }
            if (__any_sync(0xFFFFFFFFu, mc != 0)) {
                const u32 c = __reduce_add_sync(0xFFFFFFFFu, mc);
                const u32 f = __reduce_min_sync(0xFFFFFFFFu, mf);
                const u32 l = __reduce_max_sync(0xFFFFFFFFu, ml);
                if (lane == 0) {
                    pss_power_result* r = P.results + jj;
                    atomicAdd(reinterpret_cast<unsigned long long*>(&r->match_count), (unsigned long long)c);
                    atomicMin(reinterpret_cast<unsigned long long*>(&r->first_match_n), (unsigned long long)(tile_n0 + f));
                    atomicMax(reinterpret_cast<unsigned long long*>(&r->last_match_n), (unsigned long long)(tile_n0 + l - 1));
                }
            }
        }
    };

    // work units: reduce pass = warp tiles; compare pass = scan tiles (8 warp tiles), pruned as a whole when possible; the
    // triangular predicate = warp tiles again (only the first few blocks of a range can hold a sum <= tri(m_max), and they are
    // the densest: one warp per scan tile walked two warp tiles of ~400 primes per lane one after the other)
    constexpr bool kUnitIsWarpTile = (MODE == kBsReduce) || (MODE == kBsTri);
    const u64 n_units = kUnitIsWarpTile ? n_warp_tiles : (u64)P.n_tiles;
    for (u64 unit = (u64)blockIdx.x * NWARP + warp; unit < n_units; unit += warp_stride) {
        u32 tile, wit_lo, wit_hi;
        if constexpr (MODE == kBsReduce) {
            tile = (u32)(unit / NWARP);
            wit_lo = (u32)(unit % NWARP);
            wit_hi = wit_lo + 1u;
        } else {
            if constexpr (MODE == kBsTri) {
                tile = (u32)(unit / NWARP);
                wit_lo = (u32)(unit % NWARP);
                wit_hi = wit_lo + 1u;
            } else {
                tile = (u32)unit;
                wit_lo = 0u;
                wit_hi = NWARP;
            }
            if (n_first + P.col_prefix[tile] > P.n_max) break;   // prefixes are monotone in tile
            // ---- interval pruning, scan-tile level (737 280 integers): same argument as below on the tile's exact
            //      prefix / end sums; a pruned tile costs one warp ~100 instructions and touches no record
            if (!P.exhaustive) {
                const u64 t_n0 = n_first + P.col_prefix[tile];
                const u64 tcnt64 = P.col_prefix[tile + 1] - P.col_prefix[tile];
                const u32 tcnt = (u32)tcnt64;
                const u32 t_lmin = (P.n_min > t_n0) ? (u32)min(P.n_min - t_n0, (u64)0xFFFFFFFFull) : 0u;
                const u32 t_lmax = (u32)min(P.n_max - t_n0, (u64)0xFFFFFFFEull);
                const u32 tquota = min(tcnt, t_lmax + 1u);
                const u32 tskip = min(t_lmin, tquota);
                const bool t_nmax = tquota > 0 && (t_n0 + tquota - 1 == P.n_max);
                // (a window edge inside the tile does not matter: the in-window primes [tskip, tquota) are one contiguous run)
                bool simple = !t_nmax;
                if (simple) {
                    u32 tpre[TOTAL];
                    tile_prefix(tile, tpre);
                    if constexpr (MODE == kBsCompare) {
                        u32 tend[TOTAL];
                        tile_prefix(tile + 1, tend);
                        int c0[NP], c1[NP];
                        OpCmpInit o0{tpre, P.target, P.target_over_mask, c0};
                        for_each_power<0, Lay>(o0);
                        OpCmpInit o1{tend, P.target, P.target_over_mask, c1};
                        for_each_power<0, Lay>(o1);
#pragma unroll
                        for (int jj = 0; jj < NP; ++jj) simple = simple && !(c0[jj] < 0 && c1[jj] >= 0);
                        if (simple) {
                            u32 t_nlt[NP], t_eq[NP];
                            const u32 t_in = (lane == 0) ? tquota - tskip : 0u;
#pragma unroll
                            for (int jj = 0; jj < NP; ++jj) {
                                t_nlt[jj] = (c0[jj] < 0) ? t_in : 0u;
                                t_eq[jj] = 0xFFFFFFFFu;
                            }
                            emit(tskip, t_in, t_nlt, t_eq, t_n0);
                            continue;
                        }
                    } else {
                        bool beyond = false;
#pragma unroll
                        for (int i = 2; i < Lay::width(0); ++i) beyond = beyond || (tpre[i] != 0u);
                        beyond = beyond || ((((u64)tpre[1] << 32) | tpre[0]) > P.tri_max_value);
                        if (beyond) continue;
                    }
                }
            }
        }
      for (u32 wit = wit_lo; wit < wit_hi; ++wit) {          // warp index inside the scan tile
        const u64 wt = (u64)tile * NWARP + wit;
        const u64 word0 = wt * WARP_WORDS;
        // ---- warp-local staging: coalesced uint4 loads -> padded blocked layout, bit-reversed
        auto stage = [&]() {
            const uint4* g4 = reinterpret_cast<const uint4*>(P.bitmap + word0);
            __syncwarp();
#pragma unroll
            for (int r = 0; r < WPT / 4; ++r) {
                const u32 q = r * 32 + lane;          // uint4 index inside the warp tile
                uint4 v = make_uint4(0, 0, 0, 0);
                if (word0 + 4ull * q + 4 <= P.n_words) v = __ldg(g4 + q);
                const u32 wl = 4u * q;                // first word (warp-tile local); 4 | WPT so one owner thread
                const u32 dst = (wl / WPT) * PAD + (wl % WPT);
                my_warp_words[dst] = __brev(v.x); my_warp_words[dst + 1] = __brev(v.y);
                my_warp_words[dst + 2] = __brev(v.z); my_warp_words[dst + 3] = __brev(v.w);
            }
            __syncwarp();
        };

        const u64 value_base = (P.B_start + 4ull * (word0 + (u64)lane * WPT)) * 30ull;
        const bool has_small = (wt == 0 && lane == 0 && P.n_small > 0);
        const u64 gthread = wt * 32 + lane;

        u32 acc[TOTAL];
        u32 cnt = 0;
        if constexpr (MODE == kBsReduce) {
            stage();
            // ---- thread-local count and sums
#pragma unroll
            for (int i = 0; i < TOTAL; ++i) acc[i] = 0;
#pragma unroll
            for (int i = 0; i < WPT; ++i) cnt += __popc(my_words[i]);
            BsAccFn<K, MULTI> fn{acc};
            bool chain = true;
            if constexpr (K <= 5) chain = P.reduce_chain != 0;
            if (chain) {
                walk_words(my_words, value_base, fn);               // per prime: p, p^2, .., p^K in multi-limb arithmetic
            } else if constexpr (K <= 5) {
                BsMomFn<K> mf{(u32)value_base, 0u, 0ull, 0ull, 0ull, 0ull, 0u};
                walk_words(my_words, value_base, mf);               // per prime: small moments of p - B
                expand_all<0, K, MULTI>(acc, mf, cnt, (u32)value_base, (u32)(value_base >> 32));
            }
            if (has_small) {
                for (u32 i = 0; i < P.n_small; ++i) fn((u32)P.small_primes[i], 0u);
                cnt += P.n_small;
            }
            bs_reduce_tail<K, MULTI>(P, acc, cnt, wt, tile, lane);
        } else {
            // ---- warp prefix (uniform) = carry-in + normalised deferred-carry tile prefix + earlier warps of the scan tile
            u32 wpre[TOTAL];
            u32 wli0 = 0;                // tile-local index of the warp's first prime
            {
                tile_prefix(tile, wpre);
                // earlier warps of this scan tile: lane l < wit fetches warp record l, butterfly sum
                u32 v[TOTAL];
                u32 c = 0;
                const u64 wrec = (u64)tile * NWARP + lane;
#pragma unroll
                for (int i = 0; i < TOTAL; ++i) v[i] = (lane < wit) ? __ldg(P.warp_rec + (u64)(1 + i) * P.warp_stride + wrec) : 0u;
                if (lane < wit) c = __ldg(P.warp_rec + wrec);
#pragma unroll
                for (int d = 4; d >= 1; d >>= 1) {   // records live in lanes 0..7
                    OpShflXorAdd op{v, d};
                    for_each_power<0, Lay>(op);
                }
                c += __shfl_xor_sync(0xFFFFFFFFu, c, 4);
                c += __shfl_xor_sync(0xFFFFFFFFu, c, 2);
                c += __shfl_xor_sync(0xFFFFFFFFu, c, 1);
#pragma unroll
                for (int i = 0; i < TOTAL; ++i) v[i] = __shfl_sync(0xFFFFFFFFu, v[i], 0);
                wli0 = __shfl_sync(0xFFFFFFFFu, c, 0);
                OpAdd op{wpre, v};
                for_each_power<0, Lay>(op);
            }
            const u64 tile_n0 = n_first + P.col_prefix[tile];
            // window bounds in tile-local indices (a tile holds < 2^20 primes, so saturating to u32 is exact)
            const u32 lmin = (P.n_min > tile_n0) ? (u32)min(P.n_min - tile_n0, (u64)0xFFFFFFFFull) : 0u;
            const u32 lmax = (u32)min(P.n_max - tile_n0, (u64)0xFFFFFFFEull);   // n_max >= tile_n0 (checked above)
            constexpr int PMODE = (MODE == kBsTri) ? kModeTri : kModeCompare;

            // per-lane outcome of this warp tile: in-window indices [a, a + n_in), of which n_lt below the target and
            // eq_li the (tile-local) index of an exact hit.  S is strictly increasing, so inside any run of consecutive
            // primes the outcomes are  < ... < [==] > ... >.
            u32 a = 0, n_in = 0, n_lt[NP], eq_li[NP];
#pragma unroll
            for (int jj = 0; jj < NP; ++jj) {
                n_lt[jj] = 0;
                eq_li[jj] = 0xFFFFFFFFu;
            }

            // ---- interval pruning, warp level.  If the warp tile lies entirely inside (or outside) the window, does
            //      not hold n_max, and the target is not inside (S_before, S_after] of the warp tile for any power, every
            //      partial sum of the tile compares like its prefix: no prime of it has to be touched.
            bool warp_done = false;
            if (!P.exhaustive) {
                const u32 wcnt = __ldg(P.warp_rec + wt);
                const u32 wquota = (wli0 > lmax) ? 0u : min(wcnt, lmax - wli0 + 1u);
                const u32 wskip = (lmin > wli0) ? min(lmin - wli0, wquota) : 0u;
                const bool holds_nmax = wquota > 0 && (tile_n0 + wli0 + wquota - 1 == P.n_max);
                bool simple = !holds_nmax;
                if constexpr (MODE == kBsCompare) {
                    if (simple) {
                        u32 wend[TOTAL];
#pragma unroll
                        for (int i = 0; i < TOTAL; ++i) wend[i] = wpre[i];
                        u32 wsum[TOTAL];
#pragma unroll
                        for (int i = 0; i < TOTAL; ++i) wsum[i] = __ldg(P.warp_rec + (u64)(1 + i) * P.warp_stride + wt);
                        OpAdd op{wend, wsum};
                        for_each_power<0, Lay>(op);
                        int c0[NP], c1[NP];
                        OpCmpInit o0{wpre, P.target, P.target_over_mask, c0};
                        for_each_power<0, Lay>(o0);
                        OpCmpInit o1{wend, P.target, P.target_over_mask, c1};
                        for_each_power<0, Lay>(o1);
#pragma unroll
                        for (int jj = 0; jj < NP; ++jj) simple = simple && !(c0[jj] < 0 && c1[jj] >= 0);
                        if (simple) {
                            warp_done = true;
                            if (lane == 0) {
                                a = wli0 + wskip;
                                n_in = wquota - wskip;
#pragma unroll
                                for (int jj = 0; jj < NP; ++jj) n_lt[jj] = (c0[jj] < 0) ? n_in : 0u;
                            }
                        }
                    }
                } else {
                    // triangular predicate: only sums <= tri(m_max) (< 2^59) can match; S grows monotonically
                    bool beyond = false;
#pragma unroll
                    for (int i = 2; i < Lay::width(0); ++i) beyond = beyond || (wpre[i] != 0u);
                    beyond = beyond || ((((u64)wpre[1] << 32) | wpre[0]) > P.tri_max_value);
                    // ... and a warp tile whose last prime lies inside the head pass has been tested there
                    if (simple && (beyond || tile_n0 + wli0 + wcnt <= tri_head_n + 1ull)) warp_done = true;
                }
            }

            if (!warp_done) {
                // ---- thread level: exclusive prefix of this thread = warp prefix + earlier lanes
                cnt = __ldg(P.thread_rec + gthread);
                u32 own[TOTAL];
#pragma unroll
                for (int i = 0; i < TOTAL; ++i) own[i] = acc[i] = __ldg(P.thread_rec + (u64)(1 + i) * P.thread_stride + gthread);
                u32 cinc = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    OpShflUpAdd op{acc, d, lane >= (u32)d};
                    for_each_power<0, Lay>(op);
                    const u32 o = __shfl_up_sync(0xFFFFFFFFu, cinc, d);
                    if (lane >= (u32)d) cinc += o;
                }
                u32 run[TOTAL];
#pragma unroll
                for (int i = 0; i < TOTAL; ++i) {
                    const u32 up = __shfl_up_sync(0xFFFFFFFFu, acc[i], 1);
                    run[i] = lane ? up : 0u;
                }
                {
                    OpAdd op{run, wpre};
                    for_each_power<0, Lay>(op);
                }
                const u32 li0 = wli0 + cinc - cnt;
                // quota = primes this thread may process (<= n_max), skip = primes of this thread still below n_min
                const u32 have = cnt;
                const u32 quota = (li0 > lmax) ? 0u : min(have, lmax - li0 + 1u);
                const u32 skip0 = (lmin > li0) ? min(lmin - li0, quota) : 0u;
                const bool owns_nmax = quota > 0 && (tile_n0 + li0 + quota - 1 == P.n_max);
                int c0[NP];
#pragma unroll
                for (int jj = 0; jj < NP; ++jj) c0[jj] = 0;
                if constexpr (MODE == kBsCompare) {
                    OpCmpInit op{run, P.target, P.target_over_mask, c0};
                    for_each_power<0, Lay>(op);
                }
                // ---- interval pruning, thread level (same argument on the thread's 2880 integers)
                bool walk = quota > 0;
                bool coop = false;     // this thread holds n_max and nothing else forces a per-prime walk: the warp finishes it together
                if (!P.exhaustive && walk) {
                    if constexpr (MODE == kBsCompare) {
                        u32 end[TOTAL];
#pragma unroll
                        for (int i = 0; i < TOTAL; ++i) end[i] = run[i];
                        OpAdd op{end, own};
                        for_each_power<0, Lay>(op);
                        int c1[NP];
                        OpCmpInit o1{end, P.target, P.target_over_mask, c1};
                        for_each_power<0, Lay>(o1);
                        bool bracket = false;
#pragma unroll
                        for (int jj = 0; jj < NP; ++jj) bracket = bracket || (c0[jj] < 0 && c1[jj] >= 0);
                        // (for the thread that holds n_max the interval (S_before, S_before + whole block] is a superset of
                        //  (S_before, S(n_max)]: no bracket there means no bracket up to n_max)
                        walk = bracket;
                        coop = owns_nmax && !bracket;
                    } else {
                        // triangular predicate: a block needs its primes tested unless its prefix is already beyond
                        // tri(m_max) or the head pass covered it; the n_max block then only owes S(n_max) and p_(n_max),
                        // which the warp computes together (below) instead of one lane walking ~135 .. 418 primes
                        bool beyond = false;
#pragma unroll
                        for (int i = 2; i < Lay::width(0); ++i) beyond = beyond || (run[i] != 0u);
                        beyond = beyond || ((((u64)run[1] << 32) | run[0]) > P.tri_max_value);
                        const bool covered = beyond || (tile_n0 + li0 + have <= tri_head_n + 1ull);
                        walk = !covered;
                        coop = owns_nmax && covered;
                    }
                }
                a = li0 + skip0;
                u64 last_p = 0;
                if (__any_sync(0xFFFFFFFFu, walk || coop)) stage();   // the bitmap is only read where a walk is needed
                bool localised = false;
                if constexpr (MODE == kBsCompare && !MULTI) {
                    // ---- the block that brackets the target, localised by the whole warp.  Walking it alone is a serial chain
                    //      of up to 418 compared primes (tools/hit_probe.py: +0.09 ms for a hit in the densest block).  Lane 0
                    //      takes the primes below 7 (first block of a range only), lane 1 + i word i of the owner's block:
                    //      count prefix -> which primes lie at or below n_max -> their power sums -> multi-limb scan ->
                    //      the ONE chunk whose (S_before, S_after] holds the target is walked by its lane (<= 32 primes);
                    //      every in-window prime before it compares below, every one after it above.
                    const u32 loc_mask = P.exhaustive ? 0u : __ballot_sync(0xFFFFFFFFu, walk);
                    if (loc_mask) {
                        const int owner = __ffs(loc_mask) - 1;
                        const u32 o_quota = __shfl_sync(0xFFFFFFFFu, quota, owner);
                        const u32 o_skip = __shfl_sync(0xFFFFFFFFu, skip0, owner);
                        const u32 o_li0 = __shfl_sync(0xFFFFFFFFu, li0, owner);
                        const u32 vb_lo = __shfl_sync(0xFFFFFFFFu, (u32)value_base, owner);
                        const u32 vb_hi = __shfl_sync(0xFFFFFFFFu, (u32)(value_base >> 32), owner);
                        const bool o_small = __shfl_sync(0xFFFFFFFFu, has_small ? 1u : 0u, owner) != 0u;
                        const u64 vb = ((u64)vb_hi << 32) | vb_lo;
                        u32 o_run[TOTAL];
#pragma unroll
                        for (int i = 0; i < TOTAL; ++i) o_run[i] = __shfl_sync(0xFFFFFFFFu, run[i], owner);
                        u32 w = 0, c = 0;
                        if (lane == 0) c = o_small ? P.n_small : 0u;
                        else if ((int)lane <= WPT) { w = my_warp_words[owner * PAD + (int)lane - 1]; c = __popc(w); }
                        u32 incl_c = c;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const u32 up = __shfl_up_sync(0xFFFFFFFFu, incl_c, d);
                            if (lane >= (u32)d) incl_c += up;
                        }
                        const u32 pre = incl_c - c;
                        const u32 take = (o_quota > pre) ? min(c, o_quota - pre) : 0u;
                        u32 part[TOTAL];
#pragma unroll
                        for (int i = 0; i < TOTAL; ++i) part[i] = 0;
                        {
                            BsAccFn<K, MULTI> pf{part};
                            if (lane == 0) {
                                for (u32 t = 0; t < take; ++t) pf((u32)P.small_primes[t], 0u);
                            } else {
                                u32 ww = w;
                                for (u32 t = 0; t < take; ++t) {
                                    const u32 m = msb_index(ww);
                                    ww ^= (1u << m);
                                    const u64 pv = vb + 120u * (lane - 1u) + msb_value_offset(m);
                                    pf((u32)pv, (u32)(pv >> 32));
                                }
                            }
                        }
                        u32 sa[TOTAL];                       // inclusive scan of the chunk sums -> S after the chunk
#pragma unroll
                        for (int i = 0; i < TOTAL; ++i) sa[i] = part[i];
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            OpShflUpAdd op{sa, d, lane >= (u32)d};
                            for_each_power<0, Lay>(op);
                        }
                        u32 sb[TOTAL];                       // S before the chunk
#pragma unroll
                        for (int i = 0; i < TOTAL; ++i) {
                            const u32 up = __shfl_up_sync(0xFFFFFFFFu, sa[i], 1);
                            sb[i] = lane ? up : 0u;
                        }
                        {
                            OpAdd o1{sa, o_run};
                            for_each_power<0, Lay>(o1);
                            OpAdd o2{sb, o_run};
                            for_each_power<0, Lay>(o2);
                        }
                        int cb[NP], ca[NP];
                        {
                            OpCmpInit o1{sb, P.target, P.target_over_mask, cb};
                            for_each_power<0, Lay>(o1);
                            OpCmpInit o2{sa, P.target, P.target_over_mask, ca};
                            for_each_power<0, Lay>(o2);
                        }
                        const bool mine = take > 0 && cb[0] < 0 && ca[0] >= 0;
                        const u32 bmask = __ballot_sync(0xFFFFFFFFu, mine);
                        u32 r_nlt = 0, r_eq = 0xFFFFFFFFu;
                        if (mine) {
                            const u32 c_skip = (o_skip > pre) ? min(o_skip - pre, take) : 0u;
                            auto chunk_walk = [&](auto& fn) {
                                fn.n_lt[0] = 0;
                                fn.eq_li[0] = 0xFFFFFFFFu;
                                fn.below[0] = true;
                                if (lane == 0) {
                                    bool go = true;
                                    for (u32 t = 0; t < take && go; ++t) go = fn((u32)P.small_primes[t], 0u);
                                } else {
                                    u32 ww = w;
                                    bool go = true;
                                    while (ww && go) {
                                        const u32 m = msb_index(ww);
                                        ww ^= (1u << m);
                                        const u64 pv = vb + 120u * (lane - 1u) + msb_value_offset(m);
                                        go = fn((u32)pv, (u32)(pv >> 32));
                                    }
                                }
                                r_nlt = fn.n_lt[0];
                                r_eq = fn.eq_li[0];
                            };
                            if (c_skip != 0) {
                                BsCmpFn<K, MULTI, kModeCompare, true> fn{sb, &P, tile_n0, o_li0 + pre, take, c_skip, 0, {}, {}, {}, 0, 0};
                                chunk_walk(fn);
                            } else {
                                BsCmpFn<K, MULTI, kModeCompare, false> fn{sb, &P, tile_n0, o_li0 + pre, take, 0, 0, {}, {}, {}, 0, 0};
                                chunk_walk(fn);
                            }
                        }
                        __syncwarp();
                        const int b = bmask ? (__ffs(bmask) - 1) : 0;
                        const u32 b_pre = __shfl_sync(0xFFFFFFFFu, pre, b);
                        r_nlt = __shfl_sync(0xFFFFFFFFu, r_nlt, b);
                        r_eq = __shfl_sync(0xFFFFFFFFu, r_eq, b);
                        if ((int)lane == owner) {
                            n_in = quota - skip0;
                            if (bmask) {
                                n_lt[0] = ((b_pre > skip0) ? (b_pre - skip0) : 0u) + r_nlt;
                                eq_li[0] = r_eq;
                            } else {
                                n_lt[0] = n_in;     // the crossing lies beyond this thread's quota (n_max): all of the window is below
                            }
                            localised = true;
                            walk = false;
                            coop = owns_nmax;       // S(n_max) and p_(n_max) by the warp-wide finish below
                        }
                    }
                }
                const u32 coop_mask = __ballot_sync(0xFFFFFFFFu, coop);
                {
                    if (coop_mask) {
                        // ---- S_k(n_max) and p_(n_max) by the whole warp.  One thread of the grid holds n_max; walking its
                        //      block alone is a serial chain of up to ~400 primes (the longest latency of a pruned query), so
                        //      lane l takes word l of that thread: popcount prefix -> which of its primes lie at or below
                        //      n_max -> their powers -> butterfly sum.  No per-prime compare is needed (no bracket in the block).
                        const int owner = __ffs(coop_mask) - 1;
                        u32 q = __shfl_sync(0xFFFFFFFFu, quota, owner);
                        const u32 vb_lo = __shfl_sync(0xFFFFFFFFu, (u32)value_base, owner);
                        const u32 vb_hi = __shfl_sync(0xFFFFFFFFu, (u32)(value_base >> 32), owner);
                        const bool o_small = __shfl_sync(0xFFFFFFFFu, has_small ? 1u : 0u, owner) != 0u;
                        u32 part[TOTAL];
#pragma unroll
                        for (int i = 0; i < TOTAL; ++i) part[i] = 0;
                        BsAccFn<K, MULTI> pf{part};
                        u64 my_last = 0;
                        u32 ns = 0;
                        if (o_small) {                      // 2, 3, 5 come first in that thread's sequence
                            ns = min(P.n_small, q);
                            if (lane == 0)
                                for (u32 i = 0; i < ns; ++i) {
                                    my_last = P.small_primes[i];
                                    pf((u32)my_last, 0u);
                                }
                            q -= ns;
                        }
                        u32 w = ((int)lane < WPT) ? my_warp_words[owner * PAD + lane] : 0u;   // bit-reversed: lowest prime = top bit
                        const u32 c = __popc(w);
                        u32 incl = c;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const u32 up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                            if (lane >= (u32)d) incl += up;
                        }
                        const u32 pre = incl - c;
                        const u32 take = (q > pre) ? min(c, q - pre) : 0u;
                        const u64 vb = ((u64)vb_hi << 32) | vb_lo;
                        for (u32 t = 0; t < take; ++t) {
                            const u32 m = msb_index(w);
                            w ^= (1u << m);
                            const u64 pv = vb + 120u * lane + msb_value_offset(m);
                            my_last = pv;
                            pf((u32)pv, (u32)(pv >> 32));
                        }
#pragma unroll
                        for (int d = 16; d >= 1; d >>= 1) {
                            OpShflXorAdd op{part, d};
                            for_each_power<0, Lay>(op);
                        }
                        // the lane whose word holds the q-th prime of the block knows p_(n_max) (lane 0 if n_max is one of 2, 3, 5)
                        const u32 last_mask = __ballot_sync(0xFFFFFFFFu, take > 0 && pre + take == q);
                        const int last_lane = (q > 0 && last_mask) ? (__ffs(last_mask) - 1) : 0;
                        const u32 lp_lo = __shfl_sync(0xFFFFFFFFu, (u32)my_last, last_lane);
                        const u32 lp_hi = __shfl_sync(0xFFFFFFFFu, (u32)(my_last >> 32), last_lane);
                        if ((int)lane == owner) {
                            OpAdd op{run, part};
                            for_each_power<0, Lay>(op);
                            *P.last_prime_out = ((u64)lp_hi << 32) | lp_lo;
#pragma unroll
                            for (int jj = 0; jj < NP; ++jj)
#pragma unroll
                                for (int i = 0; i < PSS_LIMBS; ++i)
                                    P.results[jj].final_sum[i] = (i < Lay::width(jj)) ? run[Lay::offset(jj) + i] : 0u;
                        }
                    }
                }
                if (walk) {
                    // ---- every S_k(n) of this thread formed and compared
                    auto run_walk = [&](auto& fn) {
#pragma unroll
                        for (int jj = 0; jj < NP; ++jj) {
                            fn.n_lt[jj] = 0;
                            fn.eq_li[jj] = 0xFFFFFFFFu;
                            fn.below[jj] = (MODE == kBsCompare) ? (c0[jj] < 0) : false;
                        }
                        bool go = true;
                        if (has_small)
                            for (u32 i = 0; i < P.n_small && go; ++i) {
                                last_p = P.small_primes[i];
                                go = fn((u32)last_p, 0u);
                            }
                        if (go) last_p = walk_words(my_words, value_base, fn);
                        n_in = fn.n_in;
#pragma unroll
                        for (int jj = 0; jj < NP; ++jj) {
                            n_lt[jj] = fn.n_lt[jj];
                            eq_li[jj] = fn.eq_li[jj];
                        }
                    };
                    if (skip0 != 0) {
                        BsCmpFn<K, MULTI, PMODE, true> fn{run, &P, tile_n0, li0, quota, skip0, 0, {}, {}, {}, 0, 0};
                        fn.tri_head_n = tri_head_n;
                        run_walk(fn);
                    } else {
                        BsCmpFn<K, MULTI, PMODE, false> fn{run, &P, tile_n0, li0, quota, 0, 0, {}, {}, {}, 0, 0};
                        fn.tri_head_n = tri_head_n;
                        run_walk(fn);
                    }
                    if (owns_nmax) {   // exactly one thread in the grid: S(n_max) and p_{n_max}
                        *P.last_prime_out = last_p;
#pragma unroll
                        for (int jj = 0; jj < NP; ++jj)
#pragma unroll
                            for (int i = 0; i < PSS_LIMBS; ++i)
                                P.results[jj].final_sum[i] = (i < Lay::width(jj)) ? run[Lay::offset(jj) + i] : 0u;
                    }
                } else if (quota > 0 && !localised) {
                    n_in = quota - skip0;
#pragma unroll
                    for (int jj = 0; jj < NP; ++jj) n_lt[jj] = (c0[jj] < 0) ? n_in : 0u;
                }
            }
            __syncwarp();
            if constexpr (MODE == kBsCompare) emit(a, n_in, n_lt, eq_li, tile_n0);
        }
      }
    }
}

// ---- reduce pass for k <= 5 (single power or all powers 1..K) from the small moments, as its own kernel --------------
// Same records as k_bitmap_scan<K, MULTI, kBsReduce> with reduce_chain = 0 (bit-identical), but compiled on its own: the
// walk needs ~10 accumulator registers and the powers are expanded and written out ONE AT A TIME, so the kernel fits 64
// registers and four CTAs per SM stay resident (the generic instantiation also carries the multi-limb power chain and
// holds all 29 limbs of a K = 5 record at once: 128 registers, two CTAs, issue-starved).
template <int J, int K, bool MULTI>
__device__ __forceinline__ void moments_emit(const BScanParams& P, const BsMomFn<K>& f, u32 m0, u32 blo, u32 bhi, bool has_small,
                                              u64 gthread, u64 wt, u32 tile, u32 lane) {
    using Lay = Layout<K, MULTI>;
    if constexpr (J < Lay::NP) {
        constexpr int KK = MULTI ? J + 1 : K;
        constexpr int W = Lay::width(J), OFF = Lay::offset(J);
        u32 out[W];
        expand_moments<KK, K>(out, f, m0, blo, bhi);
        if (has_small) {                      // 2, 3, 5 (one thread of the whole grid)
            for (u32 i = 0; i < P.n_small; ++i) {
                u64 pw = 1;
                for (int e = 0; e < KK; ++e) pw *= P.small_primes[i];   // <= 5^5
                const u32 t[2] = {(u32)pw, (u32)(pw >> 32)};
                acc_add<0, W, 2>(out, t);
            }
        }
#pragma unroll
        for (int i = 0; i < W; ++i) P.thread_rec[(u64)(1 + OFF + i) * P.thread_stride + gthread] = out[i];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            u32 o[W];
#pragma unroll
            for (int i = 0; i < W; ++i) o[i] = __shfl_xor_sync(0xFFFFFFFFu, out[i], d);
            acc_add<0, W, W>(out, o);
        }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < W; ++i) {
                P.warp_rec[(u64)(1 + OFF + i) * P.warp_stride + wt] = out[i];
                atomicAdd(reinterpret_cast<unsigned long long*>(P.cols + (u64)(1 + OFF + i) * P.col_stride + tile), (unsigned long long)out[i]);
            }
        }
        moments_emit<J + 1, K, MULTI>(P, f, m0, blo, bhi, has_small, gthread, wt, tile, lane);
    }
}

// Lean walk for the moments kernel: only the offset o = p - B of every prime is needed (no 64-bit prime value), the word
// refill is four predicated instructions, and the accumulators are fed with IMAD.WIDE accumulate forms / one carry chain:
// about 30 instructions per prime for K = 5 (the generic walk_words + BsMomFn pair compiles to 49).
template <int K>
__device__ __forceinline__ void walk_moments(u32 row_saddr, BsMomFn<K>& f) {
    u32 addr = row_saddr, w = 0, woff = 0u - 120u;
    const u32 end = row_saddr + 4u * (u32)kWordsPerThread;     // the sentinel slot
    u32 m1 = 0, m5a = 0, m5b = 0, m5c = 0;
    u64 m2 = 0, m3 = 0, m4 = 0;
    bool alive = true;
    while (alive) {
        // exhausted word: fetch the next one (the row ends with a zero sentinel)
        asm("{\n\t.reg .pred q;\n\tsetp.eq.u32 q, %0, 0;\n\t@q ld.shared.u32 %0, [%1];\n\t@q add.u32 %1, %1, 4;\n\t@q add.u32 %2, %2, 120;\n\t}"
            : "+r"(w), "+r"(addr), "+r"(woff));
        // a plain if-region, so the warp reconverges every iteration (a lane that meets an empty word loses one trip; a
        // `continue` here makes the compiler park that lane until every other lane leaves the loop: 19 of 32 lanes busy)
        if (w == 0) {
            alive = (addr <= end);      // empty word (rare): try the next one; past the sentinel: done
        } else {
        const u32 m = msb_index(w);
        w ^= (1u << m);
        const u32 o = woff + msb_value_offset(m);     // < 2880
        m1 += o;
        if constexpr (K >= 2) {
            const u32 o2 = o * o;                     // < 2^23
            asm("mad.wide.u32 %0, %1, %1, %0;" : "+l"(m2) : "r"(o));
            if constexpr (K >= 3) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(m3) : "r"(o2), "r"(o));
            if constexpr (K >= 4) {
                const u64 o4 = (u64)o2 * o2;          // < 2^46
                m4 += o4;
                if constexpr (K >= 5) {
                    const u64 t = (u64)(u32)o4 * o + ((u64)((u32)(o4 >> 32) * o) << 32);   // o^5 < 2^58
                    asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;"
                        : "+r"(m5a), "+r"(m5b), "+r"(m5c)
                        : "r"((u32)t), "r"((u32)(t >> 32)));
                }
            }
        }
        }
    }
    f.m1 = m1; f.m2 = m2; f.m3 = m3; f.m4 = m4;
    f.m5lo = ((u64)m5b << 32) | m5a;
    f.m5hi = m5c;
}

template <int K, bool MULTI>
__global__ void __launch_bounds__(kBsThreads, 4) k_bitmap_moments(const __grid_constant__ BScanParams P) {
    constexpr int T = kBsThreads, NWARP = T / 32, WPT = kWordsPerThread, PAD = WPT + 1, WARP_WORDS = 32 * WPT;
    __shared__ u32 s_words[T * PAD];
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    u32* my_warp_words = s_words + warp * 32 * PAD;
    const u32* my_words = my_warp_words + lane * PAD;
    my_warp_words[lane * PAD + WPT] = 0u;        // sentinel word of this thread's row (never overwritten)
    const u64 n_warp_tiles = (u64)P.n_tiles * NWARP;
    const u64 warp_stride = (u64)gridDim.x * NWARP;
    for (u64 wt = (u64)blockIdx.x * NWARP + warp; wt < n_warp_tiles; wt += warp_stride) {
        const u32 tile = (u32)(wt / NWARP);
        const u64 word0 = wt * WARP_WORDS;
        {   // warp-local staging: coalesced uint4 loads -> padded blocked layout, bit-reversed
            const uint4* g4 = reinterpret_cast<const uint4*>(P.bitmap + word0);
            __syncwarp();
#pragma unroll
            for (int r = 0; r < WPT / 4; ++r) {
                const u32 q = r * 32 + lane;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (word0 + 4ull * q + 4 <= P.n_words) v = __ldg(g4 + q);
                const u32 wl = 4u * q;
                const u32 dst = (wl / WPT) * PAD + (wl % WPT);
                my_warp_words[dst] = __brev(v.x); my_warp_words[dst + 1] = __brev(v.y);
                my_warp_words[dst + 2] = __brev(v.z); my_warp_words[dst + 3] = __brev(v.w);
            }
            __syncwarp();
        }
        const u64 value_base = (P.B_start + 4ull * (word0 + (u64)lane * WPT)) * 30ull;
        const bool has_small = (wt == 0 && lane == 0 && P.n_small > 0);
        const u64 gthread = wt * 32 + lane;
        u32 cnt = 0;
#pragma unroll
        for (int i = 0; i < WPT; ++i) cnt += __popc(my_words[i]);
        BsMomFn<K> mf{(u32)value_base, 0u, 0ull, 0ull, 0ull, 0ull, 0u};
        walk_moments<K>((u32)__cvta_generic_to_shared(my_words), mf);   // per prime: small moments of p - B
        const u32 m0 = cnt;
        if (has_small) cnt += P.n_small;
        P.thread_rec[gthread] = cnt;
        moments_emit<0, K, MULTI>(P, mf, m0, (u32)value_base, (u32)(value_base >> 32), has_small, gthread, wt, tile, lane);
        cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
        if (lane == 0) {
            P.warp_rec[wt] = cnt;
            atomicAdd(reinterpret_cast<unsigned long long*>(P.cols + tile), (unsigned long long)cnt);
        }
    }
}

// ---- reduce pass for k <= 5 WITHOUT touching individual primes (generalisation of k_bitmap_moments2) --------------------
// Byte table: for a bitmap byte v, mu_i(v) = sum of R[j]^i over its set bits, i = 0..5, packed in one uint4
//     .x = mu0 | mu3 << 12      .y = mu1 | mu2 << 14      .z = mu4      .w = mu5
// The low field of .x / .y has head-room for every byte-index weight it is used with (sum_q q^5 mu0 <= 2208 < 2^12,
// sum_q q^4 mu1 <= 11760 < 2^14), the high field may run off the top of the word where it is not needed, so ONE 32-bit
// multiply-add forms the weighted sum of both fields:  X_e = sum_q q^e E_q  gives  A(e,i) = sum_q q^e mu_i(byte q).
// Word moments about the word origin (byte q sits at 30 q):  w_j = sum_i C(j,i) 30^(j-i) A(j-i, i)  -- all below 2^32 except
// w_5 (< 2^37).  Thread moments: words are folded from the last to the first; before a word is added the running moments
// are re-based by 120 with the triangular scheme  M_j += 120 M_(j-1)  (j = K..t, passes t = 1..K), i.e. only
// multiplications by the constant 120.  ~150 instructions per 32-bit word (5.3 primes) with every lane busy, against
// ~44 per prime and 27 busy lanes for the walk.  The table is replicated 8x (32 KiB): lane l reads copy l & 7, so the
// eight lanes of every LDS.128 phase hit eight different 16-byte bank groups (conflict free).
__device__ __forceinline__ uint4 lut4_at(u32 lut_lane, u32 w, int q) {
    u32 b;
    uint4 e;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(b) : "r"(w), "r"(0x4440u + (u32)q));   // byte q, zero extended
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(e.x), "=r"(e.y), "=r"(e.z), "=r"(e.w) : "r"(lut_lane + (b << 7)));
    return e;
}

template <int K, bool MULTI>
__global__ void __launch_bounds__(kBsThreads, 3) k_bitmap_moments_tab(const __grid_constant__ BScanParams P) {
    constexpr int NWARP = kBsThreads / 32, WPT = kWordsPerThread, WARP_WORDS = 32 * WPT;
    __shared__ uint4 s_lut[256 * 8];
    {
        const u32 v = threadIdx.x;     // kBsThreads == 256: one byte value per thread
        u32 mu[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (u32 j = 0; j < 8; ++j)
            if ((v >> j) & 1u) {
                const u32 r = wheel_residue(j);
                u32 pw = 1;
#pragma unroll
                for (int i = 0; i < 6; ++i) { mu[i] += pw; pw *= r; }
            }
        const uint4 e = make_uint4(mu[0] | (mu[3] << 12), mu[1] | (mu[2] << 14), mu[4], mu[5]);
        for (u32 k = 0; k < 8; ++k) s_lut[v * 8 + ((k + v) & 7u)] = e;
    }
    __syncthreads();
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u32 lut_lane = (u32)__cvta_generic_to_shared(s_lut) + 16u * (lane & 7u);
    const u64 n_warp_tiles = (u64)P.n_tiles * NWARP;
    const u64 warp_stride = (u64)gridDim.x * NWARP;
    for (u64 wt = (u64)blockIdx.x * NWARP + warp; wt < n_warp_tiles; wt += warp_stride) {
        const u32 tile = (u32)(wt / NWARP);
        const u64 word0 = wt * WARP_WORDS + (u64)lane * WPT;      // this thread's first word
        const uint4* g4 = reinterpret_cast<const uint4*>(P.bitmap + word0);
        // D(e,i) = sum over the thread's words of (word index)^e * w_i(word); rigorous worst cases (every bit set) put
        // D04, D05, D14, D23, D32 beyond 32 bits, everything else below 2^32 (largest: D13 <= 3.81e9)
        u32 D00 = 0, D01 = 0, D02 = 0, D03 = 0, D10 = 0, D11 = 0, D12 = 0, D13 = 0, D20 = 0, D21 = 0, D22 = 0, D30 = 0, D31 = 0,
            D40 = 0, D41 = 0, D50 = 0;
        u64 D04 = 0, D05 = 0, D14 = 0, D23 = 0, D32 = 0;
#pragma unroll 1
        for (int g = 0; g < WPT / 4; ++g) {
            uint4 vv = make_uint4(0, 0, 0, 0);
            if (word0 + 4ull * g + 4 <= P.n_words) vv = __ldg(g4 + g);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const u32 w = c == 0 ? vv.x : c == 1 ? vv.y : c == 2 ? vv.z : vv.w;
                const u32 wi = 4u * (u32)g + (u32)c, wi2 = wi * wi, wi3 = wi2 * wi, wi4 = wi2 * wi2, wi5 = wi4 * wi;   // <= 23^5
                const uint4 e0 = lut4_at(lut_lane, w, 0), e1 = lut4_at(lut_lane, w, 1), e2 = lut4_at(lut_lane, w, 2), e3 = lut4_at(lut_lane, w, 3);
                // X(field, e) = sum_q q^e E_q
                const u32 x00 = e0.x + e1.x + e2.x + e3.x;
                const u32 x01 = e1.x + 2u * e2.x + 3u * e3.x;
                const u32 y00 = e0.y + e1.y + e2.y + e3.y;
                const u32 A00 = x00 & 0xFFFu, A10 = x01 & 0xFFFu;
                const u32 A01 = y00 & 0x3FFFu;
                const u32 w0 = A00, w1 = A01 + 30u * A10;
                D00 += w0; D01 += w1;
                D10 += wi * w0;
                if constexpr (K >= 2) {
                    const u32 x02 = e1.x + 4u * e2.x + 9u * e3.x;
                    const u32 y01 = e1.y + 2u * e2.y + 3u * e3.y;
                    const u32 A20 = x02 & 0xFFFu, A11 = y01 & 0x3FFFu, A02 = y00 >> 14;
                    const u32 w2 = A02 + 60u * A11 + 900u * A20;
                    D02 += w2; D11 += wi * w1; D20 += wi2 * w0;
                    if constexpr (K >= 3) {
                        const u32 x03 = e1.x + 8u * e2.x + 27u * e3.x;
                        const u32 y02 = e1.y + 4u * e2.y + 9u * e3.y;
                        const u32 A30 = x03 & 0xFFFu, A21 = y02 & 0x3FFFu, A12 = y01 >> 14, A03 = x00 >> 12;
                        const u32 w3 = A03 + 90u * A12 + 2700u * A21 + 27000u * A30;
                        D03 += w3; D12 += wi * w2; D21 += wi2 * w1; D30 += wi3 * w0;
                        if constexpr (K >= 4) {
                            const u32 x04 = e1.x + 16u * e2.x + 81u * e3.x;
                            const u32 y03 = e1.y + 8u * e2.y + 27u * e3.y;
                            const u32 z00 = e0.z + e1.z + e2.z + e3.z;
                            const u32 A40 = x04 & 0xFFFu, A31 = y03 & 0x3FFFu, A22 = y02 >> 14, A13 = x01 >> 12, A04 = z00;
                            const u32 w4 = A04 + 120u * A13 + 5400u * A22 + 108000u * A31 + 810000u * A40;   // <= 1.33e9
                            D04 += w4; D13 += wi * w3; D22 += wi2 * w2; D31 += wi3 * w1; D40 += wi4 * w0;
                            if constexpr (K >= 5) {
                                const u32 x05 = e1.x + 32u * e2.x + 243u * e3.x;
                                const u32 y04 = e1.y + 16u * e2.y + 81u * e3.y;
                                const u32 z01 = e1.z + 2u * e2.z + 3u * e3.z;
                                const u32 t00 = e0.w + e1.w + e2.w + e3.w;
                                const u32 A50 = x05 & 0xFFFu, A41 = y04 & 0x3FFFu, A32 = y03 >> 14, A23 = x02 >> 12, A14 = z01, A05 = t00;
                                const u32 inner = A14 + 60u * A23 + 1800u * A32 + 27000u * A41 + 162000u * A50;   // < 8.8e8
                                const u64 w5 = (u64)A05 + 150ull * inner;
                                D05 += w5; D14 += (u64)wi * w4; D23 += (u64)wi2 * w3; D32 += (u64)wi3 * w2; D41 += wi4 * w1; D50 += wi5 * w0;
                            }
                        }
                    }
                }
            }
        }
        // thread moments about the thread origin (word i sits at 120 i):  M_j = sum_i C(j,i) 120^(j-i) D(j-i, i)
        const u32 M0 = D00;
        const u32 M1 = D01 + 120u * D10;
        const u64 M2 = (u64)D02 + 240ull * D11 + 14400ull * D20;
        const u64 M3 = (u64)D03 + 360ull * D12 + 43200ull * D21 + 1728000ull * D30;
        const u64 M4 = D04 + 480ull * D13 + 86400ull * D22 + 6912000ull * D31 + 207360000ull * D40;
        const unsigned __int128 M5 = (unsigned __int128)D05 + (unsigned __int128)D14 * 600u + (unsigned __int128)D23 * 144000u +
                                     (unsigned __int128)D32 * 17280000u + (unsigned __int128)D41 * 1036800000u +
                                     (unsigned __int128)D50 * 24883200000ull;
        BsMomFn<K> mf{0u, M1, M2, M3, M4, (u64)M5, (u32)(M5 >> 64)};
        u32 cnt = M0;
        const u64 B = (P.B_start + 4ull * word0) * 30ull;
        const bool has_small = (wt == 0 && lane == 0 && P.n_small > 0);
        const u64 gthread = wt * 32 + lane;
        const u32 m0 = cnt;
        if (has_small) cnt += P.n_small;
        P.thread_rec[gthread] = cnt;
        moments_emit<0, K, MULTI>(P, mf, m0, (u32)B, (u32)(B >> 32), has_small, gthread, wt, tile, lane);
        cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
        if (lane == 0) {
            P.warp_rec[wt] = cnt;
            atomicAdd(reinterpret_cast<unsigned long long*>(P.cols + tile), (unsigned long long)cnt);
        }
    }
}

// ---- the same pass with position-specific byte tables (default; PSS_MOMENTS_WIDE=0 selects the kernel above) --------------
// Entry (q, v): sigma_i = sum of (R[j] + 30 q)^i over the set bits of v -- the moments of byte v AT byte position q about the
// WORD origin -- packed so that the four entries of a word add up field by field without overflow (maxima with every bit of
// the word set: sigma_1 1920 < 2^12, sigma_2 153 440 < 2^18, sigma_3 13 795 200 < 2^24, sigma_4 1 322 586 272 < 2^31,
// sigma_5 132 046 281 600 < 2^38):
//     .x = sigma_4      (.y, .z) as one u64 = sigma_5 | sigma_3 << 38      .w = sigma_1 | sigma_2 << 12
// so the word moments are w_0 = POPC, w_1..w_5 = the fields of the sum: 4 lookups (PRMT + LDS.128 each), 3 + 3 + 6 adds, 5
// field extractions -- all the byte-index weighting of the kernel above (14 weighted sums, 15 field extractions, 15
// multiply-adds per word) is gone.  Layout as in k_bitmap_moments2w: 256-byte rows, row v of a table pair = 8 replicas of
// (q even, v) then 8 replicas of (q odd, v), lane l reads replica l & 7 (the eight lanes of an LDS.128 phase hit eight
// different 16-byte bank groups); 2 x 64 KiB, one 768-thread CTA per SM (85 registers per thread available).
constexpr int kMomwThreads = 768;
constexpr size_t kMomwSmemBytes = 2 * 65536;

template <int Q>
__device__ __forceinline__ uint4 lutw_at(const char* tab, u32 w, u32 lane16) {
    u32 d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(lane16), "n"(0x5504 + 0x10 * Q));   // v << 8 | (lane & 7) << 4
    uint4 e = *reinterpret_cast<const uint4*>(tab + d + (Q >> 1) * 65536 + (Q & 1) * 128);
    // keep it ONE LDS.128 even where a power configuration ignores some fields: narrower loads of the same entry would put
    // the four lanes that share a replica on one bank (measured: k = 3 at 0.79 ms against 0.51 ms)
    asm volatile("" : "+r"(e.x), "+r"(e.y), "+r"(e.z), "+r"(e.w));
    return e;
}

// k <= 3: the three moments fit one 64-bit entry (sigma_1 | sigma_2 << 12 | sigma_3 << 30, 54 bits), 16 replicas of 8 bytes per
// half row: the 16 lanes of an LDS.64 phase read 16 different bank pairs
template <int Q>
__device__ __forceinline__ u64 lutn_at(const char* tab, u32 w, u32 lane8) {
    u32 d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(lane8), "n"(0x5504 + 0x10 * Q));    // v << 8 | (lane & 15) << 3
    uint2 e = *reinterpret_cast<const uint2*>(tab + d + (Q >> 1) * 65536 + (Q & 1) * 128);
    asm volatile("" : "+r"(e.x), "+r"(e.y));     // one LDS.64 whatever fields the power configuration uses
    return ((u64)e.y << 32) | e.x;
}

template <int K, bool MULTI>
__global__ void __launch_bounds__(kMomwThreads, 1) k_bitmap_moments_tabw(const __grid_constant__ BScanParams P) {
    constexpr bool NARROW = K <= 3;
    constexpr int NWARP = kMomwThreads / 32, TILE_WARPS = kBsThreads / 32, WPT = kWordsPerThread, WARP_WORDS = 32 * WPT;
    extern __shared__ uint4 momw_smem[];
    char* const tab = reinterpret_cast<char*>(momw_smem);
    for (u32 idx = threadIdx.x; idx < 1024u; idx += kMomwThreads) {
        const u32 v = idx & 255u, q = idx >> 8;
        u64 sg[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
        for (u32 j = 0; j < 8; ++j)
            if ((v >> j) & 1u) {
                const u64 r = wheel_residue(j) + 30u * q;
                u64 pw = 1;
#pragma unroll
                for (int i = 0; i < 6; ++i) { sg[i] += pw; pw *= r; }
            }
        char* const row_b = tab + (q >> 1) * 65536u + v * 256u + (q & 1u) * 128u;
        if constexpr (NARROW) {
            const u64 e = sg[1] | (sg[2] << 12) | (sg[3] << 30);
            u64* row = reinterpret_cast<u64*>(row_b);
            for (u32 k = 0; k < 16; ++k) row[(k + v) & 15u] = e;
        } else {
            const u64 yz = sg[5] | (sg[3] << 38);
            const uint4 e = make_uint4((u32)sg[4], (u32)yz, (u32)(yz >> 32), (u32)sg[1] | ((u32)sg[2] << 12));
            uint4* row = reinterpret_cast<uint4*>(row_b);
            for (u32 k = 0; k < 8; ++k) row[(k + v) & 7u] = e;
        }
    }
    __syncthreads();
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u32 lane16 = NARROW ? (lane & 15u) << 3 : (lane & 7u) << 4;
    const u64 n_warp_tiles = (u64)P.n_tiles * TILE_WARPS;
    const u64 warp_stride = (u64)gridDim.x * NWARP;
    for (u64 wt = (u64)blockIdx.x * NWARP + warp; wt < n_warp_tiles; wt += warp_stride) {
        const u32 tile = (u32)(wt / TILE_WARPS);
        const u64 word0 = wt * WARP_WORDS + (u64)lane * WPT;      // this thread's first word
        const uint4* g4 = reinterpret_cast<const uint4*>(P.bitmap + word0);
        // D(e,i) = sum over the thread's words of (word index)^e * w_i(word); same accumulator widths as the kernel above
        u32 D00 = 0, D01 = 0, D02 = 0, D03 = 0, D10 = 0, D11 = 0, D12 = 0, D13 = 0, D20 = 0, D21 = 0, D22 = 0, D30 = 0, D31 = 0,
            D40 = 0, D41 = 0, D50 = 0;
        u64 D04 = 0, D05 = 0, D14 = 0, D23 = 0, D32 = 0;
#pragma unroll 1
        for (int g = 0; g < WPT / 4; ++g) {
            uint4 vv = make_uint4(0, 0, 0, 0);
            if (word0 + 4ull * g + 4 <= P.n_words) vv = __ldg(g4 + g);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const u32 w = c == 0 ? vv.x : c == 1 ? vv.y : c == 2 ? vv.z : vv.w;
                const u32 wi = 4u * (u32)g + (u32)c, wi2 = wi * wi, wi3 = wi2 * wi, wi4 = wi2 * wi2, wi5 = wi4 * wi;   // <= 23^5
                const u32 w0 = __popc(w);
                if constexpr (NARROW) {
                    const u64 sn = lutn_at<0>(tab, w, lane16) + lutn_at<1>(tab, w, lane16) + lutn_at<2>(tab, w, lane16) +
                                   lutn_at<3>(tab, w, lane16);
                    const u32 w1 = (u32)sn & 0xFFFu;
                    D00 += w0; D01 += w1;
                    D10 += wi * w0;
                    if constexpr (K >= 2) {
                        const u32 w2 = (u32)(sn >> 12) & 0x3FFFFu;
                        D02 += w2; D11 += wi * w1; D20 += wi2 * w0;
                        if constexpr (K >= 3) {
                            const u32 w3 = (u32)(sn >> 30);
                            D03 += w3; D12 += wi * w2; D21 += wi2 * w1; D30 += wi3 * w0;
                        }
                    }
                    continue;
                }
                const uint4 e0 = lutw_at<0>(tab, w, lane16), e1 = lutw_at<1>(tab, w, lane16), e2 = lutw_at<2>(tab, w, lane16),
                            e3 = lutw_at<3>(tab, w, lane16);
                const u32 sw = e0.w + e1.w + e2.w + e3.w;
                const u32 w1 = sw & 0xFFFu;
                D00 += w0; D01 += w1;
                D10 += wi * w0;
                if constexpr (K >= 2) {
                    const u32 w2 = sw >> 12;
                    D02 += w2; D11 += wi * w1; D20 += wi2 * w0;
                    if constexpr (K >= 3) {
                        const u64 syz = (((u64)e0.z << 32) | e0.y) + (((u64)e1.z << 32) | e1.y) + (((u64)e2.z << 32) | e2.y) +
                                        (((u64)e3.z << 32) | e3.y);
                        const u32 w3 = (u32)(syz >> 38);
                        D03 += w3; D12 += wi * w2; D21 += wi2 * w1; D30 += wi3 * w0;
                        if constexpr (K >= 4) {
                            const u32 w4 = e0.x + e1.x + e2.x + e3.x;
                            D04 += w4; D13 += wi * w3; D22 += wi2 * w2; D31 += wi3 * w1; D40 += wi4 * w0;
                            if constexpr (K >= 5) {
                                const u64 w5 = syz & ((1ull << 38) - 1ull);
                                D05 += w5; D14 += (u64)wi * w4; D23 += (u64)wi2 * w3; D32 += (u64)wi3 * w2; D41 += wi4 * w1; D50 += wi5 * w0;
                            }
                        }
                    }
                }
            }
        }
        // thread moments about the thread origin (word i sits at 120 i):  M_j = sum_i C(j,i) 120^(j-i) D(j-i, i)
        const u32 M0 = D00;
        const u32 M1 = D01 + 120u * D10;
        const u64 M2 = (u64)D02 + 240ull * D11 + 14400ull * D20;
        const u64 M3 = (u64)D03 + 360ull * D12 + 43200ull * D21 + 1728000ull * D30;
        const u64 M4 = D04 + 480ull * D13 + 86400ull * D22 + 6912000ull * D31 + 207360000ull * D40;
        const unsigned __int128 M5 = (unsigned __int128)D05 + (unsigned __int128)D14 * 600u + (unsigned __int128)D23 * 144000u +
                                     (unsigned __int128)D32 * 17280000u + (unsigned __int128)D41 * 1036800000u +
                                     (unsigned __int128)D50 * 24883200000ull;
        BsMomFn<K> mf{0u, M1, M2, M3, M4, (u64)M5, (u32)(M5 >> 64)};
        u32 cnt = M0;
        const u64 B = (P.B_start + 4ull * word0) * 30ull;
        const bool has_small = (wt == 0 && lane == 0 && P.n_small > 0);
        const u64 gthread = wt * 32 + lane;
        const u32 m0 = cnt;
        if (has_small) cnt += P.n_small;
        P.thread_rec[gthread] = cnt;
        moments_emit<0, K, MULTI>(P, mf, m0, (u32)B, (u32)(B >> 32), has_small, gthread, wt, tile, lane);
        cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
        if (lane == 0) {
            P.warp_rec[wt] = cnt;
            atomicAdd(reinterpret_cast<unsigned long long*>(P.cols + tile), (unsigned long long)cnt);
        }
    }
}

// s_lut[(byte q of w) * 32 + lane] through its 32-bit shared address (lut_lane = &s_lut[lane])
__device__ __forceinline__ u32 lut_at(u32 lut_lane, u32 w, int q) {
    u32 b, e;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(b) : "r"(w), "r"(0x4440u + (u32)q));   // byte q, zero extended
    asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(lut_lane + (b << 7)));
    return e;
}

// ---- reduce pass for the sum of SQUARES without touching individual primes ---------------------------------------
// A thread's 24 words cover the 2880 integers [B, B + 2880).  With o = p - B:
//     sum p^2 = M0 * B^2 + 2 B * M1 + M2,      M0 = #primes, M1 = sum o, M2 = sum o^2        (exact)
// and the moments come from a 256-entry byte table (count, sum r, sum r^2 of the wheel residues r of the set bits,
// packed 7 + 10 + 14 bits so that the weighted sums of four entries used below cannot overflow the fields they are
// read from) plus constant byte / word position weights.
// ~32 instructions per 32-bit word (~5 primes) instead of ~48 per prime for decode + square + accumulate.  The output
// (thread / warp records, scan-tile columns) is bit-identical to k_bitmap_scan<2, false, kBsReduce>.
__global__ void __launch_bounds__(kBsThreads) k_bitmap_moments2(const __grid_constant__ BScanParams P) {
    constexpr int NWARP = kBsThreads / 32, WPT = kWordsPerThread, WARP_WORDS = 32 * WPT;
    // the byte table, replicated once per bank: lane l only ever reads word v * 32 + l, so the four lookups per
    // word are bank-conflict free whatever the byte values are (32 KiB; a single copy costs ~4.5 wavefronts per LDS)
    __shared__ u32 s_lut[256 * 32];
    {
        const u32 v = threadIdx.x;     // kBsThreads == 256: one byte value per thread
        u32 n = 0, r1 = 0, r2 = 0;
#pragma unroll
        for (u32 j = 0; j < 8; ++j)
            if ((v >> j) & 1u) {
                const u32 r = wheel_residue(j);
                n += 1u; r1 += r; r2 += r * r;
            }
        const u32 e = n | (r1 << 7) | (r2 << 17);
        for (u32 k = 0; k < 32; ++k) s_lut[v * 32 + ((k + v) & 31u)] = e;   // rotated: conflict-free stores too
    }
    __syncthreads();
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u64 n_warp_tiles = (u64)P.n_tiles * NWARP;
    const u64 warp_stride = (u64)gridDim.x * NWARP;
    for (u64 wt = (u64)blockIdx.x * NWARP + warp; wt < n_warp_tiles; wt += warp_stride) {
        const u32 tile = (u32)(wt / NWARP);
        const u64 word0 = wt * WARP_WORDS + (u64)lane * WPT;      // this thread's first word
        const uint4* g4 = reinterpret_cast<const uint4*>(P.bitmap + word0);
        uint4 v[WPT / 4];
#pragma unroll
        for (int q = 0; q < WPT / 4; ++q) {
            v[q] = make_uint4(0, 0, 0, 0);
            if (word0 + 4ull * q + 4 <= P.n_words) v[q] = __ldg(g4 + q);
        }
        // per word i (origin 120 i inside the thread's range): m0, m1, m2 about the word origin; the shift to the
        // thread origin is deferred: M1 = sum m1 + 120 sum i m0,  M2 = sum m2 + 240 sum i m1 + 14400 sum i^2 m0
        // (all partial sums fit 32 bits: 768 bits set at most, i <= 23)
        u32 M0 = 0, A1 = 0, A2 = 0, C1 = 0, C2 = 0, D1 = 0;
        const u32 lut_lane = (u32)__cvta_generic_to_shared(s_lut) + 4u * lane;
#pragma unroll
        for (int i = 0; i < WPT; ++i) {
            const uint4 vv = v[i >> 2];
            const u32 w = (i & 3) == 0 ? vv.x : (i & 3) == 1 ? vv.y : (i & 3) == 2 ? vv.z : vv.w;
            // table word (byte value << 5) + lane: one PRMT (byte extract) + one LEA (address) + one LDS per byte
            const u32 e0 = lut_at(lut_lane, w, 0), e1 = lut_at(lut_lane, w, 1), e2 = lut_at(lut_lane, w, 2), e3 = lut_at(lut_lane, w, 3);
            const u32 W = e0 + e1 + e2 + e3;                       // moments about each byte's own origin
            const u32 V = e1 + 2u * e2 + 3u * e3;                  // byte-index weighted: count and sum r fields are valid
            const u32 U = e1 + 4u * e2 + 9u * e3;                  // byte-index^2 weighted: only the count field (<= 112) is valid
            const u32 m0 = W & 127u;
            const u32 m1 = ((W >> 7) & 1023u) + 30u * (V & 127u);                    // about the word origin
            const u32 m2 = (W >> 17) + 60u * ((V >> 7) & 1023u) + 900u * (U & 127u);
            M0 += m0;
            A1 += m1;
            A2 += m2;
            C1 += (u32)i * m0;
            C2 += (u32)(i * i) * m0;
            D1 += (u32)i * m1;
        }
        const u32 M1 = A1 + 120u * C1;
        const u64 M2 = (u64)A2 + 240ull * D1 + 14400ull * C2;
        // thread sum = M0 * B^2 + 2 B M1 + M2
        const u64 B = (P.B_start + 4ull * word0) * 30ull;
        const u32 blo = (u32)B, bhi = (u32)(B >> 32);
        u32 acc[4] = {0, 0, 0, 0};
        {
            Big<2> b1;  b1.v[0] = blo;  b1.v[1] = bhi;
            Big<3> b2;  big_mul_p<3, 2>(b2, b1, blo, bhi);
            Big<4> t;   big_mul_p<4, 3>(t, b2, M0, 0u);
            acc_add<0, 4, 4>(acc, t.v);
            Big<2> m;   m.v[0] = 2u * M1;  m.v[1] = 0u;
            Big<3> x;   big_mul_p<3, 2>(x, m, blo, bhi);
            acc_add<0, 4, 3>(acc, x.v);
            const u32 m2v[2] = {(u32)M2, (u32)(M2 >> 32)};
            acc_add<0, 4, 2>(acc, m2v);
        }
        u32 cnt = M0;
        if (wt == 0 && lane == 0 && P.n_small > 0) {
            for (u32 i = 0; i < P.n_small; ++i) square_acc4(acc, (u32)P.small_primes[i], 0u);
            cnt += P.n_small;
        }
        bs_reduce_tail<2, false>(P, acc, cnt, wt, tile, lane);
    }
}

// ---- the same pass with position-specific byte tables (default; PSS_MOMENTS2_WIDE=0 selects the kernel above) -------------
// Four tables, one per byte position q of a word: entry (q, v) = sum r' | sum r'^2 << 12 over the set bits of v with
// r' = wheel residue + 30 q, i.e. the moments about the WORD origin, so the four entries of a word simply add up (fields sized
// for the sum: 2368 < 2^12, 453 152 < 2^19) and the count is a POPC.  That removes the byte-index weighting (V, U and the
// field arithmetic around them: ~36 -> ~20 instructions per word).  Layout: 256-byte rows, row v of a table PAIR holds the 32
// per-lane replicas of (q even, v) in its first 128 bytes and those of (q odd, v) in the second, so the address of a lookup
// is (v << 8 | lane << 2) + constant: ONE PRMT builds it from the bitmap word and the lane offset, the constant goes into
// the LDS immediate.  2 x 64 KiB of shared memory, one 1024-thread CTA per SM; records bit-identical to the kernel above.
constexpr int kMom2wThreads = 1024;
constexpr size_t kMom2wSmemBytes = 2 * 65536;

template <int Q>
__device__ __forceinline__ u32 lut4_at(const char* tab, u32 w, u32 lane4) {
    u32 d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(lane4), "n"(0x5504 + 0x10 * Q));   // v << 8 | lane << 2
    return *reinterpret_cast<const u32*>(tab + d + (Q >> 1) * 65536 + (Q & 1) * 128);
}

__global__ void __launch_bounds__(kMom2wThreads, 1) k_bitmap_moments2w(const __grid_constant__ BScanParams P) {
    constexpr int NWARP = kMom2wThreads / 32, TILE_WARPS = kBsThreads / 32, WPT = kWordsPerThread, WARP_WORDS = 32 * WPT;
    extern __shared__ uint4 mom2w_smem[];
    char* const tab = reinterpret_cast<char*>(mom2w_smem);
    {
        const u32 v = threadIdx.x & 255u, q = threadIdx.x >> 8;     // 1024 threads: one (position, byte value) each
        u32 r1 = 0, r2 = 0;
#pragma unroll
        for (u32 j = 0; j < 8; ++j)
            if ((v >> j) & 1u) {
                const u32 r = wheel_residue(j) + 30u * q;
                r1 += r; r2 += r * r;
            }
        const u32 e = r1 | (r2 << 12);
        u32* row = reinterpret_cast<u32*>(tab + (q >> 1) * 65536u + v * 256u + (q & 1u) * 128u);
        for (u32 k = 0; k < 32; ++k) row[(k + v) & 31u] = e;
    }
    __syncthreads();
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u32 lane4 = lane << 2;
    const u64 n_warp_tiles = (u64)P.n_tiles * TILE_WARPS;
    const u64 warp_stride = (u64)gridDim.x * NWARP;
    for (u64 wt = (u64)blockIdx.x * NWARP + warp; wt < n_warp_tiles; wt += warp_stride) {
        const u32 tile = (u32)(wt / TILE_WARPS);
        const u64 word0 = wt * WARP_WORDS + (u64)lane * WPT;      // this thread's first word
        const uint4* g4 = reinterpret_cast<const uint4*>(P.bitmap + word0);
        uint4 v[WPT / 4];
#pragma unroll
        for (int q = 0; q < WPT / 4; ++q) {
            v[q] = make_uint4(0, 0, 0, 0);
            if (word0 + 4ull * q + 4 <= P.n_words) v[q] = __ldg(g4 + q);
        }
        // M1 = sum m1 + 120 sum i m0,  M2 = sum m2 + 240 sum i m1 + 14400 sum i^2 m0 (all partial sums fit 32 bits)
        u32 M0 = 0, A1 = 0, A2 = 0, C1 = 0, C2 = 0, D1 = 0;
#pragma unroll
        for (int i = 0; i < WPT; ++i) {
            const uint4 vv = v[i >> 2];
            const u32 w = (i & 3) == 0 ? vv.x : (i & 3) == 1 ? vv.y : (i & 3) == 2 ? vv.z : vv.w;
            const u32 W = lut4_at<0>(tab, w, lane4) + lut4_at<1>(tab, w, lane4) + lut4_at<2>(tab, w, lane4) + lut4_at<3>(tab, w, lane4);
            const u32 m0 = __popc(w), m1 = W & 4095u, m2 = W >> 12;    // about the word origin
            M0 += m0;
            A1 += m1;
            A2 += m2;
            C1 += (u32)i * m0;
            C2 += (u32)(i * i) * m0;
            D1 += (u32)i * m1;
        }
        const u32 M1 = A1 + 120u * C1;
        const u64 M2 = (u64)A2 + 240ull * D1 + 14400ull * C2;
        // thread sum = M0 * B^2 + 2 B M1 + M2
        const u64 B = (P.B_start + 4ull * word0) * 30ull;
        const u32 blo = (u32)B, bhi = (u32)(B >> 32);
        u32 acc[4] = {0, 0, 0, 0};
        {
            Big<2> b1;  b1.v[0] = blo;  b1.v[1] = bhi;
            Big<3> b2;  big_mul_p<3, 2>(b2, b1, blo, bhi);
            Big<4> t;   big_mul_p<4, 3>(t, b2, M0, 0u);
            acc_add<0, 4, 4>(acc, t.v);
            Big<2> m;   m.v[0] = 2u * M1;  m.v[1] = 0u;
            Big<3> x;   big_mul_p<3, 2>(x, m, blo, bhi);
            acc_add<0, 4, 3>(acc, x.v);
            const u32 m2v[2] = {(u32)M2, (u32)(M2 >> 32)};
            acc_add<0, 4, 2>(acc, m2v);
        }
        u32 cnt = M0;
        if (wt == 0 && lane == 0 && P.n_small > 0) {
            for (u32 i = 0; i < P.n_small; ++i) square_acc4(acc, (u32)P.small_primes[i], 0u);
            cnt += P.n_small;
        }
        bs_reduce_tail<2, false>(P, acc, cnt, wt, tile, lane);
    }
}

// ---- K1c, warp-autonomous form: compaction of bitmap chunks into the coalesced u64 prime buffer ------------------------------
// One warp per 1 KiB chunk (30 720 integers), no block barriers.  Lane l owns the chunk's words 8l .. 8l+7 (two coalesced
// LDG.128 per lane), so the lanes' primes are consecutive runs of the output: popc + warp scan give every lane its offset,
// the lane walks its 256-bit stream MSB-first (bit-reversed words in a padded shared-memory row, branch-free refill, one
// FLO + one PRMT per prime) and parks the 15-bit in-chunk offsets as u16 in the warp's staging area; the warp then writes
// the chunk's primes as consecutive u64 (value base + offset): 256 B per store instruction, fully coalesced.  HBM traffic
// is the algorithmic minimum (bitmap read once, 8 B per prime written); shared memory per warp 9.3 KiB (24 warps per SM).
constexpr int kCwWarps = 8, kCwRow = 9;
constexpr size_t kCwSmemBytes = (size_t)kCwWarps * (32 * kCwRow * sizeof(u32) + kMaxChunkPrimes * sizeof(uint16_t)) + 32 * sizeof(u32);

__global__ void __launch_bounds__(kCwWarps * 32, 3) k_compact_warp(const CompactParams P) {
    extern __shared__ u32 cw_smem[];
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    u32* my_rows = cw_smem + warp * (32 * kCwRow);
    uint16_t* my_stage = reinterpret_cast<uint16_t*>(cw_smem + kCwWarps * 32 * kCwRow) + warp * kMaxChunkPrimes;
    const u32 cpt = P.chunks_per_tile;
    const u64 n_chunks = (u64)P.n_groups;      // this kernel: one work item per chunk
    my_rows[lane * kCwRow + 8] = 0u;           // sentinel word of the lane's row
    const u32 row_saddr = (u32)__cvta_generic_to_shared(my_rows + lane * kCwRow);
    const u32 stage_saddr = (u32)__cvta_generic_to_shared(my_stage);
    // in-word offset of the prime whose (bit-reversed) bit index is m: one 32-entry table, one entry per bank, so any mix
    // of indices across the lanes is conflict free
    u32* tab = cw_smem + kCwWarps * 32 * kCwRow + kCwWarps * (kMaxChunkPrimes / 2);
    if (tid < 32) tab[tid] = msb_value_offset(tid);
    __syncthreads();
    const u32 tab_saddr = (u32)__cvta_generic_to_shared(tab);
    for (u64 g = (u64)blockIdx.x * kCwWarps + warp; g < n_chunks; g += (u64)gridDim.x * kCwWarps) {
        const u32 t = (u32)(g / cpt), c = (u32)(g - (u64)t * cpt);
        // output index of the chunk's first prime = tile prefix + counts of the tile's earlier chunks
        u32 pre = 0;
        for (u32 i = lane; i < c; i += 32) pre += __ldg(P.chunk_counts + (u64)t * cpt + i);
        pre = __reduce_add_sync(0xFFFFFFFFu, pre);
        const u64 base = __ldg(P.tile_prefix + t) + pre;
        if (base >= P.out_cap) continue;       // warp-uniform (first-n truncation)
        const uint4* src = reinterpret_cast<const uint4*>(P.bitmap + g * kChunkWords) + 2u * lane;
        const uint4 a = __ldg(src), b = __ldg(src + 1);
        const u32 cnt = __popc(a.x) + __popc(a.y) + __popc(a.z) + __popc(a.w) + __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
        u32 incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (u32)d) incl += o;
        }
        u32 total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (total > (u32)kMaxChunkPrimes) {    // cannot happen for a correctly sieved bitmap
            if (lane == 0) atomicExch(P.error_flag, 2u);
            total = kMaxChunkPrimes;
        }
        u32* row = my_rows + lane * kCwRow;
        row[0] = __brev(a.x); row[1] = __brev(a.y); row[2] = __brev(a.z); row[3] = __brev(a.w);
        row[4] = __brev(b.x); row[5] = __brev(b.y); row[6] = __brev(b.z); row[7] = __brev(b.w);
        // walk the lane's eight words (own row: no synchronisation needed before reading it back)
        u32 off = incl - cnt;
        u32 addr = row_saddr, w = 0, woff = 960u * lane - 120u;
        const u32 end = row_saddr + 32u;       // the sentinel slot
        bool alive = true;
        while (alive) {
            asm("{\n\t.reg .pred q;\n\tsetp.eq.u32 q, %0, 0;\n\t@q ld.shared.u32 %0, [%1];\n\t@q add.u32 %1, %1, 4;\n\t@q add.u32 %2, %2, 120;\n\t}"
                : "+r"(w), "+r"(addr), "+r"(woff));
            if (w == 0) {
                alive = (addr <= end);
            } else {
                const u32 m = msb_index(w);
                u32 below, o;
                asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(below) : "r"(m));              // bits under the leading one
                asm("ld.shared.u32 %0, [%1];" : "=r"(o) : "r"(tab_saddr + 4u * m));
                w &= below;
                // (off is below kMaxChunkPrimes for every valid bitmap; the mask only keeps a corrupt one inside the stage)
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(stage_saddr + 2u * (off & (u32)(kMaxChunkPrimes - 1))), "h"((unsigned short)(woff + o)) : "memory");
                ++off;
            }
        }
        __syncwarp();
        const u64 vbase = (P.B_start + g * (u64)kChunkBytes) * 30ull;
        for (u32 i = lane; i < total; i += 32) {
            const u64 o = base + i;
            if (o < P.out_cap) P.out[o] = vbase + my_stage[i];
        }
        __syncwarp();
    }
}

// ---- K1c, table form: the same warp-per-chunk compaction without a data-dependent loop -----------------------------------
// A 256-entry byte table lists the wheel residues of a byte's set bits in ascending order (5 bits each) and their count.
// Every lane runs the SAME 32 steps, one per bitmap byte it owns: one table lookup, up to four predicated 16-bit stores into
// the warp's staging area (bytes with five or more primes take a branch: common only near the origin), offset += count.
// No FLO, no refill, no lane ever idles: ~25 issue slots per byte instead of ~26 per prime with 25 of 32 lanes busy.
// The table is replicated 8x (16 KiB): lane l reads copy l & 7.
constexpr int kClStage = 3584;      // >= pi(30720) = 3316 primes in the densest chunk
constexpr size_t kClSmemBytes = 256 * 8 * sizeof(uint2) + (size_t)kCwWarps * kClStage * sizeof(uint16_t);

__global__ void __launch_bounds__(kCwWarps * 32, 3) k_compact_lut(const CompactParams P) {
    extern __shared__ u32 cl_smem[];
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint2* lut = reinterpret_cast<uint2*>(cl_smem);
    uint16_t* my_stage = reinterpret_cast<uint16_t*>(cl_smem + 256 * 8 * 2) + warp * kClStage;
    {
        const u32 v = tid;               // 256 threads: one byte value each
        u32 lo = 0, hi = 0, n = 0;
#pragma unroll
        for (u32 j = 0; j < 8; ++j)
            if ((v >> j) & 1u) {
                const u32 r = wheel_residue(j);
                if (n < 6) lo |= r << (5 * n); else hi |= r << (5 * (n - 6));
                ++n;
            }
        hi |= n << 16;
        for (u32 c = 0; c < 8; ++c) lut[v * 8 + ((c + v) & 7u)] = make_uint2(lo, hi);
    }
    __syncthreads();
    const u32 lut_lane = (u32)__cvta_generic_to_shared(lut) + 8u * (lane & 7u);
    const u32 cpt = P.chunks_per_tile;
    const u64 n_chunks = (u64)P.n_groups;      // one work item per chunk
    const u64 g_stride = (u64)gridDim.x * kCwWarps;
    auto load_chunk = [&](u64 g, uint4& a, uint4& b) {      // the lane's eight words of chunk g (zeros past the end)
        a = make_uint4(0, 0, 0, 0); b = a;
        if (g < n_chunks) {
            const uint4* src = reinterpret_cast<const uint4*>(P.bitmap + g * kChunkWords) + 2u * lane;
            a = __ldg(src); b = __ldg(src + 1);
        }
    };
    uint4 na, nb;
    load_chunk((u64)blockIdx.x * kCwWarps + warp, na, nb);
    for (u64 g = (u64)blockIdx.x * kCwWarps + warp; g < n_chunks; g += g_stride) {
        const uint4 a = na, b = nb;
        load_chunk(g + g_stride, na, nb);       // the next chunk's words fly while this one is decoded and written out
        const u32 t = (u32)(g / cpt), c = (u32)(g - (u64)t * cpt);
        u32 pre = 0;
        for (u32 i = lane; i < c; i += 32) pre += __ldg(P.chunk_counts + (u64)t * cpt + i);
        pre = __reduce_add_sync(0xFFFFFFFFu, pre);
        const u64 base = __ldg(P.tile_prefix + t) + pre;
        if (base >= P.out_cap) continue;       // warp-uniform (first-n truncation)
        const u32 wd[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        u32 cnt = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) cnt += __popc(wd[i]);
        u32 incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (u32)d) incl += o;
        }
        const u32 total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (total > (u32)kClStage) {            // cannot happen for a correctly sieved bitmap
            if (lane == 0) atomicExch(P.error_flag, 2u);
            continue;
        }
        u32 st = (u32)__cvta_generic_to_shared(my_stage) + 2u * (incl - cnt);   // shared address of the lane's next slot
        const u32 obase = 960u * lane;          // in-chunk offset of the lane's first byte
#pragma unroll
        for (int wi = 0; wi < 8; ++wi) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                u32 byte;
                uint2 e;
                asm("prmt.b32 %0, %1, 0, %2;" : "=r"(byte) : "r"(wd[wi]), "r"(0x4440u + (u32)q));
                asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(lut_lane + (byte << 6)));
                const u32 n = e.y >> 16;
                const u32 x = obase + 30u * (u32)(4 * wi + q);
                // four predicated stores (inline PTX: the compiler turns `if (n > k) st[k] = ..` into a branch each)
                asm volatile("{\n\t.reg .pred p0, p1, p2, p3;\n\t"
                             "setp.gt.u32 p0, %0, 0;\n\tsetp.gt.u32 p1, %0, 1;\n\tsetp.gt.u32 p2, %0, 2;\n\tsetp.gt.u32 p3, %0, 3;\n\t"
                             "@p0 st.shared.u16 [%1], %2;\n\t@p1 st.shared.u16 [%1+2], %3;\n\t@p2 st.shared.u16 [%1+4], %4;\n\t@p3 st.shared.u16 [%1+6], %5;\n\t}"
                             ::"r"(n), "r"(st), "h"((unsigned short)(x + (e.x & 31u))), "h"((unsigned short)(x + ((e.x >> 5) & 31u))),
                               "h"((unsigned short)(x + ((e.x >> 10) & 31u))), "h"((unsigned short)(x + ((e.x >> 15) & 31u)))
                             : "memory");
                if (n > 4) {                     // rare away from the origin (at 4 % prime density about one byte in 10^4)
                    asm volatile("st.shared.u16 [%0+8], %1;" ::"r"(st), "h"((unsigned short)(x + ((e.x >> 20) & 31u))) : "memory");
                    if (n > 5) asm volatile("st.shared.u16 [%0+10], %1;" ::"r"(st), "h"((unsigned short)(x + ((e.x >> 25) & 31u))) : "memory");
                    if (n > 6) asm volatile("st.shared.u16 [%0+12], %1;" ::"r"(st), "h"((unsigned short)(x + (e.y & 31u))) : "memory");
                    if (n > 7) asm volatile("st.shared.u16 [%0+14], %1;" ::"r"(st), "h"((unsigned short)(x + ((e.y >> 5) & 31u))) : "memory");
                }
                st += 2u * n;
            }
        }
        __syncwarp();
        const u64 vbase = (P.B_start + g * (u64)kChunkBytes) * 30ull;
        const u32 n_out = (u32)min((u64)total, P.out_cap - base);      // first-n truncation
        // 16-byte stores: entry pairs (i, i+1) with base + i even; a leading odd entry and a trailing single go alone
        const u32 head = (u32)(base & 1ull) & (n_out ? 1u : 0u);
        if (lane == 0 && head) P.out[base] = vbase + my_stage[0];
        const u32 n_pairs = (n_out - head) >> 1;
        ulonglong2* out2 = reinterpret_cast<ulonglong2*>(P.out + base + head);
        for (u32 i = lane; i < n_pairs; i += 32) {
            const u32 k = head + 2u * i;
            out2[i] = make_ulonglong2(vbase + my_stage[k], vbase + my_stage[k + 1]);
        }
        if (lane == 0 && ((n_out - head) & 1u)) P.out[base + n_out - 1] = vbase + my_stage[n_out - 1];
        __syncwarp();
    }
}

// ---- grid level of the scan: exclusive u64 prefix of every (deferred-carry) column, one CTA per column --------------
// out[c * stride + t] = sum_{i<t} in[c * stride + i]  for t = 0..n  (entry n = column total)
__global__ void __launch_bounds__(1024) k_scan_columns(const u64* in, u64* out, u32 n, u64 stride) {
    __shared__ u64 s_warp[32];
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u64* col = in + (u64)blockIdx.x * stride;
    u64* o = out + (u64)blockIdx.x * stride;
    // every warp owns a contiguous run of rows of 32 elements (coalesced loads): warp total -> one block-wide scan of
    // the 32 totals -> row-by-row warp scans seeded with the warp's offset
    const u32 rows = (n + 31u) / 32u;
    const u32 rows_per_warp = (rows + 31u) / 32u;
    const u32 r0 = min(rows, warp * rows_per_warp), r1 = min(rows, r0 + rows_per_warp);
    u64 sum = 0;
#pragma unroll 8
    for (u32 r = r0; r < r1; ++r) {
        const u32 i = r * 32u + lane;
        sum += (i < n) ? col[i] : 0ull;
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    if (lane == 0) s_warp[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        const u64 w = s_warp[lane];
        u64 wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 up = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= (u32)d) wi += up;
        }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    u64 run = s_warp[warp];
    // (the loads do not depend on the running sum: unrolled so that eight of them are in flight per warp)
#pragma unroll 8
    for (u32 r = r0; r < r1; ++r) {
        const u32 i = r * 32u + lane;
        const u64 v = (i < n) ? col[i] : 0ull;
        u64 incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (u32)d) incl += up;
        }
        if (i < n) o[i] = run + incl - v;
        run += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (warp == 31 && lane == 0) o[n] = run;     // the last warp's running sum is the column total (also for empty runs)
}

// register-resident form for n <= 16384 rows-of-32 per warp budget (host: launch_scan_columns)
__global__ void __launch_bounds__(512) k_scan_columns_small(const u64* in, u64* out, u32 n, u64 stride) {
    __shared__ u64 s_warp[32];          // 16 warps: entries 16..31 stay zero
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u64* col = in + (u64)blockIdx.x * stride;
    u64* o = out + (u64)blockIdx.x * stride;
    // every warp owns a contiguous run of rows of 32 elements (coalesced loads): warp total -> one block-wide scan of
    // the 32 totals -> row-by-row warp scans seeded with the warp's offset
    const u32 rows = (n + 31u) / 32u;
    const u32 rows_per_warp = (rows + 15u) / 16u;     // 16 warps
    const u32 r0 = min(rows, warp * rows_per_warp), r1 = min(rows, r0 + rows_per_warp);
    // up to 8192 scan tiles (6e9 integers per slice: the slices of the headline job from 4 GPUs up): the warp's rows are
    // loaded ONCE, all loads in flight together (the kernel is pure latency: one CTA per column), and both passes run
    // out of registers (16 x u64 per thread, 512-thread CTA)
    u64 v[16];
#pragma unroll
    for (u32 k = 0; k < 16u; ++k) {
        const u32 i = (r0 + k) * 32u + lane;
        v[k] = (r0 + k < r1 && i < n) ? col[i] : 0ull;
    }
    u64 sum = 0;
#pragma unroll
    for (u32 k = 0; k < 16u; ++k) sum += v[k];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    if (lane == 0) s_warp[warp] = sum;
    if (tid < 16) s_warp[16 + tid] = 0;
    __syncthreads();
    if (warp == 0) {
        const u64 w = s_warp[lane];
        u64 wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 up = __shfl_up_sync(0xFFFFFFFFu, wi, d);
            if (lane >= (u32)d) wi += up;
        }
        s_warp[lane] = wi - w;
    }
    __syncthreads();
    u64 run = s_warp[warp];
#pragma unroll
    for (u32 k = 0; k < 16u; ++k) {
        if (r0 + k < r1) {             // warp-uniform
            const u32 i = (r0 + k) * 32u + lane;
            u64 incl = v[k];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u64 up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= (u32)d) incl += up;
            }
            if (i < n) o[i] = run + incl - v[k];
            run += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
    }
    if (warp == 15 && lane == 0) o[n] = run;
}

}  // namespace pss
