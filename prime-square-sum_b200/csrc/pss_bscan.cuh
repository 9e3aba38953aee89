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
    u64 n_min, n_max;
    u32 cmp_mask, target_over_mask;
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

// Decode loop: one prime per iteration for every lane that still has one (lanes own equal word counts, so
// the trip counts differ only by the local density fluctuation).  The refill of an exhausted word is
// branch-free (unconditional shared-memory load + selects) and the body is a plain if-region, so the warp
// reconverges every iteration and never splits around the power chain.
// f(plo, phi) -> false stops this lane (past n_max).
template <class F>
__device__ __forceinline__ void walk_words(const u32* my_words, u32 base_lo, u32 base_hi, F& f) {
    u32 wi = 0, w = 0, woff = 0;
    bool alive = true;
    while (alive) {
        const bool need = (w == 0);
        const u32 idx = min(wi, (u32)(kWordsPerThread - 1));
        const u32 nw = my_words[idx];
        w = need ? ((wi < (u32)kWordsPerThread) ? nw : 0u) : w;
        woff = need ? idx * 120u : woff;
        wi += need ? 1u : 0u;
        if (w != 0) {
            const u32 m = 31u - (u32)__clz(w);
            w ^= (1u << m);
            const u32 off = woff + msb_value_offset(m);
            const u32 plo = base_lo + off;
            const u32 phi = base_hi + (plo < off ? 1u : 0u);
            alive = f(plo, phi);
        } else {
            alive = (wi <= (u32)kWordsPerThread);   // empty word (rare): try the next one; else done
        }
    }
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

template <int K, bool MULTI, int PMODE>
struct BsCmpFn {
    using Lay = Layout<K, MULTI>;
    Pass2Fn<K, MULTI, PMODE, BScanParams> inner;
    const BScanParams* P;
    u64 tile_n0;     // global index of the tile's first prime
    u32 li;          // tile-local index of the next prime
    u32 lmin, lmax;  // window [n_min, n_max] in tile-local indices (saturated)
    __device__ __forceinline__ bool operator()(u32 plo, u32 phi) {
        if (li > lmax) return false;
        inner.p = ((u64)phi << 32) | plo;
        inner.local_idx = li;
        inner.in_window = (li >= lmin);
        power_chain2<K, MULTI>(plo, phi, inner);
        if (li == lmax && tile_n0 + li == P->n_max) {   // exactly one thread in the grid: S(n_max), p_{n_max}
            *P->last_prime_out = inner.p;
#pragma unroll
            for (int jj = 0; jj < Lay::NP; ++jj)
#pragma unroll
                for (int i = 0; i < PSS_LIMBS; ++i)
                    P->results[jj].final_sum[i] = (i < Lay::width(jj)) ? inner.run[Lay::offset(jj) + i] : 0u;
        }
        ++li;
        return true;
    }
};

// Warp-autonomous kernel: a warp owns one "warp tile" = 32 threads x 24 words (768 words, 92 160 integers);
// 8 consecutive warp tiles form a scan tile (the granularity of the column prefix).  No __syncthreads.
template <int K, bool MULTI, int MODE>
__global__ void __launch_bounds__(kBsThreads) k_bitmap_scan(const __grid_constant__ BScanParams P) {
    using Lay = Layout<K, MULTI>;
    constexpr int TOTAL = Lay::TOTAL, NP = Lay::NP, T = kBsThreads, NWARP = T / 32;
    constexpr int WPT = kWordsPerThread, PAD = WPT + 1, WARP_WORDS = 32 * WPT;

    __shared__ u32 s_words[T * PAD];

    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    u32* my_warp_words = s_words + warp * 32 * PAD;
    const u32* my_words = my_warp_words + lane * PAD;
    const u64 n_warp_tiles = (u64)P.n_tiles * NWARP;
    const u64 warp_stride = (u64)gridDim.x * NWARP;

    for (u64 wt = (u64)blockIdx.x * NWARP + warp; wt < n_warp_tiles; wt += warp_stride) {
        const u32 tile = (u32)(wt / NWARP);
        const u32 wit = (u32)(wt % NWARP);          // warp index inside the scan tile
        if constexpr (MODE != kBsReduce) {
            if (P.n_first + P.col_prefix[tile] > P.n_max) break;   // prefixes are monotone in tile
        }
        const u64 word0 = wt * WARP_WORDS;
        // ---- warp-local staging: coalesced uint4 loads -> padded blocked layout, bit-reversed
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

        const u64 value_base = (P.B_start + 4ull * (word0 + (u64)lane * WPT)) * 30ull;
        const u32 base_lo = (u32)value_base, base_hi = (u32)(value_base >> 32);
        const bool has_small = (wt == 0 && lane == 0 && P.n_small > 0);
        const u64 gthread = wt * 32 + lane;

        u32 acc[TOTAL];
        u32 cnt = 0;
        if constexpr (MODE == kBsReduce) {
            // ---- thread-local count and sums
#pragma unroll
            for (int i = 0; i < TOTAL; ++i) acc[i] = 0;
#pragma unroll
            for (int i = 0; i < WPT; ++i) cnt += __popc(my_words[i]);
            BsAccFn<K, MULTI> fn{acc};
            if (has_small) {
                for (u32 i = 0; i < P.n_small; ++i) fn((u32)P.small_primes[i], 0u);
                cnt += P.n_small;
            }
            walk_words(my_words, base_lo, base_hi, fn);
            P.thread_rec[gthread] = cnt;
#pragma unroll
            for (int i = 0; i < TOTAL; ++i) P.thread_rec[(u64)(1 + i) * P.thread_stride + gthread] = acc[i];
            // ---- warp totals (exact limbs) -> warp record; deferred-carry u64 atomics -> tile record
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
                    atomicAdd(reinterpret_cast<unsigned long long*>(P.cols + (u64)(1 + i) * P.col_stride + tile),
                              (unsigned long long)acc[i]);
                }
            }
        } else {
            // ---- exclusive prefix of this thread = carry-in + tile prefix + earlier warps + earlier lanes
            cnt = __ldg(P.thread_rec + gthread);
#pragma unroll
            for (int i = 0; i < TOTAL; ++i) acc[i] = __ldg(P.thread_rec + (u64)(1 + i) * P.thread_stride + gthread);
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
            u32 li0 = cinc - cnt;
            // earlier warps of this scan tile: lane l < wit fetches warp record l, butterfly sum
            {
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
                c = __shfl_sync(0xFFFFFFFFu, c, 0);
                OpAdd op{run, v};
                for_each_power<0, Lay>(op);
                li0 += c;
            }
            // tile prefix = carry-in + normalised deferred-carry column prefixes (uniform across the warp)
            {
                u32 ex[TOTAL];
#pragma unroll
                for (int jj = 0; jj < NP; ++jj) {
                    u64 c = 0;
#pragma unroll
                    for (int i = 0; i < PSS_LIMBS; ++i) {
                        if (i < Lay::width(jj)) {
                            const int col = 1 + Lay::offset(jj) + i;
                            const u64 v = P.col_prefix[(u64)col * P.col_stride + tile];
                            const u64 lo = (v & 0xFFFFFFFFull) + (c & 0xFFFFFFFFull) + P.carry_in[jj * PSS_LIMBS + i];
                            ex[Lay::offset(jj) + i] = (u32)lo;
                            c = (v >> 32) + (c >> 32) + (lo >> 32);
                        }
                    }
                }
                OpAdd op{run, ex};
                for_each_power<0, Lay>(op);
            }
            const u64 tile_n0 = P.n_first + P.col_prefix[tile];
            // ---- every S_k(n) formed and compared
            int prevc[NP];
            u32 mcount[NP], mfirst[NP], mlast[NP];
#pragma unroll
            for (int jj = 0; jj < NP; ++jj) {
                mcount[jj] = 0;
                mfirst[jj] = 0xFFFFFFFFu;
                mlast[jj] = 0;
                prevc[jj] = 0;
            }
            if constexpr (MODE == kBsCompare) {
                OpCmpInit op{run, P.target, P.target_over_mask, prevc};
                for_each_power<0, Lay>(op);
            }
            // window in tile-local indices: a tile holds < 2^20 primes, so saturating to u32 is exact
            const u32 lmin = (P.n_min > tile_n0) ? (u32)min(P.n_min - tile_n0, (u64)0xFFFFFFFFull) : 0u;
            const u32 lmax = (u32)min(P.n_max - tile_n0, (u64)0xFFFFFFFEull);   // n_max >= tile_n0 (checked above)
            constexpr int PMODE = (MODE == kBsTri) ? kModeTri : kModeCompare;
            BsCmpFn<K, MULTI, PMODE> fn{{run, &P, prevc, mcount, mfirst, mlast, tile_n0, 0, 0, 0, false}, &P, tile_n0, li0, lmin, lmax};
            bool go = true;
            if (has_small)
                for (u32 i = 0; i < P.n_small && go; ++i) go = fn((u32)P.small_primes[i], 0u);
            if (go) walk_words(my_words, base_lo, base_hi, fn);
            __syncwarp();
#pragma unroll
            for (int jj = 0; jj < NP; ++jj) {
                if (__any_sync(0xFFFFFFFFu, mcount[jj] != 0)) {
                    const u32 c = __reduce_add_sync(0xFFFFFFFFu, mcount[jj]);
                    const u32 f = __reduce_min_sync(0xFFFFFFFFu, mfirst[jj]);
                    const u32 l = __reduce_max_sync(0xFFFFFFFFu, mlast[jj]);
                    if (lane == 0) {
                        pss_power_result* r = P.results + jj;
                        atomicAdd(reinterpret_cast<unsigned long long*>(&r->match_count), (unsigned long long)c);
                        atomicMin(reinterpret_cast<unsigned long long*>(&r->first_match_n), (unsigned long long)(tile_n0 + f));
                        atomicMax(reinterpret_cast<unsigned long long*>(&r->last_match_n), (unsigned long long)(tile_n0 + l - 1));
                    }
                }
            }
        }
    }
}

// ---- grid level of the scan: exclusive u64 prefix of every (deferred-carry) column, one CTA per column --------------
// out[c * stride + t] = sum_{i<t} in[c * stride + i]  for t = 0..n  (entry n = column total)
__global__ void __launch_bounds__(1024) k_scan_columns(const u64* in, u64* out, u32 n, u64 stride) {
    __shared__ u64 s_warp[32];
    __shared__ u64 s_running;
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u64* col = in + (u64)blockIdx.x * stride;
    u64* o = out + (u64)blockIdx.x * stride;
    if (tid == 0) s_running = 0;
    __syncthreads();
    for (u32 base = 0; base < n; base += 1024) {
        const u32 i = base + tid;
        const u64 v = (i < n) ? col[i] : 0;
        u64 incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (u32)d) incl += up;
        }
        if (lane == 31) s_warp[warp] = incl;
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
        const u64 excl = s_running + s_warp[warp] + incl - v;
        if (i < n) o[i] = excl;
        __syncthreads();
        if (tid == 1023) s_running = excl + v;
        __syncthreads();
    }
    if (tid == 0) o[n] = s_running;
}

}  // namespace pss
