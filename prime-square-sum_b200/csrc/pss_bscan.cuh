// pss_bscan.cuh — the fused search path: K2 + K3 reading the sieve's bit-packed output directly.
//
// The wheel-30 bitmap IS the compact prime buffer (0.76 B / prime at 2.28e10 instead of 8 B / prime), so the
// query path never materialises u64 primes: each thread owns kWordsPerThread consecutive bitmap words
// (blocked order, staged through padded shared memory from coalesced uint4 loads), decodes its primes with
// ffs, raises them to the k-th power with the same IMAD carry chains as pss_scan.cuh and
//   kBsReduce   block-reduces (count, S_k limbs) per tile  -> column-major u32 tile records
//   k_scan_columns  exclusive u64 prefix of every column (count, each 32-bit limb as a deferred-carry
//                   64-bit lane) across tiles — the grid level of the warp -> block -> grid scan
//   kBsCompare  re-walks the tile from its exact prefix (carry-in + normalised column prefixes + block
//               scan), forms EVERY partial sum S_k(n) and compares it with the target limbs.
// No CTA ever waits for another CTA (three launches instead of a look-back chain), so the multi-GPU
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
enum BsMode { kBsReduce = 0, kBsCompare = 1 };

struct BScanParams {
    const u32* bitmap;       // word 0 = absolute byte index B_start
    u64 n_words;
    u64 B_start;
    u32 n_tiles;
    u32 n_small;             // 2, 3, 5 inside [lo, hi): handled by thread 0 of tile 0 (not representable)
    u64 small_primes[3];
    // tile records, column major: col c of tile t at cols[c * col_stride + t]; col 0 = count, 1.. = limbs
    u32* cols;
    const u64* col_prefix;   // exclusive prefixes, same layout, (n_tiles + 1) entries per column
    u64 col_stride;
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
// branch-free (unconditional shared-memory load + selects) so the warp never splits around the power chain.
template <class F>
__device__ __forceinline__ void walk_words(const u32* my_words, u64 value_base, F& f) {
    u32 wi = 0, w = 0, woff = 0;
    for (;;) {
        const bool need = (w == 0);
        const u32 idx = min(wi, (u32)(kWordsPerThread - 1));
        const u32 nw = my_words[idx];
        w = need ? ((wi < (u32)kWordsPerThread) ? nw : 0u) : w;
        woff = need ? idx * 120u : woff;
        wi += need ? 1u : 0u;
        if (w == 0) {                       // empty word (rare) or end of this thread's range
            if (wi > (u32)kWordsPerThread) return;
            continue;
        }
        const u32 m = 31u - (u32)__clz(w);
        w ^= (1u << m);
        if (!f(value_base + (woff + msb_value_offset(m)))) return;
    }
}

template <int K, bool MULTI>
struct BsAccFn {
    u32* acc;
    __device__ __forceinline__ bool operator()(u64 p) {
        AccFn<K, MULTI> fn{acc};
        power_chain<K, MULTI>(p, fn);
        return true;
    }
};

template <int K, bool MULTI>
struct BsCmpFn {
    using Lay = Layout<K, MULTI>;
    Pass2Fn<K, MULTI, kModeCompare, BScanParams> inner;
    const BScanParams* P;
    u64 tile_n0;     // global index of the tile's first prime
    u32 li;          // tile-local index of the next prime
    u32 lmin, lmax;  // window [n_min, n_max] in tile-local indices (saturated)
    __device__ __forceinline__ bool operator()(u64 p) {
        if (li > lmax) return false;
        inner.p = p;
        inner.local_idx = li;
        inner.in_window = (li >= lmin);
        power_chain<K, MULTI>(p, inner);
        if (li == lmax && tile_n0 + li == P->n_max) {   // exactly one thread in the grid: S(n_max), p_{n_max}
            *P->last_prime_out = p;
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

template <int K, bool MULTI, int MODE>
__global__ void __launch_bounds__(kBsThreads) k_bitmap_scan(const __grid_constant__ BScanParams P) {
    using Lay = Layout<K, MULTI>;
    constexpr int TOTAL = Lay::TOTAL, NP = Lay::NP, T = kBsThreads, NWARP = T / 32;
    constexpr int WPT = kWordsPerThread, PAD = WPT + 1;

    __shared__ u32 s_words[T * PAD];
    __shared__ u32 s_warp[NWARP][TOTAL + 1];   // [..][TOTAL] = count
    __shared__ u32 s_excl[TOTAL];
    __shared__ u64 s_tile_n0;

    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    for (u32 tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        if constexpr (MODE == kBsCompare) {
            // tiles that start beyond n_max have nothing to compare
            if (P.n_first + P.col_prefix[tile] > P.n_max) break;   // prefixes are monotone in tile
        }
        const u64 word0 = (u64)tile * kBsTileWords;
        // ---- coalesced uint4 loads -> padded blocked layout (word w of thread t at t*PAD + w)
        const uint4* g4 = reinterpret_cast<const uint4*>(P.bitmap + word0);
#pragma unroll
        for (int r = 0; r < WPT / 4; ++r) {
            const u32 q = r * T + tid;            // uint4 index inside the tile
            const u64 wq = word0 + 4ull * q;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (wq + 4 <= P.n_words) v = __ldg(g4 + q);
            const u32 wl = 4u * q;                // first word (tile local); 4 | WPT so one owner thread
            const u32 dst = (wl / WPT) * PAD + (wl % WPT);
            s_words[dst] = __brev(v.x); s_words[dst + 1] = __brev(v.y);
            s_words[dst + 2] = __brev(v.z); s_words[dst + 3] = __brev(v.w);
        }
        __syncthreads();

        const u32* my_words = s_words + tid * PAD;
        const u64 value_base = (P.B_start + 4ull * (word0 + (u64)tid * WPT)) * 30ull;
        const bool has_small = (tile == 0 && tid == 0 && P.n_small > 0);

        // ---- pass 1: thread-local count and sums (reduce pass: decode + powers; compare pass: reload the
        //      reduce pass's per-thread record)
        u32 acc[TOTAL];
        u32 cnt = 0;
        const u64 gthread = (u64)tile * T + tid;
        if constexpr (MODE == kBsReduce) {
#pragma unroll
            for (int i = 0; i < TOTAL; ++i) acc[i] = 0;
#pragma unroll
            for (int i = 0; i < WPT; ++i) cnt += __popc(my_words[i]);
            BsAccFn<K, MULTI> fn{acc};
            if (has_small) {
                for (u32 i = 0; i < P.n_small; ++i) fn(P.small_primes[i]);
                cnt += P.n_small;
            }
            walk_words(my_words, value_base, fn);
            P.thread_rec[gthread] = cnt;
#pragma unroll
            for (int i = 0; i < TOTAL; ++i) P.thread_rec[(u64)(1 + i) * P.thread_stride + gthread] = acc[i];
        } else {
            cnt = __ldg(P.thread_rec + gthread);
#pragma unroll
            for (int i = 0; i < TOTAL; ++i) acc[i] = __ldg(P.thread_rec + (u64)(1 + i) * P.thread_stride + gthread);
        }

        if constexpr (MODE == kBsReduce) {
            // ---- block reduce -> one column-major record per tile
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) {
                OpShflXorAdd op{acc, d};
                for_each_power<0, Lay>(op);
            }
            cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < TOTAL; ++i) s_warp[warp][i] = acc[i];
                s_warp[warp][TOTAL] = cnt;
            }
            __syncthreads();
            if (warp == 0) {
                u32 v[TOTAL];
#pragma unroll
                for (int i = 0; i < TOTAL; ++i) v[i] = (lane < NWARP) ? s_warp[lane][i] : 0u;
                u32 c = (lane < NWARP) ? s_warp[lane][TOTAL] : 0u;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    OpShflXorAdd op{v, d};
                    for_each_power<0, Lay>(op);
                }
                c = __reduce_add_sync(0xFFFFFFFFu, c);
                if (lane == 0) {
                    P.cols[tile] = c;
#pragma unroll
                    for (int i = 0; i < TOTAL; ++i) P.cols[(u64)(1 + i) * P.col_stride + tile] = v[i];
                }
            }
            __syncthreads();
        } else {
            // ---- warp inclusive scan of (count, sums); exclusive values for this lane
            u32 cinc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                OpShflUpAdd op{acc, d, lane >= (u32)d};
                for_each_power<0, Lay>(op);
                const u32 o = __shfl_up_sync(0xFFFFFFFFu, cinc, d);
                if (lane >= (u32)d) cinc += o;
            }
            if (lane == 31) {
#pragma unroll
                for (int i = 0; i < TOTAL; ++i) s_warp[warp][i] = acc[i];
                s_warp[warp][TOTAL] = cinc;
            }
            u32 run[TOTAL];
#pragma unroll
            for (int i = 0; i < TOTAL; ++i) {
                const u32 up = __shfl_up_sync(0xFFFFFFFFu, acc[i], 1);
                run[i] = lane ? up : 0u;
            }
            u32 cexc = cinc - cnt;
            __syncthreads();
            if (warp == 0) {
                // exclusive scan of the warp totals
                u32 wi_[TOTAL];
#pragma unroll
                for (int i = 0; i < TOTAL; ++i) wi_[i] = (lane < NWARP) ? s_warp[lane][i] : 0u;
                u32 wc = (lane < NWARP) ? s_warp[lane][TOTAL] : 0u;
                const u32 wc0 = wc;
#pragma unroll
                for (int d = 1; d < NWARP; d <<= 1) {
                    OpShflUpAdd op{wi_, d, lane >= (u32)d};
                    for_each_power<0, Lay>(op);
                    const u32 o = __shfl_up_sync(0xFFFFFFFFu, wc, d);
                    if (lane >= (u32)d) wc += o;
                }
#pragma unroll
                for (int i = 0; i < TOTAL; ++i) {
                    const u32 up = __shfl_up_sync(0xFFFFFFFFu, wi_[i], 1);
                    if (lane < NWARP) s_warp[lane][i] = lane ? up : 0u;
                }
                if (lane < NWARP) s_warp[lane][TOTAL] = wc - wc0;
                // tile prefix = carry-in + normalised deferred-carry column prefixes
                if (lane == 0) {
                    u32 ex[TOTAL];
#pragma unroll
                    for (int jj = 0; jj < NP; ++jj) {
                        u64 c = 0;
#pragma unroll
                        for (int i = 0; i < PSS_LIMBS; ++i) {
                            if (i < Lay::width(jj)) {
                                const int col = 1 + Lay::offset(jj) + i;
                                const u64 v = P.col_prefix[(u64)col * P.col_stride + tile];
                                // v < 2^63, carry < 2^32: split the add to stay inside 64 bits
                                const u64 lo = (v & 0xFFFFFFFFull) + (c & 0xFFFFFFFFull) + P.carry_in[jj * PSS_LIMBS + i];
                                ex[Lay::offset(jj) + i] = (u32)lo;
                                c = (v >> 32) + (c >> 32) + (lo >> 32);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < TOTAL; ++i) s_excl[i] = ex[i];
                    s_tile_n0 = P.n_first + P.col_prefix[tile];
                }
            }
            __syncthreads();
            // ---- pass 2: exact exclusive prefix of this thread, every S_k(n) formed and compared
            {
                OpAdd op1{run, s_excl};
                for_each_power<0, Lay>(op1);
                OpAdd op2{run, s_warp[warp]};
                for_each_power<0, Lay>(op2);
            }
            int prevc[NP];
            u32 mcount[NP], mfirst[NP], mlast[NP];
#pragma unroll
            for (int jj = 0; jj < NP; ++jj) {
                mcount[jj] = 0;
                mfirst[jj] = 0xFFFFFFFFu;
                mlast[jj] = 0;
                prevc[jj] = 0;
            }
            {
                OpCmpInit op{run, P.target, P.target_over_mask, prevc};
                for_each_power<0, Lay>(op);
            }
            const u64 tile_n0 = s_tile_n0;
            // window in tile-local indices: a tile holds < 2^20 primes, so saturating to u32 is exact
            const u32 lmin = (P.n_min > tile_n0) ? (u32)min(P.n_min - tile_n0, (u64)0xFFFFFFFFull) : 0u;
            const u32 lmax = (u32)min(P.n_max - tile_n0, (u64)0xFFFFFFFEull);   // n_max >= tile_n0 (checked above)
            BsCmpFn<K, MULTI> fn{{run, &P, prevc, mcount, mfirst, mlast, tile_n0, 0, 0, 0, false}, &P, tile_n0,
                                 s_warp[warp][TOTAL] + cexc, lmin, lmax};
            bool go = true;
            if (has_small)
                for (u32 i = 0; i < P.n_small && go; ++i) go = fn(P.small_primes[i]);
            if (go) walk_words(my_words, value_base, fn);
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
            __syncthreads();
        }
    }
}

// ---- grid level of the scan: exclusive u64 prefix of every u32 column, one CTA per column --------------
// out[c * stride + t] = sum_{i<t} in[c * stride + i]  for t = 0..n  (entry n = column total)
__global__ void __launch_bounds__(1024) k_scan_columns(const u32* in, u64* out, u32 n, u64 stride) {
    __shared__ u64 s_warp[32];
    __shared__ u64 s_running;
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u32* col = in + (u64)blockIdx.x * stride;
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
