// pss_scan.cuh — K2 + K3: exact multi-limb p^k, warp -> block -> grid (decoupled look-back) prefix scan
// of the running sums S_k(n), and the fused compare of EVERY partial sum against the target.
//
// Replaces (behaviour, not code):
//   primes_gpu ** power + cp.sum  (utils/gpu.py:169-232, int64 only, wraps silently)
//   sum(int(p) ** power for p in primes)  (utils/sequences.py:75, utils/gpu.py:244-247)
//   IncrementalSumCache.add_prime running sums  (utils/sum_cache.py:116-141)
//   ExpressionEvaluator._compare on each n     (utils/grammar.py:686-701,770-787)
//
// Arithmetic: 32-bit limbs, fixed widths per power (pow_limbs/sum_limbs in pss_common.cuh), schoolbook
// p^(e+1) = p^e * p as IMAD.WIDE.U32 + carry chains, no tensor cores (integer work, not a contraction).
// With MULTI the powers 1..K share one chain (config "--max p:5").
//
// Scan (warp -> block -> grid): pass A k_power_tiles forms every tile's exact sums (streaming, registers only),
// k_scan_columns (pss_bscan.cuh) scans them across tiles, pass B k_power_walk gives every thread its exact exclusive
// prefix (carry-in + tile prefix + block scan + warp scan) and forms / compares every S_k(n) of the tiles that can
// matter.  See the kernels below.
#pragma once
#include "pss_common.cuh"
#include "../../include/pss_b200.h"

namespace pss {

constexpr int kScanThreads = 256;

enum ScanMode { kModeReduce = 0, kModeCompare = 1, kModePrefix = 2, kModeTri = 3 };

template <int K, bool MULTI>
PSS_HD constexpr int lay_width(int j) { return sum_limbs(MULTI ? j + 1 : K); }
template <int K, bool MULTI>
PSS_HD constexpr int lay_offset(int j) {
    int o = 0;
    for (int i = 0; i < j; ++i) o += lay_width<K, MULTI>(i);
    return o;
}
template <int K, bool MULTI>
struct Layout {
    static constexpr int NP = MULTI ? K : 1;
    static constexpr int TOTAL = lay_offset<K, MULTI>(MULTI ? K : 1);
    PSS_HD static constexpr int width(int j) { return lay_width<K, MULTI>(j); }
    PSS_HD static constexpr int offset(int j) { return lay_offset<K, MULTI>(j); }
};

struct ScanParams {
    const u64* primes;     // element 0 of the range to scan
    u64 count;             // elements
    u64 n_first;           // global 1-based prime index of element 0
    u64 n_min, n_max;      // compare window (global indices, inclusive)
    u32 cmp_mask;          // bit0: S<T matches, bit1: S==T, bit2: S>T
    u32 target_over_mask;  // bit j: target >= 2^(32*width(j)) -> S_j < target always
    u32 n_tiles;
    u32 target[PSS_LIMBS];
    const u32* carry_in;   // NP * PSS_LIMBS limbs (device)
    u64* cols;             // pass A output: deferred-carry tile sums, column major (limb column c of tile t at cols[c * col_stride + t])
    const u64* col_prefix; // their exclusive prefixes over tiles (k_scan_columns), n_tiles + 1 entries per column
    u64 col_stride;
    u32 exhaustive;        // compare mode: 1 = walk every tile, 0 = skip tiles whose sum interval cannot contain the target
    pss_power_result* results;  // NP records (device)
    u32* prefix_out;       // kModePrefix: count * PSS_LIMBS limbs
    u32* error_flag;
};

// acc[OFF .. OFF+W) += b[0 .. LB)
template <int OFF, int W, int LB>
__device__ __forceinline__ void acc_add(u32* acc, const u32* b) {
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        u64 t = (u64)acc[OFF + i] + (i < LB ? (u64)b[i] : 0ull) + c;
        acc[OFF + i] = (u32)t;
        c = t >> 32;
    }
}
// r[0..W) = a[0..W) - b[0..LB)
template <int W, int LB>
__device__ __forceinline__ void limbs_sub(u32* r, const u32* a, const u32* b) {
    u32 borrow = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
        const u64 bi = (i < LB ? (u64)b[i] : 0ull) + borrow;
        const u64 ai = a[i];
        r[i] = (u32)(ai - bi);
        borrow = (ai < bi) ? 1u : 0u;
    }
}
template <int W>
__device__ __forceinline__ int limbs_cmp(const u32* a, const u32* t) {
    int res = 0;
#pragma unroll
    for (int i = 0; i < W; ++i)
        if (a[i] != t[i]) res = (a[i] < t[i]) ? -1 : 1;
    return res;
}

template <int J, class Lay, class Op>
__device__ __forceinline__ void for_each_power(Op& op) {
    if constexpr (J < Lay::NP) {
        op.template apply<Lay::offset(J), Lay::width(J), J>();
        for_each_power<J + 1, Lay>(op);
    }
}

struct OpShflUpAdd {   // a += shfl_up(a, delta) for lanes >= delta
    u32* a; int delta; bool active;
    template <int OFF, int W, int J>
    __device__ __forceinline__ void apply() {
        u32 o[W];
#pragma unroll
        for (int i = 0; i < W; ++i) o[i] = __shfl_up_sync(0xFFFFFFFFu, a[OFF + i], delta);
        if (active) acc_add<OFF, W, W>(a, o);
    }
};
struct OpShflXorAdd {  // butterfly reduction step
    u32* a; int delta;
    template <int OFF, int W, int J>
    __device__ __forceinline__ void apply() {
        u32 o[W];
#pragma unroll
        for (int i = 0; i < W; ++i) o[i] = __shfl_xor_sync(0xFFFFFFFFu, a[OFF + i], delta);
        acc_add<OFF, W, W>(a, o);
    }
};
struct OpAdd {         // a += b (same layout)
    u32* a; const u32* b;
    template <int OFF, int W, int J>
    __device__ __forceinline__ void apply() { acc_add<OFF, W, W>(a, b + OFF); }
};

struct OpCmpInit {     // three-way compare of every power's prefix with the target
    const u32* run; const u32* target; u32 over; int* prevc;
    template <int OFF, int W, int J>
    __device__ __forceinline__ void apply() {
        prevc[J] = ((over >> J) & 1u) ? -1 : limbs_cmp<W>(run + OFF, target);
    }
};

// ---- power chain ---------------------------------------------------------------------------------
template <int E, int K, bool MULTI, class F>
struct ChainStep {
    __device__ __forceinline__ static void run(const Big<pow_limbs(E)>& pe, u32 plo, u32 phi, F& f) {
        if constexpr (MULTI || E == K) f.template apply<E>(pe);
        if constexpr (E < K) {
            Big<pow_limbs(E + 1)> nx;
            big_mul_p<pow_limbs(E + 1), pow_limbs(E)>(nx, pe, plo, phi);
            ChainStep<E + 1, K, MULTI, F>::run(nx, plo, phi, f);
        }
    }
};
// a[0..3] += p^2 for p = plo + phi * 2^32 (phi < 2^8): multiply and accumulate fused in one carry chain
// (mad.lo.cc / madc.hi.cc pairs become IMAD.WIDE with carry-in/out; 9 instructions instead of ~19)
__device__ __forceinline__ void square_acc4(u32* a, u32 plo, u32 phi) {
    const u32 phi2 = phi + phi;
    asm("mad.lo.cc.u32  %0, %4, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %4, %4, %1;\n\t"
        "addc.cc.u32    %2, %2, 0;\n\t"
        "addc.u32       %3, %3, 0;\n\t"
        "mad.lo.cc.u32  %1, %4, %5, %1;\n\t"
        "madc.hi.cc.u32 %2, %4, %5, %2;\n\t"
        "addc.u32       %3, %3, 0;\n\t"
        "mad.lo.cc.u32  %2, %6, %6, %2;\n\t"
        "addc.u32       %3, %3, 0;\n\t"
        : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3])
        : "r"(plo), "r"(phi2), "r"(phi));
}

template <int K, bool MULTI, class F>
__device__ __forceinline__ void power_chain(u64 p, F& f) {
    if constexpr (K == 2 && !MULTI) {
        f.apply_square((u32)p, (u32)(p >> 32));   // the namesake case: sum of prime SQUARES
    } else {
        Big<2> p1;
        p1.v[0] = (u32)p;
        p1.v[1] = (u32)(p >> 32);
        ChainStep<1, K, MULTI, F>::run(p1, p1.v[0], p1.v[1], f);
    }
}

// same chain from a pre-split prime (plo, phi)
template <int K, bool MULTI, class F>
__device__ __forceinline__ void power_chain2(u32 plo, u32 phi, F& f) {
    if constexpr (K == 2 && !MULTI) {
        f.apply_square(plo, phi);
    } else {
        Big<2> p1;
        p1.v[0] = plo;
        p1.v[1] = phi;
        ChainStep<1, K, MULTI, F>::run(p1, plo, phi, f);
    }
}

template <int K, bool MULTI>
struct AccFn {
    using Lay = Layout<K, MULTI>;
    u32* acc;
    template <int E>
    __device__ __forceinline__ void apply(const Big<pow_limbs(E)>& pe) {
        constexpr int J = MULTI ? E - 1 : 0;
        acc_add<Lay::offset(J), Lay::width(J), pow_limbs(E)>(acc, pe.v);
    }
    __device__ __forceinline__ void apply_square(u32 plo, u32 phi) {
        static_assert(sum_limbs(2) == 4, "square_acc4 assumes a 4-limb accumulator");
        square_acc4(acc, plo, phi);
    }
};

template <int K, bool MULTI, int MODE, class PT = ScanParams>
struct Pass2Fn {
    using Lay = Layout<K, MULTI>;
    u32* run;
    const PT* P;
    int* prevc;
    u32* mcount;
    u32* mfirst;
    u32* mlast;
    u64 n_base;      // global index of local_idx 0
    u64 p, elem;
    u32 local_idx;
    bool in_window;
    template <int E>
    __device__ __forceinline__ void apply(const Big<pow_limbs(E)>& pe) {
        constexpr int J = MULTI ? E - 1 : 0;
        constexpr int OFF = Lay::offset(J), W = Lay::width(J);
        acc_add<OFF, W, pow_limbs(E)>(run, pe.v);
        post<E>(&pe);
    }
    __device__ __forceinline__ void apply_square(u32 plo, u32 phi) {
        square_acc4(run, plo, phi);
        post<2>(nullptr);
    }
    // run already holds S_E(n): record it (prefix mode) or compare it with the target
    template <int E>
    __device__ __forceinline__ void post(const Big<pow_limbs(E)>* pe) {
        constexpr int J = MULTI ? E - 1 : 0;
        constexpr int OFF = Lay::offset(J), W = Lay::width(J);
        if constexpr (MODE == kModePrefix) {
            u32* o = P->prefix_out + elem * PSS_LIMBS;
#pragma unroll
            for (int i = 0; i < PSS_LIMBS; ++i) o[i] = (i < W) ? run[OFF + i] : 0u;
        } else if constexpr (MODE == kModeTri) {
            // S == tri(m) for some m in [tri_m_min, tri_m_max]  <=>  8S+1 is a perfect square s^2 and
            // m = (s-1)/2 is in range (reference: utils/number_theory.py:292-354 qtri).  tri(m_max) < 2^59,
            // so only sums that fit 64 bits can match; the square root is taken in double and corrected.
            bool small = true;
#pragma unroll
            for (int i = 2; i < W; ++i) small = small && (run[OFF + i] == 0u);
            const u64 S = ((u64)run[OFF + 1] << 32) | run[OFF];
            if (in_window && small && S <= P->tri_max_value) {
                const u64 d = 8ull * S + 1ull;
                u64 r = (u64)sqrt((double)d);
                while (r * r > d) --r;
                while ((r + 1) * (r + 1) <= d) ++r;
                if (r * r == d) {
                    const u64 m = (r - 1) >> 1;
                    if (m >= P->tri_m_min && m <= P->tri_m_max) {
                        const u32 slot = atomicAdd(P->tri_count, 1u);
                        if (slot < P->tri_cap) {
                            P->tri_out[2 * slot] = n_base + local_idx;
                            P->tri_out[2 * slot + 1] = m;
                        }
                    }
                }
            }
        } else {
            const int c = ((P->target_over_mask >> J) & 1u) ? -1 : limbs_cmp<W>(run + OFF, P->target);
            if (in_window && ((P->cmp_mask >> (c + 1)) & 1u)) {
                mcount[J] += 1;
                mfirst[J] = min(mfirst[J], local_idx);
                mlast[J] = max(mlast[J], local_idx + 1);
            }
            if (prevc[J] < 0 && c >= 0) {   // the unique crossing of this power (S is strictly increasing)
                pss_power_result* r = P->results + J;
                r->cross_n = n_base + local_idx;
                r->cross_prime = p;
                Big<pow_limbs(E)> pw;
                if (pe) {
                    pw = *pe;
                } else {   // fused-square path: rebuild p^2 for the (once per query) bracket record
                    Big<2> p1;
                    p1.v[0] = (u32)p;
                    p1.v[1] = (u32)(p >> 32);
                    big_mul_p<pow_limbs(2), 2>(reinterpret_cast<Big<pow_limbs(2)>&>(pw), p1, p1.v[0], p1.v[1]);
                }
                u32 before[W];
                limbs_sub<W, pow_limbs(E)>(before, run + OFF, pw.v);
#pragma unroll
                for (int i = 0; i < PSS_LIMBS; ++i) {
                    r->sum_at_cross[i] = (i < W) ? run[OFF + i] : 0u;
                    r->sum_before_cross[i] = (i < W) ? before[i] : 0u;
                }
            }
            prevc[J] = c;
        }
    }
};

// ---- K2/K3 over materialised u64 primes: three launches, no CTA ever waits for another --------------------------------
//   k_power_tiles     pass A: every CTA streams tiles of T * ITEMS primes (coalesced 16-byte loads straight into
//                     registers: the order inside a tile is irrelevant for a sum), forms the tile's exact sums of p^k and
//                     stores them as deferred-carry u64 columns (col c of tile t at cols[c * col_stride + t]).  Pure
//                     streaming: 8 B per prime read once — the HBM-bound form of power_sum / slice sums.
//   k_scan_columns    (pss_bscan.cuh) exclusive u64 prefix of every column = the grid level of the scan; entry n_tiles of
//                     a column is the slice total.
//   k_power_walk      pass B (compare / prefix modes): tile prefix = carry-in + normalised column prefix; the tile is
//                     staged in blocked order, thread sums -> warp scan -> block scan give every thread its exact
//                     exclusive prefix, then every S_k(n) is formed and compared / written.  With interval pruning
//                     (default) a tile whose sum interval (S_before, S_after] cannot contain the target, that lies
//                     entirely inside or outside the window and does not hold n_max, compares like its prefix and is not
//                     even loaded (S_k is strictly increasing) — same argument and same results as the fused path.
// This replaces the single-pass decoupled look-back of round 1 (6.1 ms per 1e9 primes: one warp per CTA chased
// predecessor records while seven waited at a barrier).
template <int K, bool MULTI, int ITEMS>
__global__ void __launch_bounds__(kScanThreads) k_power_tiles(const __grid_constant__ ScanParams P) {
    using Lay = Layout<K, MULTI>;
    constexpr int TOTAL = Lay::TOTAL, T = kScanThreads, NWARP = T / 32, TILE = T * ITEMS;
    static_assert(ITEMS % 2 == 0, "16-byte loads");
    __shared__ u32 s_part[NWARP][TOTAL];
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (u32 tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
        const u64 base = (u64)tile * TILE;
        u32 acc[TOTAL];
#pragma unroll
        for (int i = 0; i < TOTAL; ++i) acc[i] = 0;
        AccFn<K, MULTI> fn{acc};
        const bool full = base + TILE <= P.count;
        if (full && ((reinterpret_cast<size_t>(P.primes + base) & 15u) == 0)) {
            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(P.primes + base);
            ulonglong2 v[ITEMS / 2];
#pragma unroll
            for (int r = 0; r < ITEMS / 2; ++r) v[r] = __ldg(src + r * T + tid);     // all loads in flight, then the arithmetic
#pragma unroll
            for (int r = 0; r < ITEMS / 2; ++r) {
                power_chain<K, MULTI>(v[r].x, fn);
                power_chain<K, MULTI>(v[r].y, fn);
            }
        } else {
#pragma unroll 4
            for (int r = 0; r < ITEMS; ++r) {
                const u64 g = base + (u64)r * T + tid;
                if (g < P.count) power_chain<K, MULTI>(__ldg(P.primes + g), fn);
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            OpShflXorAdd op{acc, d};
            for_each_power<0, Lay>(op);
        }
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < TOTAL; ++i) s_part[warp][i] = acc[i];
        }
        __syncthreads();
        if (tid < (u32)TOTAL) {
            u64 sum = 0;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) sum += s_part[w][tid];
            P.cols[(u64)tid * P.col_stride + tile] = sum;
        }
        __syncthreads();
    }
}

template <int K, bool MULTI, int MODE, int ITEMS>
__global__ void __launch_bounds__(kScanThreads) k_power_walk(const __grid_constant__ ScanParams P) {
    using Lay = Layout<K, MULTI>;
    constexpr int TOTAL = Lay::TOTAL, NP = Lay::NP, T = kScanThreads, NWARP = T / 32, TILE = T * ITEMS;
    static_assert(NWARP <= 32 && (NWARP & (NWARP - 1)) == 0, "warp count must be a power of two");
    constexpr int kListCap = 64;           // tiles of one batch that need a walk (a query has a handful in total)

    __shared__ u64 s_p[TILE + T];          // element e lives at e + e / ITEMS (bank-conflict padding)
    __shared__ u32 s_warp[NWARP][TOTAL];   // inclusive warp totals, then exclusive warp offsets
    __shared__ u32 s_excl[TOTAL];          // tile exclusive prefix (includes carry-in)
    __shared__ u32 s_list[kListCap];
    __shared__ u32 s_nlist;

    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    // normalised prefix (carry-in + deferred-carry column prefix) of power jj before tile `idx` (idx = n_tiles: slice end)
    auto prefix_of = [&](u32 idx, int jj, u32* out) {
        u64 c = 0;
        const int off = lay_offset<K, MULTI>(jj), w = lay_width<K, MULTI>(jj);
        for (int i = 0; i < w; ++i) {
            const u64 v = P.col_prefix[(u64)(off + i) * P.col_stride + idx];
            const u64 lo = (v & 0xFFFFFFFFull) + (c & 0xFFFFFFFFull) + P.carry_in[jj * PSS_LIMBS + i];
            out[i] = (u32)lo;
            c = (v >> 32) + (c >> 32) + (lo >> 32);
        }
    };
    auto cmp_target = [&](const u32* a, int jj) -> int {
        if ((P.target_over_mask >> jj) & 1u) return -1;
        int r = 0;
        const int w = lay_width<K, MULTI>(jj);
        for (int i = 0; i < w; ++i)
            if (a[i] != P.target[i]) r = (a[i] < P.target[i]) ? -1 : 1;
        return r;
    };

    // ---- the full walk of one tile by the whole CTA
    auto walk_tile = [&](u32 tile) {
        const u64 base = (u64)tile * TILE;
        const u64 t_n0 = P.n_first + base;               // global index of the tile's first prime
        if (tid < (u32)NP) prefix_of(tile, (int)tid, s_excl + lay_offset<K, MULTI>((int)tid));
        // coalesced load -> padded shared memory (blocked order for the owner threads)
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const u32 e = r * T + tid;
            const u64 g = base + e;
            s_p[e + e / ITEMS] = (g < P.count) ? P.primes[g] : 0ull;
        }
        __syncthreads();

        // pass 1: thread-local sums of p^k
        u32 acc[TOTAL];
#pragma unroll
        for (int i = 0; i < TOTAL; ++i) acc[i] = 0;
        {
            AccFn<K, MULTI> fn{acc};
#pragma unroll 2
            for (int i = 0; i < ITEMS; ++i) power_chain<K, MULTI>(s_p[tid * (ITEMS + 1) + i], fn);
        }
        // warp inclusive scan (multi-limb adds with carry), exclusive value for this lane
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            OpShflUpAdd op{acc, d, lane >= (u32)d};
            for_each_power<0, Lay>(op);
        }
        if (lane == 31) {
#pragma unroll
            for (int i = 0; i < TOTAL; ++i) s_warp[warp][i] = acc[i];
        }
        u32 run[TOTAL];   // becomes this thread's exclusive prefix
#pragma unroll
        for (int i = 0; i < TOTAL; ++i) {
            const u32 up = __shfl_up_sync(0xFFFFFFFFu, acc[i], 1);
            run[i] = lane ? up : 0u;
        }
        __syncthreads();
        if (warp == 0) {   // scan of the warp totals -> exclusive warp offsets
            u32 wi[TOTAL];
#pragma unroll
            for (int i = 0; i < TOTAL; ++i) wi[i] = (lane < NWARP) ? s_warp[lane][i] : 0u;
#pragma unroll
            for (int d = 1; d < NWARP; d <<= 1) {
                OpShflUpAdd op{wi, d, lane >= (u32)d};
                for_each_power<0, Lay>(op);
            }
#pragma unroll
            for (int i = 0; i < TOTAL; ++i) {
                const u32 up = __shfl_up_sync(0xFFFFFFFFu, wi[i], 1);
                if (lane < NWARP) s_warp[lane][i] = lane ? up : 0u;
            }
        }
        __syncthreads();

        // pass 2: exact exclusive prefix of this thread, then every S_k(n) formed and compared / written
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
        if constexpr (MODE == kModeCompare) {
            OpCmpInit op{run, P.target, P.target_over_mask, prevc};
            for_each_power<0, Lay>(op);
        }
        Pass2Fn<K, MULTI, MODE> fn{run, &P, prevc, mcount, mfirst, mlast, t_n0, 0, 0, 0, false};
#pragma unroll 2
        for (int i = 0; i < ITEMS; ++i) {
            const u64 p = s_p[tid * (ITEMS + 1) + i];
            if (p == 0) continue;   // padding past the end of the range
            const u32 li = tid * ITEMS + i;
            fn.local_idx = li;
            fn.elem = base + li;
            fn.p = p;
            const u64 n = t_n0 + li;
            fn.in_window = (n >= P.n_min) && (n <= P.n_max);
            power_chain<K, MULTI>(p, fn);
        }
        if constexpr (MODE == kModeCompare) {
#pragma unroll
            for (int jj = 0; jj < NP; ++jj) {
                if (__any_sync(0xFFFFFFFFu, mcount[jj] != 0)) {
                    const u32 c = __reduce_add_sync(0xFFFFFFFFu, mcount[jj]);
                    const u32 f = __reduce_min_sync(0xFFFFFFFFu, mfirst[jj]);
                    const u32 l = __reduce_max_sync(0xFFFFFFFFu, mlast[jj]);
                    if (lane == 0) {
                        pss_power_result* r = P.results + jj;
                        atomicAdd(reinterpret_cast<unsigned long long*>(&r->match_count), (unsigned long long)c);
                        atomicMin(reinterpret_cast<unsigned long long*>(&r->first_match_n), (unsigned long long)(t_n0 + f));
                        atomicMax(reinterpret_cast<unsigned long long*>(&r->last_match_n), (unsigned long long)(t_n0 + l - 1));
                    }
                }
            }
        }
        __syncthreads();   // shared memory is reused by the next tile
    };

    if (MODE != kModeCompare || P.exhaustive) {
        for (u32 tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) walk_tile(tile);
        return;
    }

    // ---- compare mode with interval pruning.  S_k is strictly increasing, so a tile whose sum interval (S_before, S_after]
    // does not contain the target for any power, and that lies entirely inside (or outside) the window, compares like its
    // prefix: its matches are known without loading a prime.  The test needs only the scanned columns, so EVERY THREAD
    // tests one tile (256 tiles per CTA step, all loads independent); the few tiles that do need a walk — the one bracketing
    // the target, the ones cut by the window — are listed in shared memory and walked by the whole CTA afterwards.
    for (u32 batch = blockIdx.x * T; batch < P.n_tiles; batch += gridDim.x * T) {
        if (tid == 0) s_nlist = 0;
        __syncthreads();
        const u32 tile = batch + tid;
        u64 mc[NP], mf[NP], ml[NP];
#pragma unroll
        for (int jj = 0; jj < NP; ++jj) mc[jj] = 0, mf[jj] = ~0ull, ml[jj] = 0;
        if (tile < P.n_tiles) {
            const u64 base = (u64)tile * TILE;
            const u32 tcnt = (u32)min((u64)TILE, P.count - base);
            const u64 t_n0 = P.n_first + base, t_last = t_n0 + tcnt - 1;
            const bool inside = t_n0 >= P.n_min && t_last <= P.n_max, outside = t_last < P.n_min || t_n0 > P.n_max;
            bool need_walk = !(inside || outside);
            int c0[NP];
#pragma unroll
            for (int jj = 0; jj < NP; ++jj) {
                u32 a[PSS_LIMBS];
                prefix_of(tile, jj, a);
                c0[jj] = cmp_target(a, jj);
                prefix_of(tile + 1, jj, a);
                need_walk = need_walk || (c0[jj] < 0 && cmp_target(a, jj) >= 0);
            }
            if (need_walk) {
                const u32 slot = atomicAdd(&s_nlist, 1u);
                if (slot < (u32)kListCap) s_list[slot] = tile;
            } else if (inside) {
#pragma unroll
                for (int jj = 0; jj < NP; ++jj) {
                    // every S of the tile compares like the tile's prefix ("==" cannot occur inside it)
                    const bool hit = (c0[jj] < 0) ? (P.cmp_mask & 1u) : (P.cmp_mask & 4u);
                    if (hit) mc[jj] = tcnt, mf[jj] = t_n0, ml[jj] = t_last;
                }
            }
        }
#pragma unroll
        for (int jj = 0; jj < NP; ++jj) {
            if (__any_sync(0xFFFFFFFFu, mc[jj] != 0)) {
                u64 c = mc[jj], f = mf[jj], l = ml[jj];
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    c += __shfl_xor_sync(0xFFFFFFFFu, c, d);
                    f = min(f, __shfl_xor_sync(0xFFFFFFFFu, f, d));
                    l = max(l, __shfl_xor_sync(0xFFFFFFFFu, l, d));
                }
                if (lane == 0) {
                    pss_power_result* r = P.results + jj;
                    atomicAdd(reinterpret_cast<unsigned long long*>(&r->match_count), (unsigned long long)c);
                    atomicMin(reinterpret_cast<unsigned long long*>(&r->first_match_n), (unsigned long long)f);
                    atomicMax(reinterpret_cast<unsigned long long*>(&r->last_match_n), (unsigned long long)l);
                }
            }
        }
        __syncthreads();
        const u32 n_list = s_nlist;
        if (n_list > (u32)kListCap) {      // cannot happen for a monotone S (at most 2 + NP tiles of a query need a walk): fail loudly
            if (tid == 0) atomicExch(P.error_flag, 4u);
        } else {
            for (u32 i = 0; i < n_list; ++i) walk_tile(s_list[i]);
        }
        __syncthreads();
    }
}

}  // namespace pss
