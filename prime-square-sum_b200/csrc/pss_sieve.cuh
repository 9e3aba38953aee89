// pss_sieve.cuh — K1: segmented, bit-packed, wheel-30 sieve of Eratosthenes for sm_100a.
//
// Replaces (behaviour, not code) primesieve.* / _basic_sieve / _segmented_sieve of the reference
// (utils/sieve.py:224-327 and the primesieve call sites :499-639): "all primes in [lo, hi)".
//
//   k_build_pattern   once per handle: pre-sieved periodic byte patterns for the tiny primes
//                     {7,11,13,17,19} {23,29,31} {37,41,43} {47,53,59}; 16x replicated so every tile
//                     start (16-byte aligned) maps to a 16-byte aligned pattern offset (L2 resident).
//   k_sieve           one shared-memory tile per CTA per segment.  Tile init = AND of the patterns
//                     (uint4), then every base prime p >= 61 (staged once in HBM/L2) is crossed off by
//                     shared-memory atomicAnd on the packed words: 8 strands per prime (one per wheel
//                     residue of the cofactor), each a byte-stride-p progression with a fixed bit.
//                     Work split by hits-per-tile: warp-wide (dense primes) and 8-lane groups (sparse primes);
//                     per-prime reciprocal and strand geometry are staged once so the per-(prime, tile)
//                     set-up is a dozen instructions.
//                     Tiles are independent (first offsets are recomputed from the tile base), so any CTA
//                     on any GPU can take any tile.  Output: bitmap tile (coalesced uint4 stores) +
//                     per-chunk and per-tile popcounts.
//   k_scan_tile_counts single CTA exclusive scan of the per-tile counts (<= a few 10^4 values).
//   k_compact         popc + block scan stream compaction of bitmap chunks into the coalesced u64 prime
//                     buffer (staged through shared memory so global stores are fully coalesced).
#pragma once
#include "pss_common.cuh"

namespace pss {

constexpr int kMaxPatterns = 6;

struct SieveParams {
    u64 lo, hi;          // value range [lo, hi)
    u64 B_start;         // absolute byte index of bitmap byte 0 (multiple of 16)
    u32 tile_bytes;      // multiple of kChunkBytes
    u32 n_tiles;
    // base primes staged once in HBM (L2 resident): ascending, first entry = smallest prime not covered
    // by a pattern.  Per prime: p, floor(2^32/p) for the mul-hi modulo, and per strand j the packed
    // geometry  c_j | bit_j << 24  (c_j = floor(p*R[j]/30) < p < 2^24).
    const u32* base_p;
    const u32* base_inv;
    const u32* base_strand;   // n_base * 8
    u32 n_base;
    u32 n_warp_wide;     // base primes [0, n_warp_wide) are "dense" (one warp per prime), the rest go in batches of 32
    const u32* batch_geo;     // per batch b of sparse primes [n_warp_wide + 32b, +32): n_uni | n_rag << 16, the warp-uniform
                              // trip counts per strand (n_uni = tile/p_hi - 1 unguarded, n_rag = (tile-1)/p_lo + 1 - n_uni guarded)
    const u32* tile_nact;     // per tile: number of base primes with p*p < tile end (the only ones that can
                              // still cross something off there)
    u32 sparse_lead;     // the first sparse_lead batches are queued BEFORE the dense primes: a batch near the dense limit
                         // is ~800 conflicted atomics per lane, the longest item of the tile by far (a dense prime is
                         // 26 .. 940 conflict-free ones), so it has to start first or it becomes the tail of phase B
    u32 wide_index;      // 1: tile byte indices may exceed 2^32 -> 64-bit modulo path
    u32 n_pat;
    const uint4* pat[kMaxPatterns];
    u64 pat_period16[kMaxPatterns];   // 16 * period (bytes)
    u32* bitmap;         // n_tiles * tile_bytes / 4 words
    u32* chunk_counts;   // n_tiles * tile_bytes / kChunkBytes
    u32* tile_counts;    // n_tiles
    unsigned long long* prof;   // PSS_SIEVE_PROF=1: cycle counters (phase A / B / C per CTA; warp time per item class), else null
};

// ---- pattern construction -------------------------------------------------------------------
// pattern byte i (absolute byte index i mod period): bit j cleared iff 30*i + R[j] is divisible by one
// of the group's primes.  (The primes themselves are cleared too; k_sieve restores bytes 0 and 1.)
__global__ void k_build_pattern(uint8_t* out, u64 n_bytes, u32 q0, u32 q1, u32 q2, u32 q3, u32 q4) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    u64 stride = (u64)gridDim.x * blockDim.x;
    const u32 qs[5] = {q0, q1, q2, q3, q4};
    for (; i < n_bytes; i += stride) {
        u32 m = 0xFF;
#pragma unroll
        for (u32 j = 0; j < 8; ++j) {
            u64 v = i * 30ull + wheel_residue(j);
#pragma unroll
            for (int k = 0; k < 5; ++k)
                if (qs[k] && (v % qs[k]) == 0) m &= ~(1u << j);
        }
        out[i] = (uint8_t)m;
    }
}

// per tile: number of base primes with p*p < tile end (exact 64-bit compare, binary search; ascending table)
__global__ void k_tile_nact(const u32* base_p, u32 n_base, u64 B_start, u32 tile_bytes, u64 hi, u32 n_tiles, u32* out) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    u64 tile_end = (B_start + (u64)(t + 1) * tile_bytes) * 30ull;
    if (tile_end > hi) tile_end = hi;
    u32 lo = 0, hiI = n_base;            // first index with p*p >= tile_end
    while (lo < hiI) {
        const u32 mid = (lo + hiI) >> 1;
        const u64 p = base_p[mid];
        if (p * p < tile_end) lo = mid + 1; else hiI = mid;
    }
    out[t] = lo;
}

// r = B0 mod p.  32-bit byte indices (hi < 1.28e11): one mul-hi with the staged reciprocal.
__device__ __forceinline__ u32 tile_mod(u64 B0, u32 p, u32 inv, bool wide) {
    if (wide) return (u32)(B0 % p);
    const u32 b = (u32)B0;
    u32 r = b - __umulhi(b, inv) * p;   // in [0, 2p)
    if (r >= p) r -= p;
    return r;
}

// first byte offset of the strand inside the tile: (c - r) mod p, never the prime itself (cofactor 1)
__device__ __forceinline__ u32 first_hit(u32 p, u32 c, u32 r, bool is_strand0, u64 B0) {
    u32 x = (c >= r) ? (c - r) : (c + p - r);
    if (is_strand0 && B0 <= (u64)c) x += p;
    return x;
}

// inverse of an odd p modulo 64 (one Newton step from x0 = p, which is already correct modulo 8)
__device__ __forceinline__ u32 inv_mod64(u32 p) { return p * (2u - p * p); }

// one slot from the phase-B work queue.  Inline PTX on purpose: the compiler rewrites `if (lane == 0) atomicAdd(&ctr, 1)` into
// its warp-aggregated form (leader ATOMS.ADD + SHFL of the result to all lanes), and that SHFL waits for the atomic right
// where it was issued -- which defeats pulling ahead of time
__device__ __forceinline__ u32 queue_pull(u32* ctr) {
    u32 r;
    asm volatile("atom.shared.inc.u32 %0, [%1], 0x3FFFFFFF;" : "=r"(r) : "r"((u32)__cvta_generic_to_shared(ctr)));   // .inc with a limit below 2^32 - 1: ptxas warp-aggregates .add (and .inc 0xFFFFFFFF) too
    return r;
}

// the sieve kernel's dynamic shared memory: the tile followed by a 128-byte scratch line (one word per bank) that
// absorbs the cross-offs of lanes that are outside the tile in a given iteration (see the dense phase)
constexpr u32 kSieveScratchBytes = 128;

template <int THREADS, bool PROF = false>
__global__ void __launch_bounds__(THREADS, (THREADS <= 768) ? 2 : 1) k_sieve(const SieveParams P) {
    extern __shared__ uint4 smem4[];
    __shared__ u64 s_off16[kMaxPatterns];
    __shared__ u32 s_cnt[THREADS / 32];
    __shared__ u32 s_next;       // work queue of phase B (dense primes first, then batches of 32 sparse primes)

    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr u32 NW = THREADS / 32;
    const u32 tile_bytes = P.tile_bytes;
    const u32 tile_vec = tile_bytes >> 4;
    const u32 chunks_per_tile = tile_bytes / kChunkBytes;
    const bool wide = P.wide_index != 0;
    char* const tile_b = reinterpret_cast<char*>(smem4);
    // AND-reduction on the tile byte offset `off` (multiple of 4).  Plain atomicAnd with the result unused compiles
    // to ATOMS.AND RZ; unlike an asm volatile statement it leaves the enclosing regions reconvergent.
    auto cross = [&](u32 off, u32 mask) { atomicAnd(reinterpret_cast<u32*>(tile_b + off), mask); };

    for (u32 t = blockIdx.x; t < P.n_tiles; t += gridDim.x) {
        const u64 B0 = P.B_start + (u64)t * tile_bytes;
        constexpr bool prof = PROF;   // instrumented instantiation (PSS_SIEVE_PROF=1)
        long long pc0 = 0, pc1 = 0, pc2 = 0;
        unsigned long long pw[4] = {0, 0, 0, 0};
        if (prof) pc0 = clock64();
        if (tid < P.n_pat) s_off16[tid] = (B0 % P.pat_period16[tid]) >> 4;
        if (tid == 32) s_next = 0;
        __syncthreads();

        // ---- phase A: tile init from the pre-sieved patterns (L2-resident, aligned uint4 loads)
        for (u32 i = tid; i < tile_vec; i += THREADS) {
            uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u);
            for (u32 k = 0; k < P.n_pat; ++k) {
                uint4 w = __ldg(P.pat[k] + s_off16[k] + i);
                v.x &= w.x; v.y &= w.y; v.z &= w.z; v.w &= w.w;
            }
            smem4[i] = v;
        }
        __syncthreads();
        if (prof) pc1 = clock64();

        // ---- phase B: cross-offs.  Warps pull work items from a shared-memory queue, longest items first:
        //        slots [0, lead)              the first `lead` batches of 32 consecutive sparse primes, one prime per lane
        //        slots [lead, lead + n_dense) one dense prime each (p < tile/104), whole warp, bank-conflict free
        //        slots [lead + n_dense, ..)   the remaining sparse batches (few hits per lane)
        long long pc1w = 0, plast = 0;
        unsigned long long pgap0 = 0, pgap = 0;
        if (prof) pc1w = clock64();
        const u32 nact = __ldg(P.tile_nact + t);
        const u32 n_dense = min(P.n_warp_wide, nact);
        const u32 n_batches = (nact - n_dense + 31u) >> 5;
        const u32 n_items = n_dense + n_batches;
        const u32 lead = min(P.sparse_lead, n_batches);
        // the queue counter AND the shuffle that broadcasts it go through the pipe the cross-offs keep full: either takes
        // several hundred cycles to come back.  So the queue runs two items ahead: while a warp works on slot k, the
        // broadcast of slot k + 1 and the pull of slot k + 2 are in flight (instrumented: 520 cycles between two items
        // of a warp with the pull one ahead and the shuffle on the critical path, 10 % of phase B)
        // (pulling the first slot before phase A, with nact loaded there too, costs 8 registers -- 64, the limit of a
        // 1024-thread CTA -- and the cross-off loops 15 %: measured 2.015 -> 2.317 ms)
        u32 pulled = 0;
        if (lane == 0) pulled = queue_pull(&s_next);
        u32 slot = __shfl_sync(0xFFFFFFFFu, pulled, 0);
        if (lane == 0) pulled = queue_pull(&s_next);
        for (;;) {
            if (slot >= n_items) break;
            const u32 cur = slot;
            slot = __shfl_sync(0xFFFFFFFFu, pulled, 0);      // consumed after this item
            if (lane == 0) pulled = queue_pull(&s_next);
            // item < n_dense: dense prime `item`; else sparse batch item - n_dense
            const u32 item = (cur < lead) ? cur + n_dense : (cur < lead + n_dense) ? cur - lead : cur;
            long long pi0 = 0;
            u32 pcls = 0;
            if (prof) {
                pi0 = clock64();
                if (plast == 0) pgap0 += (unsigned long long)(pi0 - pc1w); else pgap += (unsigned long long)(pi0 - plast);
            }
            if (item < n_dense) {
                // dense prime (p <= tile/72): lane -> (strand j, sub-sequence s), word stride p (byte stride 4p),
                // constant mask; every lane has count = H-1 .. H+1 >= 17 hits (H = tile / 4p).
                // Bank-conflict-free order: lane l visits its hits rotated by e_l = (w_l - l) / p mod 32, i.e. first the
                // last e_l hits (indices count-e_l .. count-1), then 0, 1, ...  Once a lane is past its wrap-around it
                // addresses bank (l + it*p) mod 32 in iteration it: from iteration 32 on the 32 lanes hit the 32 different
                // banks (one wavefront per warp instruction, measured 1.00); in the first 32 iterations the lanes still in
                // their wrapped part are shifted by count*p and a bank sees at most three accesses.  No lane ever waits:
                // count iterations cover count hits.
                const u32 j = lane & 7u, s = lane >> 3;
                const u32 p = __ldg(P.base_p + item), inv = __ldg(P.base_inv + item);
                const u32 cb = __ldg(P.base_strand + item * 8u + j);
                const u32 r = tile_mod(B0, p, inv, wide);
                const u32 c = cb & 0xFFFFFFu;
                const bool special = (j == 0) && (B0 <= (u64)c);      // q = 0 would be p itself: first hit is one period later
                const u32 x = first_hit(p, c, r, j == 0, B0) + s * p;  // x < 4p unless special (then x < 5p)
                const u32 mask = ~(1u << (((x & 3u) << 3) + (cb >> 24)));
                const u32 w = x >> 2;
                const u32 step = p << 2;
                const u32 H = tile_bytes / step;                       // >= 18 (host: n_warp_wide)
                // full rotation (mod 32) needs 33 hits per lane; with 17..32 hits the rotation is taken mod 16: lanes l and
                // l + 16 then share a bank class and collide half of the time (2 wavefronts instead of ~3.5 unordered)
                const bool full = H >= 34u;
                pcls = full ? 0u : 1u;
                const u32 e = special ? 0u : (((w - lane) * inv_mod64(p)) & (full ? 31u : 15u));
                const u32 b0 = w << 2;                                 // byte offset of hit 0 (< 2 * step)
                const u32 count = H - 1u + ((b0 + (H - 1u) * step < tile_bytes) ? 1u : 0u) + ((b0 + H * step < tile_bytes) ? 1u : 0u);
                const u32 span = count * step;
                u32 a = b0 + span - e * step;                          // hit count - e  (e = 0: wraps before the first access)
                if (full) {
#pragma unroll
                    for (u32 it = 0; it < 32u; ++it) {                 // lanes wrap around at it == e
                        if (it == e) a -= span;
                        cross(a, mask);
                        a += step;
                    }
#pragma unroll 4
                    for (u32 it = 32u; it + 1u < H; ++it) {            // steady state: hits e.. of every lane, conflict free
                        cross(a, mask);
                        a += step;
                    }
                } else {
#pragma unroll
                    for (u32 it = 0; it < 16u; ++it) {
                        if (it == e) a -= span;
                        cross(a, mask);
                        a += step;
                    }
                    for (u32 it = 16u; it + 1u < H; ++it) {
                        cross(a, mask);
                        a += step;
                    }
                }
#pragma unroll
                for (u32 it = H - 1u; it <= H; ++it) {                 // lanes with count = H, H + 1; the others AND into the
                    cross((it < count) ? a : ((a & 127u) | tile_bytes), mask);   // scratch word of the bank they would have used
                    a += step;
                }
            } else {
                // sparse primes: one prime per lane, its 8 strands one after the other; byte stride p (odd), so the
                // byte position inside the word advances by p & 3: the mask rotates by a per-prime constant.
                // Trip counts are warp-uniform (no votes, no divergence): every lane of the batch has at least
                // n_uni = T / p_hi - 1 hits per strand and at most h_max = (T - 1) / p_lo + 1; the iterations in
                // between are guarded by redirecting lanes that have left the tile to their scratch word.
                const u32 first = n_dense + ((item - n_dense) << 5);
                const u32 n_on = min(32u, nact - first);
                const bool on = lane < n_on;
                const u32 park = tile_bytes + 4u * lane;     // this lane's scratch word
                u32 p = 0, inv = 0;
                uint4 s0 = make_uint4(park, park, park, park), s1 = s0;   // idle lane: c = park, bit 0, p = 0: stays parked
                if (on) {
                    const u32 i = first + lane;
                    p = __ldg(P.base_p + i);
                    inv = __ldg(P.base_inv + i);
                    s0 = __ldg(reinterpret_cast<const uint4*>(P.base_strand) + 2u * i);
                    s1 = __ldg(reinterpret_cast<const uint4*>(P.base_strand) + 2u * i + 1u);
                }
                const u32 geo = __ldg(P.batch_geo + (item - n_dense));   // host-staged: n_uni | n_rag << 16 (uniform load)
                const u32 n_uni = geo & 0xFFFFu, n_rag = geo >> 16;
                const u32 cbs[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                const u32 r = on ? tile_mod(B0, p, inv, wide) : 0u;
                const u32 rot = (p & 3u) << 3;
                pcls = (n_uni == 0u && n_rag <= 2u) ? 3u : 2u;
                if (n_uni == 0u && n_rag <= 2u) {
                    // large primes (p > tile/2): at most two hits per strand -- straight-line code
                    const bool two = n_rag == 2u;
#pragma unroll
                    for (u32 j = 0; j < 8; ++j) {
                        const u32 cb = cbs[j];
                        u32 x = first_hit(p, cb & 0xFFFFFFu, r, j == 0, B0);
                        u32 mask = ~(1u << (((x & 3u) << 3) + (cb >> 24)));
                        cross((x < tile_bytes) ? (x & ~3u) : park, mask);
                        if (two) {
                            x += p;
                            mask = __funnelshift_l(mask, mask, rot);
                            cross((x < tile_bytes) ? (x & ~3u) : park, mask);
                        }
                    }
                } else {
#pragma unroll
                    for (u32 j = 0; j < 8; ++j) {
                        const u32 cb = cbs[j];
                        u32 x = first_hit(p, cb & 0xFFFFFFu, r, j == 0, B0);
                        u32 mask = ~(1u << (((x & 3u) << 3) + (cb >> 24)));
#pragma unroll 4
                        for (u32 k = 0; k < n_uni; ++k) {
                            cross(x & ~3u, mask);
                            x += p;
                            mask = __funnelshift_l(mask, mask, rot);
                        }
#pragma unroll 1
                        for (u32 k = 0; k < n_rag; ++k) {
                            cross((x < tile_bytes) ? (x & ~3u) : park, mask);
                            x += p;
                            mask = __funnelshift_l(mask, mask, rot);
                        }
                    }
                }
            }
            if (prof) {
                plast = clock64();
                pw[pcls] += (unsigned long long)(plast - pi0);
            }
        }
        long long pend = 0;
        if (prof) pend = clock64();
        __syncthreads();
        if (prof) {
            pc2 = clock64();
            if (lane == 0) {
                for (u32 c = 0; c < 4; ++c)
                    if (pw[c]) atomicAdd(P.prof + 4 + c, pw[c]);
                atomicAdd(P.prof + 8, (unsigned long long)(pc2 - pend));   // this warp's wait for the slowest one
                atomicAdd(P.prof + 9, pgap0);                              // entry of B -> first item
                atomicAdd(P.prof + 10, pgap);                              // between items (queue)
                atomicAdd(P.prof + 11, (unsigned long long)(pend - (plast ? plast : pc1w)));   // last item -> loop exit
            }
        }

        // ---- phase C: edge masks, popcounts, coalesced write-out of the bitmap tile
        const bool interior = (B0 * 30ull >= P.lo) && ((B0 + tile_bytes) * 30ull <= P.hi) && (B0 >= 4);
        u32 my_tile_cnt = 0;
        uint4* gout = reinterpret_cast<uint4*>(P.bitmap) + (u64)t * tile_vec;
        for (u32 ch = warp; ch < chunks_per_tile; ch += NW) {
            u32 cnt = 0;
#pragma unroll
            for (u32 h = 0; h < 2; ++h) {
                const u32 vi = ch * (kChunkWords / 4) + h * 32u + lane;
                uint4 v = smem4[vi];
                if (!interior) {
                    u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (u32 k = 0; k < 4; ++k) {
                        const u64 Bw = B0 + (u64)vi * 16ull + 4ull * k;
                        u32 x = w[k];
                        if (Bw == 0) x = 0x7EEFDFFEu;   // the integers below 120: all wheel residues are prime except 1, 49, 77, 91, 119
                                                         // (the patterns also clear their own primes, 7..89)
                        u32 m = edge_byte_mask(Bw, P.lo, P.hi) | (edge_byte_mask(Bw + 1, P.lo, P.hi) << 8) |
                                (edge_byte_mask(Bw + 2, P.lo, P.hi) << 16) | (edge_byte_mask(Bw + 3, P.lo, P.hi) << 24);
                        w[k] = x & m;
                    }
                    v = make_uint4(w[0], w[1], w[2], w[3]);
                }
                gout[vi] = v;
                cnt += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
            }
            cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
            if (lane == 0) P.chunk_counts[(u64)t * chunks_per_tile + ch] = cnt;
            my_tile_cnt += cnt;  // identical in every lane
        }
        if (lane == 0) s_cnt[warp] = my_tile_cnt;
        __syncthreads();
        if (tid == 0) {
            u32 s = 0;
            for (u32 w = 0; w < NW; ++w) s += s_cnt[w];
            P.tile_counts[t] = s;
        }
        // s_off16 / s_cnt / tile are rewritten only after the next iteration's first barrier
        __syncthreads();
        if (prof && tid == 0) {
            const long long pc3 = clock64();
            atomicAdd(P.prof + 0, (unsigned long long)(pc1 - pc0));
            atomicAdd(P.prof + 1, (unsigned long long)(pc2 - pc1));
            atomicAdd(P.prof + 2, (unsigned long long)(pc3 - pc2));
            atomicAdd(P.prof + 3, 1ull);
        }
    }
}

// ---- K1 variant "individual": per-candidate trial division (reference: utils/sieve.py:330-387 _individual_sieve) --------
// The reference's minimal-memory variant tests every candidate on its own by trial division with the primes up to its
// square root.  Here: one CTA per 1 KiB bitmap chunk, one thread per bitmap word (32 wheel-30 candidates); a candidate is
// prime iff no prime 7 <= p <= sqrt(v) divides it (2, 3, 5 are excluded by the wheel).  Same outputs as k_sieve (bitmap,
// chunk and tile popcounts), so counting, compaction and the fused scan run unchanged on top of it.  O(n sqrt(n) / log n)
// like the reference's: a selectable cross-check of the sieve, not a fast path.
__global__ void __launch_bounds__(kChunkWords) k_individual(const SieveParams P) {
    __shared__ u32 s_cnt[kChunkWords / 32];
    const u32 chunks_per_tile = P.tile_bytes / kChunkBytes;
    const u64 n_chunks = (u64)P.n_tiles * chunks_per_tile;
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // every prime a pre-sieve pattern group can cover (pss_api.cu: kPatGroups), so the test does not depend on where the
    // staged base-prime table starts
    const u32 kSmall[21] = {7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89};
    for (u64 ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const u64 Bw = P.B_start + ch * (u64)kChunkBytes + 4ull * tid;
        u32 word = 0;
        for (u32 bit = 0; bit < 32; ++bit) {
            const u64 v = word_bit_value(Bw, bit);
            if (v < 7 || v < P.lo || v >= P.hi) continue;
            bool prime = true;
            for (u32 i = 0; i < 21 && prime; ++i) {
                const u64 q = kSmall[i];
                if (q * q > v) break;
                if (v % q == 0) prime = false;
            }
            // base_p holds the primes from the first one not covered by a pre-sieve pattern (the host stages them up to
            // sqrt(hi)); entries up to 89 were tested above
            for (u32 i = 0; i < P.n_base && prime; ++i) {
                const u64 q = __ldg(P.base_p + i);
                if (q * q > v) break;
                if (q > 89 && v % q == 0) prime = false;
            }
            if (prime) word |= 1u << bit;
        }
        P.bitmap[ch * kChunkWords + tid] = word;
        u32 cnt = __reduce_add_sync(0xFFFFFFFFu, __popc(word));
        if (lane == 0) s_cnt[warp] = cnt;
        __syncthreads();
        if (tid == 0) {
            u32 total = 0;
            for (u32 w = 0; w < kChunkWords / 32; ++w) total += s_cnt[w];
            P.chunk_counts[ch] = total;
            atomicAdd(P.tile_counts + (u32)(ch / chunks_per_tile), total);
        }
        __syncthreads();
    }
}

// ---- K1b: exclusive scan of per-tile counts (u32 -> u64), single CTA ----------------------------
// prefix[t] = init + sum_{i<t} counts[i];  prefix[n] = grand total.
__global__ void __launch_bounds__(1024) k_scan_tile_counts(const u32* counts, u32 n, u64 init, u64* prefix) {
    __shared__ u64 s_warp[32];
    __shared__ u64 s_running;
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_running = init;
    __syncthreads();
    for (u32 base = 0; base < n; base += 1024) {
        const u32 i = base + tid;
        const u64 v = (i < n) ? counts[i] : 0;
        u64 incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u64 o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (u32)d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            u64 w = s_warp[lane];
            u64 wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u64 o = __shfl_up_sync(0xFFFFFFFFu, wi, d);
                if (lane >= (u32)d) wi += o;
            }
            s_warp[lane] = wi - w;  // exclusive warp offsets
        }
        __syncthreads();
        const u64 run = s_running;
        const u64 excl = run + s_warp[warp] + incl - v;
        if (i < n) prefix[i] = excl;
        __syncthreads();
        if (tid == 1023) s_running = excl + v;
        __syncthreads();
    }
    if (tid == 0) prefix[n] = s_running;
}

// ---- K1c: compaction of bitmap chunks into the u64 prime buffer ----------------------------------
// CTA (256 threads) <-> groups of `group` consecutive chunks of one tile; thread <-> word.
struct CompactParams {
    const u32* bitmap;
    const u32* chunk_counts;
    const u64* tile_prefix;   // exclusive prefix of tile counts (includes the 2,3,5 offset)
    u32 chunks_per_tile;
    u32 group;                // chunks per CTA work item (divides chunks_per_tile)
    u32 n_groups;             // total work items = n_tiles * chunks_per_tile / group
    u64 B_start;
    u64* out;
    u64 out_cap;              // only indices < out_cap are written (first-n truncation)
    u32* error_flag;
};

__global__ void __launch_bounds__(256) k_compact(const CompactParams P) {
    __shared__ u64 stage[kMaxChunkPrimes];
    __shared__ u32 s_wsum[8];
    __shared__ u32 s_pre[8];
    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u32 groups_per_tile = P.chunks_per_tile / P.group;

    for (u32 g = blockIdx.x; g < P.n_groups; g += gridDim.x) {
        const u32 t = g / groups_per_tile;
        const u32 c0 = (g - t * groups_per_tile) * P.group;   // first chunk (within tile) of this item
        const u64 chunk0 = (u64)t * P.chunks_per_tile + c0;
        // prefix of the first chunk = tile prefix + counts of the tile's earlier chunks
        u32 pre = 0;
        for (u32 i = tid; i < c0; i += 256) pre += P.chunk_counts[(u64)t * P.chunks_per_tile + i];
        pre = __reduce_add_sync(0xFFFFFFFFu, pre);
        if (lane == 0) s_pre[warp] = pre;
        __syncthreads();
        u64 base = P.tile_prefix[t];
#pragma unroll
        for (u32 w = 0; w < 8; ++w) base += s_pre[w];
        if (base >= P.out_cap) { __syncthreads(); continue; }

        for (u32 c = 0; c < P.group; ++c) {
            const u64 widx = (chunk0 + c) * kChunkWords + tid;
            u32 w = P.bitmap[widx];
            const u32 cnt = __popc(w);
            u32 incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                u32 o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= (u32)d) incl += o;
            }
            if (lane == 31) s_wsum[warp] = incl;
            __syncthreads();
            u32 woff = 0, total = 0;
#pragma unroll
            for (u32 k = 0; k < 8; ++k) {
                const u32 s = s_wsum[k];
                if (k < warp) woff += s;
                total += s;
            }
            if (total > kMaxChunkPrimes) {  // cannot happen for a correctly sieved bitmap
                if (tid == 0) atomicExch(P.error_flag, 2u);
                total = kMaxChunkPrimes;
            }
            u32 off = woff + incl - cnt;
            const u64 Bw = P.B_start + widx * 4ull;
            while (w) {
                const u32 b = __ffs(w) - 1;
                w &= w - 1;
                if (off < kMaxChunkPrimes) stage[off] = word_bit_value(Bw, b);
                ++off;
            }
            __syncthreads();
            for (u32 i = tid; i < total; i += 256) {
                const u64 o = base + i;
                if (o < P.out_cap) P.out[o] = stage[i];
            }
            base += total;
            __syncthreads();
        }
    }
}

}  // namespace pss
