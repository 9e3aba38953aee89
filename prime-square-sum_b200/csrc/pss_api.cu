// pss_api.cu — the C ABI of libpss_b200.so (include/pss_b200.h): handle, HBM workspace, kernel
// launches and CUDA-event timing.  No CPU fallback: every compute entry point runs the sm_100a kernels
// in pss_sieve.cuh / pss_scan.cuh or fails.
//
// HBM layout per handle (grow-only, reused across calls so steady-state calls do not allocate):
//   pattern[4]      16x-replicated pre-sieve byte patterns (9 MB total, L2 resident)
//   base_primes     u32 primes 61 .. sqrt(hi)        (56 KB at hi = 2.28e10)
//   bitmap          wheel-30 bit-packed sieve result, 1 byte / 30 integers (760 MB at hi = 2.28e10)
//   chunk_counts    u32 popcount per 1 KiB bitmap chunk;  tile_counts / tile_prefix per sieve tile
//   primes          compacted u64 primes (8 B / prime; 8 GB for the first 1e9 primes)
//   scan state      per scan tile: status word + aggregate + inclusive limb records
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/pss_b200.h"
#include "pss_bscan.cuh"
#include "pss_trihead.cuh"
#include "pss_peer.cuh"
#include "pss_scan.cuh"
#include "pss_sieve.cuh"

using namespace pss;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

constexpr u64 kMaxValue = 1ull << PSS_MAX_PRIME_BITS;
constexpr u32 kPatternPad = 256 * 1024;   // >= largest tile
constexpr u32 kMaxTileBytes = 224 * 1024; // + scratch line + static shared memory < 227 KiB opt-in limit
static const u32 kPatGroups[kMaxPatterns][5] = {{7, 11, 13, 17, 19}, {23, 29, 31, 0, 0}, {37, 41, 43, 0, 0}, {47, 53, 59, 0, 0},
                                                {61, 67, 71, 0, 0}, {73, 79, 83, 0, 0}};
static const u32 kFirstSievePrime[kMaxPatterns + 1] = {7, 23, 37, 47, 61, 73, 89};

struct MiscDev {   // small device-side block, one H2D / D2H per scan
    u32 ticket;
    u32 error_flag;
    u32 pad[2];
    pss_power_result results[PSS_MAX_POWER];
    u32 carry[PSS_MAX_POWER * PSS_LIMBS];
    u64 last_prime;
    u64 n_first;     // peer path: global index of the slice's first prime (written by k_peer_gather)
    u64 tl[8];       // peer path: globaltimer stamps of the exchange kernels (pss_peer.cuh: stamp)
};

struct HostBack {   // pinned: asynchronous D2H targets of one query
    u64 sv_total;
    u64 col_totals[1 + PSS_MAX_POWER * PSS_LIMBS];
};

// A query chain keeps the copy engine out of its critical path: a copy between two kernels costs ~8 us of stream latency
// (profiles/r02: per-rank timeline), a dependent kernel ~2 us.  So the MiscDev block is INITIALISED by a one-CTA kernel from
// by-value arguments instead of an H2D copy, and the query's results (MiscDev block, column totals, peer outcome records)
// are stored by a one-CTA kernel straight into mapped pinned host memory instead of two or three D2H copies.
struct MiscInit {
    u32 np;
    u32 power[PSS_MAX_POWER];
    u32 carry[PSS_MAX_POWER * PSS_LIMBS];
};
__global__ void k_init_misc(MiscDev* dm, const MiscInit in) {
    MiscDev* d = dm + blockIdx.x;
    u32* w = reinterpret_cast<u32*>(d);
    for (u32 i = threadIdx.x; i < sizeof(MiscDev) / 4; i += blockDim.x) w[i] = 0u;
    __syncthreads();
    if (threadIdx.x < in.np) {
        d->results[threadIdx.x].power = in.power[threadIdx.x];
        d->results[threadIdx.x].first_match_n = ~0ull;
    }
    for (u32 i = threadIdx.x; i < PSS_MAX_POWER * PSS_LIMBS; i += blockDim.x) d->carry[i] = in.carry[i];
}
struct ExportParams {
    const MiscDev* dm;       // device block ...
    MiscDev* hm;             // ... -> mapped pinned host block
    const u64* col_prefix;   // column totals at col_prefix[c * col_stride + n_tiles], c < ncols (ncols == 0: none)
    u64 col_stride;
    u32 n_tiles, ncols;
    u64* totals_host;
    const PeerResult* peer;  // peer outcome records (world == 0: none)
    PeerResult* peer_host;
    u32 world;
};
__global__ void k_export(const ExportParams P) {
    const u32* src = reinterpret_cast<const u32*>(P.dm);
    u32* dst = reinterpret_cast<u32*>(P.hm);
    for (u32 i = threadIdx.x; i < sizeof(MiscDev) / 4; i += blockDim.x) dst[i] = src[i];
    for (u32 c = threadIdx.x; c < P.ncols; c += blockDim.x) P.totals_host[c] = P.col_prefix[(u64)c * P.col_stride + P.n_tiles];
    const u64* ps = reinterpret_cast<const u64*>(P.peer);
    u64* pd = reinterpret_cast<u64*>(P.peer_host);
    for (u32 i = threadIdx.x; i < P.world * (u32)(sizeof(PeerResult) / 8); i += blockDim.x) pd[i] = ps[i];
    __threadfence_system();
}

}  // namespace

struct pss_handle {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    char name[256] = {0};

    // tuning
    u32 tile_bytes = 224 * 1024;   // one 1024-thread CTA per SM (tools/tune_sieve.py, profiles/r01/tune_sieve_*.json)
    u32 sieve_threads = 1024;
    u32 n_pat = 4;
    u32 sparse_lead = 0xFFFFFFFFu;   // sparse batches queued ahead of the dense primes (PSS_SPARSE_LEAD; all of them by default)
    u32 warp_wide_div = 104;    // p < tile_bytes / div -> dense (one warp per prime, skewed conflict-free lanes), else one lane per prime
    int kernel_events = 1;     // PSS_KERNEL_EVENTS
    int sieve_variant = 0;     // 0: segmented wheel-30 sieve (k_sieve), 1: "individual" per-candidate trial division (k_individual)
    int search_path = 0;       // 0: fused (scan straight from the bitmap), 1: materialised u64 primes
    int compare_exhaustive = 0;  // 1: the compare pass walks every prime; 0: interval pruning (PSS_COMPARE_EXHAUSTIVE / pss_set_compare_mode)
    int reduce_moments = 2;    // block sums from moments (k = 2: per-byte tables; other k <= 5: per-prime small moments);
                               // 0: multi-limb power chain per prime (PSS_REDUCE_MOMENTS)

    // patterns
    uint4* d_pat[kMaxPatterns] = {};
    u64 pat_period16[kMaxPatterns] = {};

    // base primes
    std::vector<u32> h_base;   // all primes >= 7 up to base_limit
    u64 base_limit = 0;
    DevBuf d_base, d_base_inv, d_base_strand, d_batch_geo;
    std::vector<u32> h_batch_geo;   // host copy of the staged sparse-batch geometry (re-uploaded only when it changes)
    u32 base_first_prime = 0;  // first sieving prime the device copy starts at
    u32 n_base_dev = 0;

    // sieve workspace / state of the last sieve
    DevBuf bitmap, chunk_counts, tile_counts, tile_prefix, tile_nact;
    u64 nact_key[6] = {0, 0, 0, 0, 0, 0};   // geometry the cached tile_nact table was computed for
    DevBuf tri_head;     // head pass of the triangular predicate: per-word prefix records + the head's prime count
    DevBuf sieve_prof;   // PSS_SIEVE_PROF=1: k_sieve cycle counters, printed by pss_close
    u64 sv_lo = 0, sv_hi = 0, sv_B_start = 0, sv_total = 0, sv_c235 = 0;
    u32 sv_tile_bytes = 0, sv_n_tiles = 0;
    bool sv_valid = false;
    bool sv_pending = false;   // sieve enqueued, count not read back yet (sieve_finish)
    int occ_cache = 0;                   // resident sieve CTAs per SM for (occ_cache_threads, occ_cache_tile)
    u32 occ_cache_threads = 0, occ_cache_tile = 0;
    bool sv_total_from_reduce = false;   // deferred sieve: no tile-count scan, the total comes from the block sums
    bool bs_pending = false;   // reduce enqueued, column totals not read back yet (reduce_finish)
    bool bs_totals_copied = false;   // ... and their D2H copy is enqueued (copy_totals)
    bool misc_uploaded = false;      // the query's MiscDev block is already on its way to the device (prepare_misc)

    // resident primes
    DevBuf primes;
    u64 n_resident = 0;
    u64 resident_last = 0;

    // fused path: column-major tile records of the last bitmap reduce and their prefixes
    DevBuf cols, col_prefix, thread_rec, warp_rec;
    bool bs_valid = false;
    u32 bs_K = 0, bs_n_tiles = 0, bs_total_cols = 0;
    bool bs_multi = false;
    u64 bs_col_stride = 0, bs_count = 0;
    u32 bs_sums[PSS_MAX_POWER * PSS_LIMBS] = {0};   // slice sums per power record (normalised limbs)

    // scan workspace
    DevBuf status, agg, incl, misc, prefix_out, user_vals, tri_buf, misc_multi;
    MiscDev* h_misc_multi = nullptr;   // pinned, h_misc_multi_cap blocks (pss_search_targets)
    size_t h_misc_multi_cap = 0;
    MiscDev* h_misc = nullptr;   // pinned + mapped (k_export stores into it)
    HostBack* h_back = nullptr;  // pinned + mapped
    MiscDev* h_misc_dev = nullptr;    // device-side addresses of the two
    HostBack* h_back_dev = nullptr;
    PeerResult* h_peer_results_dev = nullptr;

    // peer-memory exchange (pss_peer_*): own mailbox, mapped mailboxes of all ranks, query epoch
    PeerBox* box_local = nullptr;
    PeerBox* box[kMaxPeers] = {nullptr};
    int peer_rank = -1, peer_world = 0;
    u64 peer_epoch = 0;
    PeerResult* h_peer_results = nullptr;   // pinned, kMaxPeers records

    // resident CTAs per SM, per kernel and launch shape.  Kept in the handle: occupancy and the opt-in shared-memory
    // attribute belong to the handle's device (a process may hold handles on several GPUs)
    std::unordered_map<const void*, int> occ_map;

    // CUDA graph of the single-GPU query chain (pss_search, fused path).  Keyed by the chain's GEOMETRY (range, power
    // configuration, tuning, buffer generation): a query with the same geometry replays the instantiated graph with the
    // compare kernel's argument (target, window, operator) updated in place -- one driver call instead of ten.
    int use_graph = 1;                   // PSS_GRAPH=0: always launch kernel by kernel
    bool dry = false;                    // bookkeeping pass of a replay: launches, memsets and event records are skipped
    bool capturing = false;              // the stream is being captured: timing events become external event nodes
    cudaGraph_t g_graph = nullptr;       // kept alive: node handles of the template graph address the nodes of g_exec
    cudaGraphExec_t g_exec = nullptr;
    cudaGraphNode_t g_cmp_node = nullptr;
    bool dry_miss = false;               // the dry pass met a launch that is not part of the captured chain
    cudaKernelNodeParams g_cmp_params = {};   // the compare node as captured (function, grid, block, shared memory)
    std::vector<u64> g_key, g_seen_key;  // geometry of the instantiated graph / of the previous kernel-by-kernel query
    u64 g_kernels = 0;                   // kernel nodes in the graph (launch counter)
    u64 alloc_generation = 0;            // bumped whenever a device buffer or a staged table changes
    BScanParams g_cmp_args;              // argument of the current query's compare launch
    const void* g_cmp_func = nullptr;

    cudaEvent_t ev[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // [8], [9]: first / last launch of a query chain
    u64 launches = 0;
    pss_stats acc = {};          // accumulated per-kernel device time since pss_stats_reset
};

namespace {

thread_local std::string g_err;

int fail(pss_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    g_err = buf;
    return code;
}

#define CU_TRY(h, expr)                                                                                 \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(h, e__ == cudaErrorMemoryAllocation ? PSS_ENOMEM : PSS_EINTERNAL, "%s: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                                   \
    } while (0)

#define PSS_TRY(expr)              \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != PSS_OK) return rc__; \
    } while (0)

int ensure(pss_handle* h, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return PSS_OK;
    h->alloc_generation++;
    if (b.p) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        CU_TRY(h, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 16 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&b.p, want);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        b.p = nullptr;
        return fail(h, PSS_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    b.cap = want;
    return PSS_OK;
}

u32 env_u32(const char* name, u32 dflt) {
    const char* s = getenv(name);
    if (!s || !*s) return dflt;
    return (u32)strtoul(s, nullptr, 10);
}

u64 isqrt_u64(u64 x) {
    u64 r = (u64)std::sqrt((long double)x);
    while (r * r > x) --r;
    while ((r + 1) * (r + 1) <= x) ++r;
    return r;
}

// Timing events of a query.  The span events ([8] first launch, [9] end of the last kernel) are always recorded; the
// per-kernel events ([0..7]) cost ~2 us of stream time and ~2 us of host time EACH (A/B in profiles/r02: a C3 query is
// 19 us shorter on the device and 37 us shorter end to end without them, a C4 query 25 % shorter), so they are levelled:
//   0  none        1 (default)  the sieve only ([0], [1]: the dominant kernel of every query)        2  every kernel
// pss_set_kernel_events / PSS_KERNEL_EVENTS select the level; times of events that were not recorded read 0.
bool ev_on(const pss_handle* h, int i) { return i >= 8 || h->kernel_events >= 2 || (h->kernel_events == 1 && i <= 1); }
cudaError_t mark(pss_handle* h, int i) {
    if (h->dry || !ev_on(h, i)) return cudaSuccess;
    if (h->capturing) return i >= 8 ? cudaSuccess : cudaEventRecordWithFlags(h->ev[i], h->stream, cudaEventRecordExternal);
    return cudaEventRecord(h->ev[i], h->stream);
}

double ev_ms(pss_handle* h, int a, int b) {
    if (!ev_on(h, a) || !ev_on(h, b)) return 0.0;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]) != cudaSuccess) {
        cudaGetLastError();
        return 0.0;
    }
    return ms;
}

// ---- base primes (staged once, regrown only when a larger sqrt(hi) is needed) --------------------
int ensure_base_primes(pss_handle* h, u64 hi) {
    const u64 need = isqrt_u64(hi) + 1;
    const u32 first = kFirstSievePrime[h->n_pat];
    if (need > h->base_limit) {
        // simple odd-only host sieve of [0, limit]; limit <= 2^20 for hi <= 2^40.  Staging only: these are
        // the sieving primes, not part of the sieved range's output.
        u64 limit = std::max<u64>(need, 1024);
        std::vector<uint8_t> comp(limit + 1, 0);
        for (u64 i = 3; i * i <= limit; i += 2)
            if (!comp[i])
                for (u64 m = i * i; m <= limit; m += 2 * i) comp[m] = 1;
        h->h_base.clear();
        for (u64 i = 7; i <= limit; i += 2)
            if (!comp[i]) h->h_base.push_back((u32)i);
        h->base_limit = limit;
        h->base_first_prime = 0;   // force re-upload
        h->alloc_generation++;
    }
    if (h->base_first_prime != first) {
        auto it = std::lower_bound(h->h_base.begin(), h->h_base.end(), first);
        const size_t n = (size_t)(h->h_base.end() - it);
        std::vector<u32> inv(std::max<size_t>(n, 1)), strand(std::max<size_t>(n, 1) * 8);
        for (size_t i = 0; i < n; ++i) {
            const u32 p = it[i];
            inv[i] = (u32)((1ull << 32) / p);
            for (u32 j = 0; j < 8; ++j) {
                u32 c, bit;
                strand_geometry(p, j, c, bit);
                strand[i * 8 + j] = c | (bit << 24);
            }
        }
        PSS_TRY(ensure(h, h->d_base, std::max<size_t>(n, 1) * sizeof(u32)));
        PSS_TRY(ensure(h, h->d_base_inv, std::max<size_t>(n, 1) * sizeof(u32)));
        PSS_TRY(ensure(h, h->d_base_strand, std::max<size_t>(n, 1) * 8 * sizeof(u32)));
        if (n) {
            CU_TRY(h, cudaMemcpyAsync(h->d_base.p, &*it, n * sizeof(u32), cudaMemcpyHostToDevice, h->stream));
            CU_TRY(h, cudaMemcpyAsync(h->d_base_inv.p, inv.data(), n * sizeof(u32), cudaMemcpyHostToDevice, h->stream));
            CU_TRY(h, cudaMemcpyAsync(h->d_base_strand.p, strand.data(), n * 8 * sizeof(u32), cudaMemcpyHostToDevice, h->stream));
        }
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        h->n_base_dev = (u32)n;
        h->base_first_prime = first;
    }
    return PSS_OK;
}

int build_patterns(pss_handle* h) {
    for (int g = 0; g < kMaxPatterns; ++g) {
        u64 period = 1;
        for (int k = 0; k < 5; ++k)
            if (kPatGroups[g][k]) period *= kPatGroups[g][k];
        const u64 bytes = 16 * period + kPatternPad;
        CU_TRY(h, cudaMalloc(&h->d_pat[g], bytes));
        h->pat_period16[g] = 16 * period;
        k_build_pattern<<<h->sm_count * 8, 256, 0, h->stream>>>(reinterpret_cast<uint8_t*>(h->d_pat[g]), bytes, kPatGroups[g][0],
                                                                 kPatGroups[g][1], kPatGroups[g][2], kPatGroups[g][3],
                                                                 kPatGroups[g][4]);
        CU_TRY(h, cudaGetLastError());
    }
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return PSS_OK;
}

// resident CTAs per SM of `kernel` for a launch shape, queried once per handle (= per device) and kernel
template <class Kern>
int occ_of(pss_handle* h, Kern kernel, int threads, size_t smem, int* occ) {
    const void* key = reinterpret_cast<const void*>(kernel);
    auto it = h->occ_map.find(key);
    if (it != h->occ_map.end()) {
        *occ = it->second;
        return PSS_OK;
    }
    int v = 0;
    CU_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, threads, smem));
    if (v < 1) v = 1;
    h->occ_map[key] = v;
    *occ = v;
    return PSS_OK;
}

// opt-in dynamic shared memory limits of the kernels that need more than 48 KiB: cudaFuncSetAttribute acts on the CURRENT
// device only, so pss_open runs this for every handle (a second handle on another GPU would otherwise find a 48 KiB limit)
int set_kernel_attributes(pss_handle* h) {
    const int big = (int)(h->smem_optin - 1024);
    CU_TRY(h, cudaFuncSetAttribute(k_sieve<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU_TRY(h, cudaFuncSetAttribute(k_sieve<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU_TRY(h, cudaFuncSetAttribute(k_sieve<768>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU_TRY(h, cudaFuncSetAttribute(k_sieve<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU_TRY(h, cudaFuncSetAttribute(k_sieve<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    CU_TRY(h, cudaFuncSetAttribute(k_compact_lut, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kClSmemBytes));
    CU_TRY(h, cudaFuncSetAttribute(k_compact_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCwSmemBytes));
    CU_TRY(h, cudaFuncSetAttribute(k_bitmap_moments2w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMom2wSmemBytes));
    const void* momw[] = {(const void*)k_bitmap_moments_tabw<1, false>, (const void*)k_bitmap_moments_tabw<2, false>,
                          (const void*)k_bitmap_moments_tabw<3, false>, (const void*)k_bitmap_moments_tabw<4, false>,
                          (const void*)k_bitmap_moments_tabw<5, false>, (const void*)k_bitmap_moments_tabw<2, true>,
                          (const void*)k_bitmap_moments_tabw<3, true>,  (const void*)k_bitmap_moments_tabw<4, true>,
                          (const void*)k_bitmap_moments_tabw<5, true>};
    for (const void* f : momw) CU_TRY(h, cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMomwSmemBytes));
    return PSS_OK;
}

// resident CTAs per SM of k_sieve<THREADS> with a tile of tile_bytes
template <int THREADS>
int sieve_occupancy_t(pss_handle* h, u32 tile_bytes, int* occ) {
    *occ = 0;
    CU_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_sieve<THREADS>, THREADS, (size_t)tile_bytes + kSieveScratchBytes));
    return PSS_OK;
}
int sieve_occupancy(pss_handle* h, u32 threads, u32 tile_bytes, int* occ) {
    switch (threads) {
        case 256: return sieve_occupancy_t<256>(h, tile_bytes, occ);
        case 512: return sieve_occupancy_t<512>(h, tile_bytes, occ);
        case 768: return sieve_occupancy_t<768>(h, tile_bytes, occ);
        default: return sieve_occupancy_t<1024>(h, tile_bytes, occ);
    }
}

template <int THREADS>
int launch_sieve(pss_handle* h, const SieveParams& sp, u32 grid) {
    if (h->dry) return PSS_OK;
    if (sp.prof && THREADS == 1024) {
        k_sieve<1024, true><<<grid, 1024, (size_t)sp.tile_bytes + kSieveScratchBytes, h->stream>>>(sp);
    } else {
        k_sieve<THREADS><<<grid, THREADS, (size_t)sp.tile_bytes + kSieveScratchBytes, h->stream>>>(sp);
    }
    CU_TRY(h, cudaGetLastError());
    h->launches++;
    return PSS_OK;
}

// completes an asynchronous sieve_range after the stream has been synchronised
void sieve_finish(pss_handle* h) {
    if (!h->sv_pending) return;
    h->sv_pending = false;
    h->sv_total = h->sv_total_from_reduce ? 0 : h->h_back->sv_total;
    h->sv_valid = !h->sv_total_from_reduce;    // without the tile prefix the bitmap cannot be compacted
    h->acc.ms_sieve += ev_ms(h, 0, 1);
    h->acc.ms_count_scan += ev_ms(h, 1, 2);
}

// K1a + K1b: sieve [lo,hi) into the bitmap, count.  Leaves sv_* describing the bitmap.  With defer_sync the
// launches are only enqueued: the caller synchronises the stream later and calls sieve_finish().
int sieve_range(pss_handle* h, u64 lo, u64 hi, bool defer_sync = false) {
    h->sv_pending = false;
    h->sv_valid = false;
    h->bs_valid = false;
    if (hi > kMaxValue) return fail(h, PSS_EINVAL, "sieve bound %llu exceeds 2^%d", (unsigned long long)hi, PSS_MAX_PRIME_BITS);
    if (lo >= hi || hi <= 2) {
        h->sv_lo = lo; h->sv_hi = hi; h->sv_total = 0; h->sv_c235 = 0; h->sv_n_tiles = 0; h->sv_valid = true;
        return PSS_OK;
    }
    PSS_TRY(ensure_base_primes(h, hi));
    const u64 B_start = (lo / 30) & ~15ull;
    const u64 B_end = (hi + 29) / 30;
    const u64 n_bytes = B_end - B_start;
    // tile size: at most the tuned size, chosen so that the tile count is a whole number of rounds of the resident
    // grid (tiles are handed out round-robin: no partial last round, no idle SMs at the tail)
    int occ = h->occ_cache;
    if (h->occ_cache_threads != h->sieve_threads || h->occ_cache_tile != h->tile_bytes) {   // queried once per tuning
        PSS_TRY(sieve_occupancy(h, h->sieve_threads, h->tile_bytes, &occ));
        h->occ_cache = occ;
        h->occ_cache_threads = h->sieve_threads;
        h->occ_cache_tile = h->tile_bytes;
    }
    if (occ < 1) return fail(h, PSS_EINTERNAL, "sieve tile of %u bytes does not fit shared memory", h->tile_bytes);
    const u32 grid_max = (u32)(h->sm_count * occ);
    u32 tile_bytes = h->tile_bytes;
    {
        const u64 per_round = (u64)grid_max * tile_bytes;
        const u64 rounds = std::max<u64>(1, (n_bytes + per_round - 1) / per_round);
        u64 per = (n_bytes + grid_max * rounds - 1) / (grid_max * rounds);
        per = ((per + kChunkBytes - 1) / kChunkBytes) * kChunkBytes;
        per = std::max<u64>(per, 8 * 1024);
        if (per < tile_bytes) tile_bytes = (u32)per;
    }
    const u64 n_tiles64 = (n_bytes + tile_bytes - 1) / tile_bytes;
    if (n_tiles64 > 0x7FFFFFFFull) return fail(h, PSS_EINVAL, "range too large");
    const u32 n_tiles = (u32)n_tiles64;
    const u32 cpt = tile_bytes / kChunkBytes;

    PSS_TRY(ensure(h, h->bitmap, (size_t)n_tiles * tile_bytes));
    PSS_TRY(ensure(h, h->chunk_counts, (size_t)n_tiles * cpt * sizeof(u32)));
    PSS_TRY(ensure(h, h->tile_counts, (size_t)n_tiles * sizeof(u32)));
    PSS_TRY(ensure(h, h->tile_prefix, ((size_t)n_tiles + 1) * sizeof(u64)));

    SieveParams sp;
    memset(&sp, 0, sizeof sp);
    sp.lo = lo;
    sp.hi = hi;
    sp.B_start = B_start;
    sp.tile_bytes = tile_bytes;
    sp.n_tiles = n_tiles;
    sp.base_p = static_cast<const u32*>(h->d_base.p);
    sp.base_inv = static_cast<const u32*>(h->d_base_inv.p);
    sp.base_strand = static_cast<const u32*>(h->d_base_strand.p);
    sp.wide_index = (B_start + (u64)n_tiles * tile_bytes > 0xFFFFFFFFull) ? 1u : 0u;
    // only primes up to sqrt(hi) are ever needed
    {
        const u64 lim = isqrt_u64(hi - 1);
        auto b = std::lower_bound(h->h_base.begin(), h->h_base.end(), h->base_first_prime);
        auto e = std::upper_bound(h->h_base.begin(), h->h_base.end(), (u32)std::min<u64>(lim, 0xFFFFFFFFull));
        sp.n_base = (e > b) ? (u32)(e - b) : 0;
        // dense primes need at least 17 hits per lane (18 <= tile / 4p): p <= tile / 72
        const u32 warp_lim = tile_bytes / std::max<u32>(h->warp_wide_div, 72);
        sp.n_warp_wide = (e > b) ? (u32)(std::lower_bound(b, e, warp_lim) - b) : 0;
        // sparse batches: warp-uniform trip counts per strand, staged once per (tile size, prime table)
        {
            const u32 n_sparse = sp.n_base - sp.n_warp_wide;
            const u32 n_batches = (n_sparse + 31) / 32;
            std::vector<u32> geo(std::max<u32>(n_batches, 1), 0);
            for (u32 q = 0; q < n_batches; ++q) {
                const u32 p_lo = b[sp.n_warp_wide + 32 * q];
                const u32 p_hi = b[std::min<u32>(sp.n_warp_wide + 32 * q + 31, sp.n_base - 1)];
                const u32 U = tile_bytes / p_hi;
                const u32 n_uni = U > 1 ? U - 1 : 0;
                const u32 n_rag = (tile_bytes - 1) / p_lo + 1 - n_uni;
                geo[q] = n_uni | (n_rag << 16);
            }
            if (geo != h->h_batch_geo) {
                PSS_TRY(ensure(h, h->d_batch_geo, geo.size() * sizeof(u32)));
                CU_TRY(h, cudaStreamSynchronize(h->stream));   // the previous table may still be in use
                CU_TRY(h, cudaMemcpy(h->d_batch_geo.p, geo.data(), geo.size() * sizeof(u32), cudaMemcpyHostToDevice));
                h->h_batch_geo.swap(geo);
                h->alloc_generation++;
            }
            sp.batch_geo = static_cast<const u32*>(h->d_batch_geo.p);
            sp.sparse_lead = h->sparse_lead;
        }
        // per tile: how many base primes satisfy p*p < tile end (tiny device kernel, no host work)
        // (the table only depends on the geometry; PSS_NACT_CACHE=1 reuses it for a repeated query instead of launching again)
        const u64 nact_key[6] = {B_start, tile_bytes, hi, n_tiles, sp.n_base, h->base_first_prime};
        static const bool nact_cache = env_u32("PSS_NACT_CACHE", 1) != 0;   // a repeated geometry reuses the table (one launch less per query)
        if (!nact_cache || memcmp(nact_key, h->nact_key, sizeof nact_key) != 0 || !h->tile_nact.p) {
            PSS_TRY(ensure(h, h->tile_nact, (size_t)n_tiles * sizeof(u32)));
            if (h->dry)
                h->dry_miss = true;      // the captured chain relied on the cached table: this query cannot be a replay
            else
                k_tile_nact<<<(n_tiles + 255) / 256, 256, 0, h->stream>>>(sp.base_p, sp.n_base, B_start, tile_bytes, hi, n_tiles,
                                                                          static_cast<u32*>(h->tile_nact.p));
            CU_TRY(h, cudaGetLastError());
            h->launches++;
            memcpy(h->nact_key, nact_key, sizeof nact_key);
        }
        sp.tile_nact = static_cast<const u32*>(h->tile_nact.p);
    }
    sp.n_pat = h->n_pat;
    for (u32 g = 0; g < h->n_pat; ++g) {
        sp.pat[g] = h->d_pat[g];
        sp.pat_period16[g] = h->pat_period16[g];
    }
    sp.bitmap = static_cast<u32*>(h->bitmap.p);
    sp.chunk_counts = static_cast<u32*>(h->chunk_counts.p);
    sp.tile_counts = static_cast<u32*>(h->tile_counts.p);
    sp.prof = static_cast<unsigned long long*>(h->sieve_prof.p);

    CU_TRY(h, mark(h, 0));
    if (h->sieve_variant == 1) {   // per-candidate trial division (selectable cross-check; same outputs as the sieve)
        CU_TRY(h, cudaMemsetAsync(sp.tile_counts, 0, (size_t)n_tiles * sizeof(u32), h->stream));
        const u64 n_chunks = (u64)n_tiles * cpt;
        k_individual<<<(u32)std::min<u64>(n_chunks, (u64)h->sm_count * 8), kChunkWords, 0, h->stream>>>(sp);
        CU_TRY(h, cudaGetLastError());
        h->launches++;
    } else {
        switch (h->sieve_threads) {
            case 256: PSS_TRY(launch_sieve<256>(h, sp, std::min(n_tiles, grid_max))); break;
            case 512: PSS_TRY(launch_sieve<512>(h, sp, std::min(n_tiles, grid_max))); break;
            case 768: PSS_TRY(launch_sieve<768>(h, sp, std::min(n_tiles, grid_max))); break;
            default: PSS_TRY(launch_sieve<1024>(h, sp, std::min(n_tiles, grid_max))); break;
        }
    }
    CU_TRY(h, mark(h, 1));

    u64 c235 = 0;
    for (u64 q : {2ull, 3ull, 5ull})
        if (q >= lo && q < hi) ++c235;
    // the per-tile prefix of the prime counts is only needed by the compaction path; a deferred (query-chain) sieve
    // takes its total from the block-sum pass instead (reduce_finish)
    h->sv_total_from_reduce = defer_sync;
    if (!defer_sync) {
        k_scan_tile_counts<<<1, 1024, 0, h->stream>>>(sp.tile_counts, n_tiles, c235, static_cast<u64*>(h->tile_prefix.p));
        CU_TRY(h, cudaGetLastError());
        h->launches++;
    }
    CU_TRY(h, mark(h, 2));
    if (!defer_sync)
        CU_TRY(h, cudaMemcpyAsync(&h->h_back->sv_total, static_cast<u64*>(h->tile_prefix.p) + n_tiles, sizeof(u64), cudaMemcpyDeviceToHost, h->stream));
    h->sv_lo = lo; h->sv_hi = hi; h->sv_B_start = B_start; h->sv_total = 0; h->sv_c235 = c235;
    h->sv_tile_bytes = tile_bytes; h->sv_n_tiles = n_tiles;
    h->acc.sieve_lo = lo;
    h->acc.sieve_hi = hi;
    h->sv_pending = true;
    if (!defer_sync) {
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        sieve_finish(h);
    }
    return PSS_OK;
}

// K1c: compact the first min(total, cap) primes of the last sieve into the resident buffer
int compact_resident(pss_handle* h, u64 cap) {
    if (!h->sv_valid) return fail(h, PSS_EINTERNAL, "no sieve result to compact");
    const u64 n_out = std::min(h->sv_total, cap);
    h->n_resident = 0;
    if (n_out == 0) {
        CU_TRY(h, mark(h, 3));
        return PSS_OK;
    }
    PSS_TRY(ensure(h, h->primes, (size_t)n_out * sizeof(u64)));
    u64* out = static_cast<u64*>(h->primes.p);
    CU_TRY(h, mark(h, 2));
    if (h->sv_c235) {
        u64 small[3];
        u32 k = 0;
        for (u64 q : {2ull, 3ull, 5ull})
            if (q >= h->sv_lo && q < h->sv_hi && k < n_out) small[k++] = q;
        CU_TRY(h, cudaMemcpyAsync(out, small, k * sizeof(u64), cudaMemcpyHostToDevice, h->stream));
        CU_TRY(h, cudaStreamSynchronize(h->stream));
    }
    if (h->sv_n_tiles) {
        CompactParams cp;
        cp.bitmap = static_cast<const u32*>(h->bitmap.p);
        cp.chunk_counts = static_cast<const u32*>(h->chunk_counts.p);
        cp.tile_prefix = static_cast<const u64*>(h->tile_prefix.p);
        cp.chunks_per_tile = h->sv_tile_bytes / kChunkBytes;
        u32 group = 8;
        while (cp.chunks_per_tile % group) --group;
        cp.group = group;
        cp.n_groups = h->sv_n_tiles * (cp.chunks_per_tile / group);
        cp.B_start = h->sv_B_start;
        cp.out = out;
        cp.out_cap = n_out;
        cp.error_flag = &static_cast<MiscDev*>(h->misc.p)->error_flag;
        CU_TRY(h, cudaMemsetAsync(cp.error_flag, 0, sizeof(u32), h->stream));
        if (env_u32("PSS_COMPACT_BLOCK", 0)) {     // the older block-synchronous form (A/B comparison)
            const u32 grid = std::min<u32>(cp.n_groups, (u32)h->sm_count * 16);
            k_compact<<<grid, 256, 0, h->stream>>>(cp);
        } else if (!env_u32("PSS_COMPACT_WALK", 0)) {   // default: table form
            int occ_l = 1;
            PSS_TRY(occ_of(h, k_compact_lut, kCwWarps * 32, kClSmemBytes, &occ_l));
            cp.group = 1;
            cp.n_groups = h->sv_n_tiles * cp.chunks_per_tile;     // one work item per chunk
            const u32 grid = std::min<u32>((cp.n_groups + kCwWarps - 1) / kCwWarps, (u32)(h->sm_count * occ_l));
            k_compact_lut<<<grid, kCwWarps * 32, kClSmemBytes, h->stream>>>(cp);
        } else {
            int occ = 1;
            PSS_TRY(occ_of(h, k_compact_warp, kCwWarps * 32, kCwSmemBytes, &occ));
            cp.group = 1;
            cp.n_groups = h->sv_n_tiles * cp.chunks_per_tile;     // one work item per chunk
            const u32 grid = std::min<u32>((cp.n_groups + kCwWarps - 1) / kCwWarps, (u32)(h->sm_count * occ));
            k_compact_warp<<<grid, kCwWarps * 32, kCwSmemBytes, h->stream>>>(cp);
        }
        CU_TRY(h, cudaGetLastError());
        h->launches++;
    }
    CU_TRY(h, mark(h, 3));
    u32 flag = 0;
    u64 last = 0;
    CU_TRY(h, cudaMemcpyAsync(&flag, &static_cast<MiscDev*>(h->misc.p)->error_flag, sizeof(u32), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaMemcpyAsync(&last, out + (n_out - 1), sizeof(u64), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    if (flag) return fail(h, PSS_EINTERNAL, "compaction overflow (flag %u): bitmap is not a valid sieve", flag);
    h->n_resident = n_out;
    h->resident_last = last;
    h->acc.ms_compact += ev_ms(h, 2, 3);
    return PSS_OK;
}

// ---- K2/K3 dispatch --------------------------------------------------------------------------------
// column scan: one CTA per column; up to 8192 scan tiles the register-resident kernel (rows loaded once)
void launch_scan_columns(pss_handle* h, u32 ncols, const u64* in, u64* out, u32 n, u64 stride) {
    const u32 rows = (n + 31u) / 32u;
    if (h->dry) return;
    if ((rows + 15u) / 16u <= 16u) k_scan_columns_small<<<ncols, 512, 0, h->stream>>>(in, out, n, stride);
    else k_scan_columns<<<ncols, 1024, 0, h->stream>>>(in, out, n, stride);
}

// pass A (tile sums) -> column scan -> pass B (walk; compare / prefix modes only).  Leaves the column totals on their way
// to h->h_back->col_totals (TOTAL entries, deferred-carry lanes).
template <int K, bool MULTI, int MODE, int ITEMS>
int launch_scan_t(pss_handle* h, ScanParams& sp) {
    using Lay = Layout<K, MULTI>;
    constexpr int TILE = kScanThreads * ITEMS;
    const u64 n_tiles64 = (sp.count + TILE - 1) / TILE;
    if (n_tiles64 > 0x7FFFFFFFull) return fail(h, PSS_EINVAL, "range too large for one scan");
    sp.n_tiles = (u32)n_tiles64;
    sp.col_stride = (u64)sp.n_tiles + 1;
    const size_t col_bytes = (size_t)Lay::TOTAL * sp.col_stride * sizeof(u64);
    PSS_TRY(ensure(h, h->agg, col_bytes));
    PSS_TRY(ensure(h, h->incl, col_bytes));
    sp.cols = static_cast<u64*>(h->agg.p);
    sp.col_prefix = static_cast<const u64*>(h->incl.p);
    int occ = 1;
    PSS_TRY(occ_of(h, k_power_tiles<K, MULTI, ITEMS>, kScanThreads, 0, &occ));
    k_power_tiles<K, MULTI, ITEMS><<<std::min<u32>(sp.n_tiles, (u32)(h->sm_count * occ)), kScanThreads, 0, h->stream>>>(sp);
    CU_TRY(h, cudaGetLastError());
    launch_scan_columns(h, (u32)Lay::TOTAL, sp.cols, static_cast<u64*>(h->incl.p), sp.n_tiles, sp.col_stride);
    CU_TRY(h, cudaGetLastError());
    h->launches += 2;
    if constexpr (MODE != kModeReduce) {
        int occ_w = 1;
        PSS_TRY(occ_of(h, k_power_walk<K, MULTI, MODE, ITEMS>, kScanThreads, 0, &occ_w));
        k_power_walk<K, MULTI, MODE, ITEMS><<<std::min<u32>(sp.n_tiles, (u32)(h->sm_count * occ_w)), kScanThreads, 0, h->stream>>>(sp);
        CU_TRY(h, cudaGetLastError());
        h->launches++;
    }
    CU_TRY(h, cudaMemcpy2DAsync(h->h_back->col_totals, sizeof(u64), static_cast<const u64*>(h->incl.p) + sp.n_tiles, sp.col_stride * sizeof(u64),
                                sizeof(u64), (size_t)Lay::TOTAL, cudaMemcpyDeviceToHost, h->stream));
    return PSS_OK;
}

template <int MODE>
int launch_scan(pss_handle* h, ScanParams& sp, u32 K, bool multi) {
    if (!multi) {
        switch (K) {
            case 1: return launch_scan_t<1, false, MODE, 16>(h, sp);
            case 2: return launch_scan_t<2, false, MODE, 16>(h, sp);
            case 3: return launch_scan_t<3, false, MODE, 16>(h, sp);
            case 4: return launch_scan_t<4, false, MODE, 8>(h, sp);
            case 5: return launch_scan_t<5, false, MODE, 8>(h, sp);
            case 6: return launch_scan_t<6, false, MODE, 8>(h, sp);
            case 7: return launch_scan_t<7, false, MODE, 8>(h, sp);
            case 8: return launch_scan_t<8, false, MODE, 8>(h, sp);
        }
    } else if constexpr (MODE != kModePrefix) {
        switch (K) {
            case 2: return launch_scan_t<2, true, MODE, 8>(h, sp);
            case 3: return launch_scan_t<3, true, MODE, 8>(h, sp);
            case 4: return launch_scan_t<4, true, MODE, 8>(h, sp);
            case 5: return launch_scan_t<5, true, MODE, 8>(h, sp);
            case 6: return launch_scan_t<6, true, MODE, 8>(h, sp);
        }
    }
    return fail(h, PSS_EINVAL, "unsupported power configuration k=%u all_powers=%d", K, (int)multi);
}

inline u32 n_powers_of(u32 K, bool multi) { return multi ? K : 1; }
inline u32 expo_of(u32 K, bool multi, u32 j) { return multi ? j + 1 : K; }

// Core scan over device values.  carry_in: n_powers * PSS_LIMBS limbs (host) or NULL.
int run_scan(pss_handle* h, const u64* d_vals, u64 count, u64 n_first, const u32* carry_in, u32 K, bool multi, int mode,
             u64 n_min, u64 n_max, u32 cmp, const u32* target, u32 target_nlimbs, u32* d_prefix_out,
             pss_power_result* out /* n_powers */) {
    if (K < 1 || K > PSS_MAX_POWER) return fail(h, PSS_EINVAL, "power %u out of range 1..%d", K, PSS_MAX_POWER);
    if (multi && K > 6) return fail(h, PSS_EINVAL, "all_powers supports power_max <= 6");
    if (multi && K == 1) multi = false;
    const u32 np = n_powers_of(K, multi);
    static const u32 kMasks[6] = {2u, 5u, 1u, 3u, 4u, 6u};   // EQ NE LT LE GT GE over bits {lt,eq,gt}
    if (cmp > 5) return fail(h, PSS_EINVAL, "invalid comparison operator %u", cmp);

    MiscDev* hm = h->h_misc;
    memset(hm, 0, sizeof(MiscDev));
    for (u32 j = 0; j < np; ++j) {
        hm->results[j].power = expo_of(K, multi, j);
        hm->results[j].first_match_n = ~0ull;
        if (carry_in) {
            memcpy(&hm->carry[j * PSS_LIMBS], carry_in + j * PSS_LIMBS, PSS_LIMBS * sizeof(u32));
            const int w = sum_limbs((int)expo_of(K, multi, j));
            for (int i = w; i < PSS_LIMBS; ++i)
                if (hm->carry[j * PSS_LIMBS + i]) return fail(h, PSS_EINVAL, "carry-in of power %u exceeds %d limbs", hm->results[j].power, w);
        }
        memcpy(hm->results[j].final_sum, &hm->carry[j * PSS_LIMBS], PSS_LIMBS * sizeof(u32));
    }
    ScanParams sp;
    memset(&sp, 0, sizeof sp);
    sp.primes = d_vals;
    sp.count = count;
    sp.n_first = n_first;
    sp.n_min = n_min;
    sp.n_max = n_max;
    sp.cmp_mask = kMasks[cmp];
    if (mode == kModeCompare) {
        for (u32 i = 0; i < target_nlimbs && i < PSS_LIMBS; ++i) sp.target[i] = target[i];
        for (u32 j = 0; j < np; ++j) {
            const u32 w = (u32)sum_limbs((int)expo_of(K, multi, j));
            bool over = false;
            for (u32 i = w; i < target_nlimbs; ++i)
                if (target[i]) over = true;
            if (over) sp.target_over_mask |= (1u << j);
        }
    }
    MiscDev* dm = static_cast<MiscDev*>(h->misc.p);
    sp.carry_in = dm->carry;
    sp.error_flag = &dm->error_flag;
    sp.results = dm->results;
    sp.prefix_out = d_prefix_out;
    sp.exhaustive = h->compare_exhaustive ? 1u : 0u;

    CU_TRY(h, mark(h, 4));
    if (count > 0) {
        CU_TRY(h, cudaMemcpyAsync(dm, hm, sizeof(MiscDev), cudaMemcpyHostToDevice, h->stream));
        switch (mode) {
            case kModeReduce: PSS_TRY(launch_scan<kModeReduce>(h, sp, K, multi)); break;
            case kModeCompare: PSS_TRY(launch_scan<kModeCompare>(h, sp, K, multi)); break;
            default: PSS_TRY(launch_scan<kModePrefix>(h, sp, K, multi)); break;
        }
        CU_TRY(h, mark(h, 5));
        CU_TRY(h, cudaMemcpyAsync(hm, dm, sizeof(MiscDev), cudaMemcpyDeviceToHost, h->stream));
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        if (hm->error_flag) return fail(h, PSS_EINTERNAL, "scan kernel reported error flag %u", hm->error_flag);
        // S_k(last element) = carry-in + column totals (deferred-carry lanes, normalised here)
        const u64* totals = h->h_back->col_totals;
        u32 col = 0;
        for (u32 j = 0; j < np; ++j) {
            const u32 w = (u32)sum_limbs((int)expo_of(K, multi, j));
            u64 c = 0;
            for (u32 i = 0; i < (u32)PSS_LIMBS; ++i) {
                const u64 v = (i < w) ? totals[col + i] : 0ull;
                const u64 lo = (v & 0xFFFFFFFFull) + (c & 0xFFFFFFFFull) + hm->carry[j * PSS_LIMBS + i];
                hm->results[j].final_sum[i] = (u32)lo;
                c = (v >> 32) + (c >> 32) + (lo >> 32);
            }
            col += w;
        }
    } else {
        CU_TRY(h, mark(h, 5));
        CU_TRY(h, cudaStreamSynchronize(h->stream));
    }
    for (u32 j = 0; j < np; ++j) {
        out[j] = hm->results[j];
        if (out[j].match_count == 0) out[j].first_match_n = 0, out[j].last_match_n = 0;
    }
    h->acc.ms_power_scan += ev_ms(h, 4, 5);
    h->acc.primes += count;
    return PSS_OK;
}


// ---- fused path: K2/K3 straight from the bitmap -------------------------------------------------------
template <int K, bool MULTI, int MODE>
int launch_bscan_t(pss_handle* h, const BScanParams& bp) {
    int occ = 1;
    PSS_TRY(occ_of(h, k_bitmap_scan<K, MULTI, MODE>, kBsThreads, 0, &occ));
    const u32 grid = std::min<u32>(bp.n_tiles, (u32)(h->sm_count * occ));
    if constexpr (MODE == kBsCompare) {     // the one launch of a query chain whose argument changes from query to query
        h->g_cmp_args = bp;
        h->g_cmp_func = reinterpret_cast<const void*>(k_bitmap_scan<K, MULTI, MODE>);
    }
    if (h->dry) return PSS_OK;
    k_bitmap_scan<K, MULTI, MODE><<<grid, kBsThreads, 0, h->stream>>>(bp);
    CU_TRY(h, cudaGetLastError());
    h->launches++;
    return PSS_OK;
}

template <int K, bool MULTI>
int launch_moments_t(pss_handle* h, const BScanParams& bp) {
    static const bool moments_wide = env_u32("PSS_MOMENTS_WIDE", 1) != 0;
    // position-specific byte tables (128 KiB, one 768-thread CTA per SM; 64-bit entries for k <= 3, 128-bit for k = 4, 5; k = 2
    // alone has its own kernel): only where every SM gets at least two rounds of warp tiles -- below that the 128 KiB table
    // fill per CTA costs more than the shorter loop saves (tools/moments_probe.py: 46 / 54 / 69 us against 48 / 64 / 78 us at
    // 9.8e8 for k = 4, 5 and all powers 1..5; 24 - 30 against 20 - 24 us at 8.6e7).  PSS_MOMENTS_WIDE_MIN_K raises the lowest k.
    const u64 n_warp_tiles = (u64)bp.n_tiles * (kBsThreads / 32);
    static const u32 moments_wide_min_k = env_u32("PSS_MOMENTS_WIDE_MIN_K", 1);
    if (h->reduce_moments >= 2 && moments_wide && K >= (int)moments_wide_min_k && n_warp_tiles >= 2ull * h->sm_count * (kMomwThreads / 32)) {
        const u32 grid_w = std::min<u32>((bp.n_tiles * (kBsThreads / 32) + kMomwThreads / 32 - 1) / (kMomwThreads / 32), (u32)h->sm_count);
        if (h->dry) return PSS_OK;
        k_bitmap_moments_tabw<K, MULTI><<<grid_w, kMomwThreads, kMomwSmemBytes, h->stream>>>(bp);
        CU_TRY(h, cudaGetLastError());
        h->launches++;
        return PSS_OK;
    }
    if (h->reduce_moments >= 2) {   // byte-table form; PSS_REDUCE_MOMENTS=1 selects the per-prime walk
        int occ_t = 1;
        PSS_TRY(occ_of(h, k_bitmap_moments_tab<K, MULTI>, kBsThreads, 0, &occ_t));
        const u32 grid_t = std::min<u32>(bp.n_tiles, (u32)(h->sm_count * occ_t));
        if (h->dry) return PSS_OK;
        k_bitmap_moments_tab<K, MULTI><<<grid_t, kBsThreads, 0, h->stream>>>(bp);
        CU_TRY(h, cudaGetLastError());
        h->launches++;
        return PSS_OK;
    }
    int occ = 1;
    PSS_TRY(occ_of(h, k_bitmap_moments<K, MULTI>, kBsThreads, 0, &occ));
    const u32 grid = std::min<u32>(bp.n_tiles, (u32)(h->sm_count * occ));
    if (h->dry) return PSS_OK;
    k_bitmap_moments<K, MULTI><<<grid, kBsThreads, 0, h->stream>>>(bp);
    CU_TRY(h, cudaGetLastError());
    h->launches++;
    return PSS_OK;
}
int launch_moments(pss_handle* h, const BScanParams& bp, u32 K, bool multi) {
    if (!multi) {
        switch (K) {
            case 1: return launch_moments_t<1, false>(h, bp);
            case 2: return launch_moments_t<2, false>(h, bp);
            case 3: return launch_moments_t<3, false>(h, bp);
            case 4: return launch_moments_t<4, false>(h, bp);
            case 5: return launch_moments_t<5, false>(h, bp);
        }
    } else {
        switch (K) {
            case 2: return launch_moments_t<2, true>(h, bp);
            case 3: return launch_moments_t<3, true>(h, bp);
            case 4: return launch_moments_t<4, true>(h, bp);
            case 5: return launch_moments_t<5, true>(h, bp);
        }
    }
    return fail(h, PSS_EINVAL, "unsupported power");
}

template <int MODE>
int launch_bscan(pss_handle* h, const BScanParams& bp, u32 K, bool multi) {
    if (!multi) {
        switch (K) {
            case 1: return launch_bscan_t<1, false, MODE>(h, bp);
            case 2: return launch_bscan_t<2, false, MODE>(h, bp);
            case 3: return launch_bscan_t<3, false, MODE>(h, bp);
            case 4: return launch_bscan_t<4, false, MODE>(h, bp);
            case 5: return launch_bscan_t<5, false, MODE>(h, bp);
            case 6: return launch_bscan_t<6, false, MODE>(h, bp);
            case 7: return launch_bscan_t<7, false, MODE>(h, bp);
            case 8: return launch_bscan_t<8, false, MODE>(h, bp);
        }
    } else {
        switch (K) {
            case 2: return launch_bscan_t<2, true, MODE>(h, bp);
            case 3: return launch_bscan_t<3, true, MODE>(h, bp);
            case 4: return launch_bscan_t<4, true, MODE>(h, bp);
            case 5: return launch_bscan_t<5, true, MODE>(h, bp);
            case 6: return launch_bscan_t<6, true, MODE>(h, bp);
        }
    }
    return fail(h, PSS_EINVAL, "unsupported power configuration k=%u all_powers=%d", K, (int)multi);
}

int launch_bscan_tri(pss_handle* h, const BScanParams& bp, u32 K) {
    switch (K) {
        case 1: return launch_bscan_t<1, false, kBsTri>(h, bp);
        case 2: return launch_bscan_t<2, false, kBsTri>(h, bp);
        case 3: return launch_bscan_t<3, false, kBsTri>(h, bp);
        case 4: return launch_bscan_t<4, false, kBsTri>(h, bp);
        case 5: return launch_bscan_t<5, false, kBsTri>(h, bp);
        case 6: return launch_bscan_t<6, false, kBsTri>(h, bp);
        case 7: return launch_bscan_t<7, false, kBsTri>(h, bp);
        case 8: return launch_bscan_t<8, false, kBsTri>(h, bp);
    }
    return fail(h, PSS_EINVAL, "unsupported power %u", K);
}

u32 total_limbs_of(u32 K, bool multi) {
    u32 t = 0;
    for (u32 j = 0; j < n_powers_of(K, multi); ++j) t += (u32)sum_limbs((int)expo_of(K, multi, j));
    return t;
}

void fill_bscan_common(pss_handle* h, BScanParams& bp) {
    memset(&bp, 0, sizeof bp);
    bp.bitmap = static_cast<const u32*>(h->bitmap.p);
    bp.n_words = (u64)h->sv_n_tiles * h->sv_tile_bytes / 4;
    bp.B_start = h->sv_B_start;
    bp.n_tiles = (u32)((bp.n_words + kBsTileWords - 1) / kBsTileWords);
    bp.n_small = 0;
    for (u64 q : {2ull, 3ull, 5ull})
        if (q >= h->sv_lo && q < h->sv_hi) bp.small_primes[bp.n_small++] = q;
    bp.col_stride = (u64)bp.n_tiles + 1;
    bp.thread_stride = (u64)bp.n_tiles * kBsThreads;
    bp.thread_rec = static_cast<u32*>(h->thread_rec.p);
    bp.warp_stride = (u64)bp.n_tiles * (kBsThreads / 32);
    bp.warp_rec = static_cast<u32*>(h->warp_rec.p);
}

// enqueue the D2H of the column totals of the last bitmap_reduce (once per reduce)
int copy_totals(pss_handle* h) {
    if (!h->bs_pending || h->bs_totals_copied || h->bs_n_tiles == 0) return PSS_OK;
    CU_TRY(h, cudaMemcpy2DAsync(h->h_back->col_totals, sizeof(u64), static_cast<const u64*>(h->col_prefix.p) + h->bs_n_tiles,
                                h->bs_col_stride * sizeof(u64), sizeof(u64), h->bs_total_cols, cudaMemcpyDeviceToHost, h->stream));
    h->bs_totals_copied = true;
    return PSS_OK;
}

// completes an asynchronous bitmap_reduce after the stream has been synchronised
int reduce_finish(pss_handle* h) {
    if (!h->bs_pending) return PSS_OK;
    h->bs_pending = false;
    const u32 K = h->bs_K;
    const bool multi = h->bs_multi;
    const u32 np = n_powers_of(K, multi);
    h->acc.ms_power_scan += ev_ms(h, 4, 5);
    h->acc.ms_bitmap_reduce += ev_ms(h, 4, 5);
    // normalise the deferred-carry column totals into limbs
    const u64* totals = h->h_back->col_totals;
    h->bs_count = totals[0];
    u32 col = 1;
    for (u32 j = 0; j < np; ++j) {
        const u32 w = (u32)sum_limbs((int)expo_of(K, multi, j));
        u64 c = 0;
        for (u32 i = 0; i < w; ++i, ++col) {
            const u64 v = totals[col];
            const u64 lo = (v & 0xFFFFFFFFull) + (c & 0xFFFFFFFFull);
            h->bs_sums[j * PSS_LIMBS + i] = (u32)lo;
            c = (v >> 32) + (c >> 32) + (lo >> 32);
        }
    }
    if (h->sv_total_from_reduce) {
        h->sv_total = h->bs_count;
    } else if (h->bs_count != h->sv_total) {
        return fail(h, PSS_EINTERNAL, "bitmap reduce counted %llu primes, sieve counted %llu", (unsigned long long)h->bs_count,
                    (unsigned long long)h->sv_total);
    }
    h->bs_valid = true;
    return PSS_OK;
}

// phase A of the fused path: per-tile (count, sums) records, column prefixes, slice totals.  The sieve may still be
// pending (sv_pending): only its host-known geometry is used to enqueue the launches.
int bitmap_reduce(pss_handle* h, u32 K, bool multi, bool defer_sync = false) {
    if (!h->sv_valid && !h->sv_pending) return fail(h, PSS_EINTERNAL, "no sieve result");
    if (K < 1 || K > PSS_MAX_POWER) return fail(h, PSS_EINVAL, "power %u out of range 1..%d", K, PSS_MAX_POWER);
    if (multi && K > 6) return fail(h, PSS_EINVAL, "all_powers supports power_max <= 6");
    if (multi && K == 1) multi = false;
    h->bs_valid = false;
    h->bs_pending = false;
    const u32 total = total_limbs_of(K, multi), ncols = 1 + total;
    memset(h->bs_sums, 0, sizeof h->bs_sums);
    h->bs_count = 0;
    h->bs_K = K; h->bs_multi = multi; h->bs_total_cols = ncols;
    if (h->sv_n_tiles == 0) {   // empty range
        if (h->sv_pending) {
            CU_TRY(h, cudaStreamSynchronize(h->stream));
            sieve_finish(h);
        }
        h->bs_n_tiles = 0; h->bs_col_stride = 1; h->bs_valid = true;
        return PSS_OK;
    }
    BScanParams bp;
    fill_bscan_common(h, bp);
    PSS_TRY(ensure(h, h->thread_rec, (size_t)ncols * bp.thread_stride * sizeof(u32)));
    bp.thread_rec = static_cast<u32*>(h->thread_rec.p);
    PSS_TRY(ensure(h, h->warp_rec, (size_t)ncols * bp.warp_stride * sizeof(u32)));
    bp.warp_rec = static_cast<u32*>(h->warp_rec.p);
    PSS_TRY(ensure(h, h->cols, (size_t)ncols * bp.col_stride * sizeof(u64)));
    PSS_TRY(ensure(h, h->col_prefix, (size_t)ncols * bp.col_stride * sizeof(u64)));
    bp.cols = static_cast<u64*>(h->cols.p);
    bp.col_prefix = static_cast<const u64*>(h->col_prefix.p);
    CU_TRY(h, mark(h, 4));
    if (!h->dry) CU_TRY(h, cudaMemsetAsync(bp.cols, 0, (size_t)ncols * bp.col_stride * sizeof(u64), h->stream));
    static const bool moments2_wide = env_u32("PSS_MOMENTS2_WIDE", 1) != 0;
    if (K == 2 && !multi && h->reduce_moments && moments2_wide) {
        // one 1024-thread CTA per SM (128 KiB of tables); a scan tile is 8 warp tiles whatever the CTA size
        const u32 grid = std::min<u32>((bp.n_tiles * (kBsThreads / 32) + kMom2wThreads / 32 - 1) / (kMom2wThreads / 32), (u32)h->sm_count);
        if (!h->dry) k_bitmap_moments2w<<<grid, kMom2wThreads, kMom2wSmemBytes, h->stream>>>(bp);
        CU_TRY(h, cudaGetLastError());
        h->launches++;
    } else if (K == 2 && !multi && h->reduce_moments) {
        int occ = 1;
        PSS_TRY(occ_of(h, k_bitmap_moments2, kBsThreads, 0, &occ));
        const u32 grid = std::min<u32>(bp.n_tiles, (u32)(h->sm_count * occ));
        if (!h->dry) k_bitmap_moments2<<<grid, kBsThreads, 0, h->stream>>>(bp);
        CU_TRY(h, cudaGetLastError());
        h->launches++;
    } else if (K <= 5 && h->reduce_moments) {
        PSS_TRY(launch_moments(h, bp, K, multi));
    } else {
        bp.reduce_chain = 1u;
        PSS_TRY(launch_bscan<kBsReduce>(h, bp, K, multi));
    }
    launch_scan_columns(h, ncols, bp.cols, static_cast<u64*>(h->col_prefix.p), bp.n_tiles, bp.col_stride);
    CU_TRY(h, cudaGetLastError());
    h->launches++;
    CU_TRY(h, mark(h, 5));
    h->bs_n_tiles = bp.n_tiles;
    h->bs_col_stride = bp.col_stride;
    h->bs_pending = true;
    h->bs_totals_copied = false;
    // A copy-engine operation between two kernels costs ~8 us of stream latency (measured: profiles/r02 timeline), so in a
    // query chain the read-back of the column totals is enqueued after the LAST kernel (copy_totals), not here.
    if (!defer_sync) {
        PSS_TRY(copy_totals(h));
        CU_TRY(h, cudaStreamSynchronize(h->stream));
        sieve_finish(h);
        PSS_TRY(reduce_finish(h));
    }
    return PSS_OK;
}

void limbs_add(u32* a, const u32* b) {
    u64 c = 0;
    for (int i = 0; i < PSS_LIMBS; ++i) {
        const u64 t = (u64)a[i] + b[i] + c;
        a[i] = (u32)t;
        c = t >> 32;
    }
}

// Host part of a compare launch that does not depend on the sieve: the MiscDev block (result records, carry-in limbs) is
// filled and its H2D copy enqueued.  A query chain calls this BEFORE the sieve so that the copy engine works under the
// sieve kernel instead of sitting between the column scan and the compare kernel.
int prepare_misc(pss_handle* h, const pss_search_args* a, const u32* carry_in) {
    const bool multi = a->all_powers && a->power_max > 1;
    const u32 K = a->power_max, np = n_powers_of(K, multi);
    MiscInit in;
    memset(&in, 0, sizeof in);
    in.np = np;
    for (u32 j = 0; j < np; ++j) {
        in.power[j] = expo_of(K, multi, j);
        if (carry_in) {
            memcpy(&in.carry[j * PSS_LIMBS], carry_in + j * PSS_LIMBS, PSS_LIMBS * sizeof(u32));
            const int w = sum_limbs((int)expo_of(K, multi, j));
            for (int i = w; i < PSS_LIMBS; ++i)
                if (in.carry[j * PSS_LIMBS + i]) return fail(h, PSS_EINVAL, "carry-in of power %u exceeds %d limbs", in.power[j], w);
        }
    }
    CU_TRY(h, mark(h, 8));     // a query chain starts here (recorded around the graph launch when the chain is a graph)
    if (!h->dry) k_init_misc<<<1, 128, 0, h->stream>>>(static_cast<MiscDev*>(h->misc.p), in);
    CU_TRY(h, cudaGetLastError());
    h->launches++;
    h->misc_uploaded = true;
    return PSS_OK;
}

// last launch of a query chain: MiscDev block, pending column totals and (peer path) all ranks' outcome records -> mapped
// pinned host memory; the caller synchronises the stream afterwards
int export_results(pss_handle* h, u32 peer_world, u64 peer_parity) {
    ExportParams ep;
    memset(&ep, 0, sizeof ep);
    ep.dm = static_cast<const MiscDev*>(h->misc.p);
    ep.hm = h->h_misc_dev;
    if (h->bs_pending && !h->bs_totals_copied && h->bs_n_tiles > 0) {
        ep.col_prefix = static_cast<const u64*>(h->col_prefix.p);
        ep.col_stride = h->bs_col_stride;
        ep.n_tiles = h->bs_n_tiles;
        ep.ncols = h->bs_total_cols;
        ep.totals_host = h->h_back_dev->col_totals;
        h->bs_totals_copied = true;
    }
    if (peer_world) {
        ep.peer = &h->box_local->r[peer_parity][0];
        ep.peer_host = h->h_peer_results_dev;
        ep.world = peer_world;
    }
    if (!h->dry) k_export<<<1, 256, 0, h->stream>>>(ep);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, mark(h, 9));     // ... and ends here
    h->launches++;
    return PSS_OK;
}

// phase B of the fused path: compare every partial sum, seeded with carry_in (host limbs or NULL).  Two halves so that a
// query chain can be captured into / replayed from a CUDA graph: the enqueue half only launches, the finish half runs after
// the stream has been synchronised.
int bitmap_compare_enqueue(pss_handle* h, u64 n_first, const u32* carry_in, const pss_search_args* a, u64* scanned_out, bool* pending_out) {
    const bool multi = a->all_powers && a->power_max > 1;
    const u32 K = a->power_max;
    const bool pending = h->bs_pending;   // single-query chain: the sieve and the reduce are enqueued, not read back yet
    if ((!h->bs_valid && !pending) || h->bs_K != K || h->bs_multi != multi)
        return fail(h, PSS_EINVAL, "pss_slice_sieve_sums must be called first with the same power configuration");
    const u32 np = n_powers_of(K, multi);
    static const u32 kMasks[6] = {2u, 5u, 1u, 3u, 4u, 6u};
    if (!h->misc_uploaded) PSS_TRY(prepare_misc(h, a, carry_in));
    h->misc_uploaded = false;
    // with a pending reduce the slice count is not known yet: the caller (pss_search) guarantees by a proven bound
    // that the slice holds n_max, and checks it after the synchronisation
    u64 scanned = (a->n_max >= n_first) ? (pending ? a->n_max - n_first + 1 : std::min<u64>(h->bs_count, a->n_max - n_first + 1)) : 0;
    if (h->bs_n_tiles == 0) scanned = 0;
    CU_TRY(h, mark(h, 6));
    if (scanned > 0) {
        BScanParams bp;
        fill_bscan_common(h, bp);
        bp.cols = static_cast<u64*>(h->cols.p);
        bp.col_prefix = static_cast<const u64*>(h->col_prefix.p);
        bp.n_first = n_first;
        bp.n_min = a->n_min;
        bp.n_max = a->n_max;
        bp.cmp_mask = kMasks[a->cmp];
        bp.exhaustive = h->compare_exhaustive ? 1u : 0u;
        for (u32 i = 0; i < a->target_nlimbs && i < PSS_LIMBS; ++i) bp.target[i] = a->target[i];
        for (u32 j = 0; j < np; ++j) {
            const u32 w = (u32)sum_limbs((int)expo_of(K, multi, j));
            for (u32 i = w; i < a->target_nlimbs; ++i)
                if (a->target[i]) bp.target_over_mask |= (1u << j);
        }
        MiscDev* dm = static_cast<MiscDev*>(h->misc.p);
        bp.carry_in = dm->carry;
        bp.results = dm->results;
        bp.last_prime_out = &dm->last_prime;
        bp.error_flag = &dm->error_flag;
        PSS_TRY(launch_bscan<kBsCompare>(h, bp, K, multi));
    }
    CU_TRY(h, mark(h, 7));
    PSS_TRY(export_results(h, 0, 0));
    *scanned_out = scanned;
    *pending_out = pending;
    return PSS_OK;
}

int bitmap_compare_finish(pss_handle* h, u64 n_first, const pss_search_args* a, pss_search_result* res, u64 scanned, bool pending) {
    const bool multi = a->all_powers && a->power_max > 1;
    const u32 K = a->power_max, np = n_powers_of(K, multi);
    MiscDev* hm = h->h_misc;
    if (pending) {
        sieve_finish(h);
        PSS_TRY(reduce_finish(h));
        if (h->sv_total < a->n_max - n_first + 1)
            return fail(h, PSS_EINTERNAL, "prime count bound failed: the slice holds %llu primes, n_max = %llu", (unsigned long long)h->sv_total,
                        (unsigned long long)a->n_max);
    }
    h->acc.ms_power_scan += ev_ms(h, 6, 7);
    h->acc.ms_bitmap_compare += ev_ms(h, 6, 7);
    if (pending) h->acc.ms_span += ev_ms(h, 8, 9);   // first launch (k_init_misc) -> end of the last kernel (k_export): the whole query
    h->acc.primes += scanned;
    const bool reached_n_max = scanned > 0 && (n_first + scanned - 1 == a->n_max);
    for (u32 j = 0; j < np; ++j) {
        res->powers[j] = hm->results[j];
        if (res->powers[j].match_count == 0) res->powers[j].first_match_n = 0, res->powers[j].last_match_n = 0;
        if (!reached_n_max) {   // slice exhausted before n_max: S(last scanned) = carry + slice sums
            u32 fs[PSS_LIMBS];
            memcpy(fs, &hm->carry[j * PSS_LIMBS], sizeof fs);
            if (scanned) limbs_add(fs, &h->bs_sums[j * PSS_LIMBS]);
            memcpy(res->powers[j].final_sum, fs, sizeof fs);
        }
    }
    res->n_powers = np;
    res->primes_scanned = scanned;
    res->last_prime = reached_n_max ? hm->last_prime : 0;
    return PSS_OK;
}

int bitmap_compare(pss_handle* h, u64 n_first, const u32* carry_in, const pss_search_args* a, pss_search_result* res) {
    u64 scanned = 0;
    bool pending = false;
    PSS_TRY(bitmap_compare_enqueue(h, n_first, carry_in, a, &scanned, &pending));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return bitmap_compare_finish(h, n_first, a, res, scanned, pending);
}

// BScanParams of a compare launch over the bitmap left by sieve_range / bitmap_reduce
void fill_compare_params(pss_handle* h, BScanParams& bp, const pss_search_args* a, u64 n_first, MiscDev* dm = nullptr) {
    static const u32 kMasks[6] = {2u, 5u, 1u, 3u, 4u, 6u};
    const bool multi = a->all_powers && a->power_max > 1;
    const u32 K = a->power_max, np = n_powers_of(K, multi);
    fill_bscan_common(h, bp);
    bp.cols = static_cast<u64*>(h->cols.p);
    bp.col_prefix = static_cast<const u64*>(h->col_prefix.p);
    bp.n_first = n_first;
    bp.n_min = a->n_min;
    bp.n_max = a->n_max;
    bp.cmp_mask = kMasks[a->cmp];
    bp.exhaustive = h->compare_exhaustive ? 1u : 0u;
    for (u32 i = 0; i < a->target_nlimbs && i < PSS_LIMBS; ++i) bp.target[i] = a->target[i];
    for (u32 j = 0; j < np; ++j) {
        const u32 w = (u32)sum_limbs((int)expo_of(K, multi, j));
        for (u32 i = w; i < a->target_nlimbs; ++i)
            if (a->target[i]) bp.target_over_mask |= (1u << j);
    }
    if (!dm) dm = static_cast<MiscDev*>(h->misc.p);
    bp.carry_in = dm->carry;
    bp.results = dm->results;
    bp.last_prime_out = &dm->last_prime;
    bp.error_flag = &dm->error_flag;
}

// The launch chain of a single-GPU query: init -> sieve -> block sums -> column scan -> compare -> export, enqueued back to back
// (no copy-engine operation inside, the small H2D / D2H traffic travels as kernel arguments / mapped pinned stores).
int enqueue_query_chain(pss_handle* h, const pss_search_args* a, u64 bound, u64* scanned, bool* pending) {
    const bool multi = a->all_powers && a->power_max > 1;
    PSS_TRY(prepare_misc(h, a, nullptr));
    PSS_TRY(sieve_range(h, 0, bound, true));
    PSS_TRY(bitmap_reduce(h, a->power_max, multi, true));
    return bitmap_compare_enqueue(h, 1, nullptr, a, scanned, pending);
}

void drop_query_graph(pss_handle* h) {
    if (h->g_exec) cudaGraphExecDestroy(h->g_exec);
    if (h->g_graph) cudaGraphDestroy(h->g_graph);
    h->g_graph = nullptr;
    h->g_exec = nullptr;
    h->g_cmp_node = nullptr;
    h->g_key.clear();
}

// Runs the chain kernel by kernel, or — from the second consecutive query with the same geometry on — as a CUDA graph: the
// first repeat captures the chain (stream capture of the very same enqueue code) and instantiates it; later queries with that
// geometry run the enqueue code "dry" (host bookkeeping only), point the compare node at the new argument (target, window,
// operator) and launch the graph: one driver call instead of ten.  Any failure on the graph route drops the graph, switches
// the feature off for the handle and runs the query kernel by kernel.
int run_query_chain(pss_handle* h, const pss_search_args* a, pss_search_result* res, u64 bound) {
    const bool multi = a->all_powers && a->power_max > 1;
    std::vector<u64> key = {bound, a->power_max, (u64)multi, (u64)h->kernel_events, h->tile_bytes, h->sieve_threads, h->n_pat,
                            h->warp_wide_div, (u64)h->reduce_moments, (u64)h->sieve_variant, h->alloc_generation,
                            (u64)(h->sieve_prof.p != nullptr)};
    u64 scanned = 0;
    bool pending = false;
    if (h->use_graph && h->g_exec && key == h->g_key) {
        // ---- replay
        const u64 launches_before = h->launches;
        h->dry = true;
        h->dry_miss = false;
        const int rc = enqueue_query_chain(h, a, bound, &scanned, &pending);
        h->dry = false;
        h->launches = launches_before;
        PSS_TRY(rc);
        if (h->dry_miss || key[10] != h->alloc_generation) {
            // a staged table or buffer changed since the capture (another geometry ran in between): this query runs kernel
            // by kernel and the graph is rebuilt on the next repeat
            drop_query_graph(h);
            h->sv_pending = h->bs_pending = false;
            key[10] = h->alloc_generation;
            h->g_seen_key = key;
            PSS_TRY(enqueue_query_chain(h, a, bound, &scanned, &pending));
            CU_TRY(h, cudaStreamSynchronize(h->stream));
            return bitmap_compare_finish(h, 1, a, res, scanned, pending);
        }
        bool ok = h->g_cmp_node != nullptr && h->g_cmp_func == h->g_cmp_params.func;
        if (ok) {
            cudaKernelNodeParams np = h->g_cmp_params;
            void* args[1] = {&h->g_cmp_args};
            np.kernelParams = args;
            np.extra = nullptr;
            ok = cudaGraphExecKernelNodeSetParams(h->g_exec, h->g_cmp_node, &np) == cudaSuccess &&
                 cudaEventRecord(h->ev[8], h->stream) == cudaSuccess && cudaGraphLaunch(h->g_exec, h->stream) == cudaSuccess &&
                 cudaEventRecord(h->ev[9], h->stream) == cudaSuccess;
        }
        if (ok) {
            h->launches += h->g_kernels;
            CU_TRY(h, cudaStreamSynchronize(h->stream));
            return bitmap_compare_finish(h, 1, a, res, scanned, pending);
        }
        cudaGetLastError();
        drop_query_graph(h);
        h->use_graph = 0;          // and fall through: kernel by kernel
    } else if (h->use_graph && key == h->g_seen_key) {
        // ---- second query in a row with this geometry: capture, instantiate, launch
        drop_query_graph(h);
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
        int rc = PSS_OK;
        if (ok) {
            const u64 launches_before = h->launches;
            h->capturing = true;
            rc = enqueue_query_chain(h, a, bound, &scanned, &pending);
            h->capturing = false;
            h->g_kernels = h->launches - launches_before;
            h->launches = launches_before;
            ok = cudaStreamEndCapture(h->stream, &graph) == cudaSuccess && graph != nullptr && rc == PSS_OK &&
                 key[10] == h->alloc_generation;     // a buffer that grew during the capture invalidates it
        }
        if (ok) ok = cudaGraphInstantiate(&h->g_exec, graph, 0) == cudaSuccess;
        if (ok) {   // find the compare node (the only kernel node running g_cmp_func)
            size_t n = 0;
            ok = cudaGraphGetNodes(graph, nullptr, &n) == cudaSuccess && n > 0;
            std::vector<cudaGraphNode_t> nodes(n);
            if (ok) ok = cudaGraphGetNodes(graph, nodes.data(), &n) == cudaSuccess;
            h->g_cmp_node = nullptr;
            for (size_t i = 0; ok && i < n; ++i) {
                cudaGraphNodeType ty;
                if (cudaGraphNodeGetType(nodes[i], &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
                cudaKernelNodeParams kp;
                if (cudaGraphKernelNodeGetParams(nodes[i], &kp) != cudaSuccess) continue;
                if (kp.func == h->g_cmp_func) {
                    h->g_cmp_node = nodes[i];
                    h->g_cmp_params = kp;
                }
            }
            ok = ok && (h->g_cmp_node != nullptr || scanned == 0);
        }
        h->g_graph = graph;          // destroyed with the executable graph (drop_query_graph)
        if (ok) ok = cudaEventRecord(h->ev[8], h->stream) == cudaSuccess && cudaGraphLaunch(h->g_exec, h->stream) == cudaSuccess &&
                     cudaEventRecord(h->ev[9], h->stream) == cudaSuccess;
        if (ok) {
            h->g_key = key;
            h->launches += h->g_kernels;
            CU_TRY(h, cudaStreamSynchronize(h->stream));
            return bitmap_compare_finish(h, 1, a, res, scanned, pending);
        }
        cudaGetLastError();
        drop_query_graph(h);
        h->use_graph = 0;
        h->sv_pending = h->bs_pending = false;
        if (rc != PSS_OK) return rc;
        // fall through: kernel by kernel
    }
    h->g_seen_key = key;
    PSS_TRY(enqueue_query_chain(h, a, bound, &scanned, &pending));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return bitmap_compare_finish(h, 1, a, res, scanned, pending);
}

int check_handle(pss_handle* h) {
    if (!h) return fail(nullptr, PSS_EINVAL, "null handle");
    h->misc_uploaded = false;   // every API call starts without a staged MiscDev block
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return fail(h, PSS_ENODEV, "cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e));
    return PSS_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

int pss_abi_version(void) { return PSS_ABI_VERSION; }

uint64_t pss_nth_prime_upper(uint64_t n) {
    // Rosser: p_n < n (ln n + ln ln n) for n >= 6;  Dusart 1999: p_n <= n (ln n + ln ln n - 0.9484), n >= 39017.
    // (the reference over-sieves by 1.3x and regrows by 1.5x instead: utils/sieve.py:574-583)
    if (n < 6) return 15;
    const long double x = (long double)n;
    const long double ln = logl(x), lnln = logl(ln);
    long double b = (n >= 39017) ? x * (ln + lnln - 0.9484L) : x * (ln + lnln);
    return (uint64_t)b + 3;
}

int pss_open(int device, pss_handle** out) {
    if (!out) return fail(nullptr, PSS_EINVAL, "null out pointer");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        cudaGetLastError();
        return fail(nullptr, PSS_ENODEV, "GPU not available: %s", e != cudaSuccess ? cudaGetErrorString(e) : "no CUDA device");
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= n_dev) return fail(nullptr, PSS_ENODEV, "GPU not available: device %d of %d", device, n_dev);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, PSS_ENODEV, "GPU not available: cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, PSS_ENODEV, "GPU not available: %s", cudaGetErrorString(e));
    if (prop.major < 10)
        return fail(nullptr, PSS_ENODEV, "GPU not available: %s is sm_%d%d, this library is built for sm_100a only", prop.name, prop.major,
                    prop.minor);
    pss_handle* h = new pss_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    snprintf(h->name, sizeof h->name, "%s (sm_%d%d, %d SMs, %.1f GB)", prop.name, prop.major, prop.minor, prop.multiProcessorCount,
             (double)prop.totalGlobalMem / 1e9);
    h->tile_bytes = env_u32("PSS_TILE_BYTES", h->tile_bytes);
    h->sieve_threads = env_u32("PSS_SIEVE_THREADS", h->sieve_threads);
    h->n_pat = std::min<u32>(env_u32("PSS_NPAT", h->n_pat), kMaxPatterns);
    if (const char* sp_ = getenv("PSS_SEARCH_PATH")) h->search_path = (strcmp(sp_, "materialised") == 0 || strcmp(sp_, "materialized") == 0) ? 1 : 0;
    h->warp_wide_div = env_u32("PSS_WARP_WIDE_DIV", h->warp_wide_div);
    h->sparse_lead = env_u32("PSS_SPARSE_LEAD", h->sparse_lead);
    h->reduce_moments = (int)env_u32("PSS_REDUCE_MOMENTS", (u32)h->reduce_moments);
    h->compare_exhaustive = (int)env_u32("PSS_COMPARE_EXHAUSTIVE", (u32)h->compare_exhaustive);
    h->kernel_events = (int)std::min<u32>(env_u32("PSS_KERNEL_EVENTS", 1), 2);
    h->use_graph = (int)env_u32("PSS_GRAPH", 1);
    h->tile_bytes = std::max<u32>(8192, std::min<u32>((h->tile_bytes / kChunkBytes) * kChunkBytes, kMaxTileBytes));
    if (env_u32("PSS_SIEVE_PROF", 0) && cudaMalloc(&h->sieve_prof.p, 128) == cudaSuccess) {
        h->sieve_prof.cap = 128;
        cudaMemset(h->sieve_prof.p, 0, 128);
    }
    if (h->sieve_threads != 256 && h->sieve_threads != 512 && h->sieve_threads != 768) h->sieve_threads = 1024;
    auto cleanup = [&](int rc) {
        pss_close(h);
        return rc;
    };
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return cleanup(fail(nullptr, PSS_EINTERNAL, "stream create failed"));
    for (auto& ev : h->ev)
        if (cudaEventCreate(&ev) != cudaSuccess) return cleanup(fail(nullptr, PSS_EINTERNAL, "event create failed"));
    if (cudaHostAlloc(reinterpret_cast<void**>(&h->h_misc), sizeof(MiscDev), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostAlloc(reinterpret_cast<void**>(&h->h_back), sizeof(HostBack), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->h_misc_dev), h->h_misc, 0) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->h_back_dev), h->h_back, 0) != cudaSuccess)
        return cleanup(fail(nullptr, PSS_EINTERNAL, "pinned alloc failed"));
    memset(h->h_misc, 0, sizeof(MiscDev));
    memset(h->h_back, 0, sizeof(HostBack));
    int rc = ensure(h, h->misc, sizeof(MiscDev));
    if (rc == PSS_OK) rc = set_kernel_attributes(h);
    if (rc == PSS_OK) rc = build_patterns(h);
    if (rc != PSS_OK) {
        std::string msg = h->err;
        pss_close(h);
        return fail(nullptr, rc, "%s", msg.c_str());
    }
    *out = h;
    return PSS_OK;
}

int pss_close(pss_handle* h) {
    if (!h) return PSS_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    drop_query_graph(h);
    if (h->sieve_prof.p) {   // instrumented sieve (PSS_SIEVE_PROF=1): where a tile's cycles went
        unsigned long long c[16] = {0};
        cudaMemcpy(c, h->sieve_prof.p, sizeof c, cudaMemcpyDeviceToHost);
        const double n = c[3] ? (double)c[3] : 1.0;
        fprintf(stderr, "[pss sieve prof] tiles=%llu cycles/tile: A=%.0f B=%.0f C=%.0f | warp-cycles/tile in B: dense32=%.0f dense16=%.0f sparse=%.0f large=%.0f, waiting at the end of B=%.0f, entry->first item=%.0f, between items=%.0f, last item->exit=%.0f\n", c[3],
                c[0] / n, c[1] / n, c[2] / n, c[4] / n, c[5] / n, c[6] / n, c[7] / n, c[8] / n, c[9] / n, c[10] / n, c[11] / n);
        cudaFree(h->sieve_prof.p);
    }
    for (DevBuf* b : {&h->d_base, &h->d_base_inv, &h->d_base_strand, &h->d_batch_geo, &h->cols, &h->col_prefix, &h->thread_rec, &h->warp_rec, &h->tile_nact, &h->tri_buf, &h->tri_head, &h->bitmap, &h->chunk_counts, &h->tile_counts, &h->tile_prefix, &h->primes, &h->status, &h->agg,
                      &h->incl, &h->misc, &h->prefix_out, &h->user_vals})
        if (b->p) cudaFree(b->p);
    for (auto& p : h->d_pat)
        if (p) cudaFree(p);
    if (h->h_misc) cudaFreeHost(h->h_misc);
    if (h->h_back) cudaFreeHost(h->h_back);
    if (h->h_misc_multi) cudaFreeHost(h->h_misc_multi);
    if (h->misc_multi.p) cudaFree(h->misc_multi.p);
    if (h->h_peer_results) cudaFreeHost(h->h_peer_results);
    for (int r = 0; r < h->peer_world; ++r)
        if (h->box[r] && h->box[r] != h->box_local) cudaIpcCloseMemHandle(h->box[r]);
    if (h->box_local) cudaFree(h->box_local);
    for (auto& ev : h->ev)
        if (ev) cudaEventDestroy(ev);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return PSS_OK;
}

const char* pss_last_error(pss_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

int pss_device_info(pss_handle* h, char* buf, size_t cap) {
    PSS_TRY(check_handle(h));
    if (!buf || !cap) return fail(h, PSS_EINVAL, "null buffer");
    snprintf(buf, cap, "%s", h->name);
    return PSS_OK;
}

int pss_set_tuning(pss_handle* h, uint32_t tile_bytes, uint32_t sieve_threads, uint32_t n_patterns) {
    PSS_TRY(check_handle(h));
    if (tile_bytes) {
        if (tile_bytes % kChunkBytes || tile_bytes < 8192 || tile_bytes > kMaxTileBytes)
            return fail(h, PSS_EINVAL, "tile_bytes must be a multiple of %d in [8192, %u]", kChunkBytes, kMaxTileBytes);
        h->tile_bytes = tile_bytes;
    }
    if (sieve_threads) {
        if (sieve_threads != 256 && sieve_threads != 512 && sieve_threads != 768 && sieve_threads != 1024)
            return fail(h, PSS_EINVAL, "sieve_threads must be 256, 512, 768 or 1024");
        h->sieve_threads = sieve_threads;
    }
    if (n_patterns) {
        if (n_patterns > kMaxPatterns) return fail(h, PSS_EINVAL, "n_patterns must be <= %d", kMaxPatterns);
        h->n_pat = n_patterns;
    }
    return PSS_OK;
}

int pss_count_primes(pss_handle* h, uint64_t lo, uint64_t hi_excl, uint64_t* count) {
    PSS_TRY(check_handle(h));
    if (!count) return fail(h, PSS_EINVAL, "null count pointer");
    PSS_TRY(sieve_range(h, lo, hi_excl));
    *count = h->sv_total;
    return PSS_OK;
}

int pss_slice_sieve(pss_handle* h, uint64_t lo, uint64_t hi_excl, uint64_t* count, uint64_t* last_prime) {
    PSS_TRY(check_handle(h));
    PSS_TRY(sieve_range(h, lo, hi_excl));
    PSS_TRY(compact_resident(h, ~0ull));
    if (count) *count = h->n_resident;
    if (last_prime) *last_prime = h->n_resident ? h->resident_last : 0;
    return PSS_OK;
}

int pss_generate_primes(pss_handle* h, uint64_t lo, uint64_t hi_excl, uint64_t* out_host, uint64_t cap, uint64_t* count) {
    PSS_TRY(check_handle(h));
    if (!count) return fail(h, PSS_EINVAL, "null count pointer");
    PSS_TRY(sieve_range(h, lo, hi_excl));
    *count = h->sv_total;
    if (h->sv_total > cap) return fail(h, PSS_ENOMEM, "output buffer too small: need %llu entries", (unsigned long long)h->sv_total);
    if (h->sv_total == 0) return PSS_OK;
    if (!out_host) return fail(h, PSS_EINVAL, "null output buffer");
    PSS_TRY(compact_resident(h, ~0ull));
    CU_TRY(h, cudaMemcpyAsync(out_host, h->primes.p, (size_t)h->n_resident * sizeof(u64), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return PSS_OK;
}

static int sieve_first_n(pss_handle* h, u64 n) {
    if (n == 0) return fail(h, PSS_EINVAL, "n must be positive");
    u64 bound = pss_nth_prime_upper(n) + 1;
    for (int attempt = 0; attempt < 4; ++attempt) {
        if (bound > kMaxValue) return fail(h, PSS_EINVAL, "n = %llu needs primes beyond 2^%d", (unsigned long long)n, PSS_MAX_PRIME_BITS);
        PSS_TRY(sieve_range(h, 0, bound));
        if (h->sv_total >= n) return compact_resident(h, n);
        bound += bound / 2;   // cannot happen with a proven bound; kept as a guard
    }
    return fail(h, PSS_EINTERNAL, "prime count bound failed for n = %llu", (unsigned long long)n);
}

int pss_generate_n_primes(pss_handle* h, uint64_t n, uint64_t* out_host) {
    PSS_TRY(check_handle(h));
    if (n == 0) return PSS_OK;
    if (!out_host) return fail(h, PSS_EINVAL, "null output buffer");
    PSS_TRY(sieve_first_n(h, n));
    CU_TRY(h, cudaMemcpyAsync(out_host, h->primes.p, (size_t)n * sizeof(u64), cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    return PSS_OK;
}

int pss_nth_prime(pss_handle* h, uint64_t n_1based, uint64_t* prime) {
    PSS_TRY(check_handle(h));
    if (n_1based == 0) return fail(h, PSS_EINVAL, "n must be positive (1-indexed)");
    if (!prime) return fail(h, PSS_EINVAL, "null output pointer");
    PSS_TRY(sieve_first_n(h, n_1based));
    *prime = h->resident_last;
    return PSS_OK;
}

int pss_power_sum(pss_handle* h, const uint64_t* values_host, uint64_t count, uint32_t power, uint32_t* limbs_out) {
    PSS_TRY(check_handle(h));
    if (!limbs_out) return fail(h, PSS_EINVAL, "null output");
    memset(limbs_out, 0, PSS_LIMBS * sizeof(u32));
    if (count == 0) return PSS_OK;
    if (!values_host) return fail(h, PSS_EINVAL, "null input");
    if (power == 0) {
        limbs_out[0] = (u32)count;
        limbs_out[1] = (u32)(count >> 32);
        return PSS_OK;
    }
    if (count >= (1ull << 36)) return fail(h, PSS_EINVAL, "too many values");
    for (u64 i = 0; i < count; ++i)
        if (values_host[i] >= kMaxValue) return fail(h, PSS_EINVAL, "value %llu at index %llu exceeds 2^%d", (unsigned long long)values_host[i],
                                                      (unsigned long long)i, PSS_MAX_PRIME_BITS);
    PSS_TRY(ensure(h, h->user_vals, (size_t)count * sizeof(u64)));
    CU_TRY(h, cudaMemcpyAsync(h->user_vals.p, values_host, (size_t)count * sizeof(u64), cudaMemcpyHostToDevice, h->stream));
    pss_power_result r[PSS_MAX_POWER];
    PSS_TRY(run_scan(h, static_cast<const u64*>(h->user_vals.p), count, 1, nullptr, power, false, kModeReduce, 0, 0, 0, nullptr, 0, nullptr, r));
    memcpy(limbs_out, r[0].final_sum, PSS_LIMBS * sizeof(u32));
    return PSS_OK;
}

int pss_primesum(pss_handle* h, uint64_t n, uint32_t power, uint32_t* limbs_out) {
    PSS_TRY(check_handle(h));
    if (!limbs_out) return fail(h, PSS_EINVAL, "null output");
    memset(limbs_out, 0, PSS_LIMBS * sizeof(u32));
    if (n == 0) return PSS_OK;
    if (power == 0) {
        limbs_out[0] = (u32)n;
        limbs_out[1] = (u32)(n >> 32);
        return PSS_OK;
    }
    if (power > PSS_MAX_POWER) return fail(h, PSS_EINVAL, "power %u out of range 0..%d", power, PSS_MAX_POWER);
    if (h->search_path == 1) {   // materialised route (kept as a cross-check): u64 primes + streaming reduction
        PSS_TRY(sieve_first_n(h, n));
        pss_power_result r[PSS_MAX_POWER];
        PSS_TRY(run_scan(h, static_cast<const u64*>(h->primes.p), n, 1, nullptr, power, false, kModeReduce, 0, 0, 0, nullptr, 0, nullptr, r));
        memcpy(limbs_out, r[0].final_sum, PSS_LIMBS * sizeof(u32));
        return PSS_OK;
    }
    // fused route: sieve -> block sums -> truncating scan to the n-th prime; no u64 prime buffer (8 GB at n = 1e9) is written
    const u64 bound = pss_nth_prime_upper(n) + 1;
    if (bound > kMaxValue) return fail(h, PSS_EINVAL, "n = %llu needs primes beyond 2^%d", (unsigned long long)n, PSS_MAX_PRIME_BITS);
    pss_search_args a;
    memset(&a, 0, sizeof a);
    a.n_min = n; a.n_max = n; a.power_max = power; a.cmp = 2;   // S < 0: never true; only S_k(n_max) is wanted
    pss_search_result res;
    memset(&res, 0, sizeof res);
    PSS_TRY(prepare_misc(h, &a, nullptr));
    PSS_TRY(sieve_range(h, 0, bound, true));
    PSS_TRY(bitmap_reduce(h, power, false, true));
    PSS_TRY(bitmap_compare(h, 1, nullptr, &a, &res));
    memcpy(limbs_out, res.powers[0].final_sum, PSS_LIMBS * sizeof(u32));
    return PSS_OK;
}

static pss_stats stats_snapshot(pss_handle* h) {
    pss_stats s = h->acc;
    s.launches = h->launches;
    return s;
}
static void fill_stats(pss_handle* h, pss_stats* s, const pss_stats& before) {
    const pss_stats now = stats_snapshot(h);
    memset(s, 0, sizeof *s);
    s->ms_sieve = now.ms_sieve - before.ms_sieve;
    s->ms_count_scan = now.ms_count_scan - before.ms_count_scan;
    s->ms_compact = now.ms_compact - before.ms_compact;
    s->ms_power_scan = now.ms_power_scan - before.ms_power_scan;
    s->ms_bitmap_reduce = now.ms_bitmap_reduce - before.ms_bitmap_reduce;
    s->ms_bitmap_compare = now.ms_bitmap_compare - before.ms_bitmap_compare;
    s->ms_span = now.ms_span - before.ms_span;
    s->ms_peer_wait_totals = now.ms_peer_wait_totals - before.ms_peer_wait_totals;
    s->ms_peer_wait_results = now.ms_peer_wait_results - before.ms_peer_wait_results;
    s->ms_peer_compare_region = now.ms_peer_compare_region - before.ms_peer_compare_region;
    s->ms_peer_pre_publish = now.ms_peer_pre_publish - before.ms_peer_pre_publish;
    s->ms_total_device = s->ms_sieve + s->ms_count_scan + s->ms_compact + s->ms_power_scan;
    s->primes = now.primes - before.primes;
    s->sieve_lo = now.sieve_lo;
    s->sieve_hi = now.sieve_hi;
    s->launches = now.launches - before.launches;
}

static int validate_search(pss_handle* h, const pss_search_args* a, pss_search_result* res) {
    if (!a || !res) return fail(h, PSS_EINVAL, "null argument");
    if (a->power_max < 1 || a->power_max > PSS_MAX_POWER) return fail(h, PSS_EINVAL, "power_max %u out of range 1..%d", a->power_max, PSS_MAX_POWER);
    if (a->cmp > 5) return fail(h, PSS_EINVAL, "invalid comparison operator %u", a->cmp);
    if (a->target_nlimbs && !a->target) return fail(h, PSS_EINVAL, "null target");
    if (a->n_min < 1 || a->n_max < a->n_min) return fail(h, PSS_EINVAL, "need 1 <= n_min <= n_max");
    return PSS_OK;
}

int pss_search(pss_handle* h, const pss_search_args* a, pss_search_result* res) {
    PSS_TRY(check_handle(h));
    PSS_TRY(validate_search(h, a, res));
    const pss_stats before = stats_snapshot(h);
    memset(res, 0, sizeof *res);
    const bool multi = a->all_powers && a->power_max > 1;
    if (h->search_path == 1) {
        // materialised path: K1c compaction into u64 primes, then tile sums + column scan + (pruned) walk over them
        CU_TRY(h, cudaEventRecord(h->ev[8], h->stream));
        PSS_TRY(sieve_first_n(h, a->n_max));
        PSS_TRY(run_scan(h, static_cast<const u64*>(h->primes.p), a->n_max, 1, nullptr, a->power_max, multi, kModeCompare, a->n_min, a->n_max,
                         a->cmp, a->target, a->target_nlimbs, nullptr, res->powers));
        CU_TRY(h, cudaEventRecord(h->ev[9], h->stream));
        CU_TRY(h, cudaEventSynchronize(h->ev[9]));
        h->acc.ms_span += ev_ms(h, 8, 9);     // includes the host synchronisations between the three stages of this path
        res->n_powers = n_powers_of(a->power_max, multi);
        res->primes_scanned = a->n_max;
        res->last_prime = h->resident_last;
    } else {
        // fused path: the bit-packed sieve output is scanned directly
        const u64 bound = pss_nth_prime_upper(a->n_max) + 1;
        if (bound > kMaxValue) return fail(h, PSS_EINVAL, "n = %llu needs primes beyond 2^%d", (unsigned long long)a->n_max, PSS_MAX_PRIME_BITS);
        PSS_TRY(run_query_chain(h, a, res, bound));
    }
    fill_stats(h, &res->stats, before);
    return PSS_OK;
}

int pss_search_targets(pss_handle* h, const pss_search_args* args, uint32_t n_targets, pss_search_result* results) {
    PSS_TRY(check_handle(h));
    if (n_targets == 0) return PSS_OK;
    if (!args || !results) return fail(h, PSS_EINVAL, "null argument");
    if (n_targets > 65536) return fail(h, PSS_EINVAL, "too many targets (%u)", n_targets);
    for (u32 t = 0; t < n_targets; ++t) {
        PSS_TRY(validate_search(h, &args[t], &results[t]));
        if (args[t].n_min != args[0].n_min || args[t].n_max != args[0].n_max || args[t].power_max != args[0].power_max ||
            (args[t].all_powers != 0) != (args[0].all_powers != 0))
            return fail(h, PSS_EINVAL, "all targets of one call must share n_min, n_max, power_max and all_powers");
    }
    if (h->search_path == 1) return fail(h, PSS_EINVAL, "pss_search_targets runs on the fused search path only");
    const pss_stats before = stats_snapshot(h);
    const pss_search_args* a0 = &args[0];
    const bool multi = a0->all_powers && a0->power_max > 1;
    const u32 K = a0->power_max, np = n_powers_of(K, multi);
    const u64 bound = pss_nth_prime_upper(a0->n_max) + 1;
    if (bound > kMaxValue) return fail(h, PSS_EINVAL, "n = %llu needs primes beyond 2^%d", (unsigned long long)a0->n_max, PSS_MAX_PRIME_BITS);
    if (n_targets > h->h_misc_multi_cap) {
        if (h->h_misc_multi) cudaFreeHost(h->h_misc_multi);
        h->h_misc_multi = nullptr;
        h->h_misc_multi_cap = 0;
        CU_TRY(h, cudaMallocHost(reinterpret_cast<void**>(&h->h_misc_multi), sizeof(MiscDev) * n_targets));
        h->h_misc_multi_cap = n_targets;
    }
    PSS_TRY(ensure(h, h->misc_multi, sizeof(MiscDev) * n_targets));
    MiscDev* dmv = static_cast<MiscDev*>(h->misc_multi.p);
    MiscDev* hmv = h->h_misc_multi;
    // one sieve + one block-sum pass, then one (pruned) compare launch per target on the resident records
    memset(hmv, 0, sizeof(MiscDev) * n_targets);
    for (u32 t = 0; t < n_targets; ++t)
        for (u32 j = 0; j < np; ++j) {
            hmv[t].results[j].power = expo_of(K, multi, j);
            hmv[t].results[j].first_match_n = ~0ull;
        }
    CU_TRY(h, cudaMemcpyAsync(dmv, hmv, sizeof(MiscDev) * n_targets, cudaMemcpyHostToDevice, h->stream));
    PSS_TRY(sieve_range(h, 0, bound, true));
    PSS_TRY(bitmap_reduce(h, K, multi, true));
    CU_TRY(h, mark(h, 6));
    if (h->bs_n_tiles > 0)
        for (u32 t = 0; t < n_targets; ++t) {
            BScanParams bp;
            fill_compare_params(h, bp, &args[t], 1, dmv + t);
            PSS_TRY(launch_bscan<kBsCompare>(h, bp, K, multi));
        }
    CU_TRY(h, mark(h, 7));
    PSS_TRY(copy_totals(h));
    CU_TRY(h, cudaMemcpyAsync(hmv, dmv, sizeof(MiscDev) * n_targets, cudaMemcpyDeviceToHost, h->stream));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    sieve_finish(h);
    PSS_TRY(reduce_finish(h));
    if (h->sv_total < a0->n_max)
        return fail(h, PSS_EINTERNAL, "prime count bound failed: %llu primes sieved, n_max = %llu", (unsigned long long)h->sv_total,
                    (unsigned long long)a0->n_max);
    h->acc.ms_power_scan += ev_ms(h, 6, 7);
    h->acc.ms_bitmap_compare += ev_ms(h, 6, 7);
    h->acc.ms_span += ev_ms(h, 0, 7);
    h->acc.primes += a0->n_max;
    for (u32 t = 0; t < n_targets; ++t) {
        pss_search_result* res = &results[t];
        memset(res, 0, sizeof *res);
        if (hmv[t].error_flag) return fail(h, PSS_EINTERNAL, "device error flag %u (target %u)", hmv[t].error_flag, t);
        for (u32 j = 0; j < np; ++j) {
            res->powers[j] = hmv[t].results[j];
            if (res->powers[j].match_count == 0) res->powers[j].first_match_n = 0, res->powers[j].last_match_n = 0;
        }
        res->n_powers = np;
        res->primes_scanned = a0->n_max;
        res->last_prime = hmv[t].last_prime;
    }
    fill_stats(h, &results[0].stats, before);
    for (u32 t = 1; t < n_targets; ++t) results[t].stats = results[0].stats;
    return PSS_OK;
}

int pss_search_tri(pss_handle* h, uint64_t n_min, uint64_t n_max, uint32_t power, uint64_t m_min, uint64_t m_max,
                   uint64_t* pairs_out, uint32_t cap, uint32_t* count, pss_stats* stats) {
    PSS_TRY(check_handle(h));
    if (!count || (cap && !pairs_out)) return fail(h, PSS_EINVAL, "null output");
    if (n_min < 1 || n_max < n_min) return fail(h, PSS_EINVAL, "need 1 <= n_min <= n_max");
    if (power < 1 || power > PSS_MAX_POWER) return fail(h, PSS_EINVAL, "power %u out of range 1..%d", power, PSS_MAX_POWER);
    if (m_max >= (1ull << 29)) return fail(h, PSS_EINVAL, "m_max must be below 2^29 (tri(m) is tested in 64-bit arithmetic)");
    *count = 0;
    const pss_stats before = stats_snapshot(h);
    const u64 bound = pss_nth_prime_upper(n_max) + 1;
    if (bound > kMaxValue) return fail(h, PSS_EINVAL, "n = %llu needs primes beyond 2^%d", (unsigned long long)n_max, PSS_MAX_PRIME_BITS);
    // one launch chain like pss_search: init -> sieve -> block sums -> column scan -> predicate walk -> export, one sync
    pss_search_args ia;
    memset(&ia, 0, sizeof ia);
    ia.power_max = power;
    PSS_TRY(prepare_misc(h, &ia, nullptr));
    h->misc_uploaded = false;
    PSS_TRY(sieve_range(h, 0, bound, true));
    PSS_TRY(bitmap_reduce(h, power, false, true));
    MiscDev* hm = h->h_misc;
    const bool want = m_max >= m_min && m_max >= 1 && h->bs_n_tiles > 0;
    CU_TRY(h, mark(h, 6));
    if (want) {
        const u32 dev_cap = std::max<u32>(cap, 1);
        PSS_TRY(ensure(h, h->tri_buf, (size_t)dev_cap * 2 * sizeof(u64) + 16));
        MiscDev* dm = static_cast<MiscDev*>(h->misc.p);
        BScanParams bp;
        fill_bscan_common(h, bp);
        bp.cols = static_cast<u64*>(h->cols.p);
        bp.col_prefix = static_cast<const u64*>(h->col_prefix.p);
        bp.n_first = 1;
        bp.n_min = n_min;
        bp.n_max = n_max;
        bp.carry_in = dm->carry;
        bp.results = dm->results;
        bp.last_prime_out = &dm->last_prime;
        bp.error_flag = &dm->error_flag;
        bp.exhaustive = h->compare_exhaustive ? 1u : 0u;
        bp.tri_m_min = std::max<u64>(m_min, 1);
        bp.tri_m_max = m_max;
        bp.tri_max_value = m_max * (m_max + 1) / 2;
        bp.tri_out = static_cast<u64*>(h->tri_buf.p);
        bp.tri_count = &dm->ticket;   // reused as the match counter (zeroed by k_init_misc)
        bp.tri_cap = cap;
        static const bool tri_head = env_u32("PSS_TRI_HEAD", 1) != 0;
        if (tri_head && power == 2 && !h->compare_exhaustive) {
            // word-parallel head pass over the first 32 warp tiles (every sum that can equal a tri(m) < 2^59 lies there); the
            // scan kernel below then only classifies, walks what the head did not cover and finishes S(n_max)
            TriHeadParams tp;
            memset(&tp, 0, sizeof tp);
            tp.bitmap = bp.bitmap; tp.n_words = bp.n_words; tp.B_start = bp.B_start;
            const u64 warp_tile_words = 32ull * kWordsPerThread;
            tp.n_head_words = (u32)std::min<u64>(kTriHeadWords, ((bp.n_words + warp_tile_words - 1) / warp_tile_words) * warp_tile_words);
            {   // no sum beyond x can be <= tri(m_max) once x^3 / (12 ln x) exceeds it (at least x / (3 ln x) primes in (x/2, x],
                // each squared >= x^2 / 4); only a size hint -- whatever the head does not cover is tested by the scan kernel
                double x = 1000.0;
                while (x * x * x / (12.0 * std::log(x)) <= (double)bp.tri_max_value && x < 4e6) x *= 1.25;
                const u64 hint_words = (u64)(x / 120.0) + 1;
                const u64 hint = ((hint_words + warp_tile_words - 1) / warp_tile_words) * warp_tile_words;
                tp.n_head_words = (u32)std::min<u64>(tp.n_head_words, hint);
            }
            tp.n_small = bp.n_small;
            for (u32 i = 0; i < bp.n_small; ++i) tp.small_primes[i] = bp.small_primes[i];
            PSS_TRY(ensure(h, h->tri_head, (size_t)kTriHeadWords * (sizeof(u32) + sizeof(u64)) + 16));
            tp.sum = static_cast<u64*>(h->tri_head.p);
            tp.head_n = tp.sum + kTriHeadWords;
            tp.cnt = reinterpret_cast<u32*>(tp.head_n + 1);
            tp.n_min = n_min; tp.n_max = n_max;
            tp.tri_m_min = bp.tri_m_min; tp.tri_m_max = bp.tri_m_max; tp.tri_max_value = bp.tri_max_value;
            tp.tri_out = bp.tri_out; tp.tri_count = bp.tri_count; tp.tri_cap = bp.tri_cap;
            const u32 grid = (tp.n_head_words + 255u) / 256u;
            k_tri_head_sums<<<grid, 256, 0, h->stream>>>(tp);
            k_tri_head_scan<<<1, 1024, 0, h->stream>>>(tp);
            k_tri_head_test<<<grid, 256, 0, h->stream>>>(tp);
            CU_TRY(h, cudaGetLastError());
            h->launches += 3;
            bp.tri_head_n = tp.head_n;
        }
        PSS_TRY(launch_bscan_tri(h, bp, power));
    }
    CU_TRY(h, mark(h, 7));
    PSS_TRY(export_results(h, 0, 0));
    CU_TRY(h, cudaStreamSynchronize(h->stream));
    sieve_finish(h);
    PSS_TRY(reduce_finish(h));
    if (h->sv_total < n_max) return fail(h, PSS_EINTERNAL, "prime count bound failed");
    h->acc.ms_power_scan += ev_ms(h, 6, 7);
    h->acc.ms_bitmap_compare += ev_ms(h, 6, 7);
    h->acc.ms_span += ev_ms(h, 8, 9);
    h->acc.primes += n_max;
    if (want) {
        *count = hm->ticket;
        const u32 n_copy = std::min<u32>(hm->ticket, cap);
        if (n_copy) CU_TRY(h, cudaMemcpy(pairs_out, h->tri_buf.p, (size_t)n_copy * 2 * sizeof(u64), cudaMemcpyDeviceToHost));
    }
    if (stats) fill_stats(h, stats, before);
    return PSS_OK;
}

int pss_set_compare_mode(pss_handle* h, int exhaustive) {
    PSS_TRY(check_handle(h));
    h->compare_exhaustive = exhaustive ? 1 : 0;
    return PSS_OK;
}

int pss_set_kernel_events(pss_handle* h, int level) {
    PSS_TRY(check_handle(h));
    if (level < 0 || level > 2) return fail(h, PSS_EINVAL, "kernel event level must be 0, 1 or 2");
    h->kernel_events = level;
    return PSS_OK;
}

int pss_set_sieve_variant(pss_handle* h, int variant) {
    PSS_TRY(check_handle(h));
    if (variant != 0 && variant != 1) return fail(h, PSS_EINVAL, "sieve variant must be 0 (segmented) or 1 (individual)");
    h->sieve_variant = variant;
    return PSS_OK;
}

int pss_set_search_path(pss_handle* h, int materialised) {
    PSS_TRY(check_handle(h));
    h->search_path = materialised ? 1 : 0;
    return PSS_OK;
}

int pss_slice_sieve_sums(pss_handle* h, uint64_t lo, uint64_t hi_excl, uint32_t power_max, uint32_t all_powers, uint64_t* count,
                         uint32_t* sums_out) {
    PSS_TRY(check_handle(h));
    PSS_TRY(sieve_range(h, lo, hi_excl));
    const bool multi = all_powers && power_max > 1;
    PSS_TRY(bitmap_reduce(h, power_max, multi));
    if (count) *count = h->bs_count;
    if (sums_out) memcpy(sums_out, h->bs_sums, (size_t)n_powers_of(power_max, multi) * PSS_LIMBS * sizeof(u32));
    return PSS_OK;
}

int pss_slice_scan_bitmap(pss_handle* h, uint64_t n_first, const uint32_t* carry_in, const pss_search_args* a, pss_search_result* res) {
    PSS_TRY(check_handle(h));
    PSS_TRY(validate_search(h, a, res));
    if (n_first < 1) return fail(h, PSS_EINVAL, "n_first must be >= 1");
    const pss_stats before = stats_snapshot(h);
    memset(res, 0, sizeof *res);
    PSS_TRY(bitmap_compare(h, n_first, carry_in, a, res));
    fill_stats(h, &res->stats, before);
    return PSS_OK;
}

int pss_peer_export(pss_handle* h, void* ipc_handle_out) {
    PSS_TRY(check_handle(h));
    if (!ipc_handle_out) return fail(h, PSS_EINVAL, "null output");
    static_assert(sizeof(cudaIpcMemHandle_t) == PSS_IPC_HANDLE_BYTES, "IPC handle size");
    static_assert(kMaxPeers == PSS_MAX_PEERS, "peer limit");
    if (!h->box_local) {
        CU_TRY(h, cudaMalloc(reinterpret_cast<void**>(&h->box_local), sizeof(PeerBox)));
        CU_TRY(h, cudaMemset(h->box_local, 0, sizeof(PeerBox)));
        CU_TRY(h, cudaHostAlloc(reinterpret_cast<void**>(&h->h_peer_results), sizeof(PeerResult) * kMaxPeers, cudaHostAllocMapped));
        CU_TRY(h, cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->h_peer_results_dev), h->h_peer_results, 0));
    }
    cudaIpcMemHandle_t hd;
    CU_TRY(h, cudaIpcGetMemHandle(&hd, h->box_local));
    memcpy(ipc_handle_out, &hd, sizeof hd);
    return PSS_OK;
}

int pss_peer_connect(pss_handle* h, int rank, int world, const void* ipc_handles) {
    PSS_TRY(check_handle(h));
    if (!h->box_local) return fail(h, PSS_EINVAL, "pss_peer_export must be called first");
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return fail(h, PSS_EINVAL, "invalid rank %d / world %d (max %d)", rank, world, kMaxPeers);
    if (!ipc_handles) return fail(h, PSS_EINVAL, "null handles");
    for (int r = 0; r < h->peer_world; ++r)
        if (h->box[r] && h->box[r] != h->box_local) cudaIpcCloseMemHandle(h->box[r]);
    h->peer_world = 0;
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            h->box[r] = h->box_local;
            continue;
        }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, static_cast<const char*>(ipc_handles) + (size_t)r * PSS_IPC_HANDLE_BYTES, sizeof hd);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(h, PSS_ENODEV, "cannot map the mailbox of rank %d over peer memory: %s", r, cudaGetErrorString(e));
        }
        h->box[r] = static_cast<PeerBox*>(ptr);
    }
    h->peer_rank = rank;
    h->peer_world = world;
    h->peer_epoch = 0;
    CU_TRY(h, cudaMemset(h->box_local, 0, sizeof(PeerBox)));
    CU_TRY(h, cudaDeviceSynchronize());
    return PSS_OK;
}

int pss_peer_search(pss_handle* h, uint64_t lo, uint64_t hi_excl, const pss_search_args* a, pss_peer_record* records_out, pss_stats* stats) {
    PSS_TRY(check_handle(h));
    if (!records_out) return fail(h, PSS_EINVAL, "null output");
    pss_search_result dummy;
    PSS_TRY(validate_search(h, a, &dummy));
    if (h->peer_world < 1) return fail(h, PSS_EINVAL, "pss_peer_connect must be called first");
    if (h->search_path == 1) return fail(h, PSS_EINVAL, "the peer path runs on the fused search path only");
    const pss_stats before = stats_snapshot(h);
    const bool multi = a->all_powers && a->power_max > 1;
    const u32 K = a->power_max, np = n_powers_of(K, multi);
    const u32 world = (u32)h->peer_world, rank = (u32)h->peer_rank;
    const u64 epoch = ++h->peer_epoch;
    // watchdog of the device-side waits: ranks may enter a query skewed (first-call allocations, host jitter), but a
    // rank that died must not leave the others spinning forever.  PSS_PEER_TIMEOUT_MS, default 10 s.
    const long long watchdog = (long long)env_u32("PSS_PEER_TIMEOUT_MS", 10000) * 2000000ll;   // ~2e6 cycles per ms

    // Everything below can fail on the host (allocations, launches).  The epoch is open: a failure before this rank's two
    // publishes must still release the peers (k_peer_publish_failure), or they would spin until the watchdog fires.
    int stage = 0;           // publishes done: 1 = slice totals, 2 = outcome record
    auto run = [&]() -> int {
        // the MiscDev block goes up first: its copy runs under the sieve instead of between two kernels of the chain
        MiscDev* hm = h->h_misc;
        MiscDev* dm = static_cast<MiscDev*>(h->misc.p);
        PSS_TRY(prepare_misc(h, a, nullptr));
        h->misc_uploaded = false;
        // ---- phase A: sieve + block sums + column scan of this rank's slice (enqueued, not synchronised)
        PSS_TRY(sieve_range(h, lo, hi_excl, true));
        const bool sv_was_pending = h->sv_pending;
        PSS_TRY(bitmap_reduce(h, K, multi, true));
        const bool has_tiles = h->bs_n_tiles > 0;
        if (!sv_was_pending) CU_TRY(h, mark(h, 0));   // empty slice: the chain starts here

        // ---- the exchange: publish the slice totals into every rank's mailbox, gather the lower ranks' totals
        PeerPublishParams pp;
        memset(&pp, 0, sizeof pp);
        pp.col_prefix = has_tiles ? static_cast<const u64*>(h->col_prefix.p) : nullptr;
        pp.col_stride = h->bs_col_stride;
        pp.n_tiles = h->bs_n_tiles;
        pp.np = np;
        for (u32 j = 0; j < np; ++j) pp.width[j] = (u32)sum_limbs((int)expo_of(K, multi, j));
        for (u32 r = 0; r < world; ++r) pp.box[r] = h->box[r];
        pp.world = world;
        pp.rank = rank;
        pp.epoch = epoch;
        pp.tl = dm->tl;
        PeerGatherParams pg;
        memset(&pg, 0, sizeof pg);
        pg.box = h->box_local;
        pg.world = world; pg.rank = rank; pg.np = np;
        pg.epoch = epoch;
        pg.watchdog = watchdog;
        pg.n_first_out = &dm->n_first;
        pg.carry_out = dm->carry;
        pg.error_flag = &dm->error_flag;
        pg.tl = dm->tl;
        static const bool peer_split = env_u32("PSS_PEER_MERGE", 0) == 0;   // default: the two halves as separate launches (8-GPU A/B)
        if (peer_split) {
            k_peer_publish<<<1, 32, 0, h->stream>>>(pp);
            k_peer_gather<<<1, 32, 0, h->stream>>>(pg);
            h->launches += 2;
        } else {
            k_peer_exchange<<<1, 32, 0, h->stream>>>(pp, pg);
            h->launches += 1;
        }
        CU_TRY(h, cudaGetLastError());
        stage = 1;

        // ---- phase B: compare, seeded from device memory (n_first, carry)
        CU_TRY(h, mark(h, 6));
        if (has_tiles) {
            BScanParams bp;
            fill_compare_params(h, bp, a, 0);
            bp.n_first_dev = &dm->n_first;
            PSS_TRY(launch_bscan<kBsCompare>(h, bp, K, multi));
        }
        CU_TRY(h, mark(h, 7));

        // ---- outcome records: publish to every rank, wait for every rank's
        PeerResultParams pr;
        memset(&pr, 0, sizeof pr);
        for (u32 r = 0; r < world; ++r) pr.box[r] = h->box[r];
        pr.world = world; pr.rank = rank; pr.np = np;
        pr.epoch = epoch;
        pr.n_max = a->n_max;
        pr.n_first = &dm->n_first;
        pr.carry = dm->carry;
        pr.results = dm->results;
        pr.last_prime = &dm->last_prime;
        pr.error_flag = &dm->error_flag;
        pr.col_prefix = pp.col_prefix;
        pr.col_stride = pp.col_stride;
        pr.n_tiles = pp.n_tiles;
        for (u32 j = 0; j < np; ++j) pr.width[j] = pp.width[j];
        pr.tl = dm->tl;
        if (peer_split) {
            k_peer_publish_result<<<1, 256, 0, h->stream>>>(pr);
            k_peer_gather_results<<<1, 32, 0, h->stream>>>(h->box_local, world, epoch, watchdog, &dm->error_flag, dm->tl);
            h->launches += 2;
        } else {
            k_peer_exchange_results<<<1, 256, 0, h->stream>>>(pr, h->box_local, world, epoch, watchdog, &dm->error_flag, dm->tl);
            h->launches += 1;
        }
        CU_TRY(h, cudaGetLastError());
        stage = 2;
        PSS_TRY(export_results(h, world, epoch & 1ull));
        CU_TRY(h, cudaStreamSynchronize(h->stream));

        sieve_finish(h);
        PSS_TRY(reduce_finish(h));
        h->acc.ms_power_scan += ev_ms(h, 6, 7);
        h->acc.ms_bitmap_compare += ev_ms(h, 6, 7);
        h->acc.ms_span += ev_ms(h, 8, 9);     // first launch -> all ranks' outcomes gathered and stored to the host (waits on peers included)
        {   // where the span went, from the exchange kernels' own globaltimer stamps
            const u64* tl = hm->tl;
            auto ms = [](u64 a, u64 b) { return b > a ? (double)(b - a) * 1e-6 : 0.0; };
            h->acc.ms_peer_wait_totals += ms(tl[1], tl[2]);
            h->acc.ms_peer_wait_results += ms(tl[4], tl[5]);
            h->acc.ms_peer_compare_region += ms(tl[2], tl[3]);
            h->acc.ms_peer_pre_publish += std::max(0.0, ev_ms(h, 8, 9) - ms(tl[0], tl[5]));
        }
        if (hm->error_flag) return fail(h, PSS_EINTERNAL, "peer exchange failed on the device (flags 0x%x): a rank did not publish in time", hm->error_flag);
        u64 scanned_here = 0;
        for (u32 r = 0; r < world; ++r) {
            const PeerResult& s = h->h_peer_results[r];
            pss_peer_record& o = records_out[r];
            o.valid = s.valid; o.n_first = s.n_first; o.scanned = s.scanned; o.last_prime = s.last_prime; o.count = s.count; o.error = s.error;
            memcpy(o.powers, s.powers, sizeof o.powers);
            if (s.error) return fail(h, PSS_EINTERNAL, "rank %u reported device error flags 0x%llx", r, (unsigned long long)s.error);
            if (r == rank) scanned_here = s.scanned;
        }
        h->acc.primes += scanned_here;
        if (stats) fill_stats(h, stats, before);
        return PSS_OK;
    };
    const int rc = run();
    if (rc != PSS_OK && stage < 2) {
        const std::string keep = h->err;
        PeerFailParams pf;
        memset(&pf, 0, sizeof pf);
        for (u32 r = 0; r < world; ++r) pf.box[r] = h->box[r];
        pf.world = world; pf.rank = rank; pf.stage = (u32)stage; pf.epoch = epoch;
        pf.code = 0x400ull | ((u64)rc << 16);
        k_peer_publish_failure<<<1, 32, 0, h->stream>>>(pf);
        cudaStreamSynchronize(h->stream);
        cudaGetLastError();
        h->err = keep;
    }
    return rc;
}

int pss_slice_reduce(pss_handle* h, uint64_t first, uint64_t count, uint32_t power_max, uint32_t all_powers, uint32_t* sums_out) {
    PSS_TRY(check_handle(h));
    if (!sums_out) return fail(h, PSS_EINVAL, "null output");
    if (first + count > h->n_resident) return fail(h, PSS_EINVAL, "range [%llu,+%llu) outside the %llu resident primes", (unsigned long long)first,
                                                    (unsigned long long)count, (unsigned long long)h->n_resident);
    const bool multi = all_powers && power_max > 1;
    pss_power_result r[PSS_MAX_POWER];
    PSS_TRY(run_scan(h, static_cast<const u64*>(h->primes.p) + first, count, 1, nullptr, power_max, multi, kModeReduce, 0, 0, 0, nullptr, 0, nullptr, r));
    const u32 np = n_powers_of(power_max, multi);
    for (u32 j = 0; j < np; ++j) memcpy(sums_out + j * PSS_LIMBS, r[j].final_sum, PSS_LIMBS * sizeof(u32));
    return PSS_OK;
}

int pss_slice_scan(pss_handle* h, uint64_t first, uint64_t count, uint64_t n_first, const uint32_t* carry_in, const pss_search_args* a,
                   pss_search_result* res) {
    PSS_TRY(check_handle(h));
    PSS_TRY(validate_search(h, a, res));
    if (first + count > h->n_resident) return fail(h, PSS_EINVAL, "range [%llu,+%llu) outside the %llu resident primes", (unsigned long long)first,
                                                    (unsigned long long)count, (unsigned long long)h->n_resident);
    const pss_stats before = stats_snapshot(h);
    memset(res, 0, sizeof *res);
    const bool multi = a->all_powers && a->power_max > 1;
    PSS_TRY(run_scan(h, static_cast<const u64*>(h->primes.p) + first, count, n_first, carry_in, a->power_max, multi, kModeCompare, a->n_min, a->n_max,
                     a->cmp, a->target, a->target_nlimbs, nullptr, res->powers));
    res->n_powers = n_powers_of(a->power_max, multi);
    res->primes_scanned = count;
    if (count) {
        u64 last = 0;
        CU_TRY(h, cudaMemcpy(&last, static_cast<const u64*>(h->primes.p) + first + count - 1, sizeof(u64), cudaMemcpyDeviceToHost));
        res->last_prime = last;
    }
    fill_stats(h, &res->stats, before);
    return PSS_OK;
}

int pss_slice_prefix_sums(pss_handle* h, uint64_t first, uint64_t count, uint32_t power, const uint32_t* carry_in, uint32_t* out_host) {
    PSS_TRY(check_handle(h));
    if (count == 0) return PSS_OK;
    if (!out_host) return fail(h, PSS_EINVAL, "null output");
    if (first + count > h->n_resident) return fail(h, PSS_EINVAL, "range outside the resident primes");
    if (power < 1 || power > PSS_MAX_POWER) return fail(h, PSS_EINVAL, "power %u out of range 1..%d", power, PSS_MAX_POWER);
    PSS_TRY(ensure(h, h->prefix_out, (size_t)count * PSS_LIMBS * sizeof(u32)));
    pss_power_result r[PSS_MAX_POWER];
    PSS_TRY(run_scan(h, static_cast<const u64*>(h->primes.p) + first, count, 1, carry_in, power, false, kModePrefix, 0, 0, 0, nullptr, 0,
                     static_cast<u32*>(h->prefix_out.p), r));
    CU_TRY(h, cudaMemcpy(out_host, h->prefix_out.p, (size_t)count * PSS_LIMBS * sizeof(u32), cudaMemcpyDeviceToHost));
    return PSS_OK;
}

int pss_slice_primes(pss_handle* h, uint64_t first, uint64_t count, uint64_t* out_host) {
    PSS_TRY(check_handle(h));
    if (count == 0) return PSS_OK;
    if (!out_host) return fail(h, PSS_EINVAL, "null output");
    if (first + count > h->n_resident) return fail(h, PSS_EINVAL, "range outside the resident primes");
    CU_TRY(h, cudaMemcpy(out_host, static_cast<const u64*>(h->primes.p) + first, (size_t)count * sizeof(u64), cudaMemcpyDeviceToHost));
    return PSS_OK;
}

int pss_stats_reset(pss_handle* h) {
    PSS_TRY(check_handle(h));
    memset(&h->acc, 0, sizeof h->acc);
    h->launches = 0;
    return PSS_OK;
}

int pss_stats_get(pss_handle* h, pss_stats* out) {
    PSS_TRY(check_handle(h));
    if (!out) return fail(h, PSS_EINVAL, "null output");
    *out = stats_snapshot(h);
    out->ms_total_device = out->ms_sieve + out->ms_count_scan + out->ms_compact + out->ms_power_scan;
    return PSS_OK;
}

int pss_slice_device_buffer(pss_handle* h, void** dev_ptr, uint64_t* count) {
    PSS_TRY(check_handle(h));
    if (dev_ptr) *dev_ptr = h->primes.p;
    if (count) *count = h->n_resident;
    return PSS_OK;
}

}  // extern "C"
