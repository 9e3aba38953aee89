// pss_peer.cuh — the multi-GPU exchange of the search path over NVLink peer memory (no host, no NCCL on the
// data path).
//
// One process per GPU; every rank owns a value slice.  The scan needs, per rank, the exclusive prefix over
// ranks of (prime count, slice sums): < 500 bytes per rank.  Every rank maps every other rank's
// mailbox (cudaIpc* handles exchanged once at set-up) and
//   k_peer_publish          normalises its slice totals and stores them into ITS slot of EVERY rank's mailbox
//                           (remote stores over NVLink, payload, then a system-scope release of the epoch flag)
//   k_peer_gather           spins (acquire loads, bounded by a watchdog) until the slots of all lower ranks carry
//                           the current epoch, adds them up: n_first and the limb carry, written straight into the
//                           device block the compare kernel reads -- the host never sees the exchange
//   k_peer_publish_result / k_peer_gather_results   the same pattern for the per-rank outcome records, so that
//                           every rank ends the query holding all ranks' records (merged on the host after the
//                           one synchronisation of the query).
// A rank's launch chain is  sieve -> block sums -> column scan -> publish -> gather -> compare -> publish result
// -> gather results -> one D2H: publish always precedes the rank's own wait, so the waits cannot cycle.
//
// Replaces the host-side exchange of the reference's roadmap for multi-device runs (the reference itself is
// single-process: utils/grammar.py:937-999 iterates n on the host).
#pragma once
#include "pss_common.cuh"
#include "../../include/pss_b200.h"

namespace pss {

constexpr int kMaxPeers = 16;

struct alignas(512) PeerSlot {          // phase 1: slice totals of one rank
    u64 flag;                           // epoch of the query the payload belongs to
    u64 count;                          // primes in the rank's slice
    u32 sums[PSS_MAX_POWER * PSS_LIMBS];  // normalised limbs, one record per power
};

struct alignas(256) PeerResult {        // phase 2: outcome of one rank
    u64 flag;
    u64 valid;                          // the rank compared at least one partial sum
    u64 n_first;                        // global index of the slice's first prime
    u64 scanned;                        // primes of the slice with index <= n_max
    u64 last_prime;                     // p_{n_max} if the slice holds n_max, else 0
    u64 count;                          // primes in the slice
    u64 error;                          // watchdog / kernel error flags of the rank
    u64 pad;
    pss_power_result powers[PSS_MAX_POWER];
};

struct PeerBox {                        // one per rank, in that rank's HBM, mapped by every rank
    PeerSlot a[2][kMaxPeers];           // [epoch parity][source rank]
    PeerResult r[2][kMaxPeers];
};

__device__ __forceinline__ u64 ld_acquire_sys(const u64* p) {
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(u64* p, u64 v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ u64 global_ns() {
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// timeline of the exchange (this rank's own globaltimer, ns): [0] publish entry, [1] gather entry, [2] gather exit,
// [3] publish_result entry, [4] gather_results entry, [5] gather_results exit
__device__ __forceinline__ void stamp(u64* tl, int i) {
    if (tl && threadIdx.x == 0) tl[i] = global_ns();
}

// spin until *flag == epoch; false after `limit` clock cycles (the watchdog: a rank that died must not hang the box)
__device__ __forceinline__ bool wait_flag(const u64* flag, u64 epoch, long long limit) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) != epoch) {
        if (clock64() - t0 > limit) return false;
        __nanosleep(64);
    }
    return true;
}

struct PeerPublishParams {
    const u64* col_prefix;   // column totals at col_prefix[c * col_stride + n_tiles] (deferred-carry lanes); NULL: empty slice
    u64 col_stride;
    u32 n_tiles;
    u32 np;                  // power records
    u32 width[PSS_MAX_POWER];   // limbs per record
    PeerBox* box[kMaxPeers];
    u32 world, rank;
    u64 epoch;
    u64* tl;                 // timeline stamps (device, 8 entries) or NULL
};

// one CTA of 32 * ceil(world / 32) threads: thread t stores this rank's record into rank t's mailbox
__device__ __forceinline__ void peer_publish_body(const PeerPublishParams& P) {
    __shared__ u64 s_count;
    stamp(P.tl, 0);
    __shared__ u32 s_sums[PSS_MAX_POWER * PSS_LIMBS];
    if (threadIdx.x == 0) {
        for (u32 i = 0; i < PSS_MAX_POWER * PSS_LIMBS; ++i) s_sums[i] = 0;
        u64 count = 0;
        if (P.col_prefix) {
            count = P.col_prefix[P.n_tiles];
            u32 col = 1;
            for (u32 j = 0; j < P.np; ++j) {
                u64 c = 0;
                for (u32 i = 0; i < P.width[j]; ++i, ++col) {
                    const u64 v = P.col_prefix[(u64)col * P.col_stride + P.n_tiles];
                    const u64 lo = (v & 0xFFFFFFFFull) + (c & 0xFFFFFFFFull);
                    s_sums[j * PSS_LIMBS + i] = (u32)lo;
                    c = (v >> 32) + (c >> 32) + (lo >> 32);
                }
            }
        }
        s_count = count;
    }
    __syncthreads();
    const u32 t = threadIdx.x;
    if (t < P.world) {
        PeerSlot* s = &P.box[t]->a[P.epoch & 1ull][P.rank];
        s->count = s_count;
        for (u32 i = 0; i < PSS_MAX_POWER * PSS_LIMBS; ++i) s->sums[i] = s_sums[i];
        __threadfence_system();
        st_release_sys(&s->flag, P.epoch);
    }
}

struct PeerGatherParams {
    PeerBox* box;            // this rank's own mailbox
    u32 world, rank, np;
    u64 epoch;
    long long watchdog;      // clock cycles
    u64* n_first_out;        // device: 1 + primes of all lower ranks
    u32* carry_out;          // device: np * PSS_LIMBS limbs = sums of all lower ranks
    u32* error_flag;
    u64* tl;
};

__device__ __forceinline__ void peer_gather_body(const PeerGatherParams& P) {
    __shared__ u32 s_bad;
    stamp(P.tl, 1);
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    const u32 t = threadIdx.x;
    if (t < P.rank) {
        if (!wait_flag(&P.box->a[P.epoch & 1ull][t].flag, P.epoch, P.watchdog)) atomicOr(&s_bad, 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_bad) atomicOr(P.error_flag, 0x100u);
        u64 n = 1;
        for (u32 j = 0; j < P.np; ++j) {
            u32 acc[PSS_LIMBS];
            for (u32 i = 0; i < PSS_LIMBS; ++i) acc[i] = 0;
            for (u32 r = 0; r < P.rank; ++r) {
                const PeerSlot* s = &P.box->a[P.epoch & 1ull][r];
                u64 c = 0;
                for (u32 i = 0; i < PSS_LIMBS; ++i) {
                    const u64 v = (u64)acc[i] + s->sums[j * PSS_LIMBS + i] + c;
                    acc[i] = (u32)v;
                    c = v >> 32;
                }
            }
            for (u32 i = 0; i < PSS_LIMBS; ++i) P.carry_out[j * PSS_LIMBS + i] = acc[i];
        }
        for (u32 r = 0; r < P.rank; ++r) n += P.box->a[P.epoch & 1ull][r].count;
        *P.n_first_out = n;
    }
    stamp(P.tl, 2);
}

struct PeerResultParams {
    PeerBox* box[kMaxPeers];
    u32 world, rank, np;
    u64 epoch;
    u64 n_max;
    const u64* n_first;              // device (k_peer_gather)
    const u32* carry;                // device, np * PSS_LIMBS
    const pss_power_result* results; // device, compare kernel output
    const u64* last_prime;           // device
    const u32* error_flag;           // device
    const u64* col_prefix;           // slice totals as in PeerPublishParams (NULL: empty slice)
    u64 col_stride;
    u32 n_tiles;
    u32 width[PSS_MAX_POWER];
    u64* tl;
};

// one CTA: builds this rank's outcome record and stores it into every rank's mailbox
__device__ __forceinline__ void peer_publish_result_body(const PeerResultParams& P) {
    __shared__ PeerResult s_rec;
    stamp(P.tl, 3);
    if (threadIdx.x == 0) {
        const u64 n_first = *P.n_first;
        const u64 count = P.col_prefix ? P.col_prefix[P.n_tiles] : 0;
        const u64 scanned = (P.n_max >= n_first) ? min(count, P.n_max - n_first + 1) : 0;
        const bool reached = scanned > 0 && (n_first + scanned - 1 == P.n_max);
        s_rec.flag = 0;
        s_rec.valid = scanned > 0 ? 1 : 0;
        s_rec.n_first = n_first;
        s_rec.scanned = scanned;
        s_rec.last_prime = reached ? *P.last_prime : 0;
        s_rec.count = count;
        s_rec.error = *P.error_flag;
        s_rec.pad = 0;
        u32 col = 1;
        for (u32 j = 0; j < PSS_MAX_POWER; ++j) {
            pss_power_result r = P.results[j];
            if (j < P.np) {
                if (r.match_count == 0) r.first_match_n = 0, r.last_match_n = 0;
                if (!reached) {   // slice exhausted before n_max (or skipped): S(last scanned) = carry (+ slice sums)
                    u64 c = 0, k = 0;
                    for (u32 i = 0; i < PSS_LIMBS; ++i) {
                        u64 add = 0;
                        if (scanned && i < P.width[j]) {
                            const u64 v = P.col_prefix[(u64)(col + i) * P.col_stride + P.n_tiles];
                            const u64 lo = (v & 0xFFFFFFFFull) + (k & 0xFFFFFFFFull);
                            add = lo & 0xFFFFFFFFull;
                            k = (v >> 32) + (k >> 32) + (lo >> 32);
                        }
                        const u64 t = (u64)P.carry[j * PSS_LIMBS + i] + add + c;
                        r.final_sum[i] = (u32)t;
                        c = t >> 32;
                    }
                }
                col += P.width[j];
            }
            s_rec.powers[j] = r;
        }
    }
    __syncthreads();
    constexpr u32 kWords = sizeof(PeerResult) / 8;
    const u64* src = reinterpret_cast<const u64*>(&s_rec);
    for (u32 peer = 0; peer < P.world; ++peer) {
        u64* dst = reinterpret_cast<u64*>(&P.box[peer]->r[P.epoch & 1ull][P.rank]);
        for (u32 i = 1 + threadIdx.x; i < kWords; i += blockDim.x) dst[i] = src[i];   // word 0 is the flag
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < P.world) st_release_sys(&P.box[threadIdx.x]->r[P.epoch & 1ull][P.rank].flag, P.epoch);
}

__device__ __forceinline__ void peer_gather_results_body(PeerBox* box, u32 world, u64 epoch, long long watchdog, u32* error_flag, u64* tl) {
    stamp(tl, 4);
    if (threadIdx.x < world) {
        if (!wait_flag(&box->r[epoch & 1ull][threadIdx.x].flag, epoch, watchdog)) atomicOr(error_flag, 0x200u);
    }
    __syncthreads();
    stamp(tl, 5);
}

// The two halves of an exchange step run in ONE launch: publish first (remote stores + release of the epoch flag), then wait
// for the peers.  Every rank publishes before it waits, so the waits cannot cycle; one launch less per step matters because a
// rank's whole query is a few hundred microseconds at 8 GPUs.
__global__ void k_peer_publish(const PeerPublishParams P) { peer_publish_body(P); }
__global__ void k_peer_gather(const PeerGatherParams P) { peer_gather_body(P); }
__global__ void k_peer_publish_result(const PeerResultParams P) { peer_publish_result_body(P); }
__global__ void k_peer_gather_results(PeerBox* box, u32 world, u64 epoch, long long watchdog, u32* error_flag, u64* tl) {
    peer_gather_results_body(box, world, epoch, watchdog, error_flag, tl);
}
// A rank whose host side failed after the query's epoch was opened (allocation, launch error) still has to release its
// peers: it publishes empty slice totals (stage 0) and an outcome record whose error field carries `code` (stages 0, 1), so
// the other ranks leave their waits at once and report the failing rank instead of running into the watchdog.
struct PeerFailParams {
    PeerBox* box[kMaxPeers];
    u32 world, rank;
    u32 stage;               // publishes already done by this rank: 0 none, 1 slice totals
    u64 epoch;
    u64 code;
};
__global__ void k_peer_publish_failure(const PeerFailParams P) {
    const u32 t = threadIdx.x;
    if (t >= P.world) return;
    if (P.stage < 1) {
        PeerSlot* s = &P.box[t]->a[P.epoch & 1ull][P.rank];
        s->count = 0;
        for (u32 i = 0; i < PSS_MAX_POWER * PSS_LIMBS; ++i) s->sums[i] = 0;
        __threadfence_system();
        st_release_sys(&s->flag, P.epoch);
    }
    PeerResult* r = &P.box[t]->r[P.epoch & 1ull][P.rank];
    u64* w = reinterpret_cast<u64*>(r);
    for (u32 i = 1; i < sizeof(PeerResult) / 8; ++i) w[i] = 0;
    r->error = P.code;
    __threadfence_system();
    st_release_sys(&r->flag, P.epoch);
}

__global__ void k_peer_exchange(const PeerPublishParams PP, const PeerGatherParams PG) {
    peer_publish_body(PP);
    __syncthreads();
    peer_gather_body(PG);
}
__global__ void k_peer_exchange_results(const PeerResultParams PR, PeerBox* box, u32 world, u64 epoch, long long watchdog, u32* error_flag, u64* tl) {
    peer_publish_result_body(PR);
    __syncthreads();
    peer_gather_results_body(box, world, epoch, watchdog, error_flag, tl);
}

}  // namespace pss
