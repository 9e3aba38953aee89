/*
 * pss_b200.h — C ABI of libpss_b200.so
 *
 * B200-native (sm_100a) implementation of the ONE hot path of djdarcy/Prime-Square-Sum:
 *   enumerate primes -> exact k-th powers -> exact running prefix sums
 *   primesum(n,k) = sum_{i<=n} p_i^k -> compare every partial sum with a fixed big-integer target.
 *
 * Plain C: pointers and sizes only, no torch / CUDA types in any signature.  Every entry point
 * names the reference interface (file:line under the reference tree) whose work it replaces.
 * The reference has no native FFI of its own (it is 100 % Python); the binding a maintainer adds
 * is a ctypes stub — see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns a pss_status; on failure pss_last_error(h) describes it.
 *       PSS_OK            -> success
 *       PSS_EINVAL        -> Python ValueError          (utils/sieve.py:485-489, utils/sequences.py:61-62)
 *       PSS_ENODEV        -> Python RuntimeError("GPU not available ...") (utils/gpu.py:160-161,207-208)
 *       PSS_ENOMEM        -> buffer too small / out of device memory; required size written back
 *       PSS_EINTERNAL     -> Python RuntimeError (CUDA error, look-back watchdog, ...)
 *   - big integers cross the boundary as little-endian arrays of 32-bit limbs
 *     (Python: int.from_bytes(bytes(limbs), "little")).  PSS_LIMBS = 12 limbs (384 bit) per value.
 *   - prime arrays are uint64 (bit-compatible with the reference's np.int64 arrays since p < 2^63).
 *   - all value ranges are [lo, hi) — hi EXCLUSIVE like generate_primes(limit) (utils/sieve.py:444);
 *     all prime INDICES n are 1-based and inclusive like the reference's iterators
 *     (utils/iterators.py:99,187-204).
 *   - there is NO CPU fallback: without a usable CUDA device pss_open fails with PSS_ENODEV.
 *   - a handle is single-threaded; CUDA is initialised lazily inside pss_open (fork safety,
 *     utils/gpu.py:263-279 forks a Pool).
 */
#ifndef PSS_B200_H
#define PSS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSS_ABI_VERSION 3
#define PSS_LIMBS 12        /* 32-bit limbs per big integer at the boundary (384 bit) */
#define PSS_MAX_POWER 8     /* highest exponent k supported by the device kernels */
#define PSS_MAX_PRIME_BITS 40 /* sieve value bound: primes < 2^40 */

typedef enum pss_status {
    PSS_OK = 0,
    PSS_EINVAL = 1,
    PSS_ENODEV = 2,
    PSS_ENOMEM = 3,
    PSS_EINTERNAL = 4
} pss_status;

/* comparison operator of the fused compare (utils/grammar.py:770-787 _compare) */
typedef enum pss_cmp {
    PSS_CMP_EQ = 0, PSS_CMP_NE = 1, PSS_CMP_LT = 2, PSS_CMP_LE = 3, PSS_CMP_GT = 4, PSS_CMP_GE = 5
} pss_cmp;

typedef struct pss_handle pss_handle;

/* ---- per-power outcome of a fused scan+compare ------------------------------------------- */
typedef struct pss_power_result {
    uint32_t power;              /* exponent k this record is for */
    uint32_t reserved;
    uint64_t match_count;        /* #n in [n_min,n_max] with S_k(n) <cmp> target */
    uint64_t first_match_n;      /* smallest such n, 0 if none */
    uint64_t last_match_n;       /* largest such n, 0 if none */
    /* bracket: the n with S_k(n-1) < target <= S_k(n) when that crossing lies inside the scanned primes (S_k before the
     * first scanned prime = carry_in, 0 for a search from n = 1).  0 when there is no crossing inside the scan: either every
     * scanned S_k(n) < target, or carry_in >= target already (then final_sum >= target tells the two apart: the crossing
     * lies at or before the first scanned n; for a search from n = 1 that only happens for target = 0). */
    uint64_t cross_n;
    uint64_t cross_prime;        /* p_{cross_n} */
    uint32_t sum_at_cross[PSS_LIMBS];      /* S_k(cross_n) */
    uint32_t sum_before_cross[PSS_LIMBS];  /* S_k(cross_n - 1) */
    uint32_t final_sum[PSS_LIMBS];         /* S_k(last scanned n) */
} pss_power_result;

typedef struct pss_stats {
    double ms_sieve;        /* K1a segmented wheel-30 sieve (CUDA events, this call) */
    double ms_count_scan;   /* K1b exclusive scan of chunk counts */
    double ms_compact;      /* K1c ballot/popc compaction into the u64 prime buffer */
    double ms_power_scan;   /* K2+K3: tile sums + column scan + walk/compare (materialised path), or the fused bitmap passes */
    double ms_total_device; /* first kernel start -> last kernel end */
    uint64_t primes;        /* primes power-summed */
    uint64_t sieve_lo, sieve_hi; /* value range sieved */
    uint64_t launches;      /* kernels launched by this call */
    double ms_bitmap_reduce;  /* fused path: per-tile (count, sums) from the bitmap + column prefix scan (part of ms_power_scan) */
    double ms_bitmap_compare; /* fused path: re-walk with exact prefixes + compare (part of ms_power_scan) */
    double ms_span;           /* pss_search, fused path: first launch -> last kernel end of the query on the launch stream,
                                 gaps between kernels included (0 for multi-call sequences) */
    /* pss_peer_search: where ms_span went, from the exchange kernels' own %globaltimer stamps */
    double ms_peer_pre_publish;    /* first launch -> this rank's slice totals published (sieve + block sums + column scan + gaps) */
    double ms_peer_wait_totals;    /* k_peer_gather: waiting for the lower ranks' slice totals */
    double ms_peer_compare_region; /* totals gathered -> outcome record published (compare kernel + launch gaps) */
    double ms_peer_wait_results;   /* k_peer_gather_results: waiting for every rank's outcome record */
} pss_stats;

/* ---- fused search: replaces the O(n^2) host loop utils/grammar.py:937-999 (find_matches) over
 *      primesum(n,k) (utils/sequences.py:30-75) for expressions  primesum(n,k) <cmp> <constant>  ---- */
typedef struct pss_search_args {
    uint64_t n_min;             /* first prime index compared (1-based, inclusive) */
    uint64_t n_max;             /* last prime index compared (inclusive) */
    uint32_t power_max;         /* k (1..PSS_MAX_POWER) */
    uint32_t all_powers;        /* 0: only k = power_max; 1: every k in 1..power_max in one pass (shared power chain) */
    uint32_t cmp;               /* pss_cmp */
    uint32_t target_nlimbs;     /* significant limbs in target (may exceed PSS_LIMBS: then nothing can reach it) */
    const uint32_t* target;     /* little-endian limbs, target_nlimbs of them */
} pss_search_args;

typedef struct pss_search_result {
    uint32_t n_powers;                       /* records filled in powers[] */
    uint32_t reserved;
    uint64_t primes_scanned;                 /* = n_max */
    uint64_t last_prime;                     /* p_{n_max} */
    pss_power_result powers[PSS_MAX_POWER];  /* index j <-> power (all_powers ? j+1 : power_max) */
    pss_stats stats;
} pss_search_result;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* device < 0: use the CUDA current device (or LOCAL_RANK-selected by the caller). */
int pss_abi_version(void);
int pss_open(int device, pss_handle** out);
int pss_close(pss_handle* h);
const char* pss_last_error(pss_handle* h);           /* valid until the next call on h; h may be NULL */
int pss_device_info(pss_handle* h, char* buf, size_t cap);   /* replaces utils/gpu.py:27-73 init_gpu() info */
/* tuning knobs (0 = keep default): sieve tile bytes, sieve block threads, number of pre-sieve patterns */
int pss_set_tuning(pss_handle* h, uint32_t tile_bytes, uint32_t sieve_threads, uint32_t n_patterns);

/* ---- K1: sieve.  Replaces primesieve.* / _basic_sieve / _segmented_sieve
 *      (utils/sieve.py:224-327, call sites :499,504,527,532,562,567,609,613,635,639) -------------- */
int pss_count_primes(pss_handle* h, uint64_t lo, uint64_t hi_excl, uint64_t* count);         /* prime_count / count_primes */
int pss_generate_primes(pss_handle* h, uint64_t lo, uint64_t hi_excl,
                        uint64_t* out_host, uint64_t cap, uint64_t* count);                   /* generate_primes(_range) */
int pss_generate_n_primes(pss_handle* h, uint64_t n, uint64_t* out_host);                    /* generate_n_primes :541-585 */
int pss_nth_prime(pss_handle* h, uint64_t n_1based, uint64_t* prime);                        /* nth_prime :588-616 */

/* ---- K2: exact power sums over a caller-supplied prime array.  Replaces power_sum /
 *      gpu_power_sum / gpu_power_values / cpu_power_sum (utils/gpu.py:140-334) ------------------ */
int pss_power_sum(pss_handle* h, const uint64_t* values_host, uint64_t count, uint32_t power,
                  uint32_t* limbs_out /* PSS_LIMBS */);
/* primesum(n, power) (utils/sequences.py:30-75): sieve + power + reduce, all on device */
int pss_primesum(pss_handle* h, uint64_t n, uint32_t power, uint32_t* limbs_out /* PSS_LIMBS */);

/* ---- K1+K2+K3 fused search (single GPU) ---------------------------------------------------- */
int pss_search(pss_handle* h, const pss_search_args* args, pss_search_result* res);

/* Several targets / operators over the same n range and power configuration in one call: one sieve and one
 * block-sum pass, then one pruned compare launch per target on the resident records (a few tens of microseconds
 * each).  Serves right-hand sides that are themselves sequences -- primesum(n,k) == fibonacci(m), factorial(m),
 * catalan(m) over an m range (utils/sequences.py:83-200, equations.json) -- and `--limit` enumerations.
 * All args[t] must share n_min, n_max, power_max, all_powers. */
int pss_search_targets(pss_handle* h, const pss_search_args* args, uint32_t n_targets, pss_search_result* results /* n_targets */);

/* `for_any primesum(n,power) == tri(m)` (config: n <= 1e8, m <= 1e7): every partial sum S(n), n in
 * [n_min,n_max], is tested for triangularity on the device (8S+1 perfect square, utils/number_theory.py:292-354)
 * with m restricted to [m_min,m_max].  pairs_out receives up to cap (n, m) pairs (unordered); *count is the
 * total number found.  Replaces the m x n double loop of utils/grammar.py:937-999 for this RHS. */
int pss_search_tri(pss_handle* h, uint64_t n_min, uint64_t n_max, uint32_t power, uint64_t m_min, uint64_t m_max,
                   uint64_t* pairs_out, uint32_t cap, uint32_t* count, pss_stats* stats);

/* sieve variant (the reference's --algorithm sieve:<variant>, utils/sieve.py:56-62,391-437): 0 (default) = segmented
 * bit-packed wheel-30 sieve (k_sieve); 1 = "individual", per-candidate trial division by the primes up to sqrt(v)
 * (utils/sieve.py:330-387 _individual_sieve) as a device kernel with the same outputs — a selectable cross-check of the
 * sieve with the reference's O(n sqrt(n) / log n) complexity, not a fast path. */
int pss_set_sieve_variant(pss_handle* h, int variant);

/* per-kernel CUDA timing events of a query: 0 = none, 1 (default) = around the sieve only (the dominant kernel:
 * pss_stats.ms_sieve), 2 = around every kernel (all pss_stats.ms_* fields).  pss_stats.ms_span (first launch -> end of the
 * last kernel) is always measured.  Each event costs ~2 us of stream time and ~2 us of host time per query. */
int pss_set_kernel_events(pss_handle* h, int level);

/* search path selection: 0 (default) = fused, K2/K3 read the sieve's bit-packed output directly;
 * 1 = materialised, K1c compacts u64 primes first (the layout generate_primes / slice_* serve). */
int pss_set_search_path(pss_handle* h, int materialised);

/* compare-pass strategy of the fused path.  0 (default) = interval pruning: S_k(n) is strictly increasing in n, so a
 * block of consecutive primes whose exact sum interval (S_before, S_after] does not contain the target compares
 * like its prefix; only the block that brackets the target (and the one holding n_max) is walked prime by prime.
 * 1 = exhaustive: every prime is decoded, raised to the k-th power and every S_k(n) is formed and compared, as the
 * reference's per-n loop does (utils/grammar.py:977-988).  Results are identical in both modes. */
int pss_set_compare_mode(pss_handle* h, int exhaustive);

/* fused slice API (multi-GPU phase A / B without a u64 prime buffer):
 * sieve [lo,hi) + per-tile sums; returns the slice's prime count and sum p^k per power record */
int pss_slice_sieve_sums(pss_handle* h, uint64_t lo, uint64_t hi_excl, uint32_t power_max, uint32_t all_powers,
                         uint64_t* count, uint32_t* sums_out /* n_powers * PSS_LIMBS */);
/* compare every partial sum of the slice sieved by pss_slice_sieve_sums; its first prime has global index
 * n_first, carry_in = sums of all primes before it; stops at args->n_max */
int pss_slice_scan_bitmap(pss_handle* h, uint64_t n_first, const uint32_t* carry_in, const pss_search_args* args,
                          pss_search_result* res);

/* ---- multi-GPU query with the exchange on the device (NVLink peer memory, one process per GPU).
 *      Every rank owns a value slice [lo, hi).  The only data the ranks exchange is the per-slice (prime count,
 *      sums) record and the per-rank outcome record; both travel as remote stores into mailboxes
 *      that every rank maps through CUDA IPC, inside the rank's launch chain
 *        sieve -> block sums -> column scan -> publish -> gather -> compare -> publish outcome -> gather outcomes
 *      with ONE host synchronisation per query and no collective library call on the data path.
 *      Set-up: every rank calls pss_peer_export, the handles are all-gathered by the caller (any transport:
 *      torch.distributed in pss_b200.search), every rank calls pss_peer_connect.  pss_peer_search must then be
 *      called by all ranks for every query, in the same order (the mailbox epoch counts queries).
 *      pss_peer_connect zeroes this rank's mailbox: it MUST be followed by a barrier over all ranks (any transport)
 *      before the first pss_peer_search, and a re-connect must not overlap a query of any rank.  A rank whose
 *      pss_peer_search fails on the host after the query was opened publishes an error record (pss_peer_record.error
 *      bit 0x400, status code in bits 16..) so that the other ranks return at once instead of waiting for the
 *      watchdog (PSS_PEER_TIMEOUT_MS, default 10 s). --------------------------------------------------------- */
#define PSS_IPC_HANDLE_BYTES 64
#define PSS_MAX_PEERS 16
typedef struct pss_peer_record {
    uint64_t valid;          /* the rank compared at least one partial sum */
    uint64_t n_first;        /* global index of the first prime of the rank's slice */
    uint64_t scanned;        /* primes of the slice with index <= n_max */
    uint64_t last_prime;     /* p_{n_max} if the slice holds n_max, else 0 */
    uint64_t count;          /* primes in the slice */
    uint64_t error;          /* device-side error flags of the rank (0 = ok) */
    pss_power_result powers[PSS_MAX_POWER];
} pss_peer_record;
int pss_peer_export(pss_handle* h, void* ipc_handle_out /* PSS_IPC_HANDLE_BYTES */);
int pss_peer_connect(pss_handle* h, int rank, int world, const void* ipc_handles /* world * PSS_IPC_HANDLE_BYTES, rank order */);
/* records_out[r] = outcome of rank r (identical on every rank after the call); stats: this rank's device times */
int pss_peer_search(pss_handle* h, uint64_t lo, uint64_t hi_excl, const pss_search_args* args,
                    pss_peer_record* records_out /* world */, pss_stats* stats);

/* ---- slice API: the multi-GPU decomposition (one handle per rank).  The per-slice state
 *      (prime_count, last_prime, {k: S_k}) is exactly IncrementalSumCache's
 *      (utils/sum_cache.py:58-79,104-141) -------------------------------------------------------- */
/* sieve [lo,hi) and keep the compacted primes resident in HBM; returns how many */
int pss_slice_sieve(pss_handle* h, uint64_t lo, uint64_t hi_excl, uint64_t* count, uint64_t* last_prime);
/* sums of p^k over resident primes [first, first+count): sums_out[j*PSS_LIMBS ..] for the n_powers records */
int pss_slice_reduce(pss_handle* h, uint64_t first, uint64_t count, uint32_t power_max, uint32_t all_powers,
                     uint32_t* sums_out /* n_powers * PSS_LIMBS */);
/* scan resident primes [first, first+count) whose global 1-based index starts at n_first, seeded with
 * carry_in (n_powers*PSS_LIMBS limbs = sums of all primes before n_first), fused compare */
int pss_slice_scan(pss_handle* h, uint64_t first, uint64_t count, uint64_t n_first,
                   const uint32_t* carry_in, const pss_search_args* args, pss_search_result* res);
/* materialise S_k(n) for resident primes [first, first+count): out_host[i*PSS_LIMBS ..] = carry + prefix.
 * (serves windows of primesum(n,k) values to the generic evaluator / sampled parity checks) */
int pss_slice_prefix_sums(pss_handle* h, uint64_t first, uint64_t count, uint32_t power,
                          const uint32_t* carry_in /* PSS_LIMBS or NULL */, uint32_t* out_host /* count*PSS_LIMBS */);
/* copy resident primes [first, first+count) to the host */
int pss_slice_primes(pss_handle* h, uint64_t first, uint64_t count, uint64_t* out_host);
/* raw device pointer + count of the resident prime buffer (for zero-copy consumers, e.g. torch.from_blob) */
int pss_slice_device_buffer(pss_handle* h, void** dev_ptr, uint64_t* count);

/* per-kernel device time (CUDA events on the handle's stream) accumulated over all calls since the last
 * reset; ms_total_device = sum of the four kernel phases.  (The reference only has wall-clock timing:
 * prime-square-sum.py:342,450-457.) */
int pss_stats_reset(pss_handle* h);
int pss_stats_get(pss_handle* h, pss_stats* out);

/* upper bound for the n-th prime used to size the sieve (n >= 1) */
uint64_t pss_nth_prime_upper(uint64_t n);

#ifdef __cplusplus
}
#endif
#endif /* PSS_B200_H */
