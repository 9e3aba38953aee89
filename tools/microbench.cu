// tools/microbench.cu — measured per-SM instruction and shared-memory rates of the B200 (sm_100a) that the
// kernel-local rooflines in bench.py / DESIGN.md are quoted against (SURVEY §8d asked for the integer peak to be
// measured instead of assumed).  Prints one JSON object; build: `make -C prime-square-sum_b200/csrc tools`.
//
//   imad / imad_wide / iadd3 / lop3 / mixed   thread-instructions per clock per SM, 32 warps per SM, 8 independent chains
//   atoms_and_free    shared ATOMS.AND (result unused) with 32 lanes on 32 different banks: warp-instructions / clk / SM
//   atoms_and_random  the same with 32 pseudo-random words per instruction (the sieve's sparse batches)
//   lds32 / lds128    conflict-free shared loads
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) {                                                                   \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                              \
            exit(1);                                                                               \
        }                                                                                          \
    } while (0)

constexpr int kIters = 4096;

template <int OP>
__global__ void __launch_bounds__(1024, 1) k_alu(unsigned* out, unsigned seed) {
    unsigned a[8];
    unsigned long long w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = seed + threadIdx.x * 8u + i;
        w[i] = a[i];
    }
    const unsigned m = seed | 1u, c = seed * 3u + 7u;
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(c));
            if (OP == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(m));
            if (OP == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));
            if (OP == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m), "r"(c));
            if (OP == 4) {   // one multiply-add and one logic op per pair: both pipes busy
                if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(c));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m), "r"(c));
            }
        }
    }
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i] ^ (unsigned)w[i] ^ (unsigned)(w[i] >> 32);
    if (r == 0x12345u) out[0] = r;
}

// MODE 0: ATOMS.AND, lane l -> bank l (conflict free); 1: ATOMS.AND, 32 pseudo-random words per instruction (fixed per
// lane, the whole pattern shifted every iteration); 2: LDS.32 conflict free; 3: LDS.128; 4: ATOMS.AND, all lanes of a warp
// on ONE bank (32 different words): the serialisation rate.  Addressing is one add per 8 accesses (immediate offsets),
// so the loop issues ~1.4 instructions per access and the shared-memory pipe is what is measured.
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_smem(unsigned* out, unsigned seed, unsigned words) {
    extern __shared__ unsigned sm[];
    for (unsigned i = threadIdx.x; i < words; i += blockDim.x) sm[i] = ~0u;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned acc = 0;
    // per-lane word offset inside a 4096-word window; 8 accesses at +0, +4096, .. (same bank), window start moves by 32 words
    unsigned w0;
    if (MODE == 1) w0 = ((seed * 2654435761u + threadIdx.x * 40503u) * 1664525u + 1013904223u) >> 20;   // 0..4095, random bank
    else if (MODE == 4) w0 = (lane * 32u + warp) & 4095u;                                                // bank = warp & 31 for all lanes
    else if (MODE == 3) w0 = (lane * 4u + warp * 128u) & 4095u;                                          // 16-byte vector per lane
    else w0 = (lane + warp * 32u) & 4095u;                                                                // bank = lane
    const unsigned mask = ~(1u << (seed & 31u));
    unsigned base = w0;
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
        unsigned* p = sm + base;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0 || MODE == 1 || MODE == 4) atomicAnd(p + u * 4096, mask);
            else if (MODE == 2) acc += p[u * 4096];
            else {
                const uint4 v = *reinterpret_cast<const uint4*>(p + u * 4096);
                acc += v.x ^ v.y ^ v.z ^ v.w;
            }
        }
        base = (base + (MODE == 3 ? 128u : 32u)) & 4095u;       // keeps every lane's bank
        base += w0 & ((MODE == 3) ? 3u : 0u);                   // (no-op; keeps w0 live for MODE 3 alignment)
    }
    if (acc == 0x12345u) out[0] = acc;
}

template <class F>
double time_ms(F&& launch, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const int sms = prop.multiProcessorCount;
    unsigned* d;
    CK(cudaMalloc(&d, 64));
    const double clk = clk_khz * 1e3;   // nominal boost clock; the measured rates are also reported per second
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_hz_nominal\": %.0f", prop.name, sms, clk);
    const char* names[5] = {"imad", "imad_wide", "iadd", "lop3", "mixed_imad_lop3"};
    for (int op = 0; op < 5; ++op) {
        auto go = [&]() {
            switch (op) {
                case 0: k_alu<0><<<sms, 1024>>>(d, 12345u); break;
                case 1: k_alu<1><<<sms, 1024>>>(d, 12345u); break;
                case 2: k_alu<2><<<sms, 1024>>>(d, 12345u); break;
                case 3: k_alu<3><<<sms, 1024>>>(d, 12345u); break;
                default: k_alu<4><<<sms, 1024>>>(d, 12345u); break;
            }
        };
        const double ms = time_ms(go);
        const double instr = (double)sms * 1024.0 * kIters * 8.0;
        printf(", \"%s\": {\"thread_instr_per_s\": %.4e, \"per_clk_per_sm_at_nominal\": %.2f}", names[op], instr / (ms * 1e-3),
               instr / (ms * 1e-3) / clk / sms);
    }
    const unsigned words = 9 * 4096;   // 144 KiB: 8 immediate offsets of 4096 words + a 4096-word window
    const size_t smem = words * 4;
    CK(cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_smem<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_smem<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const char* sn[5] = {"atoms_and_conflict_free", "atoms_and_random_words", "lds32_conflict_free", "lds128", "atoms_and_single_bank"};
    for (int mode = 0; mode < 5; ++mode) {
        auto go = [&]() {
            switch (mode) {
                case 0: k_smem<0><<<sms, 1024, smem>>>(d, 777u, words); break;
                case 1: k_smem<1><<<sms, 1024, smem>>>(d, 777u, words); break;
                case 2: k_smem<2><<<sms, 1024, smem>>>(d, 777u, words); break;
                case 3: k_smem<3><<<sms, 1024, smem>>>(d, 777u, words); break;
                default: k_smem<4><<<sms, 1024, smem>>>(d, 777u, words); break;
            }
        };
        const double ms = time_ms(go);
        const double winstr = (double)sms * 32.0 * kIters * 8.0;
        printf(", \"%s\": {\"warp_instr_per_s\": %.4e, \"warp_instr_per_clk_per_sm_at_nominal\": %.3f, \"lane_ops_per_s\": %.4e}", sn[mode],
               winstr / (ms * 1e-3), winstr / (ms * 1e-3) / clk / sms, 32.0 * winstr / (ms * 1e-3));
    }
    printf("}\n");
    return 0;
}
