// host_harness.cpp — compiles the kernels' shared __host__ __device__ arithmetic (pss_common.cuh) for the
// HOST and replays the sieve-tile index arithmetic serially, so the wheel geometry, first-offset,
// edge-mask and limb-multiply code can be checked against the oracle on a machine without a GPU.
// Test infrastructure only (built by tests/test_host_arith.py with g++); it is NOT a CPU fallback and is
// never loaded by the product package.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../prime-square-sum_b200/csrc/pss_common.cuh"

using namespace pss;

static const u32 kGroups[4][5] = {{7, 11, 13, 17, 19}, {23, 29, 31, 0, 0}, {37, 41, 43, 0, 0}, {47, 53, 59, 0, 0}};
static const u32 kFirst[5] = {7, 23, 37, 47, 61};

static uint8_t pattern_byte(u64 i, const u32* qs) {
    u32 m = 0xFF;
    for (u32 j = 0; j < 8; ++j) {
        u64 v = i * 30ull + wheel_residue(j);
        for (int k = 0; k < 5; ++k)
            if (qs[k] && v % qs[k] == 0) m &= ~(1u << j);
    }
    return (uint8_t)m;
}

extern "C" {

// serial replay of k_sieve + k_compact for [lo, hi): returns the number of primes, writes up to cap
u64 emul_sieve(u64 lo, u64 hi, u32 tile_bytes, u32 n_pat, u64* out, u64 cap) {
    u64 count = 0;
    auto emit = [&](u64 v) {
        if (out && count < cap) out[count] = v;
        ++count;
    };
    if (lo >= hi || hi <= 2) return 0;
    for (u64 q : {2ull, 3ull, 5ull})
        if (q >= lo && q < hi) emit(q);
    // base primes
    u64 lim = (u64)std::sqrt((long double)(hi - 1));
    while (lim * lim > hi - 1) --lim;
    while ((lim + 1) * (lim + 1) <= hi - 1) ++lim;
    std::vector<u32> base;
    {
        std::vector<uint8_t> comp(lim + 2, 0);
        for (u64 i = 2; i <= lim; ++i) {
            if (comp[i]) continue;
            if (i >= kFirst[n_pat]) base.push_back((u32)i);
            for (u64 m = i * i; m <= lim; m += i) comp[m] = 1;
        }
    }
    const u64 B_start = (lo / 30) & ~15ull;
    const u64 B_end = (hi + 29) / 30;
    const u64 n_tiles = (B_end - B_start + tile_bytes - 1) / tile_bytes;
    std::vector<u32> tile(tile_bytes / 4);
    for (u64 t = 0; t < n_tiles; ++t) {
        const u64 B0 = B_start + t * tile_bytes;
        uint8_t* tb = reinterpret_cast<uint8_t*>(tile.data());
        for (u32 i = 0; i < tile_bytes; ++i) {
            uint8_t m = 0xFF;
            for (u32 g = 0; g < n_pat; ++g) m &= pattern_byte(B0 + i, kGroups[g]);
            tb[i] = m;
        }
        const double tile_end = (double)(B0 + tile_bytes) * 30.0;
        const u32 plim = (u32)std::fmin(std::sqrt(tile_end) + 2.0, 4.0e9);
        for (u32 p : base) {
            if (p > plim) break;
            const u32 r = (u32)(B0 % p);
            for (u32 j = 0; j < 8; ++j) {
                u32 c, bit;
                strand_geometry(p, j, c, bit);
                for (u32 x = strand_first_offset(p, j, c, B0, r); x < tile_bytes; x += p)
                    tile[x >> 2] &= ~(1u << (((x & 3u) << 3) + bit));
            }
        }
        for (u32 w = 0; w < tile_bytes / 4; ++w) {
            const u64 Bw = B0 + 4ull * w;
            u32 x = tile[w];
            if (Bw == 0) x = (x & 0xFFFF0000u) | 0xDFFEu;
            u32 m = edge_byte_mask(Bw, lo, hi) | (edge_byte_mask(Bw + 1, lo, hi) << 8) | (edge_byte_mask(Bw + 2, lo, hi) << 16) |
                    (edge_byte_mask(Bw + 3, lo, hi) << 24);
            x &= m;
            while (x) {
                u32 b = __builtin_ctz(x);
                x &= x - 1;
                emit(word_bit_value(Bw, b));
            }
        }
    }
    return count;
}

// out[(k-1)*16 ..] = p^k for k = 1..8 through the kernels' big_mul_p chain (16 limbs each, zero padded)
void emul_power_chain(u64 p, u32* out) {
    memset(out, 0, 8 * 16 * sizeof(u32));
    const u32 plo = (u32)p, phi = (u32)(p >> 32);
    Big<pow_limbs(1)> p1; p1.v[0] = plo; p1.v[1] = phi;
    Big<pow_limbs(2)> p2; big_mul_p<pow_limbs(2), pow_limbs(1)>(p2, p1, plo, phi);
    Big<pow_limbs(3)> p3; big_mul_p<pow_limbs(3), pow_limbs(2)>(p3, p2, plo, phi);
    Big<pow_limbs(4)> p4; big_mul_p<pow_limbs(4), pow_limbs(3)>(p4, p3, plo, phi);
    Big<pow_limbs(5)> p5; big_mul_p<pow_limbs(5), pow_limbs(4)>(p5, p4, plo, phi);
    Big<pow_limbs(6)> p6; big_mul_p<pow_limbs(6), pow_limbs(5)>(p6, p5, plo, phi);
    Big<pow_limbs(7)> p7; big_mul_p<pow_limbs(7), pow_limbs(6)>(p7, p6, plo, phi);
    Big<pow_limbs(8)> p8; big_mul_p<pow_limbs(8), pow_limbs(7)>(p8, p7, plo, phi);
    memcpy(out + 0 * 16, p1.v, sizeof p1.v);
    memcpy(out + 1 * 16, p2.v, sizeof p2.v);
    memcpy(out + 2 * 16, p3.v, sizeof p3.v);
    memcpy(out + 3 * 16, p4.v, sizeof p4.v);
    memcpy(out + 4 * 16, p5.v, sizeof p5.v);
    memcpy(out + 5 * 16, p6.v, sizeof p6.v);
    memcpy(out + 6 * 16, p7.v, sizeof p7.v);
    memcpy(out + 7 * 16, p8.v, sizeof p8.v);
}

int emul_limb_widths(int k, int which) { return which ? sum_limbs(k) : pow_limbs(k); }

}  // extern "C"
