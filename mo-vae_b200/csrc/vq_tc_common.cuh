// Device helpers of K4 (vq_argmin_tc.cu) and K4x (vq_argmin_exact.cu):
// bf16 hi/lo splitting, the TMEM-load wait, and the top-2 (min / second-min) trackers of the epilogue.
#pragma once
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace movae {

struct Top2 {
    float best, second;
    int chunk;      // 32-column chunk the current best came from
    int trk;        // tracker it came from = bits 1, 2 of its column within the chunk
};

// two float32 values -> packed bf16x2 "hi" (round to nearest) and bf16x2 "lo" = bf16(x - hi);
// one packed F2FP conversion per pair instead of two scalar F2F
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xFFFF0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - ah, b - bh);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ uint32_t bf16_bits_rn(float x) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x)); }

__device__ __forceinline__ void top2_push(Top2& tr, float a, float b) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    tr.second = fminf(fminf(tr.second, hi), fmaxf(tr.best, lo));
    tr.best = fminf(tr.best, lo);
}

__device__ __forceinline__ void tmem_ld_wait_for(uint32_t (&v)[32]) {
    // wait::ld with the destination registers as in/out operands so that no use of v[] can be
    // scheduled above the wait
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}

// one 32-column chunk of one row: the accumulator entries ARE the keys; keep the two smallest in
// four independent trackers (pairs -> 2.5 min/max per entry)
template <bool DBG>
__device__ __forceinline__ void epi_chunk(const uint32_t (&v)[32], int chunk, Top2 (&tr)[4], float* __restrict__ dbg_row) {
    float prev[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) prev[k] = tr[k].best;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float k0 = __uint_as_float(v[4 * q + 0]), k1 = __uint_as_float(v[4 * q + 1]);
        const float k2 = __uint_as_float(v[4 * q + 2]), k3 = __uint_as_float(v[4 * q + 3]);
        if (DBG && dbg_row) {
            dbg_row[chunk * 32 + 4 * q + 0] = k0;
            dbg_row[chunk * 32 + 4 * q + 1] = k1;
            dbg_row[chunk * 32 + 4 * q + 2] = k2;
            dbg_row[chunk * 32 + 4 * q + 3] = k3;
        }
        top2_push(tr[(2 * q) & 3], k0, k1);
        top2_push(tr[(2 * q + 1) & 3], k2, k3);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (tr[k].best != prev[k]) tr[k].chunk = chunk;
}

// (distance, index) -> one 64-bit key whose unsigned order is "smaller distance first, then smaller index"
__device__ __forceinline__ unsigned long long pack_dist_index(float dist, int j) {
    unsigned int b = __float_as_uint(dist);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);           // order-preserving map float -> uint
    return ((unsigned long long)b << 32) | (unsigned int)j;
}

__device__ __forceinline__ void top2_merge(Top2& a, const Top2& b) {
    const float nb = fminf(a.best, b.best);
    a.second = fminf(fminf(a.second, b.second), fmaxf(a.best, b.best));
    if (b.best < a.best) { a.chunk = b.chunk; a.trk = b.trk; }
    a.best = nb;
}

}  // namespace movae
