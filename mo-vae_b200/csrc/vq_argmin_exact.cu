// K4x `vq_argmin_exact`: nearest-codebook search evaluated with the reference's own formula and
// float32 roundings, for (a) the rows the tensor-core kernel could not decide (worklist mode) and
// (b) every row when (K, D) is outside the tensor-core kernel's shape (full mode).
//
// Follows /root/reference/models/vq_vae.py:34-39:
//     dist = (sum(z^2) + sum(E^2)) - 2 * (z @ E^T)          two float32 roundings after the GEMM
//     idx  = argmin(dist)                                     first minimal index
// The three inner sums are accumulated in float64 and rounded to float32 once (the correctly
// rounded value every float32 library result is within its own accumulation error of), then
// combined with float32 add / mul / sub exactly in the reference's order.
//
// One warp per row: lanes own codes lane, lane+32, ...; the row is staged in shared memory.
// Two-stage evaluation: a float32 FMA pass over all K codes selects the candidates whose score is
// within a safe bound of the minimum, only those are re-evaluated in float64.
#include <math.h>

#include "common.cuh"

namespace movae {

constexpr int kExWarps = 8;
constexpr int kExThreads = kExWarps * 32;

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(kExThreads)
vq_argmin_exact_kernel(const float* __restrict__ z, int64_t N, int D, int64_t HW, const float* __restrict__ E, int K,
                       const int* __restrict__ list, const unsigned int* __restrict__ list_count,
                       long long* __restrict__ idx_out) {
    extern __shared__ float zs_all[];                       // kExWarps x D
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* zs = zs_all + (size_t)warp * D;
    const int64_t total = list ? (int64_t)(*list_count) : N;
    const int64_t stride = (int64_t)gridDim.x * kExWarps;
    const bool vec = (D % 4 == 0) && (reinterpret_cast<uintptr_t>(E) % 16 == 0);

    for (int64_t w = (int64_t)blockIdx.x * kExWarps + warp; w < total; w += stride) {
        const int64_t n = list ? (int64_t)list[w] : w;
        const int64_t b = n / HW, hw = n - b * HW;
        const float* zp = z + (b * D) * HW + hw;
        double z2p = 0.0;
        for (int d = lane; d < D; d += 32) {
            const float v = zp[(int64_t)d * HW];
            zs[d] = v;
            z2p += (double)v * (double)v;
        }
        __syncwarp();
        const double z2d = warp_sum(z2p);
        const float z2f = (float)z2d;
        const float znorm = sqrtf(z2f);

        // ---- pass 1: float32 scores for every code, row minimum and max code norm ----------------
        float best32 = __uint_as_float(0x7f800000u);
        float emax2 = 0.f;
        for (int j = lane; j < K; j += 32) {
            const float* e = E + (size_t)j * D;
            float dot = 0.f, e2 = 0.f;
            if (vec) {
                for (int d = 0; d < D; d += 4) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(e + d));
                    const float4 y = *reinterpret_cast<const float4*>(zs + d);
                    dot = fmaf(x.x, y.x, dot); dot = fmaf(x.y, y.y, dot); dot = fmaf(x.z, y.z, dot); dot = fmaf(x.w, y.w, dot);
                    e2 = fmaf(x.x, x.x, e2); e2 = fmaf(x.y, x.y, e2); e2 = fmaf(x.z, x.z, e2); e2 = fmaf(x.w, x.w, e2);
                }
            } else {
                for (int d = 0; d < D; ++d) {
                    const float x = __ldg(e + d);
                    dot = fmaf(x, zs[d], dot);
                    e2 = fmaf(x, x, e2);
                }
            }
            best32 = fminf(best32, e2 - 2.f * dot);
            emax2 = fmaxf(emax2, e2);
        }
        best32 = warp_min_f(best32);
        emax2 = -warp_min_f(-emax2);
        // float32 pass error per score <= (D+2) 2^-24 (2|z||e| + |e|^2); candidates within twice that (+ the
        // reference formula's own quantum, 4 ulp of |z|^2 + |e|^2) of the minimum can be the float32-rounded argmin
        const float emax = sqrtf(emax2);
        const float bound = 2.f * (float)(D + 2) * 5.9604645e-08f * (2.f * znorm * emax + emax2) +
                            8.f * 1.1920929e-07f * (z2f + emax2);

        // ---- pass 2: candidates re-evaluated with float64 sums and the reference's float32 formula
        float bestd = __uint_as_float(0x7f800000u);
        int bestj = 0x7fffffff;
        for (int j = lane; j < K; j += 32) {
            const float* e = E + (size_t)j * D;
            float dot = 0.f, e2 = 0.f;
            if (vec) {
                for (int d = 0; d < D; d += 4) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(e + d));
                    const float4 y = *reinterpret_cast<const float4*>(zs + d);
                    dot = fmaf(x.x, y.x, dot); dot = fmaf(x.y, y.y, dot); dot = fmaf(x.z, y.z, dot); dot = fmaf(x.w, y.w, dot);
                    e2 = fmaf(x.x, x.x, e2); e2 = fmaf(x.y, x.y, e2); e2 = fmaf(x.z, x.z, e2); e2 = fmaf(x.w, x.w, e2);
                }
            } else {
                for (int d = 0; d < D; ++d) {
                    const float x = __ldg(e + d);
                    dot = fmaf(x, zs[d], dot);
                    e2 = fmaf(x, x, e2);
                }
            }
            if (e2 - 2.f * dot <= best32 + bound) {
                double dd = 0.0, ee = 0.0;
                for (int d = 0; d < D; ++d) {
                    const double x = (double)__ldg(e + d);
                    dd = fma(x, (double)zs[d], dd);
                    ee = fma(x, x, ee);
                }
                const float dist = __fsub_rn(__fadd_rn(z2f, (float)ee), __fmul_rn(2.f, (float)dd));
                if (dist < bestd) { bestd = dist; bestj = j; }    // ascending j per lane: first minimum kept
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bestd, o);
            const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
            if (od < bestd || (od == bestd && oj < bestj)) { bestd = od; bestj = oj; }
        }
        if (lane == 0) idx_out[n] = (long long)bestj;
        __syncwarp();
    }
}

// list == nullptr: all N rows.  Otherwise the rows in list[0 .. *list_count) (count read on the device).
int launch_vq_argmin_exact(const float* z, int64_t N, int D, int64_t HW, const float* E, int K, const int* list,
                           const unsigned int* list_count, long long* idx, cudaStream_t st) {
    const int sms = sm_count();
    MOVAE_REQUIRE(sms > 0, MOVAE_ERR_CUDA, "CUDA device query failed (no GPU?)");
    const size_t smem = (size_t)kExWarps * D * sizeof(float);
    MOVAE_REQUIRE(smem <= 48 * 1024, MOVAE_ERR_UNSUPPORTED, "vq_argmin: embedding_dim %d too large", D);
    int64_t grid = (N + kExWarps - 1) / kExWarps;
    const int64_t cap = (int64_t)sms * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    vq_argmin_exact_kernel<<<(unsigned)grid, kExThreads, smem, st>>>(z, N, D, HW, E, K, list, list_count, idx);
    MOVAE_CUDA_TRY(cudaGetLastError());
    return MOVAE_OK;
}

}  // namespace movae
