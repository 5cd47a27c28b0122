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
// One warp per row: lanes own codes lane, lane+32, ...; the row is staged in shared memory.  When it
// fits, the codebook is staged TRANSPOSED in shared memory (Es[d][j], row stride K+1 floats) so that a
// warp's 32 codes are 32 consecutive banks (a per-lane walk over global rows costs 32 L1 wavefronts
// per load and ran 14x slower).  Two-stage evaluation: a float32 FMA pass over all K codes selects
// the candidates whose score is within a safe bound of the minimum, only those are re-evaluated in
// float64.
#include <math.h>

#include "common.cuh"

namespace movae {

constexpr int kExWarps = 16;
constexpr int kExThreads = kExWarps * 32;
constexpr size_t kExMaxSmem = 200 * 1024;

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// shared memory: zs[kExWarps][D] | sc[kExWarps][K] | e2s[K] | (STAGE) Es[D][K+1]
template <bool STAGE>
__global__ void __launch_bounds__(kExThreads)
vq_argmin_exact_kernel(const float* __restrict__ z, int64_t N, int D, int64_t HW, const float* __restrict__ E, int K,
                       const int* __restrict__ list, const unsigned int* __restrict__ list_count,
                       long long* __restrict__ idx_out) {
    extern __shared__ float smem_f[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total = list ? (int64_t)(*list_count) : N;
    if ((int64_t)blockIdx.x * kExWarps >= total) return;      // nothing for this CTA: skip the staging
    float* zs = smem_f + (size_t)warp * D;
    float* sc = smem_f + (size_t)kExWarps * D + (size_t)warp * K;
    float* e2s = smem_f + (size_t)kExWarps * (D + K);
    float* Es = e2s + K;
    const int ldE = K + 1;
    const int64_t stride = (int64_t)gridDim.x * kExWarps;

    // ---- per-CTA: |e_j|^2 (float32 chain, used by the candidate filter only) and the staged codebook
    if (STAGE) {
        for (int i = threadIdx.x; i < K * D; i += kExThreads) {
            const int j = i / D, d = i - j * D;
            Es[d * ldE + j] = __ldg(E + i);
        }
        __syncthreads();
    }
    float emax2 = 0.f;
    for (int j = threadIdx.x; j < K; j += kExThreads) {
        float e2 = 0.f;
        for (int d = 0; d < D; ++d) {
            const float x = STAGE ? Es[d * ldE + j] : __ldg(E + (size_t)j * D + d);
            e2 = fmaf(x, x, e2);
        }
        e2s[j] = e2;
    }
    __syncthreads();
    for (int j = lane; j < K; j += 32) emax2 = fmaxf(emax2, e2s[j]);
    emax2 = -warp_min_f(-emax2);
    const float emax = sqrtf(emax2);

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
        const float z2f = (float)warp_sum(z2p);
        const float znorm = sqrtf(z2f);

        // ---- pass 1: float32 scores |e|^2 - 2 z.e for every code; four codes per lane at a time so
        // that one broadcast load of z[d] feeds four independent FMA chains
        float best32 = __uint_as_float(0x7f800000u);
        for (int j0 = lane; j0 < K; j0 += 128) {
            const int j1 = min(j0 + 32, K - 1), j2 = min(j0 + 64, K - 1), j3 = min(j0 + 96, K - 1);
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
            if (STAGE) {
                const float* e = Es;
#pragma unroll 4
                for (int d = 0; d < D; ++d, e += ldE) {
                    const float zv = zs[d];
                    d0 = fmaf(e[j0], zv, d0);
                    d1 = fmaf(e[j1], zv, d1);
                    d2 = fmaf(e[j2], zv, d2);
                    d3 = fmaf(e[j3], zv, d3);
                }
            } else {
                const float *e0 = E + (size_t)j0 * D, *e1 = E + (size_t)j1 * D, *e2p = E + (size_t)j2 * D, *e3 = E + (size_t)j3 * D;
                for (int d = 0; d < D; ++d) {
                    const float zv = zs[d];
                    d0 = fmaf(__ldg(e0 + d), zv, d0);
                    d1 = fmaf(__ldg(e1 + d), zv, d1);
                    d2 = fmaf(__ldg(e2p + d), zv, d2);
                    d3 = fmaf(__ldg(e3 + d), zv, d3);
                }
            }
            const float s0 = e2s[j0] - 2.f * d0, s1 = e2s[j1] - 2.f * d1, s2 = e2s[j2] - 2.f * d2, s3 = e2s[j3] - 2.f * d3;
            sc[j0] = s0;
            best32 = fminf(best32, s0);
            if (j0 + 32 < K) { sc[j1] = s1; best32 = fminf(best32, s1); }
            if (j0 + 64 < K) { sc[j2] = s2; best32 = fminf(best32, s2); }
            if (j0 + 96 < K) { sc[j3] = s3; best32 = fminf(best32, s3); }
        }
        best32 = warp_min_f(best32);
        // float32 pass error per score <= (D+2) 2^-24 (2|z||e| + |e|^2); a code within twice that (+ the
        // reference formula's own quantum, 8 ulp of |z|^2 + |e|^2) of the minimum can be the float32-rounded argmin
        const float bound = 2.f * (float)(D + 2) * 5.9604645e-08f * (2.f * znorm * emax + emax2) +
                            8.f * 1.1920929e-07f * (z2f + emax2);

        // ---- pass 2: candidates re-evaluated with float64 sums and the reference's float32 formula
        float bestd = __uint_as_float(0x7f800000u);
        int bestj = 0x7fffffff;
        for (int j = lane; j < K; j += 32) {
            if (sc[j] <= best32 + bound) {
                double dd = 0.0, ee = 0.0;
                if (STAGE) {
                    for (int d = 0; d < D; ++d) {
                        const double x = (double)Es[d * ldE + j];
                        dd = fma(x, (double)zs[d], dd);
                        ee = fma(x, x, ee);
                    }
                } else {
                    const float* e = E + (size_t)j * D;
                    for (int d = 0; d < D; ++d) {
                        const double x = (double)__ldg(e + d);
                        dd = fma(x, (double)zs[d], dd);
                        ee = fma(x, x, ee);
                    }
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
    const size_t base = ((size_t)kExWarps * (D + K) + K) * sizeof(float);
    const size_t staged = base + (size_t)D * (K + 1) * sizeof(float);
    MOVAE_REQUIRE(base <= kExMaxSmem, MOVAE_ERR_UNSUPPORTED, "vq_argmin: K=%d, D=%d too large for the exact kernel", K, D);
    const bool stage = staged <= kExMaxSmem;
    int64_t grid = (N + kExWarps - 1) / kExWarps;
    const int64_t cap = stage ? (int64_t)sms : (int64_t)sms * 4;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    static thread_local int configured_dev = -1;
    int dev = 0;
    MOVAE_CUDA_TRY(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        MOVAE_CUDA_TRY(cudaFuncSetAttribute(vq_argmin_exact_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kExMaxSmem));
        MOVAE_CUDA_TRY(cudaFuncSetAttribute(vq_argmin_exact_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kExMaxSmem));
        configured_dev = dev;
    }
    if (stage)
        vq_argmin_exact_kernel<true><<<(unsigned)grid, kExThreads, staged, st>>>(z, N, D, HW, E, K, list, list_count, idx);
    else
        vq_argmin_exact_kernel<false><<<(unsigned)grid, kExThreads, base, st>>>(z, N, D, HW, E, K, list, list_count, idx);
    MOVAE_CUDA_TRY(cudaGetLastError());
    return MOVAE_OK;
}

}  // namespace movae
