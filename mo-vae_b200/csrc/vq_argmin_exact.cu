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
// One warp per group of R rows (R = 4 when the codebook fits in shared memory): lanes own codes lane,
// lane+32, ...; the rows are staged interleaved in shared memory so that one 16-byte broadcast load
// feeds the same channel of all R rows.  The codebook is staged TRANSPOSED (Es[d][j], row stride K+1
// floats) so that a warp's 32 codes are 32 consecutive banks (a per-lane walk over global rows costs 32
// L1 wavefronts per load and ran 14x slower); each codebook element read from shared memory is used for
// R rows x 1 FMA and four codes per lane run as independent chains (the one-row version was
// shared-memory-bandwidth bound: 1300 wavefronts per row).  Two-stage evaluation: a float32 FMA pass
// over all K codes selects the candidates whose score is within a safe bound of the row minimum, only
// those are re-evaluated with float64 sums.
#include <math.h>

#include "common.cuh"

namespace movae {

constexpr int kExWarps = 8;
constexpr int kExThreads = kExWarps * 32;
constexpr size_t kExMaxSmem = 220 * 1024;

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// shared memory: zs[kExWarps][D][R] | sc[kExWarps][R][K] | e2s[K] | (STAGE) Es[D][K+1]
template <bool STAGE, int R>
__global__ void __launch_bounds__(kExThreads)
vq_argmin_exact_kernel(const float* __restrict__ z, int64_t N, int D, int64_t HW, const float* __restrict__ E, int K,
                       const int* __restrict__ list, const unsigned int* __restrict__ list_count,
                       long long* __restrict__ idx_out) {
    extern __shared__ __align__(16) float smem_f[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t total = list ? (int64_t)(*list_count) : N;
    const int64_t n_groups = (total + R - 1) / R;
    if ((int64_t)blockIdx.x * kExWarps >= n_groups) return;   // nothing for this CTA: skip the staging
    float* zs = smem_f + (size_t)warp * D * R;
    float* sc = smem_f + (size_t)kExWarps * D * R + (size_t)warp * R * K;
    float* e2s = smem_f + (size_t)kExWarps * R * (D + K);
    float* Es = e2s + K;
    const int ldE = K + 1;
    const int64_t stride = (int64_t)gridDim.x * kExWarps;

    // ---- per-CTA: the staged codebook and |e_j|^2 (float32 chain, used by the candidate filter only) ----
    if (STAGE) {
        // transposed copy, 8 independent 16-byte loads in flight per thread (the 4-byte / one-at-a-time version made
        // this fixed cost ~20 us per CTA -- more than the re-check itself at the BASELINE sizes)
        if ((D & 3) == 0 && (reinterpret_cast<uintptr_t>(E) & 15) == 0) {
            const int n4 = K * D / 4;
            for (int i0 = threadIdx.x; i0 < n4; i0 += kExThreads * 8) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * kExThreads;
                    v[u] = i < n4 ? __ldg(reinterpret_cast<const float4*>(E) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int i = i0 + u * kExThreads;
                    if (i < n4) {
                        const int e0 = i * 4, j = e0 / D, d = e0 - j * D;
                        Es[d * ldE + j] = v[u].x;
                        Es[(d + 1) * ldE + j] = v[u].y;
                        Es[(d + 2) * ldE + j] = v[u].z;
                        Es[(d + 3) * ldE + j] = v[u].w;
                    }
                }
            }
        } else {
            for (int i = threadIdx.x; i < K * D; i += kExThreads) {
                const int j = i / D, d = i - j * D;
                Es[d * ldE + j] = __ldg(E + i);
            }
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < K; j += kExThreads) {
        float e2 = 0.f;
        for (int d = 0; d < D; ++d) {
            const float x = STAGE ? Es[d * ldE + j] : __ldg(E + (size_t)j * D + d);
            e2 = fmaf(x, x, e2);
        }
        e2s[j] = e2;
    }
    __syncthreads();
    float emax2 = 0.f;
    for (int j = lane; j < K; j += 32) emax2 = fmaxf(emax2, e2s[j]);
    emax2 = -warp_min_f(-emax2);
    const float emax = sqrtf(emax2);

    for (int64_t g = (int64_t)blockIdx.x * kExWarps + warp; g < n_groups; g += stride) {
        // ---- stage the R rows of this group (the last group repeats its final row) --------------------
        int64_t rows[R];
        float z2f[R];
        constexpr int kMaxPerLane = 8;                       // D <= 256 handled with loads in flight; larger D loops
        {
            const float* zp[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int64_t w = g * R + r;
                w = w < total ? w : total - 1;
                rows[r] = list ? (int64_t)list[w] : w;
                const int64_t b = rows[r] / HW, hw = rows[r] - b * HW;
                zp[r] = z + (b * D) * HW + hw;
            }
            double z2p[R];
#pragma unroll
            for (int r = 0; r < R; ++r) z2p[r] = 0.0;
            for (int d0 = 0; d0 < D; d0 += 32 * kMaxPerLane) {
                float v[kMaxPerLane][R];
#pragma unroll
                for (int i = 0; i < kMaxPerLane; ++i) {
                    const int d = d0 + i * 32 + lane;
#pragma unroll
                    for (int r = 0; r < R; ++r) v[i][r] = d < D ? __ldg(zp[r] + (int64_t)d * HW) : 0.f;   // all rows' loads in flight together
                }
#pragma unroll
                for (int i = 0; i < kMaxPerLane; ++i) {
                    const int d = d0 + i * 32 + lane;
                    if (d < D) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            zs[d * R + r] = v[i][r];
                            z2p[r] += (double)v[i][r] * (double)v[i][r];
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) z2f[r] = (float)warp_sum(z2p[r]);
        }
        __syncwarp();

        // ---- pass 1: float32 scores |e|^2 - 2 z.e for every code and row --------------------------------
        float best32[R];
#pragma unroll
        for (int r = 0; r < R; ++r) best32[r] = __uint_as_float(0x7f800000u);
        for (int j0 = lane; j0 < K; j0 += 128) {
            int jj[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) jj[c] = min(j0 + 32 * c, K - 1);
            float acc[4][R];
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int r = 0; r < R; ++r) acc[c][r] = 0.f;
            if (STAGE) {
                const float* e = Es;
#pragma unroll 2
                for (int d = 0; d < D; ++d, e += ldE) {
                    float zv[R];
                    if constexpr (R == 4) {
                        const float4 q = *reinterpret_cast<const float4*>(zs + d * 4);
                        zv[0] = q.x; zv[1] = q.y; zv[2] = q.z; zv[3] = q.w;
                    } else {
#pragma unroll
                        for (int r = 0; r < R; ++r) zv[r] = zs[d * R + r];
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float ev = e[jj[c]];
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[c][r] = fmaf(ev, zv[r], acc[c][r]);
                    }
                }
            } else {
                for (int d = 0; d < D; ++d) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float ev = __ldg(E + (size_t)jj[c] * D + d);
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[c][r] = fmaf(ev, zs[d * R + r], acc[c][r]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (j0 + 32 * c < K) {
                    const float e2 = e2s[jj[c]];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float s = e2 - 2.f * acc[c][r];
                        sc[r * K + jj[c]] = s;
                        best32[r] = fminf(best32[r], s);
                    }
                }
            }
        }
        __syncwarp();

        // ---- pass 2 (per row): candidates re-evaluated with float64 sums and the reference's float32 formula
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float b32 = warp_min_f(best32[r]);
            const float znorm = sqrtf(z2f[r]);
            // float32 pass error per score <= (D+2) 2^-24 (2|z||e| + |e|^2); a code within twice that (+ the reference
            // formula's own quantum, 8 ulp of |z|^2 + |e|^2) of the minimum can be the float32-rounded argmin
            const float bound = 2.f * (float)(D + 2) * 5.9604645e-08f * (2.f * znorm * emax + emax2) +
                                8.f * 1.1920929e-07f * (z2f[r] + emax2);
            float bestd = __uint_as_float(0x7f800000u);
            int bestj = 0x7fffffff;
            for (int jb = 0; jb < K; jb += 32) {
                const int j = jb + lane;
                unsigned cand = __ballot_sync(0xffffffffu, j < K && sc[r * K + j] <= b32 + bound);
                while (cand) {                                  // the whole warp evaluates one candidate: lanes over d
                    const int jc = jb + __ffs(cand) - 1;
                    cand &= cand - 1;
                    double dd = 0.0, ee = 0.0;
                    for (int d = lane; d < D; d += 32) {
                        const double x = (double)(STAGE ? Es[d * ldE + jc] : __ldg(E + (size_t)jc * D + d));
                        dd = fma(x, (double)zs[d * R + r], dd);
                        ee = fma(x, x, ee);
                    }
                    dd = warp_sum(dd);
                    ee = warp_sum(ee);
                    const float dist = __fsub_rn(__fadd_rn(z2f[r], (float)ee), __fmul_rn(2.f, (float)dd));
                    if (dist < bestd) { bestd = dist; bestj = jc; }   // candidates visited in ascending j: first minimum kept
                }
            }
            // every distance inf / NaN (non-finite latents): torch.argmin of such a row is its first element
            if (lane == 0) idx_out[rows[r]] = (long long)(bestj == 0x7fffffff ? 0 : bestj);
        }
        __syncwarp();
    }
}

// Small-worklist variant (the BASELINE model shapes: N <= 262,144 rows leave 20 .. 800 rows to re-check): the staged
// kernel above spends ~20-35 us per CTA on its fixed prologue (128 KB transposed codebook copy + |e|^2) whatever the
// list length, which is more than the tensor-core search itself at N = 8,192.  Here ONE CTA takes ONE group of R = 4
// rows and nothing is staged: warp w evaluates codes [w K / 8, (w + 1) K / 8), each lane walks its own codebook rows
// with 16-byte loads straight from L2 (the search kernel has just read them), float32 scores and |e|^2 stay in
// registers, the row minima are combined through shared memory, and each lane re-evaluates its own candidates with
// float64 sums and the reference's float32 formula; the winner (smallest distance, then smallest index) is an
// order-free 64-bit integer atomicMin.  Same candidate bound and the same result as the staged kernel.
constexpr int kDxThreads = 256, kDxWarps = kDxThreads / 32, kDxR = 4, kDxMaxPerLane = 4, kDxMaxD = 256;

__device__ __forceinline__ unsigned long long pack_dist_index(float dist, int j) {
    unsigned int b = __float_as_uint(dist);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);           // order-preserving map float -> uint
    return ((unsigned long long)b << 32) | (unsigned int)j;
}

__global__ void __launch_bounds__(kDxThreads)
vq_argmin_exact_direct_kernel(const float* __restrict__ z, int64_t N, int D, int64_t HW, const float* __restrict__ E, int K,
                              const int* __restrict__ list, const unsigned int* __restrict__ list_count,
                              long long* __restrict__ idx_out) {
    __shared__ __align__(16) float zs[kDxMaxD * kDxR];        // [d][r]
    __shared__ double z2d[kDxR];
    __shared__ float wmin[kDxWarps][kDxR], wemax[kDxWarps];
    __shared__ unsigned long long best[kDxR];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t total = list ? (int64_t)(*list_count) : N;
    const int64_t n_groups = (total + kDxR - 1) / kDxR;
    const bool vec = (D % 4 == 0) && (reinterpret_cast<uintptr_t>(E) % 16 == 0);
    const int per_warp = (K + kDxWarps - 1) / kDxWarps;       // codes per warp; lane handles j0 + lane + 32 c

    for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        // ---- stage the group's rows (the last group repeats its final row), |z|^2 in float64 --------------
        int64_t rows[kDxR];
#pragma unroll
        for (int r = 0; r < kDxR; ++r) {
            int64_t wv = g * kDxR + r;
            wv = wv < total ? wv : total - 1;
            rows[r] = list ? (int64_t)list[wv] : wv;
        }
        __syncthreads();                                      // previous group's readers are done
        for (int t = tid; t < D * kDxR; t += kDxThreads) {
            const int d = t / kDxR, r = t - d * kDxR;
            const int64_t b = rows[r] / HW, hw = rows[r] - b * HW;
            zs[t] = __ldg(z + (b * D + d) * HW + hw);
        }
        if (tid < kDxR) best[tid] = ~0ull;
        __syncthreads();
        if (warp < kDxR) {
            double s = 0.0;
            for (int d = lane; d < D; d += 32) s += (double)zs[d * kDxR + warp] * (double)zs[d * kDxR + warp];
            s = warp_sum(s);
            if (lane == 0) z2d[warp] = s;
        }

        // ---- pass 1: float32 scores |e|^2 - 2 z.e of this lane's codes, all R rows ------------------------
        float sc[kDxMaxPerLane][kDxR], e2v[kDxMaxPerLane];
        float mn[kDxR], emax2 = 0.f;
#pragma unroll
        for (int r = 0; r < kDxR; ++r) mn[r] = __uint_as_float(0x7f800000u);
        const int j_lo = warp * per_warp, j_hi = min(K, j_lo + per_warp);
#pragma unroll
        for (int c = 0; c < kDxMaxPerLane; ++c) {
            const int j = j_lo + lane + 32 * c;
            e2v[c] = 0.f;
#pragma unroll
            for (int r = 0; r < kDxR; ++r) sc[c][r] = __uint_as_float(0x7f800000u);
            if (j < j_hi) {
                float acc[kDxR] = {0.f, 0.f, 0.f, 0.f}, e2 = 0.f;
                const float* ep = E + (size_t)j * D;
                if (vec) {
                    for (int d = 0; d < D; d += 4) {
                        const float4 e = __ldg(reinterpret_cast<const float4*>(ep + d));
                        const float ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float4 q = *reinterpret_cast<const float4*>(zs + (d + u) * kDxR);
                            acc[0] = fmaf(ev[u], q.x, acc[0]); acc[1] = fmaf(ev[u], q.y, acc[1]);
                            acc[2] = fmaf(ev[u], q.z, acc[2]); acc[3] = fmaf(ev[u], q.w, acc[3]);
                            e2 = fmaf(ev[u], ev[u], e2);
                        }
                    }
                } else {
                    for (int d = 0; d < D; ++d) {
                        const float ev = __ldg(ep + d);
                        const float4 q = *reinterpret_cast<const float4*>(zs + d * kDxR);
                        acc[0] = fmaf(ev, q.x, acc[0]); acc[1] = fmaf(ev, q.y, acc[1]);
                        acc[2] = fmaf(ev, q.z, acc[2]); acc[3] = fmaf(ev, q.w, acc[3]);
                        e2 = fmaf(ev, ev, e2);
                    }
                }
                e2v[c] = e2;
                emax2 = fmaxf(emax2, e2);
#pragma unroll
                for (int r = 0; r < kDxR; ++r) {
                    sc[c][r] = e2 - 2.f * acc[r];
                    mn[r] = fminf(mn[r], sc[c][r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kDxR; ++r) mn[r] = warp_min_f(mn[r]);
        emax2 = -warp_min_f(-emax2);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < kDxR; ++r) wmin[warp][r] = mn[r];
            wemax[warp] = emax2;
        }
        __syncthreads();
        float em2 = 0.f;
#pragma unroll
        for (int w = 0; w < kDxWarps; ++w) em2 = fmaxf(em2, wemax[w]);
        const float emax = sqrtf(em2);

        // ---- pass 2: this lane's candidates with float64 sums and the reference's float32 formula -----------
#pragma unroll
        for (int r = 0; r < kDxR; ++r) {
            float b32 = wmin[0][r];
#pragma unroll
            for (int w = 1; w < kDxWarps; ++w) b32 = fminf(b32, wmin[w][r]);
            const float z2f = (float)z2d[r];
            const float znorm = sqrtf(z2f);
            const float bound = 2.f * (float)(D + 2) * 5.9604645e-08f * (2.f * znorm * emax + em2) +
                                8.f * 1.1920929e-07f * (z2f + em2);
#pragma unroll
            for (int c = 0; c < kDxMaxPerLane; ++c) {
                const int j = j_lo + lane + 32 * c;
                if (j < j_hi && sc[c][r] <= b32 + bound) {
                    const float* ep = E + (size_t)j * D;
                    double dd = 0.0, ee = 0.0;
                    for (int d = 0; d < D; ++d) {
                        const double x = (double)__ldg(ep + d);
                        dd = fma(x, (double)zs[d * kDxR + r], dd);
                        ee = fma(x, x, ee);
                    }
                    const float dist = __fsub_rn(__fadd_rn(z2f, (float)ee), __fmul_rn(2.f, (float)dd));
                    atomicMin(&best[r], pack_dist_index(dist, j));          // smallest distance, then smallest index
                }
            }
        }
        __syncthreads();
        if (tid < kDxR && g * kDxR + tid < total)   // no candidate at all (every distance inf / NaN): index 0 like torch.argmin
            idx_out[rows[tid]] = best[tid] == ~0ull ? 0ll : (long long)(unsigned int)(best[tid] & 0xffffffffull);
    }
}

// list == nullptr: all N rows.  Otherwise the rows in list[0 .. *list_count) (count read on the device).
int launch_vq_argmin_exact(const float* z, int64_t N, int D, int64_t HW, const float* E, int K, const int* list,
                           const unsigned int* list_count, long long* idx, cudaStream_t st) {
    const int sms = sm_count();
    MOVAE_REQUIRE(sms > 0, MOVAE_ERR_CUDA, "CUDA device query failed (no GPU?)");
    const size_t staged = ((size_t)kExWarps * 4 * (D + K) + K + (size_t)D * (K + 1)) * sizeof(float);
    const size_t plain = ((size_t)kExWarps * 1 * (D + K) + K) * sizeof(float);
    const bool stage = staged <= kExMaxSmem && D % 1 == 0;
    MOVAE_REQUIRE(stage || plain <= kExMaxSmem, MOVAE_ERR_UNSUPPORTED, "vq_argmin: K=%d, D=%d too large for the exact kernel", K, D);
    // few rows to re-check (worklist of a search over <= 2^19 rows: ~0.3% of them): one CTA per 4-row group, nothing staged
    if (list != nullptr && N <= ((int64_t)1 << 19) && D <= kDxMaxD && K <= kDxMaxPerLane * 32 * kDxWarps) {
        int64_t g = (N / 64 + kDxR - 1) / kDxR;               // room for ~1.5% of the rows before CTAs take a second group
        if (g > (int64_t)sms * 8) g = (int64_t)sms * 8;
        if (g < 8) g = 8;
        vq_argmin_exact_direct_kernel<<<(unsigned)g, kDxThreads, 0, st>>>(z, N, D, HW, E, K, list, list_count, idx);
        MOVAE_CUDA_TRY(cudaGetLastError());
        return MOVAE_OK;
    }
    const int R = stage ? 4 : 1;
    int64_t grid = ((N + R - 1) / R + kExWarps - 1) / kExWarps;
    const int64_t cap = stage ? (int64_t)sms : (int64_t)sms * 4;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    static thread_local int configured_dev = -1;
    int dev = 0;
    MOVAE_CUDA_TRY(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        MOVAE_CUDA_TRY(cudaFuncSetAttribute(vq_argmin_exact_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kExMaxSmem));
        MOVAE_CUDA_TRY(cudaFuncSetAttribute(vq_argmin_exact_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kExMaxSmem));
        configured_dev = dev;
    }
    if (stage)
        vq_argmin_exact_kernel<true, 4><<<(unsigned)grid, kExThreads, staged, st>>>(z, N, D, HW, E, K, list, list_count, idx);
    else
        vq_argmin_exact_kernel<false, 1><<<(unsigned)grid, kExThreads, plain, st>>>(z, N, D, HW, E, K, list, list_count, idx);
    MOVAE_CUDA_TRY(cudaGetLastError());
    return MOVAE_OK;
}

}  // namespace movae
