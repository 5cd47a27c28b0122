// Device-side pieces of the Gramian pass, shared by K1 (`gram_kernel`, gram.cu) and phase 1 of the
// fused aggregation kernel (aggregate.cu).
//
//   * every CTA owns one contiguous span of columns (equal bytes per CTA) and walks it in tiles of 256 threads x U float4
//     per row; every warp load instruction covers 512 contiguous bytes of one row, all k rows of a tile are in flight
//     together (k*U independent 16-byte loads per thread);
//   * k(k+1)/2 float32 FMA chains per thread, at most 64 columns long, then promoted into float64
//     registers (the reference's own float32 SGEMM is 3e-5..8e-3 off at P=2.4M..1e8, SURVEY App. C.2;
//     the parity contract is rtol 1e-5 against a float64-accumulated oracle);
//   * warp shuffle -> shared memory -> one partial per CTA; the partials are summed in a fixed order
//     by ONE CTA: bit-reproducible for a given (k, P, grid), no float atomics.
#pragma once
#include "common.cuh"

namespace movae {

constexpr int kGramThreads = 256;
constexpr int kGramMaxBlocks = 2048;
constexpr int kGramHeaderBytes = 256;
constexpr int kGramChain = 64;   // float32 FMA chain length (columns) between promotions to float64

template <int K>
struct GramAcc {
    static constexpr int N = K * (K + 1) / 2;
};

template <int K>
__device__ __forceinline__ void gram_fma(float (&acc)[GramAcc<K>::N], const float (&x)[K]) {
    int a = 0;
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = i; j < K; ++j) { acc[a] = fmaf(x[i], x[j], acc[a]); ++a; }
}

// [lo, hi) float4 (or float) items of one row owned by CTA `b` of `g`: contiguous spans of equal byte size (so every CTA
// finishes together whatever P is -- a round-robin of whole tiles leaves a tail of up to one tile time in which only
// a fraction of the CTAs still stream: 9 % at P = 1e7), boundaries on multiples of 8 items (128-byte lines).
__device__ __forceinline__ void cta_span(int64_t n_items, int64_t& lo, int64_t& hi) {
    const int64_t b = blockIdx.x, g = gridDim.x;
    lo = b == 0 ? 0 : ((n_items * b / g) & ~(int64_t)7);
    hi = b == g - 1 ? n_items : ((n_items * (b + 1) / g) & ~(int64_t)7);
}

// Streams this CTA's span of J front to back and adds the products into acc64.
// VEC: J base 16-byte aligned and ldJ % 4 == 0 -> float4 path; otherwise scalar path.
template <int K, int U, bool VEC>
__device__ __forceinline__ void gram_stream_tiles(const float* __restrict__ J, int64_t P, int64_t ldJ,
                                                  double (&acc64)[GramAcc<K>::N]) {
    constexpr int NACC = GramAcc<K>::N;
    constexpr int W = VEC ? 4 : 1;                       // columns per item
    constexpr int FLUSH = kGramChain / (W * U) > 0 ? kGramChain / (W * U) : 1;
    const int tid = threadIdx.x;
    const int64_t n_items = P / W;                       // float4 (or float) items per row
    constexpr int64_t tile_items = (int64_t)kGramThreads * U;
    int64_t t0, hi;
    cta_span(n_items, t0, hi);

    while (t0 < hi) {
        float acc[NACC];
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc[a] = 0.f;
#pragma unroll 1
        for (int f = 0; f < FLUSH && t0 < hi; ++f, t0 += tile_items) {
            const int64_t base = t0 + tid;
            if constexpr (VEC) {
                float4 v[K][U];
                if (t0 + tile_items <= hi) {
#pragma unroll
                    for (int i = 0; i < K; ++i)
#pragma unroll
                        for (int u = 0; u < U; ++u)
                            v[i][u] = ld_stream_f4(reinterpret_cast<const float4*>(J + i * ldJ) + base + u * kGramThreads);
                } else {
#pragma unroll
                    for (int i = 0; i < K; ++i)
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int64_t idx = base + u * kGramThreads;
                            v[i][u] = idx < hi ? ld_stream_f4(reinterpret_cast<const float4*>(J + i * ldJ) + idx)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    float x[K];
#pragma unroll
                    for (int i = 0; i < K; ++i) x[i] = v[i][u].x;
                    gram_fma<K>(acc, x);
#pragma unroll
                    for (int i = 0; i < K; ++i) x[i] = v[i][u].y;
                    gram_fma<K>(acc, x);
#pragma unroll
                    for (int i = 0; i < K; ++i) x[i] = v[i][u].z;
                    gram_fma<K>(acc, x);
#pragma unroll
                    for (int i = 0; i < K; ++i) x[i] = v[i][u].w;
                    gram_fma<K>(acc, x);
                }
            } else {
                float v[K][U];
#pragma unroll
                for (int i = 0; i < K; ++i)
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int64_t idx = base + u * kGramThreads;
                        v[i][u] = idx < hi ? ld_stream_f1(J + i * ldJ + idx) : 0.f;
                    }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    float x[K];
#pragma unroll
                    for (int i = 0; i < K; ++i) x[i] = v[i][u];
                    gram_fma<K>(acc, x);
                }
            }
        }
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc64[a] += (double)acc[a];
    }

    // ragged tail of the float4 path: columns 4*(P/4) .. P-1, one thread each in CTA 0
    if (VEC && blockIdx.x == 0 && tid < (int)(P - n_items * W)) {
        float x[K];
        float acc[NACC];
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc[a] = 0.f;
#pragma unroll
        for (int i = 0; i < K; ++i) x[i] = J[i * ldJ + n_items * W + tid];
        gram_fma<K>(acc, x);
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc64[a] += (double)acc[a];
    }
}

// CTA reduce (shuffle within warps, fixed-order sum across the 8 warps), store this CTA's partial, take a
// ticket.  Returns true (to every thread) in the LAST CTA of the grid to arrive.  `red` is shared scratch.
template <int K>
__device__ __forceinline__ bool gram_cta_partial_and_ticket(const double (&acc64)[GramAcc<K>::N], double* __restrict__ partials,
                                                            unsigned int* __restrict__ counter,
                                                            double (*red)[GramAcc<K>::N], int* is_last_smem) {
    constexpr int NACC = GramAcc<K>::N;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        const double s = warp_sum(acc64[a]);
        if (lane == 0) red[warp][a] = s;
    }
    __syncthreads();
    if (tid < NACC) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kGramThreads / 32; ++w) s += red[w][tid];
        partials[(int64_t)blockIdx.x * NACC + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) *is_last_smem = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    return *is_last_smem != 0;
}

// Last CTA: deterministic combine of all CTA partials into the full symmetric k x k matrix `Gs` (shared memory,
// row-major K x K).  Ends with a CTA barrier.
template <int K>
__device__ __forceinline__ void gram_combine_partials(const double* __restrict__ partials, double* __restrict__ Gs) {
    constexpr int NACC = GramAcc<K>::N;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __threadfence();
    for (int a = warp; a < NACC; a += kGramThreads / 32) {
        double s = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(&partials[(int64_t)b * NACC + a]);
        s = warp_sum(s);
        if (lane == 0) {
            int i = 0, rem = a;          // a -> (i, j), i <= j, row-major upper triangle
            while (rem >= K - i) { rem -= K - i; ++i; }
            const int j = i + rem;
            Gs[i * K + j] = s;
            Gs[j * K + i] = s;
        }
    }
    __syncthreads();
}

}  // namespace movae
