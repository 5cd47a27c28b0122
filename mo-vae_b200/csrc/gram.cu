// K1 `gram_stream`:  G = J J^T  for a float32 row-major Jacobian J[k, P] (k <= 8 objectives).
//
// Replaces torchjd `compute_gramian` (`J @ J.T`, a cuBLAS SGEMM with M=N=k in the reference; call
// sites /root/reference/main.py:189-196 through GramianWeightedAggregator, aligned_mtl.py:39,
// mgda.py:12, nupgrad.py:37).
//
// Roofline: HBM.  Algorithmic traffic 4*k*P bytes read, k*k doubles written.  One pass:
//   * persistent grid (SMs x resident CTAs), CTA tile = 256 threads x U float4 per row, every warp
//     load instruction covers 512 contiguous bytes of one row, all k rows of a tile are in flight
//     together (k*U independent 16-byte loads per thread);
//   * k(k+1)/2 float32 FMA chains per thread, at most 64 columns long, then promoted into float64
//     registers (the reference's own float32 SGEMM is 3e-5..8e-3 off at P=2.4M..1e8, SURVEY App. C.2;
//     the parity contract is rtol 1e-5 against a float64-accumulated oracle);
//   * warp shuffle -> shared memory -> one partial per CTA -> the last CTA to finish (atomic ticket)
//     sums the partials in a fixed order: bit-reproducible for a given (k, P, grid), no float atomics.
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

// VEC: J base 16-byte aligned and ldJ % 4 == 0 -> float4 path; otherwise scalar path.
template <int K, int U, bool VEC, int MINB>
__global__ void __launch_bounds__(kGramThreads, MINB)
gram_kernel(const float* __restrict__ J, int64_t P, int64_t ldJ, double* __restrict__ partials,
            unsigned int* __restrict__ counter, double* __restrict__ G, int accumulate, P2PArgs px) {
    constexpr int NACC = GramAcc<K>::N;
    constexpr int W = VEC ? 4 : 1;                       // columns per item
    constexpr int FLUSH = kGramChain / (W * U) > 0 ? kGramChain / (W * U) : 1;
    const int tid = threadIdx.x;
    const int64_t n_items = P / W;                       // float4 (or float) items per row
    const int64_t tile_items = (int64_t)kGramThreads * U;
    const int64_t n_tiles = (n_items + tile_items - 1) / tile_items;

    double acc64[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc64[a] = 0.0;

    int64_t tile = blockIdx.x;
    while (tile < n_tiles) {
        float acc[NACC];
#pragma unroll
        for (int a = 0; a < NACC; ++a) acc[a] = 0.f;
#pragma unroll 1
        for (int f = 0; f < FLUSH && tile < n_tiles; ++f, tile += gridDim.x) {
            const int64_t base = tile * tile_items + tid;
            if constexpr (VEC) {
                float4 v[K][U];
                if (base - tid + tile_items <= n_items) {
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
                            v[i][u] = idx < n_items ? ld_stream_f4(reinterpret_cast<const float4*>(J + i * ldJ) + idx)
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
                        v[i][u] = idx < n_items ? ld_stream_f1(J + i * ldJ + idx) : 0.f;
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

    // ---- CTA reduce: shuffle within warps, fixed-order sum across the 8 warps -------------------
    __shared__ double red[kGramThreads / 32][NACC];
    __shared__ int is_last;
    const int warp = tid >> 5, lane = tid & 31;
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
    if (tid == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;

    // ---- last CTA: deterministic combine of all CTA partials -------------------------------------
    __threadfence();
    for (int a = warp; a < NACC; a += kGramThreads / 32) {
        double s = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(&partials[(int64_t)b * NACC + a]);
        s = warp_sum(s);
        if (lane == 0) {
            int i = 0, rem = a;          // a -> (i, j), i <= j, row-major upper triangle
            while (rem >= K - i) { rem -= K - i; ++i; }
            const int j = i + rem;
            if (accumulate) {
                G[i * K + j] += s;
                if (i != j) G[j * K + i] += s;
            } else {
                G[i * K + j] = s;
                G[j * K + i] = s;
            }
        }
    }
    if (tid == 0) *counter = 0u;   // self-reset: the workspace is reusable by the next launch

    // ---- fused exchange tail (P-sharded aggregation): publish this rank's Gramian partial to every peer ----
    if (px.world > 0) {
        __syncthreads();                                   // G complete and visible to this CTA
        const int par = (int)(px.seq & 1ull);
        if (tid < K * K) {
            const double v = G[tid];
            for (int r = 0; r < px.world; ++r) px.peers[r]->slots[par][px.rank][tid] = v;   // peer-to-peer stores over NVLink
        }
        __threadfence_system();
        __syncthreads();
        if (tid < px.world) st_release_sys_u64(&px.peers[tid]->flags[par][px.rank], px.seq);
    }
}

template <int K, int U, bool VEC, int MINB>
static int launch_gram(const float* J, int64_t P, int64_t ldJ, double* G, int accumulate, void* ws, cudaStream_t st,
                       const P2PArgs& px) {
    auto kern = gram_kernel<K, U, VEC, MINB>;
    static thread_local int occ = 0;
    if (occ == 0) {
        MOVAE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kGramThreads, 0));
        if (occ < 1) occ = 1;
    }
    const int sms = sm_count();
    MOVAE_REQUIRE(sms > 0, MOVAE_ERR_CUDA, "CUDA device query failed (no GPU?)");
    const int64_t n_items = P / (VEC ? 4 : 1);
    const int64_t tile_items = (int64_t)kGramThreads * U;
    int64_t n_tiles = (n_items + tile_items - 1) / tile_items;
    if (n_tiles < 1) n_tiles = 1;
    int64_t grid = (int64_t)sms * occ;
    if (grid > n_tiles) grid = n_tiles;
    if (grid > kGramMaxBlocks) grid = kGramMaxBlocks;
    unsigned int* counter = reinterpret_cast<unsigned int*>(ws);
    double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + kGramHeaderBytes);
    kern<<<(unsigned)grid, kGramThreads, 0, st>>>(J, P, ldJ, partials, counter, G, accumulate, px);
    MOVAE_CUDA_TRY(cudaGetLastError());
    return MOVAE_OK;
}

template <int K>
static int dispatch_gram(const float* J, int64_t P, int64_t ldJ, double* G, int accumulate, void* ws, cudaStream_t st,
                         const P2PArgs& px) {
    const bool vec = (reinterpret_cast<uintptr_t>(J) % 16 == 0) && (ldJ % 4 == 0 || K == 1);
    // registers: float64 accumulators cost 2*K(K+1)/2; small k affords deeper unroll and 2+ CTAs/SM
    if (vec) {
        if constexpr (K <= 2) return launch_gram<K, 8, true, 2>(J, P, ldJ, G, accumulate, ws, st, px);
        else if constexpr (K <= 4) return launch_gram<K, 4, true, 2>(J, P, ldJ, G, accumulate, ws, st, px);
        else return launch_gram<K, 2, true, 1>(J, P, ldJ, G, accumulate, ws, st, px);
    } else {
        if constexpr (K <= 4) return launch_gram<K, 8, false, 2>(J, P, ldJ, G, accumulate, ws, st, px);
        else return launch_gram<K, 4, false, 1>(J, P, ldJ, G, accumulate, ws, st, px);
    }
}

}  // namespace movae

extern "C" {

size_t movae_gram_workspace_bytes(int k) {
    if (k < 1 || k > MOVAE_MAX_K) return 0;
    return (size_t)movae::kGramHeaderBytes + (size_t)movae::kGramMaxBlocks * (k * (k + 1) / 2) * sizeof(double);
}

static int gram_entry(const float* d_J, int k, int64_t P, int64_t ldJ, double* d_G, int accumulate, void* d_ws, size_t ws_bytes,
                      void* stream, const movae::P2PArgs& px) {
    using namespace movae;
    MOVAE_REQUIRE(k >= 1, MOVAE_ERR_INVALID, "gram: k must be >= 1 (got %d)", k);
    MOVAE_REQUIRE(k <= MOVAE_MAX_K, MOVAE_ERR_UNSUPPORTED, "gram: k=%d > MOVAE_MAX_K=%d", k, MOVAE_MAX_K);
    MOVAE_REQUIRE(P >= 0 && ldJ >= P, MOVAE_ERR_INVALID, "gram: need 0 <= P <= ldJ (P=%lld ldJ=%lld)", (long long)P,
                  (long long)ldJ);
    MOVAE_REQUIRE(d_G != nullptr && (d_J != nullptr || P == 0), MOVAE_ERR_INVALID, "gram: null pointer");
    MOVAE_REQUIRE(d_ws != nullptr && ws_bytes >= movae_gram_workspace_bytes(k), MOVAE_ERR_WORKSPACE,
                  "gram: workspace too small (%zu < %zu)", ws_bytes, movae_gram_workspace_bytes(k));
    MOVAE_REQUIRE(reinterpret_cast<uintptr_t>(d_ws) % 8 == 0, MOVAE_ERR_WORKSPACE, "gram: workspace must be 8-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (k) {
        case 1: return dispatch_gram<1>(d_J, P, ldJ, d_G, accumulate, d_ws, st, px);
        case 2: return dispatch_gram<2>(d_J, P, ldJ, d_G, accumulate, d_ws, st, px);
        case 3: return dispatch_gram<3>(d_J, P, ldJ, d_G, accumulate, d_ws, st, px);
        case 4: return dispatch_gram<4>(d_J, P, ldJ, d_G, accumulate, d_ws, st, px);
        case 5: return dispatch_gram<5>(d_J, P, ldJ, d_G, accumulate, d_ws, st, px);
        case 6: return dispatch_gram<6>(d_J, P, ldJ, d_G, accumulate, d_ws, st, px);
        case 7: return dispatch_gram<7>(d_J, P, ldJ, d_G, accumulate, d_ws, st, px);
        default: return dispatch_gram<8>(d_J, P, ldJ, d_G, accumulate, d_ws, st, px);
    }
}

int movae_gram_f32(const float* d_J, int k, int64_t P, int64_t ldJ, double* d_G, int accumulate, void* d_ws,
                   size_t ws_bytes, void* stream) {
    return gram_entry(d_J, k, P, ldJ, d_G, accumulate, d_ws, ws_bytes, stream, movae::p2p_disabled());
}

int movae_gram_publish_f32(const float* d_J, int k, int64_t P, int64_t ldJ, double* d_G, int accumulate, void* d_ws,
                           size_t ws_bytes, const movae_p2p_ctx* ctx, uint64_t seq, void* stream) {
    movae::P2PArgs px;
    const int rc = movae::make_p2p_args(ctx, seq, &px);
    if (rc != MOVAE_OK) return rc;
    return gram_entry(d_J, k, P, ldJ, d_G, accumulate, d_ws, ws_bytes, stream, px);
}

}  // extern "C"
