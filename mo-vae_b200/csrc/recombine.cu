// K3 `recombine_writeback`:  grad (=|+=) w @ J   for float32 J[k, P], w[k] resident on the device.
//
// Replaces torchjd `WeightedAggregator.forward` (`weights @ J`, cuBLAS SGEMV in the reference) fused
// with the split / reshape / `param.grad = slice.clone()` (or `+=`) chain that follows it (call sites
// /root/reference/main.py:189-196): the parameters' .grad tensors are views of ONE flat buffer, so
// the write-back is a single streaming store of that buffer instead of one copy kernel per tensor.
//
// Roofline: HBM.  Algorithmic traffic 4*k*P read + 4*P written (+4*P read when accumulating).
// Tiles are walked from the END of J backwards: K1 just streamed J front-to-back, so the last
// ~100 MB of it are still L2-resident on a 126 MB L2 and are consumed first.
#include "common.cuh"

namespace movae {

constexpr int kRecThreads = 256;

template <int K, int U, bool VEC>
__global__ void __launch_bounds__(kRecThreads)
recombine_kernel(const float* __restrict__ J, int64_t P, int64_t ldJ, const float* __restrict__ w_dev,
                 float* __restrict__ out, int accumulate) {
    constexpr int W = VEC ? 4 : 1;
    const int tid = threadIdx.x;
    const int64_t n_items = P / W;
    const int64_t tile_items = (int64_t)kRecThreads * U;
    const int64_t n_tiles = (n_items + tile_items - 1) / tile_items;

    float w[K];
#pragma unroll
    for (int i = 0; i < K; ++i) w[i] = w_dev[i];

    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t tile = n_tiles - 1 - t;
        const int64_t base = tile * tile_items + tid;
        if constexpr (VEC) {
            float4 v[K][U];
            const bool full = (base - tid + tile_items <= n_items);
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int64_t idx = base + u * kRecThreads;
                    v[i][u] = (full || idx < n_items) ? ld_stream_f4(reinterpret_cast<const float4*>(J + i * ldJ) + idx)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t idx = base + u * kRecThreads;
                if (!full && idx >= n_items) continue;
                float4 o;
                o.x = w[0] * v[0][u].x; o.y = w[0] * v[0][u].y; o.z = w[0] * v[0][u].z; o.w = w[0] * v[0][u].w;
#pragma unroll
                for (int i = 1; i < K; ++i) {
                    o.x = fmaf(w[i], v[i][u].x, o.x); o.y = fmaf(w[i], v[i][u].y, o.y);
                    o.z = fmaf(w[i], v[i][u].z, o.z); o.w = fmaf(w[i], v[i][u].w, o.w);
                }
                float4* dst = reinterpret_cast<float4*>(out) + idx;
                if (accumulate) {
                    const float4 old = *dst;
                    o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                st_stream_f4(dst, o);
            }
        } else {
            float v[K][U];
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int64_t idx = base + u * kRecThreads;
                    v[i][u] = idx < n_items ? ld_stream_f1(J + i * ldJ + idx) : 0.f;
                }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t idx = base + u * kRecThreads;
                if (idx >= n_items) continue;
                float o = w[0] * v[0][u];
#pragma unroll
                for (int i = 1; i < K; ++i) o = fmaf(w[i], v[i][u], o);
                if (accumulate) o += out[idx];
                out[idx] = o;
            }
        }
    }
    // ragged tail of the float4 path
    if (VEC && blockIdx.x == 0 && tid < (int)(P - n_items * W)) {
        const int64_t c = n_items * W + tid;
        float o = w[0] * J[c];
#pragma unroll
        for (int i = 1; i < K; ++i) o = fmaf(w[i], J[i * ldJ + c], o);
        if (accumulate) o += out[c];
        out[c] = o;
    }
}

template <int K, int U, bool VEC>
static int launch_recombine(const float* J, int64_t P, int64_t ldJ, const float* w, float* out, int accumulate,
                            cudaStream_t st) {
    auto kern = recombine_kernel<K, U, VEC>;
    static thread_local int occ = 0;
    if (occ == 0) {
        MOVAE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kRecThreads, 0));
        if (occ < 1) occ = 1;
    }
    const int sms = sm_count();
    MOVAE_REQUIRE(sms > 0, MOVAE_ERR_CUDA, "CUDA device query failed (no GPU?)");
    const int64_t n_items = P / (VEC ? 4 : 1);
    const int64_t tile_items = (int64_t)kRecThreads * U;
    int64_t n_tiles = (n_items + tile_items - 1) / tile_items;
    if (n_tiles < 1) n_tiles = 1;
    int64_t grid = (int64_t)sms * occ;
    if (grid > n_tiles) grid = n_tiles;
    kern<<<(unsigned)grid, kRecThreads, 0, st>>>(J, P, ldJ, w, out, accumulate);
    MOVAE_CUDA_TRY(cudaGetLastError());
    return MOVAE_OK;
}

template <int K>
static int dispatch_recombine(const float* J, int64_t P, int64_t ldJ, const float* w, float* out, int accumulate,
                              cudaStream_t st) {
    const bool vec = (reinterpret_cast<uintptr_t>(J) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
                     (ldJ % 4 == 0 || K == 1);
    if (vec) {
        if constexpr (K <= 4) return launch_recombine<K, 4, true>(J, P, ldJ, w, out, accumulate, st);
        else return launch_recombine<K, 2, true>(J, P, ldJ, w, out, accumulate, st);
    } else {
        if constexpr (K <= 4) return launch_recombine<K, 8, false>(J, P, ldJ, w, out, accumulate, st);
        else return launch_recombine<K, 4, false>(J, P, ldJ, w, out, accumulate, st);
    }
}

}  // namespace movae

extern "C" int movae_recombine_f32(const float* d_J, int k, int64_t P, int64_t ldJ, const float* d_w, float* d_grad,
                                   int accumulate, void* stream) {
    using namespace movae;
    MOVAE_REQUIRE(k >= 1, MOVAE_ERR_INVALID, "recombine: k must be >= 1 (got %d)", k);
    MOVAE_REQUIRE(k <= MOVAE_MAX_K, MOVAE_ERR_UNSUPPORTED, "recombine: k=%d > MOVAE_MAX_K=%d", k, MOVAE_MAX_K);
    MOVAE_REQUIRE(P >= 0 && ldJ >= P, MOVAE_ERR_INVALID, "recombine: need 0 <= P <= ldJ");
    if (P == 0) return MOVAE_OK;
    MOVAE_REQUIRE(d_J && d_w && d_grad, MOVAE_ERR_INVALID, "recombine: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (k) {
        case 1: return dispatch_recombine<1>(d_J, P, ldJ, d_w, d_grad, accumulate, st);
        case 2: return dispatch_recombine<2>(d_J, P, ldJ, d_w, d_grad, accumulate, st);
        case 3: return dispatch_recombine<3>(d_J, P, ldJ, d_w, d_grad, accumulate, st);
        case 4: return dispatch_recombine<4>(d_J, P, ldJ, d_w, d_grad, accumulate, st);
        case 5: return dispatch_recombine<5>(d_J, P, ldJ, d_w, d_grad, accumulate, st);
        case 6: return dispatch_recombine<6>(d_J, P, ldJ, d_w, d_grad, accumulate, st);
        case 7: return dispatch_recombine<7>(d_J, P, ldJ, d_w, d_grad, accumulate, st);
        default: return dispatch_recombine<8>(d_J, P, ldJ, d_w, d_grad, accumulate, st);
    }
}
