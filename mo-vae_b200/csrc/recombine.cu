// K3 `recombine_writeback`:  grad (=|+=) w @ J   for float32 J[k, P], w[k] resident on the device.
//
// Replaces torchjd `WeightedAggregator.forward` (`weights @ J`, cuBLAS SGEMV in the reference) fused
// with the split / reshape / `param.grad = slice.clone()` (or `+=`) chain that follows it (call sites
// /root/reference/main.py:189-196): the parameters' .grad tensors are views of ONE flat buffer, so
// the write-back is a single streaming store of that buffer instead of one copy kernel per tensor.
//
// Roofline: HBM.  Algorithmic traffic 4*k*P read + 4*P written (+4*P read when accumulating).
// J is walked from its END backwards: K1 just streamed J front-to-back, so the last ~100 MB of it are still
// L2-resident on a 126 MB L2 and are consumed first.
#include "recombine_device.cuh"

namespace movae {

template <int K, int U, bool VEC>
__global__ void __launch_bounds__(kRecThreads)
recombine_kernel(const float* __restrict__ J, int64_t P, int64_t ldJ, const float* __restrict__ w_dev,
                 float* __restrict__ out, int accumulate) {
    recombine_tiles<K, U, VEC>(J, P, ldJ, w_dev, out, accumulate, [] {});
}

template <int K, int U, bool VEC>
static int launch_recombine(const float* J, int64_t P, int64_t ldJ, const float* w, float* out, int accumulate,
                            cudaStream_t st) {
    auto kern = recombine_kernel<K, U, VEC>;
    static thread_local int occ_dev = -1, occ = 0;      // cached per (host thread, device)
    int dev = 0;
    MOVAE_CUDA_TRY(cudaGetDevice(&dev));
    if (occ_dev != dev) {
        MOVAE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kRecThreads, 0));
        if (occ < 1) occ = 1;
        occ_dev = dev;
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
