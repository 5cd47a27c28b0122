// The host-buffer pipeline behind bench.py's `e2e` number: Jacobian rows
// that live in (pinned) HOST memory are streamed to the GPU in column chunks while K1 consumes the
// chunks already landed; after the solve, K3 chunks are streamed back the same way.  PCIe-bound by
// construction (4kP bytes in, 4P bytes out); the two copy directions and the kernels overlap.
#include "common.cuh"

namespace movae {

constexpr int kMaxEvents = 256;

static cudaEvent_t* event_pool() {
    // one pool per (host thread, device): an event belongs to the device that was current when it was created
    constexpr int kMaxDevices = 16;
    static thread_local cudaEvent_t pool[kMaxDevices][kMaxEvents];
    static thread_local bool ready[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    if (!ready[dev]) {
        for (int i = 0; i < kMaxEvents; ++i)
            if (cudaEventCreateWithFlags(&pool[dev][i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        ready[dev] = true;
    }
    return pool[dev];
}

}  // namespace movae

extern "C" {

int movae_host_gram_f32(const float* h_J, int k, int64_t P, int64_t h_ld, float* d_J, int64_t d_ld, double* d_G,
                        void* d_ws, size_t ws_bytes, int64_t chunk_cols, void* compute_stream, void* copy_stream) {
    using namespace movae;
    MOVAE_REQUIRE(k >= 1 && k <= MOVAE_MAX_K, MOVAE_ERR_UNSUPPORTED, "host_gram: k=%d outside 1..%d", k, MOVAE_MAX_K);
    MOVAE_REQUIRE(P >= 0 && h_ld >= P && d_ld >= P && d_ld % 4 == 0, MOVAE_ERR_INVALID,
                  "host_gram: need P <= h_ld, P <= d_ld, d_ld %% 4 == 0");
    MOVAE_REQUIRE(h_J && d_J && d_G, MOVAE_ERR_INVALID, "host_gram: null pointer");
    MOVAE_REQUIRE(chunk_cols >= 4 && chunk_cols % 4 == 0, MOVAE_ERR_INVALID, "host_gram: chunk_cols must be a positive multiple of 4");
    cudaStream_t cs = static_cast<cudaStream_t>(compute_stream), xs = static_cast<cudaStream_t>(copy_stream);
    cudaEvent_t* ev = event_pool();
    MOVAE_REQUIRE(ev != nullptr, MOVAE_ERR_CUDA, "CUDA event pool creation failed");
    MOVAE_CUDA_TRY(cudaMemsetAsync(d_G, 0, sizeof(double) * k * k, cs));
    // copies must not start before earlier work on the compute stream that may still read d_J
    MOVAE_CUDA_TRY(cudaEventRecord(ev[kMaxEvents - 1], cs));
    MOVAE_CUDA_TRY(cudaStreamWaitEvent(xs, ev[kMaxEvents - 1], 0));
    int n = 0;
    for (int64_t c0 = 0; c0 < P; c0 += chunk_cols, ++n) {
        const int64_t cols = (P - c0 < chunk_cols) ? (P - c0) : chunk_cols;
        MOVAE_CUDA_TRY(cudaMemcpy2DAsync(d_J + c0, sizeof(float) * d_ld, h_J + c0, sizeof(float) * h_ld, sizeof(float) * cols,
                                         (size_t)k, cudaMemcpyHostToDevice, xs));
        cudaEvent_t e = ev[n % (kMaxEvents - 1)];
        MOVAE_CUDA_TRY(cudaEventRecord(e, xs));
        MOVAE_CUDA_TRY(cudaStreamWaitEvent(cs, e, 0));
        const int rc = movae_gram_f32(d_J + c0, k, cols, d_ld, d_G, 1, d_ws, ws_bytes, compute_stream);
        if (rc != MOVAE_OK) return rc;
    }
    return MOVAE_OK;
}

static int host_recombine(const float* d_J, int k, int64_t P, int64_t d_ld, const float* d_w, float* d_grad, float* h_grad,
                          int64_t chunk_cols, void* compute_stream, void* copy_stream, bool synchronous) {
    using namespace movae;
    MOVAE_REQUIRE(k >= 1 && k <= MOVAE_MAX_K, MOVAE_ERR_UNSUPPORTED, "host_recombine: k=%d outside 1..%d", k, MOVAE_MAX_K);
    MOVAE_REQUIRE(d_J && d_w && d_grad && h_grad, MOVAE_ERR_INVALID, "host_recombine: null pointer");
    MOVAE_REQUIRE(chunk_cols >= 4 && chunk_cols % 4 == 0, MOVAE_ERR_INVALID, "host_recombine: chunk_cols must be a positive multiple of 4");
    cudaStream_t cs = static_cast<cudaStream_t>(compute_stream), xs = static_cast<cudaStream_t>(copy_stream);
    cudaEvent_t* ev = event_pool();
    MOVAE_REQUIRE(ev != nullptr, MOVAE_ERR_CUDA, "CUDA event pool creation failed");
    int n = 0;
    for (int64_t c0 = 0; c0 < P; c0 += chunk_cols, ++n) {
        const int64_t cols = (P - c0 < chunk_cols) ? (P - c0) : chunk_cols;
        const int rc = movae_recombine_f32(d_J + c0, k, cols, d_ld, d_w, d_grad + c0, 0, compute_stream);
        if (rc != MOVAE_OK) return rc;
        cudaEvent_t e = ev[n % (kMaxEvents - 1)];
        MOVAE_CUDA_TRY(cudaEventRecord(e, cs));
        MOVAE_CUDA_TRY(cudaStreamWaitEvent(xs, e, 0));
        MOVAE_CUDA_TRY(cudaMemcpyAsync(h_grad + c0, d_grad + c0, sizeof(float) * cols, cudaMemcpyDeviceToHost, xs));
    }
    if (synchronous) {
        MOVAE_CUDA_TRY(cudaStreamSynchronize(xs));
        MOVAE_CUDA_TRY(cudaStreamSynchronize(cs));
    }
    return MOVAE_OK;
}

int movae_host_recombine_f32(const float* d_J, int k, int64_t P, int64_t d_ld, const float* d_w, float* d_grad,
                             float* h_grad, int64_t chunk_cols, void* compute_stream, void* copy_stream) {
    return host_recombine(d_J, k, P, d_ld, d_w, d_grad, h_grad, chunk_cols, compute_stream, copy_stream, true);
}

int movae_host_recombine_async_f32(const float* d_J, int k, int64_t P, int64_t d_ld, const float* d_w, float* d_grad,
                                   float* h_grad, int64_t chunk_cols, void* compute_stream, void* d2h_stream) {
    return host_recombine(d_J, k, P, d_ld, d_w, d_grad, h_grad, chunk_cols, compute_stream, d2h_stream, false);
}

}  // extern "C"
