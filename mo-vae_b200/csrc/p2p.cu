// Exchange-buffer management for the P-sharded aggregation path (CUDA IPC over NVLink peer memory)
// and the gather + solve entry point.  The kernels that use the buffers are K1's tail (gram.cu) and
// K2's head (solve.cu).
#include "common.cuh"

namespace movae {

void set_solve_p2p(const P2PArgs& px, double* G_sum);   // solve.cu

int make_p2p_args(const movae_p2p_ctx* ctx, uint64_t seq, P2PArgs* out) {
    MOVAE_REQUIRE(ctx != nullptr && out != nullptr, MOVAE_ERR_INVALID, "p2p: null context");
    MOVAE_REQUIRE(ctx->world >= 1 && ctx->world <= MOVAE_MAX_WORLD, MOVAE_ERR_UNSUPPORTED, "p2p: world size %d outside 1..%d",
                  ctx->world, MOVAE_MAX_WORLD);
    MOVAE_REQUIRE(ctx->rank >= 0 && ctx->rank < ctx->world, MOVAE_ERR_INVALID, "p2p: rank %d outside world %d", ctx->rank, ctx->world);
    MOVAE_REQUIRE(seq >= 1, MOVAE_ERR_INVALID, "p2p: seq must start at 1");
    *out = p2p_disabled();
    out->rank = ctx->rank;
    out->world = ctx->world;
    out->seq = seq;
    for (int r = 0; r < ctx->world; ++r) {
        MOVAE_REQUIRE(ctx->peers[r] != nullptr, MOVAE_ERR_INVALID, "p2p: peer %d buffer is null", r);
        out->peers[r] = static_cast<XchgBuffer*>(ctx->peers[r]);
    }
    return MOVAE_OK;
}

}  // namespace movae

extern "C" {

size_t movae_p2p_exchange_bytes(void) { return sizeof(movae::XchgBuffer); }

int movae_p2p_alloc(size_t bytes, void** d_ptr, unsigned char ipc_handle[64]) {
    using namespace movae;
    MOVAE_REQUIRE(d_ptr != nullptr && ipc_handle != nullptr && bytes > 0, MOVAE_ERR_INVALID, "p2p_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    MOVAE_CUDA_TRY(cudaMalloc(&p, bytes));
    MOVAE_CUDA_TRY(cudaMemset(p, 0, bytes));
    MOVAE_CUDA_TRY(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    MOVAE_CUDA_TRY(cudaIpcGetMemHandle(&h, p));
    memcpy(ipc_handle, &h, sizeof(h));
    *d_ptr = p;
    return MOVAE_OK;
}

int movae_p2p_open(const unsigned char ipc_handle[64], void** d_ptr) {
    using namespace movae;
    MOVAE_REQUIRE(d_ptr != nullptr && ipc_handle != nullptr, MOVAE_ERR_INVALID, "p2p_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    void* p = nullptr;
    MOVAE_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *d_ptr = p;
    return MOVAE_OK;
}

int movae_p2p_close(void* d_ptr) {
    if (d_ptr) MOVAE_CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
    return MOVAE_OK;
}

int movae_p2p_free(void* d_ptr) {
    if (d_ptr) MOVAE_CUDA_TRY(cudaFree(d_ptr));
    return MOVAE_OK;
}

int movae_solve_p2p(const movae_p2p_ctx* ctx, uint64_t seq, int k, const movae_solve_spec* spec, const float* d_vec,
                    float* d_w, double* d_diag, double* d_G_sum, void* stream) {
    movae::P2PArgs px;
    const int rc = movae::make_p2p_args(ctx, seq, &px);
    if (rc != MOVAE_OK) return rc;
    movae::set_solve_p2p(px, d_G_sum);
    const int rc2 = movae_solve(nullptr, k, spec, d_vec, d_w, d_diag, stream);
    movae::set_solve_p2p(movae::p2p_disabled(), nullptr);
    return rc2;
}

}  // extern "C"
