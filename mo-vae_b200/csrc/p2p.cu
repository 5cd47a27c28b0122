// Exchange-buffer management for the P-sharded aggregation path (CUDA IPC over NVLink peer memory)
// and the device-side barrier.  The exchange itself happens inside the fused aggregation kernel (aggregate.cu).
#include "common.cuh"

namespace movae {

int make_p2p_args(const movae_p2p_ctx* ctx, P2PArgs* out) {
    MOVAE_REQUIRE(ctx != nullptr && out != nullptr, MOVAE_ERR_INVALID, "p2p: null context");
    MOVAE_REQUIRE(ctx->world >= 1 && ctx->world <= MOVAE_MAX_WORLD, MOVAE_ERR_UNSUPPORTED, "p2p: world size %d outside 1..%d",
                  ctx->world, MOVAE_MAX_WORLD);
    MOVAE_REQUIRE(ctx->rank >= 0 && ctx->rank < ctx->world, MOVAE_ERR_INVALID, "p2p: rank %d outside world %d", ctx->rank, ctx->world);
    *out = p2p_disabled();
    out->rank = ctx->rank;
    out->world = ctx->world;
    for (int r = 0; r < ctx->world; ++r) {
        MOVAE_REQUIRE(ctx->peers[r] != nullptr, MOVAE_ERR_INVALID, "p2p: peer %d buffer is null", r);
        out->peers[r] = static_cast<XchgBuffer*>(ctx->peers[r]);
    }
    return MOVAE_OK;
}

// Device-side barrier over the exchange buffers: every rank raises its epoch in every peer's buffer and waits for all
// of them in its own.  One tiny CTA; aligns the GPUs' streams to within a flag round trip (~2 us over NVLink) without
// involving the hosts, and is CUDA-graph capturable.  d_status (may be NULL) receives 1 when a peer never arrived.
__global__ void p2p_barrier_kernel(P2PArgs px, int* __restrict__ d_status) {
    __shared__ unsigned long long epoch;
    __shared__ int failed;
    XchgBuffer* own = px.peers[px.rank];
    const int tid = threadIdx.x;
    if (tid == 0) { epoch = own->bar_epoch + 1; failed = 0; }
    __syncthreads();
    if (tid < px.world) {
        st_release_sys_u64(&px.peers[tid]->bar_flags[px.rank], epoch);
        if (!wait_flag_sys(&own->bar_flags[tid], epoch, kExchangeTimeoutNs)) failed = 1;
    }
    __syncthreads();
    if (tid == 0) {
        own->bar_epoch = epoch;
        if (d_status) *d_status = failed;
    }
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

int movae_p2p_barrier(const movae_p2p_ctx* ctx, int* d_status, void* stream) {
    movae::P2PArgs px;
    const int rc = movae::make_p2p_args(ctx, &px);
    if (rc != MOVAE_OK) return rc;
    movae::p2p_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(px, d_status);
    MOVAE_CUDA_TRY(cudaGetLastError());
    return MOVAE_OK;
}

}  // extern "C"
