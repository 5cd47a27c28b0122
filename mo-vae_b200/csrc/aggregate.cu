// Fused aggregation step:  g (=|+=) w(J J^T) @ J  in ONE persistent launch per GPU.
//
// Replaces the whole chain the reference runs per training step behind `aggregator(J)` (call sites
// /root/reference/main.py:189-196): torchjd compute_gramian (`J @ J.T`), the weighting (UPGrad QPs on the host,
// utils/torchmoo/mgda.py:221-272, aligned_mtl.py:97-133, nupgrad.py:122-158, comfort.py:148-158), `weights @ J` and the
// split / `.grad` write-back -- and, across GPUs, the k x k exchange of the P-sharded path.
//
// One grid of (SMs x resident CTAs) co-resident CTAs (cooperative launch), three phases:
//   1. every CTA streams its share of J (tiles dealt round-robin, front to back) into k(k+1)/2 Gramian partials
//      (gram_device.cuh) and takes a ticket;
//   2. the LAST CTA to arrive sums the partials in a fixed order, (multi-GPU) stores this rank's k x k float64 partial into
//      every peer's exchange buffer over NVLink (peer-to-peer stores + release flag), waits for the peers' flags, sums
//      the ranks' partials in RANK ORDER -- bit-identical input on every rank --, runs the small solve (solve_device.cuh)
//      and publishes w with a release store; the other CTAs have meanwhile issued the loads of their first phase-3 tile
//      and poll that flag;
//   3. every CTA walks its tiles back to front (the end of J is still in L2), FMAs with w and streams the result into
//      the flat .grad buffer (recombine_device.cuh).
// (A dedicated solver CTA that pre-runs the solve on a dummy Gramian to warm its instruction cache was tried: the solve
// phase stayed at 10.6 us for k = 3 -- it is bound by its dependent float64 divide / sqrt chains, not by instruction fetch.)
// Versus K1 -> K2 -> K3 as three launches this removes two launch gaps and a kernel ramp per step (they are 40 % of a
// step at P = 1e7) and, in the P-sharded path, the separate collective.  Nothing about a step is a host-side kernel
// argument (the exchange sequence number lives in the exchange buffer): the launch is CUDA-graph capturable.
//
// Roofline: HBM.  Algorithmic traffic 4*k*P (phase 1) + 4*k*P + 4*P (phase 3) bytes.
// Workspace: movae_gram_workspace_bytes(k), zero-filled once; header: byte 0 ticket, 4 exit counter, 128 ready flag,
// 192..255 timestamps of the last launch (globaltimer ns: start, all partials in, weights published, end, partials
// combined, exchange done).
#include "recombine_device.cuh"
#include "solve_device.cuh"

namespace movae {

int fill_solve_params(int k, const movae_solve_spec* spec, SolveParams* out);          // solve.cu
int check_solve_vectors(const SolveParams& p, const float* d_vec, const float* d_aux);  // solve.cu

constexpr int kAggThreads = 256;
static_assert(kAggThreads == kGramThreads && kAggThreads == kRecThreads && kAggThreads == kSolveThreads, "one CTA shape");

struct AggArgs {
    const float* J;
    int64_t P, ldJ;
    unsigned char* ws;
    const float* vec;
    const float* aux;
    float* out;           // nullptr: weights only (phases 1 and 2)
    int accumulate;
    float* w_out;         // [k] (COMFORT: [2k], the MGDA weights second)
    double* diag_out;     // [MOVAE_DIAG_DOUBLES] or nullptr
    double* G_out;        // [k*k] summed Gramian, or nullptr
};

// Phase 2 of the fused launches, executed by the LAST CTA to arrive: fixed-order combine of the CTA partials, (multi-GPU) the
// k x k exchange over peer memory, the small solve, publication of w through a release store on `ready`.
template <int K>
__device__ __forceinline__ void aggregate_solve_phase(const SolveParams& sp, const P2PArgs& px, const float* vec, const float* aux,
                                                      float* w_out, double* diag_out, double* G_out, const double* partials,
                                                      unsigned int* ticket, unsigned int* ready, unsigned int ready0,
                                                      unsigned long long* stamps, unsigned long long t_start) {
    __shared__ SolveSmem S;
    __shared__ double Gs[K * K];
    __shared__ unsigned long long seq_s;
    __shared__ int failed_s;
    const int tid = threadIdx.x;
    const unsigned long long t_in = global_timer_ns();
    gram_combine_partials<K>(partials, Gs);
    if (tid == 0) { *ticket = 0u; failed_s = 0; stamps[4] = global_timer_ns(); }   // self-reset: the workspace is reusable
    if (px.world > 0) {
        XchgBuffer* own = px.peers[px.rank];
        if (tid == 0) seq_s = own->step + 1;
        __syncthreads();
        const unsigned long long seq = seq_s;
        const int par = (int)(seq & 1ull);
        if (tid < K * K) {
            const double v = Gs[tid];
            for (int r = 0; r < px.world; ++r) st_relaxed_sys_f64(&px.peers[r]->slots[par][px.rank][tid], v);   // NVLink stores
        }
        __threadfence_system();
        __syncthreads();
        if (tid < px.world) st_release_sys_u64(&px.peers[tid]->flags[par][px.rank], seq);
        if (tid == 0) own->step = seq;
        // one waiter per peer flag (round-2 first version: every summing thread polled all flags in turn, 8 dependent
        // system-scope loads at 8 ranks); the CTA barrier carries the acquired visibility over to the summing threads
        if (tid < px.world && !wait_flag_sys(&own->flags[par][tid], seq, kExchangeTimeoutNs)) failed_s = 1;
        __syncthreads();
        if (tid < K * K) {
            double g = 0.0;
            for (int r = 0; r < px.world; ++r) g += ld_relaxed_sys_f64(&own->slots[par][r][tid]);   // rank order: the same sum on every rank
            Gs[tid] = g;
        }
        __syncthreads();
    }
    if (tid == 0) stamps[5] = global_timer_ns();
    if (tid < MK * MK) {
        const int i = tid / MK, j = tid % MK;
        S.G[i][j] = (i < K && j < K) ? Gs[i * K + j] : 0.0;
    }
    if (G_out && tid < K * K) G_out[tid] = Gs[tid];
    __syncthreads();
    solve_block<K>(sp, S, vec, aux, failed_s != 0, tid);
    if (tid < K) {
        w_out[tid] = S.w[tid];
        if (sp.comfort) w_out[K + tid] = S.w2[tid];
    }
    __syncthreads();
    if (tid == 0) {
        stamps[0] = t_start;
        stamps[1] = t_in;
        stamps[2] = global_timer_ns();
        __threadfence();
        st_release_gpu_u32(ready, ready0 + 1u);          // the other CTAs start their recombination pass now
    }
    solve_diagnostics(sp, S, tid);                       // nobody waits for these
    __syncthreads();
    if (tid < MOVAE_DIAG_DOUBLES && diag_out) diag_out[tid] = S.dg[tid];
}

// The other CTAs' wait for the weights (thread 0 polls, the CTA follows).
__device__ __forceinline__ void aggregate_wait_ready(const unsigned int* ready, unsigned int ready0) {
    if (threadIdx.x == 0) {
        // the solving CTA is resident (cooperative launch) and its only unbounded wait -- the peers' flags -- times out by
        // itself; the bound here is a last line of defence against a hung device, not a code path
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_gpu_u32(ready) == ready0) {
            __nanosleep(40);
            if (global_timer_ns() - t0 > 2ull * kExchangeTimeoutNs) break;
        }
    }
    __syncthreads();
}

template <int K, int U1, int U2, bool VEC, int MINB>
__global__ void __launch_bounds__(kAggThreads, MINB)
aggregate_kernel(AggArgs a, SolveParams sp, P2PArgs px) {
    constexpr int NACC = GramAcc<K>::N;
    const int tid = threadIdx.x;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(a.ws);
    unsigned int* done = reinterpret_cast<unsigned int*>(a.ws + 4);
    unsigned int* ready = reinterpret_cast<unsigned int*>(a.ws + 128);
    unsigned long long* stamps = reinterpret_cast<unsigned long long*>(a.ws + 192);
    double* partials = reinterpret_cast<double*>(a.ws + kGramHeaderBytes);

    __shared__ unsigned int ready0;
    __shared__ unsigned long long t_start;
    __shared__ double red[kGramThreads / 32][NACC];
    __shared__ int is_last;
    if (tid == 0) {
        ready0 = ld_acquire_gpu_u32(ready);      // the flag's value before this launch publishes its weights
        t_start = global_timer_ns();
    }

    // ---- phase 1: Gramian partials ----------------------------------------------------------------------------------
    double acc64[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc64[i] = 0.0;
    gram_stream_tiles<K, U1, VEC>(a.J, a.P, a.ldJ, acc64);
    const bool last = gram_cta_partial_and_ticket<K>(acc64, partials, ticket, red, &is_last);

    // ---- phase 2 (last CTA to arrive): combine, exchange, solve, publish --------------------------------------------------
    if (last) aggregate_solve_phase<K>(sp, px, a.vec, a.aux, a.w_out, a.diag_out, a.G_out, partials, ticket, ready, ready0, stamps, t_start);
    if (a.out == nullptr) return;

    // ---- phase 3: recombine + write-back -----------------------------------------------------------------------------------
    recombine_tiles<K, U2, VEC>(a.J, a.P, a.ldJ, a.w_out, a.out, a.accumulate, [&] {
        if (!last) aggregate_wait_ready(ready, ready0);
    });
    // end-of-launch timestamp by the last CTA to leave (diagnostics only)
    __syncthreads();
    if (tid == 0) {
        if (atomicAdd(done, 1u) == gridDim.x - 1) {
            stamps[3] = global_timer_ns();
            *done = 0u;
        }
    }
}

template <int K, int U1, int U2, bool VEC, int MINB>
static int launch_aggregate(const AggArgs& a, const SolveParams& sp, const P2PArgs& px, cudaStream_t st) {
    auto kern = aggregate_kernel<K, U1, U2, VEC, MINB>;
    static thread_local int occ_dev = -1, occ = 0;
    int dev = 0;
    MOVAE_CUDA_TRY(cudaGetDevice(&dev));
    if (occ_dev != dev) {
        MOVAE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kAggThreads, 0));
        if (occ < 1) occ = 1;
        occ_dev = dev;
    }
    const int sms = sm_count();
    MOVAE_REQUIRE(sms > 0, MOVAE_ERR_CUDA, "CUDA device query failed (no GPU?)");
    const int64_t n_items = a.P / (VEC ? 4 : 1);
    const int64_t tile_items = (int64_t)kAggThreads * (U1 > U2 ? U1 : U2);
    int64_t n_tiles = (n_items + tile_items - 1) / tile_items;
    if (n_tiles < 1) n_tiles = 1;
    int64_t grid = (int64_t)sms * occ;       // every CTA must be resident: the CTAs wait for each other
    if (grid > n_tiles) grid = n_tiles;
    if (grid > kGramMaxBlocks) grid = kGramMaxBlocks;
    AggArgs a_ = a;
    SolveParams sp_ = sp;
    P2PArgs px_ = px;
    void* params[] = {&a_, &sp_, &px_};
    MOVAE_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3((unsigned)grid), dim3(kAggThreads), params, 0, st));
    return MOVAE_OK;
}

template <int K>
static int dispatch_aggregate(const AggArgs& a, const SolveParams& sp, const P2PArgs& px, cudaStream_t st) {
    const bool vec = (reinterpret_cast<uintptr_t>(a.J) % 16 == 0) && (a.out == nullptr || reinterpret_cast<uintptr_t>(a.out) % 16 == 0) &&
                     (a.ldJ % 4 == 0 || K == 1);
    // registers: float64 accumulators cost 2*K(K+1)/2; small k affords deeper unroll and 2 CTAs/SM
    if (vec) {
        if constexpr (K <= 2) return launch_aggregate<K, 8, 4, true, 2>(a, sp, px, st);
        else if constexpr (K == 3) return launch_aggregate<K, 4, 4, true, 2>(a, sp, px, st);
        else if constexpr (K == 4) return launch_aggregate<K, 4, 2, true, 2>(a, sp, px, st);     // two phase-3 register tiles of U2 = 4 spill at 128 regs
        else return launch_aggregate<K, 2, 2, true, 1>(a, sp, px, st);
    } else {
        if constexpr (K <= 4) return launch_aggregate<K, 8, 8, false, 2>(a, sp, px, st);
        else return launch_aggregate<K, 4, 4, false, 1>(a, sp, px, st);
    }
}

// ---- segmented Jacobian: the rows are the gradient tensors autograd produced, wherever they live -----------------------
// Replaces the flatten + `cat` of torchjd's autojac (call sites main.py:189-196) WITHOUT the copy into a flat J: segment s
// (one shared parameter tensor) has k row pointers, n[s] columns and an offset into the flat gradient buffer.  Same three
// phases as aggregate_kernel; virtual tiles of 256 x U float4 are numbered segment after segment and dealt round-robin.
struct SegArgs {
    movae_jac_segments segs;
    unsigned char* ws;
    const float* vec;
    const float* aux;
    float* out;
    int accumulate;
    float* w_out;
    double* diag_out;
    double* G_out;
};

template <int K, int U, int MINB>
__global__ void __launch_bounds__(kAggThreads, MINB)
aggregate_seg_kernel(const __grid_constant__ SegArgs a, SolveParams sp, P2PArgs px) {
    constexpr int NACC = GramAcc<K>::N;
    constexpr int kTile = kAggThreads * U;                 // float4 items per tile and row
    const int tid = threadIdx.x;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(a.ws);
    unsigned int* done = reinterpret_cast<unsigned int*>(a.ws + 4);
    unsigned int* ready = reinterpret_cast<unsigned int*>(a.ws + 128);
    unsigned long long* stamps = reinterpret_cast<unsigned long long*>(a.ws + 192);
    double* partials = reinterpret_cast<double*>(a.ws + kGramHeaderBytes);
    const int n_seg = a.segs.n_segments;

    __shared__ unsigned int ready0;
    __shared__ unsigned long long t_start;
    __shared__ double red[kGramThreads / 32][NACC];
    __shared__ int is_last;
    __shared__ int tile_start[MOVAE_MAX_SEGMENTS + 1];
    if (tid == 0) {
        ready0 = ld_acquire_gpu_u32(ready);
        t_start = global_timer_ns();
        int acc = 0;
        for (int s = 0; s < n_seg; ++s) {
            tile_start[s] = acc;
            acc += (int)(((a.segs.n[s] >> 2) + kTile - 1) / kTile);
        }
        tile_start[n_seg] = acc;
    }
    __syncthreads();
    const int n_tiles = tile_start[n_seg];
    auto segment_of = [&](int v) { int s = 0; while (v >= tile_start[s + 1]) ++s; return s; };

    // ---- phase 1 ----
    double acc64[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc64[i] = 0.0;
    for (int v = blockIdx.x; v < n_tiles; v += gridDim.x) {
        const int s = segment_of(v);
        const int64_t n_items = a.segs.n[s] >> 2;
        const int64_t base = (int64_t)(v - tile_start[s]) * kTile + tid;
        float4 x[K][U];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const float4* row = reinterpret_cast<const float4*>(a.segs.rows[s][i]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t idx = base + u * kAggThreads;
                x[i][u] = idx < n_items ? ld_stream_f4(row + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        float acc[NACC];
#pragma unroll
        for (int q = 0; q < NACC; ++q) acc[q] = 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float c[K];
#pragma unroll
            for (int i = 0; i < K; ++i) c[i] = x[i][u].x;
            gram_fma<K>(acc, c);
#pragma unroll
            for (int i = 0; i < K; ++i) c[i] = x[i][u].y;
            gram_fma<K>(acc, c);
#pragma unroll
            for (int i = 0; i < K; ++i) c[i] = x[i][u].z;
            gram_fma<K>(acc, c);
#pragma unroll
            for (int i = 0; i < K; ++i) c[i] = x[i][u].w;
            gram_fma<K>(acc, c);
        }
#pragma unroll
        for (int q = 0; q < NACC; ++q) acc64[q] += (double)acc[q];
    }
    // ragged tails: columns 4 * (n / 4) .. n - 1 of every segment, one thread each in CTA 0
    if (blockIdx.x == 0 && tid < 4 * n_seg) {
        const int s = tid >> 2, e = tid & 3;
        const int64_t c0 = a.segs.n[s] & ~(int64_t)3;
        if (c0 + e < a.segs.n[s]) {
            float c[K];
            float acc[NACC];
#pragma unroll
            for (int q = 0; q < NACC; ++q) acc[q] = 0.f;
#pragma unroll
            for (int i = 0; i < K; ++i) c[i] = a.segs.rows[s][i][c0 + e];
            gram_fma<K>(acc, c);
#pragma unroll
            for (int q = 0; q < NACC; ++q) acc64[q] += (double)acc[q];
        }
    }
    const bool last = gram_cta_partial_and_ticket<K>(acc64, partials, ticket, red, &is_last);

    // ---- phase 2 ----
    if (last) aggregate_solve_phase<K>(sp, px, a.vec, a.aux, a.w_out, a.diag_out, a.G_out, partials, ticket, ready, ready0, stamps, t_start);
    if (a.out == nullptr) return;

    // ---- phase 3: the virtual tiles back to front ----
    float w[K];
    bool have_w = false;
    for (int v = n_tiles - 1 - (int)blockIdx.x; v >= 0; v -= gridDim.x) {
        const int s = segment_of(v);
        const int64_t n_items = a.segs.n[s] >> 2;
        const int64_t base = (int64_t)(v - tile_start[s]) * kTile + tid;
        float4 x[K][U];
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const float4* row = reinterpret_cast<const float4*>(a.segs.rows[s][i]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t idx = base + u * kAggThreads;
                x[i][u] = idx < n_items ? ld_stream_f4(row + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (!have_w) {
            if (!last) aggregate_wait_ready(ready, ready0);
#pragma unroll
            for (int i = 0; i < K; ++i) w[i] = __ldcg(a.w_out + i);
            have_w = true;
        }
        float4* dst = reinterpret_cast<float4*>(a.out + a.segs.out_off[s]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t idx = base + u * kAggThreads;
            if (idx >= n_items) continue;
            float4 o;
            o.x = w[0] * x[0][u].x; o.y = w[0] * x[0][u].y; o.z = w[0] * x[0][u].z; o.w = w[0] * x[0][u].w;
#pragma unroll
            for (int i = 1; i < K; ++i) {
                o.x = fmaf(w[i], x[i][u].x, o.x); o.y = fmaf(w[i], x[i][u].y, o.y);
                o.z = fmaf(w[i], x[i][u].z, o.z); o.w = fmaf(w[i], x[i][u].w, o.w);
            }
            if (a.accumulate) {
                const float4 old = dst[idx];
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            st_stream_f4(dst + idx, o);
        }
    }
    if (!have_w) {                                           // a CTA without tiles still has to join the protocol
        if (!last) aggregate_wait_ready(ready, ready0);
#pragma unroll
        for (int i = 0; i < K; ++i) w[i] = __ldcg(a.w_out + i);
    }
    if (blockIdx.x == 0 && tid < 4 * n_seg) {
        const int s = tid >> 2, e = tid & 3;
        const int64_t c = (a.segs.n[s] & ~(int64_t)3) + e;
        if (c < a.segs.n[s]) {
            float o = w[0] * a.segs.rows[s][0][c];
#pragma unroll
            for (int i = 1; i < K; ++i) o = fmaf(w[i], a.segs.rows[s][i][c], o);
            float* dst = a.out + a.segs.out_off[s] + c;
            *dst = a.accumulate ? *dst + o : o;
        }
    }
    __syncthreads();
    if (tid == 0) {
        if (atomicAdd(done, 1u) == gridDim.x - 1) {
            stamps[3] = global_timer_ns();
            *done = 0u;
        }
    }
}

template <int K, int U, int MINB>
static int launch_aggregate_seg(const SegArgs& a, const SolveParams& sp, const P2PArgs& px, cudaStream_t st) {
    auto kern = aggregate_seg_kernel<K, U, MINB>;
    static thread_local int occ_dev = -1, occ = 0;
    int dev = 0;
    MOVAE_CUDA_TRY(cudaGetDevice(&dev));
    if (occ_dev != dev) {
        MOVAE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kAggThreads, 0));
        if (occ < 1) occ = 1;
        occ_dev = dev;
    }
    const int sms = sm_count();
    MOVAE_REQUIRE(sms > 0, MOVAE_ERR_CUDA, "CUDA device query failed (no GPU?)");
    int64_t n_tiles = 0;
    for (int s = 0; s < a.segs.n_segments; ++s) n_tiles += ((a.segs.n[s] >> 2) + kAggThreads * U - 1) / (kAggThreads * U);
    if (n_tiles < 1) n_tiles = 1;
    int64_t grid = (int64_t)sms * occ;
    if (grid > n_tiles) grid = n_tiles;
    if (grid > kGramMaxBlocks) grid = kGramMaxBlocks;
    SegArgs a_ = a;
    SolveParams sp_ = sp;
    P2PArgs px_ = px;
    void* params[] = {&a_, &sp_, &px_};
    MOVAE_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3((unsigned)grid), dim3(kAggThreads), params, 0, st));
    return MOVAE_OK;
}

}  // namespace movae

extern "C" {

int movae_aggregate_f32(const float* d_J, int k, int64_t P, int64_t ldJ, const movae_solve_spec* spec, const float* d_vec,
                        const float* d_aux, float* d_grad, int accumulate, float* d_w, double* d_diag, double* d_G, void* d_ws,
                        size_t ws_bytes, const movae_p2p_ctx* ctx, void* stream) {
    using namespace movae;
    SolveParams sp;
    int rc = fill_solve_params(k, spec, &sp);
    if (rc != MOVAE_OK) return rc;
    rc = check_solve_vectors(sp, d_vec, d_aux);
    if (rc != MOVAE_OK) return rc;
    MOVAE_REQUIRE(P >= 0 && ldJ >= P, MOVAE_ERR_INVALID, "aggregate: need 0 <= P <= ldJ (P=%lld ldJ=%lld)", (long long)P, (long long)ldJ);
    MOVAE_REQUIRE(d_w != nullptr && (d_J != nullptr || P == 0), MOVAE_ERR_INVALID, "aggregate: null pointer");
    MOVAE_REQUIRE(d_ws != nullptr && ws_bytes >= movae_gram_workspace_bytes(k), MOVAE_ERR_WORKSPACE,
                  "aggregate: workspace too small (%zu < %zu)", ws_bytes, movae_gram_workspace_bytes(k));
    MOVAE_REQUIRE(reinterpret_cast<uintptr_t>(d_ws) % 8 == 0, MOVAE_ERR_WORKSPACE, "aggregate: workspace must be 8-byte aligned");
    P2PArgs px = p2p_disabled();
    if (ctx != nullptr) {
        rc = make_p2p_args(ctx, &px);
        if (rc != MOVAE_OK) return rc;
    }
    AggArgs a;
    a.J = d_J;
    a.P = P;
    a.ldJ = ldJ;
    a.ws = static_cast<unsigned char*>(d_ws);
    a.vec = d_vec;
    a.aux = d_aux;
    a.out = d_grad;
    a.accumulate = accumulate;
    a.w_out = d_w;
    a.diag_out = d_diag;
    a.G_out = d_G;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (k) {
        case 1: return dispatch_aggregate<1>(a, sp, px, st);
        case 2: return dispatch_aggregate<2>(a, sp, px, st);
        case 3: return dispatch_aggregate<3>(a, sp, px, st);
        case 4: return dispatch_aggregate<4>(a, sp, px, st);
        case 5: return dispatch_aggregate<5>(a, sp, px, st);
        case 6: return dispatch_aggregate<6>(a, sp, px, st);
        case 7: return dispatch_aggregate<7>(a, sp, px, st);
        default: return dispatch_aggregate<8>(a, sp, px, st);
    }
}

int movae_aggregate_segments_f32(const movae_jac_segments* segs, const movae_solve_spec* spec, const float* d_vec,
                                 const float* d_aux, float* d_grad, int accumulate, float* d_w, double* d_diag, double* d_G,
                                 void* d_ws, size_t ws_bytes, const movae_p2p_ctx* ctx, void* stream) {
    using namespace movae;
    MOVAE_REQUIRE(segs != nullptr, MOVAE_ERR_INVALID, "aggregate_segments: null segment table");
    const int k = segs->k;
    SolveParams sp;
    int rc = fill_solve_params(k, spec, &sp);
    if (rc != MOVAE_OK) return rc;
    rc = check_solve_vectors(sp, d_vec, d_aux);
    if (rc != MOVAE_OK) return rc;
    MOVAE_REQUIRE(segs->n_segments >= 1 && segs->n_segments <= MOVAE_MAX_SEGMENTS, MOVAE_ERR_UNSUPPORTED,
                  "aggregate_segments: %d segments outside 1..%d", segs->n_segments, MOVAE_MAX_SEGMENTS);
    MOVAE_REQUIRE(d_w != nullptr, MOVAE_ERR_INVALID, "aggregate_segments: null pointer");
    MOVAE_REQUIRE(d_grad == nullptr || reinterpret_cast<uintptr_t>(d_grad) % 16 == 0, MOVAE_ERR_INVALID,
                  "aggregate_segments: the gradient buffer must be 16-byte aligned");
    for (int s = 0; s < segs->n_segments; ++s) {
        MOVAE_REQUIRE(segs->n[s] >= 0 && segs->out_off[s] >= 0 && segs->out_off[s] % 4 == 0, MOVAE_ERR_INVALID,
                      "aggregate_segments: segment %d needs n >= 0 and an output offset that is a multiple of 4", s);
        for (int i = 0; i < k; ++i)
            MOVAE_REQUIRE(segs->rows[s][i] != nullptr && reinterpret_cast<uintptr_t>(segs->rows[s][i]) % 16 == 0, MOVAE_ERR_INVALID,
                          "aggregate_segments: row %d of segment %d must be a 16-byte aligned device pointer", i, s);
    }
    MOVAE_REQUIRE(d_ws != nullptr && ws_bytes >= movae_gram_workspace_bytes(k), MOVAE_ERR_WORKSPACE,
                  "aggregate_segments: workspace too small (%zu < %zu)", ws_bytes, movae_gram_workspace_bytes(k));
    P2PArgs px = p2p_disabled();
    if (ctx != nullptr) {
        rc = make_p2p_args(ctx, &px);
        if (rc != MOVAE_OK) return rc;
    }
    SegArgs a;
    a.segs = *segs;
    a.ws = static_cast<unsigned char*>(d_ws);
    a.vec = d_vec;
    a.aux = d_aux;
    a.out = d_grad;
    a.accumulate = accumulate;
    a.w_out = d_w;
    a.diag_out = d_diag;
    a.G_out = d_G;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (k) {
        case 1: return launch_aggregate_seg<1, 4, 2>(a, sp, px, st);
        case 2: return launch_aggregate_seg<2, 4, 2>(a, sp, px, st);
        case 3: return launch_aggregate_seg<3, 4, 2>(a, sp, px, st);
        case 4: return launch_aggregate_seg<4, 2, 2>(a, sp, px, st);
        case 5: return launch_aggregate_seg<5, 2, 1>(a, sp, px, st);
        case 6: return launch_aggregate_seg<6, 2, 1>(a, sp, px, st);
        case 7: return launch_aggregate_seg<7, 2, 1>(a, sp, px, st);
        default: return launch_aggregate_seg<8, 2, 1>(a, sp, px, st);
    }
}

int movae_aggregate_timestamps(const void* d_ws, uint64_t h_stamps[6], void* stream) {
    using namespace movae;
    MOVAE_REQUIRE(d_ws != nullptr && h_stamps != nullptr, MOVAE_ERR_INVALID, "aggregate_timestamps: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MOVAE_CUDA_TRY(cudaMemcpyAsync(h_stamps, static_cast<const unsigned char*>(d_ws) + 192, 6 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    MOVAE_CUDA_TRY(cudaStreamSynchronize(st));
    return MOVAE_OK;
}

}  // extern "C"
