// K7 `optim_step`: the optimizer step that follows the aggregation, on the FLAT
// parameter / gradient / moment buffers the aggregation path already owns.
//
// Replaces, for the four optimizers the reference can construct (/root/reference/main.py:1169-1176:
// optim.SGD / Adam / AdamW / RMSprop with lr, weight_decay and -- SGD only -- momentum), the per-tensor
// `optimizer.step()` at main.py:214 and the `clip_grad_norm_` in front of it (main.py:211-212):
//
//   * ONE streaming kernel over the flat buffers instead of ~10 multi-tensor launches over 12-60 tensors;
//     algorithmic traffic per parameter: Adam/AdamW 16 B read + 12 B written, SGD+momentum 12 + 8,
//     RMSprop 12 + 8, plain SGD 8 + 4.  Roofline: HBM.
//   * gradient clipping costs no pass over the gradients: the squared global norm arrives as a device
//     double (K1 with k = 1 over the flat gradient buffer) and the clip coefficient
//     min(1, max_norm / (norm + 1e-6)) is applied on the fly (torch multiplies .grad in place first);
//   * the step count lives on the device and is advanced by the kernel itself (last CTA, ticket), and the
//     learning rate may be read from device memory, so the kernel is CUDA-graph capturable with no host
//     state baked in.
//
// Per-element arithmetic follows torch.optim's single-tensor formulas in float32 (torch/optim/adam.py
// `_single_tensor_adam`, sgd.py, rmsprop.py; defaults amsgrad=False, maximize=False, dampening=0,
// nesterov=False, centered=False, RMSprop momentum=0) with the bias corrections evaluated in float64.
#include <math.h>

#include "common.cuh"

namespace movae {

constexpr int kOptThreads = 256;
constexpr int kOptU = 4;             // float4 items per thread and array: up to 16 x 16 B loads in flight per thread

struct OptimState {
    long long step;            // completed steps
    unsigned int ticket;       // CTAs finished in the running launch
    unsigned int pad;
};

struct OptimCoefs {
    float clip;                // gradient scale (1 when clipping is off)
    float lr, wd;
    float b1, b2, eps;
    float one_minus_b1, one_minus_b2;
    float step_size;           // Adam: lr / (1 - b1^t)
    float bc2_sqrt;            // Adam: sqrt(1 - b2^t)
    float decay;               // AdamW: 1 - lr * wd
};

__device__ __forceinline__ OptimCoefs make_coefs(const movae_optim_spec& s, const double* d_lr, const double* d_gnorm_sq,
                                                 const OptimState* st) {
    OptimCoefs c;
    const double lr = d_lr ? *d_lr : s.lr;
    c.lr = (float)lr;
    c.wd = (float)s.weight_decay;
    c.b1 = (float)s.beta1;
    c.b2 = (float)s.beta2;
    c.eps = (float)s.eps;
    c.one_minus_b1 = (float)(1.0 - s.beta1);        // formed in float64 like torch's Python floats, then rounded
    c.one_minus_b2 = (float)(1.0 - s.beta2);
    c.clip = 1.f;
    if (d_gnorm_sq != nullptr && s.max_grad_norm > 0.0) {
        const float total = (float)sqrt(*d_gnorm_sq);
        const float coef = (float)s.max_grad_norm / (total + 1e-6f);
        c.clip = coef < 1.f ? coef : 1.f;
        if (coef != coef) c.clip = coef;                        // NaN norm -> NaN coefficient, as in torch
    }
    const double t = (double)(st->step + 1);
    const double bc1 = 1.0 - pow(s.beta1, t);
    const double bc2 = 1.0 - pow(s.beta2, t);
    c.step_size = (float)(lr / bc1);
    c.bc2_sqrt = (float)sqrt(bc2);
    c.decay = (float)(1.0 - lr * s.weight_decay);
    return c;
}

template <int KIND, bool HAS_M>
__device__ __forceinline__ void update_one(float& p, float g, float& m, float& v, const OptimCoefs& c) {
    g *= c.clip;
    if constexpr (KIND == MOVAE_OPT_ADAM || KIND == MOVAE_OPT_ADAMW) {
        if constexpr (KIND == MOVAE_OPT_ADAMW) {
            p *= c.decay;                                       // param.mul_(1 - lr * weight_decay)
        } else {
            if (c.wd != 0.f) g = fmaf(c.wd, p, g);              // grad.add(param, alpha=weight_decay)
        }
        m = fmaf(g - m, c.one_minus_b1, m);                     // exp_avg.lerp_(grad, 1 - beta1)
        v = fmaf(c.one_minus_b2 * g, g, v * c.b2);              // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = sqrtf(v) / c.bc2_sqrt + c.eps;
        p = fmaf(-c.step_size, m / denom, p);                   // param.addcdiv_(exp_avg, denom, value=-step_size)
    } else if constexpr (KIND == MOVAE_OPT_SGD) {
        if (c.wd != 0.f) g = fmaf(c.wd, p, g);
        if constexpr (HAS_M) {
            m = fmaf(m, c.b1, g);                               // buf.mul_(momentum).add_(grad); first step: buf = grad
            g = m;
        }
        p = fmaf(-c.lr, g, p);
    } else {                                                    // RMSprop
        if (c.wd != 0.f) g = fmaf(c.wd, p, g);
        v = fmaf(c.one_minus_b2 * g, g, v * c.b2);              // square_avg.mul_(alpha).addcmul_(grad, grad, 1 - alpha)
        const float avg = sqrtf(v) + c.eps;
        p = fmaf(-c.lr, g / avg, p);
    }
}

__device__ __forceinline__ float4 ld_na_f4(const float4* p) {     // read-write buffers: plain (coherent) load, no L1 allocation
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <int KIND, bool HAS_M, bool HAS_V, bool VEC>
__global__ void __launch_bounds__(kOptThreads)
optim_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                  movae_optim_spec spec, const double* __restrict__ d_lr, const double* __restrict__ d_gnorm_sq,
                  OptimState* __restrict__ state) {
    const OptimCoefs c = make_coefs(spec, d_lr, d_gnorm_sq, state);
    const int tid = threadIdx.x;
    if constexpr (VEC) {
        const int64_t n4 = n / 4;
        const int64_t tile = (int64_t)kOptThreads * kOptU;
        for (int64_t base = (int64_t)blockIdx.x * tile; base < n4; base += (int64_t)gridDim.x * tile) {
            float4 pv[kOptU], gv[kOptU], mv[kOptU], vv[kOptU];
#pragma unroll
            for (int u = 0; u < kOptU; ++u) {
                const int64_t i = base + u * kOptThreads + tid;
                if (i < n4) {
                    gv[u] = ld_stream_f4(reinterpret_cast<const float4*>(g) + i);
                    pv[u] = ld_na_f4(reinterpret_cast<const float4*>(p) + i);
                    if constexpr (HAS_M) mv[u] = ld_na_f4(reinterpret_cast<const float4*>(m) + i);
                    if constexpr (HAS_V) vv[u] = ld_na_f4(reinterpret_cast<const float4*>(v) + i);
                }
            }
#pragma unroll
            for (int u = 0; u < kOptU; ++u) {
                const int64_t i = base + u * kOptThreads + tid;
                if (i >= n4) continue;
                float dm = 0.f, dv = 0.f;
                update_one<KIND, HAS_M>(pv[u].x, gv[u].x, HAS_M ? mv[u].x : dm, HAS_V ? vv[u].x : dv, c);
                update_one<KIND, HAS_M>(pv[u].y, gv[u].y, HAS_M ? mv[u].y : dm, HAS_V ? vv[u].y : dv, c);
                update_one<KIND, HAS_M>(pv[u].z, gv[u].z, HAS_M ? mv[u].z : dm, HAS_V ? vv[u].z : dv, c);
                update_one<KIND, HAS_M>(pv[u].w, gv[u].w, HAS_M ? mv[u].w : dm, HAS_V ? vv[u].w : dv, c);
                reinterpret_cast<float4*>(p)[i] = pv[u];
                if constexpr (HAS_M) st_stream_f4(reinterpret_cast<float4*>(m) + i, mv[u]);
                if constexpr (HAS_V) st_stream_f4(reinterpret_cast<float4*>(v) + i, vv[u]);
            }
        }
        if (blockIdx.x == 0 && tid < (int)(n - n4 * 4)) {         // ragged tail
            const int64_t i = n4 * 4 + tid;
            float pp = p[i], mm = HAS_M ? m[i] : 0.f, vv1 = HAS_V ? v[i] : 0.f;
            update_one<KIND, HAS_M>(pp, g[i], mm, vv1, c);
            p[i] = pp;
            if constexpr (HAS_M) m[i] = mm;
            if constexpr (HAS_V) v[i] = vv1;
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * kOptThreads + tid; i < n; i += (int64_t)gridDim.x * kOptThreads) {
            float pp = p[i], mm = HAS_M ? m[i] : 0.f, vv1 = HAS_V ? v[i] : 0.f;
            update_one<KIND, HAS_M>(pp, g[i], mm, vv1, c);
            p[i] = pp;
            if constexpr (HAS_M) m[i] = mm;
            if constexpr (HAS_V) v[i] = vv1;
        }
    }
    // every CTA read state->step before it got here; the last one to finish advances it
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned int done = atomicAdd(&state->ticket, 1u);
        if (done == gridDim.x - 1) {
            if (!spec.hold_step) state->step += 1;
            state->ticket = 0;
            __threadfence();
        }
    }
}

template <int KIND, bool HAS_M, bool HAS_V>
static int launch_optim(float* p, const float* g, float* m, float* v, int64_t n, const movae_optim_spec& spec, const double* d_lr,
                        const double* d_gnorm_sq, OptimState* state, cudaStream_t st) {
    const int sms = sm_count();
    MOVAE_REQUIRE(sms > 0, MOVAE_ERR_CUDA, "CUDA device query failed (no GPU?)");
    auto aligned = [](const void* q) { return q == nullptr || reinterpret_cast<uintptr_t>(q) % 16 == 0; };
    const bool vec = aligned(p) && aligned(g) && aligned(m) && aligned(v);
    // persistent grid: exactly the CTAs that are resident at once (a second partial wave costs ~10% here)
    static thread_local int occ_vec = 0, occ_scalar = 0;
    int& occ = vec ? occ_vec : occ_scalar;
    if (occ == 0) {
        if (vec)
            MOVAE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, optim_step_kernel<KIND, HAS_M, HAS_V, true>, kOptThreads, 0));
        else
            MOVAE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, optim_step_kernel<KIND, HAS_M, HAS_V, false>, kOptThreads, 0));
        if (occ < 1) occ = 1;
    }
    const int64_t items = vec ? (n / 4 + (int64_t)kOptThreads * kOptU - 1) / ((int64_t)kOptThreads * kOptU)
                              : (n + kOptThreads - 1) / kOptThreads;
    int64_t grid = (int64_t)sms * occ;
    if (grid > items) grid = items;
    if (grid < 1) grid = 1;
    if (vec)
        optim_step_kernel<KIND, HAS_M, HAS_V, true><<<(unsigned)grid, kOptThreads, 0, st>>>(p, g, m, v, n, spec, d_lr, d_gnorm_sq, state);
    else
        optim_step_kernel<KIND, HAS_M, HAS_V, false><<<(unsigned)grid, kOptThreads, 0, st>>>(p, g, m, v, n, spec, d_lr, d_gnorm_sq, state);
    MOVAE_CUDA_TRY(cudaGetLastError());
    return MOVAE_OK;
}

}  // namespace movae

extern "C" {

size_t movae_optim_state_bytes(void) { return sizeof(movae::OptimState); }

int movae_optim_step_f32(float* d_p, const float* d_g, float* d_m, float* d_v, int64_t n, const movae_optim_spec* spec,
                         const double* d_lr, const double* d_gnorm_sq, void* d_state, void* stream) {
    using namespace movae;
    MOVAE_REQUIRE(spec != nullptr, MOVAE_ERR_INVALID, "optim_step: null spec");
    MOVAE_REQUIRE(n >= 0, MOVAE_ERR_INVALID, "optim_step: n must be >= 0");
    if (n == 0) return MOVAE_OK;
    MOVAE_REQUIRE(d_p && d_g && d_state, MOVAE_ERR_INVALID, "optim_step: null pointer");
    MOVAE_REQUIRE(reinterpret_cast<uintptr_t>(d_state) % 8 == 0, MOVAE_ERR_INVALID, "optim_step: state must be 8-byte aligned");
    MOVAE_REQUIRE(spec->lr >= 0.0 || d_lr != nullptr, MOVAE_ERR_INVALID, "optim_step: negative learning rate");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    OptimState* state = static_cast<OptimState*>(d_state);
    switch (spec->kind) {
        case MOVAE_OPT_ADAM:
            MOVAE_REQUIRE(d_m && d_v, MOVAE_ERR_INVALID, "optim_step: Adam needs both moment buffers");
            return launch_optim<MOVAE_OPT_ADAM, true, true>(d_p, d_g, d_m, d_v, n, *spec, d_lr, d_gnorm_sq, state, st);
        case MOVAE_OPT_ADAMW:
            MOVAE_REQUIRE(d_m && d_v, MOVAE_ERR_INVALID, "optim_step: AdamW needs both moment buffers");
            return launch_optim<MOVAE_OPT_ADAMW, true, true>(d_p, d_g, d_m, d_v, n, *spec, d_lr, d_gnorm_sq, state, st);
        case MOVAE_OPT_SGD:
            if (spec->beta1 != 0.0) {
                MOVAE_REQUIRE(d_m, MOVAE_ERR_INVALID, "optim_step: SGD with momentum needs the momentum buffer");
                return launch_optim<MOVAE_OPT_SGD, true, false>(d_p, d_g, d_m, nullptr, n, *spec, d_lr, d_gnorm_sq, state, st);
            }
            return launch_optim<MOVAE_OPT_SGD, false, false>(d_p, d_g, nullptr, nullptr, n, *spec, d_lr, d_gnorm_sq, state, st);
        case MOVAE_OPT_RMSPROP:
            MOVAE_REQUIRE(d_v, MOVAE_ERR_INVALID, "optim_step: RMSprop needs the square-average buffer");
            return launch_optim<MOVAE_OPT_RMSPROP, false, true>(d_p, d_g, nullptr, d_v, n, *spec, d_lr, d_gnorm_sq, state, st);
        default:
            MOVAE_REQUIRE(false, MOVAE_ERR_INVALID, "optim_step: unknown optimizer kind %d", spec->kind);
    }
    return MOVAE_OK;
}

}  // extern "C"
