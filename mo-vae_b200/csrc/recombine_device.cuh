// Device-side recombination pass  out (=|+=) w @ J, shared by K3 (`recombine_kernel`, recombine.cu) and
// phase 2 of the fused aggregation kernel (aggregate.cu).
//
// Every CTA walks the span of columns it owned in the Gramian pass from its END backwards: that pass streamed the
// spans front-to-back, so the last ~100 MB it touched (the tails of all spans) are still L2-resident on a 126 MB L2
// and are consumed first.  The loads of the first tile
// are issued BEFORE `ready()` is called and the weights are read after it: in the fused kernel `ready`
// waits for the solve, so the wait overlaps the first loads.
#pragma once
#include <type_traits>

#include "gram_device.cuh"

namespace movae {

constexpr int kRecThreads = 256;

template <int K, int U, bool VEC>
struct RecTile {
    using T = typename std::conditional<VEC, float4, float>::type;
    T v[K][U];
};

template <int K, int U, bool VEC>
__device__ __forceinline__ void rec_load_tile(RecTile<K, U, VEC>& r, const float* __restrict__ J, int64_t ldJ, int64_t base,
                                              int64_t lo, int64_t hi, bool full) {
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t idx = base + u * kRecThreads;
            if constexpr (VEC)
                r.v[i][u] = (full || (idx >= lo && idx < hi)) ? ld_stream_f4(reinterpret_cast<const float4*>(J + i * ldJ) + idx)
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
            else
                r.v[i][u] = (full || (idx >= lo && idx < hi)) ? ld_stream_f1(J + i * ldJ + idx) : 0.f;
        }
}

template <int K, int U, bool VEC>
__device__ __forceinline__ void rec_store_tile(const RecTile<K, U, VEC>& r, const float (&w)[K], float* __restrict__ out,
                                               int64_t base, int64_t lo, int64_t hi, bool full, int accumulate) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int64_t idx = base + u * kRecThreads;
        if (!full && (idx < lo || idx >= hi)) continue;
        if constexpr (VEC) {
            float4 o;
            o.x = w[0] * r.v[0][u].x; o.y = w[0] * r.v[0][u].y; o.z = w[0] * r.v[0][u].z; o.w = w[0] * r.v[0][u].w;
#pragma unroll
            for (int i = 1; i < K; ++i) {
                o.x = fmaf(w[i], r.v[i][u].x, o.x); o.y = fmaf(w[i], r.v[i][u].y, o.y);
                o.z = fmaf(w[i], r.v[i][u].z, o.z); o.w = fmaf(w[i], r.v[i][u].w, o.w);
            }
            float4* dst = reinterpret_cast<float4*>(out) + idx;
            if (accumulate) {
                const float4 old = *dst;
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            st_stream_f4(dst, o);
        } else {
            float o = w[0] * r.v[0][u];
#pragma unroll
            for (int i = 1; i < K; ++i) o = fmaf(w[i], r.v[i][u], o);
            if (accumulate) o += out[idx];
            out[idx] = o;
        }
    }
}

template <int K, int U, bool VEC, class Ready>
__device__ __forceinline__ void recombine_tiles(const float* __restrict__ J, int64_t P, int64_t ldJ, const float* w_dev,
                                                float* __restrict__ out, int accumulate, Ready ready) {
    constexpr int W = VEC ? 4 : 1;
    const int tid = threadIdx.x;
    const int64_t n_items = P / W;
    constexpr int64_t tile_items = (int64_t)kRecThreads * U;
    int64_t lo, hi;
    cta_span(n_items, lo, hi);              // the same span the Gramian pass gave this CTA

    // tiles [t0, t0 + tile_items) with t0 = hi - tile_items, hi - 2 tile_items, ...; the last one is clipped at lo
    RecTile<K, U, VEC> r;
    int64_t t0 = hi - tile_items;
    if (hi > lo) rec_load_tile<K, U, VEC>(r, J, ldJ, t0 + tid, lo, hi, t0 >= lo);
    ready();
    float w[K];
#pragma unroll
    for (int i = 0; i < K; ++i) w[i] = __ldcg(w_dev + i);
    while (t0 + tile_items > lo) {
        rec_store_tile<K, U, VEC>(r, w, out, t0 + tid, lo, hi, t0 >= lo, accumulate);
        t0 -= tile_items;
        if (t0 + tile_items > lo) rec_load_tile<K, U, VEC>(r, J, ldJ, t0 + tid, lo, hi, t0 >= lo);
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

}  // namespace movae
