// K4, CTA-pair version: the same search as vq_argmin_tc.cu issued as `tcgen05.mma.cta_group::2`.
//
// Two CTAs of a cluster (two SMs of one TPC) work on two consecutive 128-row tiles as ONE M = 256
// MMA: each CTA stages its own tile (A operand, its own TMEM accumulators, its own epilogue) but only
// HALF of every B operand -- for a 128-code unit, CTA r holds codes 64r .. 64r+63 of the unit and the
// tensor cores of both SMs read both halves.  Per SM and 64-cycle MMA that is 4 KB of A + 2 KB of B
// = 96 B/clk of shared-memory reads, so N = 128 units are affordable (one CTA alone needs 128 B/clk for
// them, vq_argmin_tc.cu) and TMEM holds FOUR 128-column accumulators: each epilogue group drains one
// of its two units while the tensor cores fill the other (issue order 0, 2, 1, 3), which removes the
// "drain time ~ MMA time" coupling that kept the one-CTA kernel at 73% tensor-pipe activity.  The
// codebook costs 72 KB of shared memory per CTA instead of 144 KB.
//
// Protocol (leader = cluster rank 0 issues every MMA and owns the barriers the issuer waits on):
//   a_full[s]   leader's, 256 arrivals: the 128 producers of both CTAs (rank 1 arrives remotely)
//   a_empty[s]  one per CTA, signalled by a multicast tcgen05.commit          (stage s may be refilled)
//   acc_full[q] one per CTA, multicast commit                               (unit q may be drained)
//   acc_empty[q] leader's, 256 arrivals: epilogue group q/2 of both CTAs    (unit q may be overwritten)
// Everything else (key construction by the tensor core, top-2 epilogue, worklist) is as in the
// one-CTA kernel; see its header for the arithmetic.
#include "vq_tc_common.cuh"

namespace movae {

namespace {

constexpr int kD = 64, kK = 512, kTileM = 128, kUnitN = 128, kHalfN = 256;
constexpr int kThreads = 13 * 32;
constexpr int kStages = 3;

// shared-memory carve-up per CTA (bytes from a 1024-aligned base)
constexpr uint32_t kOffBhi = 0;                          // 4 units x 64 rows x 128 B (this CTA's half of each unit)
constexpr uint32_t kOffBlo = 32768;
constexpr uint32_t kOffA = 65536;                        // kStages x (hi 16 KB + lo 16 KB)
constexpr uint32_t kOffBaug = kOffA + kStages * 32768;   // 4 units x 8 code groups x 256 B
constexpr uint32_t kOffAaug = kOffBaug + 8192;           // kStages x 3 steps x 256 B
constexpr uint32_t kOffXchg = kOffAaug + kStages * 768;  // 2 slots x 128 x {best, second, code}
constexpr uint32_t kOffBar = ((kOffXchg + 2 * 128 * 12 + 15) / 16) * 16;
constexpr uint32_t kSmemBytes = kOffBar + 512 + 1024;    // + alignment slack

struct Bars {
    uint64_t a_full[kStages], a_empty[kStages], acc_full[4], acc_empty[4], x_full[2], x_free[2];
    uint32_t tmem_base;
    uint32_t emax2_bits;
    float pmax[kStages][4];
};

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.  Default (CTA-scope) semantics on
// purpose: with .release.cluster / .acquire.cluster ptxas emits MEMBAR + CCTL.IVALL (an L1 invalidate) around every
// barrier operation (18% of the stall samples of the first version); what crosses CTAs here is shared memory read by the
// tensor core, ordered by fence.proxy.async + the mbarrier itself.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(tc::smem_u32(bar)), "r"(rank)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(tc::smem_u32(bar)), "r"(parity)
            : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mma2_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (when all MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void mma2_commit_both(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     tc::smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

}  // namespace

template <bool DBG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
vq_argmin_tc2_kernel(const float* __restrict__ z, int64_t N, int64_t HW, const float* __restrict__ E,
                     long long* __restrict__ idx_out, int* __restrict__ list, unsigned int* __restrict__ list_count,
                     float* __restrict__ dbg) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = tc::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    Bars* bars = reinterpret_cast<Bars*>(smem + kOffBar);
    float* xchg = reinterpret_cast<float*>(smem + kOffXchg);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_rank();
    const int64_t n_tiles = (N + kTileM - 1) / kTileM;
    const int64_t n_pairs = (n_tiles + 1) / 2;
    const int64_t pair0 = blockIdx.x >> 1, pair_stride = gridDim.x >> 1;

    // ---- one-time setup ---------------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            tc::mbar_init(&bars->a_full[i], 256);
            tc::mbar_init(&bars->a_empty[i], 1);
        }
        for (int i = 0; i < 4; ++i) {
            tc::mbar_init(&bars->acc_full[i], 1);
            tc::mbar_init(&bars->acc_empty[i], 256);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bars->x_full[i], 128);
            tc::mbar_init(&bars->x_free[i], 128);
        }
        bars->emax2_bits = 0u;
        tc::mbar_fence_init();
    }
    if (warp == 12) tmem_alloc2(&bars->tmem_base, 512);
    __syncthreads();

    // this CTA's half of the codebook: unit n (codes 128n .. 128n+127), local row lr = 0..63 <-> code 128n + 64 rank + lr
    // B = -2 E split into bf16 hi + lo, K-major rows of 128 B, 128B swizzle; unit n at byte offset n * 8192
    for (int t = tid; t < 256 * 8; t += kThreads) {
        const int row = t >> 3, c = t & 7;                       // row = n * 64 + lr
        const int j = (row >> 6) * 128 + (int)rank * 64 + (row & 63);
        const float4 x0 = __ldg(reinterpret_cast<const float4*>(E + (size_t)j * kD + c * 8));
        const float4 x1 = __ldg(reinterpret_cast<const float4*>(E + (size_t)j * kD + c * 8 + 4));
        const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) split_bf16x2(-2.f * x[2 * p], -2.f * x[2 * p + 1], hi[p], lo[p]);
        const uint32_t off = (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(smem + kOffBhi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(smem + kOffBlo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    // side table rows for this CTA's codes (no swizzle; group g = row / 8 at g * 256, k >= 8 half at +128 all zero);
    // max |e|^2 is taken over ALL codes (both CTAs need the same bound)
    for (int j = tid; j < kK; j += kThreads) {
        double s = 0.0;
        for (int d = 0; d < kD; d += 4) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(E + (size_t)j * kD + d));
            s += (double)x.x * x.x + (double)x.y * x.y + (double)x.z * x.z + (double)x.w * x.w;
        }
        const float sf = (float)s;
        atomicMax(&bars->emax2_bits, __float_as_uint(sf));
        if (((j >> 6) & 1) == (int)rank) {
            const int row = (j >> 7) * 64 + (j & 63);
            const uint32_t h = bf16_bits_rn(sf);
            const float r1 = sf - __uint_as_float(h << 16);
            const uint32_t m = bf16_bits_rn(r1);
            const uint32_t l = bf16_bits_rn(r1 - __uint_as_float(m << 16));
            const uint32_t one = 0x3F80u;
            const uint32_t col = bf16_bits_rn((float)((j & 1) | (((j & 31) >> 3) << 1)));
            uint8_t* p = smem + kOffBaug + (uint32_t)(row >> 3) * 256u + (uint32_t)(row & 7) * 16u;
            *reinterpret_cast<uint4*>(p) = make_uint4(h | (m << 16), l | (one << 16), one | (one << 16), col);
            *reinterpret_cast<uint4*>(p + 128) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    for (int t = tid; t < kStages * 768 / 16; t += kThreads) *reinterpret_cast<uint4*>(smem + kOffAaug + t * 16) = make_uint4(0u, 0u, 0u, 0u);
    tc::fence_proxy_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                                          // both CTAs: barriers initialised, operands staged, TMEM allocated
    tc::tc_fence_after_sync();
    const uint32_t tmem_base = bars->tmem_base;
    const float emax2 = __uint_as_float(bars->emax2_bits) * 1.000001f;
    const float emax = sqrtf(emax2) * 1.000001f;

    if (warp >= 8 && warp < 12) {
        // ===== producers: one row of this CTA's tile per thread ====================================
        const int r = tid - 256;
        float v[kD];
        const float* p_next = nullptr;
        {
            const int64_t n = (2 * pair0 + rank) * kTileM + r;
            if (pair0 < n_pairs && n < N) {
                const int64_t b = n / HW, hw = n - b * HW;
                p_next = z + (b * kD) * HW + hw;
            }
#pragma unroll
            for (int d = 0; d < kD; ++d) v[d] = p_next ? ld_stream_f1(p_next + (int64_t)d * HW) : 0.f;
        }
        uint32_t it = 0;
        for (int64_t pair = pair0; pair < n_pairs; pair += pair_stride, ++it) {
            const uint32_t s = it % kStages, fill = it / kStages;
            const int64_t pair2 = pair + pair_stride;
            const int64_t n2 = (2 * pair2 + rank) * kTileM + r;
            p_next = nullptr;
            if (pair2 < n_pairs && n2 < N) {
                const int64_t b = n2 / HW, hw = n2 - b * HW;
                p_next = z + (b * kD) * HW + hw;
            }
            tc::mbar_wait(&bars->a_empty[s], (fill & 1u) ^ 1u);
            uint8_t* ahi = smem + kOffA + s * 32768u + (uint32_t)r * 128u;
            uint8_t* alo = ahi + 16384;
            float z2 = 0.f;
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
#pragma unroll
                for (int c = 2 * qd; c < 2 * qd + 2; ++c) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float a = v[c * 8 + 2 * p], b = v[c * 8 + 2 * p + 1];
                        z2 = fmaf(a, a, z2);
                        z2 = fmaf(b, b, z2);
                        split_bf16x2(a, b, hi[p], lo[p]);
                    }
                    const uint32_t off = (uint32_t)((c ^ (r & 7)) << 4);
                    *reinterpret_cast<uint4*>(ahi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(alo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
#pragma unroll
                for (int d = 16 * qd; d < 16 * qd + 16; ++d) v[d] = p_next ? ld_stream_f1(p_next + (int64_t)d * HW) : 0.f;
            }
            // the tile's key constants (identical for its 128 rows): see vq_argmin_tc.cu
            float zm = z2;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) zm = fmaxf(zm, __shfl_xor_sync(0xffffffffu, zm, o));
            if (lane == 0) bars->pmax[s][warp - 8] = zm;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 8 && lane < 24) {
                const float z2max = fmaxf(fmaxf(bars->pmax[s][0], bars->pmax[s][1]), fmaxf(bars->pmax[s][2], bars->pmax[s][3]));
                const float Craw = 2.001f * sqrtf(z2max * 1.00001f) * emax + 1e-30f;
                const uint32_t Cb = (__float_as_uint(Craw) + 0xFFFFu) >> 16;
                const float C = __uint_as_float(Cb << 16);
                const float Rg = fmaxf(2.f * C + emax2, 1e-30f) * 1.001f;
                uint32_t mexp = (__float_as_uint(Rg) >> 23) + 1u;
                mexp = mexp < 40u ? 40u : mexp;
                const uint32_t bigM = (mexp + 3u) << 7;                      // 8 M
                const uint32_t m7 = 0x8000u | ((mexp + 2u) << 7) | 0x60u;    // -7 M
                const uint32_t uu = (mexp - 23u) << 7;                       // u = M 2^-23
                const int step = lane >> 3, row = lane & 7;
                uint4 val;
                if (step == 0) val = make_uint4(0x3F80u | (0x3F80u << 16), 0x3F80u | (Cb << 16), bigM, 0u);
                else if (step == 1) val = make_uint4(0u, 0u, m7 << 16, 0u);
                else val = make_uint4(0u, 0u, 0u, uu);
                *reinterpret_cast<uint4*>(smem + kOffAaug + s * 768u + (uint32_t)step * 256u + (uint32_t)row * 16u) = val;
            }
            tc::fence_proxy_async_smem();
            mbar_arrive_cluster(&bars->a_full[s], 0u);                   // the leader's barrier
        }
    } else if (warp == 12) {
        // ===== MMA issuer: leader CTA only =========================================================
        if (rank == 0) {
            constexpr uint32_t idesc = tc::idesc_bf16_f32(2 * kTileM, kUnitN);     // M = 256 across the pair
            const uint32_t smem_base = tc::smem_u32(smem);
            uint32_t it = 0;
            for (int64_t pair = pair0; pair < n_pairs; pair += pair_stride, ++it) {
                const uint32_t s = it % kStages, fill = it / kStages;
                mbar_wait_cluster(&bars->a_full[s], fill & 1u);
                tc::tc_fence_after_sync();
                // Two units (one of each epilogue group) are issued INTERLEAVED: consecutive MMAs then accumulate into
                // different TMEM columns.  Back-to-back MMAs into the same accumulator expose ~50 cycles of pipeline
                // latency per instruction (measured: 118 cycles per N=128 MMA, 175 per N=256 MMA instead of 64 / 128).
#pragma unroll 1
                for (uint32_t h = 0; h < 2; ++h) {                          // units (0, 2) then (1, 3)
                    const uint32_t qa = h, qb = h + 2;
                    mbar_wait_cluster(&bars->acc_empty[qa], (it & 1u) ^ 1u);
                    mbar_wait_cluster(&bars->acc_empty[qb], (it & 1u) ^ 1u);
                    tc::tc_fence_after_sync();
                    if (lane == 0) {
                        const uint32_t da = tmem_base + qa * kUnitN, db = tmem_base + qb * kUnitN;
                        const uint64_t a_hi = tc::smem_desc_kmajor_sw128(smem_base + kOffA + s * 32768u);
                        const uint64_t a_lo = tc::smem_desc_kmajor_sw128(smem_base + kOffA + s * 32768u + 16384u);
                        const uint64_t bha = tc::smem_desc_kmajor_sw128(smem_base + kOffBhi + qa * 8192u);
                        const uint64_t bla = tc::smem_desc_kmajor_sw128(smem_base + kOffBlo + qa * 8192u);
                        const uint64_t bhb = tc::smem_desc_kmajor_sw128(smem_base + kOffBhi + qb * 8192u);
                        const uint64_t blb = tc::smem_desc_kmajor_sw128(smem_base + kOffBlo + qb * 8192u);
#pragma unroll
                        for (uint32_t k = 0; k < 4; ++k) {
                            mma2_bf16_ss(da, a_lo + 2 * k, bha + 2 * k, idesc, k > 0);
                            mma2_bf16_ss(db, a_lo + 2 * k, bhb + 2 * k, idesc, k > 0);
                        }
#pragma unroll
                        for (uint32_t k = 0; k < 4; ++k) {
                            mma2_bf16_ss(da, a_hi + 2 * k, bla + 2 * k, idesc, 1u);
                            mma2_bf16_ss(db, a_hi + 2 * k, blb + 2 * k, idesc, 1u);
                        }
#pragma unroll
                        for (uint32_t k = 0; k < 4; ++k) {
                            mma2_bf16_ss(da, a_hi + 2 * k, bha + 2 * k, idesc, 1u);
                            mma2_bf16_ss(db, a_hi + 2 * k, bhb + 2 * k, idesc, 1u);
                        }
                        const uint64_t bga = tc::smem_desc_kmajor_noswizzle(smem_base + kOffBaug + qa * 2048u, 128u, 256u);
                        const uint64_t bgb = tc::smem_desc_kmajor_noswizzle(smem_base + kOffBaug + qb * 2048u, 128u, 256u);
#pragma unroll
                        for (uint32_t k = 0; k < 3; ++k) {
                            const uint64_t a_aug = tc::smem_desc_kmajor_noswizzle(smem_base + kOffAaug + s * 768u + k * 256u, 128u, 0u);
                            mma2_bf16_ss(da, a_aug, bga, idesc, 1u);
                            mma2_bf16_ss(db, a_aug, bgb, idesc, 1u);
                        }
                        mma2_commit_both(&bars->acc_full[qa]);
                        mma2_commit_both(&bars->acc_full[qb]);
                        if (h == 1) mma2_commit_both(&bars->a_empty[s]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== epilogue groups (each CTA drains its own TMEM) =======================================
        const int g = warp >> 2;                               // 0: codes 0..255 (units 0, 1), 1: codes 256..511 (units 2, 3)
        const int r = (warp & 3) * 32 + lane;
        const uint32_t tbase = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)g * kHalfN;
        const float kInf = __uint_as_float(0x7f800000u);
        uint64_t* full0 = &bars->acc_full[2 * g];
        uint64_t* full1 = &bars->acc_full[2 * g + 1];
        uint64_t* empty0 = &bars->acc_empty[2 * g];
        uint64_t* empty1 = &bars->acc_empty[2 * g + 1];
        uint32_t it = 0;
        for (int64_t pair = pair0; pair < n_pairs; pair += pair_stride, ++it) {
            const int64_t n = (2 * pair + rank) * kTileM + r;
            mbar_wait_cluster(full0, it & 1u);
            tc::tc_fence_after_sync();
            float* dbg_row = (DBG && dbg != nullptr && n < N) ? dbg + n * kK + g * kHalfN : nullptr;
            Top2 tr[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { tr[k].best = kInf; tr[k].second = kInf; tr[k].chunk = 0; tr[k].trk = k; }
            uint32_t va[32], vb[32];
            tc::tmem_ld_32x32(tbase, va);
#pragma unroll 1
            for (int c = 0; c < 8; c += 2) {
                tmem_ld_wait_for(va);
                tc::tmem_ld_32x32(tbase + (c + 1) * 32, vb);
                epi_chunk<DBG>(va, c, tr, dbg_row);
                tmem_ld_wait_for(vb);
                if (c == 2) {
                    // first unit drained: hand it back (to the leader), then wait for the second one
                    tc::tc_fence_before_sync();
                    mbar_arrive_cluster(empty0, 0u);
                    mbar_wait_cluster(full1, it & 1u);
                    tc::tc_fence_after_sync();
                }
                if (c == 6) {
                    tc::tc_fence_before_sync();
                    mbar_arrive_cluster(empty1, 0u);
                } else {
                    tc::tmem_ld_32x32(tbase + (c + 2) * 32, va);
                }
                epi_chunk<DBG>(vb, c + 1, tr, dbg_row);
            }

            top2_merge(tr[0], tr[1]);
            top2_merge(tr[2], tr[3]);
            top2_merge(tr[0], tr[2]);
            const int i3 = (int)(__float_as_uint(tr[0].best) & 7u);
            const int code = g * kHalfN + tr[0].chunk * 32 + ((i3 >> 1) << 3) + (tr[0].trk << 1) + (i3 & 1);
            const uint32_t slot = it & 1u;
            float* x = xchg + (slot * 128u + r) * 3u;
            if (g == 0) {
                tc::mbar_wait(&bars->x_free[slot], ((it >> 1) & 1u) ^ 1u);
                x[0] = tr[0].best;
                x[1] = tr[0].second;
                x[2] = __int_as_float(code);
                tc::mbar_arrive(&bars->x_full[slot]);
            } else {
                tc::mbar_wait(&bars->x_full[slot], (it >> 1) & 1u);
                const float b0 = x[0], s0 = x[1];
                const int code0 = __float_as_int(x[2]);
                tc::mbar_arrive(&bars->x_free[slot]);
                const float b1 = tr[0].best, s1 = tr[0].second;
                const float best = fminf(b0, b1);
                const float second = fminf(fminf(s0, s1), fmaxf(b0, b1));
                const int win = (b1 < b0) ? code : code0;
                if (n < N) {
                    idx_out[n] = (long long)win;
                    const float thr = 1.25f * __uint_as_float((__float_as_uint(best) & 0x7F800000u) - (16u << 23));
                    if (!(second - best > thr)) {
                        const unsigned int pos = atomicAdd(list_count, 1u);
                        list[pos] = (int)n;
                    }
                }
            }
        }
    }

    // ---- teardown: nobody may leave while the partner can still touch its barriers / TMEM ----------------
    tc::tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 12) {
        tc::tc_fence_after_sync();
        tmem_dealloc2(tmem_base, 512);
    }
}

// Host launcher.  Needs an even grid (CTA pairs); `list` must hold N ints, `list_count` one zeroed counter.
int launch_vq_argmin_tc2(const float* z, int64_t N, int64_t HW, const float* E, long long* idx, int* list,
                         unsigned int* list_count, float* dbg, cudaStream_t st) {
    static thread_local int configured_dev = -1;
    int dev = 0;
    MOVAE_CUDA_TRY(cudaGetDevice(&dev));
    if (configured_dev != dev) {
        MOVAE_CUDA_TRY(cudaFuncSetAttribute(vq_argmin_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        MOVAE_CUDA_TRY(cudaFuncSetAttribute(vq_argmin_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        configured_dev = dev;
    }
    const int sms = sm_count();
    MOVAE_REQUIRE(sms > 0, MOVAE_ERR_CUDA, "CUDA device query failed (no GPU?)");
    const int64_t n_tiles = (N + kTileM - 1) / kTileM;
    const int64_t n_pairs = (n_tiles + 1) / 2;
    int64_t pairs = sms / 2;
    if (pairs > n_pairs) pairs = n_pairs;
    const unsigned grid = (unsigned)(2 * pairs);
    if (dbg)
        vq_argmin_tc2_kernel<true><<<grid, kThreads, kSmemBytes, st>>>(z, N, HW, E, idx, list, list_count, dbg);
    else
        vq_argmin_tc2_kernel<false><<<grid, kThreads, kSmemBytes, st>>>(z, N, HW, E, idx, list, list_count, nullptr);
    MOVAE_CUDA_TRY(cudaGetLastError());
    return MOVAE_OK;
}

}  // namespace movae
