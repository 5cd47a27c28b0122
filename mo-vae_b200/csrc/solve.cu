// K2 `solve_small`:  k x k Gramian -> weights w[k], entirely on the device (one CTA, no host sync).
//
// Replaces, in the reference (/root/reference):
//   * UPGrad  : torchjd UPGradWeighting -> project_weights -> qpsolvers/quadprog on the HOST in
//               float64 (G.cpu().numpy(), k QP calls, back to device; main.py:1195; same pipeline
//               visible at utils/torchmoo/nupgrad.py:122-126)
//   * MGDA    : utils/torchmoo/mgda.py:221-272 -- <= 250 Frank-Wolfe iterations of ~10 tiny kernels
//               with >= 3 host syncs each (python `if c <= a`), normalisers :274-285, :319-367,
//               eigen clamp :287-317
//   * AlignedMTL: utils/torchmoo/aligned_mtl.py:97-133 -- cuSOLVER eigh on a k x k + host syncs
//   * the gradient-similarity hook main.py:94-122 (two more passes over J): here it is computed
//     from G alone, 0 extra bytes.
// Latency-bound by construction (k <= 8); excluded from GB/s, included in steps/s.
#include <math.h>

#include "common.cuh"

namespace movae {

constexpr int MK = MOVAE_MAX_K;
constexpr int kSolveThreads = 256;   // 2^MK active-set candidates for the UPGrad QPs
constexpr float kEps32 = 1.1920928955078125e-07f;

enum SolveKind { SOLVE_CONST = 0, SOLVE_UPGRAD = 1, SOLVE_MGDA = 2, SOLVE_AMTL = 3 };

struct SolveParams {
    int kind;
    int k;
    float value;        // CONST
    float norm_eps;     // UPGRAD
    float reg_eps;      // UPGRAD
    int upgrad_norm;    // UPGRAD: MOVAE_UPGRAD_NORM_*
    int dualproj;       // UPGRAD: 1 = torchjd DualProj (ONE QP with the whole preference vector as lower bound)
    int norm_type;      // MGDA
    float epsilon;      // MGDA
    int max_iters;      // MGDA
    int stable;         // MGDA
    float min_eig_eps;  // MGDA
    int scale_mode;     // AMTL
};

// Cyclic Jacobi eigen-decomposition of a symmetric k x k matrix held in shared memory (single
// thread; k <= 8 => a few hundred rotations at most).  On exit A's diagonal holds the eigenvalues
// (unsorted) and V's columns the eigenvectors.  Only the upper triangle of the input is trusted
// (torch.linalg.eigh(UPLO="U"), aligned_mtl.py:108): it is mirrored first.
__device__ void jacobi_eigh(double (*A)[MK], double (*V)[MK], int k) {
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            V[i][j] = (i == j) ? 1.0 : 0.0;
            if (j < i) A[i][j] = A[j][i];
        }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < k; ++i) {
            diag += A[i][i] * A[i][i];
            for (int j = i + 1; j < k; ++j) off += A[i][j] * A[i][j];
        }
        if (off <= 1e-34 * diag || off == 0.0) break;
        for (int p = 0; p < k - 1; ++p)
            for (int q = p + 1; q < k; ++q) {
                const double apq = A[p][q];
                if (apq == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int r = 0; r < k; ++r) {   // A <- A R
                    const double arp = A[r][p], arq = A[r][q];
                    A[r][p] = c * arp - s * arq;
                    A[r][q] = s * arp + c * arq;
                }
                for (int r = 0; r < k; ++r) {   // A <- R^T A
                    const double apr = A[p][r], aqr = A[q][r];
                    A[p][r] = c * apr - s * aqr;
                    A[q][r] = s * apr + c * aqr;
                }
                A[p][q] = 0.0;
                A[q][p] = 0.0;
                for (int r = 0; r < k; ++r) {   // V <- V R
                    const double vrp = V[r][p], vrq = V[r][q];
                    V[r][p] = c * vrp - s * vrq;
                    V[r][q] = s * vrp + c * vrq;
                }
            }
    }
}

// Parallel-ordering Jacobi executed by one warp: the k (k - 1) / 2 rotations of a sweep are scheduled as a round-robin
// tournament -- n - 1 rounds (n = k rounded up to even) of n / 2 DISJOINT pairs.  Disjoint rotations commute and their
// parameters only depend on their own 2 x 2 blocks, so a round is exactly the sequential application of its rotations,
// but its column / row / eigenvector updates run as three warp-wide steps (lane = pair * 8 + index) and the float64
// divisions and square roots of the n / 2 rotation parameters run side by side: k = 8 has 7 dependent rounds per sweep
// instead of 28 dependent rotations (70 us -> ~20 us).  Same fixed point as jacobi_eigh (eigenvalues on A's diagonal,
// eigenvectors in V's columns), rounding-level differences only.
__device__ void jacobi_eigh_warp_rr(double (*A)[MK], double (*V)[MK], int k, int lane, double (*cs)[2], int (*pq)[2]) {
    for (int e = lane; e < MK * MK; e += 32) {
        const int i = e / MK, j = e % MK;
        if (i < k && j < k) {
            V[i][j] = (i == j) ? 1.0 : 0.0;
            if (j < i) A[i][j] = A[j][i];
        }
    }
    __syncwarp();
    const int n = (k + 1) & ~1, half = n / 2;
    const int m = lane >> 3, r = lane & 7;                       // pair slot, row / column index
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < k; ++i) {
            diag += A[i][i] * A[i][i];
            for (int j = i + 1; j < k; ++j) off += A[i][j] * A[i][j];
        }
        if (off <= 1e-34 * diag || off == 0.0) break;            // uniform: every lane read the same values
        for (int round = 0; round < n - 1; ++round) {
            __syncwarp();
            if (lane < half) {                                   // pairing of this round + rotation parameters
                int a, b;
                if (lane == 0) { a = n - 1; b = round; }
                else { a = (round + lane) % (n - 1); b = (round - lane + (n - 1)) % (n - 1); }
                int p = a < b ? a : b, q = a < b ? b : a;
                double c = 1.0, sn = 0.0;
                if (q >= k) { p = -1; q = -1; }                  // pair with the padding index of an odd k
                else {
                    const double apq = A[p][q];
                    if (apq != 0.0) {
                        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0);
                        sn = t * c;
                    } else { p = -1; q = -1; }
                }
                pq[lane][0] = p; pq[lane][1] = q;
                cs[lane][0] = c; cs[lane][1] = sn;
            }
            __syncwarp();
            const bool on = m < half && r < k && pq[m < half ? m : 0][0] >= 0;
            const int p = on ? pq[m][0] : 0, q = on ? pq[m][1] : 0;
            const double c = on ? cs[m][0] : 1.0, sn = on ? cs[m][1] : 0.0;
            if (on) {   // A <- A R
                const double arp = A[r][p], arq = A[r][q];
                A[r][p] = c * arp - sn * arq;
                A[r][q] = sn * arp + c * arq;
            }
            __syncwarp();
            if (on) {   // A <- R^T A
                const double apr = A[p][r], aqr = A[q][r];
                A[p][r] = c * apr - sn * aqr;
                A[q][r] = sn * apr + c * aqr;
            }
            __syncwarp();
            if (on) {   // V <- V R, and the annihilated pair set exactly to zero
                if (r == 0) { A[p][q] = 0.0; A[q][p] = 0.0; }
                const double vrp = V[r][p], vrq = V[r][q];
                V[r][p] = c * vrp - sn * vrq;
                V[r][q] = sn * vrp + c * vrq;
            }
        }
        __syncwarp();
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// UPGrad: k strictly convex QPs  min 1/2 x^T H x  s.t. x >= lo_i e_i  by exhaustive active-set
// enumeration: thread s owns the active set encoded by the bits of s; the candidate with the
// smallest KKT violation is the (unique) optimum.  All float64.  The reduced system of an active set is
// embedded in a K x K system (rows / columns of the active set replaced by identity) and solved by
// Gaussian elimination without pivoting (the free block is a principal submatrix of the SPD matrix H);
// every loop bound is a template constant, so the matrix lives in registers instead of dynamically
// indexed local memory.
// ------------------------------------------------------------------------------------------------
// All k QPs of one active set at once.  The k QPs  min 1/2 x^T H x  s.t. x >= lo_i e_i  share H, and for a given active
// set the embedded K x K system matrix is the SAME for every i -- only the right-hand side changes.  So each thread
// eliminates its matrix ONCE (keeping the multipliers in the strict lower triangle) and runs k forward / backward
// substitutions: ~(1/3 K^3 + k * 2 K^2) multiply-adds instead of k * (1/3 K^3 + 2 K^2), and the k block-wide argmin
// reductions collapse into one round.  solve_i() recomputes x for the winning set of QP i from the resident factors.
template <int KT>
struct UpgradSet {
    double M[KT][KT];          // U on and above the diagonal, elimination multipliers below
    unsigned mask;

    __device__ void factor(const double (*H)[MK], unsigned m) {
        mask = m;
#pragma unroll
        for (int a = 0; a < KT; ++a) {
            const bool aa = (m >> a) & 1u;
#pragma unroll
            for (int b = 0; b < KT; ++b) {
                const bool ab = (m >> b) & 1u;
                M[a][b] = (aa || ab) ? ((a == b) ? 1.0 : 0.0) : H[a][b];
            }
        }
#pragma unroll
        for (int c = 0; c < KT; ++c) {
            const double inv = 1.0 / M[c][c];
#pragma unroll
            for (int r = c + 1; r < KT; ++r) {
                const double f = M[r][c] * inv;
#pragma unroll
                for (int cc = c + 1; cc < KT; ++cc) M[r][cc] -= f * M[c][cc];
                M[r][c] = f;
            }
        }
    }

    // x for QP i (lower bound lo_i on coordinate i, 0 elsewhere); returns the KKT violation of this active set
    __device__ double solve_i(const double (*H)[MK], int i, double lo_i, double* xs) const {
        const bool i_active = (mask >> i) & 1u;
        double rhs[KT];
#pragma unroll
        for (int a = 0; a < KT; ++a) {
            const bool aa = (mask >> a) & 1u;
            rhs[a] = aa ? ((a == i) ? lo_i : 0.0) : (i_active ? -H[a][i] * lo_i : 0.0);
        }
#pragma unroll
        for (int c = 0; c < KT; ++c)
#pragma unroll
            for (int r = c + 1; r < KT; ++r) rhs[r] -= M[r][c] * rhs[c];
#pragma unroll
        for (int r = KT - 1; r >= 0; --r) {
            double acc = rhs[r];
#pragma unroll
            for (int cc = r + 1; cc < KT; ++cc) acc -= M[r][cc] * xs[cc];
            xs[r] = acc / M[r][r];
        }
        double viol = 0.0;
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            if ((mask >> j) & 1u) {
                double g = 0.0;
#pragma unroll
                for (int c = 0; c < KT; ++c) g += H[j][c] * xs[c];
                viol = fmax(viol, -g);                                   // multiplier must be >= 0
            } else {
                viol = fmax(viol, ((j == i) ? lo_i : 0.0) - xs[j]);      // free coordinate must stay feasible
            }
        }
        return viol;
    }

    // x for the single QP  min 1/2 x^T H x  s.t. x >= lo  (every coordinate bounded: DualProj); returns the KKT violation
    __device__ double solve_vec(const double (*H)[MK], const double* lo, double* xs) const {
        double rhs[KT];
#pragma unroll
        for (int a = 0; a < KT; ++a) {
            if ((mask >> a) & 1u) {
                rhs[a] = lo[a];
            } else {
                double acc = 0.0;
#pragma unroll
                for (int b = 0; b < KT; ++b) acc -= ((mask >> b) & 1u) ? H[a][b] * lo[b] : 0.0;
                rhs[a] = acc;
            }
        }
#pragma unroll
        for (int c = 0; c < KT; ++c)
#pragma unroll
            for (int r = c + 1; r < KT; ++r) rhs[r] -= M[r][c] * rhs[c];
#pragma unroll
        for (int r = KT - 1; r >= 0; --r) {
            double acc = rhs[r];
#pragma unroll
            for (int cc = r + 1; cc < KT; ++cc) acc -= M[r][cc] * xs[cc];
            xs[r] = acc / M[r][r];
        }
        double viol = 0.0;
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            if ((mask >> j) & 1u) {
                double g = 0.0;
#pragma unroll
                for (int c = 0; c < KT; ++c) g += H[j][c] * xs[c];
                viol = fmax(viol, -g);
            } else {
                viol = fmax(viol, lo[j] - xs[j]);
            }
        }
        return viol;
    }
};

// Block-wide DualProj solve for k == KT (torchjd `DualProj`, selectable at main.py:1221-1222): the projection of the
// preference vector u (default 1/k each) onto the dual cone, ONE QP  argmin_{v >= u} v^T H v  over the same 2^k sets.
template <int KT>
__device__ void dualproj_all(const double (*H)[MK], const float* __restrict__ pref, float* w, double* dg, double* red_v, int* red_i,
                             double (*xbest)[MK], int tid) {
    constexpr unsigned n_sets = 1u << KT;
    constexpr int kWarps = kSolveThreads / 32;
    UpgradSet<KT> set;
    __shared__ double lo[MK];
    if (tid < KT) lo[tid] = (double)(pref ? pref[tid] : __fdiv_rn(1.0f, (float)KT));
    __syncthreads();
    const bool has = (unsigned)tid < n_sets;
    double x[KT];
    double bv = 1e300;
    if (has) {
        set.factor(H, (unsigned)tid);
        bv = set.solve_vec(H, lo, x);
    }
    int bi = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { red_v[tid >> 5] = bv; red_i[tid >> 5] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int q = 1; q < kWarps; ++q)
            if (red_v[q] < bv || (red_v[q] == bv && red_i[q] < bi)) { bv = red_v[q]; bi = red_i[q]; }
        red_v[0] = bv;
        red_i[0] = bi;
    }
    __syncthreads();
    if (has && tid == red_i[0]) {
#pragma unroll
        for (int j = 0; j < KT; ++j) xbest[0][j] = x[j];
    }
    __syncthreads();
    if (tid < KT) w[tid] = (float)xbest[0][tid];
    if (tid == 0) {
        dg[MOVAE_DIAG_RESIDUAL] = red_v[0];
        dg[MOVAE_DIAG_STATUS] = (red_v[0] <= 1e-9) ? 0.0 : 1.0;
    }
}

// Block-wide UPGrad solve for k == KT: writes w (float32 sums of the float32-cast projections) and the worst violation.
template <int KT>
__device__ void upgrad_all(const double (*H)[MK], const float* __restrict__ pref, float* w, double* dg, double* red_v, int* red_i,
                           double (*xbest)[MK], int tid) {
    constexpr unsigned n_sets = 1u << KT;
    constexpr int kWarps = kSolveThreads / 32;
    UpgradSet<KT> set;
    __shared__ double lo[MK];
    if (tid < KT) lo[tid] = (double)(pref ? pref[tid] : __fdiv_rn(1.0f, (float)KT));
    __syncthreads();
    const bool has = (unsigned)tid < n_sets;
    if (has) set.factor(H, (unsigned)tid);
    // The loop over the QPs is deliberately NOT unrolled: fully unrolled, the k = 8 kernel was ~10,000 straight-line
    // instructions per thread executed once each and ran instruction-fetch bound (ncu: 49% of the stall samples
    // `no_instruction`); rolled, the ~250-instruction body is fetched once and replayed k times.  Per-warp argmin by
    // shuffles (ties -> lowest candidate index) inside the loop, one block-level round after it.
#pragma unroll 1
    for (int i = 0; i < KT; ++i) {
        double x[KT];
        double bv = has ? set.solve_i(H, i, lo[i], x) : 1e300;
        int bi = tid;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { red_v[i * kWarps + (tid >> 5)] = bv; red_i[i * kWarps + (tid >> 5)] = bi; }
    }
    __syncthreads();
    if (tid < KT) {
        double bv = red_v[tid * kWarps];
        int bi = red_i[tid * kWarps];
        for (int q = 1; q < kWarps; ++q) {
            const double ov = red_v[tid * kWarps + q];
            const int oi = red_i[tid * kWarps + q];
            if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        red_v[tid * kWarps] = bv;
        red_i[tid * kWarps] = bi;
    }
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < KT; ++i) {
        if (tid == red_i[i * kWarps]) {
            double x[KT];
            set.solve_i(H, i, lo[i], x);
#pragma unroll
            for (int j = 0; j < KT; ++j) xbest[i][j] = x[j];
        }
    }
    __syncthreads();
    if (tid < KT) {
        // W.sum(dim=0) on the float32-cast rows (torchjd casts W back to G's dtype first), rows added in order i = 0..k-1
        float acc = 0.f;
        for (int i = 0; i < KT; ++i) acc = __fadd_rn(acc, (float)xbest[i][tid]);
        w[tid] = acc;
    }
    if (tid == 0) {
        double worst = 0.0;
        for (int i = 0; i < KT; ++i) worst = fmax(worst, red_v[i * kWarps]);
        dg[MOVAE_DIAG_RESIDUAL] = worst;
        dg[MOVAE_DIAG_STATUS] = (worst <= 1e-9) ? 0.0 : 1.0;      // NaN / inf Gramian -> status 1 (torchjd raises ValueError)
    }
}

__global__ void __launch_bounds__(kSolveThreads)
solve_kernel(SolveParams p, const double* __restrict__ G_in, const float* __restrict__ pref,
             const float* __restrict__ losses, float* __restrict__ w_out, double* __restrict__ diag, P2PArgs px,
             double* __restrict__ G_sum) {
    __shared__ double G[MK][MK];     // float64 Gramian as produced by K1 (+ allreduce)
    __shared__ float Gf[MK][MK];     // rounded once to float32: the reference's `J @ J.T` tensor
    __shared__ double H[MK][MK];     // work matrix
    __shared__ double V[MK][MK];
    __shared__ float w[MK];
    __shared__ double dg[MOVAE_DIAG_DOUBLES];
    __shared__ double red_v[MK * (kSolveThreads / 32)];
    __shared__ int red_i[MK * (kSolveThreads / 32)];
    __shared__ double xbest[MK][MK];
    __shared__ double rot_cs[MK / 2][2];
    __shared__ int rot_pq[MK / 2][2];
    const int k = p.k;
    const int tid = threadIdx.x;

    if (tid < MOVAE_DIAG_DOUBLES) dg[tid] = 0.0;
    __shared__ int xchg_timeout;
    if (tid == 0) xchg_timeout = 0;
    __syncthreads();
    if (tid < MK * MK) {
        const int i = tid / MK, j = tid % MK;
        double g = 0.0;
        if (i < k && j < k) {
            if (px.world > 0) {
                // fused gather head (P-sharded aggregation): wait for every rank's partial in the own exchange
                // buffer, then sum in rank order -> bit-identical Gramian on every rank
                const int par = (int)(px.seq & 1ull);
                const XchgBuffer* mine = px.peers[px.rank];
                for (int r = 0; r < px.world; ++r) {
                    unsigned int polls = 0;
                    while (ld_acquire_sys_u64(&mine->flags[par][r]) < px.seq) {
                        if (++polls > (1u << 27)) { xchg_timeout = 1; break; }
                    }
                    g += ld_relaxed_sys_f64(&mine->slots[par][r][i * k + j]);
                }
                if (G_sum) G_sum[i * k + j] = g;
            } else {
                g = G_in[i * k + j];
            }
        }
        G[i][j] = g;
        Gf[i][j] = (float)g;
    }
    if (tid < MK) w[tid] = 0.f;
    __syncthreads();

    if (p.kind == SOLVE_CONST) {
        if (tid < k) w[tid] = p.value;
    } else if (p.kind == SOLVE_UPGRAD) {
        // normalize / regularize, in float32 like the reference does on the float32 Gramian tensor
        if (tid == 0) {
            float tr = 0.f;
            for (int i = 0; i < k; ++i) tr += Gf[i][i];
            dg[MOVAE_DIAG_TRACE] = tr;
            float sc[MK];                                      // per-row scale of the two l2-based normalisations
            bool all_zero = false;
            if (p.upgrad_norm == MOVAE_UPGRAD_NORM_MIN_L2) {
                // nupgrad.py:129-158: l2 = sqrt(clamp(diag, eps)); rows with l2 > eps are scaled to the smallest such norm
                float l2[MK], mn = __uint_as_float(0x7f800000u);
                bool any = false;
                for (int i = 0; i < k; ++i) {
                    l2[i] = __fsqrt_rn(fmaxf(Gf[i][i], p.norm_eps));
                    if (l2[i] > p.norm_eps) { any = true; mn = fminf(mn, l2[i]); }
                }
                all_zero = !any;
                for (int i = 0; i < k; ++i) sc[i] = (l2[i] > p.norm_eps) ? __fdiv_rn(mn, l2[i]) : 0.f;
            } else if (p.upgrad_norm == MOVAE_UPGRAD_NORM_L2) {
                // nupgrad.py:14-24 / pnupgrad.py `normalize`: G / (|g_i| |g_j|), norms = sqrt(clamp(diag, eps))
                for (int i = 0; i < k; ++i) sc[i] = __fsqrt_rn(fmaxf(Gf[i][i], p.norm_eps));
            }
            for (int i = 0; i < k; ++i)
                for (int j = 0; j < k; ++j) {
                    float gn;
                    if (p.upgrad_norm == MOVAE_UPGRAD_NORM_MIN_L2) gn = all_zero ? 0.f : __fmul_rn(Gf[i][j], __fmul_rn(sc[i], sc[j]));
                    else if (p.upgrad_norm == MOVAE_UPGRAD_NORM_L2) gn = __fdiv_rn(Gf[i][j], __fmul_rn(sc[i], sc[j]));
                    else gn = (tr < p.norm_eps) ? 0.f : __fdiv_rn(Gf[i][j], tr);      // torchjd `normalize`: divide by the trace
                    H[i][j] = (double)__fadd_rn(gn, (i == j) ? p.reg_eps : 0.f);
                }
        }
        __syncthreads();
        if (p.dualproj) {
            switch (k) {
                case 1: dualproj_all<1>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
                case 2: dualproj_all<2>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
                case 3: dualproj_all<3>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
                case 4: dualproj_all<4>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
                case 5: dualproj_all<5>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
                case 6: dualproj_all<6>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
                case 7: dualproj_all<7>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
                default: dualproj_all<8>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
            }
        } else {
            switch (k) {
                case 1: upgrad_all<1>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
                case 2: upgrad_all<2>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
                case 3: upgrad_all<3>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
                case 4: upgrad_all<4>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
                case 5: upgrad_all<5>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
                case 6: upgrad_all<6>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
                case 7: upgrad_all<7>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
                default: upgrad_all<8>(H, pref, w, dg, red_v, red_i, xbest, tid); break;
            }
        }
    } else if (p.kind == SOLVE_MGDA) {
        if (tid == 0) {
            // --- normalisation (float32, IEEE ops, no contraction) ---
            float s[MK];
            float (*R)[MK] = Gf;
            if (p.norm_type != MOVAE_MGDA_NONE) {
                for (int i = 0; i < k; ++i) {
                    const float ell = (p.norm_type == MOVAE_MGDA_L2) ? 1.f : fmaxf(losses[i], 1e-20f);
                    const float nrm = (p.norm_type == MOVAE_MGDA_LOSS) ? 1.f : __fsqrt_rn(fmaxf(Gf[i][i], 1e-20f));
                    s[i] = (p.norm_type == MOVAE_MGDA_L2) ? nrm : (p.norm_type == MOVAE_MGDA_LOSS ? ell : __fmul_rn(ell, nrm));
                }
                for (int i = 0; i < k; ++i)
                    for (int j = 0; j < k; ++j) R[i][j] = __fdiv_rn(Gf[i][j], __fmul_rn(s[i], s[j]));
            }
            if (p.stable) {   // eigen clamp, mgda.py:287-317 (float64 Jacobi on the float32 matrix)
                for (int i = 0; i < k; ++i)
                    for (int j = 0; j < k; ++j) H[i][j] = (double)R[i][j];
                jacobi_eigh(H, V, k);
                for (int i = 0; i < k; ++i)
                    for (int j = 0; j < k; ++j) {
                        double acc = 0.0;
                        for (int c = 0; c < k; ++c) acc += V[i][c] * fmax(H[c][c], (double)p.min_eig_eps) * V[j][c];
                        R[i][j] = (float)acc;
                    }
            }
            // --- Frank-Wolfe, op-for-op in float32 (mgda.py:244-262) ---
            float alpha[MK], Ra[MK];
            for (int i = 0; i < k; ++i) alpha[i] = __fdiv_rn(1.0f, (float)k);
            float gamma = 0.f;
            int it = 0;
            for (; it < p.max_iters; ++it) {
                int t = 0;
                for (int i = 0; i < k; ++i) {
                    float acc = 0.f;
                    for (int j = 0; j < k; ++j) acc = fmaf(R[i][j], alpha[j], acc);
                    Ra[i] = acc;
                    if (acc < Ra[t]) t = i;          // first minimal index
                }
                float a = 0.f, b = 0.f;
                for (int j = 0; j < k; ++j) {
                    a = fmaf(alpha[j], R[j][t], a);
                    b = fmaf(alpha[j], Ra[j], b);
                }
                const float c = R[t][t];
                if (c <= a) gamma = 1.f;
                else if (b <= a) gamma = 0.f;
                else gamma = __fdiv_rn(__fsub_rn(b, a), __fsub_rn(__fadd_rn(b, c), __fmul_rn(2.f, a)));
                bool changed = false;
                const float om = __fsub_rn(1.f, gamma);
                for (int j = 0; j < k; ++j) {
                    const float nv = __fadd_rn(__fmul_rn(om, alpha[j]), __fmul_rn(gamma, (j == t) ? 1.f : 0.f));
                    changed |= (nv != alpha[j]);
                    alpha[j] = nv;
                }
                if (gamma < p.epsilon) { ++it; break; }
                if (!changed) { it = p.max_iters; break; }   // exact fixpoint: the reference spins to max_iters with identical state
            }
            if (p.max_iters <= 0) it = 0;
            for (int i = 0; i < k; ++i) w[i] = alpha[i];
            dg[MOVAE_DIAG_COUNT] = (double)it;
            dg[MOVAE_DIAG_GAMMA] = (double)gamma;
        }
    } else if (p.kind == SOLVE_AMTL) {
        if (tid < MK * MK) H[tid / MK][tid % MK] = (double)Gf[tid / MK][tid % MK];
        __syncthreads();
        if (tid < 32) jacobi_eigh_warp_rr(H, V, k, tid, rot_cs, rot_pq);
        __syncthreads();
        if (tid == 0) {
            double lam[MK];
            int order[MK];
            double lmax = -1e300;
            for (int i = 0; i < k; ++i) { lam[i] = H[i][i]; order[i] = i; lmax = fmax(lmax, lam[i]); }
            const double tol = lmax * (double)k * (double)kEps32;     // aligned_mtl.py:109
            int rank = 0;
            for (int i = 0; i < k; ++i) rank += (lam[i] > tol) ? 1 : 0;
            for (int i = 1; i < k; ++i) {                              // insertion sort, descending
                const int oi = order[i];
                int j = i - 1;
                while (j >= 0 && lam[order[j]] < lam[oi]) { order[j + 1] = order[j]; --j; }
                order[j + 1] = oi;
            }
            double w0[MK];
            for (int i = 0; i < k; ++i) w0[i] = (double)(pref ? pref[i] : __fdiv_rn(1.0f, (float)k));
            dg[MOVAE_DIAG_RANK] = (double)rank;
            if (rank == 0) {
                for (int i = 0; i < k; ++i) w[i] = (float)w0[i];       // B = I
            } else {
                double scale;
                if (p.scale_mode == MOVAE_AMTL_MIN) scale = lam[order[rank - 1]];
                else if (p.scale_mode == MOVAE_AMTL_MEDIAN) scale = lam[order[rank - 1 - (rank - 1) / 2]];   // lower middle
                else { scale = 0.0; for (int r = 0; r < rank; ++r) scale += lam[order[r]]; scale /= (double)rank; }
                double out[MK];
                for (int i = 0; i < k; ++i) out[i] = 0.0;
                for (int r = 0; r < rank; ++r) {
                    const int c = order[r];
                    double proj = 0.0;
                    for (int i = 0; i < k; ++i) proj += V[i][c] * w0[i];
                    proj /= sqrt(lam[c]);
                    for (int i = 0; i < k; ++i) out[i] += V[i][c] * proj;
                }
                const double ss = sqrt(scale);
                for (int i = 0; i < k; ++i) w[i] = (float)(ss * out[i]);
            }
        }
    }
    __syncthreads();

    if (tid == 0) {
        // cos(J^T w, J^T 1/k) = (w^T G m) / (|J^T w| |J^T m|)   (F.cosine_similarity clamps the norm product at 1e-8)
        double num = 0.0, ww = 0.0, mm = 0.0;
        const double m = 1.0 / (double)k;
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < k; ++j) {
                num += (double)w[i] * G[i][j] * m;
                ww += (double)w[i] * G[i][j] * (double)w[j];
                mm += m * G[i][j] * m;
            }
        dg[MOVAE_DIAG_SIMILARITY] = num / fmax(sqrt(fmax(ww, 0.0)) * sqrt(fmax(mm, 0.0)), 1e-8);
        if (p.kind != SOLVE_UPGRAD) {
            double tr = 0.0;
            for (int i = 0; i < k; ++i) tr += (double)Gf[i][i];
            dg[MOVAE_DIAG_TRACE] = tr;
        }
    }
    __syncthreads();
    if (tid < k) w_out[tid] = w[tid];
    if (tid == 0) {
        // non-finite weights (a NaN / inf Jacobian upstream): surfaced through STATUS so that UPGrad's check_status() raises
        // like torchjd does when quadprog fails, instead of passing NaN gradients on silently
        bool finite = true;
        for (int i = 0; i < k; ++i) finite = finite && (fabsf(w[i]) <= 3.4028234e38f);
        if (!finite && dg[MOVAE_DIAG_STATUS] == 0.0) dg[MOVAE_DIAG_STATUS] = 1.0;
        if (xchg_timeout) dg[MOVAE_DIAG_STATUS] = 2.0;
    }
    __syncthreads();
    if (tid < MOVAE_DIAG_DOUBLES && diag) diag[tid] = dg[tid];
}

static thread_local P2PArgs g_px = p2p_disabled();     // set by movae_solve_p2p around the generic dispatch
static thread_local double* g_G_sum = nullptr;

static int launch_solve(const SolveParams& p, const double* G, const float* pref, const float* losses, float* w,
                        double* diag, void* stream) {
    MOVAE_REQUIRE(p.k >= 1, MOVAE_ERR_INVALID, "solve: k must be >= 1 (got %d)", p.k);
    MOVAE_REQUIRE(p.k <= MOVAE_MAX_K, MOVAE_ERR_UNSUPPORTED, "solve: k=%d > MOVAE_MAX_K=%d", p.k, MOVAE_MAX_K);
    MOVAE_REQUIRE((G || g_px.world > 0) && w, MOVAE_ERR_INVALID, "solve: null pointer");
    const int threads = (p.kind == SOLVE_UPGRAD) ? kSolveThreads : 64;
    solve_kernel<<<1, threads, 0, static_cast<cudaStream_t>(stream)>>>(p, G, pref, losses, w, diag, g_px, g_G_sum);
    MOVAE_CUDA_TRY(cudaGetLastError());
    return MOVAE_OK;
}

void set_solve_p2p(const P2PArgs& px, double* G_sum) {
    g_px = px;
    g_G_sum = G_sum;
}

}  // namespace movae

extern "C" {

int movae_solve_constant(const double* d_G, int k, float value, float* d_w, double* d_diag, void* stream) {
    movae::SolveParams p{};
    p.kind = movae::SOLVE_CONST;
    p.k = k;
    p.value = value;
    return movae::launch_solve(p, d_G, nullptr, nullptr, d_w, d_diag, stream);
}

int movae_solve_upgrad(const double* d_G, int k, const float* d_pref, float norm_eps, float reg_eps, float* d_w,
                       double* d_diag, void* stream) {
    movae::SolveParams p{};
    p.kind = movae::SOLVE_UPGRAD;
    p.k = k;
    p.norm_eps = norm_eps;
    p.reg_eps = reg_eps;
    p.upgrad_norm = MOVAE_UPGRAD_NORM_TRACE;
    return movae::launch_solve(p, d_G, d_pref, nullptr, d_w, d_diag, stream);
}

int movae_solve_nupgrad(const double* d_G, int k, const float* d_pref, float norm_eps, float reg_eps, int norm_mode, float* d_w,
                        double* d_diag, void* stream) {
    using namespace movae;
    MOVAE_REQUIRE(norm_mode >= MOVAE_UPGRAD_NORM_TRACE && norm_mode <= MOVAE_UPGRAD_NORM_L2, MOVAE_ERR_INVALID,
                  "nupgrad: bad norm_mode %d", norm_mode);
    SolveParams p{};
    p.kind = SOLVE_UPGRAD;
    p.k = k;
    p.norm_eps = norm_eps;
    p.reg_eps = reg_eps;
    p.upgrad_norm = norm_mode;
    return launch_solve(p, d_G, d_pref, nullptr, d_w, d_diag, stream);
}

int movae_solve_dualproj(const double* d_G, int k, const float* d_pref, float norm_eps, float reg_eps, float* d_w, double* d_diag,
                         void* stream) {
    movae::SolveParams p{};
    p.kind = movae::SOLVE_UPGRAD;
    p.k = k;
    p.norm_eps = norm_eps;
    p.reg_eps = reg_eps;
    p.upgrad_norm = MOVAE_UPGRAD_NORM_TRACE;
    p.dualproj = 1;
    return movae::launch_solve(p, d_G, d_pref, nullptr, d_w, d_diag, stream);
}

int movae_solve_mgda(const double* d_G, int k, int norm_type, const float* d_losses, float epsilon, int max_iters,
                     int stable, float min_eigenvalue_eps, float* d_w, double* d_diag, void* stream) {
    using namespace movae;
    MOVAE_REQUIRE(norm_type >= MOVAE_MGDA_NONE && norm_type <= MOVAE_MGDA_LOSS_PLUS, MOVAE_ERR_INVALID,
                  "mgda: bad norm_type %d", norm_type);
    MOVAE_REQUIRE(d_losses || norm_type == MOVAE_MGDA_NONE || norm_type == MOVAE_MGDA_L2, MOVAE_ERR_INVALID,
                  "mgda: losses must be set for norm_type 'loss'/'loss+'");
    SolveParams p{};
    p.kind = SOLVE_MGDA;
    p.k = k;
    p.norm_type = norm_type;
    p.epsilon = epsilon;
    p.max_iters = max_iters;
    p.stable = stable;
    p.min_eig_eps = min_eigenvalue_eps;
    return launch_solve(p, d_G, nullptr, d_losses, d_w, d_diag, stream);
}

int movae_solve_aligned_mtl(const double* d_G, int k, int scale_mode, const float* d_pref, float* d_w, double* d_diag,
                            void* stream) {
    using namespace movae;
    MOVAE_REQUIRE(scale_mode >= MOVAE_AMTL_MIN && scale_mode <= MOVAE_AMTL_RMSE, MOVAE_ERR_INVALID,
                  "aligned_mtl: bad scale_mode %d", scale_mode);
    SolveParams p{};
    p.kind = SOLVE_AMTL;
    p.k = k;
    p.scale_mode = scale_mode;
    return launch_solve(p, d_G, d_pref, nullptr, d_w, d_diag, stream);
}

}  // extern "C"
