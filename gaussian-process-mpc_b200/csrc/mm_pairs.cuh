// Pair-sum kernels of the moment-matching step (the dominant cost of the whole path).
//
// For one Gaussian input N(u, diag(s)) and one GP output a, the reference evaluates
// (src/tools/uncertainty_prop.py:372-399)
//     T_a = sum_ij (Ky_a^-1 - beta_a beta_a^T)_ij L_ij,
//     L_ij = sf^4 |2 Lam^-1 S + I|^-1/2 exp(-1/8 (v_i+v_j)^T A (v_i+v_j)) exp(-1/4 (x_i-x_j)^T Lam^-1 (x_i-x_j))
// with v_i = u - x_i, A = (Lam/2 + S)^-1, through ~12 dense [n,n] temporaries and an n^3 matmul whose
// trace is taken.  Here the constant factors live in Wt (fit.cu) and one pass over the upper-triangular
// pair space computes, per (rollout, output),
//     T  = sum Wt_ij e_ij,   N1_k = sum Wt_ij e_ij q_k,   N2_k = sum Wt_ij e_ij q_k^2,
//     q_k = c_k (v_ik + v_jk),  c_k = sqrt(a_k / 8),  e_ij = exp(-sum_k q_k^2)
// (the first/second moments give d T / d u and d T / d s in closed form for the adjoint).
//
// mm_pairs_batch: lanes <-> rollouts.  Every lane of a warp owns one rollout b and keeps its
// accumulators in registers for the whole kernel; the pair data (x_i, x_j, Wt_ij of up to 4 outputs that
// share lambda and hence the exp) is warp-uniform and read from shared memory by broadcast, so the inner
// loop has no cross-lane traffic at all.  Shared memory is filled by cp.async double buffering of 32x32
// Wt tiles.  The kernel is bound by the FP64 pipe (one exp + ~14 + 12*EG DFMA-class ops per pair).
#pragma once
#include "common.cuh"

namespace gpmpc {

// exp(-S) for S >= 0 (clamped at 700): Cody-Waite reduction + degree-11 polynomial (|err| < 2e-17 before
// rounding) + exponent insertion.  Branch free; ~16 FP64-pipe ops.
__device__ __forceinline__ double exp_neg(double S)
{
    S = fmin(S, 700.0);
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52
    double t = fma(S, -1.4426950408889634, MAGIC);          // round(-S * log2 e) in the low word
    const int k = __double2loint(t);
    const double kf = t - MAGIC;
    double r = fma(kf, -6.93147180369123816490e-01, -S);     // ln2 hi
    r = fma(kf, -1.90821492927058770002e-10, r);             // ln2 lo
    double p = 2.5110037605963777e-08;
    p = fma(p, r, 2.763263963904103e-07);
    p = fma(p, r, 2.755724091857897e-06);
    p = fma(p, r, 2.4801485482328494e-05);
    p = fma(p, r, 0.00019841269890047113);
    p = fma(p, r, 0.0013888888952314775);
    p = fma(p, r, 0.008333333333319601);
    p = fma(p, r, 0.0416666666664881);
    p = fma(p, r, 0.1666666666666668);
    p = fma(p, r, 0.5000000000000019);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (int)((unsigned)k << 20), __double2loint(p));
}

__device__ __forceinline__ void cpa16(void *smem, const void *gmem)
{
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct PairArgs {
    const double *Wt[kGroupMax];   // weight matrices of the group's outputs (ld x ld, upper-tri weights)
    int out_idx[kGroupMax];        // global output index of each member
    const double *X;               // [ld, D]
    const double *cst;             // this group's per-rollout constants [4D][Bpad]: c, c*u, cm, cm*u
    double *part;                  // [P][E][nacc][Bpad]
    int ld, ntile, B, Bpad, E, P;
    long long total_tiles;
};

constexpr int PT = kPairTile;      // 32
constexpr int PAIR_THREADS = 128;
constexpr int RI = 4;              // rows of the register micro-tile

template <int D, int EG>
__host__ __device__ constexpr size_t pair_stage_doubles() { return (size_t)EG * PT * PT + 2 * PT * D; }

template <int D, int EG, bool GRAD>
__global__ void __launch_bounds__(PAIR_THREADS, 2) mm_pairs_batch(const PairArgs a)
{
    extern __shared__ __align__(16) double smem[];
    constexpr size_t STAGE = pair_stage_doubles<D, EG>();
    const int tid = threadIdx.x;
    const int b = blockIdx.y * PAIR_THREADS + tid;
    const bool active = b < a.B;

    // per-rollout constants
    double c[D], cu[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        c[k] = active ? a.cst[(size_t)k * a.Bpad + b] : 0.0;
        cu[k] = active ? a.cst[(size_t)(D + k) * a.Bpad + b] : 0.0;
    }

    double accT[EG], acc1[GRAD ? EG : 1][D], acc2[GRAD ? EG : 1][D];
#pragma unroll
    for (int g = 0; g < EG; ++g) accT[g] = 0.0;
    if (GRAD) {
#pragma unroll
        for (int g = 0; g < EG; ++g)
#pragma unroll
            for (int k = 0; k < D; ++k) acc1[g][k] = acc2[g][k] = 0.0;
    }

    // this block's contiguous range of upper-triangular tiles
    const long long t_begin = a.total_tiles * blockIdx.x / a.P;
    const long long t_end = a.total_tiles * (blockIdx.x + 1) / a.P;
    int I = 0, J = 0;
    {
        long long rem = t_begin;
        int row = 0;
        while (rem >= a.ntile - row) { rem -= a.ntile - row; ++row; }
        I = row; J = row + (int)rem;
    }

    auto issue = [&](int stage, int ti, int tj) {
        double *base = smem + (size_t)stage * STAGE;
#pragma unroll
        for (int g = 0; g < EG; ++g) {
            const double *src = a.Wt[g] + (size_t)ti * PT * a.ld + (size_t)tj * PT;
            double *dst = base + (size_t)g * PT * PT;
#pragma unroll
            for (int q = 0; q < (PT * PT / 2) / PAIR_THREADS; ++q) {
                const int chunk = tid + q * PAIR_THREADS;      // 16-byte chunk id: 16 per row
                const int r = chunk >> 4, cc = (chunk & 15) * 2;
                cpa16(dst + r * PT + cc, src + (size_t)r * a.ld + cc);
            }
        }
        double *xi = base + (size_t)EG * PT * PT;
        double *xj = xi + PT * D;
        for (int chunk = tid; chunk < PT * D / 2; chunk += PAIR_THREADS) {
            cpa16(xi + chunk * 2, a.X + (size_t)ti * PT * D + chunk * 2);
            cpa16(xj + chunk * 2, a.X + (size_t)tj * PT * D + chunk * 2);
        }
        cpa_commit();
    };

    if (t_begin < t_end) issue(0, I, J);
    int stage = 0;
    for (long long t = t_begin; t < t_end; ++t) {
        int In = I, Jn = J + 1;
        if (Jn == a.ntile) { ++In; Jn = In; }
        if (t + 1 < t_end) { issue(stage ^ 1, In, Jn); cpa_wait<1>(); }
        else cpa_wait<0>();
        __syncthreads();

        const double *Ws = smem + (size_t)stage * STAGE;
        const double *xi = Ws + (size_t)EG * PT * PT;
        const double *xj = xi + PT * D;

#pragma unroll 1
        for (int r0 = 0; r0 < PT; r0 += RI) {
            double zi[RI][D];
#pragma unroll
            for (int r = 0; r < RI; ++r)
#pragma unroll
                for (int k = 0; k < D; ++k) zi[r][k] = fma(-c[k], xi[(r0 + r) * D + k], cu[k]);
#pragma unroll 1
            for (int j = 0; j < PT; ++j) {
                double zj[D];
#pragma unroll
                for (int k = 0; k < D; ++k) zj[k] = fma(-c[k], xj[j * D + k], cu[k]);
#pragma unroll
                for (int r = 0; r < RI; ++r) {
                    double q[D], qq[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) { q[k] = zi[r][k] + zj[k]; qq[k] = q[k] * q[k]; }
                    double S = qq[0];
#pragma unroll
                    for (int k = 1; k < D; ++k) S += qq[k];
                    const double e = exp_neg(S);
#pragma unroll
                    for (int g = 0; g < EG; ++g) {
                        const double w = Ws[(size_t)g * PT * PT + (r0 + r) * PT + j] * e;
                        accT[g] += w;
                        if (GRAD) {
#pragma unroll
                            for (int k = 0; k < D; ++k) {
                                acc1[g][k] = fma(w, q[k], acc1[g][k]);
                                acc2[g][k] = fma(w, qq[k], acc2[g][k]);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        stage ^= 1; I = In; J = Jn;
    }

    if (active) {
        constexpr int NA = 1 + 2 * D;
#pragma unroll
        for (int g = 0; g < EG; ++g) {
            double *dst = a.part + (((size_t)blockIdx.x * a.E + a.out_idx[g]) * NA) * a.Bpad + b;
            dst[0] = accT[g];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                dst[(size_t)(1 + k) * a.Bpad] = GRAD ? acc1[g][k] : 0.0;
                dst[(size_t)(1 + D + k) * a.Bpad] = GRAD ? acc2[g][k] : 0.0;
            }
        }
    }
}

}  // namespace gpmpc
