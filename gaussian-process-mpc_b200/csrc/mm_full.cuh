// Pair-sum kernels of the FULL-covariance moment-matching step (BASELINE config 4).
//
// For a Gaussian input N(u, S) with a full covariance S, the covariance of the GP outputs a, b is
// (src/tools/uncertainty_prop.py:187-236, the formula-correct NumPy form; a = b: :341-399)
//     Sigma_ab = delta_ab sf_a^2 - sf_a^2 sf_b^2 |R|^-1/2 T_ab - m_a m_b,
//     T_ab = sum_ij W^ab_ij exp(-1/2 v_ij^T (S + Lam_ab)^-1 v_ij),  v_ij = u - P_a x_i - P_b x_j,
// with Lam_ab = (Lam_a^-1 + Lam_b^-1)^-1, P_a = Lam_ab Lam_a^-1, P_b = Lam_ab Lam_b^-1 (P_a + P_b = I), R = S Lam_ab^-1 + I
// and the input-independent factor exp(-1/2 (x_i-x_j)^T (Lam_a+Lam_b)^-1 (x_i-x_j)) folded into the weights at fit time
// (W^aa = (Ky^-1 - beta beta^T) o c, W^ab = -beta_a beta_b^T o c; fullcov.cu).  This is the product-of-Gaussians form
// of the reference's expression: k_a(x,x_i) k_b(x,x_j) is itself a Gaussian in x centred at P_a x_i + P_b x_j.
//
// With S + Lam_ab = C C^T and Lt = C^-1 / sqrt(2) (lower triangular) the exponent is -|q|^2, q = Lt v
//     q_ij = z_i + z_j,   z_i = Lt u - (Lt P_a) x_i,   z_j = -(Lt P_b) x_j,
// so the pair loop is the same "sum of squares -> exp" chain as the variance-only kernel (mm_pairs.cuh); only the
// per-point transforms z_i, z_j are triangular matrix-vector products instead of scalings.
//
// All outputs that share Lam_ab and P_a (a "unit": a pair of lambda groups) share the exp: one launch accumulates NP
// pair-outputs.  When both outputs come from the same lambda group everything is symmetric in (i, j) and only the
// upper-triangular tiles are swept (weights 2 / 1 / 0 as in Wt); otherwise all n x n tiles.
//
// Mapping: lanes <-> 32 rollouts (a chunk), the 8 warps of a CTA <-> 4 rows each of the 32x32 tile.  The weights of
// the NP pair-outputs are warp-uniform broadcast reads from shared memory, filled by TMA bulk copies (tile-major
// storage, one contiguous 8 KB block per tile and pair-output), double buffered.  Work items = fixed contiguous
// ranges of the tile list drawn from a ticket counter; every item writes its own partial slot, warps and items are
// summed in a fixed order => bit-reproducible.
//
// The adjoint is reverse mode with recomputation: the forward kernel keeps only T_ab; given the adjoints gbar_ab of
// the T_ab of a step, the backward kernel sweeps the pairs again with the combined weight sum_ab gbar_ab W^ab_ij and
// accumulates N1 = sum w e q (D values) and N2 = sum w e q q^T (D(D+1)/2 values), from which
//     d/du = -2 Lt^T N1,   d/dS = 2 Lt^T N2 Lt - (Lt^T Lt) sum_ab gbar_ab T_ab        (fullcov.cu, bwd_finalize)
// (forward-accumulated second moments as in the variance-only kernel would need NP (1 + D + D(D+1)/2) = 210 live
// accumulators per thread at D = 5, NP = 10).
#pragma once
#include "mm_pairs.cuh"

namespace gpmpc {

constexpr int FULL_THREADS = 256;
constexpr int FULL_WARPS = FULL_THREADS / 32;
constexpr int FULL_ROWS = PT / FULL_WARPS;          // rows of a tile per warp (4)
constexpr int kFullNPMax = 10;                      // pair-outputs per launch: 4 outputs of one lambda group -> 10

__host__ __device__ constexpr int tri_count(int D) { return D * (D + 1) / 2; }
__host__ __device__ constexpr int tri_idx(int k, int l) { return k * (k + 1) / 2 + l; }      // l <= k
// per-(rollout, unit) constants written by prep_full_kernel: Lr | Lc | off | pref
__host__ __device__ constexpr int full_nc(int D) { return 2 * tri_count(D) + D + 1; }
// moments of the backward sweep: N1 [D] | N2 [tri]
__host__ __device__ constexpr int full_nbwd(int D) { return D + tri_count(D); }

struct FullPairArgs {
    const double *W[kFullNPMax];   // weights of the launch's pair-outputs (tile-major; upper tiles if sym, else all tiles)
    const double *X;               // [ld, D]
    const double *cst;             // [full_nc][Bpad] constants of this unit for the current step
    const double *gbar;            // backward: [NP][Bpad] adjoints of T of the launch's pair-outputs
    double *part;                  // [chunks][n_items][NV][32]   NV = NP (forward) or full_nbwd(D) (backward)
    int *counters;                 // [chunks] work-item tickets, zero before the launch
    int ld, ntile, B, Bpad, n_items, chunks, total_tiles, sym;
};

template <int D, int NP>
__host__ __device__ constexpr size_t full_stage_doubles() { return (size_t)NP * PT * PT + 2 * PT * D; }
// the tile stages double as the scratch of the per-item warp reduction ([8 warps][NV][32])
template <int D, int NP>
__host__ __device__ constexpr size_t full_buf_doubles()
{
    constexpr size_t stages = 2 * full_stage_doubles<D, NP>();
    constexpr size_t nv = (size_t)(NP > D + tri_count(D) ? NP : D + tri_count(D));
    constexpr size_t scratch = (size_t)FULL_WARPS * nv * 32;
    return stages > scratch ? stages : scratch;
}
template <int D, int NP>
__host__ __device__ constexpr size_t full_smem_bytes()
{
    return (full_buf_doubles<D, NP>() + 16 + (size_t)(2 * tri_count(D) + D) * 32) * sizeof(double);
}

template <int D, int NP, bool BWD>
__global__ void __launch_bounds__(FULL_THREADS, 1)
mm_full_pairs(const FullPairArgs a)
{
    constexpr int TR = tri_count(D);
    constexpr int NV = BWD ? D + TR : NP;
    constexpr size_t STAGE = full_stage_doubles<D, NP>();
    constexpr unsigned STAGE_BYTES = (unsigned)(STAGE * sizeof(double));
    extern __shared__ __align__(128) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int chunk = blockIdx.x % a.chunks;
    const int b = chunk * 32 + lane;
    const bool active = b < a.B;
    double *tab = smem + full_buf_doubles<D, NP>();
    double *cs = tab + 16;                               // [2 TR + D][32]: Lr | Lc | off, this lane's column
    if (tid < 16) tab[tid] = kExp2Tab[tid];
    for (int e = wid; e < 2 * TR + D; e += FULL_WARPS) cs[e * 32 + lane] = active ? a.cst[(size_t)e * a.Bpad + b] : 0.0;
    double gb[BWD ? NP : 1];
    if (BWD) {
#pragma unroll
        for (int p = 0; p < NP; ++p) gb[p] = active ? a.gbar[(size_t)p * a.Bpad + b] : 0.0;
    }
    __shared__ __align__(8) unsigned long long full[2];
    __shared__ int s_item;
    if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); }
    __syncthreads();
    unsigned uses0 = 0, uses1 = 0;
    auto tile_rc = [&](int t, int &I, int &J) {          // tile number -> (row tile, column tile)
        if (a.sym) {
            int rem = t, row = 0;
            while (rem >= a.ntile - row) { rem -= a.ntile - row; ++row; }
            I = row; J = row + rem;
        } else { I = t / a.ntile; J = t - I * a.ntile; }
    };
    auto issue = [&](int stage, int t, int ti, int tj) {
        if (tid == 0) {
            double *base = smem + (size_t)stage * STAGE;
            void *bar = &full[stage];
            mbar_expect_tx(bar, STAGE_BYTES);
#pragma unroll
            for (int p = 0; p < NP; ++p)
                bulk_load_1d(base + (size_t)p * PT * PT, a.W[p] + (size_t)t * PT * PT, PT * PT * sizeof(double), bar);
            double *xi = base + (size_t)NP * PT * PT;
            bulk_load_1d(xi, a.X + (size_t)ti * PT * D, PT * D * sizeof(double), bar);
            bulk_load_1d(xi + PT * D, a.X + (size_t)tj * PT * D, PT * D * sizeof(double), bar);
        }
    };
    auto wait_stage = [&](int stage) {
        if (stage == 0) { mbar_wait(&full[0], uses0 & 1); ++uses0; }
        else { mbar_wait(&full[1], uses1 & 1); ++uses1; }
    };

    for (;;) {
        __syncthreads();                                 // previous item's reduction has left the stage buffers
        if (tid == 0) {
            // the stage buffers were last written through the generic proxy (reduction scratch); order those writes
            // before the TMA (async proxy) writes of the next item
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            s_item = atomicAdd(&a.counters[chunk], 1);
        }
        __syncthreads();
        const int item = s_item;
        if (item >= a.n_items) break;

        double acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = 0.0;

        const int t_begin = (int)((long long)a.total_tiles * item / a.n_items);
        const int t_end = (int)((long long)a.total_tiles * (item + 1) / a.n_items);
        int I = 0, J = 0;
        tile_rc(t_begin, I, J);
        if (t_begin < t_end) issue(0, t_begin, I, J);
        int stage = 0;
        for (int t = t_begin; t < t_end; ++t) {
            int In = I, Jn = J + 1;
            if (Jn == a.ntile) { ++In; Jn = a.sym ? In : 0; }
            if (t + 1 < t_end) issue(stage ^ 1, t + 1, In, Jn);
            wait_stage(stage);
            const double *Ws = smem + (size_t)stage * STAGE;
            const double *xi = Ws + (size_t)NP * PT * PT;
            const double *xj = xi + PT * D;

            // this warp's FULL_ROWS rows: z_i = off - Lr x_i (lower-triangular product)
            double zi[FULL_ROWS][D];
#pragma unroll
            for (int r = 0; r < FULL_ROWS; ++r) {
                const double *x = xi + (wid * FULL_ROWS + r) * D;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    double s = cs[(2 * TR + k) * 32 + lane];
#pragma unroll
                    for (int l = 0; l <= k; ++l) s = fma(-cs[tri_idx(k, l) * 32 + lane], x[l], s);
                    zi[r][k] = s;
                }
            }
            double Lc[TR];
#pragma unroll
            for (int e = 0; e < TR; ++e) Lc[e] = cs[(TR + e) * 32 + lane];

#pragma unroll 1
            for (int j = 0; j < PT; j += 2) {
                double zj[2][D];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const double *x = xj + (j + c) * D;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        double s = -Lc[tri_idx(k, 0)] * x[0];
#pragma unroll
                        for (int l = 1; l <= k; ++l) s = fma(-Lc[tri_idx(k, l)], x[l], s);
                        zj[c][k] = s;
                    }
                }
#pragma unroll
                for (int r = 0; r < FULL_ROWS; ++r) {
                    const int row = wid * FULL_ROWS + r;
                    double2 w2[NP];
#pragma unroll
                    for (int p = 0; p < NP; ++p)
                        w2[p] = *reinterpret_cast<const double2 *>(Ws + (size_t)p * PT * PT + row * PT + j);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        double q[D], qq[D];
#pragma unroll
                        for (int k = 0; k < D; ++k) { q[k] = zi[r][k] + zj[c][k]; qq[k] = q[k] * q[k]; }
                        double S = qq[0];
                        if (D >= 4) {
                            double S2 = qq[2] + qq[3];
                            S += qq[1];
#pragma unroll
                            for (int k = 4; k < D; k += 2) { S += qq[k]; if (k + 1 < D) S2 += qq[k + 1]; }
                            S += S2;
                        } else {
#pragma unroll
                            for (int k = 1; k < D; ++k) S += qq[k];
                        }
                        const double e = exp_neg(S, tab);
                        if (!BWD) {
#pragma unroll
                            for (int p = 0; p < NP; ++p) acc[p] = fma(c ? w2[p].y : w2[p].x, e, acc[p]);
                        } else {
                            double wb = gb[0] * (c ? w2[0].y : w2[0].x);
#pragma unroll
                            for (int p = 1; p < NP; ++p) wb = fma(gb[p], c ? w2[p].y : w2[p].x, wb);
                            const double we = wb * e;
#pragma unroll
                            for (int k = 0; k < D; ++k) {
                                const double wq = we * q[k];
                                acc[k] += wq;
#pragma unroll
                                for (int l = 0; l <= k; ++l) acc[D + tri_idx(k, l)] = fma(wq, q[l], acc[D + tri_idx(k, l)]);
                            }
                        }
                    }
                }
            }
            __syncthreads();
            stage ^= 1; I = In; J = Jn;
        }

        // warps (rows) are summed in index order through the (now idle) stage buffers, then one partial per item
        double *scr = smem;                              // [FULL_WARPS][NV][32]
#pragma unroll
        for (int v = 0; v < NV; ++v) scr[((size_t)wid * NV + v) * 32 + lane] = acc[v];
        __syncthreads();
        double *dst = a.part + (((size_t)chunk * a.n_items + item) * NV) * 32 + lane;
        for (int v = wid; v < NV; v += FULL_WARPS) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < FULL_WARPS; ++w) s += scr[((size_t)w * NV + v) * 32 + lane];
            dst[(size_t)v * 32] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Mean sums with a full input covariance (src/tools/uncertainty_prop.py:324-338): per lambda group
//     M0_a = sum_j beta_aj exp(-|q_j|^2),   q_j = Lt (u - x_j),  Lt^T Lt = (S + Lam)^-1 / 2,
// forward; backward (reverse mode, adjoints gbar_a of M0_a): N1 = sum w e q, N2 = sum w e q q^T with w_j = sum_a gbar_a beta_aj.
// O(n) per rollout: lanes <-> rollouts, grid (rollout chunks of 128, FULL_MEAN_JP partitions of the training set).
// ---------------------------------------------------------------------------------------------------------------
constexpr int FULL_MEAN_JP = 32;
constexpr int FULL_MEAN_THREADS = 128;

struct FullMeanArgs {
    const double *X;               // [ld, D]
    const double *beta[kGroupMax]; // [ld] of the group's outputs
    int EG;
    const double *cst;             // [full_nc][Bpad] constants of this mean unit (Lr = Lt, off = Lt u)
    const double *gbar[kGroupMax]; // backward: [Bpad] adjoint of M0 of each of the group's outputs
    double *part;                  // [FULL_MEAN_JP][NV][Bpad]   NV = kGroupMax (forward) or full_nbwd(D) (backward)
    int ld, B, Bpad;
};

template <int D, bool BWD>
__global__ void __launch_bounds__(FULL_MEAN_THREADS) mean_full_kernel(const FullMeanArgs a)
{
    constexpr int TR = tri_count(D);
    constexpr int NV = BWD ? D + TR : kGroupMax;
    __shared__ double xs[64 * D];
    __shared__ double bs[kGroupMax][64];
    __shared__ double etab[16];
    const int tid = threadIdx.x;
    if (tid < 16) etab[tid] = kExp2Tab[tid];
    const int b = blockIdx.x * FULL_MEAN_THREADS + tid;
    const bool active = b < a.B;
    const int jp = blockIdx.y;
    const int per = (a.ld / 64 + FULL_MEAN_JP - 1) / FULL_MEAN_JP * 64;
    const int j_begin = jp * per, j_end = min(a.ld, j_begin + per);
    double L[TR], off[D], gb[kGroupMax];
#pragma unroll
    for (int e = 0; e < TR; ++e) L[e] = active ? a.cst[(size_t)e * a.Bpad + b] : 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) off[k] = active ? a.cst[(size_t)(2 * TR + k) * a.Bpad + b] : 0.0;
#pragma unroll
    for (int g = 0; g < kGroupMax; ++g) gb[g] = (BWD && active && g < a.EG) ? a.gbar[g][b] : 0.0;
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = 0.0;
    for (int j0 = j_begin; j0 < j_end; j0 += 64) {
        __syncthreads();
        for (int e = tid; e < 64 * D; e += FULL_MEAN_THREADS) xs[e] = a.X[(size_t)j0 * D + e];
        for (int e = tid; e < 64 * kGroupMax; e += FULL_MEAN_THREADS)
            bs[e / 64][e % 64] = (e / 64 < a.EG) ? a.beta[e / 64][j0 + e % 64] : 0.0;
        __syncthreads();
        for (int j = 0; j < 64; ++j) {
            double q[D], S = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                double s = off[k];
#pragma unroll
                for (int l = 0; l <= k; ++l) s = fma(-L[tri_idx(k, l)], xs[j * D + l], s);
                q[k] = s;
                S = fma(s, s, S);
            }
            const double e = exp_neg(S, etab);
            if (!BWD) {
#pragma unroll
                for (int g = 0; g < kGroupMax; ++g) acc[g] = fma(bs[g][j], e, acc[g]);
            } else {
                double wb = 0.0;
#pragma unroll
                for (int g = 0; g < kGroupMax; ++g) wb = fma(gb[g], bs[g][j], wb);
                const double we = wb * e;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const double wq = we * q[k];
                    acc[k] += wq;
#pragma unroll
                    for (int l = 0; l <= k; ++l) acc[D + tri_idx(k, l)] = fma(wq, q[l], acc[D + tri_idx(k, l)]);
                }
            }
        }
    }
    if (active) {
#pragma unroll
        for (int v = 0; v < NV; ++v) a.part[((size_t)jp * NV + v) * a.Bpad + b] = acc[v];
    }
}

}  // namespace gpmpc
