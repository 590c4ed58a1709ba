// Full-covariance moment-matched rollout, its risk-sensitive cost and the exact adjoint (BASELINE
// config 4).  The reference propagates variances only and leaves the cross-covariances as a TODO
// (src/dynamics.py:104-121,184); its cost already accepts a full Sigma (src/mpc.py:182-185) and the formula for the
// cross-covariance of two outputs is its NumPy `covariance_prop` (src/tools/uncertainty_prop.py:187-236).
//
// Per horizon step t, on the handle's stream (B rollouts in lock step, internal layout [row][Bpad]):
//   prep_full       (b, unit)  S_t = blockdiag(Sigma_{t-1}, fp32(1e-3) I), Cholesky of S_t + Lam_unit, the triangular
//                              transforms Lr, Lc, the offset Lt u and the determinant prefactor of every unit
//   mean_full       (b, j)     M0_a per lambda group                                     (mm_full.cuh)
//   mm_full_pairs   (b, i, j)  T_ab for the pair-outputs of every unit                   (mm_full.cuh)
//   reduce_rows                fixed-order sum of the work-item partials
//   finalize_full   (b)        mu_t, Sigma_t
// The backward sweep (t = H..1) is reverse mode with recomputation: seed (adjoints of T_ab / M0_a from the adjoints of
// mu_t, Sigma_t), the backward pair / mean kernels (N1, N2), reduce, bwd_finalize (adjoints of mu_{t-1}, Sigma_{t-1},
// gradient of action t-1).
//
// A "unit" is a set of outputs pairs that share the exponent: mean units = lambda groups; pair units = pairs of lambda
// groups (g <= h).  A unit's pair-outputs are evaluated by launches of 1, 2, 3, 4, 6 or 10 pair-outputs ("pieces").
#include "common.cuh"
#include "small_linalg.cuh"
#include "mm_full.cuh"
#include <cmath>

namespace gpmpc {

#define DECL_FULL(D) cudaError_t launch_full_pairs_D##D(int, bool, const FullPairArgs &, int, cudaStream_t); \
                     cudaError_t launch_full_mean_D##D(bool, const FullMeanArgs &, cudaStream_t);
DECL_FULL(2) DECL_FULL(3) DECL_FULL(4) DECL_FULL(5) DECL_FULL(6) DECL_FULL(7) DECL_FULL(8)
#undef DECL_FULL

static cudaError_t launch_full_pairs(int D, int NP, bool bwd, const FullPairArgs &a, int ctas, cudaStream_t st)
{
    switch (D) {
        case 2: return launch_full_pairs_D2(NP, bwd, a, ctas, st); case 3: return launch_full_pairs_D3(NP, bwd, a, ctas, st);
        case 4: return launch_full_pairs_D4(NP, bwd, a, ctas, st); case 5: return launch_full_pairs_D5(NP, bwd, a, ctas, st);
        case 6: return launch_full_pairs_D6(NP, bwd, a, ctas, st); case 7: return launch_full_pairs_D7(NP, bwd, a, ctas, st);
        case 8: return launch_full_pairs_D8(NP, bwd, a, ctas, st);
    }
    return cudaErrorInvalidValue;
}
static cudaError_t launch_full_mean(int D, bool bwd, const FullMeanArgs &a, cudaStream_t st)
{
    switch (D) {
        case 2: return launch_full_mean_D2(bwd, a, st); case 3: return launch_full_mean_D3(bwd, a, st);
        case 4: return launch_full_mean_D4(bwd, a, st); case 5: return launch_full_mean_D5(bwd, a, st);
        case 6: return launch_full_mean_D6(bwd, a, st); case 7: return launch_full_mean_D7(bwd, a, st);
        case 8: return launch_full_mean_D8(bwd, a, st);
    }
    return cudaErrorInvalidValue;
}

constexpr int kMaxPO = kMaxE * (kMaxE + 1) / 2;      // pair-outputs (a <= b)
constexpr int kMaxUnits = kMaxE + kMaxPO;            // mean units + pair units (worst case: all lambdas distinct)
constexpr int kMaxPieces = kMaxPO;

// Plan descriptors read by the small per-rollout kernels (device copy in h->fc_plan).
struct FcPlanDev {
    int D, E, m, n_mean, n_units, n_po, n_pieces;
    int out_unit[kMaxE];                 // mean unit (= lambda group) of each output
    double sf[kMaxE];
    int po_a[kMaxPO], po_b[kMaxPO], po_unit[kMaxPO];
    int piece_unit[kMaxPieces];
    int unit_is_mean[kMaxUnits];
    double lam[kMaxUnits][kMaxD];        // Lam_unit (mean: lambda_g; pair: lambda_g lambda_h / (lambda_g + lambda_h))
    double Pa[kMaxUnits][kMaxD], Pb[kMaxUnits][kMaxD];
};

struct FcPiece { int unit, po0, np; };
struct FcPlan {
    FcPlanDev d;
    const double *po_W[kMaxPO];
    bool unit_sym[kMaxUnits];
    int unit_g[kMaxUnits], unit_h[kMaxUnits];
    FcPiece pieces[kMaxPieces];
};

// ---------------------------------------------------------------------------------------------
// Cross-output weights (built lazily, once per fit / hyper-parameter change).
//   same lambda group, a < b:  W^ab_ij = -w(i,j) 1/2 (beta_ai beta_bj + beta_aj beta_bi) exp(-1/4 d^T Lam^-1 d),  upper tiles,
//                              w = 2 (j > i), 1 (j == i)     [the (i,j)-symmetric part is all a symmetric sweep sees]
//   different groups:          W^ab_ij = -beta_ai beta_bj exp(-1/2 d^T (Lam_a + Lam_b)^-1 d),  all tiles, d = x_i - x_j
// The sign makes every pair-output obey  Sigma_ab = delta_ab sf_a^2 - sf_a^2 sf_b^2 |R|^-1/2 T_ab - m_a m_b
// (for a == b the existing Wt = w (Ky^-1 - beta beta^T) exp(..) is used).
// ---------------------------------------------------------------------------------------------
struct CrossArg { double inv_lam_sum[kMaxD]; int sym; };
__global__ void cross_weights_kernel(const double *__restrict__ X, int n, int D, CrossArg ca, const double *__restrict__ ba,
                                     const double *__restrict__ bb, double *__restrict__ W, int ld)
{
    const int I = blockIdx.y, J = blockIdx.x;
    if (ca.sym && J < I) return;
    const int nt = ld / kPairTile;
    double *tile = W + (ca.sym ? wt_tile_index(I, J, nt) : (size_t)I * nt + J) * kPairTile * kPairTile;
    const int c = threadIdx.x;
    const int j = J * kPairTile + c;
    for (int r = threadIdx.y; r < kPairTile; r += blockDim.y) {
        const int i = I * kPairTile + r;
        double v = 0.0;
        if (i < n && j < n && (!ca.sym || j >= i)) {
            double q = 0.0;
            for (int k = 0; k < D; ++k) {
                const double d = X[(size_t)i * D + k] - X[(size_t)j * D + k];
                q = fma(d * d, ca.inv_lam_sum[k], q);
            }
            const double e = exp(-0.5 * q);
            if (ca.sym) v = -(j > i ? 2.0 : 1.0) * 0.5 * (ba[i] * bb[j] + ba[j] * bb[i]) * e;
            else v = -ba[i] * bb[j] * e;
        }
        tile[r * kPairTile + c] = v;
    }
}

static void split_pieces(int npo, int unit, int po0, FcPlan &p)
{
    static const int sizes[] = {10, 6, 4, 3, 2, 1};
    int done = 0;
    while (done < npo) {
        for (int s : sizes)
            if (s <= npo - done) { p.pieces[p.d.n_pieces++] = FcPiece{unit, po0 + done, s}; done += s; break; }
    }
}

// Builds the plan for the handle's current lambda groups and (re)builds the cross weights if they are stale.
static int build_plan(gpmpc_ctx *h, FcPlan &p)
{
    std::memset(&p, 0, sizeof p);
    const int D = h->D, E = h->E, ld = h->ld;
    const int G = (int)h->groups.size();
    FcPlanDev &d = p.d;
    d.D = D; d.E = E; d.m = h->m; d.n_mean = G;
    for (int a = 0; a < E; ++a) d.sf[a] = h->sf_prop[a];
    for (int g = 0; g < G; ++g) {
        d.unit_is_mean[g] = 1;
        const int lead = h->groups[g].outputs[0];
        for (int k = 0; k < D; ++k) { d.lam[g][k] = h->lam_prop[lead][k]; d.Pa[g][k] = 1.0; d.Pb[g][k] = 0.0; }
        for (int i = 0; i < h->groups[g].count; ++i) d.out_unit[h->groups[g].outputs[i]] = g;
    }
    int unit = G, po = 0;
    const size_t nt = (size_t)ld / kPairTile;
    size_t cross_doubles = 0;
    std::vector<size_t> po_off;                          // offset into Wx, or (size_t)-1 for a == b (Wt)
    for (int g = 0; g < G; ++g)
        for (int hh = g; hh < G; ++hh) {
            const LambdaGroup &gg = h->groups[g], &gh = h->groups[hh];
            const int la = gg.outputs[0], lb = gh.outputs[0];
            p.unit_g[unit] = g; p.unit_h[unit] = hh; p.unit_sym[unit] = (g == hh);
            for (int k = 0; k < D; ++k) {
                const double a = h->lam_prop[la][k], b = h->lam_prop[lb][k];
                d.lam[unit][k] = a * b / (a + b);
                d.Pa[unit][k] = b / (a + b);
                d.Pb[unit][k] = a / (a + b);
            }
            const int po0 = po;
            for (int i = 0; i < gg.count; ++i)
                for (int j = (g == hh ? i : 0); j < gh.count; ++j) {
                    d.po_a[po] = gg.outputs[i]; d.po_b[po] = gh.outputs[j]; d.po_unit[po] = unit;
                    if (g == hh && i == j) po_off.push_back((size_t)-1);
                    else { po_off.push_back(cross_doubles); cross_doubles += (g == hh) ? wt_doubles(ld) : nt * nt * kPairTile * kPairTile; }
                    ++po;
                }
            split_pieces(po - po0, unit, po0, p);
            ++unit;
        }
    d.n_units = unit; d.n_po = po;
    for (int i = 0; i < d.n_pieces; ++i) d.piece_unit[i] = p.pieces[i].unit;

    const bool rebuild = h->cross_epoch != h->weights_epoch;
    if (rebuild) GP_CUDA(h, h->Wx.reserve((cross_doubles + 8) * sizeof(double)));
    for (int q = 0; q < po; ++q) {
        const int a = d.po_a[q], b = d.po_b[q];
        if (po_off[q] == (size_t)-1) { p.po_W[q] = h->Wt.as<double>() + (size_t)a * wt_doubles(ld); continue; }
        double *W = h->Wx.as<double>() + po_off[q];
        p.po_W[q] = W;
        if (!rebuild) continue;
        CrossArg ca;
        ca.sym = p.unit_sym[d.po_unit[q]] ? 1 : 0;
        for (int k = 0; k < kMaxD; ++k) ca.inv_lam_sum[k] = k < D ? 1.0 / (h->lam_prop[a][k] + h->lam_prop[b][k]) : 0.0;
        dim3 blk(32, 8), grid((unsigned)nt, (unsigned)nt);
        cross_weights_kernel<<<grid, blk, 0, h->stream>>>(h->X.as<double>(), h->n, D, ca, h->beta.as<double>() + (size_t)a * ld,
                                                          h->beta.as<double>() + (size_t)b * ld, W, ld);
        GP_LAUNCH_CHECK(h);
    }
    h->cross_epoch = h->weights_epoch;
    GP_CUDA(h, h->fc_plan.reserve(sizeof(FcPlanDev)));
    // the plan is a few KB and changes only with the hyper-parameters; an ordered copy per call keeps it simple
    GP_CUDA(h, cudaMemcpyAsync(h->fc_plan.p, &p.d, sizeof(FcPlanDev), cudaMemcpyHostToDevice, h->stream));
    GP_CUDA(h, cudaStreamSynchronize(h->stream));       // p.d lives on the caller's stack
    return GPMPC_OK;
}

// ---------------------------------------------------------------------------------------------
// prep_full: one thread per (rollout, unit).
//   mode 0: u = [mu_{t-1}, a_{t-1}], S = blockdiag(Sigma_{t-1}, act_var I)     (src/dynamics.py:154-163 with a full Sigma)
//   mode 1: u, S supplied by the caller (rollout-major U[B,D], S[B,D,D])
// out cst[unit][full_nc][Bpad]: Lr = Lt P_a | Lc = Lt P_b | off = Lt u | pref = |S Lam_unit^-1 + I|^-1/2,
//   S + Lam_unit = C C^T,  Lt = C^-1 / sqrt(2)
// ---------------------------------------------------------------------------------------------
struct PrepArgs {
    const FcPlanDev *plan; int B, Bpad, mode;
    const double *mu_prev, *cov_prev, *act;      // mode 0: [E][Bpad], [E*E][Bpad], [m][Bpad]
    const double *Uext, *Sext;                   // mode 1
    double act_var;
    double *cst;
};
__global__ void prep_full_kernel(const PrepArgs a)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int unit = blockIdx.y;
    if (b >= a.B) return;
    const FcPlanDev &P = *a.plan;
    const int D = P.D, E = P.E, Bp = a.Bpad;
    double u[kMaxD], A[kMaxD * kMaxD], Li[kMaxD * kMaxD];
    if (a.mode == 0) {
        for (int k = 0; k < D; ++k) u[k] = k < E ? a.mu_prev[(size_t)k * Bp + b] : a.act[(size_t)(k - E) * Bp + b];
        for (int r = 0; r < D; ++r)
            for (int c = 0; c < D; ++c)
                A[r * D + c] = (r < E && c < E) ? a.cov_prev[(size_t)(r * E + c) * Bp + b] : ((r == c) ? a.act_var : 0.0);
    } else {
        for (int k = 0; k < D; ++k) u[k] = a.Uext[(size_t)b * D + k];
        // symmetrised: the formulas assume a symmetric covariance
        for (int r = 0; r < D; ++r)
            for (int c = 0; c < D; ++c)
                A[r * D + c] = 0.5 * (a.Sext[((size_t)b * D + r) * D + c] + a.Sext[((size_t)b * D + c) * D + r]);
    }
    for (int k = 0; k < D; ++k) A[k * D + k] += P.lam[unit][k];
    // Cholesky A = C C^T (lower, in place); a non-positive pivot gives NaN, which propagates to the outputs as data
    double pref = 1.0;
    for (int c = 0; c < D; ++c) {
        double s = A[c * D + c];
        for (int l = 0; l < c; ++l) s -= A[c * D + l] * A[c * D + l];
        const double cc = sqrt(s);
        A[c * D + c] = cc;
        pref *= sqrt(P.lam[unit][c]) / cc;
        for (int r = c + 1; r < D; ++r) {
            double v = A[r * D + c];
            for (int l = 0; l < c; ++l) v -= A[r * D + l] * A[c * D + l];
            A[r * D + c] = v / cc;
        }
    }
    // Lt = C^-1 / sqrt(2) (lower triangular, forward substitution column by column)
    const double rs2 = 0.70710678118654752440;
    for (int c = 0; c < D; ++c) {
        Li[c * D + c] = 1.0 / A[c * D + c];
        for (int r = c + 1; r < D; ++r) {
            double s = 0.0;
            for (int l = c; l < r; ++l) s += A[r * D + l] * Li[l * D + c];
            Li[r * D + c] = -s / A[r * D + r];
        }
    }
    const int TR = tri_count(D);
    double *out = a.cst + (size_t)unit * full_nc(D) * Bp + b;
    for (int k = 0; k < D; ++k) {
        double off = 0.0;
        for (int l = 0; l <= k; ++l) {
            const double lt = Li[k * D + l] * rs2;
            out[(size_t)tri_idx(k, l) * Bp] = lt * P.Pa[unit][l];
            out[(size_t)(TR + tri_idx(k, l)) * Bp] = lt * P.Pb[unit][l];
            off = fma(lt, u[l], off);
        }
        out[(size_t)(2 * TR + k) * Bp] = off;
    }
    out[(size_t)(2 * TR + D) * Bp] = pref;
}

// ---------------------------------------------------------------------------------------------
// reduce_rows: fixed-order sum of work-item partials.  grid (rollout chunks of 32, rows, descriptors), 256 threads =
// 8 interleaved slices of the item list x 32 rollouts; slices are summed in index order.
//   layout 0 (pair kernels): part[chunk][item][nv][32];  layout 1 (mean kernels): part[item * item_stride + v * Bpad + b]
// ---------------------------------------------------------------------------------------------
struct RedDesc { const double *part; double *out; int layout, n_items, nv, pad; long long item_stride; };
constexpr int kMaxRed = kMaxPieces + kMaxE;
struct RedArgs { RedDesc d[kMaxRed]; int B, Bpad; };
__global__ void __launch_bounds__(256) reduce_rows_kernel(const RedArgs a)
{
    __shared__ double red[8][32];
    const RedDesc &d = a.d[blockIdx.z];
    const int v = blockIdx.y;
    if (v >= d.nv) return;
    const int chunk = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = chunk * 32 + lane;
    const double *src;
    size_t stride;
    if (d.layout == 0) { src = d.part + (((size_t)chunk * d.n_items) * d.nv + v) * 32 + lane; stride = (size_t)d.nv * 32; }
    else { src = d.part + (size_t)v * a.Bpad + b; stride = (size_t)d.item_stride; }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int i = w;
    for (; i + 24 < d.n_items; i += 32) {
        s0 += src[(size_t)i * stride]; s1 += src[(size_t)(i + 8) * stride];
        s2 += src[(size_t)(i + 16) * stride]; s3 += src[(size_t)(i + 24) * stride];
    }
    for (; i < d.n_items; i += 8) s0 += src[(size_t)i * stride];
    red[w][lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (w == 0) {
        double s = 0.0;
        for (int q = 0; q < 8; ++q) s += red[q][lane];
        d.out[(size_t)v * a.Bpad + b] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// finalize_full: one thread per rollout.  raw rows: [n_po] T_ab, then [E] M0_a.
//   m_a = sf_a^2 pref_mean M0_a;  Sigma_ab = delta_ab sf_a^2 - sf_a^2 sf_b^2 pref_unit T_ab - m_a m_b
// ---------------------------------------------------------------------------------------------
__global__ void finalize_full_kernel(const FcPlanDev *plan, int B, int Bp, const double *__restrict__ raw,
                                     const double *__restrict__ cst, double *__restrict__ mu_t, double *__restrict__ cov_t)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const FcPlanDev &P = *plan;
    const int D = P.D, E = P.E, NC = full_nc(D), ipref = 2 * tri_count(D) + D;
    double m[kMaxE];
    for (int a = 0; a < E; ++a) {
        const double pref = cst[((size_t)P.out_unit[a] * NC + ipref) * Bp + b];
        m[a] = P.sf[a] * P.sf[a] * pref * raw[(size_t)(P.n_po + a) * Bp + b];
        mu_t[(size_t)a * Bp + b] = m[a];
    }
    for (int q = 0; q < P.n_po; ++q) {
        const int a = P.po_a[q], c = P.po_b[q];
        const double pref = cst[((size_t)P.po_unit[q] * NC + ipref) * Bp + b];
        const double sa2 = P.sf[a] * P.sf[a], sc2 = P.sf[c] * P.sf[c];
        const double v = (a == c ? sa2 : 0.0) - sa2 * sc2 * pref * raw[(size_t)q * Bp + b] - m[a] * m[c];
        cov_t[(size_t)(a * E + c) * Bp + b] = v;
        cov_t[(size_t)(c * E + a) * Bp + b] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Cost terms with a full Sigma_t (src/mpc.py:179-198) and their partials; one thread per (rollout, t).
//   c_t = 1/gamma log det(I + gamma Q Sigma_t) + e^T (Q^-1 + gamma Sigma_t)^-1 e,   e = mu_t - x_ref
//   d c_t / d mu = (G + G^T) e;   d c_t / d Sigma_ij = ((I + gamma Q Sigma)^-1 Q)_ji - gamma (G^T e)_i (G e)_j
// ---------------------------------------------------------------------------------------------
struct FcCostArgs {
    int B, Bpad, E, m, H, has_rd, want_grad;
    const double *mu, *cov, *Uint, *gamma, *last_u;
    double Q[kMaxE * kMaxE], Qi[kMaxE * kMaxE], R[kMaxD * kMaxD], Rd[kMaxD * kMaxD], xref[kMaxE], uref[kMaxD];
    double *cterm;            // [(H+1) + H][Bpad]: state terms, then action terms
    double *seed_mu, *seed_cov, *gact;
    double *cost;             // [B]
};
template <int E>
__device__ __forceinline__ void state_cost_full_t(const FcCostArgs &a, int b, int t, double gamma)
{
    const int Bp = a.Bpad;
    double Sg[E * E], M[E * E], Minv[E * E], Gm[E * E], G[E * E], e[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        e[r] = a.mu[((size_t)t * E + r) * Bp + b] - a.xref[r];
#pragma unroll
        for (int k = 0; k < E; ++k) Sg[r * E + k] = a.cov[((size_t)t * E * E + r * E + k) * Bp + b];
    }
#pragma unroll
    for (int r = 0; r < E; ++r)
#pragma unroll
        for (int k = 0; k < E; ++k) {
            double s = 0.0;
#pragma unroll
            for (int l = 0; l < E; ++l) s += a.Q[r * E + l] * Sg[l * E + k];
            M[r * E + k] = (r == k ? 1.0 : 0.0) + gamma * s;
            Gm[r * E + k] = a.Qi[r * E + k] + gamma * Sg[r * E + k];
        }
    const double det = lu_det_inv_t<E>(M, a.want_grad ? Minv : nullptr);
    lu_det_inv_t<E>(Gm, G);
    double cost = (1.0 / gamma) * log(det);                // log of the determinant: NaN if det < 0 (src/mpc.py:183)
    double Ge[E], Gte[E];
#pragma unroll
    for (int r = 0; r < E; ++r) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int k = 0; k < E; ++k) { s1 += G[r * E + k] * e[k]; s2 += G[k * E + r] * e[k]; }
        Ge[r] = s1; Gte[r] = s2;
    }
#pragma unroll
    for (int k = 0; k < E; ++k) cost += e[k] * Ge[k];
    a.cterm[(size_t)t * Bp + b] = cost;
    if (!a.want_grad) return;
#pragma unroll
    for (int i = 0; i < E; ++i) {
        a.seed_mu[((size_t)t * E + i) * Bp + b] = Ge[i] + Gte[i];
#pragma unroll
        for (int j = 0; j < E; ++j) {
            double mq = 0.0;
#pragma unroll
            for (int r = 0; r < E; ++r) mq += Minv[j * E + r] * a.Q[r * E + i];
            a.seed_cov[((size_t)t * E * E + i * E + j) * Bp + b] = mq - gamma * Gte[i] * Ge[j];
        }
    }
}
__global__ void cost_full_terms_kernel(const FcCostArgs a)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (b >= a.B) return;
    const double gamma = a.gamma[b];
    switch (a.E) {
        case 1: state_cost_full_t<1>(a, b, t, gamma); break; case 2: state_cost_full_t<2>(a, b, t, gamma); break;
        case 3: state_cost_full_t<3>(a, b, t, gamma); break; case 4: state_cost_full_t<4>(a, b, t, gamma); break;
        case 5: state_cost_full_t<5>(a, b, t, gamma); break; case 6: state_cost_full_t<6>(a, b, t, gamma); break;
        case 7: state_cost_full_t<7>(a, b, t, gamma); break; default: state_cost_full_t<8>(a, b, t, gamma); break;
    }
    if (t >= a.H) return;
    // direct cost of action t (src/mpc.py:188-198) and its gradient (u_t also appears in delta_{t+1})
    const int m = a.m, Bp = a.Bpad, H = a.H, j = t;
    double cost = 0.0, du[kMaxD], g[kMaxD];
    for (int k = 0; k < m; ++k) { g[k] = 0.0; du[k] = a.Uint[((size_t)j * m + k) * Bp + b] - a.uref[k]; }
    for (int r = 0; r < m; ++r)
        for (int k = 0; k < m; ++k) { cost += du[r] * a.R[r * m + k] * du[k]; g[r] += (a.R[r * m + k] + a.R[k * m + r]) * du[k]; }
    if (a.has_rd) {
        double d0[kMaxD], d1[kMaxD];
        for (int k = 0; k < m; ++k) {
            const double cur = a.Uint[((size_t)j * m + k) * Bp + b];
            const double prev = (j == 0) ? a.last_u[(size_t)k * Bp + b] : a.Uint[((size_t)(j - 1) * m + k) * Bp + b];
            d0[k] = cur - prev;
            d1[k] = (j + 1 < H) ? a.Uint[((size_t)(j + 1) * m + k) * Bp + b] - cur : 0.0;
        }
        for (int r = 0; r < m; ++r)
            for (int k = 0; k < m; ++k) { cost += d0[r] * a.Rd[r * m + k] * d0[k]; g[r] += (a.Rd[r * m + k] + a.Rd[k * m + r]) * (d0[k] - d1[k]); }
    }
    a.cterm[(size_t)(H + 1 + j) * Bp + b] = cost;
    if (a.want_grad) for (int k = 0; k < m; ++k) a.gact[((size_t)j * m + k) * Bp + b] = g[k];
}
// same summation order as the variance-only path: t = H..0, each state term followed by the action term of t-1
__global__ void cost_full_sum_kernel(int B, int Bp, int H, const double *__restrict__ cterm, double *__restrict__ cost)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double c = 0.0;
    for (int t = H; t >= 0; --t) { c += cterm[(size_t)t * Bp + b]; if (t > 0) c += cterm[(size_t)(H + t) * Bp + b]; }
    cost[b] = c;
}

// ---------------------------------------------------------------------------------------------
// Backward, step t: seed.  carry_mu / carry_cov hold the adjoints of mu_t, Sigma_t (everything downstream of step t plus
// the cost / caller seeds at t).  Writes the adjoints of the raw sums: gbar[n_po] (T_ab), gbar[n_po + a] (M0_a) and the
// per-unit scalars sum gbar * raw (needed for the derivative of the determinant prefactors).
// ---------------------------------------------------------------------------------------------
__global__ void bwd_seed_kernel(const FcPlanDev *plan, int B, int Bp, const double *__restrict__ carry_mu,
                                const double *__restrict__ carry_cov, const double *__restrict__ mu_t,
                                const double *__restrict__ raw, const double *__restrict__ cst, double *__restrict__ gbar,
                                double *__restrict__ uscal)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const FcPlanDev &P = *plan;
    const int D = P.D, E = P.E, NC = full_nc(D), ipref = 2 * tri_count(D) + D;
    double Sb[kMaxE * kMaxE], m[kMaxE], us[kMaxUnits];
    for (int i = 0; i < E * E; ++i) Sb[i] = carry_cov[(size_t)i * Bp + b];
    for (int a = 0; a < E; ++a) m[a] = mu_t[(size_t)a * Bp + b];
    for (int u = 0; u < P.n_units; ++u) us[u] = 0.0;
    for (int q = 0; q < P.n_po; ++q) {
        const int a = P.po_a[q], c = P.po_b[q], u = P.po_unit[q];
        const double pref = cst[((size_t)u * NC + ipref) * Bp + b];
        const double sb = (a == c) ? Sb[a * E + a] : Sb[a * E + c] + Sb[c * E + a];
        const double g = -P.sf[a] * P.sf[a] * P.sf[c] * P.sf[c] * pref * sb;
        gbar[(size_t)q * Bp + b] = g;
        us[u] += g * raw[(size_t)q * Bp + b];
    }
    for (int a = 0; a < E; ++a) {
        double mb = carry_mu[(size_t)a * Bp + b];
        for (int c = 0; c < E; ++c) mb -= (Sb[a * E + c] + Sb[c * E + a]) * m[c];
        const int u = P.out_unit[a];
        const double pref = cst[((size_t)u * NC + ipref) * Bp + b];
        const double g = P.sf[a] * P.sf[a] * pref * mb;
        gbar[(size_t)(P.n_po + a) * Bp + b] = g;
        us[u] += g * raw[(size_t)(P.n_po + a) * Bp + b];
    }
    for (int u = 0; u < P.n_units; ++u) uscal[(size_t)u * Bp + b] = us[u];
}

// ---------------------------------------------------------------------------------------------
// Backward, step t: finalize.  red rows: [n_pieces + n_mean][full_nbwd] = N1 | N2 of every pair piece, then of every
// mean unit.  With Lt = Lr + Lc (P_a + P_b = I) and F = sym(N2):
//   d/du = -2 Lt^T N1,   d/dS = 2 Lt^T F Lt - (Lt^T Lt) s_unit,   s_unit = sum gbar * raw
// (the second term is the derivative of |S Lam^-1 + I|^-1/2: d pref / dS = -1/2 pref (S + Lam)^-1 = -pref Lt^T Lt).
// Writes the adjoints of (mu_{t-1}, Sigma_{t-1}) into carry (adding the seeds of t-1) and the gradient of action t-1.
// ---------------------------------------------------------------------------------------------
__global__ void bwd_finalize_kernel(const FcPlanDev *plan, int B, int Bp, int tm1, const double *__restrict__ red,
                                    const double *__restrict__ cst, const double *__restrict__ uscal,
                                    const double *__restrict__ seed_mu, const double *__restrict__ seed_cov,
                                    const double *__restrict__ gact, double *__restrict__ carry_mu,
                                    double *__restrict__ carry_cov, double *__restrict__ gradint)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const FcPlanDev &P = *plan;
    const int D = P.D, E = P.E, m = P.m, TR = tri_count(D), NC = full_nc(D), NB = full_nbwd(D);
    double ub[kMaxD], Sb[kMaxD * kMaxD];
    for (int k = 0; k < D; ++k) ub[k] = 0.0;
    for (int i = 0; i < D * D; ++i) Sb[i] = 0.0;
    for (int u = 0; u < P.n_units; ++u) {
        double Lt[kMaxD * kMaxD], N1[kMaxD], F[kMaxD * kMaxD];
        const double *c = cst + (size_t)u * NC * Bp + b;
        for (int k = 0; k < D; ++k)
            for (int l = 0; l < D; ++l)
                Lt[k * D + l] = l <= k ? c[(size_t)tri_idx(k, l) * Bp] + c[(size_t)(TR + tri_idx(k, l)) * Bp] : 0.0;
        for (int k = 0; k < D; ++k) N1[k] = 0.0;
        for (int i = 0; i < D * D; ++i) F[i] = 0.0;
        // the unit's moments: its pair pieces in index order, or its mean row
        for (int r = 0; r < P.n_pieces + P.n_mean; ++r) {
            const int ru = r < P.n_pieces ? P.piece_unit[r] : r - P.n_pieces;
            if (ru != u) continue;
            const double *rr = red + (size_t)r * NB * Bp + b;
            for (int k = 0; k < D; ++k) {
                N1[k] += rr[(size_t)k * Bp];
                for (int l = 0; l <= k; ++l) {
                    const double v = rr[(size_t)(D + tri_idx(k, l)) * Bp];
                    F[k * D + l] += v;
                    if (l != k) F[l * D + k] += v;
                }
            }
        }
        const double su = uscal[(size_t)u * Bp + b];
        // ub += -2 Lt^T N1
        for (int l = 0; l < D; ++l) {
            double s = 0.0;
            for (int k = l; k < D; ++k) s += Lt[k * D + l] * N1[k];
            ub[l] -= 2.0 * s;
        }
        // Sb += Lt^T (2 F - su I) Lt
        double Tm[kMaxD * kMaxD];                          // (2 F - su I) Lt
        for (int k = 0; k < D; ++k)
            for (int l = 0; l < D; ++l) {
                double s = 0.0;
                for (int q = 0; q < D; ++q) s += (2.0 * F[k * D + q] - (k == q ? su : 0.0)) * Lt[q * D + l];
                Tm[k * D + l] = s;
            }
        for (int k = 0; k < D; ++k)
            for (int l = 0; l < D; ++l) {
                double s = 0.0;
                for (int q = 0; q < D; ++q) s += Lt[q * D + k] * Tm[q * D + l];
                Sb[k * D + l] += s;
            }
    }
    for (int k = 0; k < E; ++k) {
        carry_mu[(size_t)k * Bp + b] = ub[k] + (seed_mu ? seed_mu[((size_t)tm1 * E + k) * Bp + b] : 0.0);
        for (int l = 0; l < E; ++l)
            carry_cov[(size_t)(k * E + l) * Bp + b] = Sb[k * D + l] + (seed_cov ? seed_cov[((size_t)tm1 * E * E + k * E + l) * Bp + b] : 0.0);
    }
    for (int k = 0; k < m; ++k)
        gradint[((size_t)tm1 * m + k) * Bp + b] = ub[E + k] + (gact ? gact[((size_t)tm1 * m + k) * Bp + b] : 0.0);
}

// layout shuffles between the C-ABI layout [B, inner] and the internal [inner][Bpad]
__global__ void fc_to_internal_kernel(const double *__restrict__ src, int B, int Bpad, int inner, double *__restrict__ dst)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int e = blockIdx.y;
    if (b < B) dst[(size_t)e * Bpad + b] = src[(size_t)b * inner + e];
}
// dst[b][t][inner] <- src[t*inner + e][Bpad]
__global__ void fc_to_external_kernel(const double *__restrict__ src, int B, int Bpad, int inner, double *__restrict__ dst)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int e = blockIdx.y;
    if (b < B) dst[(size_t)b * inner + e] = src[(size_t)e * Bpad + b];
}
__global__ void fc_init_state_kernel(const double *__restrict__ x0int, int B, int Bpad, int E, double *__restrict__ mu,
                                     double *__restrict__ cov, double var0)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int a = 0; a < E; ++a) {
        mu[(size_t)a * Bpad + b] = x0int[(size_t)a * Bpad + b];
        for (int c = 0; c < E; ++c) cov[(size_t)(a * E + c) * Bpad + b] = (a == c) ? var0 : 0.0;
    }
}
__global__ void fc_zero_kernel(double *p, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0.0;
}

// =============================================================================================
// Host orchestration
// =============================================================================================
struct FcGeom { int B, Bp, chunks, ctas, n_items_sym, n_items_full; long long tiles_sym, tiles_full; size_t part_doubles; };

static FcGeom make_geom(gpmpc_ctx *h, const FcPlan &p, int B)
{
    FcGeom g;
    g.B = B; g.Bp = round_up(B, 32); g.chunks = g.Bp / 32;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const int per_chunk = sms / g.chunks > 0 ? sms / g.chunks : 1;          // one CTA per SM (smem limited)
    g.ctas = per_chunk * g.chunks;
    const long long nt = h->ld / kPairTile;
    g.tiles_sym = nt * (nt + 1) / 2; g.tiles_full = nt * nt;
    auto items = [&](long long tiles) { long long it = (long long)per_chunk * 8; return (int)(it > tiles ? tiles : it); };
    g.n_items_sym = items(g.tiles_sym); g.n_items_full = items(g.tiles_full);
    // partial buffer of one step: every pair piece (backward is the larger: full_nbwd rows), then the mean units
    const int D = p.d.D;
    const size_t nvmax = (size_t)(full_nbwd(D) > kFullNPMax ? full_nbwd(D) : kFullNPMax);
    g.part_doubles = (size_t)p.d.n_pieces * g.chunks * g.n_items_full * nvmax * 32 +
                     (size_t)p.d.n_mean * FULL_MEAN_JP * nvmax * g.Bp + 64;
    return g;
}

// pair + mean launches of one step (forward or backward) and the reduction into `out`:
//   forward : out rows [n_po] T_ab | [E] M0_a            (= the raw rows of this step)
//   backward: out rows [n_pieces + n_mean][full_nbwd]
static int run_unit_kernels(gpmpc_ctx *h, const FcPlan &p, const FcGeom &g, bool bwd, const double *cst_step,
                            const double *gbar, double *out)
{
    const FcPlanDev &d = p.d;
    const int D = d.D, NB = full_nbwd(D), NC = full_nc(D);
    const size_t nvmax = (size_t)(NB > kFullNPMax ? NB : kFullNPMax);
    int *tickets = h->tickets.as<int>();
    GP_CUDA(h, cudaMemsetAsync(tickets, 0, (size_t)d.n_pieces * g.chunks * sizeof(int), h->stream));
    RedArgs ra;
    std::memset(&ra, 0, sizeof ra);
    ra.B = g.B; ra.Bpad = g.Bp;
    int nred = 0, max_nv = 0;
    double *part = h->fc_part.as<double>();
    for (int i = 0; i < d.n_pieces; ++i) {
        const FcPiece &pc = p.pieces[i];
        const bool sym = p.unit_sym[pc.unit];
        FullPairArgs a;
        std::memset(&a, 0, sizeof a);
        for (int q = 0; q < pc.np; ++q) a.W[q] = p.po_W[pc.po0 + q];
        a.X = h->X.as<double>();
        a.cst = cst_step + (size_t)pc.unit * NC * g.Bp;
        a.gbar = bwd ? gbar + (size_t)pc.po0 * g.Bp : nullptr;
        a.part = part + (size_t)i * g.chunks * g.n_items_full * nvmax * 32;
        a.counters = tickets + (size_t)i * g.chunks;
        a.ld = h->ld; a.ntile = h->ld / kPairTile; a.B = g.B; a.Bpad = g.Bp;
        a.n_items = sym ? g.n_items_sym : g.n_items_full; a.chunks = g.chunks;
        a.total_tiles = (int)(sym ? g.tiles_sym : g.tiles_full); a.sym = sym ? 1 : 0;
        cudaError_t e = launch_full_pairs(D, pc.np, bwd, a, g.ctas, h->stream);
        h->launches++;
        if (e != cudaSuccess) return fail(h, GPMPC_ERR_CUDA, std::string("mm_full_pairs: ") + cudaGetErrorString(e));
        RedDesc &rd = ra.d[nred++];
        rd.part = a.part; rd.layout = 0; rd.n_items = a.n_items; rd.nv = bwd ? NB : pc.np;
        rd.out = bwd ? out + (size_t)i * NB * g.Bp : out + (size_t)pc.po0 * g.Bp;
        if (rd.nv > max_nv) max_nv = rd.nv;
    }
    double *mpart = part + (size_t)d.n_pieces * g.chunks * g.n_items_full * nvmax * 32;
    for (int u = 0; u < d.n_mean; ++u) {
        const LambdaGroup &grp = h->groups[u];
        FullMeanArgs a;
        std::memset(&a, 0, sizeof a);
        a.X = h->X.as<double>();
        for (int i = 0; i < kGroupMax; ++i) a.beta[i] = h->beta.as<double>() + (size_t)grp.outputs[i < grp.count ? i : 0] * h->ld;
        a.EG = grp.count;
        a.cst = cst_step + (size_t)u * NC * g.Bp;
        a.part = mpart + (size_t)u * FULL_MEAN_JP * nvmax * g.Bp;
        a.ld = h->ld; a.B = g.B; a.Bpad = g.Bp;
        if (!bwd) {
            cudaError_t e = launch_full_mean(D, false, a, h->stream);
            h->launches++;
            if (e != cudaSuccess) return fail(h, GPMPC_ERR_CUDA, std::string("mean_full: ") + cudaGetErrorString(e));
            // forward partials are [JP][kGroupMax][Bp]; row i belongs to output grp.outputs[i]
            for (int i = 0; i < grp.count; ++i) {
                RedDesc &rd = ra.d[nred++];
                rd.part = a.part + (size_t)i * g.Bp; rd.layout = 1; rd.n_items = FULL_MEAN_JP; rd.nv = 1;
                rd.item_stride = (long long)kGroupMax * g.Bp;
                rd.out = out + (size_t)(d.n_po + grp.outputs[i]) * g.Bp;
                if (1 > max_nv) max_nv = 1;
            }
        } else {
            for (int i = 0; i < kGroupMax; ++i) a.gbar[i] = gbar + (size_t)(d.n_po + grp.outputs[i < grp.count ? i : 0]) * g.Bp;
            cudaError_t e = launch_full_mean(D, true, a, h->stream);
            h->launches++;
            if (e != cudaSuccess) return fail(h, GPMPC_ERR_CUDA, std::string("mean_full bwd: ") + cudaGetErrorString(e));
            RedDesc &rd = ra.d[nred++];
            rd.part = a.part; rd.layout = 1; rd.n_items = FULL_MEAN_JP; rd.nv = NB;
            rd.item_stride = (long long)NB * g.Bp;
            rd.out = out + (size_t)(d.n_pieces + u) * NB * g.Bp;
            if (NB > max_nv) max_nv = NB;
        }
    }
    reduce_rows_kernel<<<dim3(g.chunks, max_nv, nred), 256, 0, h->stream>>>(ra);
    GP_LAUNCH_CHECK(h);
    return GPMPC_OK;
}

struct FcWork {
    FcPlan plan; FcGeom g; int H;
    double *mu, *cov, *cst, *raw, *gbar, *uscal, *seed_mu, *seed_cov, *gact, *cterm, *carry_mu, *carry_cov, *gradint, *red;
    double *x0int, *Uint, *luint, *ext;
};

static int fc_check(gpmpc_ctx *h, int B, int H)
{
    if (!h) return GPMPC_ERR_INVALID;
    if (!h->fitted || h->n <= 0) return fail(h, GPMPC_ERR_NOT_FIT, "no training data: call gpmpc_fit first");
    if (B <= 0 || H < 0) return fail(h, GPMPC_ERR_INVALID, "B must be > 0 and H >= 0");
    if (h->D < 2 || h->D > kMaxD) return fail(h, GPMPC_ERR_UNSUPPORTED, "full-covariance propagation needs 2 <= D <= 8");
    return GPMPC_OK;
}

static int fc_reserve(gpmpc_ctx *h, int B, int H, FcWork &w)
{
    int rc = build_plan(h, w.plan);
    if (rc) return rc;
    w.g = make_geom(h, w.plan, B);
    w.H = H;
    const FcPlanDev &d = w.plan.d;
    const size_t Bp = w.g.Bp;
    const int D = d.D, E = d.E, m = d.m, Hs = H > 0 ? H : 1, ms = m > 0 ? m : 1;
    const int NB = full_nbwd(D), NC = full_nc(D);
    GP_CUDA(h, h->tickets.reserve(((size_t)d.n_pieces * w.g.chunks + 8) * sizeof(int)));
    GP_CUDA(h, h->fc_mu.reserve((size_t)(H + 1) * E * Bp * sizeof(double)));
    GP_CUDA(h, h->fc_cov.reserve((size_t)(H + 1) * E * E * Bp * sizeof(double)));
    GP_CUDA(h, h->fc_cst.reserve((size_t)Hs * d.n_units * NC * Bp * sizeof(double)));
    GP_CUDA(h, h->fc_raw.reserve((size_t)Hs * (d.n_po + E) * Bp * sizeof(double)));
    GP_CUDA(h, h->fc_part.reserve(w.g.part_doubles * sizeof(double)));
    GP_CUDA(h, h->fc_red.reserve((size_t)(d.n_pieces + d.n_mean) * NB * Bp * sizeof(double)));
    GP_CUDA(h, h->fc_gbar.reserve((size_t)(d.n_po + E + d.n_units) * Bp * sizeof(double)));
    // seeds: seed_mu [(H+1) E] | seed_cov [(H+1) E E] | gact [H m] | cterm [2H + 1]
    GP_CUDA(h, h->fc_seed.reserve(((size_t)(H + 1) * E + (size_t)(H + 1) * E * E + (size_t)Hs * ms + 2 * Hs + 1) * Bp * sizeof(double)));
    // carry: carry_mu [E] | carry_cov [E E] | gradint [H m]
    GP_CUDA(h, h->fc_carry.reserve(((size_t)E + (size_t)E * E + (size_t)Hs * ms) * Bp * sizeof(double)));
    // io: x0int [E] | Uint [H m] | last_u [m] | external staging (largest export: covs [B (H+1) E E])
    const size_t ext = (size_t)B * (H + 1) * (E * E + E + ms) + 4 * (size_t)B + 256;
    GP_CUDA(h, h->fc_io.reserve((((size_t)E + (size_t)Hs * ms + ms) * Bp + 2 * ext) * sizeof(double)));
    w.mu = h->fc_mu.as<double>(); w.cov = h->fc_cov.as<double>(); w.cst = h->fc_cst.as<double>(); w.raw = h->fc_raw.as<double>();
    w.red = h->fc_red.as<double>();
    w.gbar = h->fc_gbar.as<double>(); w.uscal = w.gbar + (size_t)(d.n_po + E) * Bp;
    double *p = h->fc_seed.as<double>();
    w.seed_mu = p; p += (size_t)(H + 1) * E * Bp;
    w.seed_cov = p; p += (size_t)(H + 1) * E * E * Bp;
    w.gact = p; p += (size_t)Hs * ms * Bp;
    w.cterm = p;
    p = h->fc_carry.as<double>();
    w.carry_mu = p; p += (size_t)E * Bp;
    w.carry_cov = p; p += (size_t)E * E * Bp;
    w.gradint = p;
    p = h->fc_io.as<double>();
    w.x0int = p; p += (size_t)E * Bp;
    w.Uint = p; p += (size_t)Hs * ms * Bp;
    w.luint = p; p += (size_t)ms * Bp;
    w.ext = p;
    return GPMPC_OK;
}

// host-or-device input -> device pointer (staged through `scratch` at *off)
static int fc_stage(gpmpc_ctx *h, double *scratch, size_t &off, const double *src, size_t count, const double **dev)
{
    if (is_device_ptr(src)) { *dev = src; return GPMPC_OK; }
    double *dst = scratch + off;
    GP_CUDA(h, cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    off += (count + 31) / 32 * 32;
    *dev = dst;
    return GPMPC_OK;
}

static int fc_step_forward(gpmpc_ctx *h, FcWork &w, int t, const PrepArgs &pa_in)
{
    const FcPlanDev &d = w.plan.d;
    const size_t Bp = w.g.Bp;
    const int NC = full_nc(d.D);
    PrepArgs pa = pa_in;
    pa.plan = reinterpret_cast<const FcPlanDev *>(h->fc_plan.p); pa.B = w.g.B; pa.Bpad = w.g.Bp;
    pa.cst = w.cst + (size_t)(t - 1) * d.n_units * NC * Bp;
    prep_full_kernel<<<dim3((w.g.B + 127) / 128, d.n_units), 128, 0, h->stream>>>(pa);
    GP_LAUNCH_CHECK(h);
    double *raw = w.raw + (size_t)(t - 1) * (d.n_po + d.E) * Bp;
    int rc = run_unit_kernels(h, w.plan, w.g, false, pa.cst, nullptr, raw);
    if (rc) return rc;
    finalize_full_kernel<<<(w.g.B + 127) / 128, 128, 0, h->stream>>>(pa.plan, w.g.B, w.g.Bp, raw, pa.cst,
                                                                     w.mu + (size_t)t * d.E * Bp, w.cov + (size_t)t * d.E * d.E * Bp);
    GP_LAUNCH_CHECK(h);
    return GPMPC_OK;
}

// forward rollout from device inputs x0 [B,E], U [B,H,m] into w.mu / w.cov (+ the per-step constants and raw sums)
static int fc_forward(gpmpc_ctx *h, FcWork &w, const double *x0_dev, const double *U_dev)
{
    const FcPlanDev &d = w.plan.d;
    const int B = w.g.B, Bp = w.g.Bp, E = d.E, m = d.m, H = w.H;
    fc_to_internal_kernel<<<dim3((B + 127) / 128, E), 128, 0, h->stream>>>(x0_dev, B, Bp, E, w.x0int);
    GP_LAUNCH_CHECK(h);
    if (H > 0 && m > 0) {
        fc_to_internal_kernel<<<dim3((B + 127) / 128, H * m), 128, 0, h->stream>>>(U_dev, B, Bp, H * m, w.Uint);
        GP_LAUNCH_CHECK(h);
    }
    fc_init_state_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(w.x0int, B, Bp, E, w.mu, w.cov, 1e-3);   // src/dynamics.py:148
    GP_LAUNCH_CHECK(h);
    for (int t = 1; t <= H; ++t) {
        PrepArgs pa;
        std::memset(&pa, 0, sizeof pa);
        pa.mode = 0;
        pa.mu_prev = w.mu + (size_t)(t - 1) * E * Bp;
        pa.cov_prev = w.cov + (size_t)(t - 1) * E * E * Bp;
        pa.act = w.Uint + (size_t)(t - 1) * m * Bp;
        pa.act_var = (double)1e-3f;                    // fp32 eye in the action block, src/dynamics.py:162
        int rc = fc_step_forward(h, w, t, pa);
        if (rc) return rc;
    }
    return GPMPC_OK;
}

// reverse sweep: seeds (adjoints of mu_t / Sigma_t for every t, internal layout; either may be NULL) and the direct
// action gradient gact (may be NULL) -> w.gradint [H m][Bp] and w.carry_mu = d/dx0
static int fc_backward(gpmpc_ctx *h, FcWork &w, const double *seed_mu, const double *seed_cov, const double *gact)
{
    const FcPlanDev &d = w.plan.d;
    const int B = w.g.B, E = d.E, H = w.H;
    const size_t Bp = w.g.Bp;
    const int NC = full_nc(d.D);
    const FcPlanDev *plan = reinterpret_cast<const FcPlanDev *>(h->fc_plan.p);
    if (seed_mu) GP_CUDA(h, cudaMemcpyAsync(w.carry_mu, seed_mu + (size_t)H * E * Bp, (size_t)E * Bp * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    else GP_CUDA(h, cudaMemsetAsync(w.carry_mu, 0, (size_t)E * Bp * sizeof(double), h->stream));
    if (seed_cov) GP_CUDA(h, cudaMemcpyAsync(w.carry_cov, seed_cov + (size_t)H * E * E * Bp, (size_t)E * E * Bp * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    else GP_CUDA(h, cudaMemsetAsync(w.carry_cov, 0, (size_t)E * E * Bp * sizeof(double), h->stream));
    for (int t = H; t >= 1; --t) {
        const double *cst = w.cst + (size_t)(t - 1) * d.n_units * NC * Bp;
        const double *raw = w.raw + (size_t)(t - 1) * (d.n_po + E) * Bp;
        bwd_seed_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(plan, B, (int)Bp, w.carry_mu, w.carry_cov, w.mu + (size_t)t * E * Bp,
                                                              raw, cst, w.gbar, w.uscal);
        GP_LAUNCH_CHECK(h);
        int rc = run_unit_kernels(h, w.plan, w.g, true, cst, w.gbar, w.red);
        if (rc) return rc;
        bwd_finalize_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(plan, B, (int)Bp, t - 1, w.red, cst, w.uscal, seed_mu, seed_cov,
                                                                  gact, w.carry_mu, w.carry_cov, w.gradint);
        GP_LAUNCH_CHECK(h);
    }
    return GPMPC_OK;
}

// internal rows [inner][Bp] -> caller's [B, inner] (host or device)
static int fc_export(gpmpc_ctx *h, const FcWork &w, const double *src, int inner, double *out, double *scratch)
{
    if (!out || inner == 0) return GPMPC_OK;
    const bool host = !is_device_ptr(out);
    double *dev = host ? scratch : out;
    fc_to_external_kernel<<<dim3((w.g.B + 127) / 128, inner), 128, 0, h->stream>>>(src, w.g.B, w.g.Bp, inner, dev);
    GP_LAUNCH_CHECK(h);
    if (host) {
        GP_CUDA(h, cudaMemcpyAsync(out, dev, (size_t)w.g.B * inner * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        GP_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return GPMPC_OK;
}

static void fc_host_inverse(int E, const double *Q, double *Qi)
{
    double A[kMaxE * kMaxE];
    int piv[kMaxE];
    std::memcpy(A, Q, sizeof(double) * E * E);
    for (int c = 0; c < E; ++c) {
        int p = c;
        for (int r = c + 1; r < E; ++r) if (std::abs(A[r * E + c]) > std::abs(A[p * E + c])) p = r;
        piv[c] = p;
        if (p != c) for (int k = 0; k < E; ++k) std::swap(A[c * E + k], A[p * E + k]);
        const double dinv = 1.0 / A[c * E + c];
        for (int r = c + 1; r < E; ++r) {
            const double f = A[r * E + c] * dinv;
            A[r * E + c] = f;
            for (int k = c + 1; k < E; ++k) A[r * E + k] -= f * A[c * E + k];
        }
    }
    for (int col = 0; col < E; ++col) {
        double x[kMaxE];
        for (int r = 0; r < E; ++r) x[r] = (r == col) ? 1.0 : 0.0;
        for (int c = 0; c < E; ++c) if (piv[c] != c) std::swap(x[c], x[piv[c]]);
        for (int r = 0; r < E; ++r) for (int k = 0; k < r; ++k) x[r] -= A[r * E + k] * x[k];
        for (int r = E - 1; r >= 0; --r) {
            for (int k = r + 1; k < E; ++k) x[r] -= A[r * E + k] * x[k];
            x[r] /= A[r * E + r];
        }
        for (int r = 0; r < E; ++r) Qi[r * E + col] = x[r];
    }
}

static int fc_fetch(gpmpc_ctx *h, const double *src, double *dst, size_t cnt)
{
    if (!src) { std::memset(dst, 0, cnt * sizeof(double)); return GPMPC_OK; }
    if (is_device_ptr(src)) {
        GP_CUDA(h, cudaMemcpyAsync(dst, src, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        GP_CUDA(h, cudaStreamSynchronize(h->stream));
    } else std::memcpy(dst, src, cnt * sizeof(double));
    return GPMPC_OK;
}

}  // namespace gpmpc

using namespace gpmpc;

extern "C" int gpmpc_rollout_full(gpmpc_handle h, int B, int H, const double *x0, const double *U, double *means, double *covs)
{
    int rc = fc_check(h, B, H);
    if (rc) return rc;
    if (!x0 || (H > 0 && h->m > 0 && !U)) return fail(h, GPMPC_ERR_INVALID, "gpmpc_rollout_full: null input");
    GP_CUDA(h, cudaSetDevice(h->device));
    FcWork w;
    if ((rc = fc_reserve(h, B, H, w))) return rc;
    const int E = h->E, m = h->m;
    size_t off = 0;
    const double *x0d, *Ud = nullptr;
    if ((rc = fc_stage(h, w.ext, off, x0, (size_t)B * E, &x0d))) return rc;
    if (H > 0 && m > 0 && (rc = fc_stage(h, w.ext, off, U, (size_t)B * H * m, &Ud))) return rc;
    if ((rc = fc_forward(h, w, x0d, Ud))) return rc;
    h->fc_B = B; h->fc_H = H;
    double *scratch = w.ext + off;
    if ((rc = fc_export(h, w, w.mu, (H + 1) * E, means, scratch))) return rc;
    return fc_export(h, w, w.cov, (H + 1) * E * E, covs, scratch);
}

extern "C" int gpmpc_rollout_full_vjp(gpmpc_handle h, int B, int H, const double *gmeans, const double *gcovs, double *gU,
                                      double *gx0)
{
    int rc = fc_check(h, B, H);
    if (rc) return rc;
    if (h->fc_B != B || h->fc_H != H)
        return fail(h, GPMPC_ERR_INVALID, "gpmpc_rollout_full_vjp: no tape of this shape (call gpmpc_rollout_full first)");
    GP_CUDA(h, cudaSetDevice(h->device));
    FcWork w;
    if ((rc = fc_reserve(h, B, H, w))) return rc;     // same sizes as the forward call: buffers (and the tape) are kept
    const int E = h->E, m = h->m;
    const size_t Bp = w.g.Bp;
    size_t off = 0;
    const double *gmd = nullptr, *gcd = nullptr;
    if (gmeans && (rc = fc_stage(h, w.ext, off, gmeans, (size_t)B * (H + 1) * E, &gmd))) return rc;
    if (gcovs && (rc = fc_stage(h, w.ext, off, gcovs, (size_t)B * (H + 1) * E * E, &gcd))) return rc;
    if (gmd) {
        fc_to_internal_kernel<<<dim3((B + 127) / 128, (H + 1) * E), 128, 0, h->stream>>>(gmd, B, (int)Bp, (H + 1) * E, w.seed_mu);
        GP_LAUNCH_CHECK(h);
    }
    if (gcd) {
        fc_to_internal_kernel<<<dim3((B + 127) / 128, (H + 1) * E * E), 128, 0, h->stream>>>(gcd, B, (int)Bp, (H + 1) * E * E, w.seed_cov);
        GP_LAUNCH_CHECK(h);
    }
    if ((rc = fc_backward(h, w, gmd ? w.seed_mu : nullptr, gcd ? w.seed_cov : nullptr, nullptr))) return rc;
    double *scratch = w.ext + off;
    if (H > 0 && m > 0 && (rc = fc_export(h, w, w.gradint, H * m, gU, scratch))) return rc;
    if (gx0) {
        if (H == 0) {          // no step: d/dx0 is the seed at t = 0
            if (gmd) GP_CUDA(h, cudaMemcpyAsync(w.carry_mu, w.seed_mu, (size_t)E * Bp * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        }
        if ((rc = fc_export(h, w, w.carry_mu, E, gx0, scratch))) return rc;
    }
    return GPMPC_OK;
}

extern "C" int gpmpc_rollout_cost_grad_full(gpmpc_handle h, int B, int H, const double *x0, const double *U, const double *gamma,
                                            const double *Q, const double *R, const double *Rdelta, const double *last_u,
                                            const double *xref, const double *uref, double *cost, double *grad, double *means,
                                            double *covs)
{
    int rc = fc_check(h, B, H);
    if (rc) return rc;
    if (!x0 || !gamma || !Q || !cost || (H > 0 && h->m > 0 && (!U || !R)))
        return fail(h, GPMPC_ERR_INVALID, "gpmpc_rollout_cost_grad_full: null input");
    if (Rdelta && !last_u) return fail(h, GPMPC_ERR_INVALID, "gpmpc_rollout_cost_grad_full: Rdelta needs last_u");
    GP_CUDA(h, cudaSetDevice(h->device));
    FcWork w;
    if ((rc = fc_reserve(h, B, H, w))) return rc;
    const int E = h->E, m = h->m;
    const size_t Bp = w.g.Bp;
    size_t off = 0;
    const double *x0d, *Ud = nullptr, *gd, *lud = nullptr;
    if ((rc = fc_stage(h, w.ext, off, x0, (size_t)B * E, &x0d))) return rc;
    if (H > 0 && m > 0 && (rc = fc_stage(h, w.ext, off, U, (size_t)B * H * m, &Ud))) return rc;
    if ((rc = fc_stage(h, w.ext, off, gamma, (size_t)B, &gd))) return rc;
    if (Rdelta && (rc = fc_stage(h, w.ext, off, last_u, (size_t)B * m, &lud))) return rc;
    if ((rc = fc_forward(h, w, x0d, Ud))) return rc;
    h->fc_B = h->fc_H = 0;

    const bool want_grad = grad != nullptr;
    FcCostArgs ca;
    std::memset(&ca, 0, sizeof ca);
    if ((rc = fc_fetch(h, Q, ca.Q, (size_t)E * E))) return rc;
    if ((rc = fc_fetch(h, R, ca.R, (size_t)m * m))) return rc;
    if ((rc = fc_fetch(h, Rdelta, ca.Rd, (size_t)m * m))) return rc;
    if ((rc = fc_fetch(h, xref, ca.xref, E))) return rc;
    if ((rc = fc_fetch(h, uref, ca.uref, m))) return rc;
    fc_host_inverse(E, ca.Q, ca.Qi);                    // Q^-1, src/mpc.py:179
    ca.B = B; ca.Bpad = (int)Bp; ca.E = E; ca.m = m; ca.H = H; ca.has_rd = Rdelta ? 1 : 0; ca.want_grad = want_grad ? 1 : 0;
    ca.mu = w.mu; ca.cov = w.cov; ca.Uint = w.Uint; ca.gamma = gd;
    if (Rdelta && m > 0) {
        fc_to_internal_kernel<<<dim3((B + 127) / 128, m), 128, 0, h->stream>>>(lud, B, (int)Bp, m, w.luint);
        GP_LAUNCH_CHECK(h);
    }
    ca.last_u = w.luint;
    ca.cterm = w.cterm; ca.seed_mu = w.seed_mu; ca.seed_cov = w.seed_cov; ca.gact = w.gact;
    double *scratch = w.ext + off;
    double *cost_dev = scratch; scratch += (B + 31) / 32 * 32;
    const bool cost_host = !is_device_ptr(cost);
    cost_full_terms_kernel<<<dim3((B + 127) / 128, H + 1), 128, 0, h->stream>>>(ca);
    GP_LAUNCH_CHECK(h);
    cost_full_sum_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(B, (int)Bp, H, w.cterm, cost_host ? cost_dev : cost);
    GP_LAUNCH_CHECK(h);
    if (cost_host) GP_CUDA(h, cudaMemcpyAsync(cost, cost_dev, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (want_grad && H > 0 && m > 0) {
        if ((rc = fc_backward(h, w, w.seed_mu, w.seed_cov, w.gact))) return rc;
        if ((rc = fc_export(h, w, w.gradint, H * m, grad, scratch))) return rc;
    }
    if ((rc = fc_export(h, w, w.mu, (H + 1) * E, means, scratch))) return rc;
    if ((rc = fc_export(h, w, w.cov, (H + 1) * E * E, covs, scratch))) return rc;
    if (cost_host) GP_CUDA(h, cudaStreamSynchronize(h->stream));
    return GPMPC_OK;
}

// Batched moment matching for FULL input covariances: mean[B,E] and the full output covariance cov[B,E,E].
extern "C" int gpmpc_moment_match_cov(gpmpc_handle h, int B, const double *U, const double *S, double *mean, double *cov)
{
    int rc = fc_check(h, B, 1);
    if (rc) return rc;
    if (!U || !S || !mean || !cov) return fail(h, GPMPC_ERR_INVALID, "gpmpc_moment_match_cov: null argument");
    GP_CUDA(h, cudaSetDevice(h->device));
    FcWork w;
    if ((rc = fc_reserve(h, B, 1, w))) return rc;
    h->fc_B = h->fc_H = 0;
    const int D = h->D, E = h->E;
    const size_t Bp = w.g.Bp;
    // staging for U [B,D] and S [B,D,D] (may exceed the rollout staging area): use gbuf
    GP_CUDA(h, h->gbuf.reserve(((size_t)B * D + (size_t)B * D * D + 128) * sizeof(double)));
    size_t off = 0;
    const double *Ud, *Sd;
    if ((rc = fc_stage(h, h->gbuf.as<double>(), off, U, (size_t)B * D, &Ud))) return rc;
    if ((rc = fc_stage(h, h->gbuf.as<double>(), off, S, (size_t)B * D * D, &Sd))) return rc;
    PrepArgs pa;
    std::memset(&pa, 0, sizeof pa);
    pa.mode = 1; pa.Uext = Ud; pa.Sext = Sd;
    if ((rc = fc_step_forward(h, w, 1, pa))) return rc;      // results land in slot t = 1
    if ((rc = fc_export(h, w, w.mu + (size_t)E * Bp, E, mean, w.ext))) return rc;
    return fc_export(h, w, w.cov + (size_t)E * E * Bp, E * E, cov, w.ext);
}
