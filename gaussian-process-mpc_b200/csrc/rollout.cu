// Horizon rollout, risk-sensitive cost and the exact adjoint (reverse sweep) for B independent control
// sequences advanced in lock step.  Replaces Dynamics.forward_propagate_torch (src/dynamics.py:126-191),
// RiskSensitiveMPC.cost_torch (src/mpc.py:156-200) and the autograd replay behind
// RiskSensitiveMPC.gradient (src/mpc.py:231-255).
//
// Per horizon step t (1..H), on the handle's stream:
//   prep_step      (b, group)   input mean u=[mu_{t-1}, a_{t-1}], variances s=[var_{t-1}, fp32(1e-3)],
//                               scaled constants for the pair / mean kernels
//   mean_sums      (b, a, j)    M0 = sum_j beta_j l_j and its first/second moments    (uncertainty_prop.py:324-338)
//   mm_pairs_batch (b, a, i<=j) T and its first/second moments (mm_pairs.cuh)          (uncertainty_prop.py:372-399)
//   finalize_step  (b, a)       fixed-order reduction of the partials, mean_t, var_t and the closed-form
//                               partial derivatives d(mean,var)/d(u,s) written to the tape
// then one thread per rollout evaluates the cost and runs the reverse sweep over the tape.
//
// Internal layouts put the rollout index last (coalesced across the lanes that own rollouts).
#include "common.cuh"
#include "step_common.cuh"
#include "small_linalg.cuh"
#include "mm_pairs.cuh"
#include "mm_step_single.cuh"
#include "mm_rollout_single.cuh"

namespace gpmpc {

#define DECL_LAUNCH(D) cudaError_t launch_pairs_batch_D##D(int, int, int, const PairArgs &, dim3, cudaStream_t); \
                       cudaError_t launch_step_single_D##D(int, int, int, const SingleStepArgs &, dim3, cudaStream_t); \
                       cudaError_t launch_rollout_single_D##D(int, int, const RolloutSingleArgs &, dim3, cudaStream_t);
DECL_LAUNCH(2) DECL_LAUNCH(3) DECL_LAUNCH(4) DECL_LAUNCH(5) DECL_LAUNCH(6) DECL_LAUNCH(7) DECL_LAUNCH(8)
#undef DECL_LAUNCH

// (outputs in the group, grad_mode 0/1/2, state dimensions, ...): see the moment selection in mm_pairs.cuh
typedef cudaError_t (*pair_launch_fn)(int, int, int, const SingleStepArgs &, dim3, cudaStream_t);
typedef cudaError_t (*pair_tma_launch_fn)(int, int, int, const PairArgs &, dim3, cudaStream_t);
static pair_tma_launch_fn pair_launcher(int D)
{
    switch (D) {
        case 2: return launch_pairs_batch_D2; case 3: return launch_pairs_batch_D3;
        case 4: return launch_pairs_batch_D4; case 5: return launch_pairs_batch_D5;
        case 6: return launch_pairs_batch_D6; case 7: return launch_pairs_batch_D7;
        case 8: return launch_pairs_batch_D8;
    }
    return nullptr;
}

static pair_launch_fn single_launcher(int D)
{
    switch (D) {
        case 2: return launch_step_single_D2; case 3: return launch_step_single_D3;
        case 4: return launch_step_single_D4; case 5: return launch_step_single_D5;
        case 6: return launch_step_single_D6; case 7: return launch_step_single_D7;
        case 8: return launch_step_single_D8;
    }
    return nullptr;
}

typedef cudaError_t (*rollout_single_fn)(int, int, const RolloutSingleArgs &, dim3, cudaStream_t);
static rollout_single_fn rollout_single_launcher(int D)
{
    switch (D) {
        case 2: return launch_rollout_single_D2; case 3: return launch_rollout_single_D3;
        case 4: return launch_rollout_single_D4; case 5: return launch_rollout_single_D5;
        case 6: return launch_rollout_single_D6; case 7: return launch_rollout_single_D7;
        case 8: return launch_rollout_single_D8;
    }
    return nullptr;
}

// up to this many rollouts the whole horizon runs in one persistent cooperative launch (mm_rollout_single.cuh)
constexpr int kPersistMaxB = 1;

// below this many rollouts the lanes<->pairs kernel (one rollout per CTA column) replaces the lanes<->rollouts one:
// measured on B200 at n=4096 it costs 1.55 ms per rollout and evaluation, the batched kernel 145 ms per started
// chunk of 128 rollouts (174 ms in round 1, crossover 112), i.e. the crossover is at 94
constexpr int kSingleMaxB = 96;
constexpr int MEAN_JP = 64;          // partitions of the training set in the mean kernel (64 x rollout chunks CTAs)
constexpr int MEAN_THREADS = 128;


// ---------------------------------------------------------------------------------------------
// prep_step: one thread per (rollout, lambda-group)
//   us[(k)*Bpad + b]       u_k            k < D
//   us[(D+k)*Bpad + b]     s_k
//   cst[g][4D][Bpad]:      c_k = sqrt(a_k/8), c_k u_k, cm_k = sqrt(b_k/2), cm_k u_k
//     a_k = 1/(lam_k/2 + s_k)  (uncertainty_prop.py:376),  b_k = 1/(s_k + lam_k)  (:331)
// mu/var hold step t-1 as [(t-1)*E + a][Bpad]; actions come from Uint [(t-1)*m + k][Bpad].
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void prep_one(const StepDims &d, int b, int g, int k, double u, double s,
                                         const double *__restrict__ lam_group, double *__restrict__ us, double *__restrict__ cst)
{
    if (g == 0) {
        us[(size_t)k * d.Bpad + b] = u;
        us[(size_t)(d.D + k) * d.Bpad + b] = s;
    }
    const double lam = lam_group[g * d.D + k];
    const double a = 1.0 / (0.5 * lam + s);
    const double bb = 1.0 / (s + lam);
    const double c = sqrt(0.125 * a);
    const double cm = sqrt(0.5 * bb);
    double *cg = cst + (size_t)g * 4 * d.D * d.Bpad;
    cg[(size_t)k * d.Bpad + b] = c;
    cg[(size_t)(d.D + k) * d.Bpad + b] = c * u;
    cg[(size_t)(2 * d.D + k) * d.Bpad + b] = cm;
    cg[(size_t)(3 * d.D + k) * d.Bpad + b] = cm * u;
}
__global__ void prep_step_kernel(StepDims d, int t, const double *__restrict__ mu, const double *__restrict__ var,
                                 const double *__restrict__ Uint, const double *__restrict__ lam_group,
                                 double *__restrict__ us, double *__restrict__ cst, double act_var)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (b >= d.B) return;
    for (int k = 0; k < d.D; ++k) {
        double u, s;
        if (k < d.E) {
            u = mu[((size_t)(t - 1) * d.E + k) * d.Bpad + b];
            s = var[((size_t)(t - 1) * d.E + k) * d.Bpad + b];
        } else {
            u = Uint[((size_t)(t - 1) * d.m + (k - d.E)) * d.Bpad + b];
            s = act_var;
        }
        prep_one(d, b, g, k, u, s, lam_group, us, cst);
    }
}
// Few rollouts: everything that precedes the first step in ONE launch -- both layout shuffles (x0 [B,E], U [B,H,m] ->
// internal), the initial state (mean x0, variance var0) and the constants of step 1 for every lambda group.
// Block (32 rollouts, 4 parts): part 0 does the state and the constants, all parts share the H*m actions.
__global__ void begin_rollout_small_kernel(StepDims d, int H, const double *__restrict__ x0, const double *__restrict__ U,
                                           double *__restrict__ x0int, double *__restrict__ Uint, double *__restrict__ mu,
                                           double *__restrict__ var, double var0, const double *__restrict__ lam_group,
                                           double *__restrict__ us, double *__restrict__ cst, double act_var)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= d.B) return;
    const int hm = H * d.m;
    for (int i = threadIdx.y; i < hm; i += blockDim.y) Uint[(size_t)i * d.Bpad + b] = U[(size_t)b * hm + i];
    if (threadIdx.y != 0) return;
    for (int k = 0; k < d.D; ++k) {
        double u, s = act_var;
        if (k < d.E) {
            u = x0[(size_t)b * d.E + k]; s = var0;
            x0int[(size_t)k * d.Bpad + b] = u; mu[(size_t)k * d.Bpad + b] = u; var[(size_t)k * d.Bpad + b] = var0;
        } else if (H > 0) u = U[(size_t)b * hm + (k - d.E)];
        else continue;
        if (H > 0)
            for (int g = 0; g < d.G; ++g) prep_one(d, b, g, k, u, s, lam_group, us, cst);
    }
}

// ---------------------------------------------------------------------------------------------
// mean_sums: lanes <-> rollouts, grid (rollout chunks, MEAN_JP partitions of j, groups).
//   p_k = cm_k (u_k - x_jk),  l_j = exp(-sum p_k^2),  M0 += beta_j l_j, M1_k += beta_j l_j p_k, M2_k += .. p_k^2
// ---------------------------------------------------------------------------------------------
struct MeanArgs {
    const double *X; const double *beta[kGroupMax]; int out_idx[kGroupMax]; int EG;
    const double *cst; double *mpart; StepDims d;
    int k1, k2;                   // moments kept: M1_k for k >= k1, M2_k for k < k2 (same selection as the pair kernel)
};

template <int D>
__global__ void __launch_bounds__(MEAN_THREADS) mean_sums_kernel(const MeanArgs a)
{
    __shared__ double xs[64 * D];
    __shared__ double bs[kGroupMax][64];
    __shared__ double etab[16];
    const int tid = threadIdx.x;
    if (tid < 16) etab[tid] = kExp2Tab[tid];
    const int b = blockIdx.x * MEAN_THREADS + tid;
    const bool active = b < a.d.B;
    const int jp = blockIdx.y;
    const int per = (a.d.ld / 64 + MEAN_JP - 1) / MEAN_JP * 64;    // rows per partition (multiple of 64)
    const int j_begin = jp * per;
    const int j_end = min(a.d.ld, j_begin + per);

    double cm[D], cmu[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        cm[k] = active ? a.cst[(size_t)(2 * D + k) * a.d.Bpad + b] : 0.0;
        cmu[k] = active ? a.cst[(size_t)(3 * D + k) * a.d.Bpad + b] : 0.0;
    }
    double m0[kGroupMax], m1[kGroupMax][D], m2[kGroupMax][D];
#pragma unroll
    for (int g = 0; g < kGroupMax; ++g) {
        m0[g] = 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) m1[g][k] = m2[g][k] = 0.0;
    }
    for (int j0 = j_begin; j0 < j_end; j0 += 64) {
        __syncthreads();
        for (int e = tid; e < 64 * D; e += MEAN_THREADS) xs[e] = a.X[(size_t)j0 * D + e];
        for (int e = tid; e < 64 * a.EG; e += MEAN_THREADS) bs[e / 64][e % 64] = a.beta[e / 64][j0 + e % 64];
        __syncthreads();
        for (int j = 0; j < 64; ++j) {
            double p[D], pp[D], S = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) { p[k] = fma(-cm[k], xs[j * D + k], cmu[k]); pp[k] = p[k] * p[k]; S += pp[k]; }
            const double l = exp_neg(S, etab);
#pragma unroll
            for (int g = 0; g < kGroupMax; ++g) {
                if (g < a.EG) {
                    const double w = bs[g][j] * l;
                    m0[g] += w;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        if (k >= a.k1) m1[g][k] = fma(w, p[k], m1[g][k]);
                        if (k < a.k2) m2[g][k] = fma(w, pp[k], m2[g][k]);
                    }
                }
            }
        }
    }
    if (active) {
        constexpr int NA = 1 + 2 * D;
#pragma unroll
        for (int g = 0; g < kGroupMax; ++g) {
            if (g < a.EG) {
                double *dst = a.mpart + (((size_t)jp * a.d.E + a.out_idx[g]) * NA) * a.d.Bpad + b;
                dst[0] = m0[g];
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    dst[(size_t)(1 + k) * a.d.Bpad] = m1[g][k];
                    dst[(size_t)(1 + D + k) * a.d.Bpad] = m2[g][k];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// finalize_step: sums the partials of (rollout, output) in a fixed order (deterministic), applies the
// determinant prefactors and writes mean/var of step t plus the tape entry.
//   tape[((t-1)*E + a) * (2+4D) + e][Bpad]:  e = 0 mean, 1 var, 2.. dm/du, dm/ds, dv/du, dv/ds
// ---------------------------------------------------------------------------------------------
constexpr int FIN_WARPS = 16;
__host__ __device__ inline size_t finalize_smem_bytes(int D) { return (size_t)FIN_WARPS * 2 * nacc(D) * 32 * sizeof(double); }
template <int DD>
__global__ void __launch_bounds__(32 * FIN_WARPS)
finalize_step_kernel(StepDims d, int t, int P, const double *__restrict__ part,
                     const double *__restrict__ mpart, const double *__restrict__ us,
                     const double *__restrict__ hyp, double *__restrict__ mu,
                     double *__restrict__ var, double *__restrict__ tape, int want_grad)
{
    // block = 32 rollouts (lanes) x FIN_WARPS interleaved slices of the partial-sum list.  The list can be long
    // (1184 work items for up to three rollout chunks), so a thread sums ALL statistics of an item at once and FIN_U
    // items per round: FIN_U x NA independent loads in flight instead of 4 (measured at B = 128, where this kernel was
    // 2.4 % of a step: 110 us with one statistic at a time).  DD is a template parameter so that the values stay in
    // registers.  The summation order is fixed: per warp the items p = wid, wid + FIN_WARPS, ... in order, then the
    // warps in index order.
    extern __shared__ double red[];                  // [FIN_WARPS][2 * NA][32]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.x * 32 + lane;
    const int a = blockIdx.y;
    constexpr int D = DD, NA = 1 + 2 * D, NAX = NA;
    constexpr int FIN_U = NA <= 11 ? 3 : 2;          // items per round: FIN_U * NA values + NA sums within 128 registers
    const bool live = b < d.B;
    const size_t stride = (size_t)d.E * NA * d.Bpad;     // one work item
    double acc[NAX], accm[NAX];
#pragma unroll
    for (int e = 0; e < NAX; ++e) acc[e] = accm[e] = 0.0;
    if (live) {
        const double *src = part + (size_t)a * NA * d.Bpad + b;
        int p = wid;
        for (; p + (FIN_U - 1) * FIN_WARPS < P; p += FIN_U * FIN_WARPS) {
            double v[FIN_U][NAX];
#pragma unroll
            for (int u = 0; u < FIN_U; ++u)
#pragma unroll
                for (int e = 0; e < NAX; ++e) v[u][e] = src[(size_t)(p + u * FIN_WARPS) * stride + (size_t)e * d.Bpad];
#pragma unroll
            for (int u = 0; u < FIN_U; ++u)
#pragma unroll
                for (int e = 0; e < NAX; ++e) acc[e] += v[u][e];
        }
        for (; p < P; p += FIN_WARPS) {
#pragma unroll
            for (int e = 0; e < NAX; ++e)
                if (e < NA) acc[e] += src[(size_t)p * stride + (size_t)e * d.Bpad];
        }
        for (int q = wid; q < MEAN_JP; q += FIN_WARPS) {
            const double *ms = mpart + (((size_t)q * d.E + a) * NA) * d.Bpad + b;
#pragma unroll
            for (int e = 0; e < NAX; ++e)
                if (e < NA) accm[e] += ms[(size_t)e * d.Bpad];
        }
    }
#pragma unroll
    for (int e = 0; e < NAX; ++e)
        if (e < NA) {
            red[((size_t)wid * 2 * NA + e) * 32 + lane] = acc[e];
            red[((size_t)wid * 2 * NA + NA + e) * 32 + lane] = accm[e];
        }
    __syncthreads();
    if (wid != 0 || !live) return;
    double accN[1 + 2 * kMaxD], accM[1 + 2 * kMaxD];
    for (int e = 0; e < NA; ++e) {
        double s = 0.0, sm = 0.0;
        for (int w = 0; w < FIN_WARPS; ++w) {
            s += red[((size_t)w * 2 * NA + e) * 32 + lane];
            sm += red[((size_t)w * 2 * NA + NA + e) * 32 + lane];
        }
        accN[e] = s; accM[e] = sm;
    }
    finalize_math(d, t, a, b, accN, accM, us, hyp, mu, var, tape, want_grad);
}

// ---------------------------------------------------------------------------------------------
// Layout shuffles between the C-ABI layout ([B, ...] rollout-major) and the internal one ([...][Bpad]).
// ---------------------------------------------------------------------------------------------
__global__ void to_internal_kernel(const double *__restrict__ src, int B, int Bpad, int inner, double *__restrict__ dst)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int e = blockIdx.y;
    if (b < B) dst[(size_t)e * Bpad + b] = src[(size_t)b * inner + e];
}
__global__ void to_external_kernel(const double *__restrict__ src, int B, int Bpad, int inner, double *__restrict__ dst)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int e = blockIdx.y;
    if (b < B) dst[(size_t)b * inner + e] = src[(size_t)e * Bpad + b];
}
__global__ void poison_on_error_kernel(const int *__restrict__ error, double *__restrict__ mu, double *__restrict__ var, size_t n)
{
    if (*error == 0) return;
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) { mu[i] = nan(""); var[i] = nan(""); }
}
__global__ void init_state_kernel(const double *__restrict__ x0int, int B, int Bpad, int E, double *__restrict__ mu,
                                  double *__restrict__ var, double var0)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = blockIdx.y;
    if (b < B) { mu[(size_t)a * Bpad + b] = x0int[(size_t)a * Bpad + b]; var[(size_t)a * Bpad + b] = var0; }
}

// ---------------------------------------------------------------------------------------------
// cost_adjoint: one thread per rollout.
//  mode 0 (cost): cost of src/mpc.py:179-198 and, if grad != NULL, its exact gradient w.r.t. U by a reverse
//                 sweep over the tape; seeds are the cost partials.
//  mode 1 (vjp) : seeds are caller-supplied d loss/d mean_t, d loss/d var_t (internal layout), output is
//                 d loss / d U (and d loss / d x0).
// ---------------------------------------------------------------------------------------------
struct CostArgs {
    StepDims d; int H; int mode; int has_rd; int want_grad;
    const double *mu, *var, *tape, *Uint;       // internal layouts
    const double *gamma;                         // [B]
    const double *last_u;                        // [m][Bpad] internal (iff has_rd)
    const double *seed_mu, *seed_var;            // mode 1: [(t*E+a)][Bpad], may be NULL
    double Q[kMaxE * kMaxE], Qi[kMaxE * kMaxE], R[kMaxD * kMaxD], Rd[kMaxD * kMaxD], xref[kMaxE], uref[kMaxD];
    double *cost;                                // [B] (device, final layout)
    double *gradint;                             // [(t*m + k)][Bpad]
    double *gx0int;                              // [a][Bpad] or NULL
    double *grad_ext;                            // few-rollouts kernel only: if set, the gradient goes here as [B, H, m] instead
};

// State cost of time t and its partials (src/mpc.py:179-185):
//   c_t = 1/gamma log det(I + gamma Q Sigma_t) + e^T (Q^-1 + gamma Sigma_t)^-1 e,  e = mu_t - x_ref
// d c_t / d mu = (G + G^T) e,  d c_t / d sigma_k^2 = (M^-1 Q)_kk - gamma (G^T e)_k (G e)_k
// The two halves (the determinant term with M^-1 Q, the quadratic term with G) share nothing but Sigma_t: the
// few-rollouts kernel runs them in different warps.
template <int E>
__device__ __forceinline__ double state_cost_det_t(const CostArgs &a, int b, int t, double gamma, double *mq)
{
    const int Bp = a.d.Bpad;
    double M[E * E], Minv[E * E], sg[E];
#pragma unroll
    for (int k = 0; k < E; ++k) sg[k] = a.var[((size_t)t * E + k) * Bp + b];
#pragma unroll
    for (int r = 0; r < E; ++r)
#pragma unroll
        for (int k = 0; k < E; ++k) M[r * E + k] = (r == k ? 1.0 : 0.0) + gamma * a.Q[r * E + k] * sg[k];   // I + gamma Q Sigma
    const double det = lu_det_inv_t<E>(M, a.want_grad ? Minv : nullptr);
    if (a.want_grad) {
#pragma unroll
        for (int k = 0; k < E; ++k) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < E; ++r) s += Minv[k * E + r] * a.Q[r * E + k];
            mq[k] = s;
        }
    }
    return (1.0 / gamma) * log(det);                   // log of the determinant (NaN if det < 0), mpc.py:183
}
// Quadratic half: returns e^T G e; dmu = (G + G^T) e; gg_k = gamma (G^T e)_k (G e)_k.
template <int E>
__device__ __forceinline__ double state_cost_quad_t(const CostArgs &a, int b, int t, double gamma, double *dmu, double *gg)
{
    const int Bp = a.d.Bpad;
    double Gm[E * E], G[E * E], e[E], sg[E];
#pragma unroll
    for (int k = 0; k < E; ++k) {
        sg[k] = a.var[((size_t)t * E + k) * Bp + b];
        e[k] = a.mu[((size_t)t * E + k) * Bp + b] - a.xref[k];
    }
#pragma unroll
    for (int r = 0; r < E; ++r)
#pragma unroll
        for (int k = 0; k < E; ++k) Gm[r * E + k] = a.Qi[r * E + k] + (r == k ? gamma * sg[k] : 0.0);      // Q^-1 + gamma Sigma
    lu_det_inv_t<E>(Gm, G);
    double quad = 0.0;
#pragma unroll
    for (int r = 0; r < E; ++r) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int k = 0; k < E; ++k) { s1 += G[r * E + k] * e[k]; s2 += G[k * E + r] * e[k]; }
        quad += e[r] * s1;
        if (a.want_grad) { dmu[r] = s1 + s2; gg[r] = gamma * s2 * s1; }
    }
    return quad;
}
#define GPMPC_SWITCH_E(EXPR)                                                                                      \
    switch (a.d.E) {                                                                                              \
        case 1: { constexpr int EE = 1; return EXPR; } case 2: { constexpr int EE = 2; return EXPR; }             \
        case 3: { constexpr int EE = 3; return EXPR; } case 4: { constexpr int EE = 4; return EXPR; }             \
        case 5: { constexpr int EE = 5; return EXPR; } case 6: { constexpr int EE = 6; return EXPR; }             \
        case 7: { constexpr int EE = 7; return EXPR; } default: { constexpr int EE = 8; return EXPR; }            \
    }
__device__ __noinline__ double state_cost_det(const CostArgs &a, int b, int t, double gamma, double *mq)
{
    GPMPC_SWITCH_E(state_cost_det_t<EE>(a, b, t, gamma, mq))
}
__device__ __noinline__ double state_cost_quad(const CostArgs &a, int b, int t, double gamma, double *dmu, double *gg)
{
    GPMPC_SWITCH_E(state_cost_quad_t<EE>(a, b, t, gamma, dmu, gg))
}
#undef GPMPC_SWITCH_E
__device__ __forceinline__ double state_cost_terms(const CostArgs &a, int b, int t, double gamma, double *dmu, double *dvar)
{
    double gg[kMaxE];
    const double c = state_cost_det(a, b, t, gamma, dvar) + state_cost_quad(a, b, t, gamma, dmu, gg);
    if (a.want_grad)
        for (int k = 0; k < a.d.E; ++k) dvar[k] -= gg[k];
    return c;
}

// Direct cost of action j (src/mpc.py:188-198): (u_j-u_ref)^T R (u_j-u_ref) + delta_j^T Rd delta_j, and the
// gradient w.r.t. u_j (u_j also appears in delta_{j+1}).
__device__ double action_cost_terms(const CostArgs &a, int b, int j, double *gact)
{
    const int m = a.d.m, Bp = a.d.Bpad, H = a.H;
    double cost = 0.0, du[kMaxD];
    for (int k = 0; k < m; ++k) { gact[k] = 0.0; du[k] = a.Uint[((size_t)j * m + k) * Bp + b] - a.uref[k]; }
    for (int r = 0; r < m; ++r)
        for (int k = 0; k < m; ++k) {
            cost += du[r] * a.R[r * m + k] * du[k];
            gact[r] += (a.R[r * m + k] + a.R[k * m + r]) * du[k];
        }
    if (a.has_rd) {
        double d0[kMaxD], d1[kMaxD];
        for (int k = 0; k < m; ++k) {
            const double cur = a.Uint[((size_t)j * m + k) * Bp + b];
            const double prev = (j == 0) ? a.last_u[(size_t)k * Bp + b] : a.Uint[((size_t)(j - 1) * m + k) * Bp + b];
            d0[k] = cur - prev;
            d1[k] = (j + 1 < H) ? a.Uint[((size_t)(j + 1) * m + k) * Bp + b] - cur : 0.0;
        }
        for (int r = 0; r < m; ++r)
            for (int k = 0; k < m; ++k) {
                cost += d0[r] * a.Rd[r * m + k] * d0[k];
                gact[r] += (a.Rd[r * m + k] + a.Rd[k * m + r]) * (d0[k] - d1[k]);
            }
    }
    return cost;
}

__global__ void __launch_bounds__(128) cost_adjoint_kernel(const CostArgs a)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.d.B) return;
    const int E = a.d.E, D = a.d.D, m = a.d.m, H = a.H, Bp = a.d.Bpad;
    const int NT = 2 + 4 * D;
    const double gamma = (a.mode == 0) ? a.gamma[b] : 0.0;
    double cost = 0.0;
    double mb[kMaxE], vb[kMaxE];               // adjoints of mean_t / var_t carried backwards
    for (int k = 0; k < E; ++k) mb[k] = vb[k] = 0.0;

    for (int t = H; t >= 0; --t) {
        // ---- seeds at time t ----
        if (a.mode == 0) {
            double dmu[kMaxE], dvar[kMaxE];
            cost += state_cost_terms(a, b, t, gamma, dmu, dvar);
            if (a.want_grad) for (int k = 0; k < E; ++k) { mb[k] += dmu[k]; vb[k] += dvar[k]; }
        } else {
            for (int k = 0; k < E; ++k) {
                if (a.seed_mu) mb[k] += a.seed_mu[((size_t)t * E + k) * Bp + b];
                if (a.seed_var) vb[k] += a.seed_var[((size_t)t * E + k) * Bp + b];
            }
        }
        if (t == 0) break;
        // ---- direct action cost of u_{t-1} ----
        double gact[kMaxD];
        for (int k = 0; k < m; ++k) gact[k] = 0.0;
        if (a.mode == 0) cost += action_cost_terms(a, b, t - 1, gact);
        if (!a.want_grad) continue;
        // ---- pull the adjoints of (mean_t, var_t) back through step t ----
        double ub[kMaxD], sb[kMaxD];
        for (int k = 0; k < D; ++k) ub[k] = sb[k] = 0.0;
        for (int o = 0; o < E; ++o) {
            const double *tp = a.tape + (((size_t)(t - 1) * E + o) * NT) * Bp + b;
            const double mo = mb[o], vo = vb[o];
            for (int k = 0; k < D; ++k) {
                ub[k] += mo * tp[(size_t)(2 + k) * Bp] + vo * tp[(size_t)(2 + 2 * D + k) * Bp];
                sb[k] += mo * tp[(size_t)(2 + D + k) * Bp] + vo * tp[(size_t)(2 + 3 * D + k) * Bp];
            }
        }
        for (int k = 0; k < E; ++k) { mb[k] = ub[k]; vb[k] = sb[k]; }
        for (int k = 0; k < m; ++k) a.gradint[((size_t)(t - 1) * m + k) * Bp + b] = ub[E + k] + gact[k];
    }
    if (a.mode == 0) a.cost[b] = cost;
    if (a.gx0int) for (int k = 0; k < E; ++k) a.gx0int[(size_t)k * Bp + b] = mb[k];
}

// Few rollouts: one block of 128 threads per rollout.
//   phase 1  the tape entries of this rollout go to shared memory as asynchronous 8-byte copies (all in flight at
//            once); meanwhile the H+1 state-cost terms are evaluated in parallel over t, the determinant half and
//            the quadratic half of a term in different warps, and the H action terms in the remaining warps;
//   phase 2  warp 1 adds up the cost in the order of the one-thread kernel; warp 0 runs the sequential reverse
//            sweep with lane (o mod 4, k) holding the partials of output o w.r.t. input dimension k in registers:
//            per step two FMAs, a two-level shuffle reduction over o and one shuffle that hands the new adjoints
//            back, with the partials of the next step already loaded -- ~100 cycles per horizon step.
// Measured (n=4096, H=30, B=1, ncu): 57 us for the previous version (tape loads one latency per 128 entries, all
// cost terms in one warp, sweep through shared memory with run-time loop bounds), see profiles/r02_summary.md.
__device__ __forceinline__ void cp_async8(double *dst_smem, const double *src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__global__ void __launch_bounds__(128) cost_adjoint_small_kernel(const CostArgs a)
{
    extern __shared__ double sm[];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int E = a.d.E, D = a.d.D, m = a.d.m, H = a.H, Bp = a.d.Bpad;
    const int NT = 2 + 4 * D;
    double *seed_mu = sm;                          // [(H+1) * E]
    double *seed_var = seed_mu + (size_t)(H + 1) * E;
    double *ggs = seed_var + (size_t)(H + 1) * E;  // [(H+1) * E]: gamma (G^T e)(G e), subtracted from seed_var after the barrier
    double *gact = ggs + (size_t)(H + 1) * E;      // [H * m]
    double *cpart = gact + (size_t)H * (m > 0 ? m : 1);   // [3H + 2]: determinant halves, quadratic halves, action terms
    double *tps = cpart + 3 * H + 2;               // [H * E * 4D]: the partial derivatives of this rollout's tape
    const double gamma = (a.mode == 0) ? a.gamma[b] : 0.0;
    if (a.want_grad)
        for (int i = tid; i < H * E * 4 * D; i += blockDim.x) {
            const int e = i % (4 * D), to = i / (4 * D);
            cp_async8(tps + i, a.tape + ((size_t)to * NT + 2 + e) * Bp + b);
        }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (a.mode == 0) {
        if (warp < 2) {                            // warp 0: determinant halves, warp 1: quadratic halves, lane <-> t
            for (int t = lane; t <= H; t += 32) {
                double o1[kMaxE], o2[kMaxE];
                if (warp == 0) {
                    cpart[t] = state_cost_det(a, b, t, gamma, o1);
                    if (a.want_grad) for (int k = 0; k < E; ++k) seed_var[t * E + k] = o1[k];
                } else {
                    cpart[H + 1 + t] = state_cost_quad(a, b, t, gamma, o1, o2);
                    if (a.want_grad) for (int k = 0; k < E; ++k) { seed_mu[t * E + k] = o1[k]; ggs[t * E + k] = o2[k]; }
                }
            }
        } else {
            for (int j = tid - 64; j < H; j += 64) {
                double g[kMaxD];
                cpart[2 * H + 2 + j] = action_cost_terms(a, b, j, g);
                for (int k = 0; k < m; ++k) gact[j * m + k] = g[k];
            }
        }
    } else {
        for (int i = tid; i < (H + 1) * E; i += blockDim.x) {
            seed_mu[i] = a.seed_mu ? a.seed_mu[(size_t)i * Bp + b] : 0.0;
            seed_var[i] = a.seed_var ? a.seed_var[(size_t)i * Bp + b] : 0.0;
            ggs[i] = 0.0;
        }
        for (int i = tid; i < H * m; i += blockDim.x) gact[i] = 0.0;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (warp == 1 && lane == 0 && a.mode == 0) {
        // same order as the one-thread kernel: t = H..0, each state term followed by the action term of t-1
        double c = 0.0;
        for (int t = H; t >= 0; --t) { c += cpart[t] + cpart[H + 1 + t]; if (t > 0) c += cpart[2 * H + 2 + t - 1]; }
        a.cost[b] = c;
    }
    if (warp != 0 || !a.want_grad) return;
    // ---- reverse sweep: lane = og * 8 + k, og = output mod 4, k = input dimension ----
    const int og = lane >> 3, k = lane & 7;
    const bool two = E > 4;                        // outputs og and og + 4
    const bool act0 = k < D && og < E, act1 = k < D && og + 4 < E;
    const unsigned full = 0xffffffffu;
    double p0[4], p1[4], s0[2], s1[2];             // partials dm/du, dm/ds, dv/du, dv/ds and the two seeds of the current step
    auto load = [&](int t, double *q0, double *q1, double *z0, double *z1) {
        const double *tp0 = tps + ((size_t)(t - 1) * E + og) * 4 * D, *tp1 = tp0 + (size_t)4 * 4 * D;
#pragma unroll
        for (int i = 0; i < 4; ++i) { q0[i] = act0 ? tp0[i * D + k] : 0.0; q1[i] = act1 ? tp1[i * D + k] : 0.0; }
        z0[0] = og < E ? seed_mu[t * E + og] : 0.0;       z0[1] = og < E ? seed_var[t * E + og] - ggs[t * E + og] : 0.0;
        z1[0] = og + 4 < E ? seed_mu[t * E + og + 4] : 0.0; z1[1] = og + 4 < E ? seed_var[t * E + og + 4] - ggs[t * E + og + 4] : 0.0;
    };
    if (H >= 1) load(H, p0, p1, s0, s1);
    double ub = 0.0, sb = 0.0;                     // adjoints of (u, s)_k entering step t (all lanes of a column hold the total)
    for (int t = H; t >= 1; --t) {
        double n0[4] = {0, 0, 0, 0}, n1[4] = {0, 0, 0, 0}, z0[2] = {0, 0}, z1[2] = {0, 0};
        if (t > 1) load(t - 1, n0, n1, z0, z1);
        // adjoints of (mean_t, var_t) of my outputs: what the previous step handed back plus this step's seeds
        double mo = __shfl_sync(full, ub, og) + s0[0], vo = __shfl_sync(full, sb, og) + s0[1];
        if (!act0) mo = vo = 0.0;                  // idle lanes read an action column: keep its value out of 0 * x
        double u = mo * p0[0] + vo * p0[2], v = mo * p0[1] + vo * p0[3];
        if (two) {
            mo = __shfl_sync(full, ub, og + 4) + s1[0]; vo = __shfl_sync(full, sb, og + 4) + s1[1];
            if (!act1) mo = vo = 0.0;
            u += mo * p1[0] + vo * p1[2]; v += mo * p1[1] + vo * p1[3];
        }
        u += __shfl_xor_sync(full, u, 8);  v += __shfl_xor_sync(full, v, 8);
        u += __shfl_xor_sync(full, u, 16); v += __shfl_xor_sync(full, v, 16);
        ub = u; sb = v;
        if (og == 0 && k >= E && k < D) {
            const double gk = ub + gact[(t - 1) * m + (k - E)];
            if (a.grad_ext) a.grad_ext[((size_t)b * H + (t - 1)) * m + (k - E)] = gk;
            else a.gradint[((size_t)(t - 1) * m + (k - E)) * Bp + b] = gk;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { p0[i] = n0[i]; p1[i] = n1[i]; }
        s0[0] = z0[0]; s0[1] = z0[1]; s1[0] = z1[0]; s1[1] = z1[1];
    }
    if (a.gx0int && og == 0 && k < E) a.gx0int[(size_t)k * Bp + b] = ub + seed_mu[k];
}

// =============================================================================================
// Host orchestration
// =============================================================================================
// returns true if the few-rollouts kernel ran (the one that honours CostArgs::grad_ext)
static bool launch_cost_adjoint(gpmpc_ctx *h, const CostArgs &ca, int B, int H)
{
    if (B < kSingleMaxB) {
        const size_t smem = ((size_t)3 * (H + 1) * ca.d.E + (size_t)H * (ca.d.m > 0 ? ca.d.m : 1) + 3 * H + 2 +
                             (size_t)H * ca.d.E * 4 * ca.d.D) * sizeof(double);
        if (smem <= 160 * 1024) {
            static bool configured[kMaxDevices] = {};
            if (first_use_on_device(configured))
                cudaFuncSetAttribute(cost_adjoint_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
            cost_adjoint_small_kernel<<<B, 128, smem, h->stream>>>(ca);
            return true;
        }
    }
    cost_adjoint_kernel<<<(B + 127) / 128, 128, 0, h->stream>>>(ca);
    return false;
}

static int check_ready(gpmpc_ctx *h, int B, int H)
{
    if (!h) return GPMPC_ERR_INVALID;
    if (!h->fitted || h->n <= 0) return fail(h, GPMPC_ERR_NOT_FIT, "no training data: call gpmpc_fit first");
    if (B <= 0 || H < 0) return fail(h, GPMPC_ERR_INVALID, "B must be > 0 and H >= 0");
    if (h->m < 0) return fail(h, GPMPC_ERR_INVALID, "D < E");
    if (!pair_launcher(h->D)) return fail(h, GPMPC_ERR_UNSUPPORTED, "input dimension D outside 2..8");
    return GPMPC_OK;
}

static StepDims make_dims(gpmpc_ctx *h, int B)
{
    StepDims d;
    d.B = B; d.Bpad = round_up(B, 32); d.D = h->D; d.E = h->E; d.m = h->m; d.G = (int)h->groups.size();
    d.n = h->n; d.ld = h->ld;
    return d;
}

// Launch geometry of the pair kernel: one wave of 2 CTAs per SM (register limited), split over the rollout
// chunks; each chunk's tile list is cut into ~kItemsPerCta work items per CTA that are handed out dynamically.
constexpr int kItemsPerCta = 16;    // measured in the bench: 16 items 738 evals/s, 8 items 714 (coarser tail)
static void pair_geometry(gpmpc_ctx *h, int B, long long total_tiles, int &ctas_per_chunk, int &n_items)
{
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    if (B < kSingleMaxB) {                       // mm_step_single: SINGLE_CTAS_PER_SM CTAs per SM and rollout, static tile ranges
        long long c = (long long)SINGLE_CTAS_PER_SM * sms;
        if (c > total_tiles) c = total_tiles;
        ctas_per_chunk = (int)c;
        n_items = (int)c;
        return;
    }
    const int chunks = (B + PAIR_THREADS - 1) / PAIR_THREADS;
    long long c = (2LL * sms) / chunks;          // floor: never spill into a second wave
    if (c < 1) c = 1;
    if (c > total_tiles) c = total_tiles;
    // few rollout chunks -> many CTAs per chunk and few tiles per CTA: fewer, larger items (per-item cost: a ticket,
    // two barriers and a partial-sum store); measured on B200 at n=4096, B=128: 2 items per CTA 628, 4 items 763, 8 items 712 evals/s
    // (round 2, faster kernel, tools/pair_bench.cu: B=256 -> 4 items 10.13 ms, 8 items 10.59; B=512 -> 16 items 20.00, 4 items 20.23)
    long long it = c * (chunks >= 4 ? kItemsPerCta : 4);
    if (it > total_tiles) it = total_tiles;
    ctas_per_chunk = (int)c;
    n_items = (int)it;
}

template <int D> static void launch_mean(const MeanArgs &ma, dim3 grid, cudaStream_t st)
{
    mean_sums_kernel<D><<<grid, MEAN_THREADS, 0, st>>>(ma);
}
static void launch_mean_d(int D, const MeanArgs &ma, dim3 grid, cudaStream_t st)
{
    switch (D) {
        case 2: launch_mean<2>(ma, grid, st); break; case 3: launch_mean<3>(ma, grid, st); break;
        case 4: launch_mean<4>(ma, grid, st); break; case 5: launch_mean<5>(ma, grid, st); break;
        case 6: launch_mean<6>(ma, grid, st); break; case 7: launch_mean<7>(ma, grid, st); break;
        case 8: launch_mean<8>(ma, grid, st); break;
    }
}

// Constants of the following step, written by the fused few-rollouts kernel (one lambda group only).
struct NextPrep { bool on = false; const double *Uint = nullptr; const double *lam_group = nullptr; double act_var = 0.0; };

// One moment-matching step for all rollouts: us/cst must have been prepared.  Writes mean/var of step t
// (slot t of mu/var) and, if want_grad, the tape entry t-1.
//   B <  kSingleMaxB: one launch of mm_step_single per lambda group does everything (pairs, mean, finalize)
//   B >= kSingleMaxB: mm_pairs_batch + mean_sums per group, then finalize_step
// grad_mode: 0 forward values only, 1 all moments the adjoint reads, 2 first step of a rollout whose d/dx0 is not
// requested (mm_pairs.cuh).
static int run_step(gpmpc_ctx *h, const StepDims &d, int t, int ctas, int P, long long total_tiles, int grad_mode,
                    double *us, double *cst, double *mu, double *var, double *tape, const NextPrep &next = NextPrep(),
                    int *claim_row = nullptr)
{
    const bool want_grad = grad_mode != 0;
    const size_t mat = (size_t)h->ld * h->ld;
    const bool few = d.B < kSingleMaxB;
    if (h->time_pairs) cudaEventRecord(h->ev0, h->stream);
    for (int g = 0; g < d.G; ++g) {
        const LambdaGroup &grp = h->groups[g];
        cudaError_t e;
        if (few) {
            SingleStepArgs sa;
            for (int i = 0; i < kGroupMax; ++i) {
                const int o = grp.outputs[i < grp.count ? i : 0];
                sa.Wt[i] = h->Wt.as<double>() + (size_t)o * wt_doubles(h->ld);
                sa.beta[i] = h->beta.as<double>() + (size_t)o * h->ld;
                sa.out_idx[i] = o;
            }
            sa.X = h->X.as<double>();
            sa.cst = cst + (size_t)g * 4 * d.D * d.Bpad;
            sa.spart = h->part.as<double>();
            sa.tickets = h->tickets.as<int>();
            sa.ld = h->ld; sa.ntile = h->ld / PT; sa.total_tiles = (int)total_tiles;
            sa.d = d; sa.t = t; sa.us = us; sa.hyp = h->hyp.as<double>(); sa.mu = mu; sa.var = var; sa.tape = tape;
            sa.want_grad = want_grad ? 1 : 0;
            sa.prep_next = (next.on && d.G == 1) ? 1 : 0;
            sa.Uint = next.Uint; sa.lam_group = next.lam_group; sa.us_w = us; sa.cst_w = cst; sa.act_var = next.act_var;
            sa.dbg = nullptr;
            sa.claim = nullptr; sa.tiles_big = 0;
            if (claim_row) { sa.claim = claim_row + (size_t)g * kClaimStride; sa.tiles_big = (int)(total_tiles * h->opt_single_big / 1000); }
            sa.l2_base = nullptr; sa.l2_bytes = 0; sa.l2_hit = 0.f;
            if (h->opt_l2_persist) {
                if (h->l2_persist_max < 0) {                 // first use: device limits, carve out the persisting part of L2
                    int pm = 0, wm = 0;
                    cudaDeviceGetAttribute(&pm, cudaDevAttrMaxPersistingL2CacheSize, h->device);
                    cudaDeviceGetAttribute(&wm, cudaDevAttrMaxAccessPolicyWindowSize, h->device);
                    h->l2_persist_max = pm; h->l2_window_max = wm;
                    if (pm > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)pm);
                }
                const size_t wbytes = (size_t)d.E * wt_doubles(h->ld) * sizeof(double);
                if (h->l2_persist_max > 0 && h->l2_window_max > 0 && wbytes > (size_t)h->l2_persist_max) {
                    // a working set that fits L2 anyway needs no policy
                    sa.l2_base = h->Wt.p;
                    sa.l2_bytes = wbytes < (size_t)h->l2_window_max ? wbytes : (size_t)h->l2_window_max;
                    const double hit = 0.9 * (double)h->l2_persist_max / (double)sa.l2_bytes;
                    sa.l2_hit = (float)(hit > 1.0 ? 1.0 : hit);
                }
            }
            static const bool step_debug = getenv("GPMPC_STEP_DEBUG") != nullptr;
            if (step_debug) {                            // development aid: per-CTA phase stamps of every launch
                GP_CUDA(h, h->dbg.reserve((size_t)d.B * ctas * 6 * sizeof(unsigned long long)));
                GP_CUDA(h, cudaMemsetAsync(h->dbg.p, 0, (size_t)d.B * ctas * 6 * sizeof(unsigned long long), h->stream));
                sa.dbg = h->dbg.as<unsigned long long>();
            }
            e = single_launcher(d.D)(grp.count, grad_mode, d.E, sa, dim3(d.B, ctas), h->stream);
            h->launches++;
            if (e != cudaSuccess) return fail(h, GPMPC_ERR_CUDA, std::string("mm_step_single: ") + cudaGetErrorString(e));
            if (step_debug) {
                std::vector<unsigned long long> st((size_t)d.B * ctas * 6);
                GP_CUDA(h, cudaMemcpyAsync(st.data(), h->dbg.p, st.size() * 8, cudaMemcpyDeviceToHost, h->stream));
                GP_CUDA(h, cudaStreamSynchronize(h->stream));
                unsigned long long t0 = ~0ull, tend = 0;
                for (size_t c = 0; c < st.size() / 6; ++c) { if (st[c * 6] && st[c * 6] < t0) t0 = st[c * 6]; }
                double mx[6] = {0, 0, 0, 0, 0, 0}, av[6] = {0, 0, 0, 0, 0, 0};
                for (size_t c = 0; c < st.size() / 6; ++c)
                    for (int i = 0; i < 6; ++i) {
                        if (!st[c * 6 + i]) continue;
                        const double v = (double)(st[c * 6 + i] - t0) * 1e-3;
                        if (v > mx[i]) mx[i] = v;
                        av[i] += v / (st.size() / 6);
                        if (st[c * 6 + i] > tend) tend = st[c * 6 + i];
                    }
                if (const char *path = getenv("GPMPC_STEP_DEBUG_FILE")) {
                    if (FILE *f = fopen(path, "w")) {
                        fprintf(f, "cta,start,depwait,loop_start,loop_end,ticket,done\n");
                        for (size_t c = 0; c < st.size() / 6; ++c) {
                            fprintf(f, "%zu", c);
                            for (int i = 0; i < 6; ++i) fprintf(f, ",%.3f", st[c * 6 + i] ? (double)(st[c * 6 + i] - t0) * 1e-3 : -1.0);
                            fprintf(f, "\n");
                        }
                        fclose(f);
                    }
                }
                fprintf(stderr, "step %d us since first CTA start: start avg %.1f max %.1f | dep-wait done avg %.1f max %.1f | loop "
                                "start avg %.1f max %.1f | loop end avg %.1f max %.1f | ticket avg %.1f max %.1f | last CTA done %.1f\n",
                        t, av[0], mx[0], av[1], mx[1], av[2], mx[2], av[3], mx[3], av[4], mx[4], mx[5]);
            }
            continue;
        }
        PairArgs pa;
        MeanArgs ma;
        for (int i = 0; i < kGroupMax; ++i) {
            const int o = grp.outputs[i < grp.count ? i : 0];
            pa.Wt[i] = h->Wt.as<double>() + (size_t)o * wt_doubles(h->ld);
            pa.out_idx[i] = o;
            ma.beta[i] = h->beta.as<double>() + (size_t)o * h->ld;
            ma.out_idx[i] = o;
        }
        pa.X = h->X.as<double>();
        pa.cst = cst + (size_t)g * 4 * d.D * d.Bpad;
        pa.part = h->part.as<double>();
        pa.ld = h->ld; pa.ntile = h->ld / PT; pa.B = d.B; pa.Bpad = d.Bpad; pa.E = d.E; pa.n_items = P;
        pa.total_tiles = (int)total_tiles; pa.chunks = (d.B + PAIR_THREADS - 1) / PAIR_THREADS;
        const int chunks = (d.B + PAIR_THREADS - 1) / PAIR_THREADS;
        pa.counters = h->tickets.as<int>();
        GP_CUDA(h, cudaMemsetAsync(pa.counters, 0, chunks * sizeof(int), h->stream));
        e = pair_launcher(d.D)(grp.count, grad_mode, d.E, pa, dim3(ctas * chunks), h->stream);
        h->launches++;
        if (e != cudaSuccess) return fail(h, GPMPC_ERR_CUDA, std::string("mm_pairs_batch: ") + cudaGetErrorString(e));

        ma.X = pa.X; ma.EG = grp.count; ma.cst = pa.cst; ma.mpart = h->mpart.as<double>(); ma.d = d;
        ma.k1 = grad_mode == 2 ? d.E : (grad_mode ? 0 : d.D);
        ma.k2 = grad_mode == 1 ? d.E : 0;
        dim3 mgrid((d.B + MEAN_THREADS - 1) / MEAN_THREADS, MEAN_JP);
        launch_mean_d(d.D, ma, mgrid, h->stream);
        GP_LAUNCH_CHECK(h);
    }
    if (h->time_pairs) cudaEventRecord(h->ev1, h->stream);
    if (!few) {
        dim3 fgrid((d.B + 31) / 32, d.E);
        auto launch_fin = [&](auto dc) {
            constexpr int DD = decltype(dc)::value;
            static bool fin_configured[kMaxDevices] = {};            // (one flag set per instantiation)
            if (first_use_on_device(fin_configured))
                cudaFuncSetAttribute(finalize_step_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)finalize_smem_bytes(DD));
            finalize_step_kernel<DD><<<fgrid, 32 * FIN_WARPS, finalize_smem_bytes(d.D), h->stream>>>(
                d, t, P, h->part.as<double>(), h->mpart.as<double>(), us, h->hyp.as<double>(), mu, var, tape, want_grad ? 1 : 0);
        };
        switch (d.D) {
            case 2: launch_fin(std::integral_constant<int, 2>{}); break; case 3: launch_fin(std::integral_constant<int, 3>{}); break;
            case 4: launch_fin(std::integral_constant<int, 4>{}); break; case 5: launch_fin(std::integral_constant<int, 5>{}); break;
            case 6: launch_fin(std::integral_constant<int, 6>{}); break; case 7: launch_fin(std::integral_constant<int, 7>{}); break;
            case 8: launch_fin(std::integral_constant<int, 8>{}); break;
        }
        GP_LAUNCH_CHECK(h);
    }
    return GPMPC_OK;
}

struct RolloutWork {
    StepDims d; int P; int ctas; long long total_tiles;
    double *x0int, *Uint, *us, *cst, *lamg;
};

static int reserve_rollout(gpmpc_ctx *h, int B, int H, RolloutWork &w)
{
    w.d = make_dims(h, B);
    const StepDims &d = w.d;
    const long long nt = h->ld / PT;
    w.total_tiles = nt * (nt + 1) / 2;
    pair_geometry(h, B, w.total_tiles, w.ctas, w.P);
    const bool few = B < kSingleMaxB;
    const size_t n_groups = (size_t)(w.P + SINGLE_GROUP - 1) / SINGLE_GROUP;
    // few rollouts: [B][1 + groups] arrival counters, then [B] completed steps and one error flag of the persistent kernel
    const size_t n_tickets = few ? (size_t)B * (1 + n_groups) + B + 1 : (size_t)((B + PAIR_THREADS - 1) / PAIR_THREADS);
    GP_CUDA(h, h->tickets.reserve(n_tickets * sizeof(int)));
    if (few) GP_CUDA(h, cudaMemsetAsync(h->tickets.as<int>(), 0, n_tickets * sizeof(int), h->stream));
    const size_t Bp = d.Bpad;
    const int Hs = H > 0 ? H : 1;
    GP_CUDA(h, h->mu.reserve((size_t)(H + 1) * d.E * Bp * sizeof(double)));
    GP_CUDA(h, h->var.reserve((size_t)(H + 1) * d.E * Bp * sizeof(double)));
    GP_CUDA(h, h->tape.reserve((size_t)Hs * d.E * ntape(d.D) * Bp * sizeof(double)));
    // few rollouts: [B][P + groups][2 * 4 * nacc] (pair + mean partials of one lambda group); else [P][E][nacc][Bp]
    GP_CUDA(h, h->part.reserve(few ? (size_t)B * (w.P + n_groups) * 2 * kGroupMax * nacc(d.D) * sizeof(double)
                                   : (size_t)w.P * d.E * nacc(d.D) * Bp * sizeof(double)));
    GP_CUDA(h, h->mpart.reserve((size_t)MEAN_JP * d.E * nacc(d.D) * Bp * sizeof(double)));
    // cst: x0int [E][Bp] | Uint [H*m][Bp] | us [2D][Bp] | cst [G][4D][Bp]
    const size_t cnt = (size_t)d.E * Bp + (size_t)Hs * (d.m > 0 ? d.m : 1) * Bp + 2 * (size_t)d.D * Bp +
                       (size_t)d.G * 4 * d.D * Bp + 64;
    GP_CUDA(h, h->cst.reserve(cnt * sizeof(double)));
    double *p = h->cst.as<double>();
    w.x0int = p; p += (size_t)d.E * Bp;
    w.Uint = p; p += (size_t)Hs * (d.m > 0 ? d.m : 1) * Bp;
    w.us = p; p += 2 * (size_t)d.D * Bp;
    w.cst = p;
    w.lamg = h->hyp.as<double>() + (size_t)d.E * d.D + d.E;      // group lambdas [G][D], see upload_prop_hypers
    return GPMPC_OK;
}

// stage a host-or-device array into a device buffer (returns the device pointer to use)
static int stage_in(gpmpc_ctx *h, DevBuf &buf, size_t &off, const double *src, size_t count, const double **dev)
{
    if (is_device_ptr(src)) { *dev = src; return GPMPC_OK; }
    double *dst = reinterpret_cast<double *>(reinterpret_cast<char *>(buf.p) + off);
    GP_CUDA(h, cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    off += (count * sizeof(double) + 255) / 256 * 256;
    *dev = dst;
    return GPMPC_OK;
}

// forward rollout into h->mu / h->var (+ tape).  x0_dev [B,E], U_dev [B,H,m] are device pointers.
// need_gx0: the caller may ask for d/dx0 later (gpmpc_rollout_vjp), so step 1 keeps the state-dimension moments too.
static int forward(gpmpc_ctx *h, int B, int H, const double *x0_dev, const double *U_dev, bool want_grad, bool need_gx0,
                   RolloutWork &w)
{
    int rc = reserve_rollout(h, B, H, w);
    if (rc) return rc;
    const StepDims &d = w.d;
    dim3 blk(128);
    const double act_var = (double)1e-3f;          // fp32 eye in the action block, src/dynamics.py:162
    const bool begun = B < kSingleMaxB;            // few rollouts: shuffles, initial state and the constants of step 1 in one launch
    if (begun) {
        begin_rollout_small_kernel<<<(B + 31) / 32, dim3(32, 4), 0, h->stream>>>(d, H, x0_dev, U_dev, w.x0int, w.Uint,
                                                                                h->mu.as<double>(), h->var.as<double>(), 1e-3,
                                                                                w.lamg, w.us, w.cst, act_var);
        GP_LAUNCH_CHECK(h);
    } else {
        to_internal_kernel<<<dim3((B + 127) / 128, d.E), blk, 0, h->stream>>>(x0_dev, B, d.Bpad, d.E, w.x0int);
        GP_LAUNCH_CHECK(h);
        if (H > 0 && d.m > 0) {
            to_internal_kernel<<<dim3((B + 127) / 128, H * d.m), blk, 0, h->stream>>>(U_dev, B, d.Bpad, H * d.m, w.Uint);
            GP_LAUNCH_CHECK(h);
        }
        init_state_kernel<<<dim3((B + 127) / 128, d.E), blk, 0, h->stream>>>(w.x0int, B, d.Bpad, d.E, h->mu.as<double>(),
                                                                               h->var.as<double>(), 1e-3);
        GP_LAUNCH_CHECK(h);
    }
    h->last_pair_ms = 0.0; h->last_pair_evals = 0;
    // few rollouts and one lambda group: the step kernel itself prepares the constants of the following step
    const bool fused_prep = B < kSingleMaxB && d.G == 1;
    // a single rollout: the whole horizon in one persistent cooperative launch
    static const bool no_persist = getenv("GPMPC_NO_PERSISTENT") != nullptr || getenv("GPMPC_STEP_DEBUG") != nullptr;
    if (fused_prep && B <= kPersistMaxB && H >= 2 && !no_persist && (h->opt_persistent || h->split_world > 1)) {
        // (the constants of step 1 are there: begin_rollout_small_kernel)
        const LambdaGroup &grp = h->groups[0];
        RolloutSingleArgs ra;
        for (int i = 0; i < kGroupMax; ++i) {
            const int o = grp.outputs[i < grp.count ? i : 0];
            ra.Wt[i] = h->Wt.as<double>() + (size_t)o * wt_doubles(h->ld);
            ra.beta[i] = h->beta.as<double>() + (size_t)o * h->ld;
            ra.out_idx[i] = o;
        }
        const size_t n_groups = (size_t)(w.P + SINGLE_GROUP - 1) / SINGLE_GROUP;
        ra.X = h->X.as<double>(); ra.cst = w.cst; ra.us = w.us; ra.spart = h->part.as<double>();
        ra.tickets = h->tickets.as<int>();
        ra.step_done = ra.tickets + (size_t)B * (1 + n_groups); ra.error = ra.step_done + B;
        ra.ld = h->ld; ra.ntile = h->ld / PT; ra.total_tiles = (int)w.total_tiles; ra.H = H; ra.d = d;
        ra.hyp = h->hyp.as<double>(); ra.mu = h->mu.as<double>(); ra.var = h->var.as<double>(); ra.tape = h->tape.as<double>();
        ra.want_grad = want_grad ? 1 : 0; ra.first_mode = need_gx0 ? 1 : 2;
        ra.Uint = w.Uint; ra.lam_group = w.lamg; ra.act_var = act_var;
        // one rollout split over several GPUs (split.cu): this rank sweeps 1/world of the tiles with as many CTAs as fit
        ra.world = h->split_world; ra.rank = h->split_rank; ra.seq0 = h->split_seq;
        for (int r = 0; r < kSplitMaxWorld; ++r) { ra.peer_mail[r] = h->peer_mail[r]; ra.peer_flags[r] = h->peer_flags[r]; }
        int ctas = w.ctas;
        if (ra.world > 1) {
            const long long local = w.total_tiles * (ra.rank + 1) / ra.world - w.total_tiles * ra.rank / ra.world;
            if (local < ctas) ctas = (int)(local > 0 ? local : 1);
            h->split_seq += H;                          // every rank advances identically (same calls in the same order)
        }
        ra.xstamp = nullptr;
        if (ra.world > 1 && h->opt_split_timeline) {
            GP_CUDA(h, h->dbg.reserve((size_t)H * 2 * sizeof(unsigned long long)));
            ra.xstamp = h->dbg.as<unsigned long long>();
        }
        GP_CUDA(h, cudaMemsetAsync(ra.step_done, 0, (size_t)(B + 1) * sizeof(int), h->stream));
        if (h->time_pairs) cudaEventRecord(h->ev0, h->stream);
        cudaError_t e = rollout_single_launcher(d.D)(grp.count, d.E, ra, dim3(B, ctas), h->stream);
        if (e == cudaSuccess) {
            h->launches++;
            if (h->time_pairs) {
                cudaEventRecord(h->ev1, h->stream);
                GP_CUDA(h, cudaEventSynchronize(h->ev1));
                float ms = 0.f;
                cudaEventElapsedTime(&ms, h->ev0, h->ev1);
                h->last_pair_ms = ms;
                h->last_pair_evals = (long long)H * B * d.E * ((long long)h->n * (h->n + 1) / 2);
            }
            if (ra.xstamp) {                             // development / bench aid: time from "sums ready" to "all peers' sums read"
                std::vector<unsigned long long> st((size_t)H * 2);
                GP_CUDA(h, cudaMemcpyAsync(st.data(), ra.xstamp, st.size() * 8, cudaMemcpyDeviceToHost, h->stream));
                GP_CUDA(h, cudaStreamSynchronize(h->stream));
                double sum = 0.0, mx = 0.0;
                for (int t = 0; t < H; ++t) { const double us = (double)(st[2 * t + 1] - st[2 * t]) * 1e-3; sum += us; if (us > mx) mx = us; }
                h->split_exchange_mean_us = sum / H; h->split_exchange_max_us = mx;
            }
            // a spin wait that ran out of budget leaves garbage behind: turn it into NaN, which the caller sees as data
            poison_on_error_kernel<<<1, 128, 0, h->stream>>>(ra.error, h->mu.as<double>() + (size_t)d.E * d.Bpad,
                                                            h->var.as<double>() + (size_t)d.E * d.Bpad, (size_t)H * d.E * d.Bpad);
            GP_LAUNCH_CHECK(h);
            h->tape_B = (want_grad && need_gx0) ? B : 0;
            h->tape_H = (want_grad && need_gx0) ? H : 0;
            return GPMPC_OK;
        }
        if (e != cudaErrorCooperativeLaunchTooLarge)
            return fail(h, GPMPC_ERR_CUDA, std::string("mm_rollout_single: ") + cudaGetErrorString(e));
        cudaGetLastError();                            // the grid does not fit: one launch per step instead
    }
    // one rollout on a grid of exactly two CTAs per SM: the CTAs claim unequal slices (mm_step_single.cuh)
    int *claim = nullptr;
    {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
        if (B == 1 && H > 0 && h->opt_single_big > 500 && w.ctas == SINGLE_CTAS_PER_SM * sms && SINGLE_CTAS_PER_SM == 2 && sms <= 256) {
            const size_t bytes = (size_t)H * d.G * kClaimStride * sizeof(int);
            GP_CUDA(h, h->claim.reserve(bytes));
            GP_CUDA(h, cudaMemsetAsync(h->claim.p, 0, bytes, h->stream));
            claim = h->claim.as<int>();
        }
    }
    for (int t = 1; t <= H; ++t) {
        if (!(begun && t == 1) && (!fused_prep || t == 1)) {
            prep_step_kernel<<<dim3((B + 127) / 128, d.G), blk, 0, h->stream>>>(d, t, h->mu.as<double>(), h->var.as<double>(),
                                                                                 w.Uint, w.lamg, w.us, w.cst, act_var);
            GP_LAUNCH_CHECK(h);
        }
        NextPrep next;
        next.on = fused_prep && t < H; next.Uint = w.Uint; next.lam_group = w.lamg; next.act_var = act_var;
        const int grad_mode = !want_grad ? 0 : ((t == 1 && !need_gx0) ? 2 : 1);
        rc = run_step(h, d, t, w.ctas, w.P, w.total_tiles, grad_mode, w.us, w.cst, h->mu.as<double>(),
                      h->var.as<double>(), h->tape.as<double>(), next,
                      claim ? claim + (size_t)(t - 1) * d.G * kClaimStride : nullptr);
        if (rc) return rc;
        if (h->time_pairs) {
            GP_CUDA(h, cudaEventSynchronize(h->ev1));
            float ms = 0.f;
            cudaEventElapsedTime(&ms, h->ev0, h->ev1);
            h->last_pair_ms += ms;
            h->last_pair_evals += (long long)B * d.E * ((long long)h->n * (h->n + 1) / 2);
        }
    }
    // only a tape with the step-1 state derivatives can serve gpmpc_rollout_vjp (which may be asked for d/dx0)
    h->tape_B = (want_grad && need_gx0) ? B : 0;
    h->tape_H = (want_grad && need_gx0) ? H : 0;
    return GPMPC_OK;
}

static int export_traj(gpmpc_ctx *h, const StepDims &d, int H, double *means, double *vars)
{
    // means / vars: [B, H+1, E] host or device
    const size_t count = (size_t)d.B * (H + 1) * d.E;
    for (int which = 0; which < 2; ++which) {
        double *out = which ? vars : means;
        if (!out) continue;
        const double *src = which ? h->var.as<double>() : h->mu.as<double>();
        double *dev = out;
        const bool host = !is_device_ptr(out);
        if (host) { GP_CUDA(h, h->stage_out.reserve(count * sizeof(double))); dev = h->stage_out.as<double>(); }
        to_external_kernel<<<dim3((d.B + 127) / 128, (H + 1) * d.E), 128, 0, h->stream>>>(src, d.B, d.Bpad, (H + 1) * d.E, dev);
        GP_LAUNCH_CHECK(h);
        if (host) {
            GP_CUDA(h, cudaMemcpyAsync(out, dev, count * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            GP_CUDA(h, cudaStreamSynchronize(h->stream));
        }
    }
    return GPMPC_OK;
}

}  // namespace gpmpc

using namespace gpmpc;

extern "C" int gpmpc_rollout(gpmpc_handle h, int B, int H, const double *x0, const double *U, double *means, double *vars)
{
    int rc = check_ready(h, B, H);
    if (rc) return rc;
    if (!x0 || (H > 0 && h->m > 0 && !U)) return fail(h, GPMPC_ERR_INVALID, "gpmpc_rollout: null input");
    GP_CUDA(h, cudaSetDevice(h->device));
    const size_t need = ((size_t)B * h->E + (size_t)B * H * h->m) * sizeof(double) + 1024;
    GP_CUDA(h, h->stage_in.reserve(need));
    size_t off = 0;
    const double *x0d, *Ud = nullptr;
    if ((rc = stage_in(h, h->stage_in, off, x0, (size_t)B * h->E, &x0d))) return rc;
    if (H > 0 && h->m > 0 && (rc = stage_in(h, h->stage_in, off, U, (size_t)B * H * h->m, &Ud))) return rc;
    RolloutWork w;
    if ((rc = forward(h, B, H, x0d, Ud, true, true, w))) return rc;
    return export_traj(h, w.d, H, means, vars);
}

static int upload_cost_args(gpmpc_ctx *h, CostArgs &ca, const double *Q, const double *R, const double *Rdelta,
                            const double *xref, const double *uref)
{
    const int E = h->E, m = h->m;
    auto fetch = [&](const double *src, double *dst, size_t cnt) -> cudaError_t {
        if (!src) { std::memset(dst, 0, cnt * sizeof(double)); return cudaSuccess; }
        if (is_device_ptr(src)) {
            cudaError_t e = cudaMemcpyAsync(dst, src, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
            if (e != cudaSuccess) return e;
            return cudaStreamSynchronize(h->stream);
        }
        std::memcpy(dst, src, cnt * sizeof(double));
        return cudaSuccess;
    };
    GP_CUDA(h, fetch(Q, ca.Q, (size_t)E * E));
    GP_CUDA(h, fetch(R, ca.R, (size_t)m * m));
    GP_CUDA(h, fetch(Rdelta, ca.Rd, (size_t)m * m));
    GP_CUDA(h, fetch(xref, ca.xref, E));
    GP_CUDA(h, fetch(uref, ca.uref, m));
    // Q^-1 on the host (LU with partial pivoting), src/mpc.py:179
    {
        double A[kMaxE * kMaxE];
        int piv[kMaxE];
        std::memcpy(A, ca.Q, sizeof(double) * E * E);
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
            for (int r = 0; r < E; ++r) ca.Qi[r * E + col] = x[r];
        }
    }
    return GPMPC_OK;
}

extern "C" int gpmpc_rollout_cost_grad(gpmpc_handle h, int B, int H, const double *x0, const double *U,
                                       const double *gamma, const double *Q, const double *R, const double *Rdelta,
                                       const double *last_u, const double *xref, const double *uref, double *cost,
                                       double *grad, double *means, double *vars)
{
    int rc = check_ready(h, B, H);
    if (rc) return rc;
    if (!x0 || !gamma || !Q || !cost || (H > 0 && h->m > 0 && (!U || !R)))
        return fail(h, GPMPC_ERR_INVALID, "gpmpc_rollout_cost_grad: null input");
    if (Rdelta && !last_u) return fail(h, GPMPC_ERR_INVALID, "gpmpc_rollout_cost_grad: Rdelta needs last_u");
    GP_CUDA(h, cudaSetDevice(h->device));
    const int E = h->E, m = h->m;
    const size_t n_x0 = (size_t)B * E, n_U = (H > 0 && m > 0) ? (size_t)B * H * m : 0, n_lu = Rdelta ? (size_t)B * m : 0;
    const size_t need = (n_x0 + n_U + (size_t)B + n_lu) * sizeof(double) + 2048;
    GP_CUDA(h, h->stage_in.reserve(need));
    const bool want_grad = grad != nullptr;
    const bool cost_host = !is_device_ptr(cost);
    const bool ghost = want_grad && !is_device_ptr(grad);
    const double *x0d, *Ud = nullptr, *gd, *lud = nullptr;
    // Few rollouts with host buffers (the solver-callback pattern): every input goes through ONE pinned copy, and
    // cost + gradient come back as one (a pageable cudaMemcpyAsync costs ~5 us of driver time per call).  The pinned
    // buffers are reused by the next call, so this path is taken only when this call ends with a synchronize.
    const bool few = B < kSingleMaxB;
    const bool packed_in = few && cost_host && !is_device_ptr(x0) && (!n_U || !is_device_ptr(U)) && !is_device_ptr(gamma) &&
                           (!n_lu || !is_device_ptr(last_u));
    if (packed_in) {
        const size_t total = n_x0 + n_U + (size_t)B + n_lu;
        GP_CUDA(h, h->pin_in.reserve(total * sizeof(double)));
        double *hp = h->pin_in.as<double>(), *dp = h->stage_in.as<double>();
        std::memcpy(hp, x0, n_x0 * sizeof(double));                                  x0d = dp;
        if (n_U) std::memcpy(hp + n_x0, U, n_U * sizeof(double));                    Ud = n_U ? dp + n_x0 : nullptr;
        std::memcpy(hp + n_x0 + n_U, gamma, (size_t)B * sizeof(double));             gd = dp + n_x0 + n_U;
        if (n_lu) std::memcpy(hp + n_x0 + n_U + B, last_u, n_lu * sizeof(double));   lud = n_lu ? dp + n_x0 + n_U + B : nullptr;
        GP_CUDA(h, cudaMemcpyAsync(dp, hp, total * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    } else {
        size_t off = 0;
        if ((rc = stage_in(h, h->stage_in, off, x0, n_x0, &x0d))) return rc;
        if (n_U && (rc = stage_in(h, h->stage_in, off, U, n_U, &Ud))) return rc;
        if ((rc = stage_in(h, h->stage_in, off, gamma, (size_t)B, &gd))) return rc;
        if (n_lu && (rc = stage_in(h, h->stage_in, off, last_u, n_lu, &lud))) return rc;
    }

    RolloutWork w;
    if ((rc = forward(h, B, H, x0d, Ud, want_grad, false, w))) return rc;
    const StepDims &d = w.d;

    CostArgs ca;
    std::memset(&ca, 0, sizeof ca);
    if ((rc = upload_cost_args(h, ca, Q, R, Rdelta, xref, uref))) return rc;
    ca.d = d; ca.H = H; ca.mode = 0; ca.has_rd = Rdelta ? 1 : 0; ca.want_grad = want_grad ? 1 : 0;
    ca.mu = h->mu.as<double>(); ca.var = h->var.as<double>(); ca.tape = h->tape.as<double>(); ca.Uint = w.Uint;
    ca.gamma = gd;
    // gbuf: last_u internal [m][Bp] | grad internal [H*m][Bp] | cost [Bp] | grad external [B*H*m]
    const size_t Bp = d.Bpad;
    const size_t gcount = (size_t)m * Bp + (size_t)H * m * Bp + Bp + (size_t)B * H * m + 64;
    GP_CUDA(h, h->gbuf.reserve(gcount * sizeof(double)));
    double *p = h->gbuf.as<double>();
    double *lu_int = p; p += (size_t)m * Bp;
    double *grad_int = p; p += (size_t)H * m * Bp;
    double *cost_dev = p; p += Bp;
    double *grad_ext = p;
    if (Rdelta && m > 0) {
        to_internal_kernel<<<dim3((B + 127) / 128, m), 128, 0, h->stream>>>(lud, B, d.Bpad, m, lu_int);
        GP_LAUNCH_CHECK(h);
    }
    ca.last_u = lu_int;
    ca.cost = cost_host ? cost_dev : cost;
    ca.gradint = grad_int;
    ca.gx0int = nullptr;
    const bool has_grad = want_grad && H > 0 && m > 0;
    double *gdev = ghost ? grad_ext : grad;
    ca.grad_ext = has_grad ? gdev : nullptr;
    const bool direct = launch_cost_adjoint(h, ca, B, H);
    GP_LAUNCH_CHECK(h);
    if (has_grad && !direct) {
        to_external_kernel<<<dim3((B + 127) / 128, H * m), 128, 0, h->stream>>>(grad_int, B, d.Bpad, H * m, gdev);
        GP_LAUNCH_CHECK(h);
    }
    const size_t n_grad = has_grad ? (size_t)B * H * m : 0;
    const bool packed_out = few && cost_host && (!has_grad || ghost);
    if (packed_out) {                              // cost [Bp] and the external gradient are adjacent in gbuf
        GP_CUDA(h, h->pin_out.reserve((Bp + n_grad) * sizeof(double)));
        GP_CUDA(h, cudaMemcpyAsync(h->pin_out.p, cost_dev, (Bp + n_grad) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    } else {
        if (has_grad && ghost) GP_CUDA(h, cudaMemcpyAsync(grad, gdev, n_grad * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (cost_host) GP_CUDA(h, cudaMemcpyAsync(cost, cost_dev, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    if (means || vars) { if ((rc = export_traj(h, d, H, means, vars))) return rc; }
    if (cost_host || ghost) GP_CUDA(h, cudaStreamSynchronize(h->stream));
    if (packed_out) {
        std::memcpy(cost, h->pin_out.as<double>(), (size_t)B * sizeof(double));
        if (n_grad) std::memcpy(grad, h->pin_out.as<double>() + Bp, n_grad * sizeof(double));
    }
    return GPMPC_OK;
}

extern "C" int gpmpc_rollout_vjp(gpmpc_handle h, int B, int H, const double *gmeans, const double *gvars, double *gU,
                                 double *gx0)
{
    int rc = check_ready(h, B, H);
    if (rc) return rc;
    if (h->tape_B != B || h->tape_H != H)
        return fail(h, GPMPC_ERR_INVALID, "gpmpc_rollout_vjp: no tape of this shape (call gpmpc_rollout first)");
    GP_CUDA(h, cudaSetDevice(h->device));
    const int E = h->E, m = h->m;
    StepDims d = make_dims(h, B);
    const size_t Bp = d.Bpad;
    const size_t tcount = (size_t)(H + 1) * E;
    GP_CUDA(h, h->stage_in.reserve(2 * (size_t)B * tcount * sizeof(double) + 1024));
    size_t off = 0;
    const double *gmd = nullptr, *gvd = nullptr;
    if (gmeans && (rc = stage_in(h, h->stage_in, off, gmeans, (size_t)B * tcount, &gmd))) return rc;
    if (gvars && (rc = stage_in(h, h->stage_in, off, gvars, (size_t)B * tcount, &gvd))) return rc;
    // gbuf: seed_mu [tcount][Bp] | seed_var [tcount][Bp] | grad_int [H*m][Bp] | gx0_int [E][Bp] | ext
    const size_t gcount = 2 * tcount * Bp + (size_t)H * m * Bp + (size_t)E * Bp + (size_t)B * H * m + (size_t)B * E + 64;
    GP_CUDA(h, h->gbuf.reserve(gcount * sizeof(double)));
    double *p = h->gbuf.as<double>();
    double *smu = p; p += tcount * Bp;
    double *svar = p; p += tcount * Bp;
    double *grad_int = p; p += (size_t)H * m * Bp;
    double *gx0_int = p; p += (size_t)E * Bp;
    double *ext = p;
    if (gmd) { to_internal_kernel<<<dim3((B + 127) / 128, (unsigned)tcount), 128, 0, h->stream>>>(gmd, B, d.Bpad, (int)tcount, smu); GP_LAUNCH_CHECK(h); }
    if (gvd) { to_internal_kernel<<<dim3((B + 127) / 128, (unsigned)tcount), 128, 0, h->stream>>>(gvd, B, d.Bpad, (int)tcount, svar); GP_LAUNCH_CHECK(h); }
    CostArgs ca;
    std::memset(&ca, 0, sizeof ca);
    ca.d = d; ca.H = H; ca.mode = 1; ca.has_rd = 0; ca.want_grad = 1;
    ca.mu = h->mu.as<double>(); ca.var = h->var.as<double>(); ca.tape = h->tape.as<double>(); ca.Uint = nullptr;
    ca.seed_mu = gmd ? smu : nullptr; ca.seed_var = gvd ? svar : nullptr;
    ca.gradint = grad_int; ca.gx0int = gx0_int; ca.cost = nullptr;
    launch_cost_adjoint(h, ca, B, H);
    GP_LAUNCH_CHECK(h);
    bool sync = false;
    if (gU && H > 0 && m > 0) {
        const bool host = !is_device_ptr(gU);
        double *dev = host ? ext : gU;
        to_external_kernel<<<dim3((B + 127) / 128, H * m), 128, 0, h->stream>>>(grad_int, B, d.Bpad, H * m, dev);
        GP_LAUNCH_CHECK(h);
        if (host) { GP_CUDA(h, cudaMemcpyAsync(gU, dev, (size_t)B * H * m * sizeof(double), cudaMemcpyDeviceToHost, h->stream)); sync = true; }
    }
    if (gx0) {
        const bool host = !is_device_ptr(gx0);
        double *dev = host ? ext + (size_t)B * H * m : gx0;
        to_external_kernel<<<dim3((B + 127) / 128, E), 128, 0, h->stream>>>(gx0_int, B, d.Bpad, E, dev);
        GP_LAUNCH_CHECK(h);
        if (host) { GP_CUDA(h, cudaMemcpyAsync(gx0, dev, (size_t)B * E * sizeof(double), cudaMemcpyDeviceToHost, h->stream)); sync = true; }
    }
    if (sync) GP_CUDA(h, cudaStreamSynchronize(h->stream));
    return GPMPC_OK;
}

// Batched moment matching for diagonal input variances through the same kernels (one "step").
extern "C" int gpmpc_moment_match_diag_internal(gpmpc_handle h, int B, const double *U_dev, const double *S_dev,
                                                double *mean_out, double *var_out);

namespace gpmpc {
// us layout [2D][Bpad] is filled directly from U[B,D], S[B,D]
__global__ void prep_direct_kernel(StepDims d, const double *__restrict__ U, const double *__restrict__ S,
                                   const double *__restrict__ lam_group, double *__restrict__ us, double *__restrict__ cst)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (b >= d.B) return;
    for (int k = 0; k < d.D; ++k) {
        const double u = U[(size_t)b * d.D + k], s = S[(size_t)b * d.D + k];
        if (g == 0) { us[(size_t)k * d.Bpad + b] = u; us[(size_t)(d.D + k) * d.Bpad + b] = s; }
        const double lam = lam_group[g * d.D + k];
        const double a = 1.0 / (0.5 * lam + s), bb = 1.0 / (s + lam);
        const double c = sqrt(0.125 * a), cm = sqrt(0.5 * bb);
        double *cg = cst + (size_t)g * 4 * d.D * d.Bpad;
        cg[(size_t)k * d.Bpad + b] = c;
        cg[(size_t)(d.D + k) * d.Bpad + b] = c * u;
        cg[(size_t)(2 * d.D + k) * d.Bpad + b] = cm;
        cg[(size_t)(3 * d.D + k) * d.Bpad + b] = cm * u;
    }
}
}  // namespace gpmpc

extern "C" int gpmpc_moment_match_diag_internal(gpmpc_handle h, int B, const double *U_dev, const double *S_dev,
                                                double *mean_out, double *var_out)
{
    int rc = check_ready(h, B, 1);
    if (rc) return rc;
    RolloutWork w;
    if ((rc = reserve_rollout(h, B, 1, w))) return rc;
    const StepDims &d = w.d;
    prep_direct_kernel<<<dim3((B + 127) / 128, d.G), 128, 0, h->stream>>>(d, U_dev, S_dev, w.lamg, w.us, w.cst);
    GP_LAUNCH_CHECK(h);
    h->last_pair_ms = 0.0; h->last_pair_evals = 0;
    // results land in slot t = 1 of mu / var
    if ((rc = run_step(h, d, 1, w.ctas, w.P, w.total_tiles, 0, w.us, w.cst, h->mu.as<double>(),
                       h->var.as<double>(), h->tape.as<double>()))) return rc;
    if (h->time_pairs) {
        GP_CUDA(h, cudaEventSynchronize(h->ev1));
        float ms = 0.f; cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        h->last_pair_ms = ms; h->last_pair_evals = (long long)B * d.E * ((long long)h->n * (h->n + 1) / 2);
    }
    h->tape_B = 0; h->tape_H = 0;
    for (int which = 0; which < 2; ++which) {
        double *out = which ? var_out : mean_out;
        if (!out) continue;
        const double *src = (which ? h->var.as<double>() : h->mu.as<double>()) + (size_t)d.E * d.Bpad;
        const bool host = !is_device_ptr(out);
        double *dev = out;
        if (host) { GP_CUDA(h, h->stage_out.reserve((size_t)B * d.E * sizeof(double))); dev = h->stage_out.as<double>(); }
        to_external_kernel<<<dim3((B + 127) / 128, d.E), 128, 0, h->stream>>>(src, B, d.Bpad, d.E, dev);
        GP_LAUNCH_CHECK(h);
        if (host) {
            GP_CUDA(h, cudaMemcpyAsync(out, dev, (size_t)B * d.E * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            GP_CUDA(h, cudaStreamSynchronize(h->stream));
        }
    }
    return GPMPC_OK;
}
