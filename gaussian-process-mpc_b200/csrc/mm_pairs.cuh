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
// loop has no cross-lane traffic at all.  Shared memory is filled by the TMA engine: Wt is stored tile-major
// (common.cuh), so a 32x32 tile is ONE contiguous 8 KB bulk copy (cp.async.bulk + mbarrier, SASS UBLKCP), double
// buffered.  The kernel is bound by the FP64 pipe (one exp + ~14 + 12*EG DFMA-class ops per pair).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace gpmpc {

// Build-time tuning knobs (defaults = the shipped configuration; tools/pair_bench.cu overrides them).
// Chosen from the variant sweeps on B200 (profiles/r01_pair_kernel_tuning.md): the kernel is limited by the
// register-file bandwidth of three-source DFMAs (2.76 cycles each unless an operand is reused, vs 2.0), so most
// schedules land within a few percent; the defaults are the best robust combination.
#ifndef GPMPC_EXP_VARIANT
#define GPMPC_EXP_VARIANT 3      // 0: Horner deg 11, 1: Estrin deg 11, 2: 16-entry table + deg 6, 3: 2 + integer clamp
#endif
#ifndef GPMPC_RI
#define GPMPC_RI 4               // rows of the register micro-tile: four independent q -> exp chains per column
#endif
#ifndef GPMPC_CST_SMEM
#define GPMPC_CST_SMEM 1         // 1: keep the per-rollout constants c, c*u in shared memory instead of registers
#endif
#ifndef GPMPC_DYNAMIC
#define GPMPC_DYNAMIC 1          // 1: CTAs draw work items from a ticket counter; 0: item = CTA index
#endif
#ifndef GPMPC_ROWSUM
#define GPMPC_ROWSUM 1           // 1: gradient variants form N1 = sum w (z_i + z_j) from row sums (per strip) and column sums (per
#endif                           //    micro-tile) of w instead of D FMAs per pair and output (round 2: 40.6 -> 37.0 ms per launch)
#ifndef GPMPC_RS_RI
#define GPMPC_RS_RI 2            // rows x columns of the register micro-tile of the row/column-sum form: 2 x 2 = four independent
#endif                           // exp chains like the 4 x 1 tile, but half the row sums (4 x 1: 47.8 ms, registers; 2 x 1: 39.2 ms)
#ifndef GPMPC_CJ
#define GPMPC_CJ 2
#endif
#ifndef GPMPC_MINBLOCKS
#define GPMPC_MINBLOCKS 2        // CTAs per SM promised to ptxas
#endif

// 2^(j/16), j = 0..15 (table of the exp variant 2; copied to shared memory by each CTA: 128 B = one bank row,
// so 32 lanes reading arbitrary entries never conflict)
static __device__ const double kExp2Tab[16] = {
    1.0, 1.0442737824274138, 1.0905077326652577, 1.1387886347566916, 1.189207115002721, 1.241857812073484,
    1.2968395546510096, 1.3542555469368927, 1.4142135623730951, 1.4768261459394993, 1.5422108254079407,
    1.6104903319492543, 1.681792830507429, 1.7562521603732995, 1.8340080864093424, 1.9152065613971474};

// exp(-S) for S >= 0 (clamped at 700), branch free, relative error ~1e-16.
//   variant 0/1: k = round(-S log2 e), r = -S - k ln2 (Cody-Waite), degree-11 polynomial, exponent insertion
//   variant 2  : k = round(-16 S log2 e), r = -S - k ln2/16 (|r| <= ln2/32), degree-6 polynomial,
//                result = 2^(k>>4) * tab[k&15] * p(r)     (12 FP64-pipe ops instead of 17)
__device__ __forceinline__ double exp_neg(double S, const double *__restrict__ tab)
{
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52
#if GPMPC_EXP_VARIANT == 3
    // like variant 2, but without a clamp on the FP64 dependency chain: the underflow case 700 < S <= +inf is
    // detected with integer compares on the high word of S (ALU pipe, off the chain) and the result is replaced
    // by 0 with two selects at the very end; NaN passes through and stays NaN.
    const bool big = (unsigned)(__double2hiint(S) - 0x4085E001) <= (unsigned)(0x7FF00000 - 0x4085E001);
    double t = fma(S, -23.083120654223414, MAGIC);
    const int k = __double2loint(t);
    const double kf = t - MAGIC;
    double r = fma(kf, -0.04332169867120683, -S);
    r = fma(kf, -1.1378974990650914e-10, r);
    double p = 0.0013889092532749104;
    p = fma(p, r, 0.008333496248686918);
    p = fma(p, r, 0.04166666666188925);
    p = fma(p, r, 0.16666666662844726);
    p = fma(p, r, 0.5000000000000003);
    p = fma(p, r, 1.0000000000000022);
    p = fma(p, r, 1.0);
    p *= tab[k & 15];
    return __hiloint2double(big ? 0 : __double2hiint(p) + (k >> 4) * 1048576, big ? 0 : __double2loint(p));
#elif GPMPC_EXP_VARIANT == 4 || GPMPC_EXP_VARIANT == 5
    S = S > 700.0 ? 700.0 : S;                               // NaN stays NaN, +inf is clamped
    double t = fma(S, -23.083120654223414, MAGIC);
    const int k = __double2loint(t);
    const double kf = t - MAGIC;
#if GPMPC_EXP_VARIANT == 5
    // one-step reduction: the error |kf| * ulp(ln2/16)/2 <= 6e-14 * exp(-S)-weighted is far below 1e-16 of the sums
    const double r = fma(kf, -0.04332169878499658, -S);
    const double r2 = r * r;
    const double pa = fma(1.0000000000000022, r, 1.0), pb = fma(0.16666666662844726, r, 0.5000000000000003),
                 pc = fma(0.008333496248686918, r, 0.04166666666188925);
    const double pd = fma(0.0013889092532749104, r2, pc);
    const double pe = fma(pd, r2, pb);
    double p = fma(pe, r2, pa);
#else
    double r = fma(kf, -0.04332169867120683, -S);
    r = fma(kf, -1.1378974990650914e-10, r);
    double p = 0.0013889092532749104;
    p = fma(p, r, 0.008333496248686918);
    p = fma(p, r, 0.04166666666188925);
    p = fma(p, r, 0.16666666662844726);
    p = fma(p, r, 0.5000000000000003);
    p = fma(p, r, 1.0000000000000022);
    p = fma(p, r, 1.0);
#endif
    p *= tab[k & 15];
    return __hiloint2double(__double2hiint(p) + (k >> 4) * 1048576, __double2loint(p));
#elif GPMPC_EXP_VARIANT == 2
    S = fmin(S, 700.0);
    double t = fma(S, -23.083120654223414, MAGIC);           // round(-S * 16 log2 e) in the low word
    const int k = __double2loint(t);
    const double kf = t - MAGIC;
    double r = fma(kf, -0.04332169867120683, -S);            // ln2/16 hi (24 trailing zero bits)
    r = fma(kf, -1.1378974990650914e-10, r);                 // ln2/16 lo
    double p = 0.0013889092532749104;
    p = fma(p, r, 0.008333496248686918);
    p = fma(p, r, 0.04166666666188925);
    p = fma(p, r, 0.16666666662844726);
    p = fma(p, r, 0.5000000000000003);
    p = fma(p, r, 1.0000000000000022);
    p = fma(p, r, 1.0);
    p *= tab[k & 15];
    return __hiloint2double(__double2hiint(p) + (k >> 4) * 1048576, __double2loint(p));
#else
    S = fmin(S, 700.0);
    double t = fma(S, -1.4426950408889634, MAGIC);          // round(-S * log2 e) in the low word
    const int k = __double2loint(t);
    const double kf = t - MAGIC;
    double r = fma(kf, -6.93147180369123816490e-01, -S);     // ln2 hi
    r = fma(kf, -1.90821492927058770002e-10, r);             // ln2 lo
    const double c2 = 0.5000000000000019, c3 = 0.1666666666666668, c4 = 0.0416666666664881,
                 c5 = 0.008333333333319601, c6 = 0.0013888888952314775, c7 = 0.00019841269890047113,
                 c8 = 2.4801485482328494e-05, c9 = 2.755724091857897e-06, c10 = 2.763263963904103e-07,
                 c11 = 2.5110037605963777e-08;
#if GPMPC_EXP_VARIANT == 1
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double p01 = 1.0 + r, p23 = fma(c3, r, c2), p45 = fma(c5, r, c4), p67 = fma(c7, r, c6),
                 p89 = fma(c9, r, c8), pab = fma(c11, r, c10);
    const double q0 = fma(p23, r2, p01), q1 = fma(p67, r2, p45), q2 = fma(pab, r2, p89);
    double p = fma(q1, r4, q0);
    p = fma(q2, r8, p);
#else
    double p = c11;
    p = fma(p, r, c10); p = fma(p, r, c9); p = fma(p, r, c8); p = fma(p, r, c7); p = fma(p, r, c6);
    p = fma(p, r, c5); p = fma(p, r, c4); p = fma(p, r, c3); p = fma(p, r, c2);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
#endif
    (void)tab;
    return __hiloint2double(__double2hiint(p) + (int)((unsigned)k << 20), __double2loint(p));
#endif
}

__device__ __forceinline__ void cpa16(void *smem, const void *gmem)
{
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- TMA (bulk copy) / mbarrier primitives (sm_90+ PTX; SASS: UBLKCP / SYNCS) ----

__device__ __forceinline__ unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GPMPC_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GPMPC_DONE_%=;\n"
        "bra GPMPC_WAIT_%=;\n"
        "GPMPC_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, unsigned bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct PairArgs {
    const double *Wt[kGroupMax];   // weights of the group's outputs, tile-major upper-triangular tiles (common.cuh)
    int out_idx[kGroupMax];        // global output index of each member
    const double *X;               // [ld, D]
    const double *cst;             // this group's per-rollout constants [4D][Bpad]: c, c*u, cm, cm*u
    double *part;                  // [n_items][E][nacc][Bpad]
    int *counters;                 // [rollout chunks] work-item tickets, zeroed before the launch
    int ld, ntile, B, Bpad, E, n_items, chunks;
    int total_tiles;
#ifdef GPMPC_PAIR_TIMING
    unsigned long long *cta_times;   // [grid][3]: start, end (globaltimer ns), smid
#endif
};

constexpr int PT = kPairTile;      // 32: tile edge
constexpr int PTJ = PT;
constexpr int PAIR_THREADS = 128;
constexpr int RI = GPMPC_RI;       // rows of the register micro-tile

template <int D, int EG>
__host__ __device__ constexpr size_t pair_stage_doubles() { return (size_t)EG * PT * PTJ + (PT + PTJ) * D; }
// dynamic shared memory of mm_pairs_batch: two stages + the 16-entry exp table
template <int D, int EG>
__host__ __device__ constexpr size_t pair_smem_bytes()
{
    return (2 * pair_stage_doubles<D, EG>() + 16 + (GPMPC_CST_SMEM ? 2 * D * 128 : 0)) * sizeof(double);
}

// Which moments a launch accumulates (the adjoint never reads the others, so they are not computed):
//   GRAD = 0  T only (forward values);
//   GRAD = 1  T, N1_k for every input dimension, N2_k for the NS state dimensions k < NS only: d/ds of an ACTION
//             dimension is never used, the action variance is the constant fp32(1e-3) (src/dynamics.py:162);
//   GRAD = 2  first horizon step when d/dx0 is not requested: the state part of the input (x0, 1e-3 I) is a
//             constant, so only N1_k of the action dimensions k >= NS is needed.
// NS = D accumulates everything (the general moment-matching entry points).
template <int D, int EG, int GRAD, int NS>
__global__ void __launch_bounds__(PAIR_THREADS, GPMPC_MINBLOCKS)
mm_pairs_batch(const PairArgs a)
{
    constexpr int K1 = GRAD == 2 ? NS : 0;       // N1_k for k in [K1, D)
    constexpr int K2 = GRAD == 1 ? NS : 0;       // N2_k for k in [0, K2)
    extern __shared__ __align__(128) double smem[];
    constexpr size_t STAGE = pair_stage_doubles<D, EG>();
    const int tid = threadIdx.x;
    // consecutive CTAs serve different rollout chunks, so that the early-launched (favoured) and late-launched
    // CTAs of the SMs are spread evenly over the chunks' ticket counters
    const int chunk_id = blockIdx.x % a.chunks;
    const int b = chunk_id * PAIR_THREADS + tid;
    const bool active = b < a.B;
    const bool warp_active = chunk_id * PAIR_THREADS + (tid & ~31) < a.B;   // ragged last chunk: idle warps only keep the barriers
    double *tab = smem + 2 * STAGE;
    if (tid < 16) tab[tid] = kExp2Tab[tid];      // visible after the first __syncthreads of the tile loop
#ifdef GPMPC_PAIR_TIMING
    unsigned long long t_start = 0;
    if (tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
#endif

    // per-rollout constants
#if GPMPC_CST_SMEM
    double *cs = tab + 16 + tid;                 // [2D][128], this thread's column
#pragma unroll
    for (int k = 0; k < D; ++k) {
        cs[k * PAIR_THREADS] = active ? a.cst[(size_t)k * a.Bpad + b] : 0.0;
        cs[(D + k) * PAIR_THREADS] = active ? a.cst[(size_t)(D + k) * a.Bpad + b] : 0.0;
    }
#define GP_C(k) cs[(k) * PAIR_THREADS]
#define GP_CU(k) cs[(D + (k)) * PAIR_THREADS]
#else
    double c[D], cu[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        c[k] = active ? a.cst[(size_t)k * a.Bpad + b] : 0.0;
        cu[k] = active ? a.cst[(size_t)(D + k) * a.Bpad + b] : 0.0;
    }
#define GP_C(k) c[k]
#define GP_CU(k) cu[k]
#endif

    // full[stage] completes when the TMA engine has written the stage's bytes; one thread arms and issues.
    __shared__ __align__(8) unsigned long long full[2];
    if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); }
    __syncthreads();
    unsigned uses0 = 0, uses1 = 0;               // completed uses of each stage -> wait parity
    constexpr unsigned STAGE_BYTES = (unsigned)(STAGE * sizeof(double));
    auto issue = [&](int stage, int t, int ti, int tj) {       // t = index of tile (ti, tj) in the tile-major list
        if (tid == 0) {
            double *base = smem + (size_t)stage * STAGE;
            void *bar = &full[stage];
            mbar_expect_tx(bar, STAGE_BYTES);
#pragma unroll
            for (int g = 0; g < EG; ++g)
                bulk_load_1d(base + (size_t)g * PT * PT, a.Wt[g] + (size_t)t * PT * PT, PT * PT * sizeof(double), bar);
            double *xi = base + (size_t)EG * PT * PT;
            bulk_load_1d(xi, a.X + (size_t)ti * PT * D, PT * D * sizeof(double), bar);
            bulk_load_1d(xi + PT * D, a.X + (size_t)tj * PT * D, PT * D * sizeof(double), bar);
        }
    };
    auto wait_stage = [&](int stage) {
        if (stage == 0) { mbar_wait(&full[0], uses0 & 1); ++uses0; }
        else { mbar_wait(&full[1], uses1 & 1); ++uses1; }
    };

    // Work items = fixed contiguous ranges of the upper-triangular tile list.  CTAs draw items from a ticket
    // counter (the two CTAs of an SM do not progress at the same rate: the warp scheduler favours one of
    // them), but every item writes its own partial-sum slot, so the result does not depend on who ran what.
#if GPMPC_DYNAMIC
    __shared__ int s_item;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(&a.counters[chunk_id], 1);
        __syncthreads();
        const int item = s_item;
        if (item >= a.n_items) break;
#else
    {
        const int item = blockIdx.x / a.chunks;
#endif

        double accT[EG], acc1[GRAD ? EG : 1][D], acc2[GRAD ? EG : 1][D];
#pragma unroll
        for (int g = 0; g < EG; ++g) accT[g] = 0.0;
        if (GRAD) {
#pragma unroll
            for (int g = 0; g < EG; ++g)
#pragma unroll
                for (int k = 0; k < D; ++k) acc1[g][k] = acc2[g][k] = 0.0;
        }

        const int t_begin = (int)((long long)a.total_tiles * item / a.n_items);
        const int t_end = (int)((long long)a.total_tiles * (item + 1) / a.n_items);
        int I = 0, J = 0;
        {
            int rem = t_begin;
            int row = 0;
            while (rem >= a.ntile - row) { rem -= a.ntile - row; ++row; }
            I = row; J = row + rem;
        }
        if (t_begin < t_end) issue(0, t_begin, I, J);
        int stage = 0;
        for (int t = t_begin; t < t_end; ++t) {
            int In = I, Jn = J + 1;
            if (Jn == a.ntile) { ++In; Jn = In; }
            if (t + 1 < t_end) issue(stage ^ 1, t + 1, In, Jn);
            wait_stage(stage);

            const double *Ws = smem + (size_t)stage * STAGE;
            const double *xi = Ws + (size_t)EG * PT * PTJ;
            const double *xj = xi + PT * D;

            if (warp_active) {
#if GPMPC_ROWSUM
            if constexpr (GRAD != 0) {
            // N1_k = sum w (z_ik + z_jk) = sum_i z_ik R_i + sum_j z_jk C_j with R_i / C_j the row / column sums of w inside the
            // strip / micro-tile: 1 + 1/2 + D/2 adds and FMAs per pair and output (2 x 2 tile) instead of 1 + D.
            constexpr int CJ = GPMPC_CJ, RSR = GPMPC_RS_RI;
#pragma unroll 1
            for (int r0 = 0; r0 < PT; r0 += RSR) {
                double zi[RSR][D], rs[RSR][EG];
#pragma unroll
                for (int r = 0; r < RSR; ++r) {
#pragma unroll
                    for (int k = 0; k < D; ++k) zi[r][k] = fma(-GP_C(k), xi[(r0 + r) * D + k], GP_CU(k));
#pragma unroll
                    for (int g = 0; g < EG; ++g) rs[r][g] = 0.0;
                }
#pragma unroll 1
                for (int j = 0; j < PTJ; j += CJ) {
                    double zj[CJ][D], cs_[CJ][EG];
#pragma unroll
                    for (int c = 0; c < CJ; ++c)
#pragma unroll
                        for (int k = 0; k < D; ++k) zj[c][k] = fma(-GP_C(k), xj[(j + c) * D + k], GP_CU(k));
#pragma unroll
                    for (int r = 0; r < RSR; ++r) {
#pragma unroll
                        for (int c = 0; c < CJ; ++c) {
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
#pragma unroll
                            for (int g = 0; g < EG; ++g) {
                                const double w = Ws[(size_t)g * PT * PTJ + (r0 + r) * PTJ + j + c] * e;
                                rs[r][g] += w;
                                if (r == 0) cs_[c][g] = w; else cs_[c][g] += w;
#pragma unroll
                                for (int k = 0; k < K2; ++k) acc2[g][k] = fma(w, qq[k], acc2[g][k]);
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < CJ; ++c)
#pragma unroll
                        for (int g = 0; g < EG; ++g)
#pragma unroll
                            for (int k = K1; k < D; ++k) acc1[g][k] = fma(cs_[c][g], zj[c][k], acc1[g][k]);
                }
#pragma unroll
                for (int r = 0; r < RSR; ++r)
#pragma unroll
                    for (int g = 0; g < EG; ++g) {
                        accT[g] += rs[r][g];
#pragma unroll
                        for (int k = K1; k < D; ++k) acc1[g][k] = fma(rs[r][g], zi[r][k], acc1[g][k]);
                    }
            }
            } else
#endif
            {
            // RI-row register micro-tile per column: RI independent q -> q^2 -> sum -> exp chains (15 deep each) and
            // RI x EG x (1 + moments) independent accumulation FMAs per basic block, interleaved by the compiler.
#pragma unroll 1
            for (int r0 = 0; r0 < PT; r0 += RI) {
                double zi[RI][D];
#pragma unroll
                for (int r = 0; r < RI; ++r)
#pragma unroll
                    for (int k = 0; k < D; ++k) zi[r][k] = fma(-GP_C(k), xi[(r0 + r) * D + k], GP_CU(k));
#pragma unroll 1
                for (int j = 0; j < PTJ; ++j) {
                    double zj[D];
#pragma unroll
                    for (int k = 0; k < D; ++k) zj[k] = fma(-GP_C(k), xj[j * D + k], GP_CU(k));
#pragma unroll
                    for (int r = 0; r < RI; ++r) {
                        double q[D], qq[D];
#pragma unroll
                        for (int k = 0; k < D; ++k) { q[k] = zi[r][k] + zj[k]; qq[k] = q[k] * q[k]; }
                        double S = qq[0];
                        if (D >= 4) {                       // pairwise tree: shorter dependency chain
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
#pragma unroll
                        for (int g = 0; g < EG; ++g) {
                            const double w = Ws[(size_t)g * PT * PTJ + (r0 + r) * PTJ + j] * e;
                            accT[g] += w;
                            if (GRAD) {
#pragma unroll
                                for (int k = K1; k < D; ++k) acc1[g][k] = fma(w, q[k], acc1[g][k]);
#pragma unroll
                                for (int k = 0; k < K2; ++k) acc2[g][k] = fma(w, qq[k], acc2[g][k]);
                            }
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
                double *dst = a.part + (((size_t)item * a.E + a.out_idx[g]) * NA) * a.Bpad + b;
                dst[0] = accT[g];
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    dst[(size_t)(1 + k) * a.Bpad] = (GRAD && k >= K1) ? acc1[g][k] : 0.0;
                    dst[(size_t)(1 + D + k) * a.Bpad] = (GRAD && k < K2) ? acc2[g][k] : 0.0;
                }
            }
        }
    }
#ifdef GPMPC_PAIR_TIMING
    if (tid == 0) {
        unsigned long long t_end; unsigned smid;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        const size_t cta = blockIdx.x;
        a.cta_times[cta * 3 + 0] = t_start; a.cta_times[cta * 3 + 1] = t_end; a.cta_times[cta * 3 + 2] = smid;
    }
#endif
}

#undef GP_C
#undef GP_CU

}  // namespace gpmpc
