// One whole moment-matching step for FEW rollouts in ONE launch (a single IPOPT solve evaluates one control
// sequence at a time: B = 1, src/mpc.py:202-255).
//
// mm_pairs_batch maps lanes to rollouts and needs >= 32 of them per warp.  Here lanes map to PAIRS: grid
// (B, P); CTA (b, x) streams its contiguous share of the upper-triangular 32x32 tiles of Wt (tile-major storage:
// one contiguous range of memory) through a TMA bulk-copy + mbarrier ring (full/empty barriers, no CTA-wide
// barrier in the loop), lane <-> column j of the tile,
// warp <-> 8 rows, and every thread keeps the (1+2D)*EG accumulators of rollout b.  Per rollout and step the
// kernel reads EG * n(n+1)/2 * 8 bytes of Wt exactly once (268 MB at n=4096, E=4): with one rollout it is bound
// by HBM, not by the FP64 pipe (41 us vs 34 us per step at n=4096).
//
// The same launch also does what used to be four more kernels per step:
//   * z_i = c*u - c*x_i is formed on the fly from X (per-warp rows; z_j from the X tile that travels with Wt),
//   * the mean sums (uncertainty_prop.py:324-338) over this CTA's slice of the training set,
//   * "last CTA done" finalize: the last CTA of rollout b to arrive (ticket counter) sums the P partials in a
//     fixed order (deterministic), applies the prefactors and writes mean_t, var_t and the tape entry,
//   * and, when all outputs share one lambda group, the constants of step t+1 (prep_step).
// So a B = 1 rollout is H launches instead of 5H.
#pragma once
#include "mm_pairs.cuh"
#include "step_common.cuh"

namespace gpmpc {

// Build-time geometry (defaults = shipped): threads per CTA, CTAs per SM, ring slots per CTA.
#ifndef GPMPC_SINGLE_THREADS
#define GPMPC_SINGLE_THREADS 128
#endif
#ifndef GPMPC_SINGLE_CTAS
#define GPMPC_SINGLE_CTAS 2
#endif
#ifndef GPMPC_SINGLE_STAGES
#define GPMPC_SINGLE_STAGES 3
#endif
constexpr int SINGLE_THREADS = GPMPC_SINGLE_THREADS;
constexpr int SINGLE_CTAS_PER_SM = GPMPC_SINGLE_CTAS;
constexpr int SINGLE_WARPS = SINGLE_THREADS / 32;
constexpr int SINGLE_ROWS = PT / SINGLE_WARPS;       // rows of a tile handled by one thread
constexpr int SINGLE_GROUP = 16;       // CTAs per first-level reduction group
constexpr int SINGLE_STAGES = GPMPC_SINGLE_STAGES;   // ring slots per CTA: STAGES-1 tiles in flight

static_assert(kGroupMax <= SINGLE_WARPS, "the finalize maps one warp to each output of the group");

struct SingleStepArgs {
    const double *Wt[kGroupMax];   // tile-major upper-triangular tiles of the group's outputs (common.cuh)
    const double *beta[kGroupMax];
    int out_idx[kGroupMax];
    const double *X;               // [ld, D]
    const double *cst;             // this group's per-rollout constants [4D][Bpad]: c, c*u, cm, cm*u
    double *spart;                 // [B][P][NV] partial sums (NV = 2*EG*NA: pair sums, then mean sums), then [B][NG][NV] group sums
    int *tickets;                  // [B][1 + NG] arrival counters (NG = ceil(P/16)), zero before the launch and after it
    int ld, ntile, total_tiles;
    // finalize (last CTA of each rollout)
    StepDims d;
    int t;
    const double *us;              // [2D][Bpad] input mean / variances of this step
    const double *hyp;
    double *mu, *var, *tape;
    int want_grad;
    // constants of step t+1 (only when the model has ONE lambda group and t < H)
    int prep_next;
    const double *Uint;            // actions [H*m][Bpad]
    const double *lam_group;       // [D] of group 0
    double *us_w, *cst_w;
    double act_var;
    unsigned long long *dbg;       // optional [B*P][6] globaltimer stamps (GPMPC_STEP_DEBUG=1), else NULL
    // Unequal shares for the two CTAs of an SM (B = 1, grid = 2 x SMs only; else NULL).  The warp scheduler favours the
    // CTA that arrived first: with equal tile ranges it finishes after 46 us, its neighbour after 60 us
    // (profiles/r01c_step_single_cta_timeline.csv).  Every CTA therefore CLAIMS its slice: the first arrival on an SM
    // takes one of the P/2 big slices (tiles_big tiles in total), the second a small one.  claim = this step's counters:
    // [smid % 256] arrivals per SM, [256] big slices handed out, [257] small ones (zeroed once per rollout).  A pool that
    // runs dry (an SM that got one or three CTAs) overflows into the other, so the slices are always a bijection; the
    // partial sums are stored per slice and summed in slice order, i.e. the result does not depend on who took what.
    int *claim; int tiles_big;
    // host side only (launch attribute, not read by the kernel): L2 access-policy window over the weights.  Every step
    // streams the same Wt; a plain LRU keeps none of a working set larger than L2 across steps, a persisting fraction
    // that fits stays resident
    const void *l2_base; size_t l2_bytes; float l2_hit;
};

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define GP_STAMP(i) do { if (a.dbg && tid == 0) a.dbg[((size_t)b * P + cta) * 6 + (i)] = gtime(); } while (0)

template <int D, int EG>
__host__ __device__ constexpr size_t single_stage_doubles() { return (size_t)EG * PT * PT + PT * D; }
template <int D, int EG>
__host__ __device__ constexpr size_t single_smem_bytes() { return SINGLE_STAGES * single_stage_doubles<D, EG>() * sizeof(double); }

__device__ __forceinline__ void mbar_arrive(void *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// GRAD / NS select the accumulated moments exactly as in mm_pairs_batch (mm_pairs.cuh).
template <int D, int EG, int GRAD, int NS>
__global__ void __launch_bounds__(SINGLE_THREADS, SINGLE_CTAS_PER_SM)
mm_step_single(const SingleStepArgs a)
{
    constexpr int K1 = GRAD == 2 ? NS : 0;       // N1_k for k in [K1, D)
    constexpr int K2 = GRAD == 1 ? NS : 0;       // N2_k for k in [0, K2)
    constexpr int NA = 1 + 2 * D;
    constexpr int NV = 2 * EG * NA;
    constexpr size_t STAGE = single_stage_doubles<D, EG>();
    constexpr unsigned STAGE_BYTES = (unsigned)(STAGE * sizeof(double));
    extern __shared__ __align__(128) double smem[];      // [slot][ Wt[EG][32*32] | x_j[32*D] ]
    __shared__ double tab[16];
    __shared__ double cs[4 * D];                         // c, c*u, cm, cm*u of this rollout
    __shared__ double ziw[SINGLE_WARPS][SINGLE_ROWS * D];   // z_i of the 8 rows each warp handles
    __shared__ double red[SINGLE_WARPS][EG * NA];
    __shared__ double fin[NV];
    __shared__ __align__(8) unsigned long long full[SINGLE_STAGES], empty[SINGLE_STAGES];
    __shared__ int s_last, s_slice;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // grid (B, P): the rollout index is the FAST block index, so the CTAs that are resident together work on the
    // same tile ranges for different rollouts and share the Wt tiles through L2 (2 <= B < 64)
    const int b = blockIdx.x;                            // rollout
    const int P = gridDim.y;
    const int cta = blockIdx.y;
    // Programmatic dependent launch: the next step's grid may be scheduled while this one drains; everything up
    // to griddepcontrol.wait touches only data no step kernel writes (the exp table, Wt, X).
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    GP_STAMP(0);                                         // CTA start
    if (tid < 16) tab[tid] = kExp2Tab[tid];
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SINGLE_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], SINGLE_WARPS); }
        mbar_fence_init();
        int slice = cta;
        if (a.claim) {                                   // (counters of THIS step: nothing the preceding grid writes)
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const int half = P >> 1;
            if (atomicAdd(&a.claim[smid & 255], 1) == 0) {
                const int k = atomicAdd(&a.claim[256], 1);
                slice = k < half ? k : half + atomicAdd(&a.claim[257], 1);
            } else {
                const int j = atomicAdd(&a.claim[257], 1);
                slice = j < half ? half + j : atomicAdd(&a.claim[256], 1);
            }
        }
        s_slice = slice;
    }
    __syncthreads();
    const int bx = s_slice;                              // this CTA's slice of the tile list / training set

    double accT[EG], acc1[GRAD ? EG : 1][D], acc2[GRAD ? EG : 1][D];
#pragma unroll
    for (int g = 0; g < EG; ++g) accT[g] = 0.0;
    if (GRAD) {
#pragma unroll
        for (int g = 0; g < EG; ++g)
#pragma unroll
            for (int k = 0; k < D; ++k) acc1[g][k] = acc2[g][k] = 0.0;
    }

    int t_begin = (int)((long long)a.total_tiles * bx / P);
    int t_end = (int)((long long)a.total_tiles * (bx + 1) / P);
    if (a.claim) {
        const int half = P >> 1, sm = bx < half ? bx : bx - half;
        const long long base = bx < half ? 0 : a.tiles_big, cnt = bx < half ? a.tiles_big : a.total_tiles - a.tiles_big;
        t_begin = (int)(base + cnt * sm / half);
        t_end = (int)(base + cnt * (sm + 1) / half);
    }
    int I = 0, J = 0;                                    // tile being consumed
    {
        int rem = t_begin, row = 0;
        while (rem >= a.ntile - row) { rem -= a.ntile - row; ++row; }
        I = row; J = row + rem;
    }
    int Ii = I, Ji = J, issued = t_begin;                // next tile to issue (thread 0 only); Ii is not needed for Wt
    (void)Ii;
    auto issue_next = [&]() {                            // thread 0: tile `issued` -> slot (issued - t_begin) % STAGES
        const int slot = (issued - t_begin) % SINGLE_STAGES;
        double *base = smem + (size_t)slot * STAGE;
        void *bar = &full[slot];
        mbar_expect_tx(bar, STAGE_BYTES);
#pragma unroll
        for (int g = 0; g < EG; ++g)             // tile-major Wt: tile number `issued` is one contiguous 8 KB block
            bulk_load_1d(base + (size_t)g * PT * PT, a.Wt[g] + (size_t)issued * PT * PT, PT * PT * sizeof(double), bar);
        bulk_load_1d(base + (size_t)EG * PT * PT, a.X + (size_t)Ji * PT * D, PT * D * sizeof(double), bar);
        ++issued; ++Ji;
        if (Ji == a.ntile) { ++Ii; Ji = Ii; }
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SINGLE_STAGES - 1; ++s)
            if (issued < t_end) issue_next();
    }

    // this thread's first training point of the mean sums (X and beta are constants of the fit: their global-load
    // latency is paid here, before the dependency wait, instead of on the critical path after it)
    constexpr int RCAP = (int)(STAGE / (EG * NA)) < SINGLE_THREADS ? (int)(STAGE / (EG * NA)) : SINGLE_THREADS;
    const int per = (a.ld + P - 1) / P;
    const int j_begin = bx * per;
    const int j_end = min(a.ld, j_begin + per);
    const int rows = min(max(j_end - j_begin, 0), RCAP);                 // threads that own >= 1 point
    double xpre[D], bpre[EG];
    if (tid < rows) {
#pragma unroll
        for (int k = 0; k < D; ++k) xpre[k] = a.X[(size_t)(j_begin + tid) * D + k];
#pragma unroll
        for (int g = 0; g < EG; ++g) bpre[g] = a.beta[g][j_begin + tid];
    }

    // the first tiles are in flight; now wait until the previous kernel on the stream (the preceding step) is
    // complete and its writes (this step's constants, the ticket counters, the partial buffers) are visible
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    GP_STAMP(1);                                         // previous grid complete
    if (tid < 4 * D) cs[tid] = a.cst[(size_t)tid * a.d.Bpad + b];
    __syncthreads();
    double *mine = a.spart + ((size_t)b * P + bx) * NV;

    // ---- mean sums over this CTA's slice of the training set (lanes <-> training points), done while the first
    // tiles land.  Scratch = the last ring slot (the prologue fills slots 0 .. STAGES-2 only).
    // p_k = cm_k (u_k - x_jk),  l_j = exp(-sum p_k^2),  M0 += beta_j l_j, M1_k += beta_j l_j p_k, M2_k += .. p_k^2
    {
        double *scratch = smem + (size_t)(SINGLE_STAGES - 1) * STAGE;    // [thread][EG*NA]
        if (tid < rows) {
            double m0[EG], m1[EG][D], m2[EG][D];
#pragma unroll
            for (int g = 0; g < EG; ++g) {
                m0[g] = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) m1[g][k] = m2[g][k] = 0.0;
            }
            for (int j = j_begin + tid; j < j_end; j += rows) {
                if (j != j_begin + tid) {                // only for training sets with more than P * RCAP points
#pragma unroll
                    for (int k = 0; k < D; ++k) xpre[k] = a.X[(size_t)j * D + k];
#pragma unroll
                    for (int g = 0; g < EG; ++g) bpre[g] = a.beta[g][j];
                }
                double p[D], pp[D], S = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    p[k] = fma(-cs[2 * D + k], xpre[k], cs[3 * D + k]);
                    pp[k] = p[k] * p[k];
                    S += pp[k];
                }
                const double l = exp_neg(S, tab);
#pragma unroll
                for (int g = 0; g < EG; ++g) {
                    const double w = bpre[g] * l;
                    m0[g] += w;
#pragma unroll
                    for (int k = 0; k < D; ++k) { m1[g][k] = fma(w, p[k], m1[g][k]); m2[g][k] = fma(w, pp[k], m2[g][k]); }
                }
            }
#pragma unroll
            for (int g = 0; g < EG; ++g) {
                scratch[(size_t)tid * (EG * NA) + g * NA] = m0[g];
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    scratch[(size_t)tid * (EG * NA) + g * NA + 1 + k] = m1[g][k];
                    scratch[(size_t)tid * (EG * NA) + g * NA + 1 + D + k] = m2[g][k];
                }
            }
        }
        __syncthreads();
        if (tid < EG * NA) {
            double sacc = 0.0;
            for (int r = 0; r < rows; ++r) sacc += scratch[(size_t)r * (EG * NA) + tid];
            mine[EG * NA + tid] = sacc;
        }
        __syncthreads();                                 // the slot is free again before tile STAGES-1 is issued into it
    }

    GP_STAMP(2);                                         // mean sums done, tile loop starts
    int curI = -1;
    for (int t = t_begin; t < t_end; ++t) {
        const int it = t - t_begin;
        const int slot = it % SINGLE_STAGES;
        // refill: tile t + STAGES - 1 goes into the slot tile t-1 used, once all four warps have released it
        if (tid == 0 && issued < t_end) {
            if (it > 0) mbar_wait(&empty[(it - 1) % SINGLE_STAGES], ((it - 1) / SINGLE_STAGES) & 1);
            issue_next();
        }
        if (I != curI) {                                 // new row block: this warp's z_i = c*u - c*x_i (rare)
            __syncwarp();
            for (int idx = lane; idx < SINGLE_ROWS * D; idx += 32) {
                const int m = idx / D, k = idx % D;
                ziw[wid][idx] = fma(-cs[k], a.X[(size_t)(I * PT + wid + m * SINGLE_WARPS) * D + k], cs[D + k]);
            }
            __syncwarp();
            curI = I;
        }
        mbar_wait(&full[slot], (it / SINGLE_STAGES) & 1);

        const double *Ws = smem + (size_t)slot * STAGE;
        const double *xjs = Ws + (size_t)EG * PT * PT;
        double zj[D];
#pragma unroll
        for (int k = 0; k < D; ++k) zj[k] = fma(-cs[k], xjs[lane * D + k], cs[D + k]);
#pragma unroll
        for (int m = 0; m < SINGLE_ROWS; ++m) {
            const int r = wid + m * SINGLE_WARPS;
            double q[D], qq[D];
#pragma unroll
            for (int k = 0; k < D; ++k) { q[k] = ziw[wid][m * D + k] + zj[k]; qq[k] = q[k] * q[k]; }
            double S = qq[0];
            if (D >= 4) {                                // pairwise tree: shorter dependency chain
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
                const double w = Ws[(size_t)g * PT * PT + r * PT + lane] * e;
                accT[g] += w;
                if (GRAD) {
#pragma unroll
                    for (int k = K1; k < D; ++k) acc1[g][k] = fma(w, q[k], acc1[g][k]);
#pragma unroll
                    for (int k = 0; k < K2; ++k) acc2[g][k] = fma(w, qq[k], acc2[g][k]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);        // this warp is done with the slot
        ++J;
        if (J == a.ntile) { ++I; J = I; }
    }

    GP_STAMP(3);                                         // tile loop done
    // ---- CTA reduction of the pair sums in a fixed order: lanes (xor tree), then warps in index order ----
    auto warp_sum = [](double v) {
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
#pragma unroll
    for (int g = 0; g < EG; ++g) {
        double v = warp_sum(accT[g]);
        if (lane == 0) red[wid][g * NA] = v;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const double v1 = (GRAD && k >= K1) ? warp_sum(acc1[g][k]) : 0.0;
            const double v2 = (GRAD && k < K2) ? warp_sum(acc2[g][k]) : 0.0;
            if (lane == 0) { red[wid][g * NA + 1 + k] = v1; red[wid][g * NA + 1 + D + k] = v2; }
        }
    }
    __syncthreads();
    if (tid < EG * NA) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < SINGLE_WARPS; ++w) s += red[w][tid];
        mine[tid] = s;
    }

    // ---- two-level "last one done" reduction (fixed order => deterministic): the last CTA of each group of
    // SINGLE_GROUP consecutive CTAs sums the group's partials, the last group to finish sums the group sums and
    // finalizes.  Only ~SINGLE_GROUP + P/SINGLE_GROUP loads per value sit on the critical path, all in flight at once.
    const int NG = (P + SINGLE_GROUP - 1) / SINGLE_GROUP;
    const int gidx = bx / SINGLE_GROUP;
    const int g0 = gidx * SINGLE_GROUP;
    const int gsize = min(SINGLE_GROUP, P - g0);
    int *tk = a.tickets + (size_t)b * (1 + NG);         // [0]: groups done, [1 + g]: CTAs of group g done
    double *gsum = a.spart + (size_t)a.d.B * P * NV + ((size_t)b * NG) * NV;   // [NG][NV] group sums of rollout b
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&tk[1 + gidx], 1) == gsize - 1);
    __syncthreads();
    GP_STAMP(4);                                         // partial written, first ticket drawn
    if (!s_last) return;
    __threadfence();
    {
        const double *src = a.spart + ((size_t)b * P + g0) * NV;
        for (int v = tid; v < NV; v += SINGLE_THREADS) {
            double x[SINGLE_GROUP];
#pragma unroll
            for (int q = 0; q < SINGLE_GROUP; ++q) x[q] = q < gsize ? __ldcg(src + (size_t)q * NV + v) : 0.0;
            double sacc = 0.0;
#pragma unroll
            for (int q = 0; q < SINGLE_GROUP; ++q) sacc += x[q];
            gsum[(size_t)gidx * NV + v] = sacc;
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) { tk[1 + gidx] = 0; s_last = (atomicAdd(&tk[0], 1) == NG - 1); }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int v = tid; v < NV; v += SINGLE_THREADS) {
        double sacc = 0.0;
        for (int q0 = 0; q0 < NG; q0 += SINGLE_GROUP) {
            double x[SINGLE_GROUP];
#pragma unroll
            for (int q = 0; q < SINGLE_GROUP; ++q) x[q] = q0 + q < NG ? __ldcg(gsum + (size_t)(q0 + q) * NV + v) : 0.0;
#pragma unroll
            for (int q = 0; q < SINGLE_GROUP; ++q) sacc += x[q];
        }
        fin[v] = sacc;
    }
    __syncthreads();
    if (wid < EG)                                        // warp <-> output, lane <-> input dimension
        finalize_math_lanes(a.d, a.t, a.out_idx[wid], b, lane, &fin[wid * NA], &fin[EG * NA + wid * NA], a.us, a.hyp,
                            a.mu, a.var, a.tape, a.want_grad);
    if (tid == 0) tk[0] = 0;                             // counters are zero again for the next launch on this stream
    if (a.prep_next) {
        __syncthreads();                                 // mean_t / var_t of all outputs are written
        if (tid < D) {
            const int k = tid, E = a.d.E;
            double u, s;
            if (k < E) {
                u = a.mu[((size_t)a.t * E + k) * a.d.Bpad + b];
                s = a.var[((size_t)a.t * E + k) * a.d.Bpad + b];
            } else {
                u = a.Uint[((size_t)a.t * a.d.m + (k - E)) * a.d.Bpad + b];
                s = a.act_var;
            }
            a.us_w[(size_t)k * a.d.Bpad + b] = u;
            a.us_w[(size_t)(D + k) * a.d.Bpad + b] = s;
            double c, cu, cm, cmu;
            step_constants(u, s, a.lam_group[k], c, cu, cm, cmu);
            a.cst_w[(size_t)k * a.d.Bpad + b] = c;
            a.cst_w[(size_t)(D + k) * a.d.Bpad + b] = cu;
            a.cst_w[(size_t)(2 * D + k) * a.d.Bpad + b] = cm;
            a.cst_w[(size_t)(3 * D + k) * a.d.Bpad + b] = cmu;
        }
    }
    GP_STAMP(5);                                         // last CTA: finalize and next-step constants done
}
#undef GP_STAMP

}  // namespace gpmpc
