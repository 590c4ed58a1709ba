// The whole horizon of a SINGLE rollout (B = 1: one IPOPT callback, src/mpc.py:202-255) in ONE persistent cooperative
// launch.
//
// mm_step_single (one launch per horizon step) spends ~16 % of every step outside the tile loop: launch ramp, the
// dependency wait, the "last CTA" tail (profiles/r01d_summary.md).  Here the grid (B, P) stays resident for all H steps:
//   * every CTA owns the same contiguous share of the tile-major Wt in every step, so its TMA ring simply keeps
//     running across the step boundary: while the step's reduction chain (partials -> group sums -> finalize -> next
//     constants) runs in the last-arriving CTA, all other CTAs already have the first tiles of the next step in
//     shared memory;
//   * the hand-off is a per-rollout step counter in global memory: the finalizing CTA publishes the constants of step
//     t+1 and then stores the counter with release semantics; the other CTAs spin on it with acquire loads;
//   * the training points of the mean sums (14 per CTA at n = 4096) stay in registers for the whole rollout.
// Same tile ranges, same pair accumulation order and the same fixed-order reductions as the step-by-step path (only the
// few mean-sum points of a CTA are combined by a warp tree instead of serially): deterministic, and equal to that path
// to rounding (tests/test_gpu_parity.py).  Everything a step writes for the next one (constants, input variances) is
// read back with L1-bypassing loads: the L1 of a resident CTA is not coherent across steps.
//
// A cooperative launch guarantees that all CTAs are co-resident (the spin waits would deadlock otherwise); a spin that
// exceeds its budget sets an error flag and leaves instead of hanging the GPU.
#pragma once
#include "mm_step_single.cuh"
#include <type_traits>

namespace gpmpc {

struct RolloutSingleArgs {
    const double *Wt[kGroupMax];
    const double *beta[kGroupMax];
    int out_idx[kGroupMax];
    const double *X;               // [ld, D]
    double *cst;                   // [4D][Bpad] constants of the current step (step 1: written by prep_step before the launch)
    double *us;                    // [2D][Bpad] input mean / variances of the current step
    double *spart;                 // [B][P][NV] partials, then [B][NG][NV] group sums
    int *tickets;                  // [B][1 + NG], zero before the launch and after it
    int *step_done;                // [B] steps completed so far (zero before the launch)
    int *error;                    // set to 1 if a spin wait ran out of budget
    int ld, ntile, total_tiles, H;
    StepDims d;
    const double *hyp;
    double *mu, *var, *tape;
    int want_grad, first_mode;     // first_mode: moment selection of step 1 (2 = d/dx0 not needed, else 1)
    const double *Uint;            // actions [H*m][Bpad]
    const double *lam_group;       // [D] of the (single) lambda group
    double act_var;
    // one rollout split over `world` GPUs (gpmpc_split_*): rank r sweeps tiles [T r / world, T (r+1) / world) and the
    // matching slice of the training set; after every step the ranks exchange their sums through peer-mapped
    // mailboxes written from inside the kernel (NVLink P2P stores + a release/acquire flag per rank and step)
    int world, rank;
    long long seq0;                // sequence number of step 0 of this call (identical on all ranks)
    double *peer_mail[kSplitMaxWorld];            // [rank r] -> r's mailbox [kSplitMaxWorld][2][kSplitNV]
    unsigned long long *peer_flags[kSplitMaxWorld];   // [rank r] -> r's flags [kSplitMaxWorld][2]
    unsigned long long *xstamp;    // optional [H][2] %globaltimer stamps around the exchange (gpmpc_set_option "split_timeline")
};

__device__ __forceinline__ int ld_acquire_gpu(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed_gpu(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double *p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];\n" : "=d"(v) : "l"(p) : "memory");
    return v;
}

template <int D, int EG, int NS>
__global__ void __launch_bounds__(SINGLE_THREADS, SINGLE_CTAS_PER_SM)
mm_rollout_single(const RolloutSingleArgs a)
{
    constexpr int NA = 1 + 2 * D;
    constexpr int NV = 2 * EG * NA;
    constexpr size_t STAGE = single_stage_doubles<D, EG>();
    constexpr unsigned STAGE_BYTES = (unsigned)(STAGE * sizeof(double));
    extern __shared__ __align__(128) double smem[];      // [slot][ Wt[EG][32*32] | x_j[32*D] ]
    __shared__ double tab[16];
    __shared__ double cs[4 * D];
    __shared__ double s_in[D];                           // input variances of the current step (for the finalize)
    __shared__ double ziw[SINGLE_WARPS][SINGLE_ROWS * D];
    __shared__ double red[SINGLE_WARPS][EG * NA];
    __shared__ double fin[NV];
    __shared__ __align__(8) unsigned long long full[SINGLE_STAGES], empty[SINGLE_STAGES];
    __shared__ int s_last, s_bail;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int b = blockIdx.x, P = gridDim.y, bx = blockIdx.y;
    if (tid < 16) tab[tid] = kExp2Tab[tid];
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SINGLE_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], SINGLE_WARPS); }
        mbar_fence_init();
        s_bail = 0;
    }
    __syncthreads();

    // this GPU's share of the tile list, then this CTA's share of that
    const long long T_lo = (long long)a.total_tiles * a.rank / a.world, T_hi = (long long)a.total_tiles * (a.rank + 1) / a.world;
    const int t_begin = (int)(T_lo + (T_hi - T_lo) * bx / P);
    const int t_end = (int)(T_lo + (T_hi - T_lo) * (bx + 1) / P);
    const int nt_cta = t_end - t_begin;                  // tiles of this CTA per step
    const int g_total = nt_cta * a.H;                    // tiles over the whole rollout
    int I0 = 0, J0 = 0;                                  // first tile of the share
    {
        int rem = t_begin, row = 0;
        while (rem >= a.ntile - row) { rem -= a.ntile - row; ++row; }
        I0 = row; J0 = row + rem;
    }
    // producer state (thread 0): next tile to issue, as (global sequence number, local index, column tile)
    // (ring positions are advanced incrementally: no division in the tile loop)
    int g_issue = 0, islot = 0;
    int loc_issue = 0, Ii = I0, Ji = J0;
    auto issue_next = [&]() {
        double *base = smem + (size_t)islot * STAGE;
        void *bar = &full[islot];
        mbar_expect_tx(bar, STAGE_BYTES);
        const size_t tile = (size_t)(t_begin + loc_issue);
#pragma unroll
        for (int g = 0; g < EG; ++g)
            bulk_load_1d(base + (size_t)g * PT * PT, a.Wt[g] + tile * PT * PT, PT * PT * sizeof(double), bar);
        bulk_load_1d(base + (size_t)EG * PT * PT, a.X + (size_t)Ji * PT * D, PT * D * sizeof(double), bar);
        ++g_issue; ++loc_issue; ++Ji;
        if (++islot == SINGLE_STAGES) islot = 0;
        if (Ji == a.ntile) { ++Ii; Ji = Ii; }
        if (loc_issue == nt_cta) { loc_issue = 0; Ii = I0; Ji = J0; }     // the next step sweeps the same share again
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < SINGLE_STAGES - 1; ++s)
            if (g_issue < g_total) issue_next();
    }

    // mean sums: this CTA's slice of the training set stays in registers (warp 0, one point per lane and round)
    const int per = (a.ld + a.world * P - 1) / (a.world * P);
    const int j_begin = (a.rank * P + bx) * per;
    const int j_end = min(a.ld, j_begin + per);
    constexpr int MR = 2;                                // rounds held in registers (per <= 64 at P >= ld / 64)
    double xpre[MR][D], bpre[MR][EG];
    if (wid == 0) {
#pragma unroll
        for (int r = 0; r < MR; ++r) {
            const int j = j_begin + r * 32 + lane;
            const bool ok = j < j_end;
#pragma unroll
            for (int k = 0; k < D; ++k) xpre[r][k] = ok ? a.X[(size_t)j * D + k] : 0.0;
#pragma unroll
            for (int g = 0; g < EG; ++g) bpre[r][g] = ok ? a.beta[g][j] : 0.0;
        }
    }

    const int NG = (P + SINGLE_GROUP - 1) / SINGLE_GROUP;
    const int gidx = bx / SINGLE_GROUP;
    const int g0 = gidx * SINGLE_GROUP;
    const int gsize = min(SINGLE_GROUP, P - g0);
    int *tk = a.tickets + (size_t)b * (1 + NG);
    double *gsum = a.spart + (size_t)a.d.B * P * NV + ((size_t)b * NG) * NV;
    double *mine = a.spart + ((size_t)b * P + bx) * NV;
    auto warp_sum = [](double v) {
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };

    int g_cons = 0;                                      // tiles consumed so far (all threads agree)
    int slot = 0, pslot = 0;                             // ring slot of tile g_cons / of tile g_cons - 1
    unsigned par = 0, ppar = 0;                          // ... and the mbarrier phase parities that go with them
    for (int t = 1; t <= a.H; ++t) {
        // ---- wait until step t-1 is complete (its finalizer published this step's constants) ----
        if (t > 1) {
            if (tid == 0) {
                // relaxed polling (an acquire load per poll would invalidate this SM's L1 under the other resident CTA),
                // one acquire fence once the value is there
                long long spins = 0;
                while (ld_relaxed_gpu(a.step_done + b) < t - 1) {
                    if (++spins > (1ll << 22)) { s_bail = 1; atomicExch(a.error, 1); break; }     // ~2 s: never hang the GPU
                }
                asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
            }
            __syncthreads();
            if (s_bail) return;
        }
        if (tid < 4 * D) cs[tid] = __ldcg(a.cst + (size_t)tid * a.d.Bpad + b);
        else if (tid < 5 * D) s_in[tid - 4 * D] = __ldcg(a.us + (size_t)(D + tid - 4 * D) * a.d.Bpad + b);
        __syncthreads();

        // ---- mean sums over this CTA's slice (uncertainty_prop.py:324-338), warp 0 ----
        if (wid == 0) {
            double m0[EG], m1[EG][D], m2[EG][D];
#pragma unroll
            for (int g = 0; g < EG; ++g) {
                m0[g] = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) m1[g][k] = m2[g][k] = 0.0;
            }
            for (int r = 0; r * 32 < per; ++r) {
                const int j = j_begin + r * 32 + lane;
                double x[D], bt[EG];
                if (r < MR) {
#pragma unroll
                    for (int k = 0; k < D; ++k) x[k] = r == 0 ? xpre[0][k] : xpre[MR - 1][k];
#pragma unroll
                    for (int g = 0; g < EG; ++g) bt[g] = r == 0 ? bpre[0][g] : bpre[MR - 1][g];
                } else {                                  // only for training sets with more than 64 P points
                    const bool ok = j < j_end;
#pragma unroll
                    for (int k = 0; k < D; ++k) x[k] = ok ? a.X[(size_t)j * D + k] : 0.0;
#pragma unroll
                    for (int g = 0; g < EG; ++g) bt[g] = ok ? a.beta[g][j] : 0.0;
                }
                double p[D], pp[D], S = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) { p[k] = fma(-cs[2 * D + k], x[k], cs[3 * D + k]); pp[k] = p[k] * p[k]; S += pp[k]; }
                const double l = exp_neg(S, tab);
#pragma unroll
                for (int g = 0; g < EG; ++g) {
                    const double w = bt[g] * l;              // beta = 0 for padded / out-of-slice lanes
                    m0[g] += w;
#pragma unroll
                    for (int k = 0; k < D; ++k) { m1[g][k] = fma(w, p[k], m1[g][k]); m2[g][k] = fma(w, pp[k], m2[g][k]); }
                }
            }
#pragma unroll
            for (int g = 0; g < EG; ++g) {
                const double v0 = warp_sum(m0[g]);
                if (lane == 0) mine[EG * NA + g * NA] = v0;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const double v1 = warp_sum(m1[g][k]), v2 = warp_sum(m2[g][k]);
                    if (lane == 0) { mine[EG * NA + g * NA + 1 + k] = v1; mine[EG * NA + g * NA + 1 + D + k] = v2; }
                }
            }
        }

        // ---- tile loop: the moments this step's adjoint needs (mm_pairs.cuh: GRAD 1 / 2) ----
        const int mode = !a.want_grad ? 0 : ((t == 1 && a.first_mode == 2) ? 2 : 1);
        double accT[EG], acc1[EG][D], acc2[EG][D];
#pragma unroll
        for (int g = 0; g < EG; ++g) {
            accT[g] = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) acc1[g][k] = acc2[g][k] = 0.0;
        }
        auto tile_loop = [&](auto mode_c) {
        constexpr int MODE = decltype(mode_c)::value;
        int I = I0, J = J0, curI = -1;
        for (int it = 0; it < nt_cta; ++it) {
            // refill: tile g_cons + STAGES - 1 goes into the slot tile g_cons - 1 used, once all warps have released it
            if (tid == 0 && g_issue < g_total) {
                if (g_cons > 0) mbar_wait(&empty[pslot], ppar);
                issue_next();
            }
            if (I != curI) {
                __syncwarp();
                for (int idx = lane; idx < SINGLE_ROWS * D; idx += 32) {
                    const int m = idx / D, k = idx % D;
                    ziw[wid][idx] = fma(-cs[k], a.X[(size_t)(I * PT + wid + m * SINGLE_WARPS) * D + k], cs[D + k]);
                }
                __syncwarp();
                curI = I;
            }
            mbar_wait(&full[slot], par);
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
                    const double w = Ws[(size_t)g * PT * PT + r * PT + lane] * e;
                    accT[g] += w;
                    if (MODE == 1) {
#pragma unroll
                        for (int k = 0; k < D; ++k) acc1[g][k] = fma(w, q[k], acc1[g][k]);
#pragma unroll
                        for (int k = 0; k < NS; ++k) acc2[g][k] = fma(w, qq[k], acc2[g][k]);
                    } else if (MODE == 2) {
#pragma unroll
                        for (int k = NS; k < D; ++k) acc1[g][k] = fma(w, q[k], acc1[g][k]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
            ++J;
            if (J == a.ntile) { ++I; J = I; }
            ++g_cons; pslot = slot; ppar = par;
            if (++slot == SINGLE_STAGES) { slot = 0; par ^= 1u; }
        }
        };
        if (mode == 1) tile_loop(std::integral_constant<int, 1>{});
        else if (mode == 2) tile_loop(std::integral_constant<int, 2>{});
        else tile_loop(std::integral_constant<int, 0>{});

        // ---- CTA reduction in a fixed order: lanes (xor tree), then warps in index order ----
#pragma unroll
        for (int g = 0; g < EG; ++g) {
            double v = warp_sum(accT[g]);
            if (lane == 0) red[wid][g * NA] = v;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const bool n1 = mode == 1 || (mode == 2 && k >= NS), n2 = mode == 1 && k < NS;
                const double v1 = n1 ? warp_sum(acc1[g][k]) : 0.0;
                const double v2 = n2 ? warp_sum(acc2[g][k]) : 0.0;
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

        // ---- two-level "last one done" reduction (fixed order), finalize, constants of step t+1, publish ----
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(&tk[1 + gidx], 1) == gsize - 1);
        __syncthreads();
        if (!s_last) continue;
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
        if (!s_last) continue;
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
        if (a.world > 1) {
            // ---- exchange with the other GPUs: publish this GPU's sums to every rank's mailbox, collect theirs ----
            const unsigned long long seq = (unsigned long long)(a.seq0 + t);
            const int mslot = (int)(seq & 1);
            if (a.xstamp && tid == 0) a.xstamp[(size_t)(t - 1) * 2] = gtime();
            for (int r = 0; r < a.world; ++r)
                for (int v = tid; v < NV; v += SINGLE_THREADS)
                    a.peer_mail[r][((size_t)a.rank * 2 + mslot) * kSplitNV + v] = fin[v];
            __threadfence_system();
            __syncthreads();
            if (tid < a.world) {
                st_release_sys(a.peer_flags[tid] + a.rank * 2 + mslot, seq);
                long long spins = 0;
                while (ld_acquire_sys(a.peer_flags[a.rank] + tid * 2 + mslot) < seq) {
                    if (++spins > (1ll << 22)) { atomicExch(a.error, 1); break; }                 // a missing peer: ~2-4 s, then NaN
                }
            }
            __syncthreads();
            const double *mail = a.peer_mail[a.rank];
            for (int v = tid; v < NV; v += SINGLE_THREADS) {
                double sacc = 0.0;
                for (int r = 0; r < a.world; ++r) sacc += ld_relaxed_sys(mail + ((size_t)r * 2 + mslot) * kSplitNV + v);   // rank order: same on every GPU
                fin[v] = sacc;
            }
            __syncthreads();
            if (a.xstamp && tid == 0) a.xstamp[(size_t)(t - 1) * 2 + 1] = gtime();
        }
        if (wid < EG)
            finalize_math_lanes(a.d, t, a.out_idx[wid], b, lane, &fin[wid * NA], &fin[EG * NA + wid * NA], a.us, a.hyp,
                                a.mu, a.var, a.tape, a.want_grad, s_in);
        if (tid == 0) tk[0] = 0;
        __syncthreads();                                 // mean_t / var_t of all outputs are written (same CTA)
        if (t < a.H && tid < D) {
            const int k = tid, E = a.d.E;
            double u, s;
            if (k < E) {
                u = a.mu[((size_t)t * E + k) * a.d.Bpad + b];
                s = a.var[((size_t)t * E + k) * a.d.Bpad + b];
            } else {
                u = a.Uint[((size_t)t * a.d.m + (k - E)) * a.d.Bpad + b];
                s = a.act_var;
            }
            a.us[(size_t)k * a.d.Bpad + b] = u;
            a.us[(size_t)(D + k) * a.d.Bpad + b] = s;
            double c, cu, cm, cmu;
            step_constants(u, s, a.lam_group[k], c, cu, cm, cmu);
            a.cst[(size_t)k * a.d.Bpad + b] = c;
            a.cst[(size_t)(D + k) * a.d.Bpad + b] = cu;
            a.cst[(size_t)(2 * D + k) * a.d.Bpad + b] = cm;
            a.cst[(size_t)(3 * D + k) * a.d.Bpad + b] = cmu;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release_gpu(a.step_done + b, t);
    }
}

}  // namespace gpmpc
