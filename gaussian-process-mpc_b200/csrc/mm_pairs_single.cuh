// Pair-sum kernel for FEW rollouts (a single IPOPT solve evaluates one control sequence at a time).
//
// mm_pairs_batch maps lanes to rollouts and needs >= 32 of them per warp.  Here lanes map to PAIRS: a CTA
// streams its share of the upper-triangular 32x32 tiles of Wt straight from global memory (every element is used
// exactly once per rollout, so there is nothing to stage), lane <-> column j of the tile, warp <-> rows, and each
// thread keeps the (1+2D)*EG accumulators of ONE rollout (blockIdx.y).  Per rollout and step this kernel reads
// EG * n(n+1)/2 * 8 bytes of Wt (268 MB at n=4096, E=4) and does the same FP64 work as the batched kernel, so it
// sits at the crossover of the HBM roof and the FP64-pipe roof (~41 us vs ~45 us per step at n=4096).
// Partial sums go to the same [partial][E][1+2D][Bpad] layout, reduced in a fixed order by finalize_step.
#pragma once
#include "mm_pairs.cuh"

namespace gpmpc {

constexpr int SINGLE_THREADS = 128;    // 2 CTAs per SM: one CTA's tile barrier overlaps the other's compute
constexpr int SINGLE_WARPS = SINGLE_THREADS / 32;
constexpr int SINGLE_ROWS = PT / SINGLE_WARPS;       // rows of a tile handled by one thread (4)

// z[g][b][i][k] = c_k u_k - c_k x_ik for every training point (the scaled offsets the pair kernel adds up);
// computed once per step so that mm_pairs_single can prefetch them with cp.async like any other operand.
static __global__ void zprep_kernel(const double *__restrict__ X, int ld, int D, int B, int Bpad, int G,
                             const double *__restrict__ cst, double *__restrict__ zall)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // over ld * D
    const int b = blockIdx.y, g = blockIdx.z;
    if (idx >= ld * D) return;
    const int k = idx % D;
    const double *cg = cst + (size_t)g * 4 * D * Bpad;
    const double c = cg[(size_t)k * Bpad + b], cu = cg[(size_t)(D + k) * Bpad + b];
    zall[((size_t)g * B + b) * ld * D + idx] = fma(-c, X[idx], cu);
}

constexpr int SINGLE_STAGES = 3;     // cp.async ring per CTA: 2 tiles in flight (x2 CTAs per SM = ~140 KB per SM)

template <int D, int EG>
__host__ __device__ constexpr size_t single_stage_doubles() { return (size_t)EG * PT * PT + 2 * PT * D; }
template <int D, int EG>
__host__ __device__ constexpr size_t single_smem_bytes() { return SINGLE_STAGES * single_stage_doubles<D, EG>() * sizeof(double); }

template <int D, int EG, bool GRAD>
__global__ void __launch_bounds__(SINGLE_THREADS, 2) mm_pairs_single(const PairArgs a)
{
    constexpr int NA = 1 + 2 * D;
    constexpr size_t STAGE = single_stage_doubles<D, EG>();
    extern __shared__ __align__(16) double smem[];      // [stage][ Wt[EG][32*32] | z_i[32*D] | z_j[32*D] ]
    __shared__ double tab[16];
    __shared__ double red[SINGLE_WARPS][EG * NA];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int b = blockIdx.y;                           // rollout
    if (tid < 16) tab[tid] = kExp2Tab[tid];
    const double *zb = a.zall + (size_t)b * a.ld * D;   // this rollout's z

    double accT[EG], acc1[GRAD ? EG : 1][D], acc2[GRAD ? EG : 1][D];
#pragma unroll
    for (int g = 0; g < EG; ++g) accT[g] = 0.0;
    if (GRAD) {
#pragma unroll
        for (int g = 0; g < EG; ++g)
#pragma unroll
            for (int k = 0; k < D; ++k) acc1[g][k] = acc2[g][k] = 0.0;
    }

    const int P = gridDim.x;
    const int t_begin = (int)((long long)a.total_tiles * blockIdx.x / P);
    const int t_end = (int)((long long)a.total_tiles * (blockIdx.x + 1) / P);
    int Ii = 0, Ji = 0;                                 // tile coordinates of the next tile to ISSUE
    {
        int rem = t_begin, row = 0;
        while (rem >= a.ntile - row) { rem -= a.ntile - row; ++row; }
        Ii = row; Ji = row + rem;
    }
    auto issue = [&](int slot, int ti, int tj) {
        double *base = smem + (size_t)slot * STAGE;
#pragma unroll
        for (int g = 0; g < EG; ++g) {
            const double *src = a.Wt[g] + (size_t)ti * PT * a.ld + (size_t)tj * PT;
            double *dst = base + (size_t)g * PT * PT;
#pragma unroll
            for (int q = 0; q < (PT * PT / 2) / SINGLE_THREADS; ++q) {
                const int chunk = tid + q * SINGLE_THREADS;    // 16-byte chunk id: 16 per row
                const int r = chunk >> 4, cc = (chunk & 15) * 2;
                cpa16(dst + r * PT + cc, src + (size_t)r * a.ld + cc);
            }
        }
        double *zi = base + (size_t)EG * PT * PT;
        for (int chunk = tid; chunk < PT * D; chunk += SINGLE_THREADS) {   // z_i then z_j, PT*D/2 chunks each
            const int which = chunk / (PT * D / 2), cc = (chunk % (PT * D / 2)) * 2;
            cpa16(zi + which * PT * D + cc, zb + (size_t)(which ? tj : ti) * PT * D + cc);
        }
    };
    // prologue: put STAGES-1 tiles in flight (empty commits keep the group count uniform)
    int issued = t_begin;
#pragma unroll
    for (int s = 0; s < SINGLE_STAGES - 1; ++s) {
        if (issued < t_end) {
            issue(s, Ii, Ji);
            ++issued; ++Ji;
            if (Ji == a.ntile) { ++Ii; Ji = Ii; }
        }
        cpa_commit();
    }

    for (int t = t_begin; t < t_end; ++t) {
        cpa_wait<SINGLE_STAGES - 2>();                   // tile t has landed (this thread's copies)
        __syncthreads();                                 // ... everyone's copies; and tile t-1 is fully consumed
        if (issued < t_end) {                            // refill the slot tile t-1 used
            issue((issued - t_begin) % SINGLE_STAGES, Ii, Ji);
            ++issued; ++Ji;
            if (Ji == a.ntile) { ++Ii; Ji = Ii; }
        }
        cpa_commit();

        const double *Ws = smem + (size_t)((t - t_begin) % SINGLE_STAGES) * STAGE;
        const double *zi = Ws + (size_t)EG * PT * PT, *zjs = zi + PT * D;
        double zj[D];
#pragma unroll
        for (int k = 0; k < D; ++k) zj[k] = zjs[lane * D + k];
#pragma unroll
        for (int m = 0; m < SINGLE_ROWS; ++m) {
            const int r = wid + m * SINGLE_WARPS;
            double q[D], qq[D];
#pragma unroll
            for (int k = 0; k < D; ++k) { q[k] = zi[r * D + k] + zj[k]; qq[k] = q[k] * q[k]; }
            double S = qq[0];
#pragma unroll
            for (int k = 1; k < D; ++k) S += qq[k];
            const double e = exp_neg(S, tab);
#pragma unroll
            for (int g = 0; g < EG; ++g) {
                const double w = Ws[(size_t)g * PT * PT + r * PT + lane] * e;
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
    cpa_wait<0>();

    // CTA reduction in a fixed order: lanes (xor tree), then warps in index order
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
            const double v1 = warp_sum(GRAD ? acc1[g][k] : 0.0);
            const double v2 = warp_sum(GRAD ? acc2[g][k] : 0.0);
            if (lane == 0) { red[wid][g * NA + 1 + k] = v1; red[wid][g * NA + 1 + D + k] = v2; }
        }
    }
    __syncthreads();
    if (tid < EG * NA) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < SINGLE_WARPS; ++w) s += red[w][tid];
        const int g = tid / NA, e = tid % NA;
        a.part[(((size_t)blockIdx.x * a.E + a.out_idx[g]) * NA + e) * a.Bpad + b] = s;
    }
}

}  // namespace gpmpc
