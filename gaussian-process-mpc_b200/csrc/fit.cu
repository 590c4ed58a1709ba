// GP fit on the device: Gram matrix, blocked Cholesky, Ky^-1, beta = Ky^-1 y and the precomputed
// moment-matching weight matrix.  Replaces GaussianProcessRegression.build_Ky_inv_mat
// (reference src/gpr.py:159-171: Kf, Ky, explicit LU inverse) with
//     Ky = L L^T  (right-looking blocked Cholesky, DMMA trailing updates)
//     Z  = L^-1   (blocked forward substitution, DMMA panels), Ky^-1 = Z^T Z (DMMA, symmetric)
// The explicit inverse is kept because the variance formula contracts a matrix that changes every step
// against the full Ky^-1 (src/tools/uncertainty_prop.py:399).
#include "common.cuh"

namespace gpmpc {

constexpr int NB = 64;   // Cholesky block size (== dgemm tile)

// ---------------------------------------------------------------------------------------------
// Gram matrix.  K[i][j] = sf^2 exp(-1/2 sum_k (x_ik - x_jk)^2 / lam_k) (+ noise on the diagonal).
// Rows/cols >= n are padded with the identity so that the padded matrix stays positive definite.
// ---------------------------------------------------------------------------------------------
struct HyperArg { double inv_lam[kMaxD]; double sf2; double noise; };

__global__ void gram_kernel(const double *__restrict__ X, int n, int np, int D, HyperArg hp,
                            double *__restrict__ K, int ldk, int add_noise)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= np || j >= np) return;
    double v;
    if (i < n && j < n) {
        double q = 0.0;
        for (int k = 0; k < D; ++k) {
            const double d = X[(size_t)i * D + k] - X[(size_t)j * D + k];
            q = fma(d * d, hp.inv_lam[k], q);
        }
        v = hp.sf2 * exp(-0.5 * q);
        if (add_noise && i == j) v += hp.noise;
    } else {
        v = (i == j) ? 1.0 : 0.0;
    }
    K[(size_t)i * ldk + j] = v;
}

static HyperArg make_hyper(const gpmpc_ctx *h, int a, bool prop)
{
    HyperArg hp;
    for (int k = 0; k < kMaxD; ++k) hp.inv_lam[k] = 0.0;
    for (int k = 0; k < h->D; ++k) hp.inv_lam[k] = 1.0 / (prop ? h->lam_prop[a][k] : h->lam_fit[a][k]);
    const double sf = prop ? h->sf_prop[a] : h->sf_fit[a];
    hp.sf2 = sf * sf;
    hp.noise = h->noise[a];
    return hp;
}

int gram_into(gpmpc_ctx *h, int a, double *dst, int ldd, bool add_noise)
{
    // exported matrices (n x n, ld = ldd) use the fit-time hyper-parameters, like self.Kf / self.Ky
    dim3 blk(32, 8), grid((h->n + 31) / 32, (h->n + 7) / 8);
    gram_kernel<<<grid, blk, 0, h->stream>>>(h->X.as<double>(), h->n, h->n, h->D, make_hyper(h, a, false), dst,
                                             ldd, add_noise ? 1 : 0);
    GP_LAUNCH_CHECK(h);
    return GPMPC_OK;
}

// ---------------------------------------------------------------------------------------------
// Diagonal block: Cholesky of a 64x64 block and the inverse of its factor, one CTA, matrix in REGISTERS.
// 256 threads form a 16x16 grid; thread (ty, tx) owns the 4x4 sub-block rows 4ty.., columns 4tx.. (threads above the
// diagonal idle along).  Both sweeps advance FOUR columns / rows per barrier: the 64x4 panel is published through a
// double-buffered shared array, every thread factorises the 4x4 diagonal tile redundantly (4 dependent rsqrt chains
// instead of 64 barrier + rsqrt + publish round trips), solves its own rows against it and applies a rank-4 update
// (64 FMAs) to its tile.  (History: shared-memory resident 63 us, registers with one column per barrier 46 us.)
// Writes
//   A    : L_kk in place (upper part of the block zeroed),
//   Linv : L_kk^-1 dense row-major (the panel solve A[i,k] L_kk^-T is then a tensor-core GEMM),
//   ZT   : (L_kk^-1)^T into the diagonal block of L^-T (leaf of the recursive triangular inverse).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void load4(const double *p, double (&v)[4])
{
    const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4(double *p, const double (&v)[4])
{
    *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2 *>(p + 2) = make_double2(v[2], v[3]);
}

__global__ void __launch_bounds__(256)
potrf_diag_kernel(double *__restrict__ A, int ld, int k0, double *__restrict__ Linv, double *__restrict__ ZT,
                  int ldz, int *info)
{
    __shared__ __align__(32) double pan[2][NB * 4];     // published 64x4 panel (rows x 4 columns), double buffered
    __shared__ __align__(32) double rowp[2][4 * NB];    // inverse: published 4 rows of X
    __shared__ __align__(32) double diag[16][16];       // the 16 diagonal 4x4 tiles of L (for the inverse sweep)
    __shared__ double rinv[NB];                         // 1 / L_cc
    __shared__ int bad;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    if (tid == 0) bad = 0;
#ifdef GPMPC_POTRF_TIMING
    long long tstamp[5]; tstamp[0] = clock64();
#define POTRF_STAMP(i) tstamp[i] = clock64()
#else
#define POTRF_STAMP(i)
#endif
    double a[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = 4 * ty + i, c = 4 * tx + j;
            a[i][j] = (c <= r) ? A[(size_t)(k0 + r) * ld + k0 + c] : 0.0;
        }
    __syncthreads();
    POTRF_STAMP(1);
    // ---- Cholesky, right-looking, four columns per step ----
#pragma unroll 1
    for (int tc = 0; tc < 16; ++tc) {
        double *buf = pan[tc & 1];
        if (tx == tc) {
#pragma unroll
            for (int i = 0; i < 4; ++i) store4(&buf[(4 * ty + i) * 4], a[i]);
        }
        __syncthreads();
        // the 4x4 diagonal tile, factorised by every thread
        double d[4][4], r[4], l[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) load4(&buf[(4 * tc + i) * 4], d[i]);
        bool okall = true;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            double t = d[p][p];
#pragma unroll
            for (int q = 0; q < p; ++q) t = fma(-l[p][q], l[p][q], t);
            const bool ok = t > 0.0;
            if (!ok) { t = 1.0; if (okall && tid == 0 && !bad) bad = k0 + 4 * tc + p + 1; okall = false; }
            r[p] = rsqrt(t);
            l[p][p] = ok ? t * r[p] : 1.0;
#pragma unroll
            for (int i = p + 1; i < 4; ++i) {
                double v = d[i][p];
#pragma unroll
                for (int q = 0; q < p; ++q) v = fma(-l[i][q], l[p][q], v);
                l[i][p] = v * r[p];
            }
        }
        if (tid == 0) {                        // (compile-time indices: r[] must stay in registers)
#pragma unroll
            for (int p = 0; p < 4; ++p) rinv[4 * tc + p] = r[p];
        }
        // rows of the panel this thread needs: its own row block and (as the other factor of the update) its column block
        double lr[4][4], lc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double pr[4], pc[4];
            load4(&buf[(4 * ty + i) * 4], pr);
            load4(&buf[(4 * tx + i) * 4], pc);
#pragma unroll
            for (int p = 0; p < 4; ++p) {      // x L_kk^T = row  ->  forward substitution
                double vr = pr[p], vc = pc[p];
#pragma unroll
                for (int q = 0; q < p; ++q) { vr = fma(-lr[i][q], l[p][q], vr); vc = fma(-lc[i][q], l[p][q], vc); }
                lr[i][p] = vr * r[p];
                lc[i][p] = vc * r[p];
            }
        }
        if (tx == tc) {                        // the owners keep the finished panel (rows above the diagonal: garbage, never read)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int p = 0; p < 4; ++p) a[i][p] = lr[i][p];
        } else if (tx > tc && ty >= tx) {      // rank-4 update of the lower tiles to the right of the panel
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int p = 0; p < 4; ++p) a[i][j] = fma(-lr[i][p], lc[j][p], a[i][j]);
        }
    }
    if (ty == tx) {
#pragma unroll
        for (int i = 0; i < 4; ++i) store4(&diag[ty][4 * i], a[i]);
    }
    __syncthreads();
    POTRF_STAMP(2);
    // ---- X = L^-1, right-looking, four rows per step: X[kb] = L_kk^-1 (I - sums), then pushed into the rows below ----
    double x[4][4];                            // running sums, then X
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) x[i][j] = 0.0;
#pragma unroll 1
    for (int tk = 0; tk < 16; ++tk) {
        double *rows = rowp[tk & 1], *cols = pan[tk & 1];
        if (ty == tk) {                        // owners of the row block: finalise and publish it
            double lk[4][4], rk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { load4(&diag[tk][4 * i], lk[i]); rk[i] = rinv[4 * tk + i]; }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double v[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    double t = ((tx == tk && p == j) ? 1.0 : 0.0) - x[p][j];
#pragma unroll
                    for (int q = 0; q < p; ++q) t = fma(-lk[p][q], v[q], t);
                    v[p] = t * rk[p];
                }
#pragma unroll
                for (int p = 0; p < 4; ++p) x[p][j] = tx <= tk ? v[p] : 0.0;
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) store4(&rows[p * NB + 4 * tx], x[p]);
        }
        if (tx == tk) {                        // owners of the column block of L publish it
#pragma unroll
            for (int i = 0; i < 4; ++i) store4(&cols[(4 * ty + i) * 4], a[i]);
        }
        __syncthreads();
        if (ty > tk && tx <= tk) {
            double xr[4][4];
#pragma unroll
            for (int p = 0; p < 4; ++p) load4(&rows[p * NB + 4 * tx], xr[p]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double li[4];
                load4(&cols[(4 * ty + i) * 4], li);
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int p = 0; p < 4; ++p) x[i][j] = fma(li[p], xr[p][j], x[i][j]);
            }
        }
    }
    POTRF_STAMP(3);
    // write-out through shared memory so that every global store is a coalesced row segment
    __shared__ double stage[NB][NB + 1];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = 4 * ty + i, c = 4 * tx + j;
                stage[r][c] = (c <= r) ? (pass == 0 ? a[i][j] : x[i][j]) : 0.0;      // upper part of the block is zero
            }
        __syncthreads();
        for (int e = tid; e < NB * NB; e += 256) {
            const int r = e / NB, c = e % NB;
            if (pass == 0) A[(size_t)(k0 + r) * ld + k0 + c] = stage[r][c];
            else {
                Linv[e] = stage[r][c];
                ZT[(size_t)(k0 + r) * ldz + k0 + c] = stage[c][r];                   // transposed
            }
        }
    }
    if (tid == 0 && bad) atomicCAS(info, 0, bad);
#ifdef GPMPC_POTRF_TIMING
    POTRF_STAMP(4);
    if (tid == 0) printf("potrf_diag cycles: load %lld, cholesky %lld, inverse %lld, store %lld\n", tstamp[1] - tstamp[0],
                         tstamp[2] - tstamp[1], tstamp[3] - tstamp[2], tstamp[4] - tstamp[3]);
#endif
#undef POTRF_STAMP
}

// copy the lower triangle (tiles computed by the tri_lower GEMM) into the upper triangle
__global__ void mirror_lower_kernel(double *__restrict__ A, int np, int ld)
{
    __shared__ double T[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int r = ty; r < 32; r += blockDim.y) T[r][tx] = A[(size_t)(bi * 32 + r) * ld + bj * 32 + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += blockDim.y) {
        const int i = bj * 32 + r, j = bi * 32 + tx;       // transposed position
        if (j > i) A[(size_t)i * ld + j] = T[tx][r];
    }
}

// y = A[0:n,0:n] x   (row per warp, coalesced, warp-shuffle reduction)
__global__ void gemv_kernel(const double *__restrict__ A, int ld, int n, const double *__restrict__ x,
                            double *__restrict__ y, int np)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= np) return;
    double s = 0.0;
    if (row < n)
        for (int j = lane; j < n; j += 32) s = fma(A[(size_t)row * ld + j], x[j], s);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[row] = (row < n) ? s : 0.0;
}

// ---------------------------------------------------------------------------------------------
// Moment-matching weights.  For the variance (src/tools/uncertainty_prop.py:392-399)
//     T = sum_ij (Ky^-1 - beta beta^T)_ij * exp(-1/4 (x_i-x_j)^T Lam^-1 (x_i-x_j)) * [u,S-dependent part]_ij
// the first two factors depend only on the fit and on lambda; they are folded once into
//     Wt_ij = w(i,j) (Ky^-1 - beta beta^T)_ij exp(-1/4 ...),   w = 2 for j > i, 1 for j == i, 0 for j < i
// so that the per-step kernels sweep only the upper triangle (the [u,S] part is symmetric in i,j).
// ---------------------------------------------------------------------------------------------
__global__ void derive_weights_kernel(const double *__restrict__ X, int n, int D, HyperArg hp,
                                      const double *__restrict__ Kinv, const double *__restrict__ beta,
                                      double *__restrict__ Wt, int ld)
{
    // one block per upper tile (I = blockIdx.y, J = blockIdx.x >= I); writes the tile-major layout of common.cuh
    const int I = blockIdx.y, J = blockIdx.x;
    if (J < I) return;
    const int nt = ld / kPairTile;
    double *tile = Wt + wt_tile_index(I, J, nt) * kPairTile * kPairTile;
    const int c = threadIdx.x;
    const int j = J * kPairTile + c;
    for (int r = threadIdx.y; r < kPairTile; r += blockDim.y) {
        const int i = I * kPairTile + r;
        double v = 0.0;
        if (i < n && j < n && j >= i) {
            double q = 0.0;
            for (int k = 0; k < D; ++k) {
                const double d = X[(size_t)i * D + k] - X[(size_t)j * D + k];
                q = fma(d * d, hp.inv_lam[k], q);
            }
            const double w = Kinv[(size_t)i * ld + j] - beta[i] * beta[j];
            v = (j > i ? 2.0 : 1.0) * w * exp(-0.25 * q);
        }
        tile[r * kPairTile + c] = v;
    }
}

int derive_weights(gpmpc_ctx *h, int a)
{
    const size_t mat = (size_t)h->ld * h->ld;
    const int nt = h->ld / kPairTile;
    dim3 blk(32, 8), grid(nt, nt);
    derive_weights_kernel<<<grid, blk, 0, h->stream>>>(h->X.as<double>(), h->n, h->D, make_hyper(h, a, true),
                                                       h->Kinv.as<double>() + a * mat,
                                                       h->beta.as<double>() + (size_t)a * h->ld,
                                                       h->Wt.as<double>() + a * wt_doubles(h->ld), h->ld);
    GP_LAUNCH_CHECK(h);
    h->weights_epoch++;             // cross-output weights derived from beta / lambda are stale now (fullcov.cu)
    return GPMPC_OK;
}

void rebuild_groups(gpmpc_ctx *h)
{
    h->groups.clear();
    std::vector<bool> used(h->E, false);
    for (int a = 0; a < h->E; ++a) {
        if (used[a]) continue;
        LambdaGroup g;
        for (int b = a; b < h->E && g.count < kGroupMax; ++b) {
            if (used[b]) continue;
            if (std::memcmp(h->lam_prop[a], h->lam_prop[b], sizeof(double) * h->D) == 0) {
                g.outputs[g.count++] = b;
                used[b] = true;
            }
        }
        h->groups.push_back(g);
    }
}

int upload_prop_hypers(gpmpc_ctx *h)
{
    // layout: lam[E,D] | sf[E] | group lambdas [G,D]
    const int E = h->E, D = h->D;
    rebuild_groups(h);
    std::vector<double> host((size_t)E * D + E + h->groups.size() * D);
    for (int a = 0; a < E; ++a) {
        for (int k = 0; k < D; ++k) host[(size_t)a * D + k] = h->lam_prop[a][k];
        host[(size_t)E * D + a] = h->sf_prop[a];
    }
    for (size_t g = 0; g < h->groups.size(); ++g)
        for (int k = 0; k < D; ++k)
            host[(size_t)E * D + E + g * D + k] = h->lam_prop[h->groups[g].outputs[0]][k];
    GP_CUDA(h, h->hyp.reserve(host.size() * sizeof(double)));
    // pageable source: the copy is staged before the call returns, so the vector may die afterwards
    GP_CUDA(h, cudaMemcpyAsync(h->hyp.p, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    GP_CUDA(h, cudaStreamSynchronize(h->stream));
    return GPMPC_OK;
}

// Ky (padded, in `L`) -> L in place; returns GPMPC_ERR_NOT_PD through info.  Also leaves (L_kk^-1)^T in the
// diagonal blocks of ZT.  Two-level right-looking algorithm: inside an outer panel of NB2 columns the 64-wide
// rank updates touch only the panel, the rest of the matrix gets ONE rank-NB2 update per outer step (4x fewer
// passes over the trailing matrix, 16 k-steps per tensor-core tile instead of 4).
constexpr int NB2 = 256;
// Look-ahead (round 2): the trailing update of outer step K is split into (1) the columns of the NEXT outer panel, which
// stay on the factorisation's stream because the next diagonal chain needs them, and (2) the rest of the trailing matrix,
// which runs on a side stream concurrently with that chain (the chain is ~4 x 55 us of tiny dependent kernels per outer
// step; at n = 4096 it used to be followed by an idle-machine wait for a ~60 us GEMM, at n = 16384 the GEMMs dominate and
// the chain hides behind them).  Orders: (2)_K after chain(K) [event]; (1)_K after (2)_{K-1} [event: both write the next
// panel's columns]; (2) of consecutive steps in stream order.
static int cholesky_inplace(gpmpc_ctx *h, double *L, double *ZT, int np, double *linv, int *info, LookAhead *la)
{
    const int ld = h->ld;
    cudaStream_t main_st = h->stream;
    const int n_outer = (np + NB2 - 1) / NB2;
    if (la) while ((int)la->ev.size() < 2 * n_outer) {
        cudaEvent_t e;
        GP_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        la->ev.push_back(e);
    }
    int step = 0;
    bool side_pending = false;
    for (int K0 = 0; K0 < np; K0 += NB2, ++step) {
        const int w = np - K0 < NB2 ? np - K0 : NB2;
        for (int k = K0; k < K0 + w; k += NB) {
            potrf_diag_kernel<<<1, 256, 0, h->stream>>>(L, ld, k, linv, ZT, ld, info);
            GP_LAUNCH_CHECK(h);
            const int rem = np - (k + NB);
            if (rem <= 0) break;
            double *Apan = L + (size_t)(k + NB) * ld + k;
            // panel: A[i,k] <- A[i,k] * Lkk^-T, in place (a CTA reads exactly the 64x64 tile it overwrites, K = 64)
            int rc = dgemm_nt(h, rem, NB, NB, 1.0, Apan, ld, linv, NB, 0.0, Apan, ld, false, 0);
            if (rc) return rc;
            // rank-64 update of the remaining columns of this outer panel (lower tiles only)
            const int wc = K0 + w - (k + NB);
            if (wc > 0) {
                rc = dgemm_nt(h, rem, wc, NB, -1.0, Apan, ld, Apan, ld, 1.0, L + (size_t)(k + NB) * ld + (k + NB), ld, true, 0);
                if (rc) return rc;
            }
        }
        const int rem2 = np - (K0 + w);
        if (rem2 > 0) {
            // trailing update with the whole outer panel: A[i,j] -= sum_p L[i,K0+p] L[j,K0+p], lower tiles only
            double *P = L + (size_t)(K0 + w) * ld + K0;
            double *A22 = L + (size_t)(K0 + w) * ld + (K0 + w);
            const int w2 = rem2 < NB2 ? rem2 : NB2;               // width of the next outer panel
            if (!la || rem2 <= w2) {
                if (la && side_pending) { GP_CUDA(h, cudaStreamWaitEvent(main_st, la->ev[2 * (step - 1) + 1], 0)); side_pending = false; }
                int rc = dgemm_nt(h, rem2, rem2, w, -1.0, P, ld, P, ld, 1.0, A22, ld, true, 0);
                if (rc) return rc;
            } else {
                GP_CUDA(h, cudaEventRecord(la->ev[2 * step], main_st));                 // chain(K) done: the panel is final
                // (2): everything right of the next panel, on the side stream
                GP_CUDA(h, cudaStreamWaitEvent(la->side, la->ev[2 * step], 0));
                h->stream = la->side;
                int rc = dgemm_nt(h, rem2 - w2, rem2 - w2, w, -1.0, P + (size_t)w2 * ld, ld, P + (size_t)w2 * ld, ld, 1.0,
                                  A22 + (size_t)w2 * (ld + 1), ld, true, 0);
                h->stream = main_st;
                if (rc) return rc;
                GP_CUDA(h, cudaEventRecord(la->ev[2 * step + 1], la->side));
                // (1): the next panel's columns, after the previous step's (2) (same columns)
                if (side_pending) GP_CUDA(h, cudaStreamWaitEvent(main_st, la->ev[2 * (step - 1) + 1], 0));
                rc = dgemm_nt(h, rem2, w2, w, -1.0, P, ld, P, ld, 1.0, A22, ld, true, 0);
                if (rc) return rc;
                side_pending = true;
            }
        }
    }
    if (la && side_pending) GP_CUDA(h, cudaStreamWaitEvent(main_st, la->ev[2 * (step - 1) + 1], 0));
    return GPMPC_OK;
}

// out[0] = 2 * sum_{i<n} log L[i][i]  (single block, fixed reduction order)
__global__ void __launch_bounds__(1024) logdet_kernel(const double *__restrict__ L, int ld, int n, double *__restrict__ out)
{
    __shared__ double sm[1024];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) s += log(L[(size_t)i * ld + i]);
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = 2.0 * sm[0];
}

// ZT = L^-T (upper triangular, row-major) by recursive doubling.  The diagonal 64x64 blocks are already there
// (potrf_diag_kernel).  At block size s the pairs of neighbouring diagonal blocks are independent:
//     L = [L11 0; L21 L22]  =>  ZT12 = -ZT11 * L21^T * ZT22
// i.e. two tensor-core GEMMs per pair, batched over all pairs of the level (blockIdx.z); `work` holds the
// intermediate products (needs np*np/4 doubles).
static int invert_factor(gpmpc_ctx *h, const double *L, double *ZT, double *work, int np)
{
    const int ld = h->ld;
    for (int s = NB; s < np; s *= 2) {
        const int nfull = np / (2 * s);                       // pairs with two full blocks
        const int tail0 = nfull * 2 * s;                      // start of a possible ragged pair
        const int s2 = np - tail0 - s;                        // size of its second block (<= 0: none)
        const long long sbig = (long long)2 * s * (ld + 1);
        for (int pass = 0; pass < 2; ++pass) {
            const int batch = pass == 0 ? nfull : (s2 > 0 ? 1 : 0);
            if (batch == 0) continue;
            const int base = pass == 0 ? 0 : tail0;
            const int n2 = pass == 0 ? s : s2;
            const double *ZT11 = ZT + (size_t)base * (ld + 1);
            const double *L21 = L + (size_t)(base + s) * ld + base;
            const double *ZT22 = ZT + (size_t)(base + s) * (ld + 1);
            double *ZT12 = ZT + (size_t)base * ld + base + s;
            // P[i][j] = sum_k ZT11[i][k] L21[j][k]      (ZT11 upper triangular: k >= row0)
            int rc = dgemm_batched(h, false, batch, s, n2, s, 1.0, ZT11, ld, sbig, L21, ld, sbig, 0.0, work, n2,
                                   (long long)s * n2, false, 1);
            if (rc) return rc;
            // ZT12[i][j] = - sum_k P[i][k] ZT22[k][j]   (ZT22 upper triangular: k < col0 + 64)
            rc = dgemm_batched(h, true, batch, s, n2, n2, -1.0, work, n2, (long long)s * n2, ZT22, ld, sbig, 0.0,
                               ZT12, ld, sbig, false, 3);
            if (rc) return rc;
        }
    }
    return GPMPC_OK;
}

int fit_all(gpmpc_ctx *h, const bool *which)
{
    const int ld = h->ld, np = h->ld, n = h->n, E = h->E;
    const size_t mat = (size_t)ld * ld;
    auto same_fit_hypers = [&](int a, int b) {
        for (int k = 0; k < h->D; ++k) if (h->lam_fit[a][k] != h->lam_fit[b][k]) return false;
        return h->sf_fit[a] == h->sf_fit[b] && h->noise[a] == h->noise[b];
    };
    // Outputs share X; with bit-identical kernel hyper-parameters they share Ky and hence Ky^-1 (the reference
    // factorises each of them again, src/gpr.py:159-171): one "leader" per distinct hyper-parameter set is
    // factorised, its twins copy the result.  Leaders are independent problems: they run concurrently on auxiliary
    // streams, so that one output's serial diagonal-block chain overlaps the other outputs' GEMMs.
    int twin[kMaxE];
    std::vector<int> leaders;
    for (int a = 0; a < E; ++a) {
        twin[a] = -1;
        if (!which[a]) continue;
        for (int b = 0; b < a && twin[a] < 0; ++b) if (which[b] && same_fit_hypers(a, b)) twin[a] = b;
        if (twin[a] < 0) leaders.push_back(a);
    }
    const size_t nl = leaders.size();
    GP_CUDA(h, h->Kinv.reserve(mat * E * sizeof(double)));
    GP_CUDA(h, h->Wt.reserve(wt_doubles(ld) * E * sizeof(double)));
    GP_CUDA(h, h->beta.reserve((size_t)ld * E * sizeof(double)));
    GP_CUDA(h, h->chol.reserve(mat * (nl ? nl : 1) * sizeof(double)));
    GP_CUDA(h, h->zt.reserve(mat * (nl ? nl : 1) * sizeof(double)));
    GP_CUDA(h, h->linv.reserve(((size_t)NB * NB + 8) * kMaxE * sizeof(double)));     // per leader: Lkk^-1 | log det
    GP_CUDA(h, h->info.reserve(kMaxE * sizeof(int)));
    GP_CUDA(h, cudaMemsetAsync(h->info.p, 0, kMaxE * sizeof(int), h->stream));

    cudaStream_t main_stream = h->stream;
    while (h->aux_streams.size() + 1 < nl) {
        cudaStream_t st; cudaEvent_t ev;
        GP_CUDA(h, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        GP_CUDA(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        h->aux_streams.push_back(st); h->aux_events.push_back(ev);
    }
    if (!h->ev_fork) GP_CUDA(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    if (nl > 1) GP_CUDA(h, cudaEventRecord(h->ev_fork, main_stream));

    int rc = GPMPC_OK;
    for (size_t s = 0; s < nl && rc == GPMPC_OK; ++s) {
        const int a = leaders[s];
        cudaStream_t st = s == 0 ? main_stream : h->aux_streams[s - 1];
        if (s > 0) { cudaError_t e = cudaStreamWaitEvent(st, h->ev_fork, 0); if (e != cudaSuccess) { rc = fail(h, GPMPC_ERR_CUDA, cudaGetErrorString(e)); break; } }
        h->stream = st;                          // the helpers below launch on h->stream
        double *L = h->chol.as<double>() + s * mat;
        double *ZT = h->zt.as<double>() + s * mat;
        double *linv = h->linv.as<double>() + s * ((size_t)NB * NB + 8);
        int *info = h->info.as<int>() + s;
        double *Kinv = h->Kinv.as<double>() + a * mat;
        dim3 blk(32, 8), grid((np + 31) / 32, (np + 7) / 8);
        gram_kernel<<<grid, blk, 0, st>>>(h->X.as<double>(), n, np, h->D, make_hyper(h, a, false), L, ld, 1);
        h->launches++;
        if (h->la.size() <= s) h->la.resize(s + 1);
        if (!h->la[s].side) GP_CUDA(h, cudaStreamCreateWithFlags(&h->la[s].side, cudaStreamNonBlocking));
        static const bool no_lookahead = getenv("GPMPC_NO_LOOKAHEAD") != nullptr;
        if ((rc = cholesky_inplace(h, L, ZT, np, linv, info, no_lookahead ? nullptr : &h->la[s]))) break;
        logdet_kernel<<<1, 1024, 0, st>>>(L, ld, n, linv + (size_t)NB * NB);
        h->launches++;
        if ((rc = invert_factor(h, L, ZT, Kinv, np))) break;        // Kinv doubles as workspace until the next line
        // Ky^-1[i][j] = sum_{k >= max(i,j)} ZT[i][k] ZT[j][k]; lower tiles, then mirrored
        if ((rc = dgemm_nt(h, np, np, np, 1.0, ZT, ld, ZT, ld, 0.0, Kinv, ld, true, 2))) break;
        dim3 mblk(32, 8), mgrid(np / 32, np / 32);
        mirror_lower_kernel<<<mgrid, mblk, 0, st>>>(Kinv, np, ld);
        h->launches++;
        // beta = Ky^-1 y   (src/tools/uncertainty_prop.py:327)
        gemv_kernel<<<(np + 7) / 8, 256, 0, st>>>(Kinv, ld, n, h->Y.as<double>() + (size_t)a * ld,
                                                  h->beta.as<double>() + (size_t)a * ld, np);
        h->launches++;
        if ((rc = derive_weights(h, a))) break;
        if (s > 0) {
            cudaError_t e = cudaEventRecord(h->aux_events[s - 1], st);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(main_stream, h->aux_events[s - 1], 0);
            if (e != cudaSuccess) rc = fail(h, GPMPC_ERR_CUDA, cudaGetErrorString(e));
        }
    }
    h->stream = main_stream;
    if (rc) { cudaDeviceSynchronize(); return rc; }
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(h, GPMPC_ERR_CUDA, std::string("fit kernels: ") + cudaGetErrorString(e));
    }
    // one host synchronisation for all leaders: pivots and log-determinants
    std::vector<int> info(nl ? nl : 1, 0);
    std::vector<double> ldet(nl ? nl : 1, 0.0);
    if (nl) GP_CUDA(h, cudaMemcpyAsync(info.data(), h->info.p, nl * sizeof(int), cudaMemcpyDeviceToHost, main_stream));
    for (size_t s = 0; s < nl; ++s)
        GP_CUDA(h, cudaMemcpyAsync(&ldet[s], h->linv.as<double>() + s * ((size_t)NB * NB + 8) + (size_t)NB * NB, sizeof(double),
                                   cudaMemcpyDeviceToHost, main_stream));
    GP_CUDA(h, cudaStreamSynchronize(main_stream));
    for (size_t s = 0; s < nl; ++s) {
        if (info[s] != 0) {
            char msg[160];
            snprintf(msg, sizeof msg, "gpmpc_fit: Ky of output %d is not positive definite (pivot %d)", leaders[s], info[s] - 1);
            return fail(h, GPMPC_ERR_NOT_PD, msg);
        }
        h->logdet[leaders[s]] = ldet[s];
    }
    for (int a = 0; a < E; ++a) {
        if (!which[a] || twin[a] < 0) continue;
        GP_CUDA(h, cudaMemcpyAsync(h->Kinv.as<double>() + a * mat, h->Kinv.as<double>() + twin[a] * mat, mat * sizeof(double),
                                   cudaMemcpyDeviceToDevice, main_stream));
        h->logdet[a] = h->logdet[twin[a]];
        gemv_kernel<<<(np + 7) / 8, 256, 0, main_stream>>>(h->Kinv.as<double>() + a * mat, ld, n,
                                                           h->Y.as<double>() + (size_t)a * ld,
                                                           h->beta.as<double>() + (size_t)a * ld, np);
        GP_LAUNCH_CHECK(h);
        if ((rc = derive_weights(h, a))) return rc;
    }
    h->fitted = true;
    return GPMPC_OK;
}

}  // namespace gpmpc
