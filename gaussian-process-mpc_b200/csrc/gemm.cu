// fp64 tensor-core GEMM (DMMA mma.sync.m8n8k4) used by the fit: Cholesky trailing updates, the
// triangular-inverse panels, Ky^-1 = L^-T L^-1 and the posterior products K* Ky^-1 (src/gpr.py:171,306,325).
//
//   C[M,N] = alpha * A[M,K] * B[N,K]^T + beta * C        ("NT": both operands have K contiguous)
//   C[M,N] = alpha * A[M,K] * B[K,N]   + beta * C        ("NN")
//
// 64x64 block tile, 4 warps (2x2), each warp 32x32 = 4x4 m8n8 accumulator tiles, K staged 16 at a time
// through a cp.async double buffer.  The smem row pitch is 20 doubles so that the 32 lanes of a fragment
// load (8 rows x 4 k) touch 32 distinct banks pairs.
#include "common.cuh"

namespace gpmpc {

constexpr int GB = 64;       // block tile edge
constexpr int GK = 16;       // k-step
constexpr int GP = 20;       // smem pitch in doubles

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Generalised form used by the fit (all of it on the FP64 tensor pipe):
//   * batched over blockIdx.z with element strides sA, sB, sC (independent diagonal blocks of the recursive
//     triangular inverse),
//   * B either [N][K] ("NT", BNN = false) or [K][N] ("NN", BNN = true), both row-major,
//   * a per-tile k range that skips the structural zeros of triangular operands:
//       kmode 0: [0, K)            1: k >= row0 (A upper triangular)      2: k >= max(row0, col0)
//             3: k <  col0 + 64 (B = [K][N] upper triangular: B[k][n] = 0 for k > n)
//             4: row0 <= k < col0 + 64 (A upper triangular and B as in 3)
template <bool BNN>
__global__ void __launch_bounds__(128)
dgemm_kernel(int M, int N, int K, double alpha, const double *__restrict__ A, int lda, long long sA,
             const double *__restrict__ B, int ldb, long long sB, double beta, double *__restrict__ C, int ldc,
             long long sC, int tri_lower, int kmode)
{
    const int row0 = blockIdx.y * GB, col0 = blockIdx.x * GB;
    if (tri_lower && col0 > row0) return;
    A += (size_t)blockIdx.z * sA; B += (size_t)blockIdx.z * sB; C += (size_t)blockIdx.z * sC;
    __shared__ __align__(16) double As[2][GB * GP];
    __shared__ __align__(16) double Bs[2][BNN ? GK * (GB + 4) : GB * GP];     // NN: [k][n], pitch 68 (conflict-free fragment loads)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int g = lane >> 2, t = lane & 3;

    int k_lo = 0, k_hi = K;
    if (kmode == 1 || kmode == 4) k_lo = row0;
    else if (kmode == 2) k_lo = row0 > col0 ? row0 : col0;
    if (kmode == 3 || kmode == 4) k_hi = col0 + GB < K ? col0 + GB : K;
    k_lo = (k_lo / GK) * GK;
    if (k_lo > k_hi) k_lo = k_hi;
    const int nk = (k_hi - k_lo) / GK;

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int s, int kt) {
        const int kbase = k_lo + kt * GK;
        // 64 rows x 16 doubles = 512 chunks of 16 B per operand; 4 per thread
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int chunk = tid + c * 128;
            const int r = chunk >> 3, kc = (chunk & 7) * 2;
            cp_async16(&As[s][r * GP + kc], A + (size_t)(row0 + r) * lda + kbase + kc);
            if (!BNN) {
                cp_async16(&Bs[s][r * GP + kc], B + (size_t)(col0 + r) * ldb + kbase + kc);
            } else {                                     // 16 k-rows x 64 doubles
                const int kr = chunk >> 5, nc = (chunk & 31) * 2;
                cp_async16(&Bs[s][kr * (GB + 4) + nc], B + (size_t)(kbase + kr) * ldb + col0 + nc);
            }
        }
        cp_async_commit();
    };

    if (nk > 0) load_stage(0, 0);
    for (int kt = 0; kt < nk; ++kt) {
        const int s = kt & 1;
        if (kt + 1 < nk) { load_stage(s ^ 1, kt + 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; kk += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[s][(wm + i * 8 + g) * GP + kk + t];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                b[j] = BNN ? Bs[s][(kk + t) * (GB + 4) + wn + j * 8 + g] : Bs[s][(wn + j * 8 + g) * GP + kk + t];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = row0 + wm + i * 8 + g;
            const int c = col0 + wn + j * 8 + t * 2;
            double2 *p = reinterpret_cast<double2 *>(C + (size_t)r * ldc + c);
            double2 v;
            if (beta != 0.0) {
                v = *p;
                v.x = alpha * acc[i][j][0] + beta * v.x;
                v.y = alpha * acc[i][j][1] + beta * v.y;
            } else {
                v.x = alpha * acc[i][j][0];
                v.y = alpha * acc[i][j][1];
            }
            *p = v;
        }
}

int dgemm_batched(gpmpc_ctx *h, bool b_nn, int batch, int M, int N, int K, double alpha, const double *A, int lda,
                  long long sA, const double *B, int ldb, long long sB, double beta, double *C, int ldc, long long sC,
                  bool tri_lower, int kmode)
{
    if (M <= 0 || N <= 0 || batch <= 0) return GPMPC_OK;
    if (M % GB || N % GB || K % GK || (lda & 1) || (ldb & 1) || (ldc & 1) || batch > 65535)
        return fail(h, GPMPC_ERR_INVALID, "dgemm: shape not tile aligned");
    dim3 grid(N / GB, M / GB, batch);
    if (b_nn)
        dgemm_kernel<true><<<grid, 128, 0, h->stream>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC,
                                                        tri_lower ? 1 : 0, kmode);
    else
        dgemm_kernel<false><<<grid, 128, 0, h->stream>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC,
                                                         tri_lower ? 1 : 0, kmode);
    GP_LAUNCH_CHECK(h);
    return GPMPC_OK;
}

int dgemm_nt(gpmpc_ctx *h, int M, int N, int K, double alpha, const double *A, int lda, const double *B,
             int ldb, double beta, double *C, int ldc, bool tri_lower, int kmode)
{
    return dgemm_batched(h, false, 1, M, N, K, alpha, A, lda, 0, B, ldb, 0, beta, C, ldc, 0, tri_lower, kmode);
}

}  // namespace gpmpc
