// Shared declarations for libgpmpc.so (sm_100a).  Internal header -- the public C ABI is include/gpmpc.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/gpmpc.h"

namespace gpmpc {

constexpr int kTile = 64;          // padding granule of n (fit GEMM tiles are 64x64)
constexpr int kPairTile = 32;      // pair-space tile rows of the moment-matching kernels
constexpr int kMaxD = GPMPC_MAX_D;
constexpr int kMaxE = GPMPC_MAX_E;
constexpr int kGroupMax = 4;       // outputs evaluated per pair-kernel pass (sharing one exp)
constexpr int kSplitMaxWorld = 8;  // GPUs one rollout can be split over
constexpr int kSplitNV = 2 * kGroupMax * (1 + 2 * GPMPC_MAX_D);   // doubles one rank contributes per step (pair + mean sums)

// Wt storage: only the upper-triangular 32x32 tiles, each tile contiguous (8 KB, row-major inside), tiles in
// row-major order of (I, J >= I).  A contiguous range of the tile list is a contiguous range of memory, which is
// what the per-step kernels stream (bulk copies of whole tiles instead of 32 strided 256-byte rows).
__host__ __device__ inline size_t wt_tiles(int ld) { const size_t nt = (size_t)ld / kPairTile; return nt * (nt + 1) / 2; }
__host__ __device__ inline size_t wt_tile_index(int I, int J, int nt) { return (size_t)I * nt - (size_t)I * (I - 1) / 2 + (size_t)(J - I); }
__host__ __device__ inline size_t wt_doubles(int ld) { return wt_tiles(ld) * kPairTile * kPairTile; }

// number of accumulated statistics per (rollout, output) in q-space: T, N1[D], N2[D]
__host__ __device__ constexpr int nacc(int D) { return 1 + 2 * D; }
// tape entries per (rollout, step, output): mean, var, dm/du[D], dm/ds[D], dv/du[D], dv/ds[D]
__host__ __device__ constexpr int ntape(int D) { return 2 + 4 * D; }

struct DevBuf {                    // grow-only device buffer
    void *p = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct HostBuf {                   // grow-only pinned host buffer
    void *p = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaHostAlloc(&p, need, cudaHostAllocDefault);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct LookAhead {                 // side stream + events of one factorisation slot (fit.cu: cholesky_inplace)
    cudaStream_t side = nullptr;
    std::vector<cudaEvent_t> ev;
};

struct LambdaGroup {               // outputs whose propagation length-scales are bit-identical
    int count = 0;
    int outputs[kMaxE];
};

}  // namespace gpmpc

struct gpmpc_ctx {
    int device = 0, D = 0, E = 0, m = 0;
    int n = 0, ld = 0;             // training points and padded leading dimension (multiple of 64)
    cudaStream_t stream = nullptr;
    std::string err;
    long long launches = 0;

    // hyper-parameters (host copies).  *_fit were used for Ky^-1, *_prop are used by moment matching.
    double lam_fit[gpmpc::kMaxE][gpmpc::kMaxD], sf_fit[gpmpc::kMaxE], noise[gpmpc::kMaxE];
    double lam_prop[gpmpc::kMaxE][gpmpc::kMaxD], sf_prop[gpmpc::kMaxE];
    std::vector<gpmpc::LambdaGroup> groups;
    bool fitted = false;
    double logdet[gpmpc::kMaxE];   // log det Ky per output (2 sum log diag L), set by the fit

    // device state
    gpmpc::DevBuf X;               // [ld, D]  (rows >= n are zero)
    gpmpc::DevBuf Y;               // [E, ld]
    gpmpc::DevBuf Kinv;            // [E, ld, ld]
    gpmpc::DevBuf Wt;              // [E][upper tiles][32*32]  moment-matching weights, tile-major (fit.cu: derive_weights)
    gpmpc::DevBuf beta;            // [E, ld]
    gpmpc::DevBuf chol, zt, tt;    // fit workspaces: L [ld,ld], L^-T [ld,ld], panel [ld,64]
    gpmpc::DevBuf linv;            // [64,64] inverse of the current diagonal block
    gpmpc::DevBuf info;            // int: index of first bad pivot + 1, 0 = ok
    gpmpc::DevBuf hyp;             // device copy of propagation hypers: lam[E,D], sf[E]

    // rollout workspaces (grow-only)
    gpmpc::DevBuf mu, var, tape, cst, part, mpart, stage_in, stage_out, gbuf, tickets, dbg;
    gpmpc::DevBuf claim;              // mm_step_single: per-step slice-claim counters [H][kClaimStride] ints
    // per mille of the tiles that go to the first-arrived CTA of each SM (500 = equal static slices).  Measured on B200
    // (B = 1 objective+gradient, us): n=4096, H=30: 500 -> 2050, 570 -> 1995, 630 -> 1980, 660 -> 1953..1974, 700 -> 2021,
    // 740 -> 2088; n=16384, H=20: 500 -> 18919, 600 -> 18466, 660 -> 17653, 700 -> 17961, 740 -> 18519
    int opt_single_big = 660;
    gpmpc::HostBuf pin_in, pin_out;   // small evaluations: all host inputs / outputs travel as ONE pinned copy each way
    int tape_B = 0, tape_H = 0;    // shape of the tape held from the last rollout

    // full-covariance rollout (fullcov.cu): cross-output weights (built lazily, rebuilt when Wt changes), the plan
    // descriptors on the device and the workspaces of the forward / backward sweeps
    long long weights_epoch = 0;   // bumped by derive_weights (any change of Wt / beta / propagation lambdas)
    long long cross_epoch = -1;    // weights_epoch the cross weights were built for
    gpmpc::DevBuf Wx, fc_plan, fc_mu, fc_cov, fc_cst, fc_raw, fc_part, fc_red, fc_gbar, fc_seed, fc_carry, fc_io;
    int fc_B = 0, fc_H = 0;        // shape of the full-covariance tape held from the last gpmpc_rollout_full

    // L2 persistence for the few-rollouts kernels (rollout.cu): device limits, queried once
    bool opt_l2_persist = false;   // measured on B200: no effect (2.139 vs 2.146 ms at n=4096, 18.8 vs 18.4 ms at n=16384)
    long long l2_persist_max = -1, l2_window_max = 0;

    // gpmpc_set_option
    bool opt_persistent = false;   // single rollouts: whole horizon in one persistent cooperative launch (measured 8 % slower
                                   // than one launch per step on one GPU; always used when the rollout is split over GPUs)

    // a single rollout split over several GPUs (gpmpc_split_*): mailbox of per-step sums written by the peers' kernels
    int split_world = 1, split_rank = 0;
    long long split_seq = 0;       // step sequence number, advanced identically on every rank
    gpmpc::DevBuf split_buf;       // local mailbox: [8 ranks][2 slots][kSplitNV] doubles, then [8][2] 64-bit flags
    double *peer_mail[8] = {};     // every rank's mailbox as mapped into this process (own entry = split_buf)
    unsigned long long *peer_flags[8] = {};
    bool split_ipc[8] = {};        // entries opened with cudaIpcOpenMemHandle (to be closed)
    bool opt_split_timeline = false;   // stamp the exchange of every step (synchronises after the launch)
    double split_exchange_mean_us = 0.0, split_exchange_max_us = 0.0;

    // auxiliary streams / events: independent outputs are fitted concurrently (fit.cu)
    std::vector<cudaStream_t> aux_streams;
    std::vector<cudaEvent_t> aux_events;
    cudaEvent_t ev_fork = nullptr;
    std::vector<gpmpc::LookAhead> la;   // per concurrent factorisation: look-ahead side stream

    // timing of the last pair-kernel sequence
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double last_pair_ms = 0.0;
    long long last_pair_evals = 0;
    bool time_pairs = false;
};

namespace gpmpc {

inline int fail(gpmpc_ctx *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    return code;
}

#define GP_CUDA(h, call)                                                                       \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return gpmpc::fail((h), GPMPC_ERR_CUDA,                                            \
                               std::string(#call) + ": " + cudaGetErrorString(e_));            \
    } while (0)

#define GP_LAUNCH_CHECK(h)                                                                     \
    do {                                                                                       \
        (h)->launches++;                                                                       \
        cudaError_t e_ = cudaGetLastError();                                                   \
        if (e_ != cudaSuccess)                                                                 \
            return gpmpc::fail((h), GPMPC_ERR_CUDA,                                            \
                               std::string("kernel launch: ") + cudaGetErrorString(e_));       \
    } while (0)

inline bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// copy `bytes` from a host-or-device source into a device destination on the handle's stream
inline cudaError_t to_device(gpmpc_ctx *h, void *dst, const void *src, size_t bytes) {
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, h->stream);
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// cudaFuncSetAttribute is per device: `done` is a per-call-site table indexed by the current device
constexpr int kClaimStride = 260;   // 256 per-SM arrival counters + big / small slice counters
constexpr int kMaxDevices = 64;
inline bool first_use_on_device(bool (&done)[kMaxDevices]) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
}

// ---- fit.cu -------------------------------------------------------------------------------
int fit_all(gpmpc_ctx *h, const bool *which);           // (re)fit the outputs flagged in which[E]
int derive_weights(gpmpc_ctx *h, int a);                // Wt[a] from Kinv[a], beta[a], lam_prop[a]
int gram_into(gpmpc_ctx *h, int a, double *dst, int ldd, bool add_noise);   // Kf / Ky of output a
void rebuild_groups(gpmpc_ctx *h);
int upload_prop_hypers(gpmpc_ctx *h);

// ---- gemm.cu ------------------------------------------------------------------------------
// C[M,N] = alpha * A[M,K] * B[N,K]^T + beta * C   (all row-major, fp64, DMMA m8n8k4)
//   M, N multiples of 64; K multiple of 16.
//   tri_lower: compute only tiles whose first column <= first row (+ col_off/row_off aligned at 0)
//   kmode: 0 = k in [0,K); 1 = k >= tile row0 (A upper triangular); 2 = k >= max(row0, col0)
int dgemm_nt(gpmpc_ctx *h, int M, int N, int K, double alpha, const double *A, int lda, const double *B,
             int ldb, double beta, double *C, int ldc, bool tri_lower, int kmode);
// batched over `batch` problems (element strides sA/sB/sC); b_nn: B is [K][N]; kmode 3/4: see gemm.cu
int dgemm_batched(gpmpc_ctx *h, bool b_nn, int batch, int M, int N, int K, double alpha, const double *A, int lda,
                  long long sA, const double *B, int ldb, long long sB, double beta, double *C, int ldc, long long sC,
                  bool tri_lower, int kmode);

// ---- predict.cu ---------------------------------------------------------------------------
int kernel_matrix_dev(gpmpc_ctx *h, int a, int p, const double *Xs_dev, double *out_dev, int ldo);

// ---- rollout.cu / mm_pairs.cu -------------------------------------------------------------
struct StepIO;   // defined in rollout.cu

}  // namespace gpmpc
