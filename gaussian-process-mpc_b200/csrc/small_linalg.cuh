// Small dense linear algebra on E x E matrices held in registers (shared by rollout.cu and fullcov.cu).
#pragma once
#include "common.cuh"

namespace gpmpc {

// ---------------------------------------------------------------------------------------------
// Small dense LU (partial pivoting) in local memory: determinant and inverse of an E x E matrix.
// Same algorithm as the LAPACK getrf/getri behind torch.linalg.det / inv (src/mpc.py:179-185).
// ---------------------------------------------------------------------------------------------
// E is a compile-time constant so that the matrix lives in registers (a run-time E puts it in local memory and
// every access on the dependency chain pays an L1 round trip); row swaps use predicated moves, no dynamic index.
template <int E>
__device__ __forceinline__ double lu_det_inv_t(const double *Ain, double *inv)
{
    double A[E][E];
    int piv[E];
    double rinv[E];                                // reciprocals of the pivots (as in LAPACK's trtri), reused by the back substitution
    double det = 1.0;
#pragma unroll
    for (int r = 0; r < E; ++r)
#pragma unroll
        for (int k = 0; k < E; ++k) A[r][k] = Ain[r * E + k];
#pragma unroll
    for (int c = 0; c < E; ++c) {
        int p = c;
        double best = fabs(A[c][c]);
#pragma unroll
        for (int r = c + 1; r < E; ++r) { const double v = fabs(A[r][c]); if (v > best) { best = v; p = r; } }
        piv[c] = p;
#pragma unroll
        for (int r = c + 1; r < E; ++r)
            if (r == p) {
#pragma unroll
                for (int k = 0; k < E; ++k) { const double tmp = A[c][k]; A[c][k] = A[r][k]; A[r][k] = tmp; }
            }
        if (p != c) det = -det;
        det *= A[c][c];
        const double dinv = 1.0 / A[c][c];
        rinv[c] = dinv;
#pragma unroll
        for (int r = c + 1; r < E; ++r) {
            const double f = A[r][c] * dinv;
            A[r][c] = f;
#pragma unroll
            for (int k = c + 1; k < E; ++k) A[r][k] -= f * A[c][k];
        }
    }
    if (inv) {
#pragma unroll
        for (int col = 0; col < E; ++col) {
            double x[E];
#pragma unroll
            for (int r = 0; r < E; ++r) x[r] = (r == col) ? 1.0 : 0.0;
#pragma unroll
            for (int c = 0; c < E; ++c) {
#pragma unroll
                for (int r = c + 1; r < E; ++r)
                    if (r == piv[c]) { const double tmp = x[c]; x[c] = x[r]; x[r] = tmp; }
            }
#pragma unroll
            for (int r = 0; r < E; ++r)
#pragma unroll
                for (int k = 0; k < r; ++k) x[r] -= A[r][k] * x[k];
#pragma unroll
            for (int r = E - 1; r >= 0; --r) {
#pragma unroll
                for (int k = r + 1; k < E; ++k) x[r] -= A[r][k] * x[k];
                x[r] *= rinv[r];
            }
#pragma unroll
            for (int r = 0; r < E; ++r) inv[r * E + col] = x[r];
        }
    }
    return det;
}

}  // namespace gpmpc
