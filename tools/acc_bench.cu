// Microbenchmark of the accumulation pattern of the pair kernel: 4 outputs x (1 + 5 + 5) accumulators,
// acc[g][f] += w[g] * F[f] per "pair", in different loop orders.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int EG = 4, D = 5;
template <int MODE>
__global__ void __launch_bounds__(128, 2) k(double *out, const double *in, int iters)
{
    double acc1[EG][D], acc2[EG][D], accT[EG], q[D], qq[D], w[EG];
    for (int g = 0; g < EG; ++g) { accT[g] = in[g]; w[g] = in[100 + g + threadIdx.x]; for (int kk = 0; kk < D; ++kk) { acc1[g][kk] = in[10 + g * D + kk]; acc2[g][kk] = in[40 + g * D + kk]; } }
    for (int kk = 0; kk < D; ++kk) { q[kk] = in[300 + kk + threadIdx.x]; qq[kk] = q[kk] * q[kk]; }
    const double dq = in[500], dw = in[501];
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {           // output-major (as shipped)
#pragma unroll
            for (int g = 0; g < EG; ++g) {
                accT[g] += w[g];
#pragma unroll
                for (int kk = 0; kk < D; ++kk) { acc1[g][kk] = fma(w[g], q[kk], acc1[g][kk]); acc2[g][kk] = fma(w[g], qq[kk], acc2[g][kk]); }
            }
        } else if (MODE == 1) {    // dimension-major
#pragma unroll
            for (int kk = 0; kk < D; ++kk) {
#pragma unroll
                for (int g = 0; g < EG; ++g) acc1[g][kk] = fma(w[g], q[kk], acc1[g][kk]);
#pragma unroll
                for (int g = 0; g < EG; ++g) acc2[g][kk] = fma(w[g], qq[kk], acc2[g][kk]);
            }
#pragma unroll
            for (int g = 0; g < EG; ++g) accT[g] += w[g];
        } else if (MODE == 2) {    // products first (2-operand), then adds
#pragma unroll
            for (int g = 0; g < EG; ++g) {
                accT[g] += w[g];
#pragma unroll
                for (int kk = 0; kk < D; ++kk) { const double t = w[g] * q[kk]; acc1[g][kk] += t; acc2[g][kk] = fma(t, q[kk], acc2[g][kk]); }
            }
        }
        // slow drift keeps everything live without adding much work (7 ops per 44)
#pragma unroll
        for (int kk = 0; kk < D; ++kk) q[kk] += dq;
        w[0] += dw; w[2] += dw;
    }
    double r = 0;
    for (int g = 0; g < EG; ++g) { r += accT[g]; for (int kk = 0; kk < D; ++kk) r += acc1[g][kk] + acc2[g][kk]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char *name)
{
    double *d, *in; cudaMalloc(&d, 296 * 128 * 8); cudaMalloc(&in, 4096 * 8); cudaMemset(in, 0, 4096 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 15; float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); k<MODE><<<296, 128>>>(d, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    const double instr = (MODE == 2 ? 64.0 : 44.0) + 7.0;   // fp64 instructions per iteration per thread
    // cycles per fp64 instruction per SMSP: 2 CTAs x 4 warps per SM = 2 warps per SMSP
    const double cyc = best * 1e-3 * 1.965e9 / (iters * instr * 2.0);
    printf("%-34s %.3f ms  %.2f cycles per FP64 instruction (2.00 = pipe peak)\n", name, best, cyc);
}
int main() { run<0>("output-major fma(w,q,acc)"); run<1>("dimension-major"); run<2>("mul + add + fma"); return 0; }
