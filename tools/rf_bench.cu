// Microbenchmark: FP64 FMA issue rate as a function of distinct register operands (operand-reuse / RF banking).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, const double *in, int iters, double s)
{
    double a[8], x[8], y[8];
    for (int i = 0; i < 8; ++i) { a[i] = s + in[i]; x[i] = in[8 + i + threadIdx.x]; y[i] = in[600 + i + threadIdx.x]; }
    const double w = in[1000 + threadIdx.x], c = in[2000];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fma(a[i], w, c);            // 1 distinct new operand (w, c reusable)
            if (MODE == 1) a[i] = fma(x[i], w, a[i]);         // 2 distinct + shared w
            if (MODE == 2) a[i] = fma(x[i], y[i], a[i]);      // 3 distinct
            if (MODE == 3) a[i] = fma(x[i], y[(i + 1) & 7], a[i]);   // 3 distinct, different pairing
        }
    }
    double r = 0; for (int i = 0; i < 8; ++i) r += a[i] + x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> void run(const char *name)
{
    double *d; cudaMalloc(&d, 148 * 8 * 256 * 8); double *in; cudaMalloc(&in, 4096 * 8); cudaMemset(in, 0, 4096 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 14; float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(d, in, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    printf("%-28s %.3f ms  %.2f TFLOP/s\n", name, best, 2.0 * 8 * iters * 148.0 * 8 * 256 / best / 1e9);
    cudaFree(d);
}
int main() { run<0>("fma(a,w,c) 1 new operand"); run<1>("fma(x,w,a) 2 new + shared"); run<2>("fma(x,y,a) 3 distinct"); run<3>("fma(x,y',a) 3 distinct b"); return 0; }
