// Do the FP64 vector pipe (DFMA) and the FP64 tensor pipe (DMMA m8n8k4) of sm_100a run concurrently?
// Three kernels with the same loop count: DFMA only, DMMA only, both interleaved.  If t(both) ~ max(t1, t2) the pipes
// are independent and the accumulation part of the pair kernel could be moved to DMMA; if t(both) ~ t1 + t2 they
// share the FP64 units.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/pipe_bench.cu -o tools/bin/pipe_bench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int MODE>   // 1: DFMA, 2: DMMA, 3: both
__global__ void __launch_bounds__(128, 2) k(double *out, int iters, double x, double y)
{
    double f[16], c[8][2];
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    const double a = x + threadIdx.x * 1e-9, b = y;
    for (int it = 0; it < iters; ++it) {
        if (MODE & 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fma(f[i], a, b);          // 16 independent DFMA
        }
        if (MODE & 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma(c[i][0], c[i][1], a, b);     // 8 independent DMMA (= 8 x 256 MAC per warp)
        }
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += f[i];
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> float run(int iters)
{
    double *d; cudaMalloc(&d, 296 * 128 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<MODE><<<296, 128>>>(d, iters, 0.999999, 1e-7); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaFree(d);
    return best;
}
int main()
{
    const int iters = 1 << 16;
    const float t1 = run<1>(iters), t2 = run<2>(iters), t3 = run<3>(iters);
    const double dfma = 296.0 * 128 * 16 * iters * 2 / (t1 * 1e-3) / 1e12;
    const double dm = 296.0 * 4 * 8 * 256.0 * iters * 2 / (t2 * 1e-3) / 1e12;
    printf("DFMA only %.2f ms (%.1f TFLOP/s), DMMA only %.2f ms (%.1f TFLOP/s), both %.2f ms (sum %.2f, max %.2f)\n", t1, dfma, t2, dm,
           t3, t1 + t2, t1 > t2 ? t1 : t2);
    return 0;
}
