// Which pairs of 64-bit register operands conflict in a DFMA?  acc[i] = fma(x[(i+S)%N], w, acc[i]) for shifts S.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 16;
template <int S>
__global__ void __launch_bounds__(128, 2) k(double *out, const double *in, int iters)
{
    double acc[N], x[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { acc[i] = in[i]; x[i] = in[100 + i + threadIdx.x]; }
    const double w = in[300 + threadIdx.x], dx = in[400];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) acc[i] = fma(x[(i + S) % N], w, acc[i]);
        x[0] += dx;
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) r += acc[i] + x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int S> void run()
{
    double *d, *in; cudaMalloc(&d, 296 * 128 * 8); cudaMalloc(&in, 4096 * 8); cudaMemset(in, 0, 4096 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 15; float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(e0); k<S><<<296, 128>>>(d, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    printf("shift %d: %.3f ms  %.2f cycles per DFMA\n", S, best, best * 1e-3 * 1.965e9 / (iters * (N + 1.0) * 2.0));
}
int main() { run<0>(); run<1>(); run<2>(); run<3>(); run<4>(); run<5>(); run<6>(); run<7>(); run<8>(); return 0; }
