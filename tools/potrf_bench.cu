// Timing of the 64x64 diagonal-block kernel of the fit in isolation (development tool).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/potrf_bench.cu gaussian-process-mpc_b200/csrc/gemm.cu -o tools/bin/potrf_bench
#include "../gaussian-process-mpc_b200/csrc/fit.cu"
#include <vector>
using namespace gpmpc;
int main()
{
    const int ld = 256;
    std::vector<double> A((size_t)ld * ld, 0.0);
    for (int i = 0; i < ld; ++i) for (int j = 0; j < ld; ++j) A[(size_t)i * ld + j] = exp(-0.01 * (i - j) * (i - j)) + (i == j ? 0.5 : 0.0);
    double *dA, *dL, *dZ; int *info;
    cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dL, 64 * 64 * 8); cudaMalloc(&dZ, A.size() * 8); cudaMalloc(&info, 4);
    cudaMemset(info, 0, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 5; ++rep) {
        cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        for (int k = 0; k < 4; ++k) potrf_diag_kernel<<<1, 256>>>(dA, ld, 64 * k, dL, dZ, ld, info);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("4 diagonal blocks: %.1f us each (%s)\n", ms * 1e3 / 4, cudaGetErrorString(cudaGetLastError()));
    }
    // check block 0: L L^T = A and L Linv = I
    std::vector<double> L((size_t)ld * ld), Li(64 * 64);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    potrf_diag_kernel<<<1, 256>>>(dA, ld, 0, dL, dZ, ld, info);
    cudaMemcpy(L.data(), dA, A.size() * 8, cudaMemcpyDeviceToHost); cudaMemcpy(Li.data(), dL, 64 * 64 * 8, cudaMemcpyDeviceToHost);
    double e1m = 0, e2m = 0;
    for (int i = 0; i < 64; ++i) for (int j = 0; j <= i; ++j) {
        double s = 0, t = 0;
        for (int k = 0; k <= j; ++k) s += L[(size_t)i * ld + k] * L[(size_t)j * ld + k];
        for (int k = j; k <= i; ++k) t += L[(size_t)i * ld + k] * Li[k * 64 + j];
        e1m = fmax(e1m, fabs(s - A[(size_t)i * ld + j])); e2m = fmax(e2m, fabs(t - (i == j)));
    }
    printf("max |L L^T - A| = %.2e, max |L Linv - I| = %.2e\n", e1m, e2m);
    return 0;
}
