// Stand-alone timing harness for the pair-sum kernel variants (development tool, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo [-DGPMPC_...] tools/pair_bench.cu -o pb
#include "../gaussian-process-mpc_b200/csrc/mm_pairs.cuh"
#include <vector>
#include <random>
using namespace gpmpc;
#ifndef GPMPC_BENCH_GRAD
#define GPMPC_BENCH_GRAD 1       // moment selection of mm_pairs.cuh: 0 forward, 1 all steps, 2 first step
#endif
#ifndef GPMPC_BENCH_NS
#define GPMPC_BENCH_NS 4         // state dimensions (N2 only for k < NS); 5 = all moments (the round-1 kernel)
#endif

__global__ void dfma_latency_kernel(double *out, int iters, double m, double c)
{
    double a = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { a = fma(a, m, c); a = fma(a, m, c); a = fma(a, m, c); a = fma(a, m, c); }
    long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) out[64] = (double)(t1 - t0) / (4.0 * iters);
}

__global__ void exp_check_kernel(const double *S, double *out, int n)
{
    __shared__ double tab[16];
    if (threadIdx.x < 16) tab[threadIdx.x] = kExp2Tab[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out[i] = exp_neg(S[i], tab); out[n + i] = exp(-fmin(S[i], 700.0)); }
}

int main(int argc, char **argv)
{
    {   // accuracy of exp_neg against libm exp over [0, 720]
        const int N = 1 << 20;
        std::vector<double> hs(N), ho(2 * N);
        std::mt19937_64 r2(7); std::uniform_real_distribution<double> V(0, 1);
        for (int i = 0; i < N; ++i) { double v = V(r2); hs[i] = (i & 1) ? 720.0 * v : 30.0 * v * v; }
        hs[0] = 1e300; hs[1] = INFINITY; hs[2] = 700.0; hs[3] = 0.0; hs[4] = 1e8;
        double *ds, *dout; cudaMalloc(&ds, N * 8); cudaMalloc(&dout, 2 * N * 8);
        cudaMemcpy(ds, hs.data(), N * 8, cudaMemcpyHostToDevice);
        exp_check_kernel<<<(N + 255) / 256, 256>>>(ds, dout, N);
        cudaMemcpy(ho.data(), dout, 2 * N * 8, cudaMemcpyDeviceToHost);
        double worst = 0; for (int i = 0; i < N; ++i) { double e = fabs(ho[i] - ho[N + i]) / ho[N + i]; if (e > worst) worst = e; }
        printf("exp_neg variant %d: max rel err vs libm = %.3e\n", GPMPC_EXP_VARIANT, worst);
        cudaFree(ds); cudaFree(dout);
    }
    constexpr int D = 5, EG = 4;
    const int n = argc > 1 ? atoi(argv[1]) : 4096;
    const int B = argc > 2 ? atoi(argv[2]) : 1024;
    const int ld = (n + 63) / 64 * 64;
    const int Bpad = (B + 31) / 32 * 32;
    std::mt19937_64 rng(1);
    std::uniform_real_distribution<double> U(-1, 1);
    std::vector<double> hX((size_t)ld * D, 0.0), hW(wt_doubles(ld)), hc((size_t)4 * D * Bpad);
    for (int i = 0; i < n * D; ++i) hX[i] = U(rng);
    for (int i = 0; i < ld; ++i) for (int j = (i / 32) * 32; j < ld; ++j)      // tile-major upper-triangular tiles
        hW[wt_tile_index(i / 32, j / 32, ld / 32) * 1024 + (i % 32) * 32 + (j % 32)] = (i < n && j < n && j >= i) ? U(rng) : 0.0;
    for (int k = 0; k < D; ++k) for (int b = 0; b < Bpad; ++b) { double c = 0.3 + 0.05 * U(rng), u = 0.5 * U(rng); hc[(size_t)k * Bpad + b] = c; hc[(size_t)(D + k) * Bpad + b] = c * u; }
    double *dX, *dW[EG], *dc, *dpart, *dlat;
    cudaMalloc(&dX, hX.size() * 8); cudaMemcpy(dX, hX.data(), hX.size() * 8, cudaMemcpyHostToDevice);
    for (int g = 0; g < EG; ++g) { cudaMalloc(&dW[g], hW.size() * 8); cudaMemcpy(dW[g], hW.data(), hW.size() * 8, cudaMemcpyHostToDevice); }
    cudaMalloc(&dc, hc.size() * 8); cudaMemcpy(dc, hc.data(), hc.size() * 8, cudaMemcpyHostToDevice);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int chunks = (B + PAIR_THREADS - 1) / PAIR_THREADS;
    const long long nt = ld / PT, total = nt * (nt + 1) / 2;
#ifndef GPMPC_CTAS_PER_SM
#define GPMPC_CTAS_PER_SM 2
#endif
    int ctas = (GPMPC_CTAS_PER_SM * sms) / chunks; if (ctas < 1) ctas = 1;
#ifndef GPMPC_ITEMS_PER_CTA
#define GPMPC_ITEMS_PER_CTA 8
#endif
    int P = ctas * GPMPC_ITEMS_PER_CTA;
    if (P > total) P = (int)total;
    int *dcnt; cudaMalloc(&dcnt, chunks * sizeof(int));
    cudaMalloc(&dpart, (size_t)P * EG * (1 + 2 * D) * Bpad * 8);
    cudaMalloc(&dlat, 128 * 8);
    PairArgs a;
    for (int g = 0; g < EG; ++g) { a.Wt[g] = dW[g]; a.out_idx[g] = g; }
    a.X = dX; a.cst = dc; a.part = dpart; a.ld = ld; a.ntile = (int)nt; a.B = B; a.Bpad = Bpad; a.E = EG; a.n_items = P; a.counters = dcnt; a.total_tiles = (int)total; a.chunks = chunks;
#ifdef GPMPC_PAIR_TIMING
    unsigned long long *dtimes; cudaMalloc(&dtimes, (size_t)ctas * chunks * 3 * 8); a.cta_times = dtimes;
#endif
    const size_t smem = pair_smem_bytes<D, EG>();
    cudaFuncSetAttribute(mm_pairs_batch<D, EG, GPMPC_BENCH_GRAD, GPMPC_BENCH_NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(ctas * chunks);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaMemset(dcnt, 0, chunks * sizeof(int));
        cudaEventRecord(e0);
        mm_pairs_batch<D, EG, GPMPC_BENCH_GRAD, GPMPC_BENCH_NS><<<grid, PAIR_THREADS, smem>>>(a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    std::vector<double> hp((size_t)P * EG * (1 + 2 * D) * Bpad);
    cudaMemcpy(hp.data(), dpart, hp.size() * 8, cudaMemcpyDeviceToHost);
    double chk[3] = {0, 0, 0};
    for (int p = 0; p < P; ++p) for (int g = 0; g < EG; ++g) for (int b = 0; b < B; ++b) {
        const size_t base = (((size_t)p * EG + g) * (1 + 2 * D)) * Bpad + b;
        chk[0] += hp[base]; chk[1] += hp[base + (size_t)1 * Bpad]; chk[2] += hp[base + (size_t)(1 + D) * Bpad];
    }
#ifdef GPMPC_PAIR_TIMING
    {
        std::vector<unsigned long long> ht((size_t)ctas * chunks * 3);
        cudaMemcpy(ht.data(), dtimes, ht.size() * 8, cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull, t1 = 0; for (size_t i = 0; i < ht.size() / 3; ++i) { if (ht[3*i] < t0) t0 = ht[3*i]; if (ht[3*i+1] > t1) t1 = ht[3*i+1]; }
        double smin = 1e30, smax = 0, ssum = 0, dmin = 1e30, dmax = 0, dsum = 0; int cnt[256] = {0};
        for (size_t i = 0; i < ht.size() / 3; ++i) {
            double st = (ht[3*i] - t0) * 1e-6, du = (ht[3*i+1] - ht[3*i]) * 1e-6;
            smin = fmin(smin, st); smax = fmax(smax, st); ssum += st; dmin = fmin(dmin, du); dmax = fmax(dmax, du); dsum += du; cnt[ht[3*i+2] & 255]++;
        }
        int c1 = 0, c2 = 0, c3 = 0, c0 = 0; for (int i = 0; i < 148; ++i) { if (cnt[i] == 0) c0++; else if (cnt[i] == 1) c1++; else if (cnt[i] == 2) c2++; else c3++; }
        printf("CTA timing: span %.3f ms; start min/avg/max %.3f/%.3f/%.3f ms; duration min/avg/max %.3f/%.3f/%.3f ms; SMs with 0/1/2/3+ CTAs: %d/%d/%d/%d\n",
               (t1 - t0) * 1e-6, smin, ssum / (ht.size() / 3), smax, dmin, dsum / (ht.size() / 3), dmax, c0, c1, c2, c3);
    }
#endif
    const double pairs = (double)n * (n + 1) / 2 * B;
    dfma_latency_kernel<<<1, 32>>>(dlat, 4096, 0.999999, 1e-9);
    double lat[65]; cudaMemcpy(lat, dlat, 65 * 8, cudaMemcpyDeviceToHost);
    printf("%s: n=%d B=%d P=%d smem=%zu  best %.3f ms  %.2f Gpair/s  (%s)  chk T=%.15g N1=%.15g N2=%.15g  dfma_lat=%.2f cyc\n",
           argv[0], n, B, P, smem, best, pairs / best / 1e6, cudaGetErrorString(err), chk[0], chk[1], chk[2], lat[64]);
    return 0;
}
