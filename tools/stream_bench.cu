// HBM streaming probe: how fast can 296 CTAs pull the upper-triangular 32x32 fp64 tiles of EG matrices through a
// TMA + mbarrier ring when the tiles are (a) boxes of a row-major ld x ld matrix (32 pieces of 256 B, 32 KB apart)
// or (b) contiguous 8 KB blocks (tile-major storage)?  Compute is a trivial sum so the memory system is the limit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/stream_bench.cu -o tools/bin/stream_bench -lcuda
#include "../gaussian-process-mpc_b200/csrc/mm_pairs.cuh"
#include <vector>
using namespace gpmpc;
struct alignas(64) PairTma { CUtensorMap map[4]; };
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, void *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
constexpr int EG = 4, STAGES = 3, THREADS = 128;
constexpr size_t STAGE = (size_t)EG * 1024;
__device__ __forceinline__ void mbar_arrive_l(void *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory"); }

template <int MODE>
__global__ void __launch_bounds__(THREADS, 2) stream_kernel(const double *W, int ld, int ntile, int total, double *out,
                                                            const __grid_constant__ PairTma tm)
{
    extern __shared__ __align__(128) double smem[];
    __shared__ __align__(8) unsigned long long full[STAGES], empty[STAGES];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, P = gridDim.x;
    if (tid == 0) { for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); } mbar_fence_init(); }
    __syncthreads();
    const int t_begin = (int)((long long)total * blockIdx.x / P), t_end = (int)((long long)total * (blockIdx.x + 1) / P);
    int Ii = 0, Ji = 0;
    { int rem = t_begin, row = 0; while (rem >= ntile - row) { rem -= ntile - row; ++row; } Ii = row; Ji = row + rem; }
    int issued = t_begin;
    const size_t mat = (size_t)ld * ld, tmat = (size_t)total * 1024;
    auto issue_next = [&]() {
        const int slot = (issued - t_begin) % STAGES;
        double *base = smem + (size_t)slot * STAGE;
        mbar_expect_tx(&full[slot], (unsigned)(STAGE * 8));
        for (int g = 0; g < EG; ++g) {
            if (MODE == 0) tma_load_2d(base + g * 1024, &tm.map[g], Ji * 32, Ii * 32, &full[slot]);
            else bulk_load_1d(base + g * 1024, W + g * tmat + (size_t)issued * 1024, 8192, &full[slot]);
        }
        ++issued; ++Ji; if (Ji == ntile) { ++Ii; Ji = Ii; }
    };
    if (tid == 0) for (int s = 0; s < STAGES - 1; ++s) if (issued < t_end) issue_next();
    double acc = 0.0;
    for (int t = t_begin; t < t_end; ++t) {
        const int it = t - t_begin, slot = it % STAGES;
        if (tid == 0 && issued < t_end) { if (it > 0) mbar_wait(&empty[(it - 1) % STAGES], ((it - 1) / STAGES) & 1); issue_next(); }
        mbar_wait(&full[slot], (it / STAGES) & 1);
        const double *Ws = smem + (size_t)slot * STAGE;
#pragma unroll
        for (int m = 0; m < 8; ++m)
#pragma unroll
            for (int g = 0; g < EG; ++g) acc += Ws[g * 1024 + (wid + 4 * m) * 32 + lane];
        __syncwarp();
        if (lane == 0) mbar_arrive_l(&empty[slot]);
    }
    out[blockIdx.x * THREADS + tid] = acc;
    (void)mat;
}

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 4096;
    const int ld = n, ntile = ld / 32, total = ntile * (ntile + 1) / 2;
    const size_t mat = (size_t)ld * ld;
    double *W, *out;
    cudaMalloc(&W, EG * mat * 8); cudaMemset(W, 0, EG * mat * 8);
    cudaMalloc(&out, 296 * THREADS * 8);
    PairTma tm;
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    for (int g = 0; g < EG; ++g) {
        const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)ld}, strides[1] = {(cuuint64_t)ld * 8};
        const cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
        ((encode_fn)fn)(&tm.map[g], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, W + g * mat, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    const size_t smem = STAGES * STAGE * 8;
    cudaFuncSetAttribute(stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = (double)EG * total * 8192;
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) stream_kernel<0><<<296, THREADS, smem>>>(W, ld, ntile, total, out, tm);
            else stream_kernel<1><<<296, THREADS, smem>>>(W, ld, ntile, total, out, tm);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%s: %.1f us  %.0f GB/s  (%s)\n", mode == 0 ? "2-D boxes of a row-major matrix" : "contiguous 8 KB tiles       ",
               best * 1e3, bytes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
