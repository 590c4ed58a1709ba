// Explicit instantiations of the full-covariance pair kernels (mm_full.cuh) for one input dimension D (compiled once
// per D = 2..8 with -DGPMPC_INST_D=<D>, like mm_pairs_inst.cu).
#include "mm_full.cuh"

#ifndef GPMPC_INST_D
#error "compile with -DGPMPC_INST_D=<2..8>"
#endif

namespace gpmpc {

template <int D, int NP, bool BWD>
static cudaError_t launch_full_one(const FullPairArgs &a, int ctas, cudaStream_t st)
{
    const size_t smem = full_smem_bytes<D, NP>();
    static bool configured[kMaxDevices] = {};
    if (first_use_on_device(configured)) {
        cudaError_t e = cudaFuncSetAttribute(mm_full_pairs<D, NP, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    mm_full_pairs<D, NP, BWD><<<ctas, FULL_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

template <int D, bool BWD>
static cudaError_t launch_full_np(int NP, const FullPairArgs &a, int ctas, cudaStream_t st)
{
    switch (NP) {
        case 1: return launch_full_one<D, 1, BWD>(a, ctas, st);
        case 2: return launch_full_one<D, 2, BWD>(a, ctas, st);
        case 3: return launch_full_one<D, 3, BWD>(a, ctas, st);
        case 4: return launch_full_one<D, 4, BWD>(a, ctas, st);
        case 6: return launch_full_one<D, 6, BWD>(a, ctas, st);
        case 10: return launch_full_one<D, 10, BWD>(a, ctas, st);
    }
    return cudaErrorInvalidValue;
}

#define GPMPC_CAT2(a, b) a##b
#define GPMPC_CAT(a, b) GPMPC_CAT2(a, b)
// NP must be one of 1, 2, 3, 4, 6, 10 (the caller splits a unit's pair-outputs into launches of these sizes)
cudaError_t GPMPC_CAT(launch_full_pairs_D, GPMPC_INST_D)(int NP, bool bwd, const FullPairArgs &a, int ctas, cudaStream_t st)
{
    return bwd ? launch_full_np<GPMPC_INST_D, true>(NP, a, ctas, st) : launch_full_np<GPMPC_INST_D, false>(NP, a, ctas, st);
}

cudaError_t GPMPC_CAT(launch_full_mean_D, GPMPC_INST_D)(bool bwd, const FullMeanArgs &a, cudaStream_t st)
{
    dim3 grid((a.B + FULL_MEAN_THREADS - 1) / FULL_MEAN_THREADS, FULL_MEAN_JP);
    if (bwd) mean_full_kernel<GPMPC_INST_D, true><<<grid, FULL_MEAN_THREADS, 0, st>>>(a);
    else mean_full_kernel<GPMPC_INST_D, false><<<grid, FULL_MEAN_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace gpmpc
