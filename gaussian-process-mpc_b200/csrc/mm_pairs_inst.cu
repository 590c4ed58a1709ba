// Explicit instantiations of the batched pair-sum kernel and of the fused few-rollouts step kernel for one input dimension D (compiled once per
// D = 2..8 with -DGPMPC_INST_D=<D> so that the seven translation units build in parallel).
#include "mm_pairs.cuh"
#include "mm_step_single.cuh"

#ifndef GPMPC_INST_D
#error "compile with -DGPMPC_INST_D=<2..8>"
#endif

namespace gpmpc {

template <int D, int EG, bool GRAD>
static cudaError_t launch_one(const PairArgs &a, dim3 grid, cudaStream_t st)
{
    const size_t smem = pair_smem_bytes<D, EG>();
    static bool configured[kMaxDevices] = {};
    static int resident[kMaxDevices] = {};       // CTAs of this variant that fit on the whole device (one wave)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) dev = 0;
    if (first_use_on_device(configured)) {
        cudaError_t e = cudaFuncSetAttribute(mm_pairs_batch<D, EG, GRAD>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int occ = 0, sms = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mm_pairs_batch<D, EG, GRAD>, PAIR_THREADS, smem);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        resident[dev] = occ * sms;
    }
    // Work items are drawn from ticket counters, so any number of CTAs produces the same (bit-identical) result: the
    // variants with few outputs or without the gradient moments need fewer registers and run more CTAs per SM than
    // the 2 the caller assumed.
    if (resident[dev] > 0) {
        const int per_chunk = resident[dev] / a.chunks;
        if (per_chunk * a.chunks > (int)grid.x) grid.x = per_chunk * a.chunks;
    }
    mm_pairs_batch<D, EG, GRAD><<<grid, PAIR_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

template <int D>
static cudaError_t launch_d(int EG, bool grad, const PairArgs &a, dim3 grid, cudaStream_t st)
{
    switch (EG * 2 + (grad ? 1 : 0)) {
        case 2: return launch_one<D, 1, false>(a, grid, st);
        case 3: return launch_one<D, 1, true>(a, grid, st);
        case 4: return launch_one<D, 2, false>(a, grid, st);
        case 5: return launch_one<D, 2, true>(a, grid, st);
        case 6: return launch_one<D, 3, false>(a, grid, st);
        case 7: return launch_one<D, 3, true>(a, grid, st);
        case 8: return launch_one<D, 4, false>(a, grid, st);
        case 9: return launch_one<D, 4, true>(a, grid, st);
    }
    return cudaErrorInvalidValue;
}

template <int D, int EG, bool GRAD>
static cudaError_t launch_single_one(const SingleStepArgs &a, dim3 grid, cudaStream_t st)
{
    const size_t smem = single_smem_bytes<D, EG>();
    static bool configured[kMaxDevices] = {};
    if (first_use_on_device(configured)) {
        cudaError_t e = cudaFuncSetAttribute(mm_step_single<D, EG, GRAD>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    // programmatic dependent launch: consecutive step kernels overlap this grid's prologue (barrier setup, first
    // TMA tile loads) with the previous grid's tail; the kernel orders itself with griddepcontrol.wait
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(SINGLE_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, mm_step_single<D, EG, GRAD>, a);
}

template <int D>
static cudaError_t launch_single_d(int EG, bool grad, const SingleStepArgs &a, dim3 grid, cudaStream_t st)
{
    switch (EG * 2 + (grad ? 1 : 0)) {
        case 2: return launch_single_one<D, 1, false>(a, grid, st);
        case 3: return launch_single_one<D, 1, true>(a, grid, st);
        case 4: return launch_single_one<D, 2, false>(a, grid, st);
        case 5: return launch_single_one<D, 2, true>(a, grid, st);
        case 6: return launch_single_one<D, 3, false>(a, grid, st);
        case 7: return launch_single_one<D, 3, true>(a, grid, st);
        case 8: return launch_single_one<D, 4, false>(a, grid, st);
        case 9: return launch_single_one<D, 4, true>(a, grid, st);
    }
    return cudaErrorInvalidValue;
}

#define GPMPC_CAT2(a, b) a##b
#define GPMPC_CAT(a, b) GPMPC_CAT2(a, b)
cudaError_t GPMPC_CAT(launch_pairs_batch_D, GPMPC_INST_D)(int EG, bool grad, const PairArgs &a,
                                                          dim3 grid, cudaStream_t st)
{
    return launch_d<GPMPC_INST_D>(EG, grad, a, grid, st);
}
cudaError_t GPMPC_CAT(launch_step_single_D, GPMPC_INST_D)(int EG, bool grad, const SingleStepArgs &a,
                                                          dim3 grid, cudaStream_t st)
{
    return launch_single_d<GPMPC_INST_D>(EG, grad, a, grid, st);
}

}  // namespace gpmpc
