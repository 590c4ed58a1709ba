// Explicit instantiations of the batched pair-sum kernel and of the fused few-rollouts step kernel for one input dimension D (compiled once per
// D = 2..8 with -DGPMPC_INST_D=<D> so that the seven translation units build in parallel).
#include "mm_pairs.cuh"
#include "mm_step_single.cuh"
#include "mm_rollout_single.cuh"
#include <type_traits>

#ifndef GPMPC_INST_D
#error "compile with -DGPMPC_INST_D=<2..8>"
#endif

namespace gpmpc {

template <int D, int EG, int GRAD, int NS>
static cudaError_t launch_one(const PairArgs &a, dim3 grid, cudaStream_t st)
{
    const size_t smem = pair_smem_bytes<D, EG>();
    static bool configured[kMaxDevices] = {};
    static int resident[kMaxDevices] = {};       // CTAs of this variant that fit on the whole device (one wave)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) dev = 0;
    if (first_use_on_device(configured)) {
        cudaError_t e = cudaFuncSetAttribute(mm_pairs_batch<D, EG, GRAD, NS>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int occ = 0, sms = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mm_pairs_batch<D, EG, GRAD, NS>, PAIR_THREADS, smem);
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
    mm_pairs_batch<D, EG, GRAD, NS><<<grid, PAIR_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

template <int D, int EG, int GRAD, int NS>
static cudaError_t launch_single_one(const SingleStepArgs &a, dim3 grid, cudaStream_t st)
{
    const size_t smem = single_smem_bytes<D, EG>();
    static bool configured[kMaxDevices] = {};
    if (first_use_on_device(configured)) {
        cudaError_t e = cudaFuncSetAttribute(mm_step_single<D, EG, GRAD, NS>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    // programmatic dependent launch: consecutive step kernels overlap this grid's prologue (barrier setup, first
    // TMA tile loads) with the previous grid's tail; the kernel orders itself with griddepcontrol.wait
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(SINGLE_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (a.l2_bytes > 0) {
        attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[1].val.accessPolicyWindow.base_ptr = const_cast<void *>(a.l2_base);
        attr[1].val.accessPolicyWindow.num_bytes = a.l2_bytes;
        attr[1].val.accessPolicyWindow.hitRatio = a.l2_hit;
        attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.numAttrs = 2;
    }
    return cudaLaunchKernelEx(&cfg, mm_step_single<D, EG, GRAD, NS>, a);
}

// Moment selection (mm_pairs.cuh): grad_mode 0 forward only, 1 all steps, 2 first step without d/dx0; ns = number of
// state dimensions.  The reduced variants are compiled for one or two action dimensions (ns = D-1, D-2), every
// other layout runs the variant that accumulates all moments (NS = D).
template <int D, int EG, bool SINGLE, class Args>
static cudaError_t launch_eg(int grad_mode, int ns, const Args &a, dim3 grid, cudaStream_t st)
{
    auto go = [&](auto grad_c, auto ns_c) -> cudaError_t {
        constexpr int G = decltype(grad_c)::value, N = decltype(ns_c)::value;
        if constexpr (SINGLE) return launch_single_one<D, EG, G, N>(a, grid, st);
        else return launch_one<D, EG, G, N>(a, grid, st);
    };
    using std::integral_constant;
    if (grad_mode == 0) return go(integral_constant<int, 0>{}, integral_constant<int, D>{});
    if constexpr (D >= 2) {
        if (ns == D - 1)
            return grad_mode == 2 ? go(integral_constant<int, 2>{}, integral_constant<int, D - 1>{})
                                  : go(integral_constant<int, 1>{}, integral_constant<int, D - 1>{});
    }
    if constexpr (D >= 3) {
        if (ns == D - 2)
            return grad_mode == 2 ? go(integral_constant<int, 2>{}, integral_constant<int, D - 2>{})
                                  : go(integral_constant<int, 1>{}, integral_constant<int, D - 2>{});
    }
    return go(integral_constant<int, 1>{}, integral_constant<int, D>{});
}

template <int D, bool SINGLE, class Args>
static cudaError_t launch_any(int EG, int grad_mode, int ns, const Args &a, dim3 grid, cudaStream_t st)
{
    switch (EG) {
        case 1: return launch_eg<D, 1, SINGLE>(grad_mode, ns, a, grid, st);
        case 2: return launch_eg<D, 2, SINGLE>(grad_mode, ns, a, grid, st);
        case 3: return launch_eg<D, 3, SINGLE>(grad_mode, ns, a, grid, st);
        case 4: return launch_eg<D, 4, SINGLE>(grad_mode, ns, a, grid, st);
    }
    return cudaErrorInvalidValue;
}

// Persistent whole-horizon kernel (mm_rollout_single.cuh): cooperative launch; returns cudaErrorCooperativeLaunchTooLarge
// when the grid cannot be co-resident (the caller then falls back to one launch per step).
template <int D, int EG, int NS>
static cudaError_t launch_rollout_single_one(const RolloutSingleArgs &a, dim3 grid, cudaStream_t st)
{
    const size_t smem = single_smem_bytes<D, EG>();
    static bool configured[kMaxDevices] = {};
    static int capacity[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDevices) dev = 0;
    if (first_use_on_device(configured)) {
        cudaError_t e = cudaFuncSetAttribute(mm_rollout_single<D, EG, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int occ = 0, sms = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, mm_rollout_single<D, EG, NS>, SINGLE_THREADS, smem);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        capacity[dev] = occ * sms;
    }
    if ((long long)grid.x * grid.y > capacity[dev]) return cudaErrorCooperativeLaunchTooLarge;
    RolloutSingleArgs args = a;
    void *params[] = {&args};
    return cudaLaunchCooperativeKernel((const void *)mm_rollout_single<D, EG, NS>, grid, dim3(SINGLE_THREADS), params, smem, st);
}

template <int D, int EG>
static cudaError_t launch_rollout_single_eg(int ns, const RolloutSingleArgs &a, dim3 grid, cudaStream_t st)
{
    if constexpr (D >= 2) { if (ns == D - 1) return launch_rollout_single_one<D, EG, D - 1>(a, grid, st); }
    if constexpr (D >= 3) { if (ns == D - 2) return launch_rollout_single_one<D, EG, D - 2>(a, grid, st); }
    return launch_rollout_single_one<D, EG, D>(a, grid, st);
}

#define GPMPC_CAT2(a, b) a##b
#define GPMPC_CAT(a, b) GPMPC_CAT2(a, b)
cudaError_t GPMPC_CAT(launch_pairs_batch_D, GPMPC_INST_D)(int EG, int grad_mode, int ns, const PairArgs &a,
                                                          dim3 grid, cudaStream_t st)
{
    return launch_any<GPMPC_INST_D, false>(EG, grad_mode, ns, a, grid, st);
}
cudaError_t GPMPC_CAT(launch_step_single_D, GPMPC_INST_D)(int EG, int grad_mode, int ns, const SingleStepArgs &a,
                                                          dim3 grid, cudaStream_t st)
{
    return launch_any<GPMPC_INST_D, true>(EG, grad_mode, ns, a, grid, st);
}

cudaError_t GPMPC_CAT(launch_rollout_single_D, GPMPC_INST_D)(int EG, int ns, const RolloutSingleArgs &a, dim3 grid, cudaStream_t st)
{
    switch (EG) {
        case 1: return launch_rollout_single_eg<GPMPC_INST_D, 1>(ns, a, grid, st);
        case 2: return launch_rollout_single_eg<GPMPC_INST_D, 2>(ns, a, grid, st);
        case 3: return launch_rollout_single_eg<GPMPC_INST_D, 3>(ns, a, grid, st);
        case 4: return launch_rollout_single_eg<GPMPC_INST_D, 4>(ns, a, grid, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace gpmpc
