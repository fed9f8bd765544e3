// lbm_launch.cuh -- programmatic dependent launch (PDL) for the per-step kernel chain.
//
// Every iteration is a chain of small dependent launches (bulk -> fix-up -> bulk -> ...).  With
// the programmatic-stream-serialization attribute the next kernel's blocks are scheduled while the
// previous kernel drains its last wave; each kernel begins with pdl_wait() (griddepcontrol.wait),
// which returns only when the previous grid has completed and its writes are visible, so the
// data dependencies are exactly those of plain stream order -- only the launch latency
// (~2-3 us per launch, 10 % of a 2048 x 512 step) is hidden.  LBM_B200_PDL=0 turns it off.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace lbm {

__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// All blocks of this grid have started: the dependent grid may begin to occupy freed SM slots.
__device__ __forceinline__ void pdl_release() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

inline bool pdl_enabled() {
    static const bool on = [] {
        const char* v = std::getenv("LBM_B200_PDL");
        return !(v && v[0] == '0');
    }();
    return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace lbm
