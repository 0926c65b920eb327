// qt_launch.cuh -- programmatic dependent launch (PDL) for the kernels that run back to back inside a block's
// forward (GEMM -> rope -> GEMM -> softmax -> ...).  A kernel launched with the attribute may become resident while
// its predecessor on the stream is still draining; it does its private prologue (table staging, mbarrier init,
// TMEM allocation, tensor-map prefetch -- nothing the predecessor writes), then griddepcontrol.wait blocks until the
// predecessor has completed and its writes are visible.  Captured into CUDA graphs as programmatic edges.
// QT_PDL=0 turns the attribute off (plain stream order); the device-side instructions are then no-ops.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

namespace {
inline bool qt_pdl_enabled()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("QT_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t qt_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                             Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = qt_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// the same launch inside clusters of `cluster_x` CTAs along x (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t qt_launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     unsigned cluster_x, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster_x;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = qt_pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace
