// b200_launch.cuh -- one launch path for every kernel of the backend, with Programmatic Dependent Launch (PDL).
//
// A UNet forward is ~1000 dependent kernels of 3-30 us; in a stream (or a captured CUDA graph) each boundary costs a full
// drain + launch (~2 us).  With the programmatic-stream-serialization launch attribute the NEXT kernel's CTAs are scheduled
// while the previous kernel is still running its tail; they park in `griddepcontrol.wait` (pdl_wait) until the predecessor has
// completed and flushed its memory.  Every kernel therefore
//     pdl_wait();                 // before its first global-memory access (after purely on-chip setup where there is any)
//     pdl_launch_dependents();    // immediately after: lets the successor's CTAs be scheduled as SM resources free up
// Correctness does not depend on the attribute: without it both instructions are no-ops.
#pragma once

#include <cuda_runtime.h>
#include <cstdlib>
#include <cstring>

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// ask L2 to fetch [p, p + bytes) from HBM (16-byte aligned, bytes % 16 == 0); fire and forget
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool b200_pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GGML_B200_PDL");
        v = (e && *e) ? (atoi(e) != 0) : 1;
    }
    return v != 0;
}

// fills `attrs` (room for 2) and returns the count: PDL, plus a cluster dimension when cluster_z > 1
inline unsigned b200_launch_attrs(cudaLaunchAttribute* attrs, unsigned cluster_z = 1) {
    unsigned n = 0;
    if (b200_pdl_enabled()) {
        attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster_z > 1) {
        attrs[n].id = cudaLaunchAttributeClusterDimension;
        attrs[n].val.clusterDim.x = 1;
        attrs[n].val.clusterDim.y = 1;
        attrs[n].val.clusterDim.z = cluster_z;
        ++n;
    }
    return n;
}

template <typename... KArgs, typename... Args>
inline cudaError_t b200_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[2];
    cfg.numAttrs = b200_launch_attrs(attrs);
    cfg.attrs = attrs;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
