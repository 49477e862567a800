// b200_common.h -- shared declarations of the B200-native ggml backend (libggml-b200.so).
//
// Layering inside the plugin:
//   b200_backend.cpp   ggml_backend_{reg,device,buffer_type,buffer}_i vtables + C-ABI exports
//                      (the drop-in boundary, ggml/src/ggml-backend-impl.h:17-230)
//   b200_graph.cpp     graph_compute: node walk, fusion planning, workspace, dispatch
//   kernels/*.cu       hand-written sm_100a kernels; each exposes a plain launcher declared in b200_ops.h
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "ggml.h"
#include "ggml-backend.h"

#define B200_MAX_DEVICES 16
#define B200_ALIGNMENT 256   // every tensor base: TMA needs 16 B, 128 B-swizzled smem boxes like 128 B; 256 keeps float4 + TMA happy

#define B200_CUDA_CHECK(expr)                                                                          \
    do {                                                                                               \
        cudaError_t err__ = (expr);                                                                    \
        if (err__ != cudaSuccess) {                                                                    \
            fprintf(stderr, "[ggml-b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(err__),       \
                    __FILE__, __LINE__, cudaGetErrorString(err__));                                    \
            abort();                                                                                   \
        }                                                                                              \
    } while (0)

// plain-old-data view of a ggml tensor handed to kernels (strides in BYTES, like ggml's nb[])
struct b200_td {
    void*   data;
    int32_t type;      // ggml_type
    int64_t ne[4];
    int64_t nb[4];
};

static inline b200_td b200_make_td(const ggml_tensor* t) {
    b200_td d;
    d.data = t->data;
    d.type = (int32_t)t->type;
    for (int i = 0; i < 4; ++i) {
        d.ne[i] = t->ne[i];
        d.nb[i] = (int64_t)t->nb[i];
    }
    return d;
}

// per-backend-instance statistics, exported through the proc-address extension
// "ggml_backend_b200_get_stats" (include/ggml-b200.h)
struct b200_stats {
    uint64_t graphs;            // graph_compute calls
    uint64_t kernel_launches;   // kernels launched by this backend instance (all of them are ours: no library calls)
    uint64_t nodes_executed;    // ggml nodes covered (a fused kernel covers several)
    uint64_t fused_nodes;       // nodes that were absorbed into a neighbour's kernel
    double   last_graph_ms;     // device time of the last graph_compute (CUDA events on the backend stream)
    double   total_graph_ms;
    uint64_t tc_gemm_launches;  // tcgen05 GEMM launches (subset of kernel_launches)
    uint64_t reserved[8];
    uint64_t ext[16];           // see include/ggml-b200.h
    uint64_t side_launches;     // projections launched on a side stream (beside the main stream's kernels)
};

struct b200_device_info {
    int    id;
    int    cc_major, cc_minor;
    int    sm_count;
    size_t total_mem;
    size_t smem_optin;
    char   name[64];
    char   desc[256];
};

// driver entry points fetched at run time (no link-time dependency on libcuda.so, so the plugin
// loads -- and reports zero devices -- on a host without a driver)
typedef CUresult (*b200_cuTensorMapEncodeTiled_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
b200_cuTensorMapEncodeTiled_t b200_get_tensormap_encoder();
