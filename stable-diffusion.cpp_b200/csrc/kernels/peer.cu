// peer.cu -- the CFG-batch split's exchange step as device code over NVLink peer memory (north star: "a single
// all-gather on the latent").
//
// Two processes, one per GPU, evaluate the cond / uncond branch of the same image.  Each owns a MAILBOX in its HBM
//     [ payload, parity 0 | payload, parity 1 | flag (u32, written by the peer) | seq (u32, local push counter) | error (u32) ]
// exported to the other process with cudaIpcGetMemHandle and mapped there (NVLink peer mapping).  At the end of every model call the
// producing rank stores its eps prediction straight into the PEER's mailbox -- by the epilogue of the convolution that computes it
// (gemm_tc2.cu writes every output element to its ggml tensor and to the peer mapping: compute + collective in one kernel), or by
// k_peer_push below when another kernel produced the tensor -- then publishes a sequence number in the peer's flag and waits for the
// peer's.  No NCCL call, no host staging: 64 KB (SD1.5) / 256 KB (SDXL) cross the link once per step and the wait costs one round trip.
// Payload slots alternate by the parity of the sequence number, so a rank that is already computing step t + 1 can never overwrite
// the data of step t its peer is still reading (it cannot get two steps ahead: every step needs the peer's flag).
#include "../b200_ops.h"
#include "b200_launch.cuh"

#include <algorithm>

namespace {

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// payload -> the peer's mailbox slot of parity (seq + 1) & 1   (seq = pushes completed so far)
__global__ void __launch_bounds__(256) k_peer_push(const float4* __restrict__ src, float4* peer_base, const unsigned* __restrict__ seq, size_t n_vec,
                                                   size_t slot_vec) {
    pdl_wait();
    pdl_launch_dependents();
    float4* dst = peer_base + (size_t)((*seq + 1u) & 1u) * slot_vec;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// one thread: count this push, publish the count in the peer's flag, wait until the peer has published the same count in ours
__global__ void k_peer_signal_wait(unsigned* my_seq, unsigned* peer_flag, const unsigned* my_flag, unsigned* my_err, unsigned long long timeout_ns) {
    pdl_wait();                 // the payload stores of the previous kernel(s) have completed and are visible system-wide
    pdl_launch_dependents();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned k = *my_seq + 1u;
    *my_seq = k;
    __threadfence_system();
    st_release_sys(peer_flag, k);
    const unsigned long long t0 = gtime();
    while ((int)(ld_acquire_sys(my_flag) - k) < 0) {
        if (gtime() - t0 > timeout_ns) { *my_err = k; break; }      // the peer never arrived: report instead of hanging the GPU
        __nanosleep(200);
    }
}

}  // namespace

int b200_launch_peer_push(cudaStream_t s, const void* src, void* peer_base, const unsigned* seq, size_t bytes, size_t slot_bytes) {
    if ((bytes & 15) || (slot_bytes & 15) || ((uintptr_t)src & 15) || ((uintptr_t)peer_base & 15)) return -1;
    const size_t n = bytes / 16;
    const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148);
    b200_launch(k_peer_push, dim3(blocks), dim3(256), 0, s, (const float4*)src, (float4*)peer_base, seq, n, slot_bytes / 16);
    return 1;
}

int b200_launch_peer_signal_wait(cudaStream_t s, unsigned* my_seq, unsigned* peer_flag, const unsigned* my_flag, unsigned* my_err, double timeout_s) {
    b200_launch(k_peer_signal_wait, dim3(1), dim3(32), 0, s, my_seq, peer_flag, my_flag, my_err, (unsigned long long)(timeout_s * 1e9));
    return 1;
}
