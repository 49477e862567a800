// desc_probe.cu -- does a tcgen05 shared-memory matrix descriptor (K-major, SWIZZLE_128B) read 8-row groups that start at an arbitrary
// 128-byte row of a swizzled region, with a stride between groups that is NOT a multiple of 1024 bytes?  That is what an implicit-GEMM
// convolution needs to use ONE halo tile in shared memory for several filter taps (tap (kh, kw) = the same pixels shifted by kw rows and
// kh * (BW + 2) rows).  The probe fills a 256-row x 128-byte region in the address-based 128B swizzle the TMA produces, issues one
// M = 128, N = 64, K = 64 product with the A descriptor {start = base + shift rows, SBO = sbo rows, base_offset variant} and compares the
// accumulator with the product of the rows the descriptor SHOULD have selected.  Tuning / bring-up tool, not part of the product path.
#include "../csrc/kernels/sm100_ptx.cuh"

#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace sm100;

namespace {

constexpr int AROWS = 256, N = 64, K = 64;

__global__ void __launch_bounds__(128) k_probe(const __half* __restrict__ Aall, const __half* __restrict__ B, float* __restrict__ D, int shift, int sbo_rows,
                                               int bo_mode) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_smem;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                       // AROWS x 128 B
    uint8_t* sB = smem + AROWS * 128;         // N x 128 B (1024-aligned: AROWS * 128 is a multiple of 1024)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // address-based 128B swizzle: 16-byte chunk index ^= (row index & 7), the row index counted from a 1024-byte aligned base
    for (int i = threadIdx.x; i < AROWS * 8; i += blockDim.x) {
        const int r = i >> 3, ch = i & 7;
        *(uint4*)(sA + r * 128 + ((ch ^ (r & 7)) << 4)) = *(const uint4*)(Aall + r * K + ch * 8);
    }
    for (int i = threadIdx.x; i < N * 8; i += blockDim.x) {
        const int r = i >> 3, ch = i & 7;
        *(uint4*)(sB + r * 128 + ((ch ^ (r & 7)) << 4)) = *(const uint4*)(B + r * K + ch * 8);
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_smem, 64); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (warp == 0 && lane == 0) {
        const uint32_t a_addr = smem_u32(sA) + (uint32_t)shift * 128u;
        uint64_t da = 0;
        da |= (uint64_t)((a_addr >> 4) & 0x3FFF);
        da |= (uint64_t)1 << 16;
        da |= (uint64_t)(((uint32_t)sbo_rows * 128u) >> 4) << 32;
        da |= (uint64_t)1 << 46;
        const uint64_t bo = bo_mode == 1 ? (uint64_t)((a_addr >> 7) & 7u) : 0;
        da |= bo << 49;
        da |= (uint64_t)2 << 61;
        const uint64_t db = make_smem_desc_sw128(smem_u32(sB));
        constexpr uint32_t idesc = make_idesc(0, 128, N);
        for (int k = 0; k < K / 16; ++k) mma_f16(tmem_base, da + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : 0u);
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[row * N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

}  // namespace

int main() {
    std::vector<__half> hA((size_t)AROWS * K), hB((size_t)N * K);
    std::vector<float> fA(hA.size()), fB(hB.size());
    uint32_t x = 2463534242u;
    for (size_t i = 0; i < hA.size(); ++i) { x = x * 1664525u + 1013904223u; fA[i] = (float)((int)(x >> 20) % 17 - 8); hA[i] = __float2half(fA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { x = x * 1664525u + 1013904223u; fB[i] = (float)((int)(x >> 20) % 13 - 6); hB[i] = __float2half(fB[i]); }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    const int smem = AROWS * 128 + N * 128 + 1024;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> got(128 * N);
    const int shifts[] = {0, 1, 2, 3, 10, 11, 12, 21};
    const int sbos[] = {8, 10, 12, 16};
    printf("%-8s %-8s %-12s %s\n", "shift", "sbo_rows", "base_offset", "mismatching accumulator elements (of 8192); rows that differ");
    for (int sbo : sbos)
        for (int shift : shifts) {
            if (shift + 15 * sbo + 8 > AROWS) continue;
            for (int bo = 0; bo < 2; ++bo) {
                cudaMemset(dD, 0xff, 128 * N * 4);
                k_probe<<<1, 128, smem>>>(dA, dB, dD, shift, sbo, bo);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%-8d %-8d %-12s CUDA error: %s\n", shift, sbo, bo ? "addr>>7&7" : "0", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost);
                int bad = 0, bad_rows = 0;
                for (int i = 0; i < 128; ++i) {
                    const int ar = shift + (i / 8) * sbo + (i % 8);
                    int rb = 0;
                    for (int n = 0; n < N; ++n) {
                        float ref = 0.f;
                        for (int c = 0; c < K; ++c) ref += fA[(size_t)ar * K + c] * fB[(size_t)n * K + c];
                        if (ref != got[(size_t)i * N + n]) { ++bad; rb = 1; }
                    }
                    bad_rows += rb;
                }
                printf("%-8d %-8d %-12s %d elements in %d rows\n", shift, sbo, bo ? "addr>>7&7" : "0", bad, bad_rows);
            }
        }
    return 0;
}
