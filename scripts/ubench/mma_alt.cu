// Microbenchmark (GPU box): cycles per tcgen05.mma (kind::f16, M = 128, K = 16, N = 64 / 128) as a function of the shared-memory
// LAYOUT of the A operand: 128-byte swizzle (the residual blocks), the non-swizzled core-matrix layout the plane-fed stem uses
// (8 GEMM rows = 128 contiguous bytes, the second K half one parity plane = 4480 B further), the same with other plane pitches,
// and 32-byte swizzle (8 rows x 32 B atoms).  Operands are zeros - only the timing matters.  Groups of 8 MMAs per commit, four
// commits in flight, one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_layout mma_layout.cu && timeout 60 ./mma_layout
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../feature-point-cnn_b200/csrc/tc_common.cuh"
using namespace spb200;

__global__ void __launch_bounds__(128) k(int alt, int N, uint32_t a_lo_extra, uint32_t a_hi, uint32_t a_step, int reps, unsigned long long* out) {
    extern __shared__ uint8_t dyn[];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t slot;
    uint8_t* base = (uint8_t*)(((uintptr_t)dyn + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x / 32;
    for (int i = threadIdx.x; i < (96 * 1024) / 4; i += 128) ((uint32_t*)base)[i] = 0;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 0) {
        const uint32_t a_addr = smem_u32(base), b_addr = smem_u32(base + 64 * 1024);
        const uint32_t hiB = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t alo = ((a_addr >> 4) & 0x3fffu) | a_lo_extra, blo = umma_desc_lo(b_addr);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (r >= 4) mbar_wait(&bar[r & 3], ((r >> 2) - 1) & 1);
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) umma_f16_w(tm + (alt ? ((r & 1) * 2 + (kk & 1)) : (r & 3)) * 128, alo + (kk & 3) * a_step, a_hi, blo + (kk & 3) * 2, hiB, idesc, 1u);
                umma_commit(&bar[r & 3]);
            }
            __syncwarp();
        }
        for (int r = reps; r < reps + 4; ++r) mbar_wait(&bar[r & 3], ((r >> 2) - 1) & 1);
        long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
    unsigned long long* d; cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    struct L { const char* name; uint32_t lo_extra, hi, step; };
    const L layouts[] = {
        {"SWIZZLE_128B, SBO 1024            ", 1u << 16, (1024u >> 4) | (1u << 14) | (2u << 29), 2},
        {"no swizzle, SBO 128, LBO 4480 (stem)", (4480u >> 4) << 16, (128u >> 4) | (1u << 14), 8},
        {"no swizzle, SBO 128, LBO 4544      ", (4544u >> 4) << 16, (128u >> 4) | (1u << 14), 8},
        {"no swizzle, SBO 128, LBO 2048      ", (2048u >> 4) << 16, (128u >> 4) | (1u << 14), 8},
        {"no swizzle, SBO 256, LBO 128 (dense)", (128u >> 4) << 16, (256u >> 4) | (1u << 14), 256},
        {"SWIZZLE_32B, SBO 256               ", 1u << 16, (256u >> 4) | (1u << 14) | (6u << 29), 256},
        {"SWIZZLE_64B, SBO 512               ", 1u << 16, (512u >> 4) | (1u << 14) | (4u << 29), 2},
    };
    for (const L& l : layouts)
        for (int alt : {0, 1}) for (int N : {64, 96, 128}) {
            const int reps = 1000;
            k<<<148, 128, 100 * 1024>>>(alt, N, l.lo_extra, l.hi, l.step, reps, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s N=%d: error %s\n", l.name, N, cudaGetErrorString(e)); return 1; }
            unsigned long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
            printf("alt=%d A: %s N=%3d: %.1f cycles per MMA (tensor floor %d, operand-fetch floor %d)\n", alt, l.name, N, s / 148 / (reps * 8.0), N / 2, (4096 + 32 * N) / 128);
            fflush(stdout);
        }
    return 0;
}
