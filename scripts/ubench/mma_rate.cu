// Microbenchmark (GPU box): cycles per tcgen05.mma (kind::f16, M=128, K=16) as a function of N, of where A comes
// from (shared memory vs TMEM) and of how many MMAs are issued between commits.  One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../feature-point-cnn_b200/csrc/tc_common.cuh"
using namespace spb200;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 db;\nmov.b64 db, {%2, %3};\nsetp.ne.b32 p, %5, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n}\n" ::"r"(d), "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc)
        : "memory");
}

// mode 0: SS, A SBO 1024; mode 1: SS, A SBO 1280 (halo view); mode 2: TS (A in TMEM)
template <int variant>
__global__ void __launch_bounds__(128) k(int N, int mode, int per_commit, int reps, unsigned long long* out) {
    extern __shared__ uint8_t dyn[];
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t slot;
    uint8_t* base = (uint8_t*)(((uintptr_t)dyn + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x / 32;
    for (int i = threadIdx.x; i < (64 * 1024) / 4; i += 128) ((uint32_t*)base)[i] = 0;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 0) {
        const uint32_t a_addr = smem_u32(base), b_addr = smem_u32(base + 24 * 1024);
        const uint32_t hiA = ((mode == 1 ? 1280u : 1024u) >> 4) | (1u << 14) | (2u << 29);
        const uint32_t hiB = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t alo = umma_desc_lo(a_addr), blo = umma_desc_lo(b_addr);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (variant == 0 && r >= 4) mbar_wait(&bar[r & 3], ((r >> 2) - 1) & 1);     // the commit of rep r-4 has landed
            if (variant == 2) mbar_wait(&bar[0], 1);   // fresh barrier: parity 1 is 'complete' -> returns at once
            if (elect_one()) {
                for (int i = 0; i < per_commit; i += 4) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        if (mode == 2) umma_ts(tm, tm + 256 + kk * 8, blo + kk * 2, hiB, idesc, 1u);
                        else umma_f16_w(tm, alo + kk * 2 + (i & 4) * 8, hiA, blo + kk * 2, hiB, idesc, 1u);
                    }
                }
                if (variant <= 1) umma_commit(&bar[r & 3]);
            }
            __syncwarp();
        }
        if (variant == 0) { for (int r = reps; r < reps + 4; ++r) mbar_wait(&bar[r & 3], ((r >> 2) - 1) & 1); }
        long long t1 = clock64();
        if (variant != 0) { if (elect_one()) umma_commit(&bar[0]); __syncwarp(); for (volatile int w = 0; w < 20000; ++w) {} }
        if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
    unsigned long long* d; cudaMalloc(&d, 148 * 8);
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const char* names[3] = {"SS sbo1024", "SS sbo1280", "TS (A in TMEM)"};
    const char* vn[4] = {"commit+wait", "commit only", "ready-wait only", "nothing"};
    for (int mode = 0; mode < 3; mode += 2)
        for (int N : {64, 128})
            for (int variant = 0; variant < 4; ++variant)
            for (int pc : {4, 8, 32}) {
                const int reps = 2000 / pc * 4;
                if (variant == 0) k<0><<<148, 128, 100 * 1024>>>(N, mode, pc, reps, d);
                if (variant == 1) k<1><<<148, 128, 100 * 1024>>>(N, mode, pc, reps, d);
                if (variant == 2) k<2><<<148, 128, 100 * 1024>>>(N, mode, pc, reps, d);
                if (variant == 3) k<3><<<148, 128, 100 * 1024>>>(N, mode, pc, reps, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                unsigned long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
                printf("%-16s N=%3d per_group=%2d %-16s: %.1f cycles per MMA (tensor floor %d)\n", names[mode], N, pc, vn[variant], (s / 148 - (variant ? 0 : 0)) / ((double)reps * pc), N / 2); fflush(stdout);
            }
    return 0;
}
