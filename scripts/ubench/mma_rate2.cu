// Microbenchmark 2 (GPU box): does interleaving independent accumulators hide the per-MMA latency of a dependent
// accumulation chain?  And what does an already-complete mbarrier wait cost the issuing warp (try_wait vs test_wait)?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../feature-point-cnn_b200/csrc/tc_common.cuh"
using namespace spb200;

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// NACC independent accumulators used round-robin; WAITK: 0 none, 1 try_wait (all lanes), 2 test_wait (all lanes),
// 3 try_wait by one elected lane + __syncwarp; a wait happens every `per_group` MMAs on an always-complete barrier.
template <int NACC, int WAITK>
__global__ void __launch_bounds__(128) k(int N, int per_group, int reps, unsigned long long* out) {
    extern __shared__ uint8_t dyn[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t slot;
    uint8_t* base = (uint8_t*)(((uintptr_t)dyn + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x / 32;
    for (int i = threadIdx.x; i < (64 * 1024) / 4; i += 128) ((uint32_t*)base)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 0) {
        const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t alo = umma_desc_lo(smem_u32(base)), blo = umma_desc_lo(smem_u32(base + 24 * 1024));
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (WAITK == 1) mbar_wait(&bar[0], 1);
            if (WAITK == 2) { while (!mbar_test_wait(&bar[0], 1)) {} }
            if (WAITK == 3) { if (elect_one()) mbar_wait(&bar[0], 1); __syncwarp(); }
            if (elect_one()) {
                for (int i = 0; i < per_group; i += 4) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_f16_w(tm + ((i / 4 * 4 + kk) % NACC) * (512 / NACC), alo + kk * 2, hi, blo + kk * 2, hi, idesc, 1u);
                }
            }
            __syncwarp();
        }
        long long t1 = clock64();
        if (elect_one()) umma_commit(&bar[1]);
        __syncwarp();
        mbar_wait(&bar[1], 0);
        if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    }
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

template <int NACC, int WAITK>
void run(int N, int pg, unsigned long long* d, const char* what) {
    cudaFuncSetAttribute(k<NACC, WAITK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 8192 / pg;
    k<NACC, WAITK><<<148, 128, 100 * 1024>>>(N, pg, reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    unsigned long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
    printf("N=%3d accumulators=%d group=%2d %-28s: %.1f cycles per MMA (floor %d)\n", N, NACC, pg, what, s / 148 / ((double)reps * pg), N / 2); fflush(stdout);
}

int main() {
    unsigned long long* d; cudaMalloc(&d, 148 * 8);
    for (int N : {64, 128}) {
        run<1, 0>(N, 32, d, "no wait");
        run<2, 0>(N, 32, d, "no wait");
        run<4, 0>(N, 32, d, "no wait");
        run<1, 0>(N, 8, d, "no wait");
        run<2, 0>(N, 8, d, "no wait");
        run<1, 1>(N, 8, d, "try_wait all lanes");
        run<1, 2>(N, 8, d, "test_wait all lanes");
        run<1, 3>(N, 8, d, "try_wait elected lane");
        run<2, 1>(N, 8, d, "try_wait all lanes");
        run<2, 2>(N, 8, d, "test_wait all lanes");
    }
    run<2, 0>(256, 32, d, "no wait");
    run<1, 0>(256, 32, d, "no wait");
    return 0;
}
