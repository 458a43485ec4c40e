// Microbenchmark 3 (GPU box): the mechanics of tcgen05.mma.cta_group::2 for the planned CTA-pair version of
// halo_tc.cu.  A cluster of two CTAs computes D[256, N] = A[256, 64] . B[N, 64]^T:
// every CTA holds its 128 rows of A and its N / 2 rows of B in its own shared memory (K-major, 128-byte swizzle, same
// offsets in both CTAs), the leader CTA issues the MMAs, one multicast commit arrives on a barrier in both CTAs and
// every CTA reads its 128 accumulator rows from its own tensor memory.
//   (1) correctness with small integer operands (exact in fp16 / fp32), N = 64, 128, 256;
//   (2) cycles per pair-MMA against the same shapes on one CTA (mma_rate.cu), operands resident;
//   (3) the producer side: both CTAs fetch their operands by TMA (.cta_group::2) and complete the transaction bytes on
//       ONE barrier in the leader CTA (address mapped with mapa); the epilogue warps of both CTAs arrive on a barrier
//       of the leader (mbarrier.arrive.shared::cluster) - the hand-overs the paired kernel needs.  All waits of (3) are
//       bounded, a hand-over that never happens is reported instead of hanging;
//   (4) the second GEMM of a fused block on the pair: Y = fp16(D1[:, 0:64]) packed into each CTA's own tensor memory by
//       its epilogue warps, D2 = Y . B2^T with the A operand read from tensor memory (TS form of the pair MMA).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o mma_pair mma_pair.cu ; run under `timeout 20`.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../feature-point-cnn_b200/csrc/tc_common.cuh"
using namespace spb200;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one warp of EACH CTA of the pair executes the allocation (the pattern of the libraries this was taken from)
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "setp.ne.b32 p, %6, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this shared-memory offset in every CTA of the mask when the MMAs issued so far are complete
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__host__ __device__ inline int a_val(int m, int k) { return (m + 2 * k) % 5 - 2; }
__host__ __device__ inline int b_val(int n, int k) { return (3 * n + k) % 7 - 3; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
kpair(int N, int reps, int check, unsigned long long* cyc, int* errs) {
    extern __shared__ uint8_t dyn[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    uint8_t* base = dyn + ((1024u - (smem_u32(dyn) & 1023u)) & 1023u);
    uint8_t* s_a = base;                 // [128 rows][64 K] fp16, 128-byte rows, 16-byte chunks XOR (row & 7)
    uint8_t* s_b = base + 16384;         // [N / 2 rows][64 K]
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int rank = (int)cluster_ctarank();
    const int nb = N / 2;
    for (int i = threadIdx.x; i < 128 * 64; i += 128) {
        const int r = i / 64, k = i % 64;
        const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
        *reinterpret_cast<__half*>(s_a + off) = __int2half_rn(a_val(rank * 128 + r, k));
        if (r < nb) *reinterpret_cast<__half*>(s_b + off) = __int2half_rn(b_val(rank * nb + r, k));
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) tmem_alloc2(&slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tm = slot;
    long long t0 = clock64();
    if (rank == 0 && warp == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t alo = umma_desc_lo(smem_u32(s_a)), blo = umma_desc_lo(smem_u32(s_b));
        for (int r = 0; r < reps; ++r) {
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma2_f16_w(tm, alo + kk * 2, hi, blo + kk * 2, hi, idesc, (r | kk) ? 1u : 0u);
            }
            __syncwarp();
        }
        if (elect_one()) umma2_commit_mc(&bar, 3);
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    tc_fence_after();
    if (check) {
        const int m = rank * 128 + warp * 32 + lane;
        int bad = 0;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) {
                int want = 0;
                for (int k = 0; k < 64; ++k) want += a_val(m, k) * b_val(c0 + j, k);
                if (__uint_as_float(r[j]) != (float)(want * reps)) ++bad;
            }
        }
        if (bad) atomicAdd(errs + rank, bad);
    }
    if (threadIdx.x == 0 && rank == 0) cyc[blockIdx.x / 2] = (unsigned long long)(t1 - t0);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) { tc_fence_after(); tmem_dealloc2(tm, 512); }
}

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// TMA load of a CTA pair: the data lands in THIS CTA's shared memory, the bytes complete on a barrier given by its
// shared::cluster address (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, int spins) {
    for (int i = 0; i < spins; ++i)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}

// flags[0]: operands never arrived on the leader's barrier, [1]: the commit never arrived (per CTA: [1], [2]),
// [3]: the 8 remote / local arrivals never completed the leader's barrier
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
kpair_tma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int N, int* errs, int* flags,
          uint32_t* addrs) {
    extern __shared__ uint8_t dyn[];
    __shared__ __align__(8) uint64_t full, bar, done;
    __shared__ uint32_t slot;
    uint8_t* base = dyn + ((1024u - (smem_u32(dyn) & 1023u)) & 1023u);
    uint8_t* s_a = base;
    uint8_t* s_b = base + 16384;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int rank = (int)cluster_ctarank();
    const int nb = N / 2;
    if (threadIdx.x == 0) {
        mbar_init(&full, 1);
        mbar_init(&bar, 1);
        mbar_init(&done, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        addrs[rank * 2] = smem_u32(&full);                      // what a shared-window address looks like in each CTA
        addrs[rank * 2 + 1] = mapa_u32(smem_u32(&full), 0);     // and the leader's barrier seen from here
    }
    if (warp == 0) tmem_alloc2(&slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t full_leader = mapa_u32(smem_u32(&full), 0);
    if (warp == 1 && lane == 0) {
        // producer of this CTA: the leader announces the bytes of BOTH CTAs, then each CTA fetches its own rows
        if (rank == 0) mbar_expect_tx(&full, (uint32_t)(2 * (16384 + nb * 128)));
        tma_load_2d_pair(smem_u32(s_a), &tmA, full_leader, 0, rank * 128);
        tma_load_2d_pair(smem_u32(s_b), &tmB, full_leader, 0, rank * nb);
    }
    bool ok = true;
    if (rank == 0 && warp == 0) {
        ok = mbar_wait_bounded(&full, 0, 200000);
        if (!ok && lane == 0) flags[0] = 1;
        tc_fence_after();
        if (ok) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t alo = umma_desc_lo(smem_u32(s_a)), blo = umma_desc_lo(smem_u32(s_b));
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma2_f16_w(tm, alo + kk * 2, hi, blo + kk * 2, hi, idesc, kk ? 1u : 0u);
                umma2_commit_mc(&bar, 3);
            }
            __syncwarp();
        }
    }
    const bool got = mbar_wait_bounded(&bar, 0, 300000);
    if (!got && threadIdx.x == 0) flags[1 + rank] = 1;
    tc_fence_after();
    if (got) {
        const int m = rank * 128 + warp * 32 + lane;
        int bad = 0;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) {
                int want = 0;
                for (int k = 0; k < 64; ++k) want += a_val(m, k) * b_val(c0 + j, k);
                if (__uint_as_float(r[j]) != (float)want) ++bad;
            }
        }
        if (bad) atomicAdd(errs + rank, bad);
    }
    // every warp of both CTAs reports to the leader (the y_full / accumulator-drained hand-over of the paired kernel)
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&done), 0));
    if (rank == 0 && threadIdx.x == 0 && !mbar_wait_bounded(&done, 0, 300000)) flags[3] = 1;
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) { tc_fence_after(); tmem_dealloc2(tm, 512); }
}

__device__ __forceinline__ void umma2_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 db;\n"
        "mov.b64 db, {%2, %3};\n"
        "setp.ne.b32 p, %5, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__host__ __device__ inline int b2_val(int n, int k) { return (n + 5 * k) % 3 - 1; }

// N = 128: D1 (columns 0..127) = A . B^T;  Y (columns 256..287, 64 packed fp16) = D1[:, 0:64];  D2 (columns 384..511) = Y . B2^T
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) kpair_ts(int* errs, int* flags) {
    constexpr int N = 128, nb = 64;
    extern __shared__ uint8_t dyn[];
    __shared__ __align__(8) uint64_t bar, bar2, yfull;
    __shared__ uint32_t slot;
    uint8_t* base = dyn + ((1024u - (smem_u32(dyn) & 1023u)) & 1023u);
    uint8_t* s_a = base;
    uint8_t* s_b = base + 16384;
    uint8_t* s_b2 = base + 24576;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int rank = (int)cluster_ctarank();
    for (int i = threadIdx.x; i < 128 * 64; i += 128) {
        const int r = i / 64, k = i % 64;
        const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
        *reinterpret_cast<__half*>(s_a + off) = __int2half_rn(a_val(rank * 128 + r, k));
        if (r < nb) {
            *reinterpret_cast<__half*>(s_b + off) = __int2half_rn(b_val(rank * nb + r, k));
            *reinterpret_cast<__half*>(s_b2 + off) = __int2half_rn(b2_val(rank * nb + r, k));
        }
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_init(&bar2, 1);
        mbar_init(&yfull, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) tmem_alloc2(&slot, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    if (rank == 0 && warp == 0) {
        const uint32_t alo = umma_desc_lo(smem_u32(s_a)), blo = umma_desc_lo(smem_u32(s_b));
        if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma2_f16_w(tm, alo + kk * 2, hi, blo + kk * 2, hi, idesc, kk ? 1u : 0u);
            umma2_commit_mc(&bar, 3);
        }
        __syncwarp();
    }
    bool ok = mbar_wait_bounded(&bar, 0, 300000);
    if (!ok && threadIdx.x == 0) flags[rank] = 1;
    tc_fence_after();
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    if (ok) {
        for (int blk = 0; blk < 2; ++blk) {
            uint32_t r[32], y[16];
            tmem_ld_32x32(tm + lane_off + (uint32_t)(blk * 32), r);
            tmem_ld_wait();
            for (int e = 0; e < 16; ++e) y[e] = pack2<__half>(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
            tmem_st_32x16(tm + lane_off + 256u + (uint32_t)(blk * 16), y);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&yfull), 0));
    if (rank == 0 && warp == 0) {
        const bool yok = mbar_wait_bounded(&yfull, 0, 300000);
        if (!yok && lane == 0) flags[2] = 1;
        tc_fence_after();
        if (yok) {
            const uint32_t b2lo = umma_desc_lo(smem_u32(s_b2));
            if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma2_f16_ts(tm + 384u, tm + 256u + (uint32_t)(kk * 8), b2lo + kk * 2, hi, idesc, kk ? 1u : 0u);
                umma2_commit_mc(&bar2, 3);
            }
            __syncwarp();
        }
    }
    const bool ok2 = mbar_wait_bounded(&bar2, 0, 300000);
    if (!ok2 && threadIdx.x == 0) flags[3] = 1;
    tc_fence_after();
    if (ok2) {
        const int m = rank * 128 + warp * 32 + lane;
        int yv[64];
        for (int k = 0; k < 64; ++k) {
            int v = 0;
            for (int kk = 0; kk < 64; ++kk) v += a_val(m, kk) * b_val(k, kk);
            yv[k] = v;
        }
        int bad = 0;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tm + lane_off + 384u + (uint32_t)c0, r);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) {
                int want = 0;
                for (int k = 0; k < 64; ++k) want += yv[k] * b2_val(c0 + j, k);
                if (__uint_as_float(r[j]) != (float)want) ++bad;
            }
        }
        if (bad) atomicAdd(errs + rank, bad);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) { tc_fence_after(); tmem_dealloc2(tm, 512); }
}

int main() {
    unsigned long long* d_cyc;
    int* d_err;
    cudaMalloc(&d_cyc, 74 * 8);
    cudaMalloc(&d_err, 2 * sizeof(int));
    const int smem = 33 * 1024 + 1024;
    cudaFuncSetAttribute(kpair, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int N : {64, 128, 256}) {
        cudaMemset(d_err, 0, 2 * sizeof(int));
        kpair<<<2, 128, smem>>>(N, 1, 1, d_cyc, d_err);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d check: error %s\n", N, cudaGetErrorString(e)); return 1; }
        int h[2];
        cudaMemcpy(h, d_err, sizeof(h), cudaMemcpyDeviceToHost);
        printf("N=%3d  M=256 over a CTA pair: wrong accumulator values: leader %d, peer %d (of %d each)\n", N, h[0], h[1], 128 * N);
        fflush(stdout);
    }
    for (int N : {64, 128, 256}) {
        const int reps = 2000;
        kpair<<<148, 128, smem>>>(N, reps, 0, d_cyc, d_err);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d rate: error %s\n", N, cudaGetErrorString(e)); return 1; }
        unsigned long long h[74];
        cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0;
        for (int i = 0; i < 74; ++i) s += (double)h[i];
        printf("N=%3d  %.1f cycles per pair-MMA (256 x %d x 16; tensor floor %d; operand reads per CTA %d B)\n", N,
               s / 74 / (reps * 4.0), N, N / 2, 4096 + N * 16);
        fflush(stdout);
    }
    // (3) TMA into both CTAs completing on the leader's barrier, remote arrivals
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess || !sym) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    EncodeFn encode = reinterpret_cast<EncodeFn>(sym);
    cudaFuncSetAttribute(kpair_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int* d_flags;
    uint32_t* d_addrs;
    cudaMalloc(&d_flags, 4 * sizeof(int));
    cudaMalloc(&d_addrs, 4 * sizeof(uint32_t));
    for (int N : {64, 128, 256}) {
        __half* hA = new __half[256 * 64];
        __half* hB = new __half[N * 64];
        for (int m = 0; m < 256; ++m) for (int k = 0; k < 64; ++k) hA[m * 64 + k] = __float2half((float)a_val(m, k));
        for (int n = 0; n < N; ++n) for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2half((float)b_val(n, k));
        __half *dA, *dB;
        cudaMalloc(&dA, 256 * 64 * 2); cudaMalloc(&dB, N * 64 * 2);
        cudaMemcpy(dA, hA, 256 * 64 * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * 64 * 2, cudaMemcpyHostToDevice);
        CUtensorMap tmA, tmB;
        cuuint32_t estr[2] = {1, 1};
        cuuint64_t dimsA[2] = {64, 256}, dimsB[2] = {64, (cuuint64_t)N}, str[1] = {128};
        cuuint32_t boxA[2] = {64, 128}, boxB[2] = {64, (cuuint32_t)(N / 2)};
        CUresult r1 = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dA, dimsA, str, boxA, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CUresult r2 = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dimsB, str, boxB, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode failed %d %d\n", (int)r1, (int)r2); return 1; }
        cudaMemset(d_err, 0, 2 * sizeof(int));
        cudaMemset(d_flags, 0, 4 * sizeof(int));
        kpair_tma<<<2, 128, smem>>>(tmA, tmB, N, d_err, d_flags, d_addrs);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d TMA pair: error %s\n", N, cudaGetErrorString(e)); return 1; }
        int h[2], f[4];
        uint32_t ad[4];
        cudaMemcpy(h, d_err, sizeof(h), cudaMemcpyDeviceToHost);
        cudaMemcpy(f, d_flags, sizeof(f), cudaMemcpyDeviceToHost);
        cudaMemcpy(ad, d_addrs, sizeof(ad), cudaMemcpyDeviceToHost);
        printf("N=%3d  TMA pair: operands arrived %s, commit seen leader %s peer %s, 8 arrivals on the leader %s; wrong values leader %d peer %d; "
               "barrier address leader 0x%x (mapa 0x%x), peer 0x%x (mapa to leader 0x%x)\n", N, f[0] ? "NO" : "yes", f[1] ? "NO" : "yes",
               f[2] ? "NO" : "yes", f[3] ? "NO" : "yes", h[0], h[1], ad[0], ad[1], ad[2], ad[3]);
        fflush(stdout);
        delete[] hA; delete[] hB; cudaFree(dA); cudaFree(dB);
    }
    // (4) second GEMM with A from tensor memory
    cudaFuncSetAttribute(kpair_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaMemset(d_err, 0, 2 * sizeof(int));
    cudaMemset(d_flags, 0, 4 * sizeof(int));
    kpair_ts<<<2, 128, smem>>>(d_err, d_flags);
    {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("TS pair: error %s\n", cudaGetErrorString(e)); return 1; }
        int h[2], f[4];
        cudaMemcpy(h, d_err, sizeof(h), cudaMemcpyDeviceToHost);
        cudaMemcpy(f, d_flags, sizeof(f), cudaMemcpyDeviceToHost);
        printf("N=128  TS pair (Y from each CTA's tensor memory): GEMM 1 seen leader %s peer %s, Y hand-over %s, GEMM 2 seen %s; wrong D2 values leader %d peer %d (of 16384 each)\n",
               f[0] ? "NO" : "yes", f[1] ? "NO" : "yes", f[2] ? "NO" : "yes", f[3] ? "NO" : "yes", h[0], h[1]);
    }
    return 0;
}
