// Fused residual block on tcgen05 / TMEM / TMA (sm_100a), persistent.
//
// One kernel computes a whole block of the reference network (python/src/resnet_blocks.py:14-27):
//     Y   = relu(conv3x3_s(X) * bn1)                      GEMM 1   D1[128, N] = im2col(X) . W1^T
//     OUT = relu(conv1x1(Y) * bn2 + shortcut(X))          GEMM 2   D2[128, N] = Y . W2^T (+ Xc . Wd^T)
// for a tile of 128 output pixels.  Y never leaves the SM: the epilogue warps read D1 from TMEM, add
// the folded bias, apply ReLU, round to the 16-bit operand type and write it into shared memory in the
// 128B-swizzled K-major layout that GEMM 2 reads as its A operand.  The 1x1 shortcut convolution of
// the first block of a layer is extra K of GEMM 2 (its A tiles are the centre-tap pixels of X, fetched
// by TMA); an identity shortcut is added in the second epilogue.  Channel concatenation
// (python/src/superpoint.py:59) is extra K segments of both GEMMs.
//
// Persistent CTAs (grid = SMs x residency) walk the tile list; warp 0 = TMA producer (one ring for the
// A/B tiles of both GEMMs), warp 1 = MMA issuer, warps 2-5 = epilogue.  D1 and D2 live in separate
// TMEM column ranges, so epilogue 2 of tile i overlaps GEMM 1 of tile i+1.
// With ksteps2 == 0 the kernel is a plain convolution (used for the transposed-conv phases).
#include <cuda.h>

#include <cstring>
#include <memory>

#include "kernels.h"
#include "tc_common.cuh"

namespace spb200 {

struct BlkSeg {
    int ntaps, nchunks, stride, C, kk_last;       // kk_last: K=16 MMA steps of the last chunk (1..4)
    int8_t dy[kMaxTaps], dx[kMaxTaps];
};

struct BlkParams {
    CUtensorMap tmA[kMaxSegs];
    CUtensorMap tmW1, tmW2;
    CUtensorMap tmD, tmR;      // destination / identity shortcut, box = the 32 pixel rows of an epilogue warp x 64 channels
    BlkSeg seg[kMaxSegs];
    int nseg;                  // GEMM 1 segments
    int nds;                   // GEMM 2 shortcut segments (0 = identity or none); they reuse tmA[0..nds)
    int y_chunks, y_kk_last;   // GEMM 2 K over Y
    int ksteps2;               // 0 = single GEMM
    int th, tw, tiles_x, tiles_per_img, total_tiles;
    int OH, OW;
    const float* bias1;
    const float* bias2;
    const void* residual;
    void* dst;
    int res_C, dst_H, dst_W, dst_C, dst_stride, dst_off_y, dst_off_x;
    int relu, dst_fp32, n_mma;
    int tma_store, tma_res;    // the output leaves (the shortcut arrives) through the Y buffer and TMA, see halo_tc.cu
};

constexpr int kBlkThreads = 192;
constexpr int kBlkABytes = 128 * 128;

template <int BLOCK_N, int STAGES, typename T>
__global__ void __launch_bounds__(kBlkThreads) block_tc_kernel(const __grid_constant__ BlkParams p) {
    constexpr int kBBytes = BLOCK_N * 128;
    constexpr int kStageBytes = kBlkABytes + kBBytes;
    constexpr int kYChunks = BLOCK_N / 64;
    constexpr uint32_t kTmemCols = 2 * BLOCK_N;            // D1 | D2  (128, 256 or 512: powers of two)

    extern __shared__ uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
    __shared__ __align__(8) uint64_t d1_full, d1_empty, y_full[kYChunks], y_empty, d2_full, d2_empty;   // y_full: one per 64-channel chunk of Y
    __shared__ __align__(8) uint64_t res_full[4];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_bias1[BLOCK_N], s_bias2[BLOCK_N];

    // aligned up to 1024 B by pointer ARITHMETIC on dyn_smem: the compiler keeps the shared address space (LDS/STS,
    // 32-bit addresses) instead of falling back to generic loads
    uint8_t* ring = dyn_smem + ((1024u - (smem_u32(dyn_smem) & 1023u)) & 1023u);
    uint8_t* s_y = ring + STAGES * kStageBytes;            // [kYChunks][128 rows][128 B]
    // shuffle: makes the warp index warp-uniform for the compiler (role loops then run on the uniform datapath)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x / 32), 0), lane = threadIdx.x % 32;
    const uint32_t idesc = (1u << 4) | (OperandFmt<T>::value << 7) | (OperandFmt<T>::value << 10) |
                           ((uint32_t)(p.n_mma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const bool fused = p.ksteps2 > 0;
    pdl_trigger();

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&d1_full, 1); mbar_init(&y_empty, 1); mbar_init(&d2_full, 1);
        mbar_init(&d1_empty, 4); mbar_init(&d2_empty, 4);
        for (int c = 0; c < kYChunks; ++c) mbar_init(&y_full[c], 4);
        for (int w = 0; w < 4; ++w) mbar_init(&res_full[w], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) prefetch_tmap(&p.tmA[s]);
        prefetch_tmap(&p.tmW1);
        if (fused) prefetch_tmap(&p.tmW2);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, kTmemCols);
    for (int i = threadIdx.x; i < BLOCK_N; i += kBlkThreads) {
        s_bias1[i] = p.bias1[i];
        s_bias2[i] = fused ? p.bias2[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d1 = tmem_slot, tmem_d2 = tmem_slot + BLOCK_N;

    if (warp == 0) {
        // ============================ TMA producer ============================
        // whole warp runs the loops; only the TMA issue sits under elect.sync (see halo_tc.cu)
        pdl_wait();                                        // the activations are the previous kernel's output
        {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int img = tile / p.tiles_per_img, tt = tile % p.tiles_per_img;
                const int y0 = (tt / p.tiles_x) * p.th, x0 = (tt % p.tiles_x) * p.tw;
                int k1 = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const BlkSeg& sg = p.seg[s];
                    for (int t = 0; t < sg.ntaps; ++t) {
                        int cx, cy, cp, cc;
                        if (sg.stride == 1) { cx = x0 + sg.dx[t]; cy = y0 + sg.dy[t]; cp = 0; cc = 0; }
                        else {
                            const int px = sg.dx[t] & 1, py = sg.dy[t] & 1;
                            cx = x0 + (sg.dx[t] - px) / 2; cy = y0 + (sg.dy[t] - py) / 2; cp = py; cc = px * sg.C;
                        }
                        for (int c = 0; c < sg.nchunks; ++c, ++k1) {
                            mbar_wait(&empty_bar[stage], phase ^ 1u);
                            uint8_t* a_dst = ring + stage * kStageBytes;
                            if (elect_one()) {
                                mbar_expect_tx(&full_bar[stage], (uint32_t)kStageBytes);
                                tma_load_5d(a_dst, &p.tmA[s], &full_bar[stage], cc + c * 64, cx, cp, cy, img);
                                tma_load_2d(a_dst + kBlkABytes, &p.tmW1, &full_bar[stage], k1 * 64, 0);
                            }
                            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
                if (fused) {
                    int k2 = 0;
                    for (int c = 0; c < p.y_chunks; ++c, ++k2) {           // A = Y (already in smem): weights only
                        mbar_wait(&empty_bar[stage], phase ^ 1u);
                        uint8_t* a_dst = ring + stage * kStageBytes;
                        if (elect_one()) {
                            mbar_expect_tx(&full_bar[stage], (uint32_t)kBBytes);
                            tma_load_2d(a_dst + kBlkABytes, &p.tmW2, &full_bar[stage], k2 * 64, 0);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                    for (int s = 0; s < p.nds; ++s) {                      // 1x1 shortcut: centre pixels of X
                        const BlkSeg& sg = p.seg[s];
                        for (int c = 0; c < sg.nchunks; ++c, ++k2) {
                            mbar_wait(&empty_bar[stage], phase ^ 1u);
                            uint8_t* a_dst = ring + stage * kStageBytes;
                            if (elect_one()) {
                                mbar_expect_tx(&full_bar[stage], (uint32_t)kStageBytes);
                                tma_load_5d(a_dst, &p.tmA[s], &full_bar[stage], c * 64, x0, 0, y0, img);
                                tma_load_2d(a_dst + kBlkABytes, &p.tmW2, &full_bar[stage], k2 * 64, 0);
                            }
                            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================ MMA issuer ============================
        {
            int stage = 0;
            uint32_t phase = 0, tph = 0;
            const uint32_t y_addr = smem_u32(s_y);
            constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);     // descriptor high word: SBO 1024, SWIZZLE_128B
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, tph ^= 1u) {
                mbar_wait(&d1_empty, tph ^ 1u);            // epilogue 1 of the previous tile has drained D1
                tc_fence_after();
                uint32_t acc = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const BlkSeg& sg = p.seg[s];
                    for (int t = 0; t < sg.ntaps; ++t)
                        for (int c = 0; c < sg.nchunks; ++c) {
                            mbar_wait(&full_bar[stage], phase);
                            tc_fence_after();
                            const uint32_t a_addr = smem_u32(ring + stage * kStageBytes), b_addr = a_addr + kBlkABytes;
                            const uint32_t alo = umma_desc_lo(a_addr), blo = umma_desc_lo(b_addr);
                            const int nkk = (c == sg.nchunks - 1) ? sg.kk_last : 4;
                            if (elect_one()) {
                                umma_f16_w(tmem_d1, alo, kHi, blo, kHi, idesc, acc);
                                if (nkk > 1) umma_f16_w(tmem_d1, alo + 2, kHi, blo + 2, kHi, idesc, 1u);
                                if (nkk > 2) umma_f16_w(tmem_d1, alo + 4, kHi, blo + 4, kHi, idesc, 1u);
                                if (nkk > 3) umma_f16_w(tmem_d1, alo + 6, kHi, blo + 6, kHi, idesc, 1u);
                                umma_commit(&empty_bar[stage]);
                            }
                            acc = 1u;
                            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                        }
                }
                if (elect_one()) umma_commit(&d1_full);
                if (fused) {
                    // GEMM 2 follows the first epilogue chunk by chunk: the MMAs over the first 64 channels of Y run while the
                    // epilogue warps convert the next ones (Y chunk c is complete when y_full[c] flips)
                    mbar_wait(&d2_empty, tph ^ 1u);        // epilogue 2 of the previous tile has drained D2
                    acc = 0;
                    for (int c = 0; c < p.y_chunks; ++c) {
                        mbar_wait(&y_full[c], tph);        // this chunk of the Y tile has been written by the epilogue warps
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t b_addr = smem_u32(ring + stage * kStageBytes) + kBlkABytes;
                        const uint32_t a_addr = y_addr + c * kBlkABytes;
                        const uint32_t alo = umma_desc_lo(a_addr), blo = umma_desc_lo(b_addr);
                        const int nkk = (c == p.y_chunks - 1) ? p.y_kk_last : 4;
                        if (elect_one()) {
                            umma_f16_w(tmem_d2, alo, kHi, blo, kHi, idesc, acc);
                            if (nkk > 1) umma_f16_w(tmem_d2, alo + 2, kHi, blo + 2, kHi, idesc, 1u);
                            if (nkk > 2) umma_f16_w(tmem_d2, alo + 4, kHi, blo + 4, kHi, idesc, 1u);
                            if (nkk > 3) umma_f16_w(tmem_d2, alo + 6, kHi, blo + 6, kHi, idesc, 1u);
                            umma_commit(&empty_bar[stage]);
                        }
                        acc = 1u;
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                    for (int s = 0; s < p.nds; ++s) {
                        const BlkSeg& sg = p.seg[s];
                        for (int c = 0; c < sg.nchunks; ++c) {
                            mbar_wait(&full_bar[stage], phase);
                            tc_fence_after();
                            const uint32_t a_addr = smem_u32(ring + stage * kStageBytes), b_addr = a_addr + kBlkABytes;
                            const uint32_t alo = umma_desc_lo(a_addr), blo = umma_desc_lo(b_addr);
                            const int nkk = (c == sg.nchunks - 1) ? sg.kk_last : 4;
                            if (elect_one()) {
                                umma_f16_w(tmem_d2, alo, kHi, blo, kHi, idesc, 1u);
                                if (nkk > 1) umma_f16_w(tmem_d2, alo + 2, kHi, blo + 2, kHi, idesc, 1u);
                                if (nkk > 2) umma_f16_w(tmem_d2, alo + 4, kHi, blo + 4, kHi, idesc, 1u);
                                if (nkk > 3) umma_f16_w(tmem_d2, alo + 6, kHi, blo + 6, kHi, idesc, 1u);
                                umma_commit(&empty_bar[stage]);
                            }
                            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                        }
                    }
                    if (elect_one()) {
                        umma_commit(&d2_full);
                        umma_commit(&y_empty);
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ============================ epilogue ============================
        pdl_wait();                                        // shortcut loads and output stores touch the chain's buffers
        const int q = warp % 4;
        const int row = q * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        uint32_t tph = 0, rph = 0;
        // TMA-store path (fused blocks): the second epilogue writes the output tile over Y (same swizzled [chunk][row]
        // layout, each warp its own 32 rows) and sends it as 4 KB sub-boxes; an identity shortcut arrives the same way
        const bool ts = fused && p.tma_store, tr = ts && p.tma_res;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, tph ^= 1u) {
            const int img = tile / p.tiles_per_img, tt = tile % p.tiles_per_img;
            const int bx = (tt % p.tiles_x) * p.tw, by = (tt / p.tiles_x) * p.th + (q * 32) / p.tw;   // this warp's sub-box
            const int oy = (tt / p.tiles_x) * p.th + row / p.tw, ox = (tt % p.tiles_x) * p.tw + row % p.tw;
            const bool valid = oy < p.OH && ox < p.OW;
            const size_t gpix = ((size_t)img * p.OH + oy) * p.OW + ox;
            const size_t dpix = ((size_t)img * p.dst_H + (oy * p.dst_stride + p.dst_off_y)) * p.dst_W +
                                (ox * p.dst_stride + p.dst_off_x);
            mbar_wait(&d1_full, tph);
            tc_fence_after();
            if (fused) {
                mbar_wait(&y_empty, tph ^ 1u);             // GEMM 2 of the previous tile has finished reading Y
                if (ts) { if (lane == 0) tma_store_wait_read<0>(); __syncwarp(); }   // and its output has left the buffer
                uint8_t* yrow = s_y + row * 128;
#pragma unroll 1
                for (int c0 = 0; c0 < p.n_mma; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_d1 + lane_off + (uint32_t)c0, r);
                    tmem_ld_wait();
                    uint8_t* ychunk = yrow + (c0 >> 6) * kBlkABytes;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = fmaxf(__uint_as_float(r[j * 8 + e]) + s_bias1[c0 + j * 8 + e], 0.f);
                        const int cj = ((c0 & 63) >> 3) + j;
                        *reinterpret_cast<uint4*>(ychunk + ((cj ^ (row & 7)) << 4)) =
                            make_uint4(pack2<T>(v[0], v[1]), pack2<T>(v[2], v[3]), pack2<T>(v[4], v[5]), pack2<T>(v[6], v[7]));
                    }
                    if ((c0 & 32) || c0 + 32 >= p.n_mma) {            // a 64-channel chunk of Y is complete: GEMM 2 may read it
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&y_full[c0 >> 6]);
                    }
                }
                // chunks of Y beyond n_mma (never written: all-zero weights there) still have to flip for the issuer
                for (int c = (p.n_mma + 63) >> 6; c < p.y_chunks; ++c) { __syncwarp(); if (lane == 0) mbar_arrive(&y_full[c]); }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d1_empty);
                mbar_wait(&d2_full, tph);
                tc_fence_after();
                if (tr) {
                    if (lane == 0) {
                        const int nb = (p.n_mma + 63) / 64;
                        mbar_expect_tx(&res_full[q], (uint32_t)(nb * 4096));
                        for (int k = 0; k < nb; ++k) tma_load_4d(s_y + k * kBlkABytes + q * 4096, &p.tmR, &res_full[q], k * 64, bx, by, img);
                    }
                    mbar_wait(&res_full[q], rph);
                    rph ^= 1u;
                }
            }
            const uint32_t tmem_out = fused ? tmem_d2 : tmem_d1;
            const float* sb = fused ? s_bias2 : s_bias1;
#pragma unroll 1
            for (int c0 = 0; c0 < p.n_mma; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_out + lane_off + (uint32_t)c0, r);
                tmem_ld_wait();
                if (valid || ts) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + sb[c0 + i];
                    const uint32_t orow = smem_u32(s_y) + (uint32_t)((c0 >> 6) * kBlkABytes + row * 128);
                    const uint32_t half = (uint32_t)(c0 & 63) >> 3, sw = (uint32_t)(row & 7);
                    if (p.residual) {
                        const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const T*>(p.residual) + gpix * p.res_C + c0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 u = tr ? ld_shared_v4(orow + (((half + (uint32_t)j) ^ sw) << 4)) : __ldg(rp + j);
                            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = unpack2<T>(w[e]);
                                v[j * 8 + e * 2] += f.x;
                                v[j * 8 + e * 2 + 1] += f.y;
                            }
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    if (ts) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            st_shared_v4(orow + (((half + (uint32_t)j) ^ sw) << 4), pack2<T>(v[j * 8], v[j * 8 + 1]), pack2<T>(v[j * 8 + 2], v[j * 8 + 3]),
                                         pack2<T>(v[j * 8 + 4], v[j * 8 + 5]), pack2<T>(v[j * 8 + 6], v[j * 8 + 7]));
                        if ((c0 & 32) || c0 + 32 >= p.n_mma) {
                            fence_async_smem();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_4d(&p.tmD, smem_u32(s_y) + (uint32_t)((c0 >> 6) * kBlkABytes + q * 4096), (c0 >> 6) * 64, bx, by, img);
                                tma_store_commit();
                            }
                        }
                    } else if (p.dst_fp32) {
                        float4* dp = reinterpret_cast<float4*>(static_cast<float*>(p.dst) + dpix * p.dst_C + c0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) dp[j] = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
                    } else {
                        uint4* dp = reinterpret_cast<uint4*>(static_cast<T*>(p.dst) + dpix * p.dst_C + c0);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            dp[j] = make_uint4(pack2<T>(v[j * 8], v[j * 8 + 1]), pack2<T>(v[j * 8 + 2], v[j * 8 + 3]),
                                               pack2<T>(v[j * 8 + 4], v[j * 8 + 5]), pack2<T>(v[j * 8 + 6], v[j * 8 + 7]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(fused ? &d2_empty : &d1_empty);
        }
        if (ts && lane == 0) tma_store_wait_all();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_slot, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// Host
// ------------------------------------------------------------------------------------------------
struct TcBlockPlan {
    BlkParams params;
    int block_n, operand_type, grid;
};

template <int BLOCK_N, int STAGES, typename T>
static void launch_block_t(const TcBlockPlan* plan, cudaStream_t st) {
    auto kern = block_tc_kernel<BLOCK_N, STAGES, T>;
    const size_t smem = (size_t)STAGES * (kBlkABytes + BLOCK_N * 128) + (size_t)(BLOCK_N / 64) * kBlkABytes + 1024;
    SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_pdl(kern, dim3(plan->grid), dim3(kBlkThreads), smem, st, plan->params);
}

template <typename T>
static void launch_block_n(const TcBlockPlan* plan, cudaStream_t st) {
    switch (plan->block_n) {
        case 64: launch_block_t<64, 3, T>(plan, st); break;      // 88 KB: two CTAs per SM
        case 128: launch_block_t<128, 4, T>(plan, st); break;    // 160 KB
        case 256: launch_block_t<256, 3, T>(plan, st); break;    // 208 KB
        default: throw std::invalid_argument("tcgen05 block: unsupported channel count");
    }
}

void launch_block_tc(const TcBlockPlan* plan, cudaStream_t st) {
    if (!plan) throw std::runtime_error("tcgen05 block: no plan");
    if (plan->operand_type == PREC_FP16) launch_block_n<__half>(plan, st);
    else launch_block_n<__nv_bfloat16>(plan, st);
}

void tc_block_plan_destroy(TcBlockPlan* plan) { delete plan; }

// c1: the 3x3 convolution (or a stand-alone convolution when c2 == nullptr); c2: the 1x1 convolution whose
// first segment is Y = output of c1 (never materialised) and whose other segments are the 1x1 shortcut over
// the sources of c1, in the same order.  real_cout: channels that are not padding (65 for the detector).
TcBlockPlan* tc_block_plan_create(const ConvDev& c1, const ConvDev* c2, int operand_type, int real_cout, int num_sms) {
    if (operand_type != PREC_FP16 && operand_type != PREC_BF16) throw std::invalid_argument("tcgen05 block: bad operand type");
    auto plan = std::make_unique<TcBlockPlan>();
    BlkParams& p = plan->params;
    std::memset(&p, 0, sizeof(p));
    const CUtensorMapDataType dt = operand_type == PREC_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const int N = c1.cout_pad;
    if (N != 64 && N != 128 && N != 256) throw std::invalid_argument("tcgen05 block: cout_pad must be 64, 128 or 256");
    const ConvDev& last = c2 ? *c2 : c1;
    if (last.dst_C % 8 != 0 || (last.residual && last.res_C % 8 != 0)) throw std::invalid_argument("tcgen05 block: channel strides must be multiples of 8");
    if (c2 && (c2->cout_pad != N || c2->OH != c1.OH || c2->OW != c1.OW || c2->dst_stride != 1))
        throw std::invalid_argument("tcgen05 block: the two convolutions do not chain");

    int best_th = 8, best_tw = 16;
    long best = -1;
    for (int tw = 128; tw >= 1; tw >>= 1) {
        const int th = 128 / tw;
        const long n = (long)((c1.OH + th - 1) / th) * ((c1.OW + tw - 1) / tw);
        if (best < 0 || n < best) { best = n; best_th = th; best_tw = tw; }
    }
    p.th = best_th; p.tw = best_tw;
    p.tiles_x = (c1.OW + p.tw - 1) / p.tw;
    p.tiles_per_img = p.tiles_x * ((c1.OH + p.th - 1) / p.th);
    p.total_tiles = p.tiles_per_img * c1.B;
    const int per_sm = N == 64 ? 2 : 1;
    plan->grid = std::min(p.total_tiles, num_sms * per_sm);
    plan->block_n = N;
    plan->operand_type = operand_type;

    // MMA N: real channels rounded up to 32 (epilogue granularity), K steps of the last chunk from the real channels
    p.n_mma = std::min(N, (real_cout + 31) / 32 * 32);
    auto kk_of = [](int real_c, int nchunks) {
        const int in_last = real_c - (nchunks - 1) * 64;
        return std::max(1, std::min(4, (in_last + 15) / 16));
    };

    p.nseg = c1.nseg;
    int ksteps = 0;
    for (int s = 0; s < c1.nseg; ++s) {
        const SegDev& sg = c1.seg[s];
        if (sg.cin % 64 != 0 || sg.C % 64 != 0) throw std::invalid_argument("tcgen05 block: channels must be multiples of 64");
        if (sg.koff != ksteps * 64) throw std::invalid_argument("tcgen05 block: unexpected K offset");
        BlkSeg& ts = p.seg[s];
        ts.ntaps = sg.ntaps; ts.nchunks = sg.cin / 64; ts.stride = sg.stride; ts.C = sg.C;
        ts.kk_last = kk_of(sg.cin_real > 0 ? sg.cin_real : sg.cin, ts.nchunks);
        for (int t = 0; t < sg.ntaps; ++t) { ts.dy[t] = sg.dy[t]; ts.dx[t] = sg.dx[t]; }
        ksteps += sg.ntaps * ts.nchunks;
        const cuuint64_t C = sg.C, W = sg.W, H = sg.H;
        cuuint32_t box[5] = {64, (cuuint32_t)p.tw, 1, (cuuint32_t)p.th, 1};
        if (sg.stride == 1) {
            cuuint64_t dims[5] = {C, W, 1, H, (cuuint64_t)c1.B};
            cuuint64_t str[4] = {C * 2, W * C * 2, W * C * 2, H * W * C * 2};
            tc_encode_tiled(&p.tmA[s], dt, 5, sg.src, dims, str, box);
        } else if (sg.stride == 2) {
            if (W % 2 || H % 2) throw std::invalid_argument("tcgen05 block: stride-2 source must have even height and width");
            cuuint64_t dims[5] = {2 * C, W / 2, 2, H / 2, (cuuint64_t)c1.B};
            cuuint64_t str[4] = {2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
            tc_encode_tiled(&p.tmA[s], dt, 5, sg.src, dims, str, box);
        } else {
            throw std::invalid_argument("tcgen05 block: stride must be 1 or 2");
        }
    }
    int ktail = 0;                                            // packed K tails follow the segments (common.cuh, SegDev): unused here
    for (int s = 0; s < c1.nseg; ++s) ktail += c1.seg[s].koff_tail > 0 ? 3 : 0;
    if ((ksteps + ktail) * 64 != c1.K) throw std::invalid_argument("tcgen05 block: K does not match the segments");
    {
        cuuint64_t dims[2] = {(cuuint64_t)c1.K, (cuuint64_t)N};
        cuuint64_t str[1] = {(cuuint64_t)c1.K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)N};
        tc_encode_tiled(&p.tmW1, dt, 2, c1.w, dims, str, box);
    }
    p.bias1 = c1.bias;
    if (c2) {
        // segment 0 of c2 is Y; the others must be 1x1 centre taps over c1's sources, in order
        if (c2->nseg < 1 || c2->seg[0].ntaps != 1 || c2->seg[0].cin != N) throw std::invalid_argument("tcgen05 block: bad 1x1 convolution");
        p.y_chunks = N / 64;
        p.y_kk_last = kk_of(real_cout, p.y_chunks);
        // chunks of Y beyond the real channels are all-zero: GEMM 2 still walks them (weights are zero there)
        p.nds = c2->nseg - 1;
        if (p.nds != 0 && p.nds != c1.nseg) throw std::invalid_argument("tcgen05 block: shortcut segments do not match the sources");
        int k2 = p.y_chunks;
        for (int s = 0; s < p.nds; ++s) {
            const SegDev& sg = c2->seg[s + 1];
            if (sg.src != c1.seg[s].src || sg.ntaps != 1 || sg.dy[0] != 0 || sg.dx[0] != 0 || sg.stride != c1.seg[s].stride ||
                sg.cin != c1.seg[s].cin || sg.koff != k2 * 64)
                throw std::invalid_argument("tcgen05 block: shortcut segment mismatch");
            k2 += sg.cin / 64;
        }
        if (k2 * 64 != c2->K) throw std::invalid_argument("tcgen05 block: K of the 1x1 convolution does not match");
        p.ksteps2 = k2;
        cuuint64_t dims[2] = {(cuuint64_t)c2->K, (cuuint64_t)N};
        cuuint64_t str[1] = {(cuuint64_t)c2->K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)N};
        tc_encode_tiled(&p.tmW2, dt, 2, c2->w, dims, str, box);
        p.bias2 = c2->bias;
    }
    p.OH = c1.OH; p.OW = c1.OW;
    p.residual = last.residual; p.dst = last.dst;
    p.res_C = last.res_C; p.dst_H = last.dst_H; p.dst_W = last.dst_W; p.dst_C = last.dst_C;
    p.dst_stride = last.dst_stride; p.dst_off_y = last.dst_off_y; p.dst_off_x = last.dst_off_x;
    p.relu = last.relu; p.dst_fp32 = last.dst_fp32;
    const int n64 = (p.n_mma + 63) / 64 * 64;
    if (c2 && !last.dst_fp32 && last.dst_stride == 1 && last.dst_off_y == 0 && last.dst_off_x == 0 && last.dst_H == c1.OH &&
        last.dst_W == c1.OW && p.tw <= 32 && last.dst_C % 64 == 0 && last.dst_C >= n64 && n64 <= N && !std::getenv("SPB200_NO_TMA_STORE")) {
        cuuint32_t box[4] = {64, (cuuint32_t)p.tw, (cuuint32_t)(32 / p.tw), 1};
        auto encode = [&](CUtensorMap* m, const void* base, int C) {
            cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)c1.OW, (cuuint64_t)c1.OH, (cuuint64_t)c1.B};
            cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)c1.OW * C * 2, (cuuint64_t)c1.OH * c1.OW * C * 2};
            tc_encode_tiled(m, dt, 4, base, dims, str, box);
        };
        encode(&p.tmD, last.dst, last.dst_C);
        p.tma_store = 1;
        if (last.residual && last.res_C % 64 == 0 && last.res_C >= n64 && !std::getenv("SPB200_NO_TMA_RES")) {
            encode(&p.tmR, last.residual, last.res_C);
            p.tma_res = 1;
        }
    }
    return plan.release();
}

}  // namespace spb200
