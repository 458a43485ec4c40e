// Stem on tensor cores: conv7x7/s2/p3 (+folded BatchNorm) + ReLU + maxpool3x3/s2/p1 in ONE kernel
// (reference python/src/superpoint.py:12-15,20-23).
//
// A CTA produces a 5x10 tile of pooled pixels x 64 channels.  That needs the 11x21 convolution
// outputs around it (pool windows overlap), i.e. 231 GEMM rows = two M=128 tcgen05 tiles:
//   1. the 27x47 input patch is loaded once, multiplied by 255 and rounded to the 16-bit operand type
//      (double buffered: the next tile's patch is fetched while the tensor core works),
//   2. every thread builds one im2col row straight into the 128B-swizzled K-major smem layout the UMMA
//      descriptor expects (no im2col in HBM).  K is ordered (channel, ky, kx) with kx padded 7 -> 8, so
//      the eight taps of one filter row are eight CONSECUTIVE patch pixels = one 16-byte chunk of the
//      row: a row costs 7 x (4 LDS.32 + 1 STS.128) per channel instead of 49 scalar loads + converts,
//   3. one thread issues the tcgen05.mma chain (weights arrive by TMA), accumulators in TMEM.  Precision:
//      the image is multiplied by 255 before the 16-bit rounding, so 8-bit images (value = k/255, what
//      cameras and the reference's loaders produce) are represented exactly; the weights are split into
//      16-bit hi + lo parts and two MMAs (a*w_hi + a*w_lo) are accumulated; 1/255 is applied to the fp32
//      accumulator.  The stem is then as accurate as fp32 for 8-bit inputs at negligible tensor cost,
//   4. the epilogue scales, adds the bias and stages the 231x64 tile in smem transposed (channel quad major),
//      rows outside the conv output as 0,
//   5. the 3x3/s2 max-pool runs per (pooled pixel, channel quad) with packed 16-bit maxima; its running maximum
//      starts at 0, which is the ReLU (max and ReLU commute); NHWC 16-bit out.
// HBM traffic is the algorithmic minimum: the fp32 image in, the pooled tensor out.
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"

namespace spb200 {

constexpr int kStPH = 5, kStPW = 10;                  // pooled tile
constexpr int kStCH = 2 * kStPH + 1, kStCW = 2 * kStPW + 1;   // conv region 11 x 21
constexpr int kStRows = kStCH * kStCW;                // 231 GEMM rows (<= 256)
constexpr int kStIH = 2 * kStCH + 5, kStIW = 2 * kStCW + 5;   // input patch 27 x 47
constexpr int kStIWp = 48;                            // patch row pitch in smem (elements): 8 taps from column 2*20 stay inside
constexpr int kStThreads = 256;

struct StemParams {
    CUtensorMap tmW;          // [64 cout][NCHUNK*64] K-major 16-bit
    const float* img;         // NCHW fp32
    const float* bias;        // [64]
    void* dst;                // NHWC [B][H/4][W/4][64] 16-bit
    int H, W;
    int tiles_x, tiles_per_img, total_tiles;
};

template <int CIN, bool SPLIT, typename T>
__global__ void __launch_bounds__(kStThreads, CIN == 1 ? 4 : 1) stem_tc_kernel(const __grid_constant__ StemParams p) {
    constexpr int NCHUNK = CIN;                                       // one 64-wide K chunk per channel: k = ky * 8 + kx
    constexpr int NPART = SPLIT ? 2 : 1;                              // weight parts: hi (+ lo)
    constexpr int kATile = 128 * 128;                                 // bytes of one M-tile of one K chunk of one part
    constexpr uint32_t kIdesc = (1u << 4) | (OperandFmt<T>::value << 7) | (OperandFmt<T>::value << 10) |
                                ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    extern __shared__ uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t bar_w, bar_mma;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_bias[64];

    // aligned up to 1024 B by pointer ARITHMETIC on dyn_smem: the compiler keeps the shared address space (LDS/STS,
    // 32-bit addresses) instead of falling back to generic loads
    uint8_t* base = dyn_smem + ((1024u - (smem_u32(dyn_smem) & 1023u)) & 1023u);
    uint8_t* s_a = base;                                              // [2 M-tiles][NCHUNK][128 rows][128 B]
    uint8_t* s_w = s_a + 2 * NCHUNK * kATile;                 // [NPART][NCHUNK][64 rows][128 B]
    T* s_in = reinterpret_cast<T*>(s_w + NPART * NCHUNK * 64 * 128);  // [2 buffers][CIN][27][48], image * 255 in the operand type
    constexpr int kPatch = CIN * kStIH * kStIWp;
    uint8_t* s_conv = s_a;                                            // aliases A after the MMAs: [16 channel quads][231 rows][8 B]

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid / 32, 0), lane = tid % 32;
    const int H = p.H, W = p.W, CH = H / 2, CW = W / 2, PH = H / 4, PW = W / 4;

    if (tid == 0) {
        mbar_init(&bar_w, 1);
        mbar_init(&bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        prefetch_tmap(&p.tmW);
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 128);
    if (tid < 64) s_bias[tid] = p.bias[tid];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_slot;

    if (tid == 0) {
        mbar_expect_tx(&bar_w, NPART * NCHUNK * 64 * 128);
        for (int c = 0; c < NPART * NCHUNK; ++c) tma_load_2d(s_w + c * 64 * 128, &p.tmW, &bar_w, c * 64, 0);
    }
    // persistent: this CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the input patch of the next
    // tile is fetched while the tensor core works on the current one
    // patch loader: the (channel, row, pixel pair) a thread fetches in each of its iterations does not depend on the
    // tile, so the decomposition is done once; per tile only the origin changes
    constexpr int kPairs = CIN * kStIH * (kStIWp / 2);
    constexpr int kPatchIters = (kPairs + kStThreads - 1) / kStThreads;
    int pl_dy[kPatchIters], pl_dx[kPatchIters], pl_c[kPatchIters], pl_off[kPatchIters];
#pragma unroll
    for (int it = 0; it < kPatchIters; ++it) {
        const int i = tid + it * kStThreads;
        const int r = i % (kStIH * (kStIWp / 2));
        pl_c[it] = i / (kStIH * (kStIWp / 2));
        pl_dy[it] = r / (kStIWp / 2);
        pl_dx[it] = 2 * (r % (kStIWp / 2));
        pl_off[it] = (pl_c[it] * p.H + pl_dy[it]) * p.W + pl_dx[it];          // offset from the patch origin (interior tiles)
    }
    auto load_patch = [&](int tile, int buf) {
        const int b = tile / p.tiles_per_img, tt = tile % p.tiles_per_img;
        const int iy0 = 4 * ((tt / p.tiles_x) * kStPH) - 5, ix0 = 4 * ((tt % p.tiles_x) * kStPW) - 5;
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_in + buf * kPatch);
        const float* img_b = p.img + (size_t)b * CIN * H * W;
        if (iy0 >= 0 && iy0 + kStIH <= H && ix0 >= 0 && ix0 + kStIWp <= W) {       // patch entirely inside the image
            const float* org = img_b + (size_t)iy0 * W + ix0;
#pragma unroll
            for (int it = 0; it < kPatchIters; ++it) {
                const int i = tid + it * kStThreads;
                if (i < kPairs) {
                    const float* q = org + pl_off[it];
                    dst[i] = pack2<T>(__ldg(q) * 255.f, __ldg(q + 1) * 255.f);
                }
            }
        } else {
#pragma unroll
            for (int it = 0; it < kPatchIters; ++it) {
                const int i = tid + it * kStThreads;
                if (i < kPairs) {
                    const int y = iy0 + pl_dy[it], x = ix0 + pl_dx[it];
                    float v0 = 0.f, v1 = 0.f;
                    if (y >= 0 && y < H) {
                        const float* row = img_b + ((size_t)pl_c[it] * H + y) * W;
                        if (x >= 0 && x < W) v0 = __ldg(row + x);
                        if (x + 1 >= 0 && x + 1 < W) v1 = __ldg(row + x + 1);
                    }
                    dst[i] = pack2<T>(v0 * 255.f, v1 * 255.f);
                }
            }
        }
    };
    if ((int)blockIdx.x < p.total_tiles) load_patch(blockIdx.x, 0);
    __syncthreads();
    int buf = 0;
    uint32_t mma_phase = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, mma_phase ^= 1u, first = false, buf ^= 1) {
        const int b = tile / p.tiles_per_img, tt = tile % p.tiles_per_img;
        const int py0 = (tt / p.tiles_x) * kStPH, px0 = (tt % p.tiles_x) * kStPW;
        const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;               // conv coords of the region origin
        // ---- im2col row of this thread, written in the SWIZZLE_128B K-major layout -------------------
        {
            const int row = tid;                                   // 0..255, rows >= 231 are zero rows
            const int mt = row >> 7, rr = row & 127;
            const bool live = row < kStRows;
            const int cyl = row / kStCW, cxl = row % kStCW;
            const T* pin = s_in + buf * kPatch + (2 * cyl) * kStIWp + 2 * cxl;    // 4-byte aligned
            const int sw = (rr & 7) << 4;                          // the row's swizzle phase
    #pragma unroll
            for (int ck = 0; ck < NCHUNK; ++ck) {                  // chunk = input channel
                uint8_t* arow = s_a + (mt * NCHUNK + ck) * kATile + rr * 128;
                const T* pc = pin + ck * kStIH * kStIWp;
    #pragma unroll
                for (int j = 0; j < 8; ++j) {                      // 16-byte piece j = filter row ky = j: patch pixels 2*cxl .. 2*cxl+7
                    uint4 u = make_uint4(0u, 0u, 0u, 0u);          // (a warp writes piece j of 32 rows: the XOR spreads them over all banks)
                    if (live && j < 7) {
                        const uint32_t* q = reinterpret_cast<const uint32_t*>(pc + j * kStIWp);
                        u = make_uint4(q[0], q[1], q[2], q[3]);
                    }
                    *reinterpret_cast<uint4*>(arow + ((j << 4) ^ sw)) = u;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> tensor-core reads
        __syncthreads();

        if (warp == 0) {
            if (first) mbar_wait(&bar_w, 0);
            tc_fence_after();
            const uint32_t a_lo = umma_desc_lo(smem_u32(s_a)), w_lo = umma_desc_lo(smem_u32(s_w));
            constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
            if (elect_one()) {
    #pragma unroll
                for (int mt = 0; mt < 2; ++mt)
    #pragma unroll
                    for (int ck = 0; ck < NCHUNK; ++ck)
    #pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint32_t a = a_lo + (uint32_t)(((mt * NCHUNK + ck) * kATile + kk * 32) >> 4);
                            const uint32_t w = w_lo + (uint32_t)((ck * 64 * 128 + kk * 32) >> 4);
                            umma_f16_w(tmem_acc + mt * 64, a, kHi, w, kHi, kIdesc, (ck > 0 || kk > 0) ? 1u : 0u);
                            if (SPLIT) umma_f16_w(tmem_acc + mt * 64, a, kHi, w + (uint32_t)((NCHUNK * 64 * 128) >> 4), kHi, kIdesc, 1u);
                        }
                umma_commit(&bar_mma);
            }
        }
        __syncwarp();

        if (tile + (int)gridDim.x < p.total_tiles) load_patch(tile + gridDim.x, buf ^ 1);   // the other patch buffer
        mbar_wait(&bar_mma, mma_phase);
        tc_fence_after();

        // ---- epilogue: warps 0-3 own M-tile 0 (TMEM columns 0-63), warps 4-7 own M-tile 1 (64-127) ----
        {
            const int mt = warp >> 2, q = warp & 3;
            const int row = mt * 128 + q * 32 + lane;
            const int cy = cy0 + row / kStCW, cx = cx0 + row % kStCW;
            const bool real = row < kStRows && cy >= 0 && cy < CH && cx >= 0 && cx < CW;
            // staged transposed: [channel quad 0..15][conv row 0..230] 8-byte entries, so that the pool addresses a
            // fixed quad with immediate row offsets and neither side has bank conflicts.  No ReLU here: the pool's
            // running maximum starts at 0, which is the ReLU; rows outside the conv output store 0 (neutral).
            uint8_t* cq = s_conv + row * 8;
    #pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * 64 + half * 32), r);
                float bb[32];
    #pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float4 f = *reinterpret_cast<const float4*>(s_bias + half * 32 + 4 * e);
                    bb[4 * e] = f.x; bb[4 * e + 1] = f.y; bb[4 * e + 2] = f.z; bb[4 * e + 3] = f.w;
                }
                tmem_ld_wait();
                if (row < kStRows) {
    #pragma unroll
                    for (int j = 0; j < 8; ++j) {                  // channel quad half * 8 + j
                        uint2 u = make_uint2(0u, 0u);
                        if (real) {
                            u.x = pack2<T>(fmaf(__uint_as_float(r[4 * j]), 1.f / 255.f, bb[4 * j]),
                                           fmaf(__uint_as_float(r[4 * j + 1]), 1.f / 255.f, bb[4 * j + 1]));
                            u.y = pack2<T>(fmaf(__uint_as_float(r[4 * j + 2]), 1.f / 255.f, bb[4 * j + 2]),
                                           fmaf(__uint_as_float(r[4 * j + 3]), 1.f / 255.f, bb[4 * j + 3]));
                        }
                        *reinterpret_cast<uint2*>(cq + (half * 8 + j) * (kStRows * 8)) = u;
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();

        // ---- 3x3 / stride-2 max-pool over the staged conv tile, two channels per thread --------------
        T* out = static_cast<T*>(p.dst);
        for (int o = tid; o < kStPH * kStPW * 16; o += kStThreads) {
            // one work item = one pooled pixel x four channels; the nine taps sit at immediate offsets from one base
            const int c4 = o & 15, pp = o >> 4;
            const int ppy = pp / kStPW, ppx = pp - ppy * kStPW;
            const int py = py0 + ppy, px = px0 + ppx;
            if (py >= PH || px >= PW) continue;
            const uint8_t* src = s_conv + c4 * (kStRows * 8) + ((2 * ppy) * kStCW + 2 * ppx) * 8;
            uint2 m = make_uint2(0u, 0u);                          // max with 0 = the ReLU
    #pragma unroll
            for (int dy = 0; dy < 3; ++dy)
    #pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const uint2 u = *reinterpret_cast<const uint2*>(src + (dy * kStCW + dx) * 8);
                    m.x = max2<T>(m.x, u.x);
                    m.y = max2<T>(m.y, u.y);
                }
            *reinterpret_cast<uint2*>(out + ((size_t)(b * PH + py) * PW + px) * 64 + c4 * 4) = m;
        }
        __syncthreads();           // pooling reads s_conv (aliases the A tiles) and s_in is rewritten: next build may start
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_acc, 128);
    }
}

// ------------------------------------------------------------------------------------------------
// Split-precision stem (the first stage of the engine's split levels): the same tile scheme with
//   * the image x 255 kept as hi + lo 16-bit parts (an fp32 image in [0, 1] is then represented to 2^-22),
//   * three MMAs per product, a.hi w.hi + a.hi w.lo + a.lo w.hi, one input channel (= one 64-wide K chunk) at a time,
//   * the convolution tile staged and pooled in fp32,
//   * the pooled tensor stored in the split layout of common.cuh (SegDev): per 32 channels [hi 32 | lo 32].
// ------------------------------------------------------------------------------------------------
template <int CIN, typename T>
__global__ void __launch_bounds__(kStThreads, 1) stem_wide_kernel(const __grid_constant__ StemParams p) {
    constexpr int kATile = 128 * 128;                                 // bytes of one M-tile of one part
    constexpr uint32_t kIdesc = (1u << 4) | (OperandFmt<T>::value << 7) | (OperandFmt<T>::value << 10) |
                                ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    extern __shared__ uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t bar_w, bar_mma;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_bias[64];

    uint8_t* base = dyn_smem + ((1024u - (smem_u32(dyn_smem) & 1023u)) & 1023u);
    uint8_t* s_a = base;                                              // [2 M-tiles][hi, lo][128 rows][128 B]
    uint8_t* s_w = s_a + 4 * kATile;                                  // [hi, lo][CIN][64 rows][128 B]
    constexpr int kPatch = CIN * kStIH * kStIWp;
    T* s_in = reinterpret_cast<T*>(s_w + 2 * CIN * 64 * 128);         // [2 buffers][hi, lo][CIN][27][48], image * 255
    float* s_conv = reinterpret_cast<float*>(s_a);                    // aliases A after the MMAs: [16 channel quads][231 rows][4] fp32

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid / 32, 0), lane = tid % 32;
    const int H = p.H, W = p.W, CH = H / 2, CW = W / 2, PH = H / 4, PW = W / 4;

    if (tid == 0) {
        mbar_init(&bar_w, 1);
        mbar_init(&bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        prefetch_tmap(&p.tmW);
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 128);
    if (tid < 64) s_bias[tid] = p.bias[tid];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_slot;

    if (tid == 0) {
        mbar_expect_tx(&bar_w, 2 * CIN * 64 * 128);
        for (int c = 0; c < 2 * CIN; ++c) tma_load_2d(s_w + c * 64 * 128, &p.tmW, &bar_w, c * 64, 0);
    }
    // ---- input patch of a tile as hi + lo parts, two pixels per thread and step; every load of the thread is issued
    // before the first conversion.  The patch of the next tile is fetched while the tensor core works on this one.
    constexpr int kPairs = CIN * kStIH * (kStIWp / 2);
    constexpr int kPatchIters = (kPairs + kStThreads - 1) / kStThreads;
    auto load_patch = [&](int tile, int buf) {
        const int b = tile / p.tiles_per_img, tt = tile % p.tiles_per_img;
        const int iy0 = 4 * ((tt / p.tiles_x) * kStPH) - 5, ix0 = 4 * ((tt % p.tiles_x) * kStPW) - 5;
        const float* img_b = p.img + (size_t)b * CIN * H * W;
        uint32_t* dst_hi = reinterpret_cast<uint32_t*>(s_in + buf * 2 * kPatch);
        uint32_t* dst_lo = reinterpret_cast<uint32_t*>(s_in + buf * 2 * kPatch + kPatch);
        float v0[kPatchIters], v1[kPatchIters];
#pragma unroll
        for (int it = 0; it < kPatchIters; ++it) {
            const int i = tid + it * kStThreads;
            v0[it] = v1[it] = 0.f;
            if (i < kPairs) {
                const int c = i / (kStIH * (kStIWp / 2)), r = i % (kStIH * (kStIWp / 2));
                const int y = iy0 + r / (kStIWp / 2), x = ix0 + 2 * (r % (kStIWp / 2));
                if (y >= 0 && y < H) {
                    const float* row = img_b + ((size_t)c * H + y) * W;
                    if (x >= 0 && x < W) v0[it] = __ldg(row + x);
                    if (x + 1 >= 0 && x + 1 < W) v1[it] = __ldg(row + x + 1);
                }
            }
        }
#pragma unroll
        for (int it = 0; it < kPatchIters; ++it) {
            const int i = tid + it * kStThreads;
            if (i < kPairs) {
                const float a = v0[it] * 255.f, c = v1[it] * 255.f;
                const uint32_t hw = pack2<T>(a, c);
                const float2 f = unpack2<T>(hw);
                dst_hi[i] = hw;
                dst_lo[i] = pack2<T>(a - f.x, c - f.y);
            }
        }
    };
    uint32_t mma_phase = 0;
    bool first = true;
    int buf = 0;
    if ((int)blockIdx.x < p.total_tiles) load_patch(blockIdx.x, 0);
    __syncthreads();
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, buf ^= 1) {
        const int b = tile / p.tiles_per_img, tt = tile % p.tiles_per_img;
        const int py0 = (tt / p.tiles_x) * kStPH, px0 = (tt % p.tiles_x) * kStPW;
        const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;
        for (int ck = 0; ck < CIN; ++ck) {
            // ---- im2col rows of channel ck (hi and lo), SWIZZLE_128B K-major, as in stem_tc_kernel ----
            {
                const int row = tid;
                const int mt = row >> 7, rr = row & 127;
                const bool live = row < kStRows;
                const int cyl = row / kStCW, cxl = row % kStCW;
                const int sw = (rr & 7) << 4;
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    const T* pin = s_in + (buf * 2 + part) * kPatch + ck * kStIH * kStIWp + (2 * cyl) * kStIWp + 2 * cxl;
                    uint8_t* arow = s_a + (mt * 2 + part) * kATile + rr * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        uint4 u = make_uint4(0u, 0u, 0u, 0u);
                        if (live && j < 7) {
                            const uint32_t* q = reinterpret_cast<const uint32_t*>(pin + j * kStIWp);
                            u = make_uint4(q[0], q[1], q[2], q[3]);
                        }
                        *reinterpret_cast<uint4*>(arow + ((j << 4) ^ sw)) = u;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (warp == 0) {
                if (first) mbar_wait(&bar_w, 0);
                tc_fence_after();
                const uint32_t a_lo = umma_desc_lo(smem_u32(s_a)), w_lo = umma_desc_lo(smem_u32(s_w));
                constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
                if (elect_one()) {
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint32_t ah = a_lo + (uint32_t)(((mt * 2) * kATile + kk * 32) >> 4);
                            const uint32_t al = a_lo + (uint32_t)(((mt * 2 + 1) * kATile + kk * 32) >> 4);
                            const uint32_t wh = w_lo + (uint32_t)((ck * 64 * 128 + kk * 32) >> 4);
                            const uint32_t wl = wh + (uint32_t)((CIN * 64 * 128) >> 4);
                            umma_f16_w(tmem_acc + mt * 64, ah, kHi, wh, kHi, kIdesc, (ck > 0 || kk > 0) ? 1u : 0u);
                            umma_f16_w(tmem_acc + mt * 64, ah, kHi, wl, kHi, kIdesc, 1u);
                            umma_f16_w(tmem_acc + mt * 64, al, kHi, wh, kHi, kIdesc, 1u);
                        }
                    umma_commit(&bar_mma);
                }
            }
            __syncwarp();
            first = false;
            if (ck == 0 && tile + (int)gridDim.x < p.total_tiles) load_patch(tile + gridDim.x, buf ^ 1);   // the other patch buffer
            mbar_wait(&bar_mma, mma_phase);                        // the A tiles may be rebuilt (or aliased by the staging)
            mma_phase ^= 1u;
            tc_fence_after();
        }
        // ---- epilogue: fp32 conv tile, transposed [channel quad][row] ----
        {
            const int mt = warp >> 2, q = warp & 3;
            const int row = mt * 128 + q * 32 + lane;
            const int cy = cy0 + row / kStCW, cx = cx0 + row % kStCW;
            const bool real = row < kStRows && cy >= 0 && cy < CH && cx >= 0 && cx < CW;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * 64 + half * 32), r);
                tmem_ld_wait();
                if (row < kStRows) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (real) {
                            const float* bb = s_bias + half * 32 + 4 * j;
                            u = make_float4(fmaf(__uint_as_float(r[4 * j]), 1.f / 255.f, bb[0]), fmaf(__uint_as_float(r[4 * j + 1]), 1.f / 255.f, bb[1]),
                                            fmaf(__uint_as_float(r[4 * j + 2]), 1.f / 255.f, bb[2]), fmaf(__uint_as_float(r[4 * j + 3]), 1.f / 255.f, bb[3]));
                        }
                        *reinterpret_cast<float4*>(s_conv + ((size_t)(half * 8 + j) * kStRows + row) * 4) = u;
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        // ---- 3x3 / stride-2 max-pool in fp32 (running maximum from 0 = the ReLU), split-layout store ----
        uint16_t* out = static_cast<uint16_t*>(p.dst);
        for (int o = tid; o < kStPH * kStPW * 16; o += kStThreads) {
            const int c4 = o & 15, pp = o >> 4;
            const int ppy = pp / kStPW, ppx = pp - ppy * kStPW;
            const int py = py0 + ppy, px = px0 + ppx;
            if (py >= PH || px >= PW) continue;
            const float* src = s_conv + ((size_t)c4 * kStRows + (2 * ppy) * kStCW + 2 * ppx) * 4;
            float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float4 u = *reinterpret_cast<const float4*>(src + (dy * kStCW + dx) * 4);
                    m.x = fmaxf(m.x, u.x); m.y = fmaxf(m.y, u.y); m.z = fmaxf(m.z, u.z); m.w = fmaxf(m.w, u.w);
                }
            const uint32_t h0 = pack2<T>(m.x, m.y), h1 = pack2<T>(m.z, m.w);
            const float2 f0 = unpack2<T>(h0), f1 = unpack2<T>(h1);
            const uint32_t l0 = pack2<T>(m.x - f0.x, m.y - f0.y), l1 = pack2<T>(m.z - f1.x, m.w - f1.y);
            // channels 4 c4 .. 4 c4 + 3 of 64: chunk c4 / 8, hi at (c4 % 8) * 4, lo 32 values further
            uint16_t* dst = out + ((size_t)(b * PH + py) * PW + px) * 128 + (c4 >> 3) * 64 + (c4 & 7) * 4;
            *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(dst + 32) = make_uint2(l0, l1);
        }
        __syncthreads();
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_acc, 128);
    }
}

struct StemTcPlan {
    StemParams params;
    int cin, operand_type, num_sms;
};

template <int CIN, bool SPLIT, typename T>
static void launch_stem_tc_t(const StemTcPlan* plan, const float* img, void* dst, int B, int H, int W, cudaStream_t st) {
    constexpr int NCHUNK = CIN;
    constexpr int NPART = SPLIT ? 2 : 1;
    auto kern = stem_tc_kernel<CIN, SPLIT, T>;
    const size_t smem = (size_t)2 * NCHUNK * 128 * 128 + (size_t)NPART * NCHUNK * 64 * 128 + (size_t)2 * CIN * kStIH * kStIWp * 2 + 1024;
    SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StemParams p = plan->params;
    p.img = img; p.dst = dst; p.H = H; p.W = W;
    p.tiles_x = (W / 4 + kStPW - 1) / kStPW;
    p.tiles_per_img = p.tiles_x * ((H / 4 + kStPH - 1) / kStPH);
    p.total_tiles = p.tiles_per_img * B;
    const int per_sm = std::max(1, std::min(4, (int)((227 * 1024) / (smem + 1536))));
    const int grid = std::min(p.total_tiles, plan->num_sms * per_sm);
    kern<<<grid, kStThreads, smem, st>>>(p);
    SPB_CHECK_LAUNCH();
}

template <int CIN, typename T>
static void launch_stem_wide_t(const StemTcPlan* plan, const float* img, void* dst, int B, int H, int W, cudaStream_t st) {
    auto kern = stem_wide_kernel<CIN, T>;
    const size_t smem = (size_t)4 * 128 * 128 + (size_t)2 * CIN * 64 * 128 + (size_t)4 * CIN * kStIH * kStIWp * 2 + 1024;
    SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StemParams p = plan->params;
    p.img = img; p.dst = dst; p.H = H; p.W = W;
    p.tiles_x = (W / 4 + kStPW - 1) / kStPW;
    p.tiles_per_img = p.tiles_x * ((H / 4 + kStPH - 1) / kStPH);
    p.total_tiles = p.tiles_per_img * B;
    const int per_sm = std::max(1, std::min(2, (int)((227 * 1024) / (smem + 1536))));
    const int grid = std::min(p.total_tiles, plan->num_sms * per_sm);
    kern<<<grid, kStThreads, smem, st>>>(p);
    SPB_CHECK_LAUNCH();
}

// Split-precision stem: dst is NHWC [B][H/4][W/4][128] 16-bit in the split layout (64 channels as hi + lo).
void launch_stem_wide(const StemTcPlan* plan, const float* img, void* dst, int B, int H, int W, cudaStream_t st) {
    if (plan->operand_type == PREC_FP16) {
        if (plan->cin == 1) launch_stem_wide_t<1, __half>(plan, img, dst, B, H, W, st);
        else launch_stem_wide_t<3, __half>(plan, img, dst, B, H, W, st);
    } else {
        if (plan->cin == 1) launch_stem_wide_t<1, __nv_bfloat16>(plan, img, dst, B, H, W, st);
        else launch_stem_wide_t<3, __nv_bfloat16>(plan, img, dst, B, H, W, st);
    }
}

void launch_stem_tc(const StemTcPlan* plan, const float* img, void* dst, int B, int H, int W, cudaStream_t st) {
    if (plan->operand_type == PREC_FP16) {
        if (plan->cin == 1) launch_stem_tc_t<1, true, __half>(plan, img, dst, B, H, W, st);
        else launch_stem_tc_t<3, true, __half>(plan, img, dst, B, H, W, st);
    } else {
        if (plan->cin == 1) launch_stem_tc_t<1, true, __nv_bfloat16>(plan, img, dst, B, H, W, st);
        else launch_stem_tc_t<3, true, __nv_bfloat16>(plan, img, dst, B, H, W, st);
    }
}

// w16: device [64][nparts*nchunk*64] 16-bit K-major (k = c*64 + ky*8 + kx, zero where ky or kx is 7; the hi
// parts first, then the lo parts w - float(w_hi)), bias: device [64] fp32
StemTcPlan* stem_tc_plan_create(const void* w16, const float* bias, int cin, int operand_type, int num_sms) {
    if (cin != 1 && cin != 3) throw std::invalid_argument("stem: input must have 1 or 3 channels");
    auto* plan = new StemTcPlan();
    std::memset(&plan->params, 0, sizeof(plan->params));
    plan->cin = cin;
    plan->operand_type = operand_type;
    plan->num_sms = num_sms;
    const int nchunk = cin;
    const CUtensorMapDataType dt = operand_type == PREC_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const int nparts = 2;
    cuuint64_t dims[2] = {(cuuint64_t)nparts * nchunk * 64, 64};
    cuuint64_t str[1] = {(cuuint64_t)nparts * nchunk * 64 * 2};
    cuuint32_t box[2] = {64, 64};
    tc_encode_tiled(&plan->params.tmW, dt, 2, w16, dims, str, box);
    plan->params.bias = bias;
    return plan;
}

void stem_tc_plan_destroy(StemTcPlan* plan) { delete plan; }

}  // namespace spb200
