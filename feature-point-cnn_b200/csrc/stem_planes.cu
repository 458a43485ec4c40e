// Gray stem without an im2col pass: conv7x7/s2/p3 (+folded BatchNorm) + ReLU + maxpool3x3/s2/p1
// (reference python/src/superpoint.py:12-15,20-23) with the tensor core reading the image itself.
//
// 1. planes_kernel turns the image (fp32 in [0,1] or 8-bit) into 16-bit values x255 - exact for 8-bit frames - split
//    into two row-parity planes: plane[par][y2][x] = image row 2*y2 + par.
// 2. stem_planes_kernel: a CTA produces 15x15 pooled pixels x 64 channels from the 32x32 convolution outputs around
//    them.  An im2col row of the stride-2 convolution is, per filter row ky, eight consecutive pixels starting at
//    column 2*cx - 3: for the conv columns cx = cx0 + 4m + s of one PHASE s these 16-byte pieces tile a row of the
//    plane without gaps, and with the rows split by parity the pieces of consecutive conv rows follow each other
//    too.  So a copy of the plane patch whose left edge is shifted by 2s pixels (64 pixels x 35 rows per phase and
//    row parity) IS the A operand in the canonical non-swizzled K-major layout: 8 GEMM rows = 128 contiguous
//    bytes (SBO = 128 B), the second filter row of a K = 16 step = the other parity plane (LBO = one plane).
//    TMA loads the raw 80-pixel x 35-row patches (zero fill outside the image = the convolution padding); its
//    start column has to be 16-byte aligned and the phase shifts are odd pixel counts, so two warps make the
//    four shifted copies with funnel shifts (3 LDS.128 + 7 SHF + 4 STS.128 per 16-byte chunk, all four phases):
//    one quarter of the shared-memory volume of an im2col matrix and no conversion work.
//    GEMM row u of M-tile t = conv pixel (cy0 + 16t + u/8, cx0 + 4(u%8) + s): the four phases of a thread are four
//    horizontally adjacent conv pixels, so the horizontal half of the pooling stays in registers (one shuffle for
//    the column owned by the neighbour), rows meet in shared memory for the vertical half.
//    K order k = ky*8 + kx (kx, ky padded to 8 with zero weights), weights split hi + lo (two MMAs) and 1/255
//    applied to the fp32 accumulator exactly as in stem_tc.cu; results are bit-identical to that kernel.
//    Warp 8 issues TMA and MMA (accumulators of two M-tiles double buffered in TMEM, two raw and two copy stages),
//    warps 9-10 shift, warps 0-7 run the epilogue: warp w reads TMEM lanes 32(w%4).., channels 32(w/4)...
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"

namespace spb200 {

constexpr int kSpP = 15;                         // pooled tile edge
constexpr int kSpRows = 35;                      // plane rows per copy: 32 conv rows + 3 (filter rows 2j, 2j+1, j <= 3)
constexpr int kSpCopy = kSpRows * 128;           // one parity plane of one phase copy, bytes
constexpr int kSpStage = 8 * kSpCopy;            // 4 phases x 2 parities
constexpr int kSpX = 32 * 16 * 128;              // x-pooled rows: [conv row][pooled px 16][64 ch] 16-bit
constexpr int kSpRawPlane = 44 * 128;             // 35 rows x 160 B of one parity plane as loaded, padded to 128 B
constexpr int kSpRawStage = 2 * kSpRawPlane;
constexpr int kSpRawPlaneF = 88 * 128;            // the same patch of an fp32 image (35 rows x 320 B): no plane pass, the shifters convert
constexpr int kSpRawStageF = 2 * kSpRawPlaneF;
constexpr int kSpShiftWarps = 2;
constexpr int kSpThreads = (9 + kSpShiftWarps) * 32;
constexpr int kSpSmem = 16384 + 2 * kSpStage + 2 * kSpRawStage + kSpX + 1024;
constexpr int kSpSmemF = 16384 + 2 * kSpStage + 2 * kSpRawStageF + kSpX + 1024;

struct StemPlanesParams {
    CUtensorMap tmW;          // [64 cout][128] K-major 16-bit: hi | lo, 128B-swizzled boxes of 64
    CUtensorMap tmI;          // planes [2B][H/2][W] 16-bit, box 80 x 35 x 1, no swizzle; F32SRC: the image as {W, row parity, H/2, B} fp32,
                              // box 80 x 1 x 35 x 1
    const float* bias;        // [64]
    void* dst;                // NHWC [B][PH][PW][64] 16-bit
    int PH, PW;
    int tiles_x, tiles_per_img, total_tiles;
};

// ---- image -> parity planes --------------------------------------------------------------------------
template <typename T, typename IN>
__global__ void __launch_bounds__(256) planes_kernel(const IN* __restrict__ img, T* __restrict__ planes, int H, int W, long total8) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;      // one thread = 8 pixels of one row
    pdl_trigger();
    pdl_wait();                                                      // the planes may still be read by a stem kernel in flight
    if (i >= total8) return;
    const int w8 = W / 8;
    const int x8 = (int)(i % w8);
    const long row = i / w8;                                         // b * H + y
    const int y = (int)(row % H);
    const long b = row / H;
    float v[8];
    if (sizeof(IN) == 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(img) + 2 * i), c = __ldg(reinterpret_cast<const float4*>(img) + 2 * i + 1);
        v[0] = a.x * 255.f; v[1] = a.y * 255.f; v[2] = a.z * 255.f; v[3] = a.w * 255.f;
        v[4] = c.x * 255.f; v[5] = c.y * 255.f; v[6] = c.z * 255.f; v[7] = c.w * 255.f;
    } else {
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(img) + i);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[e] = (float)((a.x >> (8 * e)) & 255u);
            v[4 + e] = (float)((a.y >> (8 * e)) & 255u);
        }
    }
    uint4 o;
    o.x = pack2<T>(v[0], v[1]); o.y = pack2<T>(v[2], v[3]); o.z = pack2<T>(v[4], v[5]); o.w = pack2<T>(v[6], v[7]);
    T* dst = planes + (((size_t)b * 2 + (y & 1)) * (H / 2) + (y >> 1)) * W + 8 * x8;
    *reinterpret_cast<uint4*>(dst) = o;
}

__device__ __forceinline__ void tma_load_3d_a(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// F32SRC: the raw patches come from the fp32 image itself (row parity is a dimension of the tensor map) and the shifter warps
// do the plane pass's conversion - pixel x 255 rounded to 16 bits, the same two operations - on the way: five 16-byte loads,
// 14 multiplications and seven conversions per 16-byte chunk of the four copies instead of 3 loads and 7 funnel shifts.
template <typename T, bool F32SRC>
__global__ void __launch_bounds__(kSpThreads, 1) stem_planes_kernel(const __grid_constant__ StemPlanesParams p) {
    constexpr int kRawPlane = F32SRC ? kSpRawPlaneF : kSpRawPlane, kRawStage = 2 * kRawPlane;
    constexpr uint32_t kIdesc = (1u << 4) | (OperandFmt<T>::value << 7) | (OperandFmt<T>::value << 10) |
                                ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    extern __shared__ uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t bar_w, bar_rfull[2], bar_rempty[2], bar_cfull[2], bar_cempty[2], bar_afull[8], bar_aempty[8];
    __shared__ uint32_t tmem_slot;

    uint8_t* base = dyn_smem + ((1024u - (smem_u32(dyn_smem) & 1023u)) & 1023u);
    uint8_t* s_w = base;                         // hi [64][128 B], lo [64][128 B] (swizzled by TMA)
    uint8_t* s_cp = s_w + 16384;                 // [2 stages][4 phases][odd rows, even rows][35 rows][128 B]: the A operand
    uint8_t* s_raw = s_cp + 2 * kSpStage;        // [2 stages][odd rows, even rows][35 rows][160 B] as loaded by TMA
    uint8_t* s_x = s_raw + 2 * kRawStage;        // [32 conv rows][16 pooled px][128 B], 16-byte chunks XORed with (px >> 1)

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid / 32, 0), lane = tid % 32;
    const int PH = p.PH, PW = p.PW, CH = 2 * PH, CW = 2 * PW;
    pdl_trigger();

    if (tid == 0) {
        mbar_init(&bar_w, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_rfull[i], 1);
            mbar_init(&bar_rempty[i], kSpShiftWarps);
            mbar_init(&bar_cfull[i], kSpShiftWarps);
            mbar_init(&bar_cempty[i], 1);
            for (int k = 0; k < 4; ++k) { mbar_init(&bar_afull[4 * i + k], 1); mbar_init(&bar_aempty[4 * i + k], 8); }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        prefetch_tmap(&p.tmW);
        prefetch_tmap(&p.tmI);
    }
    if (warp == 8) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 8) {
        // ---------------- TMA loads of the raw patches (two tiles ahead) and MMA issue, one elected lane ----------------
        if (elect_one()) {
            const uint32_t cp_addr = smem_u32(s_cp), raw_addr = smem_u32(s_raw);
            auto load_tile = [&](int tile, int stage) {
                const int b = tile / p.tiles_per_img, tt = tile % p.tiles_per_img;
                const int p0 = (tt / p.tiles_x) * kSpP, q0 = (tt % p.tiles_x) * kSpP;
                const uint32_t bar = smem_u32(&bar_rfull[stage]);
                const uint32_t dst = raw_addr + stage * kRawStage;
                // TMA start columns must be 16-byte aligned: the patch starts at the multiple of 8 below 4 q0 - 5
                const int xs = 4 * q0 - 8;                       // q0 even: 4 q0 - 5 - 3; q0 odd: 4 q0 - 5 - 7 = 4 q0 - 12
                mbar_expect_tx_a(bar, 2 * kSpRows * (F32SRC ? 320 : 160));
                // odd image rows from y2 = cy0 - 2, even rows from y2 = cy0 - 1, with cy0 = 2 p0 - 1
                if (F32SRC) {
                    tma_load_4d_a(dst, &p.tmI, bar, xs - 4 * (q0 & 1), 1, 2 * p0 - 3, b);
                    tma_load_4d_a(dst + kRawPlane, &p.tmI, bar, xs - 4 * (q0 & 1), 0, 2 * p0 - 2, b);
                } else {
                    tma_load_3d_a(dst, &p.tmI, bar, xs - 4 * (q0 & 1), 2 * p0 - 3, 2 * b + 1);
                    tma_load_3d_a(dst + kRawPlane, &p.tmI, bar, xs - 4 * (q0 & 1), 2 * p0 - 2, 2 * b);
                }
            };
            mbar_expect_tx(&bar_w, 16384);
            tma_load_2d(s_w, &p.tmW, &bar_w, 0, 0);
            tma_load_2d(s_w + 8192, &p.tmW, &bar_w, 64, 0);
            pdl_wait();                                          // the image planes are the previous kernel's output
            if ((int)blockIdx.x < p.total_tiles) load_tile(blockIdx.x, 0);
            if ((int)(blockIdx.x + gridDim.x) < p.total_tiles) load_tile(blockIdx.x + gridDim.x, 1);
            mbar_wait(&bar_w, 0);
            const uint32_t w_lo = umma_desc_lo(smem_u32(s_w));
            constexpr uint32_t kHiW = (1024u >> 4) | (1u << 14) | (2u << 29);       // SBO 1024, version 1, SWIZZLE_128B
            constexpr uint32_t kHiA = (128u >> 4) | (1u << 14);                     // SBO 128, version 1, no swizzle
            constexpr uint32_t kLboA = (uint32_t)(kSpCopy >> 4) << 16;              // LBO = one parity plane
            int it = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
                const int stage = it & 1;
                const uint32_t par = (it >> 1) & 1;
                const int ahead = tile + 2 * gridDim.x;
                if (ahead < p.total_tiles) {
                    mbar_wait(&bar_rempty[stage], par);                 // the shifters have read this tile's raw patch
                    load_tile(ahead, stage);
                }
                mbar_wait(&bar_cfull[stage], par);
                // accumulators are handed over per (M-tile, phase): the MMAs of the next tile start as soon as the epilogue
                // has read one phase of this one, so the tensor core's shared-memory traffic spreads over the whole tile
                // instead of piling up behind the vertical pooling pass
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        if (it >= 1) mbar_wait(&bar_aempty[t * 4 + s], (it - 1) & 1);
                        tc_fence_after();
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#pragma unroll
                            for (int part = 0; part < 2; ++part) {
                                const uint32_t a_addr = cp_addr + stage * kSpStage + s * 2 * kSpCopy + (16 * t + j) * 128;
                                const uint32_t a = ((a_addr >> 4) & 0x3fffu) | kLboA;
                                const uint32_t w = w_lo + (uint32_t)((part * 8192 + j * 32) >> 4);
                                umma_f16_w(tmem + t * 256 + s * 64, a, kHiA, w, kHiW, kIdesc, (j > 0 || part > 0) ? 1u : 0u);
                            }
                        umma_commit(&bar_afull[t * 4 + s]);
                    }
                umma_commit(&bar_cempty[stage]);
            }
        }
        __syncwarp();
    } else if (warp > 8) {
        // ---------------- shifter warps: the four phase copies of the raw patch ----------------
        // copy s, row r, 16-byte chunk m = raw elements 8m + o0 + 2s .. + 7 of row r, o0 = 3 (q0 even) or 7 (q0 odd):
        // an odd element offset, i.e. every output word is a funnel shift of two neighbouring raw words
        const int st = tid - 9 * 32;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int stage = it & 1;
            const uint32_t par = (it >> 1) & 1;
            const int tt = tile % p.tiles_per_img;
            const int c0 = ((tt % p.tiles_x) & 1) ? 3 : 1;               // (o0 - 1) / 2; q0 = 15 tx has the parity of tx
            mbar_wait(&bar_rfull[stage], par);
            if (it >= 2) mbar_wait(&bar_cempty[stage], par ^ 1u);        // the MMAs of tile it-2 have read the copies
            const uint8_t* raw = s_raw + stage * kRawStage;
            uint8_t* cp = s_cp + stage * kSpStage;
            for (int i = st; i < 2 * kSpRows * 8; i += 32 * kSpShiftWarps) {
                const int m = i & 7, row = i >> 3;
                const int pl = row >= kSpRows ? 1 : 0, r = row - pl * kSpRows;
                uint32_t F[7];
                if (F32SRC) {
                    // pixels 8m + o0 .. 8m + o0 + 13 of the raw row: with the first load at pixel 8m (o0 = 3) or 8m + 4 (o0 = 7) they
                    // are elements 3 .. 16 of five 16-byte loads
                    const float4* src = reinterpret_cast<const float4*>(raw + pl * kRawPlane + r * 320) + 2 * m + (c0 == 3 ? 1 : 0);
                    float f[20];
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const float4 q = src[k];
                        f[4 * k] = q.x; f[4 * k + 1] = q.y; f[4 * k + 2] = q.z; f[4 * k + 3] = q.w;
                    }
#pragma unroll
                    for (int k = 0; k < 7; ++k) F[k] = pack2<T>(f[3 + 2 * k] * 255.f, f[4 + 2 * k] * 255.f);
                } else {
                    const uint4* src = reinterpret_cast<const uint4*>(raw + pl * kRawPlane + r * 160 + m * 16);
                    const uint4 a = src[0], bq = src[1], cq = src[2];
                    const uint32_t R[12] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w, cq.x, cq.y, cq.z, cq.w};
                    if (c0 == 1) {
#pragma unroll
                        for (int k = 0; k < 7; ++k) F[k] = __funnelshift_r(R[1 + k], R[2 + k], 16);
                    } else {
#pragma unroll
                        for (int k = 0; k < 7; ++k) F[k] = __funnelshift_r(R[3 + k], R[4 + k], 16);
                    }
                }
                uint8_t* dst = cp + pl * kSpCopy + r * 128 + m * 16;
#pragma unroll
                for (int s = 0; s < 4; ++s)
                    *reinterpret_cast<uint4*>(dst + s * 2 * kSpCopy) = make_uint4(F[s], F[s + 1], F[s + 2], F[s + 3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> tensor-core reads
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bar_cfull[stage]);
                mbar_arrive(&bar_rempty[stage]);
            }
        }
    } else {
        // ---------------- epilogue warps ----------------
        pdl_wait();                                              // the pooled tensor is read by kernels of the previous step
        const int q = warp & 3, h = warp >> 2;
        const int u = q * 32 + lane, crow = u >> 3, mp = u & 7;
        float bias[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) bias[e] = __ldg(p.bias + 32 * h + e);
        T* out = static_cast<T*>(p.dst);
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int b = tile / p.tiles_per_img, tt = tile % p.tiles_per_img;
            const int p0 = (tt / p.tiles_x) * kSpP, q0 = (tt % p.tiles_x) * kSpP;
            const int cy0 = 2 * p0 - 1, cx0 = 2 * q0 - 1;
#pragma unroll 1
            for (int t = 0; t < 2; ++t) {
                const int c = 16 * t + crow, cy = cy0 + c;
                const bool rowok = cy >= 0 && cy < CH;
                uint32_t E[16], Q[16];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    uint32_t r[32];
                    mbar_wait(&bar_afull[t * 4 + s], it & 1);
                    tc_fence_after();
                    tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256 + s * 64 + h * 32), r);
                    const int cx = cx0 + 4 * mp + s;
                    const uint32_t m = (rowok && cx >= 0 && cx < CW) ? 0xffffffffu : 0u;   // outside the conv output: 0 (neutral under ReLU)
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_aempty[t * 4 + s]);
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const uint32_t v = pack2<T>(fmaf(__uint_as_float(r[2 * e]), 1.f / 255.f, bias[2 * e]),
                                                    fmaf(__uint_as_float(r[2 * e + 1]), 1.f / 255.f, bias[2 * e + 1])) & m;
                        if (s == 0) { E[e] = v; Q[e] = __shfl_down_sync(0xffffffffu, v, 1); }   // Q starts with the neighbour's first column
                        else if (s == 1) E[e] = max2<T>(E[e], v);
                        else if (s == 2) { E[e] = max2<T>(E[e], v); Q[e] = max2<T>(Q[e], v); }
                        else Q[e] = max2<T>(Q[e], v);
                    }
                }
                if (c < 31) {
                    uint8_t* xe = s_x + (c * 16 + 2 * mp) * 128;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int ch = ((4 * h + k) ^ mp) << 4;
                        *reinterpret_cast<uint4*>(xe + ch) = make_uint4(E[4 * k], E[4 * k + 1], E[4 * k + 2], E[4 * k + 3]);
                        *reinterpret_cast<uint4*>(xe + 128 + ch) = make_uint4(Q[4 * k], Q[4 * k + 1], Q[4 * k + 2], Q[4 * k + 3]);
                    }
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // vertical half of the pooling: pooled row k of the tile = conv rows 2k, 2k+1, 2k+2 of the tile; max with 0 = ReLU.
            // A thread owns one (pooled column, 8-channel chunk) and walks down half of the pooled rows with a running
            // row: two shared loads per output instead of three and no per-item index arithmetic (this pass shares the
            // shared-memory pipe with the next tile's MMA operand fetch, every load less shortens it).
            if (tid < 2 * kSpP * 8) {
                const int chunk = tid & 7, l = (tid >> 3) % kSpP, g = tid / (kSpP * 8);       // g = 0: pooled rows 0..7, 1: 8..14
                const int k0 = g * 8, k1 = g == 0 ? 8 : kSpP;
                const int px = q0 + l;
                const uint8_t* src = s_x + ((2 * k0) * 16 + l) * 128 + ((chunk ^ (l >> 1)) << 4);
                T* dst = out + ((size_t)(b * PH + p0 + k0) * PW + px) * 64 + chunk * 8;
                uint4 ev = *reinterpret_cast<const uint4*>(src);                               // conv row 2k
                for (int k = k0; k < k1; ++k) {
                    const uint4 od = *reinterpret_cast<const uint4*>(src + 2048);              // conv row 2k + 1
                    const uint4 nx = *reinterpret_cast<const uint4*>(src + 4096);              // conv row 2k + 2
                    uint4 m;
                    m.x = max2<T>(max2<T>(max2<T>(0u, ev.x), od.x), nx.x);
                    m.y = max2<T>(max2<T>(max2<T>(0u, ev.y), od.y), nx.y);
                    m.z = max2<T>(max2<T>(max2<T>(0u, ev.z), od.z), nx.z);
                    m.w = max2<T>(max2<T>(max2<T>(0u, ev.w), od.w), nx.w);
                    if (p0 + k < PH && px < PW) *reinterpret_cast<uint4*>(dst) = m;
                    ev = nx;
                    src += 4096;
                    dst += (size_t)PW * 64;
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

struct StemPlanesPlan {
    StemPlanesParams params;
    int operand_type, num_sms;
    const void* planes = nullptr;     // what tmI was encoded for
    int B = 0, H = 0, W = 0, f32 = -1;
};

static void encode_tiled_plain(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims,
                               const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    tc_encode_tiled_ex(map, dt, rank, base, dims, strides_bytes, box, false);
}

void launch_planes(const void* img, int img_is_u8, void* planes, int operand_type, int B, int H, int W, cudaStream_t st) {
    const long total8 = (long)B * H * (W / 8);
    const int grid = (int)((total8 + 255) / 256);
    if (operand_type == PREC_FP16) {
        if (img_is_u8) launch_pdl(planes_kernel<__half, uint8_t>, dim3(grid), dim3(256), 0, st, (const uint8_t*)img, (__half*)planes, H, W, total8);
        else launch_pdl(planes_kernel<__half, float>, dim3(grid), dim3(256), 0, st, (const float*)img, (__half*)planes, H, W, total8);
    } else {
        if (img_is_u8) launch_pdl(planes_kernel<__nv_bfloat16, uint8_t>, dim3(grid), dim3(256), 0, st, (const uint8_t*)img, (__nv_bfloat16*)planes, H, W, total8);
        else launch_pdl(planes_kernel<__nv_bfloat16, float>, dim3(grid), dim3(256), 0, st, (const float*)img, (__nv_bfloat16*)planes, H, W, total8);
    }
}

// src: the 16-bit parity planes of launch_planes (src_is_f32 = 0), or the fp32 image [B][H][W] itself (1; 16-byte aligned)
void launch_stem_planes(StemPlanesPlan* plan, const void* src, int src_is_f32, void* dst, int B, int H, int W, cudaStream_t st) {
    const CUtensorMapDataType dt = plan->operand_type == PREC_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if (plan->planes != src || plan->B != B || plan->H != H || plan->W != W || plan->f32 != src_is_f32) {
        if (src_is_f32) {
            if (reinterpret_cast<uintptr_t>(src) & 15) throw std::invalid_argument("stem: the fp32 image must be 16-byte aligned");
            cuuint64_t dims[4] = {(cuuint64_t)W, 2, (cuuint64_t)(H / 2), (cuuint64_t)B};
            cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * 8, (cuuint64_t)H * W * 4};
            cuuint32_t box[4] = {80, 1, (cuuint32_t)kSpRows, 1};
            encode_tiled_plain(&plan->params.tmI, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, src, dims, str, box);
        } else {
            cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)(H / 2), (cuuint64_t)2 * B};
            cuuint64_t str[2] = {(cuuint64_t)W * 2, (cuuint64_t)(H / 2) * W * 2};
            cuuint32_t box[3] = {80, (cuuint32_t)kSpRows, 1};
            encode_tiled_plain(&plan->params.tmI, dt, 3, src, dims, str, box);
        }
        plan->planes = src; plan->B = B; plan->H = H; plan->W = W; plan->f32 = src_is_f32;
    }
    StemPlanesParams p = plan->params;
    p.dst = dst;
    p.PH = H / 4; p.PW = W / 4;
    p.tiles_x = (p.PW + kSpP - 1) / kSpP;
    p.tiles_per_img = p.tiles_x * ((p.PH + kSpP - 1) / kSpP);
    p.total_tiles = p.tiles_per_img * B;
    const int grid = std::min(p.total_tiles, plan->num_sms);
    auto go = [&](auto kern, int smem) {
        SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        launch_pdl(kern, dim3(grid), dim3(kSpThreads), (size_t)smem, st, p);
    };
    if (plan->operand_type == PREC_FP16) {
        if (src_is_f32) go(stem_planes_kernel<__half, true>, kSpSmemF); else go(stem_planes_kernel<__half, false>, kSpSmem);
    } else {
        if (src_is_f32) go(stem_planes_kernel<__nv_bfloat16, true>, kSpSmemF); else go(stem_planes_kernel<__nv_bfloat16, false>, kSpSmem);
    }
}

// w16: device [64][128] 16-bit K-major, k = ky*8 + kx hi parts then lo parts (the 1-channel pack of stem_tc.cu)
StemPlanesPlan* stem_planes_plan_create(const void* w16, const float* bias, int operand_type, int num_sms) {
    auto* plan = new StemPlanesPlan();
    std::memset(&plan->params, 0, sizeof(plan->params));
    plan->operand_type = operand_type;
    plan->num_sms = num_sms;
    const CUtensorMapDataType dt = operand_type == PREC_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    cuuint64_t dims[2] = {128, 64};
    cuuint64_t str[1] = {128 * 2};
    cuuint32_t box[2] = {64, 64};
    tc_encode_tiled(&plan->params.tmW, dt, 2, w16, dims, str, box);
    plan->params.bias = bias;
    return plan;
}

void stem_planes_plan_destroy(StemPlanesPlan* plan) { delete plan; }

}  // namespace spb200
