// Fused residual block on tcgen05 / TMEM / TMA (sm_100a), haloed-tile version for stride-1 blocks.
//
// Same arithmetic as block_tc.cu (reference python/src/resnet_blocks.py:14-27):
//     Y   = relu(conv3x3(X) * bn1)                        GEMM 1   D1[128, N] = im2col(X) . W1^T
//     OUT = relu(conv1x1(Y) * bn2 + shortcut(X))          GEMM 2   D2[128, N] = Y . W2^T (+ Xc . Wd^T)
// but the activations of a 128-pixel output tile are fetched ONCE per 64-channel chunk as a haloed
// 18 x 10 pixel box (one 4-d TMA load, zero fill outside the image = the convolution's padding)
// instead of once per filter tap.  The tile is 16 "slow" rows of 8 "group" pixels; in the haloed box a
// pixel is one 128-byte swizzled row, so the A operand of tap (ds, dg) is the same box read through a
// UMMA shared-memory descriptor whose start is shifted by (ds * 10 + dg) * 128 B and whose 8-row-group
// pitch (SBO) is one haloed row = 1280 B.  That cuts the L2 -> SM traffic of the 3x3 convolution from
// 9 x 16 KB to 23 KB per chunk per tile.  The group axis is x or y, whichever tiles the image better
// (the TMA tensor map just orders the two spatial dimensions differently).
//
// Weights stream through their own ring, one [N x 64] K-slab per (chunk, tap); T = 2 output tiles share
// every slab for the 128-channel layers (halves the weight traffic per pixel), while the 64-channel
// blocks keep all 10-11 slabs resident in shared memory for the whole kernel.  The 1x1 shortcut
// convolution is issued while its chunk of X is resident (centre-tap view) straight into the GEMM 2
// accumulator; an identity shortcut is added in the second epilogue.
//
// Warp roles (352 threads, one persistent CTA per SM): warp 0 = activation TMA producer, warp 1 = weight
// TMA producer, warp 2 = TMEM allocator + MMA issuer, warps 3-10 = epilogue (two warps per TMEM lane
// quarter: one per tile when T = 2, one per column half when T = 1).  With NBUF = 2 (64-channel blocks,
// transposed-conv phases) the accumulators and Y are double buffered, so GEMM 1 of the next tile runs
// under the epilogues of the current one.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "kernels.h"
#include "tc_common.cuh"

namespace spb200 {

constexpr int kHaloG = 10;                                  // 8 + 2 pixels along the group axis
constexpr int kHaloS = 18;                                  // 16 + 2 rows along the slow axis
constexpr int kHaloLoadBytes = kHaloS * kHaloG * 128;       // 23040 B landed by one TMA box
constexpr int kHaloBufBytes = 23552;                        // rounded up to a multiple of 1024
constexpr int kHaloSbo = kHaloG * 128;                      // 8-pixel groups are one haloed row apart
constexpr int kMaxSteps = 48;
constexpr int kMaxChunks = 8;
constexpr int kHaloThreads = 352;

// One weight slab [N x 64 K] and the MMAs that consume it.
struct HaloStep {
    uint16_t kcoord;     // K coordinate / 64 in the weight tensor (W1 for gemm 0, W2 otherwise)
    uint16_t a_off16;    // gemm 0/1: byte offset / 16 of the tap view inside the haloed box; gemm 2: of the Y chunk
    uint8_t gemm;        // 0: 3x3 taps -> D1;  1: 1x1 shortcut (centre view) -> D2;  2: 1x1 over Y -> D2
    uint8_t nkk;         // K = 16 MMA steps in this slab (1..4)
    uint8_t flags;       // bit 0: first slab of an activation chunk; bit 1: last slab using it
    uint8_t pad;
};

struct HaloParams {
    CUtensorMap tmA[kMaxSegs];
    CUtensorMap tmW1, tmW2;
    HaloStep steps[kMaxSteps];
    int nsteps, n1steps;           // all slabs; slabs of GEMM 1 + shortcut (they come first)
    int chunk_seg[kMaxChunks], chunk_c0[kMaxChunks], nchunks;
    int lo_s, lo_g;                // origin of the haloed box relative to the tile origin
    int orient;                    // 0: group axis = x (tile 16 rows x 8 cols); 1: group axis = y (8 rows x 16 cols)
    int tiles_g, tiles_per_img, total_tiles, n_super;
    int OH, OW;
    int has_ds;
    const float* bias1;
    const float* bias2;
    const void* residual;
    void* dst;
    int res_C, dst_H, dst_W, dst_C, dst_stride, dst_off_y, dst_off_x;
    int relu, dst_fp32, n_mma;
    unsigned long long* stats;     // debug (SPB200_HALO_STATS=1): [grid][16] cycles spent waiting per role, else null
};

// mbarrier wait that adds the cycles it spent to a debug counter when stats are collected
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, unsigned long long* acc) {
    if (acc) {
        const long long t0 = clock64();
        mbar_wait(bar, parity);
        *acc += (unsigned long long)(clock64() - t0);
    } else {
        mbar_wait(bar, parity);
    }
}

template <int N, int T, int NBUF, int SA, int SW, bool FUSED, bool WRES, typename Tp>
__global__ void __launch_bounds__(kHaloThreads, 1) halo_tc_kernel(const __grid_constant__ HaloParams p) {
    constexpr int kWBytes = N * 128;
    constexpr int kYTile = (N / 64) * 16384;
    constexpr uint32_t kAccCols = NBUF * T * N;
    constexpr uint32_t kTmemCols = (FUSED ? 2 : 1) * kAccCols;
    static_assert(kTmemCols == 64 || kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512, "TMEM columns");
    static_assert(!(FUSED && NBUF == 2) || WRES, "pipelined GEMM 2 needs resident weights (slab order)");

    extern __shared__ uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t a_full[SA], a_empty[SA], w_full[SW], w_empty[SW];
    __shared__ __align__(8) uint64_t d1_full[NBUF], d1_empty[NBUF], y_full[NBUF], y_empty[NBUF], d2_full[NBUF], d2_empty[NBUF];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_bias1[N], s_bias2[N];

    uint8_t* a_ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dyn_smem) + 1023) & ~(uintptr_t)1023);
    uint8_t* w_ring = a_ring + SA * T * kHaloBufBytes;
    uint8_t* y_buf = w_ring + SW * kWBytes;                 // [NBUF][T][N/64][128 rows][128 B]
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t idesc = (1u << 4) | (OperandFmt<Tp>::value << 7) | (OperandFmt<Tp>::value << 10) |
                           ((uint32_t)(p.n_mma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int n_local = (p.n_super - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < SW; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(&d1_full[b], 1); mbar_init(&y_empty[b], 1); mbar_init(&d2_full[b], 1);
            mbar_init(&d1_empty[b], 8); mbar_init(&y_full[b], 8); mbar_init(&d2_empty[b], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kMaxSegs; ++s) prefetch_tmap(&p.tmA[s]);
        prefetch_tmap(&p.tmW1);
        if (FUSED) prefetch_tmap(&p.tmW2);
    }
    if (warp == 2) tmem_alloc(&tmem_slot, kTmemCols);
    for (int i = threadIdx.x; i < N; i += kHaloThreads) {
        s_bias1[i] = p.bias1[i];
        s_bias2[i] = FUSED ? p.bias2[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    unsigned long long st_acc[6] = {0, 0, 0, 0, 0, 0};
    const bool stats = p.stats != nullptr;
#define ST(i) (stats ? &st_acc[i] : nullptr)
    const long long t_begin = stats ? clock64() : 0;

    // The three issuing roles run as WHOLE warps: loop counters, ring indices and phases are warp-uniform
    // (they live in uniform registers), every lane polls the mbarriers, and only the instructions that must come
    // from one thread (TMA, tcgen05.mma, tcgen05.commit, expect_tx) sit under elect.sync.  Issuing from inside an
    // `if (lane == 0)` region instead costs ~200 cycles per tcgen05.mma (measured): the compiler has to move
    // every descriptor through R2UR and wrap each instruction in an active-lane loop.
    if (warp == 0) {
        // ============================ activation producer ============================
        int sa = 0;
        uint32_t pha = 0;
        for (int st = blockIdx.x; st < p.n_super; st += gridDim.x) {
            int cg[T], cs[T], cimg[T];
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const int tile = min(st * T + t, p.total_tiles - 1);
                const int tt = tile % p.tiles_per_img;
                cimg[t] = tile / p.tiles_per_img;
                cs[t] = (tt / p.tiles_g) * 16 + p.lo_s;
                cg[t] = (tt % p.tiles_g) * 8 + p.lo_g;
            }
            for (int ci = 0; ci < p.nchunks; ++ci) {
                mbar_wait_t(&a_empty[sa], pha ^ 1u, ST(0));
                if (elect_one()) {
                    mbar_expect_tx(&a_full[sa], (uint32_t)(T * kHaloLoadBytes));
                    const CUtensorMap* tm = &p.tmA[p.chunk_seg[ci]];
#pragma unroll
                    for (int t = 0; t < T; ++t)
                        tma_load_4d(a_ring + (sa * T + t) * kHaloBufBytes, tm, &a_full[sa], p.chunk_c0[ci], cg[t], cs[t], cimg[t]);
                }
                __syncwarp();
                if (++sa == SA) { sa = 0; pha ^= 1u; }
            }
        }
        if (stats && lane == 0) { p.stats[blockIdx.x * 16 + 6] = st_acc[0]; }
    } else if (warp == 1) {
        // ============================ weight producer ============================
        int sw = 0;
        uint32_t phw = 0;
        const int passes = WRES ? min(n_local, 1) : n_local;
        for (int j = 0; j < passes; ++j) {
            for (int e = 0; e < p.nsteps; ++e) {
                const HaloStep s = p.steps[e];
                if (!WRES) mbar_wait_t(&w_empty[sw], phw ^ 1u, ST(0));
                if (elect_one()) {
                    mbar_expect_tx(&w_full[sw], (uint32_t)kWBytes);
                    tma_load_2d(w_ring + sw * kWBytes, s.gemm == 0 ? &p.tmW1 : &p.tmW2, &w_full[sw], (int)s.kcoord * 64, 0);
                }
                __syncwarp();
                if (++sw == SW) { sw = 0; phw ^= 1u; }
            }
        }
        if (stats && lane == 0) { p.stats[blockIdx.x * 16 + 7] = st_acc[0]; }
    } else if (warp == 2) {
        // ============================ MMA issuer ============================
        int sa = 0, sw = 0;
        uint32_t pha = 0, phw = 0;
        const uint32_t a_base = smem_u32(a_ring), w_base = smem_u32(w_ring), y_base = smem_u32(y_buf);
        constexpr uint32_t kHiA = ((uint32_t)kHaloSbo >> 4) | (1u << 14) | (2u << 29);   // SBO = haloed row pitch
        constexpr uint32_t kHiB = (1024u >> 4) | (1u << 14) | (2u << 29);                // SBO = 1024 (dense tile)
        constexpr int LAG = (FUSED && NBUF == 2) ? 1 : 0;
        for (int j = 0; j < n_local + LAG; ++j) {
            if (j < n_local) {
                const int b = j % NBUF;
                const uint32_t ph = (uint32_t)(j / NBUF) & 1u;
                mbar_wait_t(&d1_empty[b], ph ^ 1u, ST(0));          // epilogue has drained D1[b]
                if (FUSED && p.has_ds) mbar_wait_t(&d2_empty[b], ph ^ 1u, ST(0));
                tc_fence_after();
                uint32_t acc1 = 0, accd = 0;
                for (int e = 0; e < p.n1steps; ++e) {
                    const HaloStep s = p.steps[e];
                    if (s.flags & 1) mbar_wait_t(&a_full[sa], pha, ST(1));
                    const int slot = WRES ? e : sw;
                    mbar_wait_t(&w_full[slot], WRES ? 0u : phw, ST(2));
                    tc_fence_after();
                    const uint32_t blo = umma_desc_lo(w_base + slot * kWBytes);
                    const uint32_t alo = umma_desc_lo(a_base + sa * T * kHaloBufBytes + (uint32_t)s.a_off16 * 16u);
                    const uint32_t acc = s.gemm == 0 ? acc1 : accd;
                    const uint32_t d0 = tmem_base + (s.gemm == 0 ? 0u : kAccCols) + (uint32_t)(b * T * N);
                    if (elect_one()) {
#pragma unroll
                        for (int t = 0; t < T; ++t) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                if (kk < s.nkk)
                                    umma_f16_w(d0 + t * N, alo + t * (kHaloBufBytes >> 4) + kk * 2, kHiA, blo + kk * 2, kHiB, idesc,
                                               acc | (uint32_t)kk);
                        }
                        if (!WRES) umma_commit(&w_empty[sw]);
                        if (s.flags & 2) umma_commit(&a_empty[sa]);
                    }
                    __syncwarp();
                    if (s.gemm == 0) acc1 = 1u; else accd = 1u;
                    if (!WRES) { if (++sw == SW) { sw = 0; phw ^= 1u; } }
                    if (s.flags & 2) { if (++sa == SA) { sa = 0; pha ^= 1u; } }
                }
                if (elect_one()) umma_commit(&d1_full[b]);
                __syncwarp();
            }
            if (FUSED && j >= LAG) {
                const int jj = j - LAG;
                const int b = jj % NBUF;
                const uint32_t ph = (uint32_t)(jj / NBUF) & 1u;
                mbar_wait_t(&y_full[b], ph, ST(3));                 // Y written by the epilogue warps
                if (!p.has_ds) mbar_wait_t(&d2_empty[b], ph ^ 1u, ST(3));
                tc_fence_after();
                uint32_t acc = p.has_ds ? 1u : 0u;
                for (int e = p.n1steps; e < p.nsteps; ++e) {
                    const HaloStep s = p.steps[e];
                    const int slot = WRES ? e : sw;
                    mbar_wait_t(&w_full[slot], WRES ? 0u : phw, ST(4));
                    tc_fence_after();
                    const uint32_t blo = umma_desc_lo(w_base + slot * kWBytes);
                    const uint32_t alo = umma_desc_lo(y_base + b * T * kYTile + (uint32_t)s.a_off16 * 16u);
                    const uint32_t d0 = tmem_base + kAccCols + (uint32_t)(b * T * N);
                    if (elect_one()) {
#pragma unroll
                        for (int t = 0; t < T; ++t) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk)
                                if (kk < s.nkk)
                                    umma_f16_w(d0 + t * N, alo + t * (kYTile >> 4) + kk * 2, kHiB, blo + kk * 2, kHiB, idesc,
                                               acc | (uint32_t)kk);
                        }
                        if (!WRES) umma_commit(&w_empty[sw]);
                    }
                    __syncwarp();
                    acc = 1u;
                    if (!WRES) { if (++sw == SW) { sw = 0; phw ^= 1u; } }
                }
                if (elect_one()) {
                    umma_commit(&d2_full[b]);
                    umma_commit(&y_empty[b]);
                }
                __syncwarp();
            }
        }
        if (stats && lane == 0) {
            unsigned long long* o = p.stats + blockIdx.x * 16;
            o[0] = (unsigned long long)(clock64() - t_begin);
            for (int i = 0; i < 5; ++i) o[1 + i] = st_acc[i];
        }
    } else {
        // ============================ epilogue ============================
        const int q = warp & 3;                                // TMEM lane quarter this warp may read
        const int eh = (warp - 3) >> 2;                        // 0/1: tile (T = 2) or column half (T = 1)
        const int row = q * 32 + lane;                         // GEMM row = TMEM lane
        const int ti = row >> 3, tr = row & 7;                 // slow row, pixel within the group
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int t = T == 2 ? eh : 0;
        const int nblk = p.n_mma / 32;
        const int blk_lo = T == 2 ? 0 : (eh == 0 ? 0 : (nblk + 1) / 2);
        const int blk_hi = T == 2 ? nblk : (eh == 0 ? (nblk + 1) / 2 : nblk);
        for (int j = 0; j < n_local; ++j) {
            const int st = blockIdx.x + j * gridDim.x;
            const int b = j % NBUF;
            const uint32_t ph = (uint32_t)(j / NBUF) & 1u;
            const int tile_raw = st * T + t;
            const int tile = min(tile_raw, p.total_tiles - 1);
            const int img = tile / p.tiles_per_img, tt = tile % p.tiles_per_img;
            const int s0 = (tt / p.tiles_g) * 16, g0 = (tt % p.tiles_g) * 8;
            const int oy = p.orient == 0 ? s0 + ti : g0 + tr;
            const int ox = p.orient == 0 ? g0 + tr : s0 + ti;
            const bool valid = tile_raw < p.total_tiles && oy < p.OH && ox < p.OW;
            const size_t gpix = ((size_t)img * p.OH + oy) * p.OW + ox;
            const size_t dpix = ((size_t)img * p.dst_H + (oy * p.dst_stride + p.dst_off_y)) * p.dst_W +
                                (ox * p.dst_stride + p.dst_off_x);
            const uint32_t tmem_d1 = tmem_base + (uint32_t)((b * T + t) * N) + lane_off;
            const uint32_t tmem_d2 = tmem_d1 + kAccCols;
            mbar_wait_t(&d1_full[b], ph, ST(0));
            tc_fence_after();
            if (FUSED) {
                mbar_wait_t(&y_empty[b], ph ^ 1u, ST(1));               // GEMM 2 of the previous use has finished reading Y[b]
                uint8_t* yrow = y_buf + (b * T + t) * kYTile + row * 128;
#pragma unroll 1
                for (int blk = blk_lo; blk < blk_hi; ++blk) {
                    const int c0 = blk * 32;
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_d1 + (uint32_t)c0, r);
                    tmem_ld_wait();
                    uint8_t* ychunk = yrow + (c0 >> 6) * 16384;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = fmaxf(__uint_as_float(r[jj * 8 + e]) + s_bias1[c0 + jj * 8 + e], 0.f);
                        const int cj = ((c0 & 63) >> 3) + jj;
                        *reinterpret_cast<uint4*>(ychunk + ((cj ^ (row & 7)) << 4)) =
                            make_uint4(pack2<Tp>(v[0], v[1]), pack2<Tp>(v[2], v[3]), pack2<Tp>(v[4], v[5]), pack2<Tp>(v[6], v[7]));
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&d1_empty[b]); mbar_arrive(&y_full[b]); }
                mbar_wait_t(&d2_full[b], ph, ST(2));
                tc_fence_after();
            }
            const uint32_t tmem_out = FUSED ? tmem_d2 : tmem_d1;
            const float* sb = FUSED ? s_bias2 : s_bias1;
#pragma unroll 1
            for (int blk = blk_lo; blk < blk_hi; ++blk) {
                const int c0 = blk * 32;
                uint32_t r[32];
                tmem_ld_32x32(tmem_out + (uint32_t)c0, r);
                tmem_ld_wait();
                if (valid) {
                    float v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + sb[c0 + i];
                    if (p.residual) {
                        const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const Tp*>(p.residual) + gpix * p.res_C + c0);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const uint4 u = __ldg(rp + jj);
                            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = unpack2<Tp>(w[e]);
                                v[jj * 8 + e * 2] += f.x;
                                v[jj * 8 + e * 2 + 1] += f.y;
                            }
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    if (p.dst_fp32) {
                        float4* dp = reinterpret_cast<float4*>(static_cast<float*>(p.dst) + dpix * p.dst_C + c0);
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) dp[jj] = make_float4(v[jj * 4], v[jj * 4 + 1], v[jj * 4 + 2], v[jj * 4 + 3]);
                    } else {
                        uint4* dp = reinterpret_cast<uint4*>(static_cast<Tp*>(p.dst) + dpix * p.dst_C + c0);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            dp[jj] = make_uint4(pack2<Tp>(v[jj * 8], v[jj * 8 + 1]), pack2<Tp>(v[jj * 8 + 2], v[jj * 8 + 3]),
                                                pack2<Tp>(v[jj * 8 + 4], v[jj * 8 + 5]), pack2<Tp>(v[jj * 8 + 6], v[jj * 8 + 7]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(FUSED ? &d2_empty[b] : &d1_empty[b]);
        }
        if (stats && warp == 3 && lane == 0) {
            unsigned long long* o = p.stats + blockIdx.x * 16;
            o[8] = st_acc[0]; o[9] = st_acc[1]; o[10] = st_acc[2];
            o[11] = (unsigned long long)(clock64() - t_begin);
        }
    }
#undef ST
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_slot, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// Host
// ------------------------------------------------------------------------------------------------
struct TcHaloPlan {
    HaloParams params;
    int variant;       // 0: N=64 fused, resident weights; 1: N=128 fused, T=2; 2: N=128 single convolution, T=2
    int operand_type, grid;
};

template <int N, int T, int NBUF, int SA, int SW, bool FUSED, bool WRES, typename Tp>
static void launch_halo_t(const TcHaloPlan* plan, cudaStream_t st) {
    auto kern = halo_tc_kernel<N, T, NBUF, SA, SW, FUSED, WRES, Tp>;
    const size_t smem = (size_t)SA * T * kHaloBufBytes + (size_t)SW * N * 128 + (FUSED ? (size_t)NBUF * T * (N / 64) * 16384 : 0) + 1024;
    static bool configured = false;       // per instantiation
    if (!configured) {
        SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    kern<<<plan->grid, kHaloThreads, smem, st>>>(plan->params);
    SPB_CHECK_LAUNCH();
    if (plan->params.stats) {
        SPB_CUDA(cudaStreamSynchronize(st));
        std::vector<unsigned long long> h((size_t)plan->grid * 16);
        SPB_CUDA(cudaMemcpy(h.data(), plan->params.stats, h.size() * 8, cudaMemcpyDeviceToHost));
        double a[16] = {0};
        for (int c = 0; c < plan->grid; ++c)
            for (int i = 0; i < 16; ++i) a[i] += (double)h[(size_t)c * 16 + i] / plan->grid;
        const HaloParams& q = plan->params;
        const double nloc = (double)q.n_super / plan->grid;
        std::fprintf(stderr,
                     "[halo stats] N=%d T=%d steps=%d supertiles/CTA=%.1f | per super-tile (cycles): mma total %.0f = wait acc-free %.0f + a_full %.0f + "
                     "w_full(g1) %.0f + y_full %.0f + w_full(g2) %.0f + issue %.0f | producers wait: a_empty %.0f w_empty %.0f | epilogue: "
                     "total %.0f wait d1_full %.0f y_empty %.0f d2_full %.0f\n",
                     N, T, q.nsteps, nloc, a[0] / nloc, a[1] / nloc, a[2] / nloc, a[3] / nloc, a[4] / nloc, a[5] / nloc,
                     (a[0] - a[1] - a[2] - a[3] - a[4] - a[5]) / nloc, a[6] / nloc, a[7] / nloc, a[11] / nloc, a[8] / nloc, a[9] / nloc,
                     a[10] / nloc);
    }
}

constexpr int kHaloResidentSlabs = 11;

template <typename Tp>
static void launch_halo_v(const TcHaloPlan* plan, cudaStream_t st) {
    switch (plan->variant) {
        case 0: launch_halo_t<64, 1, 2, 3, kHaloResidentSlabs, true, true, Tp>(plan, st); break;
        case 1: launch_halo_t<128, 2, 1, 2, 4, true, false, Tp>(plan, st); break;
        case 2: launch_halo_t<128, 2, 2, 2, 6, false, false, Tp>(plan, st); break;
        default: throw std::invalid_argument("tcgen05 halo block: bad variant");
    }
}

void launch_halo_tc(const TcHaloPlan* plan, cudaStream_t st) {
    if (!plan) throw std::runtime_error("tcgen05 halo block: no plan");
    if (plan->operand_type == PREC_FP16) launch_halo_v<__half>(plan, st);
    else launch_halo_v<__nv_bfloat16>(plan, st);
}

void tc_halo_plan_destroy(TcHaloPlan* plan) {
    if (plan && plan->params.stats) cudaFree(plan->params.stats);
    delete plan;
}

// Returns nullptr when the block does not fit this kernel (stride 2, 256 channels, ...): the caller then
// uses the per-tap kernel of block_tc.cu.  Arguments as tc_block_plan_create.
TcHaloPlan* tc_halo_plan_create(const ConvDev& c1, const ConvDev* c2, int operand_type, int real_cout, int num_sms) {
    if (operand_type != PREC_FP16 && operand_type != PREC_BF16) return nullptr;
    const int N = c1.cout_pad;
    if (N != 64 && N != 128) return nullptr;
    if (N == 64 && !c2) return nullptr;
    int min_dy = 127, max_dy = -127, min_dx = 127, max_dx = -127;
    for (int s = 0; s < c1.nseg; ++s) {
        const SegDev& sg = c1.seg[s];
        if (sg.stride != 1 || sg.cin % 64 != 0 || sg.C % 64 != 0 || sg.cin != sg.C) return nullptr;
        if (sg.H != c1.OH || sg.W != c1.OW) return nullptr;
        for (int t = 0; t < sg.ntaps; ++t) {
            min_dy = std::min<int>(min_dy, sg.dy[t]); max_dy = std::max<int>(max_dy, sg.dy[t]);
            min_dx = std::min<int>(min_dx, sg.dx[t]); max_dx = std::max<int>(max_dx, sg.dx[t]);
        }
    }
    if (c2) { min_dy = std::min(min_dy, 0); max_dy = std::max(max_dy, 0); min_dx = std::min(min_dx, 0); max_dx = std::max(max_dx, 0); }
    if (max_dy - min_dy > 2 || max_dx - min_dx > 2) return nullptr;
    const ConvDev& last = c2 ? *c2 : c1;
    if (last.dst_C % 8 != 0 || (last.residual && last.res_C % 8 != 0)) return nullptr;
    if (c2 && (c2->cout_pad != N || c2->OH != c1.OH || c2->OW != c1.OW || c2->dst_stride != 1)) return nullptr;

    auto plan = std::make_unique<TcHaloPlan>();
    HaloParams& p = plan->params;
    std::memset(&p, 0, sizeof(p));
    const CUtensorMapDataType dt = operand_type == PREC_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    plan->operand_type = operand_type;
    plan->variant = N == 64 ? 0 : (c2 ? 1 : 2);
    const int T = N == 64 ? 1 : 2;

    // tile orientation: 16 x 8 (group axis = x) or 8 x 16 (group axis = y), whichever needs fewer tiles
    const long n0 = (long)((c1.OH + 15) / 16) * ((c1.OW + 7) / 8), n1 = (long)((c1.OH + 7) / 8) * ((c1.OW + 15) / 16);
    const char* force = std::getenv("SPB200_HALO_ORIENT");
    p.orient = force ? (force[0] == '1') : (n1 < n0 ? 1 : 0);
    const int ext_g = p.orient == 0 ? c1.OW : c1.OH, ext_s = p.orient == 0 ? c1.OH : c1.OW;
    p.tiles_g = (ext_g + 7) / 8;
    p.tiles_per_img = p.tiles_g * ((ext_s + 15) / 16);
    p.total_tiles = p.tiles_per_img * c1.B;
    p.n_super = (p.total_tiles + T - 1) / T;
    plan->grid = std::min(p.n_super, num_sms);
    p.lo_s = p.orient == 0 ? min_dy : min_dx;
    p.lo_g = p.orient == 0 ? min_dx : min_dy;
    auto view16 = [&](int dy, int dx) {
        const int vs = (p.orient == 0 ? dy : dx) - p.lo_s, vg = (p.orient == 0 ? dx : dy) - p.lo_g;
        return (uint16_t)((vs * kHaloG + vg) * 128 / 16);
    };

    p.n_mma = std::min(N, (real_cout + 31) / 32 * 32);
    auto kk_of = [](int real_c, int nchunks) {
        const int in_last = real_c - (nchunks - 1) * 64;
        return std::max(1, std::min(4, (in_last + 15) / 16));
    };
    p.has_ds = c2 && c2->nseg > 1;
    if (c2) {
        if (c2->nseg < 1 || c2->seg[0].ntaps != 1 || c2->seg[0].cin != N) return nullptr;
        if (p.has_ds && c2->nseg - 1 != c1.nseg) return nullptr;
    }

    int nsteps = 0, nchunks = 0;
    for (int s = 0; s < c1.nseg; ++s) {
        const SegDev& sg = c1.seg[s];
        const int nch = sg.cin / 64;
        const int kk_last = kk_of(sg.cin_real > 0 ? sg.cin_real : sg.cin, nch);
        if (sg.koff % 64) return nullptr;
        const SegDev* ds = nullptr;
        if (p.has_ds) {
            ds = &c2->seg[s + 1];
            if (ds->src != sg.src || ds->ntaps != 1 || ds->dy[0] != 0 || ds->dx[0] != 0 || ds->stride != 1 || ds->cin != sg.cin ||
                ds->koff % 64)
                return nullptr;
        }
        for (int c = 0; c < nch; ++c) {
            if (nchunks >= kMaxChunks) return nullptr;
            p.chunk_seg[nchunks] = s;
            p.chunk_c0[nchunks] = c * 64;
            ++nchunks;
            const int nkk = c == nch - 1 ? kk_last : 4;
            for (int t = 0; t < sg.ntaps; ++t) {
                if (nsteps >= kMaxSteps) return nullptr;
                HaloStep& e = p.steps[nsteps++];
                e.kcoord = (uint16_t)(sg.koff / 64 + t * nch + c);
                e.a_off16 = view16(sg.dy[t], sg.dx[t]);
                e.gemm = 0; e.nkk = (uint8_t)nkk;
                e.flags = (uint8_t)((t == 0 ? 1 : 0) | ((t == sg.ntaps - 1 && !ds) ? 2 : 0));
            }
            if (ds) {
                if (nsteps >= kMaxSteps) return nullptr;
                HaloStep& e = p.steps[nsteps++];
                e.kcoord = (uint16_t)(ds->koff / 64 + c);
                e.a_off16 = view16(0, 0);
                e.gemm = 1; e.nkk = (uint8_t)nkk; e.flags = 2;
            }
        }
        // activations: dims {C, group axis, slow axis, image}
        const cuuint64_t C = sg.C, W = sg.W, H = sg.H;
        cuuint32_t box[4] = {64, (cuuint32_t)kHaloG, (cuuint32_t)kHaloS, 1};
        if (p.orient == 0) {
            cuuint64_t dims[4] = {C, W, H, (cuuint64_t)c1.B};
            cuuint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
            tc_encode_tiled(&p.tmA[s], dt, 4, sg.src, dims, str, box);
        } else {
            cuuint64_t dims[4] = {C, H, W, (cuuint64_t)c1.B};
            cuuint64_t str[3] = {W * C * 2, C * 2, H * W * C * 2};
            tc_encode_tiled(&p.tmA[s], dt, 4, sg.src, dims, str, box);
        }
    }
    for (int s = c1.nseg; s < kMaxSegs; ++s) p.tmA[s] = p.tmA[0];
    p.nchunks = nchunks;
    p.n1steps = nsteps;
    {
        cuuint64_t dims[2] = {(cuuint64_t)c1.K, (cuuint64_t)N};
        cuuint64_t str[1] = {(cuuint64_t)c1.K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)N};
        tc_encode_tiled(&p.tmW1, dt, 2, c1.w, dims, str, box);
    }
    p.bias1 = c1.bias;
    if (c2) {
        const int ych = N / 64;
        const int y_kk_last = kk_of(real_cout, ych);
        if (c2->seg[0].koff != 0) return nullptr;
        for (int c = 0; c < ych; ++c) {
            if (nsteps >= kMaxSteps) return nullptr;
            HaloStep& e = p.steps[nsteps++];
            e.kcoord = (uint16_t)c;
            e.a_off16 = (uint16_t)(c * 16384 / 16);
            e.gemm = 2; e.nkk = (uint8_t)(c == ych - 1 ? y_kk_last : 4); e.flags = 0;
        }
        cuuint64_t dims[2] = {(cuuint64_t)c2->K, (cuuint64_t)N};
        cuuint64_t str[1] = {(cuuint64_t)c2->K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)N};
        tc_encode_tiled(&p.tmW2, dt, 2, c2->w, dims, str, box);
        p.bias2 = c2->bias;
    } else {
        p.tmW2 = p.tmW1;
    }
    p.nsteps = nsteps;
    if (plan->variant == 0 && nsteps > kHaloResidentSlabs) return nullptr;
    p.OH = c1.OH; p.OW = c1.OW;
    p.residual = last.residual; p.dst = last.dst;
    p.res_C = last.res_C; p.dst_H = last.dst_H; p.dst_W = last.dst_W; p.dst_C = last.dst_C;
    p.dst_stride = last.dst_stride; p.dst_off_y = last.dst_off_y; p.dst_off_x = last.dst_off_x;
    p.relu = last.relu; p.dst_fp32 = last.dst_fp32;
    const char* se = std::getenv("SPB200_HALO_STATS");
    if (se && se[0] == '1') {
        SPB_CUDA(cudaMalloc((void**)&p.stats, (size_t)plan->grid * 16 * sizeof(unsigned long long)));
        SPB_CUDA(cudaMemset(p.stats, 0, (size_t)plan->grid * 16 * sizeof(unsigned long long)));
    }
    return plan.release();
}

}  // namespace spb200
