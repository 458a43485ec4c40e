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
// Warp roles (384 threads, one persistent CTA per SM): warp 0 = activation TMA producer, warp 1 = weight
// TMA producer, warps 2-3 = MMA issuers (one per tile of the pair; warp 2 also owns the TMEM allocation),
// warps 4-11 = epilogue (two warps per TMEM lane quarter, one per tile).  With NBUF = 2 (transposed-conv
// phases) the accumulator is double buffered, so the next tile pair runs under the epilogue of this one.
//
// The output leaves through shared memory: an epilogue warp owns 32 pixel rows of its tile (8 pixels along the
// group axis x 4 slow rows), writes them as 128-byte swizzled rows into its 4 KB share of a staging box and sends
// the sub-box with one TMA store per 64 channels (per 32 for fp32 output), clipped at the image border by the TMA
// unit; an identity shortcut is fetched into the same sub-boxes by TMA as soon as the previous tile's stores have
// read them and is overwritten in place.  No cross-warp synchronisation is involved.  Global loads and stores
// issued from the accumulator layout (one pixel row per lane, 32 lines per instruction) go through the same
// L1 / shared-memory pipe as the MMAs' operand reads and were measured to stretch a tile pair of the 128-channel
// block from 12.3 k to 18.3 k cycles.
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
constexpr int kMaxSteps = 96;
constexpr int kMaxChunks = 8;
__host__ __device__ constexpr int halo_threads(int T) { return (2 + T + 8) * 32; }
__host__ __device__ constexpr int halo_out_slots(int N, bool split) { return (N == 128 || split) ? 2 : 1; }
constexpr int kHaloOutBox = 16384;                          // one staged output box: 128 pixel rows x 128 B   // 2 producers, T MMA issuers, 8 epilogue warps

// One weight slab [N x 64 K] and the MMAs that consume it, packed into 64 bits so that the issuing warps fetch a
// step with one constant load and a few bit-field extracts (their instruction count is the bottleneck):
//   bits  0-15  a_lo    descriptor-low offset (bytes / 16) of the A view: tap view in the haloed box, or Y chunk
//   bits 16-31  w_lo    descriptor-low offset of the slab in the weight ring (slot * N * 128 / 16)
//   bits 32-35  slot    weight ring slot (static: step index modulo the ring size)
//   bits 36-37  gemm    0: 3x3 taps -> D1;  1: 1x1 shortcut (centre view) -> D2;  2: 1x1 over Y -> D2
//   bits 38-40  nkk     K = 16 MMA steps in this slab (1..4)
//   bit  41     first slab of an activation chunk (wait for it);  bit 42  last slab using it (release it)
//   bit  43     acc0    accumulate flag of the first MMA of the slab
//   bits 44-45  astep   what the A view advances by per K = 16 step: 0 = 32 B (the next 16 channels of the same pixels),
//                       1 = one pixel along the group axis, 2 = one haloed row - the packed K tail (common.cuh, SegDev): the
//                       K steps of such a slab are the taps of one filter row over the chunk's first 16 channels
//   bits 48-59  kcoord  K coordinate / 64 in the weight tensor (W1 for gemm 0, W2 otherwise)
//   bit  60     d1done  alternating issue: the warp that issues this slab has issued its last GEMM-1 MMA of the tile pair with it and
//                       commits D1 right after it - the shortcut slab that follows runs while the first epilogue already reads D1
using HaloStep = unsigned long long;
__host__ __device__ constexpr HaloStep halo_step(unsigned a_lo, unsigned w_lo, unsigned slot, unsigned gemm, unsigned nkk,
                                                 unsigned first, unsigned last, unsigned acc0, unsigned kcoord, unsigned astep = 0) {
    return (HaloStep)a_lo | ((HaloStep)w_lo << 16) | ((HaloStep)slot << 32) | ((HaloStep)gemm << 36) | ((HaloStep)nkk << 38) |
           ((HaloStep)first << 41) | ((HaloStep)last << 42) | ((HaloStep)acc0 << 43) | ((HaloStep)astep << 44) | ((HaloStep)kcoord << 48);
}

struct HaloParams {
    CUtensorMap tmA[kMaxSegs];
    CUtensorMap tmW1, tmW2;
    CUtensorMap tmR;               // identity-shortcut operand, fetched into the same staging boxes
    CUtensorMap tmD;               // destination, for the TMA stores of the fused variants (box = 32 pixels x 128 B)
    HaloStep steps[kMaxSteps + 1];   // one spare entry: the issuing loop prefetches step e + 1
    int nsteps, n1steps;           // all slabs; slabs of GEMM 1 + shortcut (they come first)
    int first_ds_step;             // first shortcut slab (n1steps when there is none)
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
    int w_bytes;                   // bytes of one weight slab as loaded: n_mma rows x 128 B
    int tma_res;                   // the identity shortcut arrives by TMA (needs tma_store)
    int tma_store;                 // fused variants: the output leaves through shared memory and TMA stores
    float* heat_inv;               // [B][OH][OW]: 1 / (sum_c exp(l_c) + 1e-5) per cell
    int heat;                      // detector tail (variant 1): the epilogue stores exp(l_c), c < 64, depth-to-space into the full-resolution
                                   // map (tmD is then a map of it) and the cell's normaliser into heat_inv: heat = exp * inv
    int split_out;                 // SPLIT kernels: the output is stored as [hi 32 | lo 32] per 32 channels (common.cuh, SegDev)
    int alt_issue;                 // streamed-weight variants: the two issuing warps take alternate steps, each for both tiles (below)
    int y_early;                   // with alt_issue: GEMM 2 starts on the first 64 channels of Y while the first epilogue converts the rest
    int early_d1;                  // with alt_issue: D1 is committed per warp at its d1done slab instead of after the last slab of GEMM 1 + shortcut
    long long* dbg;                // SPB200_HALO_DBG: clock stamps of CTA 0 (scripts/halo_dbg.py)
};

// ---- CTA pairs (cta_group::2): two CTAs of a cluster, one issuing thread in the leader, M = 256 MMAs ---------------------
__device__ __forceinline__ uint32_t pair_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void pair_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pair_leader_addr(uint32_t addr) {       // the same shared-memory offset in CTA 0 of the pair
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(0));
    return r;
}
__device__ __forceinline__ void pair_arrive_leader(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void pair_tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void pair_tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2,
                                                 int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void pair_tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void pair_tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void pair_mma(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                         uint32_t accumulate) {
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
__device__ __forceinline__ void pair_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
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
// arrive on the barrier at this offset in BOTH CTAs when the MMAs issued so far by this thread are complete
__device__ __forceinline__ void pair_commit_a(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// SPLIT: split-precision operands (three MMAs per product, common.cuh SegDev): the activation chunks and the weight slabs
// arrive in the split layout - which only the host-built step list knows about -, the first epilogue writes Y back as
// [hi | lo] and the second one can store the block output the same way.
//
// PAIR (launched as clusters of two CTAs): the resident-weight kernel on tcgen05.mma.cta_group::2.  Each CTA keeps what it had -
// its own tile pairs, activation stages, accumulators, epilogue warps and staging boxes - but holds only HALF of the rows of
// every weight slab, and the MMAs of both CTAs are issued by the two issuing warps of the LEADER as M = 256 instructions
// (the leader's tile in its tensor memory, the peer's tile in the peer's).  An MMA then reads 4 KB of A + 1 KB of B per CTA
// from shared memory instead of 4 + 2 KB: the operand fetch that caps the N = 64 blocks drops from 192 to 160 B/clk.
//   a_full, w_full    leader's barriers; its producers announce the bytes of both CTAs, both CTAs' TMA loads complete on them
//   a_empty, d1_full, d2_full    local in both CTAs, reached by multicast tcgen05.commit
//   y_full, d2_empty  leader's, counted over the epilogue warps of both CTAs (the peer's arrive through mapa)
template <int N, int T, int NBUF, int SA, int SW, bool FUSED, bool WRES, bool SPLIT, bool PAIR, typename Tp>
__global__ void __launch_bounds__(halo_threads(T), 1) halo_tc_kernel(const __grid_constant__ HaloParams p) {
    static_assert(!PAIR || (FUSED && WRES && !SPLIT && NBUF == 2), "CTA pairs: the resident-weight fused kernel");
    constexpr int kWBytes = N * 128;
    constexpr int kWSlot = PAIR ? kWBytes / 2 : kWBytes;      // bytes of a weight slab held by this CTA
    constexpr uint32_t kAccCols = NBUF * T * N;
    constexpr uint32_t kTmemCols = (FUSED ? 2 : 1) * kAccCols;
    static_assert(kTmemCols == 64 || kTmemCols == 128 || kTmemCols == 256 || kTmemCols == 512, "TMEM columns");
    static_assert(!(FUSED && NBUF == 2) || WRES, "pipelined GEMM 2 needs resident weights (slab order)");
    static_assert(T == 2 || !FUSED, "the in-place Y write-back assumes one epilogue warp per lane quarter and tile");
    static_assert(!SPLIT || (FUSED && !WRES), "split precision: fused blocks with streamed weights");

    extern __shared__ uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t a_full[SA], a_empty[SA], w_full[SW], w_empty[SW];
    __shared__ __align__(8) uint64_t d1_full[NBUF], d1_empty[NBUF], y_full[NBUF], d2_full[NBUF], d2_empty[NBUF];
    __shared__ __align__(8) uint64_t y_half[NBUF];         // y_early: the first 64 channels of Y (both tiles) are in tensor memory
    __shared__ __align__(8) uint64_t res_full[8];          // one per epilogue warp: its shortcut sub-boxes have landed
    __shared__ __align__(8) uint64_t tok[2];               // alternating issue: "the step before yours has been issued"
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_bias1[N], s_bias2[N];

    // aligned up to 1024 B by pointer ARITHMETIC on dyn_smem: the compiler keeps the shared address space (LDS/STS,
    // 32-bit addresses) instead of falling back to generic loads
    uint8_t* a_ring = dyn_smem + ((1024u - (smem_u32(dyn_smem) & 1023u)) & 1023u);
    uint8_t* w_ring = a_ring + SA * T * kHaloBufBytes;
    // output staging of the fused variants: per tile kOutSlots boxes of 128 pixel rows x 128 B (16 KB), each epilogue warp
    // owns the 4 KB of its 32 rows in every box
    constexpr int kOutSlots = halo_out_slots(N, SPLIT);
    uint8_t* out_stage = w_ring + SW * kWBytes;
    // warp index through a shuffle: tells the compiler it is warp-uniform, so the role loops below run on the
    // uniform datapath (loop counters, ring state, descriptors in uniform registers) instead of R2UR round trips
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x / 32), 0), lane = threadIdx.x % 32;
    const uint32_t idesc = (1u << 4) | (OperandFmt<Tp>::value << 7) | (OperandFmt<Tp>::value << 10) |
                           ((uint32_t)(p.n_mma >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
    // scheduling units: CTAs, or CTA pairs (both CTAs of a pair run the same number of steps; rank r takes super-tile 2 u + r)
    const int rank = PAIR ? (int)pair_rank() : 0;
    const int unit = PAIR ? (int)blockIdx.x >> 1 : (int)blockIdx.x, nunits = PAIR ? (int)gridDim.x >> 1 : (int)gridDim.x;
    const int n_work = PAIR ? (p.n_super + 1) / 2 : p.n_super;
    const int n_local = (n_work - unit + nunits - 1) / nunits;
    auto st_of = [&](int j) { return PAIR ? (unit + j * nunits) * 2 + rank : unit + j * nunits; };
    pdl_trigger();

    if (threadIdx.x == 0) {
        constexpr int kEpi = PAIR ? 16 : 8;                  // epilogue warps that arrive on the issuer's barriers
        for (int s = 0; s < SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], T); }
        for (int s = 0; s < SW; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], T); }
        for (int w = 0; w < 8; ++w) mbar_init(&res_full[w], 1);
        mbar_init(&tok[0], 1); mbar_init(&tok[1], 1);
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(&d1_full[b], T); mbar_init(&d2_full[b], T);
            mbar_init(&d1_empty[b], 8); mbar_init(&y_full[b], kEpi); mbar_init(&d2_empty[b], kEpi);
            mbar_init(&y_half[b], kEpi);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kMaxSegs; ++s) prefetch_tmap(&p.tmA[s]);
        prefetch_tmap(&p.tmW1);
        if (FUSED) prefetch_tmap(&p.tmW2);
    }
    if (warp == 2) { if (PAIR) pair_tmem_alloc(&tmem_slot, kTmemCols); else tmem_alloc(&tmem_slot, kTmemCols); }
    for (int i = threadIdx.x; i < T * kOutSlots * kHaloOutBox / 16; i += halo_threads(T))
        reinterpret_cast<uint4*>(out_stage)[i] = make_uint4(0u, 0u, 0u, 0u);      // channels beyond n_mma leave as zeros
    if (kOutSlots > 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int i = threadIdx.x; i < N; i += halo_threads(T)) {
        s_bias1[i] = p.bias1[i];
        s_bias2[i] = FUSED ? p.bias2[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) pair_sync();                                   // the peer's barriers exist before anything remote touches them
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    // The three issuing roles run as WHOLE warps: loop counters, ring indices and phases are warp-uniform
    // (they live in uniform registers), every lane polls the mbarriers, and only the instructions that must come
    // from one thread (TMA, tcgen05.mma, tcgen05.commit, expect_tx) sit under elect.sync.  Issuing from inside an
    // `if (lane == 0)` region instead costs ~200 cycles per tcgen05.mma (measured): the compiler has to move
    // every descriptor through R2UR and wrap each instruction in an active-lane loop.
    if (warp == 0) {
        // ============================ activation producer ============================
        pdl_wait();                                            // the activations are the previous kernel's output
        int sa = 0;
        uint32_t pha = 0;
        for (int j = 0; j < n_local; ++j) {
            const int st = st_of(j);
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
                mbar_wait(&a_empty[sa], pha ^ 1u);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(&a_full[sa], (uint32_t)((PAIR ? 2 : 1) * T * kHaloLoadBytes));
                    const CUtensorMap* tm = &p.tmA[p.chunk_seg[ci]];
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        if (PAIR) pair_tma_load_4d(smem_u32(a_ring + (sa * T + t) * kHaloBufBytes), tm, pair_leader_addr(smem_u32(&a_full[sa])),
                                                   p.chunk_c0[ci], cg[t], cs[t], cimg[t]);
                        else tma_load_4d(a_ring + (sa * T + t) * kHaloBufBytes, tm, &a_full[sa], p.chunk_c0[ci], cg[t], cs[t], cimg[t]);
                    }
                }
                if (++sa == SA) { sa = 0; pha ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ============================ weight producer ============================
        uint32_t eph = 0;                                      // per-slot parity of the next w_empty wait
        const uint32_t bar_full = smem_u32(&w_full[0]), bar_empty = smem_u32(&w_empty[0]), ring = smem_u32(w_ring);
        const int passes = WRES ? min(n_local, 1) : n_local;
        for (int j = 0; j < passes; ++j) {
            for (int e = 0; e < p.nsteps; ++e) {
                const HaloStep s = p.steps[e];
                const uint32_t hi = (uint32_t)(s >> 32), slot = hi & 15u;
                if (!WRES) { mbar_wait_a(bar_empty + slot * 8, ((eph >> slot) & 1u) ^ 1u); eph ^= 1u << slot; }
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx_a(bar_full + slot * 8, (uint32_t)p.w_bytes);     // both halves in a pair
                    const CUtensorMap* tw = ((hi >> 4) & 3u) == 0 ? &p.tmW1 : &p.tmW2;
                    if (PAIR) pair_tma_load_2d(ring + slot * kWSlot, tw, pair_leader_addr(bar_full + slot * 8), (int)((hi >> 16) & 0xfffu) * 64,
                                               rank * (p.n_mma >> 1));
                    else tma_load_2d_a(ring + slot * kWBytes, tw, bar_full + slot * 8, (int)((hi >> 16) & 0xfffu) * 64, 0);
                }
            }
        }
    } else if (warp < 2 + T && !WRES && !PAIR && T == 2 && p.alt_issue) {
        // ============================ MMA issuers, alternating steps ============================
        // With one issuing warp per tile both warps wait for the same slab, issue their four MMAs side by side and then
        // walk through the same per-step overhead (commit, step decode, barrier polls: ~180 clocks) at the same time - and
        // the tensor pipe, whose queue holds about two MMAs, runs dry once per step (step traces: 660-710 clocks per step
        // of eight MMAs where the pipe needs 384-512).  Here the warps take ALTERNATE steps, each issuing the step's MMAs
        // for both tiles: while one issues, the other is through its overhead and waiting for the token "the step before
        // yours has been issued" (an mbarrier handed back and forth), so the order of the MMAs on every accumulator - and
        // with it every result bit - is the one of the step list, as before.
        const int w = warp - 2;
        int sa = 0;
        uint32_t pha = 0, wph = 0, gpar = 0, tokph = 0;
        bool first_step = true;                                  // this warp has not issued yet (warp 0: no token before step 0)
        const uint32_t a_lo_base = umma_desc_lo(smem_u32(a_ring));
        const uint32_t w_lo_base = umma_desc_lo(smem_u32(w_ring));
        const uint32_t bar_wfull = smem_u32(&w_full[0]), bar_wempty = smem_u32(&w_empty[0]);
        const uint32_t bar_afull = smem_u32(&a_full[0]), bar_aempty = smem_u32(&a_empty[0]);
        constexpr uint32_t kHiA = ((uint32_t)kHaloSbo >> 4) | (1u << 14) | (2u << 29);
        constexpr uint32_t kHiB = (1024u >> 4) | (1u << 14) | (2u << 29);
        constexpr uint32_t kTileA = (uint32_t)(kHaloBufBytes >> 4);
        auto take_token = [&]() {
            if (!(first_step && w == 0)) { mbar_wait(&tok[w], tokph); tokph ^= 1u; }
            first_step = false;
        };
        for (int j = 0; j < n_local; ++j) {
            const bool dbg_on = p.dbg && blockIdx.x == 0 && w == 0 && lane == 0 && j < 16;
            if (dbg_on) p.dbg[j * 8 + 0] = clock64();
            const int b = j % NBUF;
            const uint32_t ph = (uint32_t)(j / NBUF) & 1u;
            if (!FUSED) mbar_wait(&d1_empty[b], ph ^ 1u);              // epilogue has drained D1[b]
            bool d2_ready = !(FUSED && p.has_ds);
            HaloStep rec = p.steps[0];
            for (int e = 0; e < p.n1steps; ++e) {
                const HaloStep s = rec;
                rec = p.steps[e + 1];
                const uint32_t lo = (uint32_t)s, hi = (uint32_t)(s >> 32);
                const uint32_t slot = hi & 15u, nkk = (hi >> 6) & 7u, acc0 = (hi >> 11) & 1u;
                if (gpar == (uint32_t)w) {
                    if (hi & (1u << 9)) mbar_wait_a(bar_afull + sa * 8, pha);
                    mbar_wait_a(bar_wfull + slot * 8, (wph >> slot) & 1u);
                    if (!d2_ready && ((hi >> 4) & 3u) != 0) mbar_wait(&d2_empty[b], ph ^ 1u);
                    take_token();
                    tc_fence_after();
                    const uint32_t alo = a_lo_base + (uint32_t)(sa * T) * kTileA + (lo & 0xffffu);
                    const uint32_t blo = w_lo_base + (lo >> 16);
                    const uint32_t d = tmem_base + (((hi >> 4) & 3u) == 0 ? 0u : kAccCols) + (uint32_t)(b * T * N);
                    const uint32_t asel = (hi >> 12) & 3u;
                    const uint32_t as = asel == 0 ? 2u : (asel == 1 ? 8u : (uint32_t)(kHaloSbo >> 4));
                    if (elect_one()) {
                        umma_f16_w(d, alo, kHiA, blo, kHiB, idesc, acc0);
                        umma_f16_w(d + N, alo + kTileA, kHiA, blo, kHiB, idesc, acc0);
                        if (nkk > 1) { umma_f16_w(d, alo + as, kHiA, blo + 2, kHiB, idesc, 1u); umma_f16_w(d + N, alo + kTileA + as, kHiA, blo + 2, kHiB, idesc, 1u); }
                        if (nkk > 2) { umma_f16_w(d, alo + 2 * as, kHiA, blo + 4, kHiB, idesc, 1u); umma_f16_w(d + N, alo + kTileA + 2 * as, kHiA, blo + 4, kHiB, idesc, 1u); }
                        if (nkk > 3) { umma_f16_w(d, alo + 3 * as, kHiA, blo + 6, kHiB, idesc, 1u); umma_f16_w(d + N, alo + kTileA + 3 * as, kHiA, blo + 6, kHiB, idesc, 1u); }
                        umma_commit_a(bar_wempty + slot * 8);               // the barriers count one arrival per tile
                        umma_commit_a(bar_wempty + slot * 8);
                        if (hi & (1u << 10)) { umma_commit_a(bar_aempty + sa * 8); umma_commit_a(bar_aempty + sa * 8); }
                        if (hi & (1u << 28)) umma_commit_a(smem_u32(&d1_full[b]));   // d1done: this warp's share of D1 is on its way
                        mbar_arrive(&tok[1 - w]);
                    }
                }
                if (!d2_ready && ((hi >> 4) & 3u) != 0) d2_ready = true;
                wph ^= 1u << slot;
                gpar ^= 1u;
                if (hi & (1u << 10)) { if (++sa == SA) { sa = 0; pha ^= 1u; } }
            }
            if (!p.early_d1 && elect_one()) umma_commit_a(smem_u32(&d1_full[b]));           // each warp for the MMAs it issued
            if (dbg_on) p.dbg[j * 8 + 1] = clock64();
            if (FUSED) {
                // y_early: the first slab of GEMM 2 reads the first 64 channels of Y only (packed columns 0-31, complete after the
                // first epilogue's second block); its eight MMAs run while the epilogue converts the remaining blocks
                mbar_wait(p.y_early ? &y_half[b] : &y_full[b], ph);                 // Y written by the epilogue warps
                if (!p.has_ds) mbar_wait(&d2_empty[b], ph ^ 1u);
                tc_fence_after();
                if (dbg_on) p.dbg[j * 8 + 2] = clock64();
                for (int e = p.n1steps; e < p.nsteps; ++e) {
                    if (p.y_early && e == p.n1steps + 1) { mbar_wait(&y_full[b], ph); tc_fence_after(); }
                    const HaloStep s = p.steps[e];
                    const uint32_t lo = (uint32_t)s, hi = (uint32_t)(s >> 32);
                    const uint32_t slot = hi & 15u, nkk = (hi >> 6) & 7u, acc0 = (hi >> 11) & 1u;
                    if (gpar == (uint32_t)w) {
                        mbar_wait_a(bar_wfull + slot * 8, (wph >> slot) & 1u);
                        take_token();
                        tc_fence_after();
                        const uint32_t ya = tmem_base + (uint32_t)(b * T * N) + (lo & 0xffffu);
                        const uint32_t blo = w_lo_base + (lo >> 16);
                        const uint32_t d = tmem_base + kAccCols + (uint32_t)(b * T * N);
                        if (elect_one()) {
                            umma_f16_ts(d, ya, blo, kHiB, idesc, acc0);
                            umma_f16_ts(d + N, ya + N, blo, kHiB, idesc, acc0);
                            if (nkk > 1) { umma_f16_ts(d, ya + 8, blo + 2, kHiB, idesc, 1u); umma_f16_ts(d + N, ya + N + 8, blo + 2, kHiB, idesc, 1u); }
                            if (nkk > 2) { umma_f16_ts(d, ya + 16, blo + 4, kHiB, idesc, 1u); umma_f16_ts(d + N, ya + N + 16, blo + 4, kHiB, idesc, 1u); }
                            if (nkk > 3) { umma_f16_ts(d, ya + 24, blo + 6, kHiB, idesc, 1u); umma_f16_ts(d + N, ya + N + 24, blo + 6, kHiB, idesc, 1u); }
                            umma_commit_a(bar_wempty + slot * 8);
                            umma_commit_a(bar_wempty + slot * 8);
                            mbar_arrive(&tok[1 - w]);
                        }
                    }
                    wph ^= 1u << slot;
                    gpar ^= 1u;
                }
                if (elect_one()) umma_commit_a(smem_u32(&d2_full[b]));
                if (dbg_on) p.dbg[j * 8 + 3] = clock64();
            }
        }
        __syncwarp();
    } else if (warp < 2 + T) {
        // ============================ MMA issuers: one warp per tile ============================
        // Non-MMA instructions of an issuing warp are not hidden behind the tensor pipe (measured: every branch, wait
        // or commit between two tcgen05.mma adds its latency), so (a) each tile of the pair has its own issuing warp -
        // while one polls a barrier or commits, the other feeds the pipe with independent MMAs - and (b) a step costs
        // one 64-bit constant load (prefetched), bit-field extracts, one wait, four MMAs whose descriptors differ by
        // immediates, and one commit; ring slots are static per step, their phases are a bit mask.
        const int mt = warp - 2;
        int sa = 0;
        uint32_t pha = 0, wph = 0;
        const uint32_t a_lo_base = umma_desc_lo(smem_u32(a_ring) + mt * kHaloBufBytes);
        const uint32_t w_lo_base = umma_desc_lo(smem_u32(w_ring));
        const uint32_t d_base = tmem_base + (uint32_t)(mt * N);
        const uint32_t bar_wfull = smem_u32(&w_full[0]), bar_wempty = smem_u32(&w_empty[0]);
        const uint32_t bar_afull = smem_u32(&a_full[0]), bar_aempty = smem_u32(&a_empty[0]);
        constexpr uint32_t kHiA = ((uint32_t)kHaloSbo >> 4) | (1u << 14) | (2u << 29);   // SBO = haloed row pitch
        constexpr uint32_t kHiB = (1024u >> 4) | (1u << 14) | (2u << 29);                // SBO = 1024 (dense tile)
        constexpr int LAG = (FUSED && NBUF == 2) ? 1 : 0;
        // a CTA pair is served by the leader's issuing warps: M = 256 instructions, multicast commits
        auto mma = [&](uint32_t d, uint32_t alo, uint32_t blo, uint32_t acc) {
            if (PAIR) pair_mma(d, alo, kHiA, blo, kHiB, idesc, acc); else umma_f16_w(d, alo, kHiA, blo, kHiB, idesc, acc);
        };
        auto mma_ts = [&](uint32_t d, uint32_t ya, uint32_t blo, uint32_t acc) {
            if (PAIR) pair_mma_ts(d, ya, blo, kHiB, idesc, acc); else umma_f16_ts(d, ya, blo, kHiB, idesc, acc);
        };
        auto commit = [&](uint32_t bar) { if (PAIR) pair_commit_a(bar); else umma_commit_a(bar); };
        auto slab_lo = [&](uint32_t lo, uint32_t hi) {           // descriptor-low of the step's weight slab in this CTA
            return PAIR ? w_lo_base + (hi & 15u) * (uint32_t)(kWSlot >> 4) : w_lo_base + (lo >> 16);
        };
        for (int j = 0; j < (PAIR && rank != 0 ? 0 : n_local + LAG); ++j) {
            const bool dbg_on = p.dbg && blockIdx.x == 0 && mt == 0 && lane == 0 && j < 16;
            if (dbg_on) p.dbg[j * 8 + 0] = clock64();
            if (j < n_local) {
                const int b = j % NBUF;
                const uint32_t ph = (uint32_t)(j / NBUF) & 1u;
                // fused: D1[b] holds Y until GEMM 2 of its previous use has read it, and that GEMM 2 was issued by this
                // warp before this point (MMAs of one thread execute in order): no barrier needed
                if (!FUSED) mbar_wait(&d1_empty[b], ph ^ 1u);          // epilogue has drained D1[b]
                // a block with a shortcut convolution accumulates it into D2 during GEMM 1: D2[b] must have been drained
                // by the epilogue of its previous use - waited for at the first shortcut step, not here, so that the
                // steps before it overlap that epilogue (the resident-weight path issues a whole tile at once and waits here)
                bool d2_ready = !(FUSED && p.has_ds);
                tc_fence_after();
                if (WRES && j > 0) {
                    // resident weights, nothing left to wait for but the activation chunk (a 64-channel block has one
                    // chunk): the tile is issued from elect regions with no barrier traffic between MMAs.  The steps
                    // before the first shortcut step do not touch D2 and go first; D2[b] is drained by the second
                    // epilogue of tile j - 2, which runs AFTER the first epilogue of tile j - 1 - waiting for it up
                    // front cost 1.9 k of the 6.8 k cycles per tile pair (timeline of CTA 0).
                    mbar_wait_a(bar_afull + sa * 8, pha);
                    tc_fence_after();
                    const uint32_t a0 = a_lo_base + (uint32_t)(sa * T * (kHaloBufBytes >> 4));
                    auto issue = [&](int e0, int e1) {
                        for (int e = e0; e < e1; ++e) {
                            const HaloStep s = p.steps[e];
                            const uint32_t lo = (uint32_t)s, hi = (uint32_t)(s >> 32);
                            const uint32_t alo = a0 + (lo & 0xffffu), blo = slab_lo(lo, hi);
                            const uint32_t d = d_base + (((hi >> 4) & 3u) == 0 ? 0u : kAccCols) + (uint32_t)(b * T * N);
                            mma(d, alo, blo, (hi >> 11) & 1u);
                            mma(d, alo + 2, blo + 2, 1u);
                            mma(d, alo + 4, blo + 4, 1u);
                            mma(d, alo + 6, blo + 6, 1u);
                        }
                    };
                    const int e_ds = d2_ready ? p.n1steps : p.first_ds_step;
                    if (elect_one()) issue(0, e_ds);
                    if (e_ds < p.n1steps) {
                        mbar_wait(&d2_empty[b], ph ^ 1u);
                        tc_fence_after();
                        if (elect_one()) issue(e_ds, p.n1steps);
                    }
                    if (elect_one()) {
                        commit(bar_aempty + sa * 8);
                        commit(smem_u32(&d1_full[b]));
                    }
                    d2_ready = true;
                    if (++sa == SA) { sa = 0; pha ^= 1u; }
                }
                HaloStep rec = p.steps[0];
                for (int e = 0; e < ((WRES && j > 0) ? 0 : p.n1steps); ++e) {
                    const HaloStep s = rec;
                    rec = p.steps[e + 1];
                    const uint32_t lo = (uint32_t)s, hi = (uint32_t)(s >> 32);
                    const uint32_t slot = hi & 15u, nkk = (hi >> 6) & 7u, acc0 = (hi >> 11) & 1u;
                    if (hi & (1u << 9)) mbar_wait_a(bar_afull + sa * 8, pha);
                    if (!WRES || j == 0) mbar_wait_a(bar_wfull + slot * 8, (wph >> slot) & 1u);
                    if (!d2_ready && ((hi >> 4) & 3u) != 0) { mbar_wait(&d2_empty[b], ph ^ 1u); d2_ready = true; }
                    tc_fence_after();
                    const uint32_t alo = a_lo_base + (uint32_t)(sa * T * (kHaloBufBytes >> 4)) + (lo & 0xffffu);
                    const uint32_t blo = slab_lo(lo, hi);
                    const uint32_t d = d_base + (((hi >> 4) & 3u) == 0 ? 0u : kAccCols) + (uint32_t)(b * T * N);
                    const uint32_t asel = (hi >> 12) & 3u;
                    const uint32_t as = asel == 0 ? 2u : (asel == 1 ? 8u : (uint32_t)(kHaloSbo >> 4));
                    if (elect_one()) {
                        mma(d, alo, blo, acc0);
                        if (nkk > 1) mma(d, alo + as, blo + 2, 1u);
                        if (nkk > 2) mma(d, alo + 2 * as, blo + 4, 1u);
                        if (nkk > 3) mma(d, alo + 3 * as, blo + 6, 1u);
                        if (!WRES) commit(bar_wempty + slot * 8);
                        if (hi & (1u << 10)) commit(bar_aempty + sa * 8);
                    }
                    if (!WRES) wph ^= 1u << slot;
                    if (hi & (1u << 10)) { if (++sa == SA) { sa = 0; pha ^= 1u; } }
                }
                if (!(WRES && j > 0) && elect_one()) commit(smem_u32(&d1_full[b]));
                if (dbg_on) p.dbg[j * 8 + 1] = clock64();
            }
            if (FUSED && j >= LAG) {
                const int jj = j - LAG;
                const int b = jj % NBUF;
                const uint32_t ph = (uint32_t)(jj / NBUF) & 1u;
                mbar_wait(&y_full[b], ph);                 // Y written by the epilogue warps
                if (!p.has_ds) mbar_wait(&d2_empty[b], ph ^ 1u);
                tc_fence_after();
                if (dbg_on) p.dbg[j * 8 + 2] = clock64();
                for (int e = p.n1steps; e < p.nsteps; ++e) {
                    const HaloStep s = p.steps[e];
                    const uint32_t lo = (uint32_t)s, hi = (uint32_t)(s >> 32);
                    const uint32_t slot = hi & 15u, nkk = (hi >> 6) & 7u, acc0 = (hi >> 11) & 1u;
                    if (!WRES || jj == 0) mbar_wait_a(bar_wfull + slot * 8, (wph >> slot) & 1u);
                    tc_fence_after();
                    // A = Y, 16-bit, packed in place over the first N/2 columns of D1[b] by the epilogue warps
                    const uint32_t ya = d_base + (uint32_t)(b * T * N) + (lo & 0xffffu);
                    const uint32_t blo = slab_lo(lo, hi);
                    const uint32_t d = d_base + kAccCols + (uint32_t)(b * T * N);
                    if (elect_one()) {
                        mma_ts(d, ya, blo, acc0);
                        if (nkk > 1) mma_ts(d, ya + 8, blo + 2, 1u);
                        if (nkk > 2) mma_ts(d, ya + 16, blo + 4, 1u);
                        if (nkk > 3) mma_ts(d, ya + 24, blo + 6, 1u);
                        if (!WRES) commit(bar_wempty + slot * 8);
                    }
                    if (!WRES) wph ^= 1u << slot;
                }
                if (elect_one()) commit(smem_u32(&d2_full[b]));
                if (dbg_on) p.dbg[j * 8 + 3] = clock64();
            }
        }
        __syncwarp();
    } else {
        // ============================ epilogue ============================
        pdl_wait();                                            // shortcut loads and output stores touch the chain's buffers
        const int q = warp & 3;                                // TMEM lane quarter this warp may read
        const int eh = (warp - (2 + T)) >> 2;                        // 0/1: tile (T = 2) or column half (T = 1)
        const int row = q * 32 + lane;                         // GEMM row = TMEM lane
        const int ti = row >> 3, tr = row & 7;                 // slow row, pixel within the group
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int t = T == 2 ? eh : 0;
        const int nblk = (p.n_mma + 31) / 32;                  // the last block may be partial (n_mma is a multiple of 16)
        const int blk_lo = T == 2 ? 0 : (eh == 0 ? 0 : (nblk + 1) / 2);
        const int blk_hi = T == 2 ? nblk : (eh == 0 ? (nblk + 1) / 2 : nblk);
        constexpr int kResBlk = N / 32;
        // per-tile coordinates of this thread's output pixel
        struct Pix { bool valid; size_t gpix, dpix; };
        auto pixel_of = [&](int jt) {
            const int st = st_of(jt);
            const int tile_raw = st * T + t;
            const int tile = min(tile_raw, p.total_tiles - 1);
            const int img = tile / p.tiles_per_img, tt = tile % p.tiles_per_img;
            const int s0 = (tt / p.tiles_g) * 16, g0 = (tt % p.tiles_g) * 8;
            const int oy = p.orient == 0 ? s0 + ti : g0 + tr;
            const int ox = p.orient == 0 ? g0 + tr : s0 + ti;
            Pix px;
            px.valid = tile_raw < p.total_tiles && oy < p.OH && ox < p.OW;
            px.gpix = ((size_t)img * p.OH + oy) * p.OW + ox;
            px.dpix = ((size_t)img * p.dst_H + (oy * p.dst_stride + p.dst_off_y)) * p.dst_W + (ox * p.dst_stride + p.dst_off_x);
            return px;
        };
        // epilogue 1: Y = relu(D1 + b1) rounded to 16 bits, written back IN PLACE: block k (fp32 columns 32k..32k+31, all
        // in registers by then) becomes packed columns 16k..16k+15 of the same lanes, which blocks < k no longer need
        auto epi1 = [&](int jt) {
            const int b = jt % NBUF;
            const uint32_t ph = (uint32_t)(jt / NBUF) & 1u;
            const uint32_t tmem_d1 = tmem_base + (uint32_t)((b * T + t) * N) + lane_off;
            mbar_wait(&d1_full[b], ph);
            tc_fence_after();
            const bool dbg_on = p.dbg && blockIdx.x == 0 && warp == 2 + T && lane == 0 && jt < 16;
            if (dbg_on) p.dbg[jt * 8 + 4] = clock64();
#pragma unroll 1
            for (int blk = blk_lo; blk < blk_hi; ++blk) {
                const int c0 = blk * 32;
                uint32_t r[32];
                tmem_ld_32x32(tmem_d1 + (uint32_t)c0, r);
                float bb[32];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float4 f = *reinterpret_cast<const float4*>(s_bias1 + c0 + 4 * e);
                    bb[4 * e] = f.x; bb[4 * e + 1] = f.y; bb[4 * e + 2] = f.z; bb[4 * e + 3] = f.w;
                }
                tmem_ld_wait();
                if (c0 + 32 > p.n_mma) {                       // columns beyond n_mma were never written by the MMAs: stale tensor memory
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c0 + i >= p.n_mma) { r[i] = 0u; bb[i] = 0.f; }
                }
                uint32_t y[16];
                if (SPLIT) {
                    // the 32 fp32 columns of the block become 16 columns of hi pairs + 16 columns of lo pairs, in place
                    uint32_t yl[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float a = fmaxf(__uint_as_float(r[2 * e]) + bb[2 * e], 0.f);
                        const float c = fmaxf(__uint_as_float(r[2 * e + 1]) + bb[2 * e + 1], 0.f);
                        y[e] = pack2<Tp>(a, c);
                        const float2 f = unpack2<Tp>(y[e]);
                        yl[e] = pack2<Tp>(a - f.x, c - f.y);
                    }
                    tmem_st_32x16(tmem_d1 + (uint32_t)c0, y);
                    tmem_st_32x16(tmem_d1 + (uint32_t)(c0 + 16), yl);
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        y[e] = pack2_relu<Tp>(__uint_as_float(r[2 * e]) + bb[2 * e], __uint_as_float(r[2 * e + 1]) + bb[2 * e + 1]);
                    tmem_st_32x16(tmem_d1 + (uint32_t)(c0 >> 1), y);
                    if (T == 2 && !SPLIT && !PAIR && p.y_early && blk == blk_lo + 1) {
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&y_half[b]);
                    }
                }
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) pair_arrive_leader(pair_leader_addr(smem_u32(&y_full[b]))); else mbar_arrive(&y_full[b]); }
            if (dbg_on) p.dbg[jt * 8 + 5] = clock64();
        };
        // the identity-shortcut operand of a tile, fetched long before it is needed (the call sites put a whole first
        // epilogue or the accumulator wait between this and its use)
        uint4 res[kResBlk][4];
        const int ew = warp - (2 + T);                         // epilogue warp 0..7
        uint32_t rph = 0;
        struct TileAt { int img, s, g; bool ok; };
        auto tile_at = [&](int jt) {                           // origin of this warp's 32 pixels (warp-uniform)
            const int tile_raw = st_of(jt) * T + t;
            const int tile = min(tile_raw, p.total_tiles - 1);
            const int tt = tile % p.tiles_per_img;
            return TileAt{tile / p.tiles_per_img, (tt / p.tiles_g) * 16 + q * 4, (tt % p.tiles_g) * 8, tile_raw < p.total_tiles};
        };
        auto prefetch_res = [&](int jt) {
            if (p.residual == nullptr) return;
            if (kOutSlots > 0 && p.tma_res) {
                // by TMA into the staging boxes the output will overwrite in place: issued as soon as the previous tile's
                // stores have read them, a whole first epilogue and accumulator wait before the use
                const TileAt ta = tile_at(jt);
                if (lane == 0) {
                    tma_store_wait_read<0>();
                    const int nb = (p.n_mma + 63) / 64;
                    mbar_expect_tx(&res_full[ew], (uint32_t)(nb * 4096));
                    for (int k = 0; k < nb; ++k)
                        tma_load_4d(out_stage + (t * kOutSlots + k) * kHaloOutBox + q * 4096, &p.tmR, &res_full[ew], k * 64, ta.g, ta.s, ta.img);
                }
                __syncwarp();
                return;
            }
            const Pix px = pixel_of(jt);
            if (!px.valid) return;
            const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const Tp*>(p.residual) + px.gpix * p.res_C);
#pragma unroll
            for (int k = 0; k < kResBlk; ++k)
                if (k >= blk_lo && k < blk_hi) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) res[k][jj] = __ldg(rp + k * 4 + jj);
                }
        };
        // epilogue 2 (or the only epilogue of a plain convolution): OUT = act(D + b (+ residual)) -> global.  The
        // residual is fetched BEFORE the wait on the accumulator so that its latency hides behind the MMAs.
        auto epi_out = [&](int jt) {
            const int b = jt % NBUF;
            const uint32_t ph = (uint32_t)(jt / NBUF) & 1u;
            const Pix px = pixel_of(jt);
            const bool has_res = p.residual != nullptr && (px.valid || (kOutSlots > 0 && p.tma_res));
            const uint32_t tmem_out = tmem_base + (uint32_t)((b * T + t) * N) + lane_off + (FUSED ? kAccCols : 0u);
            const float* sb = FUSED ? s_bias2 : s_bias1;
            // TMA-store path (fused variants): a warp stages its 32 pixel rows (128 B per row and box, 128-byte swizzle: the
            // eight lanes of a store phase hit eight different 16-byte columns) and one lane sends the 4 KB sub-box; the
            // box is clipped at the image border by the TMA unit.  Direct 16-byte stores from the accumulator layout
            // (one pixel row per lane) touch 32 lines per instruction and were measured to slow the MMAs' own
            // shared-memory reads: 18.3 k -> 12.3 k cycles per tile pair with the stores removed.
            const bool ts = kOutSlots > 0 && p.tma_store;
            const bool tr = ts && p.tma_res && p.residual != nullptr;
            int tc_img = 0, tc_s = 0, tc_g = 0;
            bool tile_ok = false;
            uint32_t stg = 0;
            if (ts) {
                const TileAt ta = tile_at(jt);
                tc_img = ta.img; tc_s = ta.s; tc_g = ta.g; tile_ok = ta.ok;
                stg = smem_u32(out_stage) + (uint32_t)(t * kOutSlots * kHaloOutBox + q * 4096);
                if (!tr) {
                    if (lane == 0) tma_store_wait_read<0>();      // the previous tile's boxes have left shared memory
                    __syncwarp();
                }
            }
            int nbox = 0;
            mbar_wait(FUSED ? &d2_full[b] : &d1_full[b], ph);
            tc_fence_after();
            if (tr) { mbar_wait(&res_full[ew], rph); rph ^= 1u; }
            const bool dbg_on = p.dbg && blockIdx.x == 0 && warp == 2 + T && lane == 0 && jt < 16;
            if (dbg_on) p.dbg[jt * 8 + 6] = clock64();
            // detector tail: exp(logit) staged per block, the sums of eight channels kept in the order of softmax_cell.cuh
            constexpr bool kHeat = N == 128 && FUSED && !WRES && !SPLIT && !PAIR;
            const bool heat = kHeat && p.heat != 0;
            float hs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dust = 0.f;
            // Staging address of half `half` (four pixels) of heat row r of this lane's cell, in the layout of the boxes that leave:
            // orient 0 (the warp's cells: 8 along x, 4 along y): two boxes {32 px, 8 rows, 4 cell rows} = the x halves;
            // orient 1 (4 along x, 8 along y): two boxes {32 px, 4 rows, 8 cell rows} = the row halves.  128-byte swizzle.
            const int cg = lane & 7, cs = lane >> 3;
            const uint32_t h_box = p.orient == 0 ? (uint32_t)(cg >> 2) * kHaloOutBox : 0u;
            const uint32_t h_row = p.orient == 0 ? (uint32_t)cs * 8u : (uint32_t)cg * 4u;
            const uint32_t h_chunk = p.orient == 0 ? 2u * (uint32_t)(cg & 3) : 2u * (uint32_t)cs;
            // lanes whose chunks would meet in a bank store the two halves of a row in the other order
            const bool h_swap = p.orient == 0 ? (cg >> 2) != 0 : ((cg >> 1) & 1) != 0;
            auto heat_addr = [&](int r, uint32_t half) {
                const uint32_t row = h_row + (uint32_t)(p.orient == 0 ? r : (r & 3));
                const uint32_t box = p.orient == 0 ? h_box : (uint32_t)(r >> 2) * kHaloOutBox;
                return stg + box + row * 128u + (((h_chunk + half) ^ (row & 7u)) << 4);
            };
#pragma unroll
            for (int blk = 0; blk < kResBlk; ++blk) {
                if (blk < blk_lo || blk >= blk_hi) continue;
                const int c0 = blk * 32;
                uint32_t r[32];
                tmem_ld_32x32(tmem_out + (uint32_t)c0, r);
                float v[32];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float4 f = *reinterpret_cast<const float4*>(sb + c0 + 4 * e);
                    v[4 * e] = f.x; v[4 * e + 1] = f.y; v[4 * e + 2] = f.z; v[4 * e + 3] = f.w;
                }
                tmem_ld_wait();
                if (c0 + 32 > p.n_mma) {                       // stale columns beyond n_mma leave as zeros (padding channels stay zero)
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c0 + i >= p.n_mma) { r[i] = 0u; v[i] = 0.f; }
                }
                if (px.valid || ts) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(r[i]);
                    if (has_res) {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const uint4 rv = tr ? ld_shared_v4(stg + (uint32_t)((blk >> 1) * kHaloOutBox) + (uint32_t)lane * 128u +
                                                               ((((uint32_t)(blk & 1) * 4u + (uint32_t)jj) ^ (uint32_t)(lane & 7)) << 4))
                                                : res[blk][jj];
                            const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = unpack2<Tp>(w[e]);
                                v[jj * 8 + e * 2] += f.x;
                                v[jj * 8 + e * 2 + 1] += f.y;
                            }
                        }
                    }
                    if (p.relu && p.dst_fp32) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    if (SPLIT && ts && p.split_out) {
                        // one 128-byte box row per block of 32 channels: 64 B of hi pairs, 64 B of lo pairs; the two
                        // slots alternate as on the fp32 path
                        const uint32_t sw = (uint32_t)(lane & 7);
                        const uint32_t box = stg + (uint32_t)((nbox & 1) * kHaloOutBox);
                        if (nbox >= 2) { if (lane == 0) tma_store_wait_read<1>(); __syncwarp(); }
                        uint32_t hw[16], lw[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const float a = p.relu ? fmaxf(v[2 * e], 0.f) : v[2 * e];
                            const float c = p.relu ? fmaxf(v[2 * e + 1], 0.f) : v[2 * e + 1];
                            hw[e] = pack2<Tp>(a, c);
                            const float2 f = unpack2<Tp>(hw[e]);
                            lw[e] = pack2<Tp>(a - f.x, c - f.y);
                        }
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            st_shared_v4(box + (uint32_t)lane * 128u + (((uint32_t)jj ^ sw) << 4), hw[jj * 4], hw[jj * 4 + 1], hw[jj * 4 + 2],
                                         hw[jj * 4 + 3]);
                            st_shared_v4(box + (uint32_t)lane * 128u + (((uint32_t)(4 + jj) ^ sw) << 4), lw[jj * 4], lw[jj * 4 + 1],
                                         lw[jj * 4 + 2], lw[jj * 4 + 3]);
                        }
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0 && tile_ok) { tma_store_4d(&p.tmD, box, blk * 64, tc_g, tc_s, tc_img); tma_store_commit(); }
                        ++nbox;
                    } else if (ts) {
                        const uint32_t sw = (uint32_t)(lane & 7);
                        if (kHeat && heat) {
                            if (blk < 2) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] = expf(v[i]);
#pragma unroll
                                for (int g = 0; g < 4; ++g) {      // eight channels = one row of the cell, summed in the order of softmax_cell.cuh
                                    float sg = v[8 * g];
#pragma unroll
                                    for (int k = 1; k < 8; ++k) sg += v[8 * g + k];
                                    hs[(blk & 1) * 4 + g] = sg;
                                    const int r = 4 * (blk & 1) + g;
                                    const float4 lo4 = make_float4(v[8 * g], v[8 * g + 1], v[8 * g + 2], v[8 * g + 3]);
                                    const float4 hi4 = make_float4(v[8 * g + 4], v[8 * g + 5], v[8 * g + 6], v[8 * g + 7]);
                                    const float4 f0 = h_swap ? hi4 : lo4, f1 = h_swap ? lo4 : hi4;
                                    st_shared_v4(heat_addr(r, h_swap ? 1u : 0u), __float_as_uint(f0.x), __float_as_uint(f0.y), __float_as_uint(f0.z), __float_as_uint(f0.w));
                                    st_shared_v4(heat_addr(r, h_swap ? 0u : 1u), __float_as_uint(f1.x), __float_as_uint(f1.y), __float_as_uint(f1.z), __float_as_uint(f1.w));
                                }
                            } else if (blk == 2) {
                                dust = expf(v[0]);                     // channel 64: the dustbin only enters the sum
                            }
                        } else if (p.dst_fp32) {
                            // one box per block of 32 fp32 channels; the two slots alternate
                            const uint32_t box = stg + (uint32_t)((nbox & 1) * kHaloOutBox);
                            if (nbox >= 2) { if (lane == 0) tma_store_wait_read<1>(); __syncwarp(); }
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj)
                                st_shared_v4(box + (uint32_t)lane * 128u + (((uint32_t)jj ^ sw) << 4), __float_as_uint(v[jj * 4]),
                                             __float_as_uint(v[jj * 4 + 1]), __float_as_uint(v[jj * 4 + 2]), __float_as_uint(v[jj * 4 + 3]));
                            fence_async_smem();
                            __syncwarp();
                            if (lane == 0 && tile_ok) { tma_store_4d(&p.tmD, box, c0, tc_g, tc_s, tc_img); tma_store_commit(); }
                            ++nbox;
                        } else {
                            // 32 16-bit channels = half a box row; the box (64 channels) leaves after its second half
                            const uint32_t box = stg + (uint32_t)((blk >> 1) * kHaloOutBox);
                            const uint32_t half = (uint32_t)(blk & 1) * 4u;
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                uint32_t w0, w1, w2, w3;
                                if (p.relu) {
                                    w0 = pack2_relu<Tp>(v[jj * 8], v[jj * 8 + 1]); w1 = pack2_relu<Tp>(v[jj * 8 + 2], v[jj * 8 + 3]);
                                    w2 = pack2_relu<Tp>(v[jj * 8 + 4], v[jj * 8 + 5]); w3 = pack2_relu<Tp>(v[jj * 8 + 6], v[jj * 8 + 7]);
                                } else {
                                    w0 = pack2<Tp>(v[jj * 8], v[jj * 8 + 1]); w1 = pack2<Tp>(v[jj * 8 + 2], v[jj * 8 + 3]);
                                    w2 = pack2<Tp>(v[jj * 8 + 4], v[jj * 8 + 5]); w3 = pack2<Tp>(v[jj * 8 + 6], v[jj * 8 + 7]);
                                }
                                st_shared_v4(box + (uint32_t)lane * 128u + (((half + (uint32_t)jj) ^ sw) << 4), w0, w1, w2, w3);
                            }
                            if ((blk & 1) || blk == blk_hi - 1) {
                                fence_async_smem();
                                __syncwarp();
                                if (lane == 0 && tile_ok) { tma_store_4d(&p.tmD, box, (blk >> 1) * 64, tc_g, tc_s, tc_img); tma_store_commit(); }
                            }
                        }
                    } else if (p.dst_fp32) {
                        float4* dp = reinterpret_cast<float4*>(static_cast<float*>(p.dst) + px.dpix * p.dst_C + c0);
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) dp[jj] = make_float4(v[jj * 4], v[jj * 4 + 1], v[jj * 4 + 2], v[jj * 4 + 3]);
                    } else {
                        uint4* dp = reinterpret_cast<uint4*>(static_cast<Tp*>(p.dst) + px.dpix * p.dst_C + c0);
#pragma unroll
                        if (p.relu) {
                            for (int jj = 0; jj < 4; ++jj)
                                dp[jj] = make_uint4(pack2_relu<Tp>(v[jj * 8], v[jj * 8 + 1]), pack2_relu<Tp>(v[jj * 8 + 2], v[jj * 8 + 3]),
                                                    pack2_relu<Tp>(v[jj * 8 + 4], v[jj * 8 + 5]), pack2_relu<Tp>(v[jj * 8 + 6], v[jj * 8 + 7]));
                        } else {
                            for (int jj = 0; jj < 4; ++jj)
                                dp[jj] = make_uint4(pack2<Tp>(v[jj * 8], v[jj * 8 + 1]), pack2<Tp>(v[jj * 8 + 2], v[jj * 8 + 3]),
                                                    pack2<Tp>(v[jj * 8 + 4], v[jj * 8 + 5]), pack2<Tp>(v[jj * 8 + 6], v[jj * 8 + 7]));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) pair_arrive_leader(pair_leader_addr(smem_u32(&d2_empty[b])));
                else mbar_arrive(FUSED ? &d2_empty[b] : &d1_empty[b]);
            }
            if (kHeat && heat) {
                // heat = exp(l) / (sum + 1e-5) (python/src/superpoint.py:111-112) = exp(l) * inv: the exponentials leave as they are
                // and the cell's inv goes to its own small map (the butterfly of softmax_cell.cuh over the eight row sums)
                hs[0] += dust;
                const float inv = 1.f / ((((hs[0] + hs[1]) + (hs[2] + hs[3])) + ((hs[4] + hs[5]) + (hs[6] + hs[7]))) + 0.00001f);
                const int oy = p.orient == 0 ? tc_s + cs : tc_g + cg;
                const int ox = p.orient == 0 ? tc_g + cg : tc_s + cs;
                if (tile_ok && oy < p.OH && ox < p.OW) p.heat_inv[((size_t)tc_img * p.OH + oy) * p.OW + ox] = inv;
                fence_async_smem();
                __syncwarp();
                if (lane == 0 && tile_ok) {
                    if (p.orient == 0) {
                        tma_store_4d(&p.tmD, stg, 8 * tc_g, 0, tc_s, tc_img);
                        tma_store_4d(&p.tmD, stg + (uint32_t)kHaloOutBox, 8 * tc_g + 32, 0, tc_s, tc_img);
                    } else {
                        tma_store_4d(&p.tmD, stg, 8 * tc_s, 0, tc_g, tc_img);
                        tma_store_4d(&p.tmD, stg + (uint32_t)kHaloOutBox, 8 * tc_s, 4, tc_g, tc_img);
                    }
                    tma_store_commit();
                }
            }
            if (dbg_on) p.dbg[jt * 8 + 7] = clock64();
        };
        if (!FUSED) {
            for (int j = 0; j < n_local; ++j) { prefetch_res(j); epi_out(j); }
        } else if (NBUF == 2) {
            // the accumulators are double buffered: the first epilogue of tile j+1 runs before the second epilogue of
            // tile j, which hides the y_full -> GEMM 2 -> d2_full round trip (and the shortcut fetch of tile j)
            for (int j = 0; j < n_local; ++j) {
                if (j > 0) prefetch_res(j - 1);
                epi1(j);
                if (j > 0) epi_out(j - 1);
            }
            if (n_local > 0) { prefetch_res(n_local - 1); epi_out(n_local - 1); }
        } else {
            for (int j = 0; j < n_local; ++j) {
                prefetch_res(j);
                epi1(j);
                epi_out(j);
            }
        }
        if (kOutSlots > 0 && lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) pair_sync();                                   // the peer has drained its accumulators; every multicast commit has landed
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) pair_tmem_dealloc(tmem_slot, kTmemCols); else tmem_dealloc(tmem_slot, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// Host
// ------------------------------------------------------------------------------------------------
static long long* g_halo_dbg = nullptr;
struct TcHaloPlan {
    HaloParams params;
    int variant;       // 0: N=64 fused, resident weights; 1: N=128 fused, T=2; 2: N=128 single convolution, T=2;
                       // 3: variant 0 on CTA pairs (tcgen05.mma.cta_group::2); 4 / 5: split precision, N = 64 / 128 fused
    int operand_type, grid;
    int real_cout;
};

template <int N, int T, int NBUF, int SA, int SW, bool FUSED, bool WRES, bool SPLIT, bool PAIR, typename Tp>
static void launch_halo_t(const TcHaloPlan* plan, cudaStream_t st, const HaloParams* override = nullptr) {
    auto kern = halo_tc_kernel<N, T, NBUF, SA, SW, FUSED, WRES, SPLIT, PAIR, Tp>;
    const size_t smem = (size_t)SA * T * kHaloBufBytes + (size_t)SW * N * 128 + (size_t)T * halo_out_slots(N, SPLIT) * kHaloOutBox + 1024;
    // function attributes are per device: set on every launch (a process may hold engines on several GPUs)
    SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const HaloParams& prm = override ? *override : plan->params;
    if (PAIR) launch_pdl_cluster(kern, dim3(plan->grid), dim3(halo_threads(T)), smem, st, 2, prm);
    else launch_pdl(kern, dim3(plan->grid), dim3(halo_threads(T)), smem, st, prm);
}

constexpr int kHaloResidentSlabs = 11;   // variant 0: every slab of a 64-channel block stays in shared memory
constexpr int kHaloRing1 = 4;            // variant 1: weight ring slots (16 KB each; 8 slots measured no faster)
constexpr int kHaloRing2 = 4;            // variant 2
constexpr bool kHaloPairDefault = false;  // variant 0 on CTA pairs unless SPB200_PAIR64 says otherwise
constexpr int kHaloRing4 = 8;            // variant 4: split precision, N = 64 (8 KB slabs)
constexpr int kHaloRing5 = 4;            // variant 5: split precision, N = 128

template <typename Tp>
static void launch_halo_v(const TcHaloPlan* plan, cudaStream_t st) {
    switch (plan->variant) {
        case 0: launch_halo_t<64, 2, 2, 2, kHaloResidentSlabs, true, true, false, false, Tp>(plan, st); break;
        case 1: launch_halo_t<128, 2, 1, 2, kHaloRing1, true, false, false, false, Tp>(plan, st); break;
        case 2: launch_halo_t<128, 2, 2, 2, kHaloRing2, false, false, false, false, Tp>(plan, st); break;
        case 3: launch_halo_t<64, 2, 2, 2, kHaloResidentSlabs, true, true, false, true, Tp>(plan, st); break;
        case 4: launch_halo_t<64, 2, 1, 2, kHaloRing4, true, false, true, false, Tp>(plan, st); break;
        case 5: launch_halo_t<128, 2, 1, 2, kHaloRing5, true, false, true, false, Tp>(plan, st); break;
        default: throw std::invalid_argument("tcgen05 halo block: bad variant");
    }
}

void launch_halo_tc(const TcHaloPlan* plan, cudaStream_t st) {
    if (!plan) throw std::runtime_error("tcgen05 halo block: no plan");
    if (plan->operand_type == PREC_FP16) launch_halo_v<__half>(plan, st);
    else launch_halo_v<__nv_bfloat16>(plan, st);
}

// The detector's last block (variant 1, 65 real channels, fp32 output through shared memory) can leave the softmax in
// depth-to-space order instead of the logits - exp(l_c) per pixel and 1 / (sum + 1e-5) per cell, heat = exp * inv -: same kernel,
// the destination map is the full-resolution map as {x, row within the cell, cell row, image}, and an epilogue warp's 32 cells
// leave as two boxes of 32 pixels x 32 rows.
bool tc_halo_heat_capable(const TcHaloPlan* plan) {
    return plan && plan->variant == 1 && plan->real_cout == 65 && plan->params.dst_fp32 && plan->params.tma_store &&
           !plan->params.tma_res && plan->params.n_mma >= 80 && plan->params.n_mma <= 96 && plan->params.relu;
}

void launch_halo_tc_heat(const TcHaloPlan* plan, float* heat_exp, float* heat_inv, int B, cudaStream_t st) {
    if (!tc_halo_heat_capable(plan)) throw std::runtime_error("tcgen05 halo block: not a detector tail");
    HaloParams p = plan->params;
    const cuuint64_t Hc = p.OH, W = 8 * (cuuint64_t)p.OW, H = 8 * Hc;
    // heatmap as {x, row within the cell, cell row, image}
    cuuint64_t dims[4] = {W, 8, Hc, (cuuint64_t)B};
    cuuint64_t str[3] = {W * 4, 8 * W * 4, H * W * 4};
    cuuint32_t box0[4] = {32, 8, 4, 1}, box1[4] = {32, 4, 8, 1};
    tc_encode_tiled(&p.tmD, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, heat_exp, dims, str, p.orient == 0 ? box0 : box1);
    p.heat = 1;
    p.heat_inv = heat_inv;
    if (plan->operand_type == PREC_FP16) launch_halo_t<128, 2, 1, 2, kHaloRing1, true, false, false, false, __half>(plan, st, &p);
    else launch_halo_t<128, 2, 1, 2, kHaloRing1, true, false, false, false, __nv_bfloat16>(plan, st, &p);
}

void tc_halo_plan_destroy(TcHaloPlan* plan) { delete plan; }

// Returns nullptr when the block does not fit this kernel (stride 2, 256 channels, ...): the caller then
// uses the per-tap kernel of block_tc.cu.  Arguments as tc_block_plan_create.
TcHaloPlan* tc_halo_plan_create(const ConvDev& c1, const ConvDev* c2, int operand_type, int real_cout, int num_sms) {
    if (operand_type != PREC_FP16 && operand_type != PREC_BF16) return nullptr;
    const int N = c1.cout_pad;
    if (N != 64 && N != 128) return nullptr;
    if (N == 64 && !c2) return nullptr;
    // split-precision block (common.cuh, SegDev): every source in the split layout, fused form only
    const bool split = c1.nseg > 0 && c1.seg[0].split != 0;
    if (split && !c2) return nullptr;
    int min_dy = 127, max_dy = -127, min_dx = 127, max_dx = -127;
    for (int s = 0; s < c1.nseg; ++s) {
        const SegDev& sg = c1.seg[s];
        if ((sg.split != 0) != split) return nullptr;
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
    plan->real_cout = real_cout;
    plan->variant = split ? (N == 64 ? 4 : 5) : (N == 64 ? 0 : (c2 ? 1 : 2));
    const int T = 2;

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

    // MMA N: the real output channels rounded up to 32, the epilogues' block width (65 -> 96 for the detector).  Rounding to 16
    // (80 columns, SPB200_N_MMA16=1; the epilogues mask the unwritten tail of the last block) was measured neutral - detector.0
    // 87.9 -> 85.8 us, detector.1 67.3 -> 69.1 us: at N <= 128 an M = 128 MMA is bound by its shared-memory operand fetch
    // ((4 KB of A + 32 N bytes of B) / 128 B per clock: 56 -> 52 clocks), not by the N / 2 clocks of the tensor pipe.
    static const bool n16 = [] { const char* e = std::getenv("SPB200_N_MMA16"); return e && e[0] == '1'; }();
    p.n_mma = std::min(N, n16 ? (real_cout + 15) / 16 * 16 : (real_cout + 31) / 32 * 32);
    auto kk_of = [](int real_c, int nchunks) {
        const int in_last = real_c - (nchunks - 1) * 64;
        return std::max(1, std::min(4, (in_last + 15) / 16));
    };
    p.has_ds = c2 && c2->nseg > 1;
    if (c2) {
        if (c2->nseg < 1 || c2->seg[0].ntaps != 1) return nullptr;
        // the second GEMM reads Y from tensor memory: N values per row, or 2 N in the split layout (the packed weights
        // may cover fewer 64-value chunks when the real output channels end earlier)
        if (!split && c2->seg[0].cin != N) return nullptr;
        if (split && (c2->seg[0].split == 0 || c2->seg[0].cin % 64 != 0 || c2->seg[0].cin < (real_cout + 31) / 32 * 64)) return nullptr;
    }
    int ds_used = 0;

    int nsteps = 0, nchunks = 0;
    const int ring = plan->variant == 0 ? kHaloResidentSlabs
                   : plan->variant == 1 ? kHaloRing1 : plan->variant == 2 ? kHaloRing2 : plan->variant == 4 ? kHaloRing4 : kHaloRing5;
    bool first_g1 = true, first_ds = true;
    auto add_step = [&](unsigned a_lo, unsigned gemm, unsigned nkk, bool first, bool last, bool acc0, int kcoord, unsigned astep = 0) {
        const unsigned slot = (unsigned)(nsteps % ring);
        p.steps[nsteps++] = halo_step(a_lo, slot * (unsigned)(N * 128 / 16), slot, gemm, nkk, first, last, acc0, (unsigned)kcoord, astep);
    };
    const bool tail_pack = [] { const char* e = std::getenv("SPB200_NO_TAIL_PACK"); return !(e && e[0] == '1'); }();   // per plan build
    // K = 16 steps of a split-layout chunk holding `real` (1..32) channels: the main slab spans the hi half and the
    // lo half (a.hi w.hi + a.lo w.hi), the lo-weight slab the hi half only (a.hi w.lo)
    auto kk_split_main = [](int real) { return 2 + (std::min(real, 32) + 15) / 16; };
    auto kk_split_lo = [](int real) { return (std::min(real, 32) + 15) / 16; };
    for (int s = 0; s < c1.nseg; ++s) {
        const SegDev& sg = c1.seg[s];
        const int nch = sg.cin / 64;
        const int kk_last = kk_of(sg.cin_real > 0 ? sg.cin_real : sg.cin, nch);
        if (sg.koff % 64) return nullptr;
        if (split && sg.koff_lo >= 0 && sg.koff_lo % 64) return nullptr;
        // the shortcut convolution of this source, if it has one (a phase-form block has it on phase (0, 0) only)
        const SegDev* ds = nullptr;
        if (p.has_ds) {
            for (int j = 1; j < c2->nseg; ++j)
                if (c2->seg[j].src == sg.src) ds = &c2->seg[j];
            if (ds) {
                if (ds->ntaps != 1 || ds->dy[0] != 0 || ds->dx[0] != 0 || ds->stride != 1 || ds->cin != sg.cin || ds->koff % 64 ||
                    ds->view != sg.view || (ds->split != 0) != split || (split && ds->koff_lo >= 0 && ds->koff_lo % 64))
                    return nullptr;
                ++ds_used;
            }
        }
        for (int c = 0; c < nch; ++c) {
            const int real_here = (sg.cin_real > 0 ? sg.cin_real : sg.cin / 2) - 32 * c;      // split layout: real channels in this chunk
            if (split && real_here <= 0) continue;                                        // a chunk of padding only
            if (nchunks >= kMaxChunks) return nullptr;
            p.chunk_seg[nchunks] = s;
            p.chunk_c0[nchunks] = c * 64;
            ++nchunks;
            // the slabs that read this chunk, in issue order: (view, gemm, K steps, K coordinate)
            struct Use { unsigned view, gemm, nkk; int kcoord; unsigned astep; };
            std::vector<Use> uses;
            const int nkk = split ? kk_split_main(real_here) : (c == nch - 1 ? kk_last : 4);
            // packed K tail: the last chunk's nine one-MMA slabs as three slabs of three MMAs, one per filter row; the K steps
            // of a slab are the row's taps dx = -1, 0, 1, i.e. A views one pixel apart along x (the group axis, or a haloed row)
            const bool packed = tail_pack && !split && c == nch - 1 && nkk == 1 && sg.koff_tail > 0 && sg.koff_tail % 64 == 0 &&
                                sg.ntaps == 9 && plan->variant != 0;
            if (packed) {
                for (int r = 0; r < 3; ++r)
                    uses.push_back({view16(r - 1, -1), 0u, 3u, sg.koff_tail / 64 + r, p.orient == 0 ? 1u : 2u});
            } else {
                for (int t = 0; t < sg.ntaps; ++t) uses.push_back({view16(sg.dy[t], sg.dx[t]), 0u, (unsigned)nkk, sg.koff / 64 + t * nch + c, 0u});
            }
            if (split && sg.koff_lo >= 0)
                for (int t = 0; t < sg.ntaps; ++t)
                    uses.push_back({view16(sg.dy[t], sg.dx[t]), 0u, (unsigned)kk_split_lo(real_here), sg.koff_lo / 64 + t * nch + c, 0u});
            if (ds) {
                uses.push_back({view16(0, 0), 1u, (unsigned)nkk, ds->koff / 64 + c, 0u});
                if (split && ds->koff_lo >= 0) uses.push_back({view16(0, 0), 1u, (unsigned)kk_split_lo(real_here), ds->koff_lo / 64 + c, 0u});
            }
            for (size_t u = 0; u < uses.size(); ++u) {
                if (nsteps >= kMaxSteps) return nullptr;
                const bool acc = uses[u].gemm == 0 ? !first_g1 : !first_ds;
                add_step(uses[u].view, uses[u].gemm, uses[u].nkk, u == 0, u + 1 == uses.size(), acc, uses[u].kcoord, uses[u].astep);
                (uses[u].gemm == 0 ? first_g1 : first_ds) = false;
            }
        }
        // activations: dims {C, group axis, slow axis, image}; a phase view steps `view` pixels of the full buffer
        const cuuint64_t C = sg.C, W = sg.W, H = sg.H, V = sg.view > 1 ? sg.view : 1;
        const cuuint64_t FW = sg.view > 1 ? sg.full_W : sg.W, FH = sg.view > 1 ? sg.full_H : sg.H;
        cuuint32_t box[4] = {64, (cuuint32_t)kHaloG, (cuuint32_t)kHaloS, 1};
        if (p.orient == 0) {
            cuuint64_t dims[4] = {C, W, H, (cuuint64_t)c1.B};
            cuuint64_t str[3] = {V * C * 2, V * FW * C * 2, FH * FW * C * 2};
            tc_encode_tiled(&p.tmA[s], dt, 4, sg.src, dims, str, box);
        } else {
            cuuint64_t dims[4] = {C, H, W, (cuuint64_t)c1.B};
            cuuint64_t str[3] = {V * FW * C * 2, V * C * 2, FH * FW * C * 2};
            tc_encode_tiled(&p.tmA[s], dt, 4, sg.src, dims, str, box);
        }
    }
    if (p.has_ds && ds_used != c2->nseg - 1) return nullptr;          // a shortcut segment without its source among the inputs
    for (int s = c1.nseg; s < kMaxSegs; ++s) p.tmA[s] = p.tmA[0];
    p.nchunks = nchunks;
    p.n1steps = nsteps;
    p.w_bytes = p.n_mma * 128;
    p.first_ds_step = nsteps;
    for (int e = nsteps - 1; e >= 0; --e)
        if (((p.steps[e] >> 36) & 3u) == 1u) p.first_ds_step = e;
    {
        cuuint64_t dims[2] = {(cuuint64_t)c1.K, (cuuint64_t)N};
        cuuint64_t str[1] = {(cuuint64_t)c1.K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)p.n_mma};        // rows beyond n_mma are never read by the MMAs
        tc_encode_tiled(&p.tmW1, dt, 2, c1.w, dims, str, box);
    }
    p.bias1 = c1.bias;
    if (c2) {
        const int ych = N / 64;
        const int y_kk_last = kk_of(real_cout, ych);
        if (c2->seg[0].koff != 0) return nullptr;
        if (split) {
            // Y sits in tensor memory as [hi 16 columns | lo 16 columns] per block of 32 channels = one 64-value K chunk
            const int ych_split = (real_cout + 31) / 32;
            if (c2->seg[0].koff_lo < 0 || c2->seg[0].koff_lo % 64) return nullptr;
            for (int c = 0; c < ych_split; ++c) {
                if (nsteps + 2 > kMaxSteps) return nullptr;
                const int real_here = real_cout - 32 * c;
                add_step((unsigned)(c * 32), 2, (unsigned)kk_split_main(real_here), 0, 0, p.has_ds || c > 0, c);
                add_step((unsigned)(c * 32), 2, (unsigned)kk_split_lo(real_here), 0, 0, true, c2->seg[0].koff_lo / 64 + c);
            }
        }
        for (int c = 0; c < ych && !split; ++c) {
            if (nsteps >= kMaxSteps) return nullptr;
            add_step((unsigned)(c * 32), 2, c == ych - 1 ? y_kk_last : 4, 0, 0, p.has_ds || c > 0, c);
        }
        cuuint64_t dims[2] = {(cuuint64_t)c2->K, (cuuint64_t)N};
        cuuint64_t str[1] = {(cuuint64_t)c2->K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)p.n_mma};        // rows beyond n_mma are never read by the MMAs
        tc_encode_tiled(&p.tmW2, dt, 2, c2->w, dims, str, box);
        p.bias2 = c2->bias;
    } else {
        p.tmW2 = p.tmW1;
    }
    p.nsteps = nsteps;
    {
        const char* e = std::getenv("SPB200_NO_ALT_ISSUE");        // per plan build
        p.alt_issue = (!(e && e[0] == '1') && plan->variant != 0 && plan->variant != 3) ? 1 : 0;
    }
    {
        // d1done (step record bit 60): the last two slabs of GEMM 1, one per issuing warp, when a shortcut slab follows them
        const char* e = std::getenv("SPB200_NO_EARLY_D1");
        int last_g1 = -1;
        for (int k = 0; k < p.n1steps; ++k)
            if (((p.steps[k] >> 36) & 3u) == 0u) last_g1 = k;
        p.early_d1 = 0;
        if (!(e && e[0] == '1') && p.alt_issue && plan->variant == 1 && c2 && last_g1 >= 1 && last_g1 < p.n1steps - 1 &&
            ((p.steps[last_g1 - 1] >> 36) & 3u) == 0u) {
            p.steps[last_g1] |= (HaloStep)1 << 60;
            p.steps[last_g1 - 1] |= (HaloStep)1 << 60;
            p.early_d1 = 1;
        }
    }
    {
        const char* e = std::getenv("SPB200_NO_Y_EARLY");
        p.y_early = (!(e && e[0] == '1') && p.alt_issue && plan->variant == 1 && c2 && p.nsteps - p.n1steps == 2 && p.n_mma > 64) ? 1 : 0;
    }
    if (plan->variant == 0 && nsteps > kHaloResidentSlabs) return nullptr;
    if (plan->variant == 0) {      // the resident-weight fast path issues four K steps per slab from one activation chunk
        if (nchunks != 1) return nullptr;
        for (int e = 0; e < p.n1steps; ++e)
            if (((p.steps[e] >> 38) & 7u) != 4u) return nullptr;
    }
    p.OH = c1.OH; p.OW = c1.OW;
    p.residual = last.residual; p.dst = last.dst;
    p.res_C = last.res_C; p.dst_H = last.dst_H; p.dst_W = last.dst_W; p.dst_C = last.dst_C;
    p.dst_stride = last.dst_stride; p.dst_off_y = last.dst_off_y; p.dst_off_x = last.dst_off_x;
    p.relu = last.relu; p.dst_fp32 = last.dst_fp32;
    p.split_out = (split && last.split_out && !last.dst_fp32) ? 1 : 0;
    if (last.split_out && !p.split_out) return nullptr;
    if (split && last.residual) return nullptr;                      // split blocks add their shortcut on the tensor core (x . I)
    // fused variants store through shared memory: destination tensor {C, group axis, slow axis, image}, one box = the 32
    // pixel rows of an epilogue warp (8 along the group axis x 4 slow rows) x 128 B of channels
    p.tma_store = 0;
    if (last.dst_stride >= 1 && (c1.OH - 1) * last.dst_stride + last.dst_off_y < last.dst_H &&
        (c1.OW - 1) * last.dst_stride + last.dst_off_x < last.dst_W &&
        last.dst_C % (last.dst_fp32 ? 32 : 64) == 0 &&
        last.dst_C >= (last.dst_fp32 ? p.n_mma : (p.split_out ? 2 * p.n_mma : (p.n_mma + 63) / 64 * 64)) &&
        (split || !std::getenv("SPB200_NO_TMA_STORE"))) {
        // a strided destination (the output phases of the transposed convolution) is the same map over every
        // dst_stride-th pixel, starting at the phase offset
        const cuuint64_t es = last.dst_fp32 ? 4 : 2;
        const cuuint64_t C = last.dst_C, W = c1.OW, H = c1.OH, S = last.dst_stride, FW = last.dst_W, FH = last.dst_H;
        const uint8_t* base = static_cast<const uint8_t*>(last.dst) + ((size_t)last.dst_off_y * FW + last.dst_off_x) * C * es;
        cuuint32_t box[4] = {(cuuint32_t)(128 / es), 8, 4, 1};
        const CUtensorMapDataType ddt = last.dst_fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : dt;
        if (p.orient == 0) {
            cuuint64_t dims[4] = {C, W, H, (cuuint64_t)c1.B};
            cuuint64_t str[3] = {S * C * es, S * FW * C * es, FH * FW * C * es};
            tc_encode_tiled(&p.tmD, ddt, 4, base, dims, str, box);
        } else {
            cuuint64_t dims[4] = {C, H, W, (cuuint64_t)c1.B};
            cuuint64_t str[3] = {S * FW * C * es, S * C * es, FH * FW * C * es};
            tc_encode_tiled(&p.tmD, ddt, 4, base, dims, str, box);
        }
        p.tma_store = 1;
        if (last.residual && last.dst_stride == 1 && last.dst_off_y == 0 && last.dst_off_x == 0 && !last.dst_fp32 && last.res_C % 64 == 0 && last.res_C >= (p.n_mma + 63) / 64 * 64 &&
            !std::getenv("SPB200_NO_TMA_RES")) {
            const cuuint64_t RC = last.res_C;
            cuuint32_t rbox[4] = {64, 8, 4, 1};
            if (p.orient == 0) {
                cuuint64_t dims[4] = {RC, W, H, (cuuint64_t)c1.B};
                cuuint64_t str[3] = {RC * 2, W * RC * 2, H * W * RC * 2};
                tc_encode_tiled(&p.tmR, dt, 4, last.residual, dims, str, rbox);
            } else {
                cuuint64_t dims[4] = {RC, H, W, (cuuint64_t)c1.B};
                cuuint64_t str[3] = {W * RC * 2, RC * 2, H * W * RC * 2};
                tc_encode_tiled(&p.tmR, dt, 4, last.residual, dims, str, rbox);
            }
            p.tma_res = 1;
        }
    }
    if (split && !p.tma_store) return nullptr;                       // the split epilogue only stores through shared memory
    // the resident-weight blocks on CTA pairs (tcgen05.mma.cta_group::2): every CTA loads its half of the rows of a weight slab
    static const bool pair_on = [] { const char* e = std::getenv("SPB200_PAIR64"); return e ? e[0] == '1' : kHaloPairDefault; }();
    if (plan->variant == 0 && pair_on && p.n_mma == 64 && num_sms >= 2) {
        plan->variant = 3;
        const cuuint32_t hbox[2] = {64, (cuuint32_t)(p.n_mma / 2)};
        {
            cuuint64_t dims[2] = {(cuuint64_t)c1.K, (cuuint64_t)N};
            cuuint64_t str[1] = {(cuuint64_t)c1.K * 2};
            tc_encode_tiled(&p.tmW1, dt, 2, c1.w, dims, str, hbox);
        }
        {
            cuuint64_t dims[2] = {(cuuint64_t)c2->K, (cuuint64_t)N};
            cuuint64_t str[1] = {(cuuint64_t)c2->K * 2};
            tc_encode_tiled(&p.tmW2, dt, 2, c2->w, dims, str, hbox);
        }
        plan->grid = 2 * std::max(1, std::min((p.n_super + 1) / 2, num_sms / 2));
    }
    {
        static int v_count[6] = {0, 0, 0, 0, 0, 0};
        const char* d = std::getenv("SPB200_HALO_DBG");         // "<variant><index>", e.g. 11 = second plan of variant 1
        if (d && plan->variant < 6 && d[0] - '0' == plan->variant && v_count[plan->variant]++ == atoi(d + 1)) {
            cudaMalloc(&g_halo_dbg, 16 * 8 * sizeof(long long));
            cudaMemset(g_halo_dbg, 0, 16 * 8 * sizeof(long long));
            p.dbg = g_halo_dbg;
            fprintf(stderr, "halo dbg on plan OH=%d OW=%d n_mma=%d nsteps=%d n1=%d has_ds=%d fp32=%d\n", p.OH, p.OW, p.n_mma, p.nsteps, p.n1steps, p.has_ds, p.dst_fp32);
        }
    }
    return plan.release();
}

}  // namespace spb200
extern "C" __attribute__((visibility("default"))) int spb200_debug_halo(long long* host) {
    if (!spb200::g_halo_dbg) return 1;
    cudaDeviceSynchronize();
    cudaMemcpy(host, spb200::g_halo_dbg, 16 * 8 * sizeof(long long), cudaMemcpyDeviceToHost);
    return 0;
}
