// Descriptor matching on the tensor cores (reference python/src/inference.py:88-96, see match.cu for the semantics).
//
// The distance matrix needs fp32-grade dot products - nearest neighbours are decided by differences far below fp16
// resolution - so every descriptor is split into two fp16 parts x = hi + lo (|lo| <= 2^-11 |x|) and the dot product is
// the three-term sum hi.hi' + hi.lo' + lo.hi' (the dropped lo.lo' is below 2^-22 relative): one K = 3 D GEMM on
// operands [hi | hi | lo] and [hi | lo | hi], fp32 accumulation in TMEM.
//
//   match_split_kernel   fp32 descriptors -> both 16-bit operand layouts + squared norms (rows past the count: zeros)
//   match_tc_kernel      a CTA keeps a 128-descriptor query tile resident in shared memory (six 64-wide K chunks,
//                        TMA, 128B swizzle) and streams train tiles of 128 through a four-stage ring; 24 tcgen05.mma
//                        (M = N = 128, K = 16) per tile into one of two TMEM accumulators; four epilogue warps turn the
//                        accumulator into squared distances and keep, per query row (= TMEM lane = thread), the
//                        minimum (distance, index) key in a register - no cross-thread reduction.  One atomicMin
//                        per row at the end (the train range may be split over several CTAs).
// The launch runs twice, a against b and b against a: the column minima of one are the row minima of the other.
#include <cuda.h>

#include <cstring>

#include "kernels.h"
#include "tc_common.cuh"

namespace spb200 {

constexpr int kMcD = 128;                       // descriptor width of this path
constexpr int kMcK = 3 * kMcD;                  // GEMM K
constexpr int kMcChunks = kMcK / 64;            // 64-element (128-byte) K chunks
constexpr int kMcTile = 128;
constexpr int kMcStages = 4;
constexpr int kMcChunkBytes = kMcTile * 128;    // 16 KB
constexpr int kMcThreads = 192;
constexpr int kMcSmem = (kMcChunks + kMcStages) * kMcChunkBytes + 1024;

// one warp per descriptor row; role 0 = [hi | hi | lo], role 1 = [hi | lo | hi]
__global__ void __launch_bounds__(256) match_split_kernel(const float* __restrict__ desc, const int* __restrict__ count, int cap,
                                                          __half* __restrict__ op0, __half* __restrict__ op1,
                                                          float* __restrict__ norms) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), b = blockIdx.y;
    const int n = min(__ldg(count + b), cap);
    // rows past the count are zero up to the end of the last 128-row tile the GEMM reads (what lies beyond only meets
    // masked rows / columns: an accumulator element depends on its own row and column vectors alone)
    if (row >= min(cap, (n + kMcTile - 1) / kMcTile * kMcTile)) return;
    const size_t r = (size_t)b * cap + row;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < n) v = __ldg(reinterpret_cast<const float4*>(desc + r * kMcD) + lane);
    const float x[4] = {v.x, v.y, v.z, v.w};
    __half hi[4], lo[4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        hi[i] = __float2half_rn(x[i]);
        lo[i] = __float2half_rn(x[i] - __half2float(hi[i]));
        s = fmaf(x[i], x[i], s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) norms[r] = s;
    const uint2 ph = make_uint2((uint32_t)__half_as_ushort(hi[0]) | ((uint32_t)__half_as_ushort(hi[1]) << 16),
                                (uint32_t)__half_as_ushort(hi[2]) | ((uint32_t)__half_as_ushort(hi[3]) << 16));
    const uint2 pl = make_uint2((uint32_t)__half_as_ushort(lo[0]) | ((uint32_t)__half_as_ushort(lo[1]) << 16),
                                (uint32_t)__half_as_ushort(lo[2]) | ((uint32_t)__half_as_ushort(lo[3]) << 16));
    uint2* o0 = reinterpret_cast<uint2*>(op0 + r * kMcK) + lane;
    uint2* o1 = reinterpret_cast<uint2*>(op1 + r * kMcK) + lane;
    o0[0] = ph; o0[32] = ph; o0[64] = pl;
    o1[0] = ph; o1[32] = pl; o1[64] = ph;
}

struct MatchTcParams {
    CUtensorMap tmQ;              // query operand (role 0) [B * cap][384] fp16, box 64 x 128
    CUtensorMap tmT;              // train operand (role 1)
    const float* nq;              // squared norms [B][cap]
    const float* nt;
    const int* cq;                // counts [B]
    const int* ct;
    unsigned long long* best;     // [B][cap] (distance bits << 32 | train index), pre-set to all ones
    int cap;
};

__global__ void __launch_bounds__(kMcThreads, 1) match_tc_kernel(const __grid_constant__ MatchTcParams p) {
    constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // fp16 x fp16 -> fp32
    extern __shared__ uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t q_full, full[kMcStages], empty[kMcStages], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_nt[2][kMcTile];
    uint8_t* s_q = dyn_smem + ((1024u - (smem_u32(dyn_smem) & 1023u)) & 1023u);     // [6 chunks][128 rows][128 B]
    uint8_t* s_t = s_q + kMcChunks * kMcChunkBytes;                                  // [4 stages][128 rows][128 B]
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x / 32), 0), lane = threadIdx.x % 32;
    const int b = blockIdx.z;
    const int nq = min(__ldg(p.cq + b), p.cap), nt = min(__ldg(p.ct + b), p.cap);
    const int q0 = blockIdx.x * kMcTile;
    if (q0 >= nq || nt <= 0) return;
    const int ntiles = (nt + kMcTile - 1) / kMcTile;
    const int my_tiles = (ntiles - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;     // tiles blockIdx.y, + gridDim.y, ...
    if (my_tiles <= 0) return;

    if (threadIdx.x == 0) {
        mbar_init(&q_full, 1);
        for (int s = 0; s < kMcStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        prefetch_tmap(&p.tmQ);
        prefetch_tmap(&p.tmT);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (elect_one()) {
            mbar_expect_tx(&q_full, kMcChunks * kMcChunkBytes);
            for (int c = 0; c < kMcChunks; ++c) tma_load_2d(s_q + c * kMcChunkBytes, &p.tmQ, &q_full, c * 64, b * p.cap + q0);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int t0 = ((int)blockIdx.y + i * (int)gridDim.y) * kMcTile;
                for (int c = 0; c < kMcChunks; ++c) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    mbar_expect_tx(&full[stage], kMcChunkBytes);
                    tma_load_2d(s_t + stage * kMcChunkBytes, &p.tmT, &full[stage], c * 64, b * p.cap + t0);
                    if (++stage == kMcStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (elect_one()) {
            constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | (2u << 29);
            const uint32_t q_lo = umma_desc_lo(smem_u32(s_q)), t_lo = umma_desc_lo(smem_u32(s_t));
            mbar_wait(&q_full, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int a = i & 1;
                if (i >= 2) mbar_wait(&acc_empty[a], ((i >> 1) - 1) & 1);
                tc_fence_after();
                for (int c = 0; c < kMcChunks; ++c) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_f16_w(tmem + a * 128, q_lo + (uint32_t)((c * kMcChunkBytes + kk * 32) >> 4), kHi,
                                   t_lo + (uint32_t)((stage * kMcChunkBytes + kk * 32) >> 4), kHi, kIdesc, (c > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(&empty[stage]);
                    if (++stage == kMcStages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(&acc_full[a]);
            }
        }
        __syncwarp();
    } else {
        // ---------------- epilogue: thread = query row ----------------
        const int q = warp & 3;                                   // TMEM lane quarter of this warp
        const int row = q * 32 + lane;
        const int et = (int)threadIdx.x - 64;                     // 0 .. 127 among the epilogue threads
        const float nrm = __ldg(p.nq + (size_t)b * p.cap + min(q0 + row, p.cap - 1));
        unsigned long long best = ~0ull;
        for (int i = 0; i < my_tiles; ++i) {
            const int a = i & 1;
            const int t0 = ((int)blockIdx.y + i * (int)gridDim.y) * kMcTile;
            // norms of the tile's train descriptors (buffer a was last read two tiles ago; the named barrier below
            // orders those reads before these writes)
            s_nt[a][et] = t0 + et < nt ? __ldg(p.nt + (size_t)b * p.cap + t0 + et) : 0.f;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(&acc_full[a], (i >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int blk = 0; blk < 4; ++blk) {
                uint32_t r[32];
                tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * 128 + blk * 32), r);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int col = blk * 32 + e;
                    const float d2 = fmaxf(nrm + s_nt[a][col] - 2.f * __uint_as_float(r[e]), 0.f);
                    const unsigned long long key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)(t0 + col);
                    if (t0 + col < nt) best = min(best, key);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
        }
        if (q0 + row < nq && best != ~0ull) atomicMin(p.best + (size_t)b * p.cap + q0 + row, best);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

static void encode_operand(CUtensorMap* map, const __half* base, long rows) {
    cuuint64_t dims[2] = {(cuuint64_t)kMcK, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)kMcK * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)kMcTile};
    tc_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, str, box);
}

size_t match_tc_workspace_bytes(int B, int cap) {
    // four operand arrays (two roles x two sets) + two norm arrays
    return (size_t)B * cap * ((size_t)4 * kMcK * sizeof(__half) + 2 * sizeof(float));
}

void match_init_launch(unsigned long long* best_a, unsigned long long* best_b, long n, cudaStream_t st);   // match.cu
void match_finish_launch(const int* count_a, int B, int cap, const unsigned long long* best_a, const unsigned long long* best_b,
                         float max_dist, int* match, float* dist, cudaStream_t st);                        // match.cu

void launch_match_tc(const float* desc_a, const int* count_a, const float* desc_b, const int* count_b, int B, int cap,
                     float max_dist, void* workspace, unsigned long long* best_a, unsigned long long* best_b, int* match,
                     float* dist, int num_sms, cudaStream_t st) {
    const size_t rows = (size_t)B * cap;
    __half* a0 = static_cast<__half*>(workspace);
    __half* a1 = a0 + rows * kMcK;
    __half* b0 = a1 + rows * kMcK;
    __half* b1 = b0 + rows * kMcK;
    float* na = reinterpret_cast<float*>(b1 + rows * kMcK);
    float* nb = na + rows;
    dim3 sg((cap + 7) / 8, B);
    match_split_kernel<<<sg, 256, 0, st>>>(desc_a, count_a, cap, a0, a1, na);
    SPB_CHECK_LAUNCH();
    match_split_kernel<<<sg, 256, 0, st>>>(desc_b, count_b, cap, b0, b1, nb);
    SPB_CHECK_LAUNCH();
    match_init_launch(best_a, best_b, (long)rows, st);
    SPB_CUDA(cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMcSmem));
    const int strips = (cap + kMcTile - 1) / kMcTile;
    const int splits = std::max(1, std::min(8, (2 * num_sms + strips * B - 1) / (strips * B)));
    MatchTcParams p;
    std::memset(&p, 0, sizeof(p));
    p.cap = cap;
    // a against b: queries = a (role 0), train = b (role 1)
    encode_operand(&p.tmQ, a0, (long)rows);
    encode_operand(&p.tmT, b1, (long)rows);
    p.nq = na; p.nt = nb; p.cq = count_a; p.ct = count_b; p.best = best_a;
    match_tc_kernel<<<dim3(strips, splits, B), kMcThreads, kMcSmem, st>>>(p);
    SPB_CHECK_LAUNCH();
    // b against a
    encode_operand(&p.tmQ, b0, (long)rows);
    encode_operand(&p.tmT, a1, (long)rows);
    p.nq = nb; p.nt = na; p.cq = count_b; p.ct = count_a; p.best = best_b;
    match_tc_kernel<<<dim3(strips, splits, B), kMcThreads, kMcSmem, st>>>(p);
    SPB_CHECK_LAUNCH();
    match_finish_launch(count_a, B, cap, best_a, best_b, max_dist, match, dist, st);
}

}  // namespace spb200
