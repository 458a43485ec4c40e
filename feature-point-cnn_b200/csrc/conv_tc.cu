// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a.
//
// One CTA computes a 128-pixel x BLOCK_N-channel output tile of one image:
//   D[128, N] (fp32, TMEM) = sum over K steps  A_step[128, 64] (smem, TMA)  *  W_step[N, 64]^T (smem, TMA)
// A K step is 64 input channels of one tap of one K segment (a segment = one source tensor of a
// channel concatenation, or the 1x1 shortcut convolution accumulated into the same tile).  The
// activations are NHWC 16-bit, so the 128 pixels x 64 channels of a tap are ONE 5-d TMA box
// {64 ch, tw, 1, th, 1} at the tap's shifted coordinates; out-of-image pixels are zero-filled by TMA,
// which is the convolution's padding.  Stride-2 convolutions read the same tensor through a parity
// view (2C, W/2, 2, H/2, N) so that a tap is again a dense box.  Weights are [cout][K] K-major, one
// 2-d TMA box {64, N} per step.  Both operands land in 128B-swizzled K-major tiles that tcgen05.mma
// (kind::f16, M=128, N=BLOCK_N, K=16) consumes through shared-memory descriptors.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> +bias (+residual) -> ReLU -> 16-bit or fp32 NHWC store).
// Reference semantics: python/src/resnet_blocks.py:14-27 (conv+BN+ReLU, shortcut, add, ReLU) with
// BatchNorm folded into weights/bias by the engine.
#include <cuda.h>

#include <cstring>
#include <memory>
#include <mutex>

#include "kernels.h"
#include "tc_common.cuh"

namespace spb200 {

// ------------------------------------------------------------------------------------------------
// Kernel
// ------------------------------------------------------------------------------------------------
struct TcSeg {
    int ntaps, nchunks, stride, C;
    int8_t dy[kMaxTaps], dx[kMaxTaps];
};

struct TcParams {
    CUtensorMap tmA[kMaxSegs];
    CUtensorMap tmW;
    TcSeg seg[kMaxSegs];
    int nseg, ksteps;
    int th, tw, tiles_x;
    int OH, OW;
    const float* bias;
    const void* residual;
    void* dst;
    int res_C, dst_H, dst_W, dst_C, dst_stride, dst_off_y, dst_off_x;
    int relu, dst_fp32, cout_pad;
};

constexpr int kTcThreads = 192;
constexpr int kTileM = 128;
constexpr int kABytes = kTileM * 128;     // 128 pixels x 64 channels x 2 B

template <int BLOCK_N, int STAGES, typename T>
__global__ void __launch_bounds__(kTcThreads) conv_tc_kernel(const __grid_constant__ TcParams p) {
    constexpr int kBBytes = BLOCK_N * 128;
    constexpr int kStageBytes = kABytes + kBBytes;
    constexpr uint32_t kTmemCols = BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;
    // instruction descriptor: fp32 accumulate, A/B format, K-major both, N>>3, M>>4
    constexpr uint32_t kIdesc = (1u << 4) | (OperandFmt<T>::value << 7) | (OperandFmt<T>::value << 10) |
                                ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

    extern __shared__ uint8_t dyn_smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t acc_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_bias[BLOCK_N];

    // aligned up to 1024 B by pointer ARITHMETIC on dyn_smem: the compiler keeps the shared address space (LDS/STS,
    // 32-bit addresses) instead of falling back to generic loads
    uint8_t* tiles = dyn_smem + ((1024u - (smem_u32(dyn_smem) & 1023u)) & 1023u);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int img = blockIdx.y;
    const int y0 = (blockIdx.x / p.tiles_x) * p.th, x0 = (blockIdx.x % p.tiles_x) * p.tw;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < p.nseg; ++s) prefetch_tmap(&p.tmA[s]);
        prefetch_tmap(&p.tmW);
    }
    if (warp == 1) tmem_alloc(&tmem_slot, kTmemCols);
    for (int i = threadIdx.x; i < BLOCK_N; i += kTcThreads) s_bias[i] = p.bias[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int kstep = 0;
            for (int s = 0; s < p.nseg; ++s) {
                const TcSeg& sg = p.seg[s];
                for (int t = 0; t < sg.ntaps; ++t) {
                    int cx, cy, cp, cc;
                    if (sg.stride == 1) {
                        cx = x0 + sg.dx[t]; cy = y0 + sg.dy[t]; cp = 0; cc = 0;
                    } else {      // parity view: pixel (2*(o+q)+par)
                        const int px = sg.dx[t] & 1, py = sg.dy[t] & 1;
                        cx = x0 + (sg.dx[t] - px) / 2; cy = y0 + (sg.dy[t] - py) / 2; cp = py; cc = px * sg.C;
                    }
                    for (int c = 0; c < sg.nchunks; ++c, ++kstep) {
                        mbar_wait(&empty_bar[stage], phase ^ 1u);
                        uint8_t* a_dst = tiles + stage * kStageBytes;
                        mbar_expect_tx(&full_bar[stage], (uint32_t)kStageBytes);
                        tma_load_5d(a_dst, &p.tmA[s], &full_bar[stage], cc + c * 64, cx, cp, cy, img);
                        tma_load_2d(a_dst + kABytes, &p.tmW, &full_bar[stage], kstep * 64, 0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int k = 0; k < p.ksteps; ++k) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(tiles + stage * kStageBytes);
                const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {       // 4 x (K = 16) per 64-channel step; +32 B inside the swizzle row
                    umma_f16(tmem_acc, umma_smem_desc(a_addr + kk * 32), umma_smem_desc(b_addr + kk * 32), kIdesc,
                             (k > 0 || kk > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);        // frees the smem stage when these MMAs retire
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
            umma_commit(&acc_bar);                     // accumulator complete
        }
        __syncwarp();
    } else {
        // ===================== epilogue =====================
        const int q = warp % 4;                        // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int oy = y0 + row / p.tw, ox = x0 + row % p.tw;
        const bool valid = oy < p.OH && ox < p.OW;
        const size_t gpix = ((size_t)img * p.OH + oy) * p.OW + ox;             // GEMM-row pixel (residual)
        const size_t dpix = ((size_t)img * p.dst_H + (oy * p.dst_stride + p.dst_off_y)) * p.dst_W +
                            (ox * p.dst_stride + p.dst_off_x);
        mbar_wait(&acc_bar, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            if (valid) {
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]) + s_bias[c0 + i];
                if (p.residual) {
                    const uint4* rp = reinterpret_cast<const uint4*>(static_cast<const T*>(p.residual) + gpix * p.res_C + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 u = __ldg(rp + j);
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
                if (p.dst_fp32) {
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
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_acc, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------
// Host: tensor maps and launch plans
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    if (!fn) throw std::runtime_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return fn;
}

void tc_encode_tiled_ex(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims,
                      const cuuint64_t* strides_bytes, const cuuint32_t* box, bool swizzle128) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode_tiled_fn()(map, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
}

void tc_encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims,
                     const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    tc_encode_tiled_ex(map, dt, rank, base, dims, strides_bytes, box, true);
}

struct TcConvPlan {
    TcParams params;
    dim3 grid;
    int block_n;
    int operand_type;
    size_t smem;
};

template <int BLOCK_N, int STAGES, typename T>
static void launch_tc_t(const TcConvPlan* plan, cudaStream_t st) {
    auto kern = conv_tc_kernel<BLOCK_N, STAGES, T>;
    const size_t smem = (size_t)STAGES * (kABytes + BLOCK_N * 128) + 1024;
    SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<plan->grid, kTcThreads, smem, st>>>(plan->params);
    SPB_CHECK_LAUNCH();
}

template <typename T>
static void launch_tc_n(const TcConvPlan* plan, cudaStream_t st) {
    switch (plan->block_n) {
        case 64: launch_tc_t<64, 4, T>(plan, st); break;
        case 128: launch_tc_t<128, 3, T>(plan, st); break;
        case 256: launch_tc_t<256, 4, T>(plan, st); break;
        default: throw std::invalid_argument("tcgen05 conv: unsupported channel count " + std::to_string(plan->block_n));
    }
}

void launch_conv_tc(const TcConvPlan* plan, cudaStream_t st) {
    if (!plan) throw std::runtime_error("tcgen05 conv: no plan");
    if (plan->operand_type == PREC_FP16) launch_tc_n<__half>(plan, st);
    else launch_tc_n<__nv_bfloat16>(plan, st);
}

void tc_plan_destroy(TcConvPlan* plan) { delete plan; }

TcConvPlan* tc_plan_create(const ConvDev& c, int operand_type) {
    if (operand_type != PREC_FP16 && operand_type != PREC_BF16) throw std::invalid_argument("tcgen05 conv: bad operand type");
    auto plan = std::make_unique<TcConvPlan>();
    TcParams& p = plan->params;
    std::memset(&p, 0, sizeof(p));
    const CUtensorMapDataType dt = operand_type == PREC_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if (c.cout_pad != 64 && c.cout_pad != 128 && c.cout_pad != 256)
        throw std::invalid_argument("tcgen05 conv: cout_pad must be 64, 128 or 256");
    if (c.dst_C % 8 != 0 || (c.residual && c.res_C % 8 != 0)) throw std::invalid_argument("tcgen05 conv: channel strides must be multiples of 8");

    // tile shape: th*tw = 128, fewest tiles; ties prefer the widest rows (better store coalescing)
    int best_th = 8, best_tw = 16;
    long best = -1;
    for (int tw = 128; tw >= 1; tw >>= 1) {
        const int th = 128 / tw;
        const long n = (long)((c.OH + th - 1) / th) * ((c.OW + tw - 1) / tw);
        if (best < 0 || n < best) { best = n; best_th = th; best_tw = tw; }
    }
    p.th = best_th; p.tw = best_tw;
    p.tiles_x = (c.OW + p.tw - 1) / p.tw;
    const int tiles_y = (c.OH + p.th - 1) / p.th;
    plan->grid = dim3(p.tiles_x * tiles_y, c.B);
    plan->block_n = c.cout_pad;
    plan->operand_type = operand_type;

    p.nseg = c.nseg;
    int ksteps = 0;
    for (int s = 0; s < c.nseg; ++s) {
        const SegDev& sg = c.seg[s];
        if (sg.cin % 64 != 0 || sg.C % 64 != 0) throw std::invalid_argument("tcgen05 conv: channels must be multiples of 64");
        if (sg.koff != ksteps * 64) throw std::invalid_argument("tcgen05 conv: unexpected K offset");
        TcSeg& ts = p.seg[s];
        ts.ntaps = sg.ntaps; ts.nchunks = sg.cin / 64; ts.stride = sg.stride; ts.C = sg.C;
        for (int t = 0; t < sg.ntaps; ++t) { ts.dy[t] = sg.dy[t]; ts.dx[t] = sg.dx[t]; }
        ksteps += sg.ntaps * ts.nchunks;
        const cuuint64_t C = sg.C, W = sg.W, H = sg.H;
        cuuint32_t box[5] = {64, (cuuint32_t)p.tw, 1, (cuuint32_t)p.th, 1};
        if (sg.stride == 1) {
            cuuint64_t dims[5] = {C, W, 1, H, (cuuint64_t)c.B};
            cuuint64_t str[4] = {C * 2, W * C * 2, W * C * 2, H * W * C * 2};
            tc_encode_tiled(&p.tmA[s], dt, 5, sg.src, dims, str, box);
        } else if (sg.stride == 2) {
            if (W % 2 || H % 2) throw std::invalid_argument("tcgen05 conv: stride-2 source must have even height and width");
            cuuint64_t dims[5] = {2 * C, W / 2, 2, H / 2, (cuuint64_t)c.B};
            cuuint64_t str[4] = {2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
            tc_encode_tiled(&p.tmA[s], dt, 5, sg.src, dims, str, box);
        } else {
            throw std::invalid_argument("tcgen05 conv: stride must be 1 or 2");
        }
    }
    int ktail = 0;                                            // packed K tails follow the segments (common.cuh, SegDev): unused here
    for (int s = 0; s < c.nseg; ++s) ktail += c.seg[s].koff_tail > 0 ? 3 : 0;
    if ((ksteps + ktail) * 64 != c.K) throw std::invalid_argument("tcgen05 conv: K does not match the segments");
    p.ksteps = ksteps;
    {
        cuuint64_t dims[2] = {(cuuint64_t)c.K, (cuuint64_t)c.cout_pad};
        cuuint64_t str[1] = {(cuuint64_t)c.K * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)c.cout_pad};
        tc_encode_tiled(&p.tmW, dt, 2, c.w, dims, str, box);
    }
    p.OH = c.OH; p.OW = c.OW;
    p.bias = c.bias; p.residual = c.residual; p.dst = c.dst;
    p.res_C = c.res_C; p.dst_H = c.dst_H; p.dst_W = c.dst_W; p.dst_C = c.dst_C;
    p.dst_stride = c.dst_stride; p.dst_off_y = c.dst_off_y; p.dst_off_x = c.dst_off_x;
    p.relu = c.relu; p.dst_fp32 = c.dst_fp32; p.cout_pad = c.cout_pad;
    return plan.release();
}

}  // namespace spb200
