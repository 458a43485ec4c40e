// CUDA-core (fp32) kernels of the spb200 engine.
//
//  * stem_pool_kernel    - 7x7 stride-2 convolution + folded BatchNorm + ReLU + 3x3 stride-2
//                          max-pool in one pass (reference python/src/superpoint.py:12-15,20-23).
//  * conv_simt_kernel    - generic fp32 implicit-GEMM convolution over up to three K segments with
//                          bias / residual / ReLU epilogue: the PREC_FP32 datapath of every residual
//                          block (python/src/resnet_blocks.py:14-27) and of the transposed conv
//                          (python/src/superpoint.py:45,55), and the on-device check for the tcgen05 path.
//  * nhwc_to_nchw_kernel - layout export for the drop-in SuperPoint.forward triple
//                          (python/src/superpoint.py:115 returns NCHW fp32 tensors).
#include "kernels.h"

namespace spb200 {

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }

// ------------------------------------------------------------------------------------------------
// Stem: conv7x7/s2/p3 + BN + ReLU + maxpool3x3/s2/p1, 8x8 pooled pixels x 64 channels per block.
// ------------------------------------------------------------------------------------------------
constexpr int kStemPT = 8;                    // pooled tile edge
constexpr int kStemCT = 2 * kStemPT + 1;      // conv tile edge (17)
constexpr int kStemIT = 2 * kStemCT + 5;      // input patch edge (39)

template <typename TOut, int CIN>
__global__ void __launch_bounds__(256)
stem_pool_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                 TOut* __restrict__ out, int H, int W) {
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;                                   // [CIN*49][64]
    float* s_in = s_w + CIN * 49 * 64;                   // [CIN][39][39]
    float* s_conv = s_in + CIN * kStemIT * kStemIT;      // [289][64] (padded start to 16 B below)
    s_conv = (float*)(((uintptr_t)s_conv + 15) & ~(uintptr_t)15);

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int CH = H / 2, CW = W / 2, PH = H / 4, PW = W / 4;
    const int py0 = blockIdx.y * kStemPT, px0 = blockIdx.x * kStemPT;
    const int iy0 = 4 * py0 - 5, ix0 = 4 * px0 - 5;

    for (int i = tid; i < CIN * 49 * 64; i += 256) s_w[i] = w[i];
    for (int i = tid; i < CIN * kStemIT * kStemIT; i += 256) {
        int c = i / (kStemIT * kStemIT), r = i % (kStemIT * kStemIT);
        int y = iy0 + r / kStemIT, x = ix0 + r % kStemIT;
        float v = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) v = img[((size_t)(b * CIN + c) * H + y) * W + x];
        s_in[i] = v;
    }
    __syncthreads();

    constexpr int NPIX = kStemCT * kStemCT;
    for (int id = tid; id < NPIX * 4; id += 256) {
        const int g = id / NPIX, cp = id % NPIX;
        const int cyl = cp / kStemCT, cxl = cp % kStemCT;
        const int cy = 2 * py0 - 1 + cyl, cx = 2 * px0 - 1 + cxl;
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = bias[g * 16 + i];
        if (cy >= 0 && cy < CH && cx >= 0 && cx < CW) {
            for (int c = 0; c < CIN; ++c) {
                const float* pin = s_in + c * kStemIT * kStemIT + (2 * cyl) * kStemIT + 2 * cxl;
                const float* pw = s_w + c * 49 * 64 + g * 16;
#pragma unroll 1
                for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
                    for (int kx = 0; kx < 7; ++kx) {
                        const float v = pin[ky * kStemIT + kx];
                        const float4* w4 = reinterpret_cast<const float4*>(pw + (ky * 7 + kx) * 64);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 ww = w4[q];
                            acc[q * 4 + 0] = fmaf(v, ww.x, acc[q * 4 + 0]);
                            acc[q * 4 + 1] = fmaf(v, ww.y, acc[q * 4 + 1]);
                            acc[q * 4 + 2] = fmaf(v, ww.z, acc[q * 4 + 2]);
                            acc[q * 4 + 3] = fmaf(v, ww.w, acc[q * 4 + 3]);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaxf(acc[i], 0.f);
        } else {
            // outside the conv output: neutral for the max because every valid value is >= 0
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = 0.f;
        }
        float4* dst = reinterpret_cast<float4*>(s_conv + cp * 64 + g * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_float4(acc[q * 4], acc[q * 4 + 1], acc[q * 4 + 2], acc[q * 4 + 3]);
    }
    __syncthreads();

    for (int o = tid; o < kStemPT * kStemPT * 64; o += 256) {
        const int ch = o % 64, pp = o / 64;
        const int ppy = pp / kStemPT, ppx = pp % kStemPT;
        const int py = py0 + ppy, px = px0 + ppx;
        if (py >= PH || px >= PW) continue;
        float m = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
                m = fmaxf(m, s_conv[((2 * ppy + dy) * kStemCT + 2 * ppx + dx) * 64 + ch]);
        out[((size_t)(b * PH + py) * PW + px) * 64 + ch] = from_float<TOut>(m);
    }
}

static size_t stem_smem_bytes(int cin) {
    return (size_t)(cin * 49 * 64 + cin * kStemIT * kStemIT + kStemCT * kStemCT * 64) * sizeof(float) + 16;
}

template <typename TOut, int CIN>
static void launch_stem_t(const float* img, int B, int H, int W, const float* w, const float* bias, void* dst,
                          cudaStream_t st) {
    auto kern = stem_pool_kernel<TOut, CIN>;
    const size_t smem = stem_smem_bytes(CIN);
    // function attributes are per device: set on every launch (a process may hold engines on several GPUs)
    SPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((W / 4 + kStemPT - 1) / kStemPT, (H / 4 + kStemPT - 1) / kStemPT, B);
    kern<<<grid, 256, smem, st>>>(img, w, bias, (TOut*)dst, H, W);
    SPB_CHECK_LAUNCH();
}

void launch_stem_pool(const float* img, int B, int C, int H, int W, const float* w, const float* bias, void* dst,
                      int dst_type, cudaStream_t st) {
    if (C != 1 && C != 3) throw std::invalid_argument("stem: input must have 1 or 3 channels");
    if (dst_type == PREC_FP32) {
        if (C == 1) launch_stem_t<float, 1>(img, B, H, W, w, bias, dst, st);
        else launch_stem_t<float, 3>(img, B, H, W, w, bias, dst, st);
    } else if (dst_type == PREC_FP16) {
        if (C == 1) launch_stem_t<__half, 1>(img, B, H, W, w, bias, dst, st);
        else launch_stem_t<__half, 3>(img, B, H, W, w, bias, dst, st);
    } else {
        if (C == 1) launch_stem_t<__nv_bfloat16, 1>(img, B, H, W, w, bias, dst, st);
        else launch_stem_t<__nv_bfloat16, 3>(img, B, H, W, w, bias, dst, st);
    }
}

// ------------------------------------------------------------------------------------------------
// Generic fp32 implicit GEMM: 64 output pixels x 64 output channels per block, K in steps of 16.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvDev p) {
    __shared__ __align__(16) float As[16][68];
    __shared__ __align__(16) float Bs[16][64];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int M = p.B * p.OH * p.OW;
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;

    // A-load role: one pixel, four consecutive channels of the 16-channel chunk
    const int a_pix = tid / 4, a_kq = (tid % 4) * 4;
    const int a_m = m0 + a_pix;
    const bool a_ok = a_m < M;
    int a_n = 0, a_oy = 0, a_ox = 0;
    if (a_ok) {
        a_n = a_m / (p.OH * p.OW);
        int r = a_m % (p.OH * p.OW);
        a_oy = r / p.OW;
        a_ox = r % p.OW;
    }
    // B-load role: one k row, four consecutive output channels
    const int b_k = tid / 16, b_n = n0 + (tid % 16) * 4;
    const float* wp = static_cast<const float*>(p.w);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int s = 0; s < p.nseg; ++s) {
        const SegDev& sg = p.seg[s];
        const float* src = static_cast<const float*>(sg.src);
        for (int t = 0; t < sg.ntaps; ++t) {
            const int iy = a_oy * sg.stride + sg.dy[t], ix = a_ox * sg.stride + sg.dx[t];
            const bool ok = a_ok && iy >= 0 && iy < sg.H && ix >= 0 && ix < sg.W;
            const float* arow = src + ((size_t)(a_n * sg.H + iy) * sg.W + ix) * sg.C + a_kq;
            const int kbase = sg.koff + t * sg.cin;
            for (int c0 = 0; c0 < sg.cin; c0 += 16) {
                float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) av = *reinterpret_cast<const float4*>(arow + c0);
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (b_n < p.cout_pad) bv = *reinterpret_cast<const float4*>(wp + (size_t)(kbase + c0 + b_k) * p.cout_pad + b_n);
                __syncthreads();
                As[a_kq + 0][a_pix] = av.x;
                As[a_kq + 1][a_pix] = av.y;
                As[a_kq + 2][a_pix] = av.z;
                As[a_kq + 3][a_pix] = av.w;
                *reinterpret_cast<float4*>(&Bs[b_k][(tid % 16) * 4]) = bv;
                __syncthreads();
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                    const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                    const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                    const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                }
            }
        }
    }

    const int n = n0 + tx * 4;
    if (n >= p.cout_pad) return;
    const float4 bias4 = *reinterpret_cast<const float4*>(p.bias + n);
    float* dst = static_cast<float*>(p.dst);
    const float* res = static_cast<const float*>(p.residual);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        float4 v = make_float4(acc[i][0] + bias4.x, acc[i][1] + bias4.y, acc[i][2] + bias4.z, acc[i][3] + bias4.w);
        if (res) {
            const float4 r = *reinterpret_cast<const float4*>(res + (size_t)m * p.res_C + n);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        const int img = m / (p.OH * p.OW), r = m % (p.OH * p.OW);
        const int oy = (r / p.OW) * p.dst_stride + p.dst_off_y, ox = (r % p.OW) * p.dst_stride + p.dst_off_x;
        *reinterpret_cast<float4*>(dst + ((size_t)(img * p.dst_H + oy) * p.dst_W + ox) * p.dst_C + n) = v;
    }
}

void launch_conv_simt(const ConvDev& p, cudaStream_t st) {
    const int M = p.B * p.OH * p.OW;
    dim3 grid((M + 63) / 64, (p.cout_pad + 63) / 64);
    conv_simt_kernel<<<grid, 256, 0, st>>>(p);
    SPB_CHECK_LAUNCH();
}

// ------------------------------------------------------------------------------------------------
// NHWC (any element type, Cs stored channels) -> NCHW fp32 (first C channels)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst,
                                                            int HW, int Cs, int C) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;   // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int p = p0 + i, c = c0 + tx;
        float v = 0.f;
        if (p < HW && c < C) v = to_float(src[((size_t)b * HW + p) * Cs + c]);
        tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, p = p0 + tx;
        if (p < HW && c < C) dst[((size_t)b * C + c) * HW + p] = tile[tx][i];
    }
}

void launch_nhwc_to_nchw(const void* src, int src_type, int B, int HW, int Cs, int C, float* dst, cudaStream_t st) {
    dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
    if (src_type == PREC_FP32) nhwc_to_nchw_kernel<float><<<grid, 256, 0, st>>>((const float*)src, dst, HW, Cs, C);
    else if (src_type == PREC_FP16) nhwc_to_nchw_kernel<__half><<<grid, 256, 0, st>>>((const __half*)src, dst, HW, Cs, C);
    else nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, dst, HW, Cs, C);
    SPB_CHECK_LAUNCH();
}

// Split-layout source (common.cuh, SegDev): channel c of a pixel = hi + lo, hi at (c / 32) * 64 + c % 32, lo 32 further.
template <typename T>
__global__ void __launch_bounds__(256) split_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int HW, int Cs, int C) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
    for (int i = ty; i < 32; i += 8) {
        const int p = p0 + i, c = c0 + tx;
        float v = 0.f;
        if (p < HW && c < C) {
            const T* px = src + ((size_t)b * HW + p) * Cs + (c / 32) * 64 + c % 32;
            v = to_float(px[0]) + to_float(px[32]);
        }
        tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, p = p0 + tx;
        if (p < HW && c < C) dst[((size_t)b * C + c) * HW + p] = tile[tx][i];
    }
}

void launch_split_to_nchw(const void* src, int src_type, int B, int HW, int Cs, int C, float* dst, cudaStream_t st) {
    dim3 grid((HW + 31) / 32, (C + 31) / 32, B);
    if (src_type == PREC_FP16) split_to_nchw_kernel<__half><<<grid, 256, 0, st>>>((const __half*)src, dst, HW, Cs, C);
    else if (src_type == PREC_BF16) split_to_nchw_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, dst, HW, Cs, C);
    else throw std::invalid_argument("split layout: 16-bit element types only");
    SPB_CHECK_LAUNCH();
}

// one thread = 16 bytes (8 hi values); a pixel has C / 8 of them
__global__ void __launch_bounds__(256) split_hi_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long total, int c8) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long pix = i / c8;
    const int k = (int)(i % c8);                       // 16-byte piece of the plain pixel: chunk k / 4, piece k % 4 of its hi half
    dst[i] = __ldg(src + pix * (2 * c8) + (k >> 2) * 8 + (k & 3));
}

void launch_split_hi(const void* src, void* dst, long npix, int C, cudaStream_t st) {
    if (C % 32) throw std::invalid_argument("split layout: channels must be a multiple of 32");
    const long total = npix * (C / 8);
    split_hi_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const uint4*)src, (uint4*)dst, total, C / 8);
    SPB_CHECK_LAUNCH();
}

__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (float)src[i] / 255.f;
}

void launch_u8_to_f32(const uint8_t* src, float* dst, long n, cudaStream_t st) {
    u8_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, dst, n);
    SPB_CHECK_LAUNCH();
}

}  // namespace spb200
