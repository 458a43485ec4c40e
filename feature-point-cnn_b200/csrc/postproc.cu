// Memory-bound post-processing kernels of the spb200 engine (sm_100a).
//
//  K3 heatmap_kernel      softmax-with-epsilon over 65 channels, drop dustbin, depth-to-space
//                         (reference python/src/superpoint.py:111-114, python/src/netutils.py:64-75)
//  K4 (nms.cu)            exact parallel form of the reference's greedy grid NMS
//  K5 (nms.cu)            block-wide LSD radix sort of the survivors by descending confidence
//                         (python/src/netutils.py:92-93) + top-k truncation, fused with the last NMS rounds
//  K6 sample_desc_kernel  bilinear sampling (align_corners=True) + L2 normalisation
//                         (python/src/netutils.py:103-121), one warp per keypoint, 128-bit loads
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "kernels.h"
#include "softmax_cell.cuh"
#include "sortkey.cuh"

namespace spb200 {

// ================================================================================================
// K3: logits -> full-resolution heatmap
// ================================================================================================
constexpr int kHeatCells = 32;     // cells of one cell-row per block: 8 lanes per cell, 256 threads

__global__ void __launch_bounds__(256)
heatmap_kernel(const float* __restrict__ logits, long batch_stride, long chan_stride, long cell_stride, int Hc, int Wc,
               float* __restrict__ heat) {
    const int tid = threadIdx.x, cell = tid >> 3, j = tid & 7;
    const int b = blockIdx.z, i = blockIdx.y, j0 = blockIdx.x * kHeatCells;
    const bool in = j0 + cell < Wc;                                   // whole octets are in or out
    const float* p = logits + (size_t)b * batch_stride + (size_t)(i * Wc + min(j0 + cell, Wc - 1)) * cell_stride;
    float l[8], l64, h[8];
    if (chan_stride == 1 && (cell_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && (batch_stride & 3) == 0) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p + 8 * j)), c = __ldg(reinterpret_cast<const float4*>(p + 8 * j) + 1);
        l[0] = a.x; l[1] = a.y; l[2] = a.z; l[3] = a.w; l[4] = c.x; l[5] = c.y; l[6] = c.z; l[7] = c.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) l[k] = __ldg(p + (size_t)(8 * j + k) * chan_stride);
    }
    l64 = __ldg(p + (size_t)64 * chan_stride);
    softmax_cell_octet(l, l64, j, h);
    if (!in) return;
    const int H = Hc * 8, W = Wc * 8;
    float* o = heat + ((size_t)b * H + (size_t)i * 8 + j) * W + (size_t)(j0 + cell) * 8;      // pixel row j of the cell
    if ((reinterpret_cast<uintptr_t>(heat) & 15) == 0) {
        reinterpret_cast<float4*>(o)[0] = make_float4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<float4*>(o)[1] = make_float4(h[4], h[5], h[6], h[7]);
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = h[k];
    }
}

// restore_prob_map alone (python/src/netutils.py:64-75): the 65-channel tensor is already softmaxed; drop the dustbin,
// pixel (8 i + c / 8, 8 j + c % 8) <- channel c of cell (i, j).  One thread per output pixel: a warp writes 128 contiguous
// bytes and reads four 32-byte runs of eight cells from each of eight channel planes.
__global__ void __launch_bounds__(256)
depth_to_space_kernel(const float* __restrict__ src, int Hc, int Wc, float* __restrict__ heat, long total) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int W = Wc * 8, H = Hc * 8;
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long b = i / ((long)W * H);
    const int c = (y & 7) * 8 + (x & 7);
    heat[i] = __ldg(src + ((b * 65 + c) * Hc + (y >> 3)) * Wc + (x >> 3));
}

void launch_depth_to_space(const float* softmax_nchw, int B, int Hc, int Wc, float* heat, cudaStream_t st) {
    const long total = (long)B * Hc * Wc * 64;
    depth_to_space_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(softmax_nchw, Hc, Wc, heat, total);
    SPB_CHECK_LAUNCH();
}

// Keypoint counts of a chunk, written straight into pinned host memory (mapped into the device's address space under
// unified addressing): a cudaMemcpyAsync would queue behind the large descriptor downloads of the previous batch on the
// device-to-host copy engine and stall the compute stream that issued it.
__global__ void counts_to_host_kernel(const int* __restrict__ src, volatile int* dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
    __threadfence_system();
}

void launch_counts_to_host(const int* src, int* dst_pinned, int n, cudaStream_t st) {
    counts_to_host_kernel<<<(n + 127) / 128, 128, 0, st>>>(src, dst_pinned, n);
    SPB_CHECK_LAUNCH();
}

// exp(l) * inv: the multiplication softmax_cell.cuh ends with, so the result equals heatmap_kernel's bit for bit
__global__ void __launch_bounds__(256) heat_scale_kernel(float* __restrict__ heat, const float* __restrict__ inv, int H, int W, long total4) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;      // four pixels of one row (W is a multiple of 8: one cell)
    pdl_trigger();
    pdl_wait();
    if (i >= total4) return;
    const int w4 = W / 4;
    const int x = (int)(i % w4) * 4;
    const long row = i / w4;
    const int y = (int)(row % H);
    const long b = row / H;
    const float s = __ldg(inv + (b * (H / 8) + (y >> 3)) * (W / 8) + (x >> 3));
    float4 v = reinterpret_cast<float4*>(heat)[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    reinterpret_cast<float4*>(heat)[i] = v;
}

void launch_heat_scale(float* heat, const float* inv, int B, int H, int W, cudaStream_t st) {
    const long total4 = (long)B * H * (W / 4);
    launch_pdl(heat_scale_kernel, dim3((unsigned)((total4 + 255) / 256)), dim3(256), 0, st, heat, inv, H, W, total4);
}

void launch_heatmap(const float* logits, long batch_stride, long chan_stride, long cell_stride, int B, int Hc, int Wc,
                    float* heat, cudaStream_t st) {
    dim3 grid((Wc + kHeatCells - 1) / kHeatCells, Hc, B);
    heatmap_kernel<<<grid, 256, 0, st>>>(logits, batch_stride, chan_stride, cell_stride, Hc, Wc, heat);
    SPB_CHECK_LAUNCH();
}

// ================================================================================================
// K6: descriptors.  One warp per keypoint; each lane owns 4 consecutive channels per 128-channel
// slab.  grid_sample(align_corners=True): ix = ((gx + 1)/2)*(Wc-1) with gx = x/(W/2) - 1 evaluated in
// double and rounded to float, exactly as the reference builds its sampling grid
// (python/src/netutils.py:110-115); those per-coordinate values come from a small host-built table.  The L2 norm has no epsilon (netutils.py:120): an all-zero sample
// yields NaN there and here.
// ================================================================================================
template <typename T>
__device__ __forceinline__ void load4(const T* p, long chan_stride, float (&v)[4]);

template <>
__device__ __forceinline__ void load4<float>(const float* p, long cs, float (&v)[4]) {
    if (cs == 1) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __ldg(p + i * cs);
    }
}
template <>
__device__ __forceinline__ void load4<__half>(const __half* p, long cs, float (&v)[4]) {
    if (cs == 1) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
        const __half2 a = *reinterpret_cast<const __half2*>(&t.x), b = *reinterpret_cast<const __half2*>(&t.y);
        v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __half2float(p[i * cs]);
    }
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, long cs, float (&v)[4]) {
    if (cs == 1) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __bfloat162float(p[i * cs]);
    }
}

// gtab: per-coordinate sampling positions precomputed on the host exactly as the reference does
// (double division, then float): gtab[x] = ix for x < W, gtab[W + y] = iy for y < H.
// OUT16: the unit vector is stored as fp16 (spb200_set_descriptor_format), otherwise fp32 like the reference's
__device__ __forceinline__ void store_desc4(void* base, size_t elem, bool out16, float a, float b, float c, float d) {
    if (out16) {
        const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
        __stcs(reinterpret_cast<uint2*>(static_cast<__half*>(base) + elem),
               make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi)));
    } else {
        __stcs(reinterpret_cast<float4*>(static_cast<float*>(base) + elem), make_float4(a, b, c, d));
    }
}

template <typename T, bool OUT16>
__global__ void __launch_bounds__(256)
sample_desc_kernel(const T* __restrict__ map, long batch_stride, long chan_stride, long cell_stride, int D, int Hc,
                   int Wc, int W, const float* __restrict__ gtab, int cap, const int* __restrict__ count,
                   const int* __restrict__ xy, void* __restrict__ out) {
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
    const int b = blockIdx.y;
    const int i = blockIdx.x * 8 + warp;
    if (i >= min(__ldg(count + b), cap)) return;
    const int2 pt = __ldg(reinterpret_cast<const int2*>(xy) + (size_t)b * cap + i);
    // caller-supplied coordinates (spb200_sample_descriptors) are clamped into the image: never an out-of-bounds read
    const float ix = __ldg(gtab + min(max(pt.x, 0), W - 1)), iy = __ldg(gtab + W + min(max(pt.y, 0), Hc * 8 - 1));
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = (fx0 + 1.f) - ix, wy0 = (fy0 + 1.f) - iy;
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float wgt[4] = {wx0 * wy0, wx1 * wy0, wx0 * wy1, wx1 * wy1};
    const int cx[4] = {x0, x0 + 1, x0, x0 + 1}, cy[4] = {y0, y0, y0 + 1, y0 + 1};
    const T* mb = map + (size_t)b * batch_stride;
    const size_t o = ((size_t)b * cap + i) * D;

    float ss = 0.f;
    constexpr int kMaxSlabs = 4;                      // D <= 512
    float acc[kMaxSlabs][4];
    const int nslab = (D + 127) / 128;
#pragma unroll
    for (int s = 0; s < kMaxSlabs; ++s) {
        if (s >= nslab) break;
        const int c = s * 128 + lane * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[s][q] = 0.f;
        if (c < D) {
            float v[4][4];
            bool ok[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {             // all four corner loads in flight before any use
                ok[k] = !(cx[k] < 0 || cx[k] >= Wc || cy[k] < 0 || cy[k] >= Hc);     // zeros padding
                if (ok[k]) load4<T>(mb + (size_t)(cy[k] * Wc + cx[k]) * cell_stride + (size_t)c * chan_stride, chan_stride, v[k]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!ok[k]) continue;
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[s][q] = fmaf(wgt[k], v[k][q], acc[s][q]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) ss = fmaf(acc[s][q], acc[s][q], ss);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    const float inv = 1.f / sqrtf(ss);                // 0-vector -> inf * 0 = NaN, like the reference's 0/0
#pragma unroll
    for (int s = 0; s < kMaxSlabs; ++s) {
        if (s >= nslab) break;
        const int c = s * 128 + lane * 4;
        if (c < D) store_desc4(out, o + c, OUT16, acc[s][0] * inv, acc[s][1] * inv, acc[s][2] * inv, acc[s][3] * inv);
    }
}

// Fast path of K6 for the engine's own descriptor map (NHWC, 16-bit, 128 channels): a HALF-warp per keypoint
// (16 lanes x 8 channels = one 16-byte load per corner per lane), two keypoints per half-warp in flight, and a
// grid of (blocks per image, images) whose blocks stride over the image's keypoints - no empty blocks however
// large `cap` is.  Same arithmetic and summation order as sample_desc_kernel.
constexpr int kDescBlocksPerImage = 24;

template <typename T, bool OUT16>
__global__ void __launch_bounds__(256)
sample_desc128_kernel(const T* __restrict__ map, long batch_stride, int Hc, int Wc, int W, const float* __restrict__ gtab,
                      int cap, const int* __restrict__ count, const int* __restrict__ xy, void* __restrict__ out) {
    const int hw = threadIdx.x >> 4, hl = threadIdx.x & 15;
    const int b = blockIdx.y;
    pdl_trigger();
    pdl_wait();                                                // counts and keypoints of the NMS kernel before it
    const int n = min(__ldg(count + b), cap);
    const T* mb = map + (size_t)b * batch_stride + hl * 8;
    const int2* pts = reinterpret_cast<const int2*>(xy) + (size_t)b * cap;
    const int stride = gridDim.x * 16;
    // the loop bound is taken on the warp's first keypoint so that both half-warps leave together (full-mask shuffles)
    for (int base = blockIdx.x * 16 + (hw & ~1); base < n; base += 2 * stride) {
        const int i0 = base + (hw & 1);
        uint4 v[2][4];
        float wgt[2][4];
        bool live[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = i0 + u * stride;
            live[u] = i < n;
#pragma unroll
            for (int k = 0; k < 4; ++k) { v[u][k] = make_uint4(0u, 0u, 0u, 0u); wgt[u][k] = 0.f; }
            if (live[u]) {
                const int2 pt = __ldg(pts + i);
                // caller-supplied coordinates (spb200_sample_descriptors) are clamped into the image: never an out-of-bounds read
                const float ix = __ldg(gtab + min(max(pt.x, 0), W - 1)), iy = __ldg(gtab + W + min(max(pt.y, 0), Hc * 8 - 1));
                const float fx0 = floorf(ix), fy0 = floorf(iy);
                const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = (fx0 + 1.f) - ix, wy0 = (fy0 + 1.f) - iy;
                const int x0 = (int)fx0, y0 = (int)fy0;
                wgt[u][0] = wx0 * wy0; wgt[u][1] = wx1 * wy0; wgt[u][2] = wx0 * wy1; wgt[u][3] = wx1 * wy1;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int cx = x0 + (k & 1), cy = y0 + (k >> 1);
                    if (cx < 0 || cx >= Wc || cy < 0 || cy >= Hc) wgt[u][k] = 0.f;     // zeros padding: the corner is skipped
                    else v[u][k] = __ldg(reinterpret_cast<const uint4*>(mb + (size_t)(cy * Wc + cx) * 128));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float acc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t w4[4] = {v[u][k].x, v[u][k].y, v[u][k].z, v[u][k].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float2 f;
                    if (sizeof(T) == 2 && std::is_same<T, __half>::value) f = __half22float2(*reinterpret_cast<const __half2*>(&w4[e]));
                    else f = make_float2(__uint_as_float(w4[e] << 16), __uint_as_float(w4[e] & 0xffff0000u));
                    // a corner outside the map contributes nothing (its weight was zeroed and its value is 0)
                    acc[2 * e] = fmaf(wgt[u][k], f.x, acc[2 * e]);
                    acc[2 * e + 1] = fmaf(wgt[u][k], f.y, acc[2 * e + 1]);
                }
            }
            float ss = 0.f;
#pragma unroll
            for (int q = 0; q < 8; q += 4) {          // same partial-sum grouping as the generic kernel: four channels at a time
                ss = fmaf(acc[q], acc[q], ss); ss = fmaf(acc[q + 1], acc[q + 1], ss);
                ss = fmaf(acc[q + 2], acc[q + 2], ss); ss = fmaf(acc[q + 3], acc[q + 3], ss);
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
            if (live[u]) {
                const float inv = 1.f / sqrtf(ss);    // 0-vector -> NaN, like the reference's 0/0
                const size_t o = ((size_t)b * cap + (i0 + u * stride)) * 128 + hl * 8;
                store_desc4(out, o, OUT16, acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
                store_desc4(out, o + 4, OUT16, acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
            }
        }
    }
}

void launch_sample_descriptors(const void* map, int map_type, long batch_stride, long chan_stride, long cell_stride,
                               int B, int D, int Hc, int Wc, int W, const float* gtab, int cap, const int* count,
                               const int* xy, void* out, int out_fp16, cudaStream_t st) {
    if (D % 4 != 0 || D > 512) throw std::invalid_argument("descriptor dimension must be a multiple of 4, at most 512");
    if (cap <= 0 || B <= 0) return;
    if (map_type != PREC_FP32 && D == 128 && chan_stride == 1 && cell_stride == 128) {
        static const int blocks = [] { const char* e = std::getenv("SPB200_DESC_BLOCKS"); return e ? std::max(1, atoi(e)) : kDescBlocksPerImage; }();
        dim3 g(blocks, B);
        if (map_type == PREC_FP16) {
            if (out_fp16) launch_pdl(sample_desc128_kernel<__half, true>, g, dim3(256), 0, st, (const __half*)map, batch_stride, Hc, Wc, W, gtab, cap, count, xy, out);
            else launch_pdl(sample_desc128_kernel<__half, false>, g, dim3(256), 0, st, (const __half*)map, batch_stride, Hc, Wc, W, gtab, cap, count, xy, out);
        } else {
            if (out_fp16) launch_pdl(sample_desc128_kernel<__nv_bfloat16, true>, g, dim3(256), 0, st, (const __nv_bfloat16*)map, batch_stride, Hc, Wc, W, gtab, cap, count, xy, out);
            else launch_pdl(sample_desc128_kernel<__nv_bfloat16, false>, g, dim3(256), 0, st, (const __nv_bfloat16*)map, batch_stride, Hc, Wc, W, gtab, cap, count, xy, out);
        }
        return;
    }
    dim3 grid((cap + 7) / 8, B);
#define SPB_SAMPLE(T, O) sample_desc_kernel<T, O><<<grid, 256, 0, st>>>((const T*)map, batch_stride, chan_stride, cell_stride, D, Hc, Wc, W, gtab, cap, count, xy, out)
    if (map_type == PREC_FP32) { if (out_fp16) SPB_SAMPLE(float, true); else SPB_SAMPLE(float, false); }
    else if (map_type == PREC_FP16) { if (out_fp16) SPB_SAMPLE(__half, true); else SPB_SAMPLE(__half, false); }
    else { if (out_fp16) SPB_SAMPLE(__nv_bfloat16, true); else SPB_SAMPLE(__nv_bfloat16, false); }
#undef SPB_SAMPLE
    SPB_CHECK_LAUNCH();
}

}  // namespace spb200
