// Frame loaders in front of the path, on the device.
//
//  * preprocess_u8:  the C++ demo's loader, cpp/src/camera.cc:12-23 - cv::resize(frame, Size(W, H)) (INTER_LINEAR on 8-bit
//    pixels: OpenCV's 11-bit fixed-point scheme), cvtColor(BGR2GRAY) (15-bit fixed point); the convertTo(CV_32FC1, 1/255)
//    that follows is what spb200_detect_u8 applies.  Bit-exact with OpenCV 4.x (tests/golden/preproc_kat.npz).
//  * preprocess_f32: the Python demo's loader, python/src/inference.py:72-85 (make_query_image) - BGR -> RGB, ratio-preserving
//    INTER_LINEAR resize in float, centre crop - followed by InferenceWrapper.prepare_input's HWC -> CHW (inferencewrapper.py:70-81).
//
// Both are HBM-bound gathers: one thread per output pixel, interpolation tables built on the host with OpenCV's own
// arithmetic (double scale, float fraction, round-half-even coefficients).
#include <cmath>
#include <vector>

#include "kernels.h"

namespace spb200 {

__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ src, int h, int w, int C, const int* __restrict__ tab, uint8_t* __restrict__ dst,
                     int H, int W, long total) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int X = (int)(i % W), Y = (int)((i / W) % H);
    const long b = i / ((long)W * H);
    // tab: [W] x0, [W] x1, [W] a0, [W] a1, [H] y0, [H] y1, [H] b0, [H] b1
    const int x0 = __ldg(tab + X), x1 = __ldg(tab + W + X), a0 = __ldg(tab + 2 * W + X), a1 = __ldg(tab + 3 * W + X);
    const int* ty = tab + 4 * W;
    const int y0 = __ldg(ty + Y), y1 = __ldg(ty + H + Y), b0 = __ldg(ty + 2 * H + Y), b1 = __ldg(ty + 3 * H + Y);
    const uint8_t* r0 = src + ((size_t)b * h + y0) * w * C;
    const uint8_t* r1 = src + ((size_t)b * h + y1) * w * C;
    int v[3];
    for (int c = 0; c < C; ++c) {
        const int s0 = (int)r0[x0 * C + c] * a0 + (int)r0[x1 * C + c] * a1;      // horizontal pass, scaled by 2048
        const int s1 = (int)r1[x0 * C + c] * a0 + (int)r1[x1 * C + c] * a1;
        v[c] = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;    // OpenCV's VResizeLinear for 8-bit pixels
    }
    // BGR -> gray: (B * 3735 + G * 19235 + R * 9798 + 2^14) >> 15 (OpenCV's 15-bit coefficients)
    dst[i] = (uint8_t)(C == 3 ? (v[0] * 3735 + v[1] * 19235 + v[2] * 9798 + (1 << 14)) >> 15 : v[0]);
}

__global__ void __launch_bounds__(256)
preprocess_f32_kernel(const float* __restrict__ src, int h, int w, const int* __restrict__ itab, const float* __restrict__ ftab,
                      float* __restrict__ dst, int H, int W, long total) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int X = (int)(i % W), Y = (int)((i / W) % H);
    const long b = i / ((long)W * H);
    const int x0 = __ldg(itab + X), x1 = __ldg(itab + W + X), y0 = __ldg(itab + 2 * W + Y), y1 = __ldg(itab + 2 * W + H + Y);
    const float fx = __ldg(ftab + X), fy = __ldg(ftab + W + Y);
    const float a0 = 1.f - fx, b0 = 1.f - fy;
    const float* r0 = src + ((size_t)b * h + y0) * w * 3;
    const float* r1 = src + ((size_t)b * h + y1) * w * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float s0 = __fadd_rn(__fmul_rn(r0[x0 * 3 + c], a0), __fmul_rn(r0[x1 * 3 + c], fx));      // no contraction: OpenCV's
        const float s1 = __fadd_rn(__fmul_rn(r1[x0 * 3 + c], a0), __fmul_rn(r1[x1 * 3 + c], fx));      // two-pass arithmetic
        dst[((size_t)b * 3 + (2 - c)) * H * W + (size_t)Y * W + X] = __fadd_rn(__fmul_rn(s0, b0), __fmul_rn(s1, fy));   // BGR -> RGB planes
    }
}

// OpenCV's linear-resize source positions (resize.cpp): scale = 1 / ((double) dst / src), f = (float)((d + 0.5) scale - 0.5),
// s = floor(f), f -= s.  Horizontally s is clamped with f = 0; vertically the two ROWS are clamped and f is kept.
static void axis_tables(int dst, int src, bool horizontal, int first, int count, std::vector<int>& p0, std::vector<int>& p1,
                        std::vector<float>& frac) {
    const double scale = 1.0 / ((double)dst / (double)src);
    p0.resize(count); p1.resize(count); frac.resize(count);
    for (int k = 0; k < count; ++k) {
        const int d = first + k;
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= (float)s;
        if (horizontal) {
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= src - 1) { s = src - 1; f = 0.f; }
            p0[k] = s; p1[k] = std::min(s + 1, src - 1);
        } else {
            p0[k] = std::min(std::max(s, 0), src - 1); p1[k] = std::min(std::max(s + 1, 0), src - 1);
        }
        frac[k] = f;
    }
}

size_t preprocess_u8_table_ints(int H, int W) { return (size_t)4 * (H + W); }

void build_preprocess_u8_table(int h, int w, int H, int W, int* tab) {
    std::vector<int> p0, p1;
    std::vector<float> f;
    axis_tables(W, w, true, 0, W, p0, p1, f);
    for (int X = 0; X < W; ++X) {
        tab[X] = p0[X]; tab[W + X] = p1[X];
        tab[2 * W + X] = (int)std::lrintf((1.f - f[X]) * 2048.f);
        tab[3 * W + X] = (int)std::lrintf(f[X] * 2048.f);
    }
    int* ty = tab + 4 * W;
    axis_tables(H, h, false, 0, H, p0, p1, f);
    for (int Y = 0; Y < H; ++Y) {
        ty[Y] = p0[Y]; ty[H + Y] = p1[Y];
        ty[2 * H + Y] = (int)std::lrintf((1.f - f[Y]) * 2048.f);
        ty[3 * H + Y] = (int)std::lrintf(f[Y] * 2048.f);
    }
}

void launch_preprocess_u8(const uint8_t* src, int B, int h, int w, int C, const int* tab_dev, uint8_t* dst, int H, int W, cudaStream_t st) {
    const long total = (long)B * H * W;
    preprocess_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, h, w, C, tab_dev, dst, H, W, total);
    SPB_CHECK_LAUNCH();
}

// make_query_image: the frame is resized to (int(w s), int(h s)) with s = max(H / h, W / w) and centre-cropped to W x H
void build_preprocess_f32_table(int h, int w, int H, int W, int* itab, float* ftab) {
    const double sh = (double)H / h, sw = (double)W / w, s = sh > sw ? sh : sw;
    const int nw = (int)(w * s), nh = (int)(h * s);
    const int cx = nw / 2 - W / 2, cy = nh / 2 - H / 2;
    if (nw < W || nh < H || cx < 0 || cy < 0) throw std::invalid_argument("preprocess: the resized frame is smaller than the crop");
    std::vector<int> p0, p1;
    std::vector<float> f;
    axis_tables(nw, w, true, cx, W, p0, p1, f);
    for (int X = 0; X < W; ++X) { itab[X] = p0[X]; itab[W + X] = p1[X]; ftab[X] = f[X]; }
    axis_tables(nh, h, false, cy, H, p0, p1, f);
    for (int Y = 0; Y < H; ++Y) { itab[2 * W + Y] = p0[Y]; itab[2 * W + H + Y] = p1[Y]; ftab[W + Y] = f[Y]; }
}

void launch_preprocess_f32(const float* src, int B, int h, int w, const int* itab_dev, const float* ftab_dev, float* dst, int H, int W,
                           cudaStream_t st) {
    const long total = (long)B * H * W;
    preprocess_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, h, w, itab_dev, ftab_dev, dst, H, W, total);
    SPB_CHECK_LAUNCH();
}

}  // namespace spb200
