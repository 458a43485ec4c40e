// Homography adaptation on the device (reference homography_adaptation, python/src/homographies.py:250-324; caller
// InferenceWrapper.run_with_homography_adaptation, python/src/inferencewrapper.py:48-68, and the COCO pseudo-labelling
// job python/src/preprocess_coco.py:64-74): the detector runs on `num` random warps of every image and the warped
// heatmaps are projected back and averaged where they are valid.  The random homographies are sampled by the caller
// (homographies.py:79-196) and arrive as flattened 8-coefficient transforms.
//
// Conventions restated from the reference's dependencies: torchvision's functional_tensor.perspective builds the
// sampling grid ((c0 x + c1 y + c2) / (c6 x + c7 y + 1), (c3 x + c4 y + c5) / (...)) at the pixel centres
// (x + 0.5, y + 0.5), normalised by the half extents, and torch's grid_sample(align_corners=False, padding zeros)
// reads it back as pixel index ((g + 1) * size - 1) / 2: bilinear for images and heatmaps, nearest (round half to
// even) for the validity maps, which are then eroded with OpenCV's elliptical structuring element of size
// 2 * margin, anchored at its centre, with a constant zero border (homographies.py:239-247).
//
//   ha_valid_raw_kernel   count_k = warp of ones by H_k^-1, mask_k = warp of ones by H_k (nearest), one byte per pixel
//   ha_erode_kernel       erosion of both from a shared-memory tile
//   ha_warp_kernel        warped image k of every batch image (bilinear), the detector's next input
//   ha_aggregate_kernel   per output pixel: sum / max over k of bilinear(prob_k * mask_k at H_k^-1) * count_k, the
//                         count total, mean or max, zero where fewer than num / 3 views saw the pixel
#include "kernels.h"

namespace spb200 {

__device__ __forceinline__ void ha_source(const float* __restrict__ c, int x, int y, int W, int H, float& ix, float& iy) {
    const float fx = (float)x + 0.5f, fy = (float)y + 0.5f;
    const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
    const float den = c[6] * fx + c[7] * fy + 1.0f;
    const float gx = (c[0] / hw * fx + c[1] / hw * fy + c[2] / hw) / den - 1.0f;
    const float gy = (c[3] / hh * fx + c[4] / hh * fy + c[5] / hh) / den - 1.0f;
    ix = ((gx + 1.0f) * (float)W - 1.0f) / 2.0f;
    iy = ((gy + 1.0f) * (float)H - 1.0f) / 2.0f;
}

// maps: [2 * num][H][W] bytes; map 2k = count_k (inverse transform), 2k + 1 = mask_k (forward transform)
__global__ void __launch_bounds__(256) ha_valid_raw_kernel(const float* __restrict__ coeffs, int num, int H, int W,
                                                           uint8_t* __restrict__ maps) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int m = blockIdx.z;                                   // 0 .. 2 num - 1
    if (x >= W || y >= H) return;
    const int k = m >> 1;
    const float* c = coeffs + ((m & 1) ? k : num + k) * 8;      // forward coefficients first, then the inverses
    float ix, iy;
    ha_source(c, x, y, W, H, ix, iy);
    const float rx = nearbyintf(ix), ry = nearbyintf(iy);
    maps[((size_t)m * H + y) * W + x] = (rx >= 0.f && rx < (float)W && ry >= 0.f && ry < (float)H) ? 1 : 0;
}

struct HaEllipse {
    int ksize;                  // 2 * margin
    signed char j1[64], j2[64]; // row i of the element holds ones in columns [j1, j2)
};

constexpr int kHaTile = 32;

__global__ void __launch_bounds__(kHaTile * kHaTile) ha_erode_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                                     int H, int W, const __grid_constant__ HaEllipse el) {
    extern __shared__ uint8_t ha_tile[];                        // [kHaTile + ksize][kHaTile + ksize]
    const int ks = el.ksize, r = ks / 2, pitch = kHaTile + ks;
    const int x0 = blockIdx.x * kHaTile, y0 = blockIdx.y * kHaTile;
    const uint8_t* s = src + (size_t)blockIdx.z * H * W;
    for (int i = threadIdx.x; i < pitch * pitch; i += kHaTile * kHaTile) {
        const int ty = i / pitch, tx = i % pitch;
        const int gy = y0 + ty - r, gx = x0 + tx - r;           // tile element (ty, tx) = source pixel (y0 + ty - r, ...)
        ha_tile[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? s[(size_t)gy * W + gx] : 0;   // constant zero border
    }
    __syncthreads();
    const int lx = threadIdx.x % kHaTile, ly = threadIdx.x / kHaTile;
    const int x = x0 + lx, y = y0 + ly;
    if (x >= W || y >= H) return;
    unsigned all = 1u;
    for (int i = 0; i < ks; ++i) {                              // element (i, j) covers source (y + i - r, x + j - r)
        const uint8_t* row = ha_tile + (ly + i) * pitch + lx;
        for (int j = el.j1[i]; j < el.j2[i]; ++j) all &= row[j];
    }
    dst[((size_t)blockIdx.z * H + y) * W + x] = (uint8_t)all;
}

// img [B][C][H][W] -> out [B][C][H][W], out(x, y) = bilinear img at the transform c of (x, y), zeros outside
__global__ void __launch_bounds__(256) ha_warp_kernel(const float* __restrict__ img, const float* __restrict__ c, int C, int H,
                                                      int W, float* __restrict__ out) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    float ix, iy;
    ha_source(c, x, y, W, H, ix, iy);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = (fx0 + 1.f) - ix, wy0 = (fy0 + 1.f) - iy;
    const int x0 = (int)fx0, y0 = (int)fy0;
    const bool okx0 = x0 >= 0 && x0 < W, okx1 = x0 + 1 >= 0 && x0 + 1 < W, oky0 = y0 >= 0 && y0 < H, oky1 = y0 + 1 >= 0 && y0 + 1 < H;
    for (int ch = 0; ch < C; ++ch) {
        const float* p = img + ((size_t)blockIdx.z * C + ch) * H * W;
        float v = 0.f;
        if (oky0 && okx0) v += p[(size_t)y0 * W + x0] * (wx0 * wy0);
        if (oky0 && okx1) v += p[(size_t)y0 * W + x0 + 1] * (wx1 * wy0);
        if (oky1 && okx0) v += p[(size_t)(y0 + 1) * W + x0] * (wx0 * wy1);
        if (oky1 && okx1) v += p[(size_t)(y0 + 1) * W + x0 + 1] * (wx1 * wy1);
        out[(((size_t)blockIdx.z * C + ch) * H + y) * W + x] = v;
    }
}

// probs [(num + 1)][B][H][W] (view 0 = the image itself), maps eroded [2 num][H][W], coeffs [2 num][8]
__global__ void __launch_bounds__(256) ha_aggregate_kernel(const float* __restrict__ probs, const uint8_t* __restrict__ maps,
                                                           const float* __restrict__ coeffs, int num, int B, int H, int W,
                                                           int use_max, float* __restrict__ out) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    if (x >= W || y >= H) return;
    const size_t plane = (size_t)H * W;
    float sum = probs[(size_t)b * plane + (size_t)y * W + x], mx = sum, cnt = 1.f;
    for (int k = 0; k < num; ++k) {
        const float count = (float)maps[(size_t)(2 * k) * plane + (size_t)y * W + x];
        float v = 0.f;
        if (count != 0.f) {
            float ix, iy;
            ha_source(coeffs + (num + k) * 8, x, y, W, H, ix, iy);
            const float fx0 = floorf(ix), fy0 = floorf(iy);
            const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = (fx0 + 1.f) - ix, wy0 = (fy0 + 1.f) - iy;
            const int x0 = (int)fx0, y0 = (int)fy0;
            const float* p = probs + ((size_t)(k + 1) * B + b) * plane;
            const uint8_t* m = maps + (size_t)(2 * k + 1) * plane;
            auto tap = [&](int yy, int xx, float w) {
                if (yy >= 0 && yy < H && xx >= 0 && xx < W) v += p[(size_t)yy * W + xx] * (float)m[(size_t)yy * W + xx] * w;
            };
            tap(y0, x0, wx0 * wy0);
            tap(y0, x0 + 1, wx1 * wy0);
            tap(y0 + 1, x0, wx0 * wy1);
            tap(y0 + 1, x0 + 1, wx1 * wy1);
            v *= count;
        }
        sum += v;
        mx = fmaxf(mx, v);
        cnt += count;
    }
    const float agg = use_max ? mx : sum / cnt;
    out[(size_t)b * plane + (size_t)y * W + x] = cnt >= (float)(num / 3) ? agg : 0.f;
}

void launch_ha_valid_maps(const float* coeffs, int num, int H, int W, int margin, uint8_t* raw, uint8_t* eroded, cudaStream_t st) {
    if (margin < 0 || margin > 32) throw std::invalid_argument("valid_border_margin must be in [0, 32]");
    dim3 grid((W + 31) / 32, (H + 7) / 8, 2 * num);
    ha_valid_raw_kernel<<<grid, 256, 0, st>>>(coeffs, num, H, W, margin ? raw : eroded);
    SPB_CHECK_LAUNCH();
    if (!margin) return;
    // cv2.getStructuringElement(MORPH_ELLIPSE, (2 margin, 2 margin)): row i spans c - dx .. c + dx,
    // dx = round(c * sqrt((r^2 - dy^2) / r^2)), r = c = margin
    HaEllipse el;
    el.ksize = 2 * margin;
    const int r = margin, c = margin;
    const double inv_r2 = 1.0 / ((double)r * r);
    for (int i = 0; i < el.ksize; ++i) {
        const int dy = i - r;
        const int dx = (int)std::nearbyint(c * std::sqrt(((double)r * r - (double)dy * dy) * inv_r2));
        el.j1[i] = (signed char)std::max(c - dx, 0);
        el.j2[i] = (signed char)std::min(c + dx + 1, el.ksize);
    }
    dim3 egrid((W + kHaTile - 1) / kHaTile, (H + kHaTile - 1) / kHaTile, 2 * num);
    const size_t smem = (size_t)(kHaTile + el.ksize) * (kHaTile + el.ksize);
    ha_erode_kernel<<<egrid, kHaTile * kHaTile, smem, st>>>(raw, eroded, H, W, el);
    SPB_CHECK_LAUNCH();
}

void launch_ha_warp(const float* img, const float* coeffs_k, int B, int C, int H, int W, float* out, cudaStream_t st) {
    dim3 grid((W + 31) / 32, (H + 7) / 8, B);
    ha_warp_kernel<<<grid, 256, 0, st>>>(img, coeffs_k, C, H, W, out);
    SPB_CHECK_LAUNCH();
}

void launch_ha_aggregate(const float* probs, const uint8_t* maps, const float* coeffs, int num, int B, int H, int W, int use_max,
                         float* out, cudaStream_t st) {
    dim3 grid((W + 31) / 32, (H + 7) / 8, B);
    ha_aggregate_kernel<<<grid, 256, 0, st>>>(probs, maps, coeffs, num, B, H, W, use_max, out);
    SPB_CHECK_LAUNCH();
}

}  // namespace spb200
