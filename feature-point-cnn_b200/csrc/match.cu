// Descriptor matching, the step right after the hot path in both demos of the reference: mutual nearest
// neighbours in L2 distance (python/src/inference.py:88-96, cv2.BFMatcher(NORM_L2, crossCheck=True)), optionally
// gated by a maximum distance (settings.py:6 nn_thresh; cpp/src/main.cc:9-29 accepts a correspondence below a
// tolerance).  Works on the batched output of spb200_detect: desc [B][cap][D] with count[B] valid rows per image.
//
// match_tile_kernel: fp32 CUDA cores.  A CTA owns a strip of 64 query descriptors (resident in shared memory, k-major)
// and walks its share of the train descriptors in tiles of 64; a thread accumulates a 4x4 block of dot products
// (two 16-byte shared loads feed sixteen FMAs), turns them into squared distances |a|^2 + |b|^2 - 2 a.b, and the
// minima of every row and column of the tile go into 64-bit keys (distance bits << 32 | index; distances are >= 0,
// so the order of the keys is the order of (distance, index) and the lowest index wins a tie as in BFMatcher) -
// shared atomics per tile, one global atomicMin per row / column and tile.  One pass gives both directions.
// match_finish_kernel: a query is matched to its nearest train descriptor when that one's nearest query is the
// query itself and the distance is below the gate.
#include "kernels.h"

namespace spb200 {

constexpr int kMtT = 64;            // tile edge (descriptors)
constexpr int kMtPitch = 68;        // floats per k-row of a tile in shared memory (16-byte aligned, few store conflicts)
constexpr int kMtThreads = 256;

__global__ void match_init_kernel(unsigned long long* __restrict__ best_a, unsigned long long* __restrict__ best_b, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { best_a[i] = ~0ull; best_b[i] = ~0ull; }
}

template <int D>
__global__ void __launch_bounds__(kMtThreads)
match_tile_kernel(const float* __restrict__ desc_a, const int* __restrict__ count_a, const float* __restrict__ desc_b,
                  const int* __restrict__ count_b, int cap, unsigned long long* __restrict__ best_a,
                  unsigned long long* __restrict__ best_b) {
    extern __shared__ __align__(16) float mt_smem[];
    float* s_a = mt_smem;                                    // [D][kMtPitch]
    float* s_b = s_a + D * kMtPitch;                         // [D][kMtPitch]
    __shared__ float s_na[kMtT], s_nb[kMtT];
    __shared__ unsigned long long s_ra[kMtT], s_cb[kMtT];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int b = blockIdx.z;
    const int na = min(__ldg(count_a + b), cap), nb = min(__ldg(count_b + b), cap);
    const int q0 = blockIdx.x * kMtT;
    if (q0 >= na) return;
    const float* A = desc_a + (size_t)b * cap * D;
    const float* Bm = desc_b + (size_t)b * cap * D;
    unsigned long long* ba = best_a + (size_t)b * cap;
    unsigned long long* bb = best_b + (size_t)b * cap;

    // the query strip, transposed to k-major; rows beyond the count are zero (and never reported)
    auto load_tile = [&](const float* src, int row0, int nrows, float* dst, float* norms) {
        constexpr int kQ = D / 4;                            // float4 per descriptor
        for (int i = tid; i < kMtT * kQ; i += kMtThreads) {
            const int r = i / kQ, kq = i % kQ;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < nrows) v = __ldg(reinterpret_cast<const float4*>(src + (size_t)(row0 + r) * D) + kq);
            dst[(kq * 4 + 0) * kMtPitch + r] = v.x;
            dst[(kq * 4 + 1) * kMtPitch + r] = v.y;
            dst[(kq * 4 + 2) * kMtPitch + r] = v.z;
            dst[(kq * 4 + 3) * kMtPitch + r] = v.w;
        }
        __syncthreads();
        if (tid < kMtT) {
            float s = 0.f;
            for (int k = 0; k < D; ++k) s = fmaf(dst[k * kMtPitch + tid], dst[k * kMtPitch + tid], s);
            norms[tid] = s;
        }
    };
    load_tile(A, q0, na, s_a, s_na);
    if (tid < kMtT) s_ra[tid] = ~0ull;
    __syncthreads();

    const int ntiles = (nb + kMtT - 1) / kMtT;
    for (int t = blockIdx.y; t < ntiles; t += gridDim.y) {
        const int t0 = t * kMtT;
        __syncthreads();                                     // the previous tile's s_b / s_cb readers are done
        load_tile(Bm, t0, nb, s_b, s_nb);
        if (tid < kMtT) s_cb[tid] = ~0ull;
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
        for (int k = 0; k < D; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(s_a + k * kMtPitch + ty * 4);
            const float4 c = *reinterpret_cast<const float4*>(s_b + k * kMtPitch + tx * 4);
            const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], cv[j], acc[i][j]);
        }
        unsigned long long rmin[4] = {~0ull, ~0ull, ~0ull, ~0ull}, cmin[4] = {~0ull, ~0ull, ~0ull, ~0ull};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = q0 + ty * 4 + i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int tr = t0 + tx * 4 + j;
                if (q < na && tr < nb) {
                    const float d2 = fmaxf(s_na[ty * 4 + i] + s_nb[tx * 4 + j] - 2.f * acc[i][j], 0.f);
                    const unsigned long long kq = ((unsigned long long)__float_as_uint(d2) << 32);
                    rmin[i] = min(rmin[i], kq | (unsigned)tr);
                    cmin[j] = min(cmin[j], kq | (unsigned)q);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            // the 16 threads of a row are the 16 lanes of a half-warp: reduce there, then one shared atomic
            unsigned long long v = rmin[i];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (tx == 0 && v != ~0ull) atomicMin(&s_ra[ty * 4 + i], v);
            if (cmin[i] != ~0ull) atomicMin(&s_cb[tx * 4 + i], cmin[i]);
        }
        __syncthreads();
        if (tid < kMtT && t0 + tid < nb && s_cb[tid] != ~0ull) atomicMin(bb + t0 + tid, s_cb[tid]);
    }
    __syncthreads();
    if (tid < kMtT && q0 + tid < na && s_ra[tid] != ~0ull) atomicMin(ba + q0 + tid, s_ra[tid]);
}

__global__ void match_finish_kernel(const int* __restrict__ count_a, int cap, const unsigned long long* __restrict__ best_a,
                                    const unsigned long long* __restrict__ best_b, float max_dist, int* __restrict__ match,
                                    float* __restrict__ dist) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= cap) return;
    int m = -1;
    float d = 0.f;
    if (q < min(__ldg(count_a + b), cap)) {
        const unsigned long long ka = best_a[(size_t)b * cap + q];
        if (ka != ~0ull) {
            const int j = (int)(unsigned)(ka & 0xffffffffull);
            d = sqrtf(__uint_as_float((unsigned)(ka >> 32)));
            const unsigned long long kb = best_b[(size_t)b * cap + j];
            if ((unsigned)(kb & 0xffffffffull) == (unsigned)q && (max_dist <= 0.f || d < max_dist)) m = j;
        }
    }
    match[(size_t)b * cap + q] = m;
    dist[(size_t)b * cap + q] = d;
}

void match_init_launch(unsigned long long* best_a, unsigned long long* best_b, long n, cudaStream_t st) {
    match_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(best_a, best_b, n);
    SPB_CHECK_LAUNCH();
}

void match_finish_launch(const int* count_a, int B, int cap, const unsigned long long* best_a, const unsigned long long* best_b,
                         float max_dist, int* match, float* dist, cudaStream_t st) {
    match_finish_kernel<<<dim3((cap + 255) / 256, B), 256, 0, st>>>(count_a, cap, best_a, best_b, max_dist, match, dist);
    SPB_CHECK_LAUNCH();
}

void launch_match(const float* desc_a, const int* count_a, const float* desc_b, const int* count_b, int B, int cap, int D,
                  float max_dist, unsigned long long* best_a, unsigned long long* best_b, int* match, float* dist, int num_sms,
                  cudaStream_t st) {
    if (D != 128 && D != 256) throw std::invalid_argument("match: descriptor dimension must be 128 or 256");
    if (B <= 0 || cap <= 0) throw std::invalid_argument("match: batch and capacity must be positive");
    const long n = (long)B * cap;
    match_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(best_a, best_b, n);
    SPB_CHECK_LAUNCH();
    const int strips = (cap + kMtT - 1) / kMtT;
    // enough train-range splits to give every SM a few CTAs even for one image pair
    int splits = std::max(1, std::min(16, (4 * num_sms + strips * B - 1) / (strips * B)));
    dim3 grid(strips, splits, B);
    const size_t smem = sizeof(float) * 2 * (size_t)D * kMtPitch;
    if (D == 128) {
        SPB_CUDA(cudaFuncSetAttribute(match_tile_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        match_tile_kernel<128><<<grid, kMtThreads, smem, st>>>(desc_a, count_a, desc_b, count_b, cap, best_a, best_b);
    } else {
        SPB_CUDA(cudaFuncSetAttribute(match_tile_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        match_tile_kernel<256><<<grid, kMtThreads, smem, st>>>(desc_a, count_a, desc_b, count_b, cap, best_a, best_b);
    }
    SPB_CHECK_LAUNCH();
    match_finish_kernel<<<dim3((cap + 255) / 256, B), 256, 0, st>>>(count_a, cap, best_a, best_b, max_dist, match, dist);
    SPB_CHECK_LAUNCH();
}

}  // namespace spb200
