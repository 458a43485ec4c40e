// Shared declarations of the spb200 engine (B200 / sm_100a SuperPoint inference).
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <string>

namespace spb200 {

enum Precision { PREC_FP32 = 0, PREC_FP16 = 1, PREC_BF16 = 2 };

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess)
        throw CudaError(std::string(what) + ": " + cudaGetErrorString(e) + " (" + file + ":" + std::to_string(line) + ")");
}
#define SPB_CUDA(x) ::spb200::cuda_check((x), #x, __FILE__, __LINE__)
#define SPB_CHECK_LAUNCH() ::spb200::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__)

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------
// The path is a chain of ~20 kernels on one stream.  Every kernel of the chain is launched with the programmatic-
// serialisation attribute: its CTAs may come up while the previous kernel is still draining (persistent kernels: as
// soon as a CTA of the previous kernel has retired from an SM), run their prologue - barrier initialisation, tensor-memory
// allocation, tensor-map prefetch, weight loads, all independent of the previous kernel's output - and then block in
// pdl_wait() until the previous kernel has completed and its writes are visible.  Rules: pdl_trigger() first thing in the
// kernel (every thread); every thread that reads or writes global memory other than constant weights calls pdl_wait()
// before it does.  SPB200_NO_PDL=1 launches without the attribute (both calls are then no-ops).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cuda_check(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...), "cudaLaunchKernelEx", __FILE__, __LINE__);
}

// the same for a kernel that runs as thread-block clusters of `cluster_x` CTAs along x
template <typename... KArgs, typename... Args>
inline void launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (unsigned)cluster_x; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 2;
    cuda_check(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...), "cudaLaunchKernelEx (cluster)", __FILE__, __LINE__);
}

constexpr int kMaxTaps = 9;
constexpr int kMaxSegs = 5;

// One K segment of an implicit-GEMM convolution: `ntaps` taps over a `cin`-channel NHWC source.
// Output pixel (oy, ox) reads source pixel (oy*stride + dy[t], ox*stride + dx[t]); pixels outside
// the source are zero (that is the convolution's zero padding).
struct SegDev {
    const void* src;      // NHWC activations, element type given by the kernel
    int H, W, C;          // source dims; C = stored channels per pixel
    int cin;              // channels consumed by this segment (<= C, multiple of 16)
    int cin_real;         // of which real (the rest is zero padding)
    int ntaps;
    int stride;
    int koff;             // first K index of this segment in the packed weights
    int view;             // > 1: src / H / W describe every `view`-th pixel of a full_H x full_W buffer in both axes (src points
    int full_H, full_W;   //      at the phase's first pixel): a stride-2 convolution read as stride-1 convolutions over four phases
    // Split-precision source (tensor-core path only): every 32 real channels are stored as 64 16-bit values
    // [hi 0..31 | lo 0..31] with value = hi + lo, so C and cin count STORED values (2 x the padded real channels) and a
    // 64-value K chunk carries a.hi and a.lo of 32 channels.  The packed weights hold w.hi at both halves of the chunk
    // from `koff` (a.hi w.hi + a.lo w.hi) and w.lo at the hi half from `koff_lo` (a.hi w.lo; < 0: no such range, the
    // weights are exact in 16 bits).
    int split;
    int koff_lo;
    // Packed K tail (tensor-core path, 3x3 segments whose last 64-channel chunk holds at most 16 real channels - the detector's
    // 65): > 0 = first K index of three extra 64-value slabs, one per filter row dy = -1, 0, 1, whose K = 16 step kk holds the
    // weights of tap (dy, dx = kk - 1) for the chunk's first 16 channels - the same values as in the main range.  A kernel may
    // run the tail as three slabs of three MMAs (A views one pixel apart) instead of nine slabs of one.  0: none.
    int koff_tail;
    int8_t dy[kMaxTaps], dx[kMaxTaps];
};

// out[p, co] = act( bias[co] + sum_seg sum_tap sum_ci src_seg[pix(p, tap), ci] * W[k, co] + residual[p, co] )
struct ConvDev {
    SegDev seg[kMaxSegs];
    int nseg;
    const void* w;        // SIMT: fp32 [K][cout_pad]; tensor-core path: 16-bit [cout_pad][K]
    const float* bias;    // [cout_pad], BatchNorm folded
    const void* residual; // NHWC, same pixel grid as the GEMM rows, res_C channels per pixel, or null
    void* dst;            // NHWC
    int B, OH, OW;        // GEMM rows = B*OH*OW output pixels
    int K, cout_pad;
    int dst_H, dst_W, dst_C;
    int dst_stride, dst_off_y, dst_off_x;   // dst pixel = (oy*dst_stride + off_y, ox*dst_stride + off_x)
    int res_C;
    int relu;
    int dst_fp32;         // tensor-core path: store fp32 instead of the 16-bit operand type
    int split_out;        // tensor-core path: store hi + lo in the split layout described at SegDev (dst_C counts stored values)
};

}  // namespace spb200
