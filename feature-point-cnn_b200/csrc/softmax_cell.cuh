// The reference's detector softmax for one cell (python/src/superpoint.py:111-112: exp(l) / (sum_c exp(l_c) + 1e-5), no
// maximum subtraction), evaluated by EIGHT adjacent lanes: lane j of the group owns channels 8j .. 8j+7 - after the
// depth-to-space of python/src/netutils.py:64-75 that is pixel row j of the cell, eight consecutive pixels - and lane 0
// adds the dustbin.  heatmap_kernel and nms_round0_kernel both call this, so the two give bit-identical values.
#pragma once

namespace spb200 {

// l[0..7]: the lane's logits, l64: the dustbin logit (used by j == 0 only).  Returns the eight heatmap values in h.
// Every lane of the warp must call it (full-mask shuffles); groups are aligned octets of lanes.
__device__ __forceinline__ void softmax_cell_octet(const float (&l)[8], float l64, int j, float (&h)[8]) {
    float e[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) e[k] = expf(l[k]);
    float s = e[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) s += e[k];
    if (j == 0) s += expf(l64);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    // one division per cell: e * (1 / den) is within an ulp of e / den, and every kernel that produces heatmap values
    // (this helper and the detector's fused epilogue, halo_tc.cu) forms them this way, so they agree bit for bit
    const float inv = 1.f / (s + 0.00001f);
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = e[k] * inv;
}

}  // namespace spb200
