// Kernel launchers of the spb200 engine (one translation unit per kernel family).
#pragma once
#include "common.cuh"

namespace spb200 {

// ---- conv_simt.cu --------------------------------------------------------------------------------
// img NCHW fp32 [B,C,H,W] (C = 1 or 3), w fp32 [C*49][64] with BatchNorm folded, dst NHWC [B,H/4,W/4,64].
void launch_stem_pool(const float* img, int B, int C, int H, int W, const float* w, const float* bias, void* dst,
                      int dst_type, cudaStream_t st);
void launch_conv_simt(const ConvDev& p, cudaStream_t st);
// 8-bit frame -> fp32 / 255 (the reference's loaders, python/src/inference.py:78-80)
void launch_u8_to_f32(const uint8_t* src, float* dst, long n, cudaStream_t st);
void launch_nhwc_to_nchw(const void* src, int src_type, int B, int HW, int Cs, int C, float* dst, cudaStream_t st);
// the same for a 16-bit source in the split layout (common.cuh, SegDev): channel c = hi + lo
void launch_split_to_nchw(const void* src, int src_type, int B, int HW, int Cs, int C, float* dst, cudaStream_t st);
// hi halves of a split-layout tensor [npix][2 C] as a plain 16-bit tensor [npix][C] (C a multiple of 32)
void launch_split_hi(const void* src, void* dst, long npix, int C, cudaStream_t st);

// ---- conv_tc.cu ----------------------------------------------------------------------------------
struct TcConvPlan;   // tensor maps + launch geometry of one tcgen05 convolution (conv_tc.cu)
TcConvPlan* tc_plan_create(const ConvDev& p, int operand_type);
void tc_plan_destroy(TcConvPlan* plan);
void launch_conv_tc(const TcConvPlan* plan, cudaStream_t st);

// ---- block_tc.cu ---------------------------------------------------------------------------------
struct TcBlockPlan;  // fused residual block (3x3 conv -> 1x1 conv + shortcut) or a single convolution, persistent
TcBlockPlan* tc_block_plan_create(const ConvDev& c1, const ConvDev* c2, int operand_type, int real_cout, int num_sms);
void tc_block_plan_destroy(TcBlockPlan* plan);
void launch_block_tc(const TcBlockPlan* plan, cudaStream_t st);

// ---- halo_tc.cu ----------------------------------------------------------------------------------
struct TcHaloPlan;   // stride-1 blocks with 64/128 output channels: haloed activation tiles, shared weight slabs
// nullptr when the block does not fit (the caller falls back to tc_block_plan_create)
TcHaloPlan* tc_halo_plan_create(const ConvDev& c1, const ConvDev* c2, int operand_type, int real_cout, int num_sms);
void tc_halo_plan_destroy(TcHaloPlan* plan);
void launch_halo_tc(const TcHaloPlan* plan, cudaStream_t st);
// the detector's last block writing the softmax in depth-to-space order instead of the logits: heat_exp [B][8 OH][8 OW] = exp(l_c) of
// channels 0..63, heat_inv [B][OH][OW] = 1 / (sum over the 65 channels + 1e-5); heatmap = heat_exp * heat_inv of the pixel's cell
bool tc_halo_heat_capable(const TcHaloPlan* plan);
void launch_halo_tc_heat(const TcHaloPlan* plan, float* heat_exp, float* heat_inv, int B, cudaStream_t st);

// ---- stem_tc.cu ----------------------------------------------------------------------------------
struct StemTcPlan;
StemTcPlan* stem_tc_plan_create(const void* w16, const float* bias, int cin, int operand_type, int num_sms);
void stem_tc_plan_destroy(StemTcPlan* plan);
void launch_stem_tc(const StemTcPlan* plan, const float* img, void* dst, int B, int H, int W, cudaStream_t st);
// split-precision variant: image hi + lo, three MMAs per product, fp32 pooling; dst [B][H/4][W/4][128] in the split layout
void launch_stem_wide(const StemTcPlan* plan, const float* img, void* dst, int B, int H, int W, cudaStream_t st);

// ---- stem_planes.cu ------------------------------------------------------------------------------
// Gray stem fed by TMA from 16-bit row-parity planes of the image (no im2col pass).
struct StemPlanesPlan;
StemPlanesPlan* stem_planes_plan_create(const void* w16, const float* bias, int operand_type, int num_sms);
void stem_planes_plan_destroy(StemPlanesPlan* plan);
// img: [B][H][W] fp32 in [0,1] (img_is_u8 = 0) or 8-bit (1) -> planes [B][2][H/2][W] 16-bit, value x255
void launch_planes(const void* img, int img_is_u8, void* planes, int operand_type, int B, int H, int W, cudaStream_t st);
// src: the planes (src_is_f32 = 0) or the fp32 image itself (1, 16-byte aligned): the kernel then converts on the way, no plane pass
void launch_stem_planes(StemPlanesPlan* plan, const void* src, int src_is_f32, void* dst, int B, int H, int W, cudaStream_t st);

// ---- preproc.cu ----------------------------------------------------------------------------------
// The frame loaders of the two demos on the device (reference cpp/src/camera.cc:12-23, python/src/inference.py:72-85).
// Tables: host-built with OpenCV's arithmetic (build_*), uploaded by the caller, read by the kernels.
size_t preprocess_u8_table_ints(int H, int W);                                        // 4 (H + W)
void build_preprocess_u8_table(int h, int w, int H, int W, int* tab);
void launch_preprocess_u8(const uint8_t* src, int B, int h, int w, int C, const int* tab_dev, uint8_t* dst, int H, int W, cudaStream_t st);
void build_preprocess_f32_table(int h, int w, int H, int W, int* itab /* 2 (W + H) */, float* ftab /* W + H */);
void launch_preprocess_f32(const float* src, int B, int h, int w, const int* itab_dev, const float* ftab_dev, float* dst, int H, int W,
                           cudaStream_t st);

// ---- postproc.cu ---------------------------------------------------------------------------------
// Softmax-with-epsilon over 65 channels, drop the dustbin, depth-to-space (reference
// python/src/superpoint.py:111-114, python/src/netutils.py:64-75).  logits element (b, c, i, j) is at
// logits[b*batch_stride + c*chan_stride + (i*Wc + j)*cell_stride].
// heat[b][y][x] *= inv[b][y / 8][x / 8] in place: the full heatmap from the detector tail's exp / normaliser pair
void launch_heat_scale(float* heat, const float* inv, int B, int H, int W, cudaStream_t st);
void launch_heatmap(const float* logits, long batch_stride, long chan_stride, long cell_stride, int B, int Hc, int Wc,
                    float* heat, cudaStream_t st);

// restore_prob_map (python/src/netutils.py:64-75) on an already softmaxed B*65*Hc*Wc tensor: dustbin drop + depth-to-space
void launch_depth_to_space(const float* softmax_nchw, int B, int Hc, int Wc, float* heat, cudaStream_t st);

// n ints from device memory into PINNED host memory by a kernel (no copy engine involved)
void launch_counts_to_host(const int* src, int* dst_pinned, int n, cudaStream_t st);

struct NmsWorkspace {
    unsigned long long* keys;      // [B][kcap] survivors as sortable keys (conf bits << 32 | ~pixel index)
    unsigned long long* keys_alt;  // [B][kcap] ping-pong buffer of the radix sort
    int* counters;                 // [B][8]: survivors, undecided after round 0, round totals (sortkey.cuh)
    int kcap;                      // capacity of keys per image: ceil(H/(r+1))*ceil(W/(r+1)), the survivor bound
    unsigned* mask;                // [B][H][mask_w] bit per pixel: undecided candidate
    int mask_w;                    // 32-bit words per image row
    unsigned* und;                 // [B][H*W] pixel indices of the candidates still undecided after round 0
    unsigned* ukey;                // [B][H*W] sortable key of a pixel, valid where its undecided bit is (or was) set
};
// Greedy-equivalent grid NMS + border removal (reference python/src/nms.py:4-53, python/src/netutils.py:59,95-99),
// descending sort (netutils.py:92-93) and top-k truncation, as two launches.
// launch_nms_round0: input `heat` [B][H][W], or - when heat is null and nms_logits_supported(radius) - the detector
// logits (channels last, cell_stride floats per cell), from which it computes the softmax / depth-to-space values
// itself; leaves the first keepers, the undecided candidates and their bit mask in the workspace.
// launch_nms_finish: remaining rounds, sort, outputs per image: count (clamped to cap and top_k), xy int32 [cap][2]
// as (x, y), conf fp32 [cap].
bool nms_logits_supported(int radius);
// zero_counters: clear the per-image counters first (a memset node); the finish kernel leaves them at zero, so only the
// first call on a fresh workspace - or the call after a failed one - needs it
// heat_inv (with heat only, may be null): [B][H/8][W/8], a pixel's value is heat * heat_inv of its cell (the detector tail's fused
// softmax, launch_halo_tc_heat)
void launch_nms_round0(const float* heat, const float* logits, int cell_stride, int B, int H, int W, float thresh, int radius,
                       int border, const NmsWorkspace& ws, bool zero_counters, cudaStream_t st, const float* heat_inv = nullptr);
void launch_nms_finish(int B, int H, int W, int radius, int border, int top_k, int cap, const NmsWorkspace& ws, int* count,
                       int* xy, float* conf, cudaStream_t st);

// Bilinear sampling (align_corners=True) of the descriptor map at the keypoints + L2 normalisation
// (reference python/src/netutils.py:103-121).  map element (b, c, i, j) is at
// map[b*batch_stride + c*chan_stride + (i*Wc + j)*cell_stride]; map_type is a Precision value; gtab[x] / gtab[W+y]
// are the sampling positions ix / iy of pixel column x / row y (host-built, device memory).
void launch_sample_descriptors(const void* map, int map_type, long batch_stride, long chan_stride, long cell_stride,
                               int B, int D, int Hc, int Wc, int W, const float* gtab, int cap, const int* count,
                               const int* xy, void* out, int out_fp16, cudaStream_t st);

// ---- homography.cu -------------------------------------------------------------------------------
// Pieces of homography adaptation (reference python/src/homographies.py:250-324).  coeffs: device [2 num][8], the num
// forward transforms followed by their inverses.  maps: [2 num][H][W] bytes, map 2k = count_k, 2k + 1 = mask_k.
void launch_ha_valid_maps(const float* coeffs, int num, int H, int W, int margin, uint8_t* raw, uint8_t* eroded, cudaStream_t st);
void launch_ha_warp(const float* img, const float* coeffs_k, int B, int C, int H, int W, float* out, cudaStream_t st);
void launch_ha_aggregate(const float* probs, const uint8_t* maps, const float* coeffs, int num, int B, int H, int W, int use_max,
                         float* out, cudaStream_t st);

// ---- match.cu ------------------------------------------------------------------------------------
// Mutual nearest neighbours in L2 distance between the descriptor sets of image pairs (reference
// python/src/inference.py:88-96, BFMatcher crossCheck; gate: settings.py:6 nn_thresh, 0 = none).
// desc_a / desc_b: [B][cap][D] fp32 with count_a[b] / count_b[b] valid rows; best_a / best_b: [B][cap] scratch;
// match[b][i] = index in b of the match of descriptor i of a, or -1; dist[b][i] = distance to its nearest.
void launch_match(const float* desc_a, const int* count_a, const float* desc_b, const int* count_b, int B, int cap, int D,
                  float max_dist, unsigned long long* best_a, unsigned long long* best_b, int* match, float* dist, int num_sms,
                  cudaStream_t st);

// ---- match_tc.cu ---------------------------------------------------------------------------------
// The same on the tensor cores for D = 128: split-precision fp16 GEMM (hi.hi + hi.lo + lo.hi, fp32 accumulation) with the
// row minima taken in the epilogue, run in both directions.  workspace: match_tc_workspace_bytes(B, cap) bytes.
size_t match_tc_workspace_bytes(int B, int cap);
void launch_match_tc(const float* desc_a, const int* count_a, const float* desc_b, const int* count_b, int B, int cap,
                     float max_dist, void* workspace, unsigned long long* best_a, unsigned long long* best_b, int* match,
                     float* dist, int num_sms, cudaStream_t st);

}  // namespace spb200
