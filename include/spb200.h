/* spb200 - B200-native (sm_100a) SuperPoint / MagicPoint inference: the C ABI.
 *
 * This is the drop-in boundary for the inference hot path of Kolkir/feature-point-cnn.  The reference
 * has no FFI of its own; its boundary is two concrete classes, and every entry point below names the
 * reference interface it replaces (paths relative to the reference repository):
 *
 *   C++    superpoint::SuperPoint::SuperPoint / ProcessFrame      cpp/src/superpoint.h:12-36,
 *                                                                 cpp/src/superpoint.cc:9-96
 *   Python SuperPoint.forward                                     python/src/superpoint.py:91-115
 *          InferenceWrapper.__init__ / run                        python/src/inferencewrapper.py:13-46
 *          load_checkpoint_for_inference                          python/src/saveutils.py:6-18
 *          restore_prob_map / get_points / get_descriptors        python/src/netutils.py:64-121
 *          corners_nms                                            python/src/nms.py:4-53
 *
 * Conventions: plain C types only; every function returns 0 on success and a non-zero SPB200_E_* code on
 * failure, after which spb200_last_error() holds a message; no exception crosses the ABI.  Unless a name
 * ends in _host, every data pointer is DEVICE memory on the engine's GPU, owned by the caller, and the
 * work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = the default stream) without a
 * host synchronisation.  An engine is bound to one GPU, owns its weights and workspace, and is not
 * thread-safe; engines are independent of each other (one per GPU / per thread).  Calls on ONE engine
 * share its workspace and are therefore ordered by the engine even across streams: a call enqueued on
 * a stream other than the previous call's waits (on the device, cudaStreamWaitEvent) until the previous
 * call's work has finished; the *_host calls run on streams of their own and obey the same rule.  The
 * caller's own buffers are the caller's to order.  There is no CPU fallback: creating an engine without
 * a Blackwell GPU fails.
 */
#ifndef SPB200_H
#define SPB200_H

#include <stdint.h>

#if defined(__GNUC__)
#define SPB200_API __attribute__((visibility("default")))
#else
#define SPB200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct spb200_engine spb200_engine;

enum {
    SPB200_OK = 0,
    SPB200_E_INVALID = 1,   /* bad argument */
    SPB200_E_RUNTIME = 2,   /* CUDA error, missing key, unreadable file, ... */
    SPB200_E_NOMEM = 3
};

/* arithmetic of the convolutions */
enum {
    SPB200_PREC_FP32 = 0,   /* fp32 CUDA cores (parity mode, also checks the tensor-core path)      */
    SPB200_PREC_FP16 = 1,   /* tcgen05 kind::f16, fp16 operands, fp32 accumulate (default)          */
    SPB200_PREC_BF16 = 2    /* tcgen05 kind::f16, bf16 operands, fp32 accumulate                    */
};

/* Engine on CUDA device `device`.  Replaces the constructors superpoint.cc:9-66 / inferencewrapper.py:13-27
 * (minus the weight load, see below). */
SPB200_API int spb200_create(int device, spb200_engine** out);
SPB200_API void spb200_destroy(spb200_engine* e);

/* Message of the last failure on `e` (or of the last failed spb200_create when e is NULL). */
SPB200_API const char* spb200_last_error(const spb200_engine* e);

/* Load snapshots/super_point.pt or magic_point.pt unchanged: a torch.save archive holding either the
 * save_checkpoint dict (saveutils.py:54-63) or a bare state_dict (inferencewrapper.py:89-91).
 * Replaces load_checkpoint_for_inference (saveutils.py:6-18) and the pickle_load branch superpoint.cc:27-53. */
SPB200_API int spb200_load_checkpoint(spb200_engine* e, const char* path);

/* Alternative to load_checkpoint for hosts that already hold the tensors (the PyTorch shim's
 * load_state_dict): one call per state_dict entry, fp32 host data, contiguous. */
SPB200_API int spb200_load_tensor(spb200_engine* e, const char* key, const float* host_data, const int64_t* shape, int rank);

/* Fold BatchNorm into the convolutions, pack and upload the weights for `precision` (SPB200_PREC_*).
 * Must follow the loads and precede any inference call; fails if a state_dict key is missing
 * (strict loading, saveutils.py:9-14). */
SPB200_API int spb200_finalize_weights(spb200_engine* e, int precision);

/* Split-precision stages: the guaranteed-parity mode of the tensor-core path (no counterpart in the reference, whose
 * arithmetic is fp32 throughout: superpoint.py:91-115).  In a split stage activations and weights are kept as 16-bit
 * hi + lo pairs and every product is three MMAs (a.hi w.hi + a.hi w.lo + a.lo w.hi, fp32 accumulation): results agree
 * with the fp32 network to ~1e-4 on the logits instead of ~5e-2, at three times the tensor work of the stage.  Errors
 * of early layers are amplified by the later ones, so only prefixes of the network are accepted. */
enum {
    SPB200_SPLIT_NONE = 0,      /* what spb200_finalize_weights does                                     */
    SPB200_SPLIT_LAYER1 = 1,    /* stem (image as hi + lo) + encoder.layer1                              */
    SPB200_SPLIT_LAYER2 = 2,    /* + encoder.layer2: the whole encoder                                   */
    SPB200_SPLIT_DETECTOR = 3   /* + detector head: every layer the heatmap depends on                   */
};
/* Same as spb200_finalize_weights with the stages up to `split_level` in split precision; precision must be
 * SPB200_PREC_FP16 or SPB200_PREC_BF16 when split_level != 0.  The descriptor head is never split. */
SPB200_API int spb200_finalize_weights_split(spb200_engine* e, int precision, int split_level);

/* settings.py:3-8 / settings.h:27-31.  top_k = 0 returns every survivor (the reference's behaviour);
 * descriptor_enabled = 0 is SuperPoint.disable_descriptor (superpoint.py:74-78, MagicPoint). */
SPB200_API int spb200_set_params(spb200_engine* e, float conf_thresh, int nms_dist, int border_remove, int top_k,
                      int descriptor_enabled);

/* SuperPoint.forward (superpoint.py:91-115).  img: B*C*H*W fp32 in [0,1], C = 1 (grayscale) or 3 (a gray
 * image replicated to RGB gives the same result, dataset_utils.py:18-20); H and W multiples of 16.
 * prob_map: B*H*W fp32 (required); desc_map: B*128*(H/8)*(W/8) fp32 NCHW, un-normalised, may be NULL;
 * logits: B*65*(H/8)*(W/8) fp32 NCHW, may be NULL. */
SPB200_API int spb200_forward(spb200_engine* e, const float* img, int B, int C, int H, int W, float* prob_map, float* desc_map,
                   float* logits, void* stream);

/* InferenceWrapper.run (inferencewrapper.py:29-46) / ProcessFrame (superpoint.cc:68-96) for every image of
 * a batch: network -> threshold -> NMS -> sort by descending confidence -> border removal -> top-k ->
 * descriptors.  Per image b: count[b] keypoints (at most `capacity`), xy[b][i] = (x, y) int32,
 * conf[b][i], desc[b][i][0..127] unit-norm fp32 (desc may be NULL; zeros when the descriptor head is
 * disabled); prob_map (B*H*W) may be NULL. */
SPB200_API int spb200_detect(spb200_engine* e, const float* img, int B, int C, int H, int W, int capacity, int* count, int* xy,
                  float* conf, float* desc, float* prob_map, void* stream);

/* Same, with HOST buffers and the copies inside the call (pinned staging, synchronous): the path a
 * caller with a cv::Mat / numpy frame takes (torchutis.cc:5-10,19-23; netutils.py:57,119). */
SPB200_API int spb200_detect_host(spb200_engine* e, const float* img_host, int B, int C, int H, int W, int capacity,
                       int* count_host, int* xy_host, float* conf_host, float* desc_host);

/* The same two calls for 8-bit grayscale frames, B*H*W bytes, value k meaning k/255: the frame as the camera or
 * cv2.imread delivers it, before the division the reference's loaders do on the host (python/src/inference.py:72-85
 * resize -> gray -> /255; cpp/src/camera.cc:12-23 convertTo(CV_32FC1, 1/255)).  A quarter of the upload bytes; with
 * the tensor-core path the result is bit-identical to passing k/255.f as fp32. */
SPB200_API int spb200_detect_u8(spb200_engine* e, const uint8_t* img, int B, int H, int W, int capacity, int* count, int* xy,
                     float* conf, float* desc, float* prob_map, void* stream);
SPB200_API int spb200_detect_host_u8(spb200_engine* e, const uint8_t* img_host, int B, int H, int W, int capacity,
                          int* count_host, int* xy_host, float* conf_host, float* desc_host);

/* The two halves of spb200_detect_host / _u8, for callers that stream batches.  submit copies (pageable) or registers
 * (pinned) the frames, enqueues upload + network + post-processing and returns at once with a ticket; the download of the
 * keypoints and descriptors into the caller's arrays is driven by a thread of the engine as the results become ready;
 * wait returns when batch `ticket` is complete in those arrays.  At most TWO batches may be in flight: submit(i + 1)
 * before wait(i) lets batch i's download - the longest stage over PCIe - run under batch i + 1's compute.  img_host (when
 * pinned) and the four output arrays must stay valid until the matching wait returns.  img_is_u8: B*H*W bytes (C = 1)
 * instead of B*C*H*W fp32.  desc_host = NULL skips the descriptors. */
SPB200_API int spb200_detect_host_submit(spb200_engine* e, const void* img_host, int img_is_u8, int B, int C, int H, int W, int capacity,
                              int* count_host, int* xy_host, float* conf_host, void* desc_host, int* ticket);
SPB200_API int spb200_detect_host_wait(spb200_engine* e, int ticket);

/* Element type of the `desc` arrays that spb200_detect*, spb200_detect_host* fill: SPB200_DESC_FP32 (default: the reference
 * returns float32, netutils.py:119-121) or SPB200_DESC_FP16 - the same unit vectors rounded to half precision, B*capacity*128
 * halves in the same layout: half the bytes over PCIe, the dominant cost of the host-buffer calls.  Pass the buffer through the
 * float* parameter.  spb200_sample_descriptors always writes fp32. */
enum { SPB200_DESC_FP32 = 0, SPB200_DESC_FP16 = 1 };
SPB200_API int spb200_set_descriptor_format(spb200_engine* e, int format);

/* The frame loaders in front of the path, on the device (all pointers device memory).
 * spb200_preprocess_u8: the C++ demo's loader (cpp/src/camera.cc:12-23): cv::resize(frame, Size(W, H)) - INTER_LINEAR on
 *   8-bit pixels - then cvtColor(BGR2GRAY); the convertTo(CV_32FC1, 1/255) that follows is what spb200_detect_u8 applies.
 *   frames: B*h*w*C bytes, C = 3 (BGR, interleaved) or 1 (gray); gray: B*H*W bytes, bit-exact with OpenCV 4.
 * spb200_preprocess_f32: the Python demo's loader (python/src/inference.py:72-85 make_query_image + the HWC -> CHW of
 *   inferencewrapper.py:70-81): frames B*h*w*3 fp32 BGR in [0,1] -> BGR2RGB, INTER_LINEAR resize by max(H/h, W/w) (ratio
 *   preserving), centre crop -> rgb: B*3*H*W fp32, the input of spb200_forward / spb200_detect with C = 3. */
SPB200_API int spb200_preprocess_u8(spb200_engine* e, const uint8_t* frames, int B, int h, int w, int C, uint8_t* gray, int H, int W,
                         void* stream);
SPB200_API int spb200_preprocess_f32(spb200_engine* e, const float* frames, int B, int h, int w, float* rgb, int H, int W, void* stream);

/* Stage-level entry points with the reference's tensor layouts (NCHW fp32). */
/* exp(l)/(sum exp(l)+1e-5), drop dustbin, depth-to-space: superpoint.py:111-114 + restore_prob_map
 * (netutils.py:64-75) / GetPoints (superpoint.cc:154-173).  logits B*65*(H/8)*(W/8) -> prob_map B*H*W. */
SPB200_API int spb200_heatmap_from_logits(spb200_engine* e, const float* logits, int B, int H, int W, float* prob_map, void* stream);
/* restore_prob_map (netutils.py:64-75) with the reference's contract: `softmax` B*65*(H/8)*(W/8) is ALREADY softmaxed
 * (what make_prob_map_from_labels and SuperPoint.forward hand it); the dustbin is dropped and the 64 channels of a cell
 * become its 8x8 pixels.  prob_map B*H*W. */
SPB200_API int spb200_restore_prob_map(spb200_engine* e, const float* softmax, int B, int H, int W, float* prob_map, void* stream);
/* get_points (netutils.py:78-100) incl. corners_nms (nms.py:4-53) / FeatureNMS (torchutis.cc:37-99). */
SPB200_API int spb200_nms(spb200_engine* e, const float* prob_map, int B, int H, int W, int capacity, int* count, int* xy,
               float* conf, void* stream);
/* get_descriptors (netutils.py:103-121) / AddDescriptors (superpoint.cc:98-152): desc_map B*D*(H/8)*(W/8). */
SPB200_API int spb200_sample_descriptors(spb200_engine* e, const float* desc_map, int B, int D, int H, int W, int capacity,
                              const int* count, const int* xy, float* desc, void* stream);

/* homography_adaptation (homographies.py:250-324) behind InferenceWrapper.run_with_homography_adaptation
 * (inferencewrapper.py:48-68) and the COCO pseudo-labelling job (preprocess_coco.py:64-74): the detector runs on the
 * images and on `num` warps of them, the warped heatmaps are projected back, masked by the eroded validity maps and
 * averaged (aggregation 0, the reference's 'sum') or maximised (1); pixels seen by fewer than num/3 views are zero.
 * img: B*C*H*W fp32 on the device; homographies_host: num*8 fp32 on the HOST, the flattened transforms as
 * sample_homography (homographies.py:79-196) returns them; prob_map: B*H*W fp32 on the device.  The call synchronises
 * `stream` once (coefficient upload). */
SPB200_API int spb200_homography_adaptation(spb200_engine* e, const float* img, int B, int C, int H, int W,
                                 const float* homographies_host, int num, int valid_border_margin, int aggregation,
                                 float* prob_map, void* stream);

/* Descriptor matching, the step that follows the path in both demos: get_best_correspondences (inference.py:88-96,
 * cv2.BFMatcher(NORM_L2, crossCheck=True)) / SearchKeyFrameCorrespondence (main.cc:9-29) for B image pairs at once,
 * on the arrays spb200_detect produced.  desc_a, desc_b: B*capacity*D fp32 with count_a[b] / count_b[b] valid rows.
 * match_ab[b][i] = index in b of the mutual nearest neighbour of descriptor i of a, or -1 (no mutual neighbour, or
 * its L2 distance is not below max_dist; max_dist <= 0 disables the gate - settings.py:6 nn_thresh = 0.7);
 * dist[b][i] = L2 distance of descriptor i to its nearest descriptor of b.  Ties go to the lowest index. */
SPB200_API int spb200_match(spb200_engine* e, const float* desc_a, const int* count_a, const float* desc_b, const int* count_b,
                 int B, int capacity, int D, float max_dist, int* match_ab, float* dist, void* stream);

/* 128 for these checkpoints (superpoint.py:49-50; the C++ demo's 256, torchutis.h:11, is stale). */
SPB200_API int spb200_descriptor_dim(const spb200_engine* e);
/* Upper bound of survivors of one image: ceil(H/(nms_dist+1)) * ceil(W/(nms_dist+1)). */
SPB200_API int spb200_max_keypoints(int H, int W, int nms_dist);
/* Kernels launched by this engine since creation (or since the last reset). */
SPB200_API long spb200_kernel_launches(const spb200_engine* e);
SPB200_API void spb200_reset_kernel_launches(spb200_engine* e);

/* Debug / parity: copy internal activation `buffer_id` of the last network run to dst (device, NCHW fp32,
 * first `channels` channels); buffer ids are listed in csrc/engine.h (BufId). */
SPB200_API int spb200_export_activation(spb200_engine* e, int buffer_id, float* dst, int channels, void* stream);
SPB200_API int spb200_activation_dims(const spb200_engine* e, int buffer_id, int* channels, int* height, int* width);

/* Per-launch timing with CUDA events on the launching stream: begin, make inference calls, end.
 * end synchronises the device and returns, per kernel launch since begin: a name (64 bytes each), its
 * duration in ms, its algorithmic FLOPs (convolutions: 2 x MACs of the reference layer) and its
 * algorithmic bytes where they do not depend on the keypoint count (else 0). */
SPB200_API int spb200_profile_begin(spb200_engine* e);
SPB200_API int spb200_profile_end(spb200_engine* e, int max_entries, char* names, float* ms, double* flops, double* bytes, int* n);

/* Checkpoint inspection on the host (no GPU needed): number of tensors in the file's model_state_dict
 * (negative on failure), and one tensor converted to fp32.  `dst` may be NULL to query shape/rank only;
 * returns SPB200_E_INVALID if `capacity` (elements) is too small or the key is absent. */
SPB200_API int spb200_checkpoint_num_tensors(const char* path);
SPB200_API int spb200_checkpoint_tensor(const char* path, const char* key, float* dst, long capacity, int64_t* shape8, int* rank);

/* Unit-test hook for the tcgen05 implicit-GEMM convolution: one 3x3 (taps=9) or 1x1 (taps=1) stride-`stride`
 * convolution + bias (+ReLU) on NHWC 16-bit input, against which the tests run the fp32 kernel.
 * x: B*H*W*cin (16-bit, operand type of `precision`), w: cout*(taps*cin) 16-bit, bias: cout fp32,
 * y: B*(H/stride)*(W/stride)*cout, 16-bit or fp32 (out_fp32). */
SPB200_API int spb200_test_conv_tc(int precision, const void* x, const void* w, const float* bias, void* y, int B, int H, int W,
                        int cin, int cout, int taps, int stride, int relu, int out_fp32, void* stream);

/* Same hook with the kernel chosen explicitly: 0 = one-convolution kernel (conv_tc.cu), 1 = persistent per-tap
 * block kernel run as a single convolution (block_tc.cu), 2 = haloed-tile kernel (halo_tc.cu; stride 1,
 * cout = 128 only; SPB200_E_INVALID when the convolution does not fit it). */
SPB200_API int spb200_test_conv_kernel(int kernel, int precision, const void* x, const void* w, const float* bias, void* y, int B,
                            int H, int W, int cin, int cout, int taps, int stride, int relu, int out_fp32, void* stream);

/* Diagnostic: with SPB200_HALO_DBG=<variant><index> in the environment one haloed-tile plan records clock stamps of
 * its CTA 0 (16 tile pairs x 8 stamps: MMA issuer start / GEMM 1 issued / Y seen / GEMM 2 issued, first epilogue warp
 * accumulator 1 seen / epilogue 1 done / accumulator 2 seen / epilogue 2 done).  Copies the 128 stamps to `host`;
 * returns 1 when no plan was instrumented.  Read by scripts/halo_dbg.py; no counterpart in the reference. */
SPB200_API int spb200_debug_halo(long long* host);

#ifdef __cplusplus
}
#endif
#endif /* SPB200_H */
