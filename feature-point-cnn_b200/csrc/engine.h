// spb200::Engine - the SuperPoint/MagicPoint inference engine behind the C ABI (include/spb200.h).
//
// Owns the folded/packed weights and the per-shape workspace on one CUDA device; runs the reference's
// hot path (python/src/superpoint.py:91-115 -> python/src/netutils.py:78-121) as a fixed sequence of
// kernel launches on the caller's stream.  Not thread-safe per instance; instances are independent.
#pragma once
#include <array>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "ckpt_reader.h"
#include "kernels.h"

namespace spb200 {

struct Params {
    float conf_thresh = 0.015f;   // python/src/settings.py:5
    int nms_dist = 4;             // python/src/settings.py:4
    int border_remove = 4;        // python/src/settings.py:8
    int top_k = 0;                // 0 = every survivor (the reference has no top-k)
    int descriptor_enabled = 1;   // SuperPoint.is_descriptor_enabled, python/src/superpoint.py:67
};

struct HostConv {                 // BatchNorm-folded convolution, [cout][cin][kh][kw]
    int cout = 0, cin = 0, kh = 0, kw = 0;
    bool synthetic = false;       // not a convolution of the reference (an identity shortcut run as MMAs): no algorithmic FLOPs
    std::vector<float> w, b;
    float at(int co, int ci, int y, int x) const { return w[(((size_t)co * cin + ci) * kh + y) * kw + x]; }
};

struct TapSpec { int dy, dx, kh, kw; };

struct SegSpec {
    int src_buf;
    const HostConv* conv;
    int ci_off, cin_real;
    int stride;
    std::vector<TapSpec> taps;
    int view = 1, vy = 0, vx = 0;   // view > 1: the segment reads the pixels (view*y + vy, view*x + vx) of its source
    int koff_tail = 0;              // > 0: first K index of the packed K tail (common.cuh, SegDev)
};

struct OpSpec {
    std::string name;
    std::vector<SegSpec> segs;
    int cout_real = 0, cout_pad = 0;
    int dst_buf = -1, res_buf = -1;
    bool relu = true, dst_fp32 = false;
    int dst_stride = 1, off_y = 0, off_x = 0;
    int K = 0;                    // packed K: every segment's taps x stored channels, then (split sources) the lo-weight ranges
    int K_main = 0;               // K without the lo-weight ranges
    std::vector<float> bias;      // [cout_pad]
    // device
    float* d_bias = nullptr;
    float* d_w32 = nullptr;       // [K][cout_pad]
    void* d_w16 = nullptr;        // [cout_pad][K]
    TcConvPlan* plan = nullptr;   // per workspace shape (unfused tensor-core path, kept for A/B runs)
    TcBlockPlan* fused = nullptr; // per workspace shape: this op fused with the next one (block) or alone
    TcHaloPlan* halo = nullptr;   // per workspace shape: haloed-tile kernel (preferred over `fused` when it applies)
    bool fused_skip = false;      // executed as the second half of the previous op's fused plan
    int side = -1;                // >= 0: one of a group of independent ops that run side by side, each on its share
                                  // of the SMs (the four transposed-convolution phases); index inside the group
};

enum BufId {
    BUF_POOL, BUF_L1A_Y, BUF_L1A, BUF_L1B_Y, BUF_L1B, BUF_L2A_Y, BUF_L2A, BUF_L2B_Y, BUF_FEAT,
    BUF_D0_Y, BUF_D0, BUF_D1_Y, BUF_LOGITS,
    BUF_I0_Y, BUF_I0, BUF_I1_Y, BUF_I1, BUF_UP, BUF_O0_Y, BUF_O0, BUF_O1_Y, BUF_DESC,
    BUF_FEAT_HI,                  // the hi halves of a split BUF_FEAT as a plain 16-bit tensor: what the descriptor head reads
    BUF_COUNT
};

// C: channels per pixel (padded); split: stored in the split-precision layout (common.cuh, SegDev), 2 C values per pixel
struct BufSpec {
    int div; int C; bool fp32; bool split;
    int stored() const { return split ? 2 * C : C; }
};

// Stages whose convolutions run with split-precision operands (three MMAs per product, fp32-grade results).  Only
// prefixes of the network are accepted: the stem, + layer1, + layer2, + the detector head.
enum SplitLevel { SPLIT_NONE = 0, SPLIT_LAYER1 = 1, SPLIT_LAYER2 = 2, SPLIT_DETECTOR = 3 };

class Engine {
public:
    explicit Engine(int device);
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    void load_checkpoint(const std::string& path);
    void load_tensor(const std::string& key, const float* data, const int64_t* shape, int rank);
    void finalize(int precision, int split_level = SPLIT_NONE);
    int split_level() const { return split_level_; }
    void set_params(const Params& p) { params_ = p; clear_graphs(); }
    const Params& params() const { return params_; }
    int precision() const { return precision_; }
    int descriptor_dim() const { return 128; }
    bool finalized() const { return finalized_; }

    // reference SuperPoint.forward (python/src/superpoint.py:91-115); all pointers are device memory,
    // desc_nchw / logits_nchw may be null
    void forward(const float* img, int B, int C, int H, int W, float* prob, float* desc_nchw, float* logits_nchw,
                 cudaStream_t st);
    // reference InferenceWrapper.run per image of the batch (python/src/inferencewrapper.py:29-46)
    void detect(const float* img, int B, int C, int H, int W, int cap, int* count, int* xy, float* conf, float* desc,
                float* prob, cudaStream_t st);
    void detect_host(const float* img, int B, int C, int H, int W, int cap, int* count, int* xy, float* conf,
                     float* desc);
    // the same for 8-bit grayscale frames [B][H][W] (what the reference's loaders divide by 255:
    // python/src/inference.py:72-85, cpp/src/camera.cc:12-23); value k means k / 255
    void detect_u8(const uint8_t* img, int B, int H, int W, int cap, int* count, int* xy, float* conf, float* desc,
                   float* prob, cudaStream_t st);
    void detect_host_u8(const uint8_t* img, int B, int H, int W, int cap, int* count, int* xy, float* conf, float* desc);
    // the two halves of detect_host, so that two batches can be in flight (the download of one under the compute of the
    // next): submit enqueues upload + compute and returns a ticket (0 / 1), wait downloads and returns the results
    int detect_host_submit(const void* img, bool img_u8, int B, int C, int H, int W, int cap, int* count, int* xy, float* conf,
                           void* desc, int chunk_pref = 0);
    void detect_host_wait(int ticket);
    // descriptors of detect / detect_host as fp32 (0, the reference's type) or fp16 (1: half the bytes over the bus)
    void set_descriptor_format(int fmt);
    int descriptor_format() const { return desc_fp16_ ? 1 : 0; }
    // stage-level entry points (reference restore_prob_map / get_points / get_descriptors)
    void heatmap_from_logits(const float* logits_nchw, int B, int H, int W, float* prob, cudaStream_t st);
    void restore_prob_map(const float* softmax_nchw, int B, int H, int W, float* prob, cudaStream_t st);
    // the demos' frame loaders (preproc.cu): frames B*h*w*C bytes -> gray B*H*W bytes; frames B*h*w*3 fp32 BGR -> B*3*H*W fp32 RGB
    void preprocess_u8(const uint8_t* frames, int B, int h, int w, int C, uint8_t* out, int H, int W, cudaStream_t st);
    void preprocess_f32(const float* frames, int B, int h, int w, float* out, int H, int W, cudaStream_t st);
    void nms(const float* prob, int B, int H, int W, int cap, int* count, int* xy, float* conf, cudaStream_t st);
    void sample_descriptors(const float* desc_nchw, int B, int D, int H, int W, int cap, const int* count,
                            const int* xy, float* out, cudaStream_t st);

    // reference get_best_correspondences (python/src/inference.py:88-96): mutual nearest neighbours of two batches of
    // descriptor sets as spb200_detect returns them; max_dist <= 0 disables the distance gate
    void match(const float* desc_a, const int* count_a, const float* desc_b, const int* count_b, int B, int cap, int D,
               float max_dist, int* match_ab, float* dist, cudaStream_t st);

    // reference homography_adaptation (python/src/homographies.py:250-324): homographies = num flattened transforms
    // [num][8] on the HOST (sampled by the caller), aggregation 0 = mean ('sum'), 1 = max; prob_out [B][H][W] device
    void homography_adaptation(const float* img, int B, int C, int H, int W, const float* homographies, int num, int margin,
                               int aggregation, float* prob_out, cudaStream_t st);

    // intermediate activation (debug / parity tests): copies buffer `id` as NCHW fp32 to dst (device)
    void export_buffer(int id, float* dst_nchw, int channels, cudaStream_t st);
    void buffer_dims(int id, int* C, int* H, int* W) const;

    // per-launch CUDA-event timing (bench.py roofline): begin, run any inference calls, end
    struct ProfEntry { std::string name; cudaEvent_t start, stop; double flops, bytes; float ms; };
    void profile_begin();
    const std::vector<ProfEntry>& profile_end();

    long launches() const { return launches_; }
    void reset_launches() { launches_ = 0; }
    static int max_keypoints(int H, int W, int nms_dist);

    std::string last_error;

private:
    void build_ops();
    void ensure_workspace(int B, int C, int H, int W, cudaStream_t st);
    void ensure_nms(int B, int H, int W);
    void release_workspace();
    void release_weights();
    // heat_out: when the detector's last block can (tc_halo_heat_capable) it writes exp(logit) there, depth-to-space, and the
    // per-cell normaliser to d_inv_ instead of the logits; returns whether it did
    bool run_network(const void* img, bool img_u8, int B, int C, int H, int W, cudaStream_t st, float* heat_out = nullptr);
    void detect_any(const void* img, bool img_u8, int B, int C, int H, int W, int cap, int* count, int* xy, float* conf,
                    float* desc, float* prob, cudaStream_t st);
    void detect_body(const void* img, bool img_u8, int B, int C, int H, int W, int cap, int* count, int* xy, float* conf,
                     float* desc, float* prob, cudaStream_t st);
    // CUDA graphs of whole detect calls: the ~20 launches (and the fork / join of the side streams) of a call repeated with
    // the same buffers - a streaming caller, the chunks of the host pipeline - are captured the second time the call is
    // seen and replayed from then on (one cudaGraphLaunch instead of ~20 launches of host work).  SPB200_NO_GRAPH=1 disables.
    struct GraphKey {
        const void* img; int u8, B, C, H, W, cap; const void *count, *xy, *conf, *desc, *prob;
        bool operator<(const GraphKey& o) const { return std::memcmp(this, &o, sizeof(GraphKey)) < 0; }
    };
    struct GraphEntry { cudaGraphExec_t exec = nullptr; int seen = 0; long launches = 0; bool failed = false; };
    std::map<GraphKey, GraphEntry> graphs_;
    bool use_graphs_ = true;
    bool zero_sum_rounding_ = true;   // SPB200_ROUND_NEAREST=1: plain round-to-nearest of the 16-bit weights
    void clear_graphs();
    void detect_host_any(const void* img, bool img_u8, int B, int C, int H, int W, int cap, int* count, int* xy, float* conf,
                         float* desc);
    ConvDev make_conv_dev(const OpSpec& op) const;
    const HostConv* fold(const std::string& conv_key, const std::string& bn_key, bool transposed, bool has_bias);
    void add_block(const std::string& prefix, std::vector<std::pair<int, int>> srcs, int y_buf, int dst_buf, int stride,
                   int cout, bool dst_fp32);

    int device_;
    int precision_ = PREC_FP32;
    int split_level_ = SPLIT_NONE;
    bool desc_fp16_ = false;
    bool finalized_ = false;
    Params params_;
    StateDict sd_;
    std::vector<std::unique_ptr<HostConv>> convs_;
    std::vector<OpSpec> ops_;
    std::array<BufSpec, BUF_COUNT> bufspec_{};
    int det_c_ = 80;
    int num_sms_ = 148;
    bool use_halo_ = true;        // SPB200_NO_HALO=1 keeps every block on the per-tap kernel
    bool fuse_blocks_ = true;     // SPB200_NO_FUSE=1 in the environment runs one kernel per convolution
    float* d_stem_w_[2] = {nullptr, nullptr};      // [0]: 1-channel (gray-folded), [1]: 3-channel
    float* d_stem_b_ = nullptr;
    void* d_stem_w16_[2] = {nullptr, nullptr};     // tensor-core stem: [64][nchunk*64] 16-bit, K-major
    StemTcPlan* stem_plan_[2] = {nullptr, nullptr};
    int cout_pad_of(int cout, int y_buf) const;
    StemPlanesPlan* stem_planes_ = nullptr;        // gray stem fed by TMA from parity planes (stem_planes.cu)
    bool use_planes_ = true;      // SPB200_OLD_STEM=1 keeps the im2col-in-shared-memory stem
    void* d_planes_ = nullptr;    // [B][2][H/2][W] 16-bit image x255, per workspace shape
    float* d_imgf_ = nullptr;     // 8-bit frames / 255 for the paths without the plane-fed stem

    // workspace: buffers sized for capB_ images of wsH_ x wsW_; the tensor-core plans (tile counts, tensor maps) depend on
    // the batch actually run, wsB_ <= capB_, and are kept per batch size so that alternating sizes costs nothing
    int wsB_ = 0, wsH_ = 0, wsW_ = 0, capB_ = 0;
    struct OpPlans { TcConvPlan* plan = nullptr; TcBlockPlan* fused = nullptr; TcHaloPlan* halo = nullptr; bool fused_skip = false; };
    std::map<int, std::vector<OpPlans>> plan_cache_;
    void build_plans();
    void stash_plans();
    void destroy_plan_cache();
    std::array<void*, BUF_COUNT> buf_{};
    float* d_prob_ = nullptr;
    float* d_inv_ = nullptr;      // [capB][H/8][W/8] per-cell softmax normaliser of the fused detector tail
    // nms workspace
    int nmsB_ = 0, nmsH_ = 0, nmsW_ = 0, nmsR_ = -1;
    bool nms_dirty_ = true;       // the NMS counters are not known to be zero (fresh workspace, or a call that did not get through)
    NmsWorkspace nms_{};
    // descriptor sampling positions per pixel column / row (see launch_sample_descriptors)
    float* d_gtab_ = nullptr;
    int gtabH_ = 0, gtabW_ = 0;
    const float* grid_table(int H, int W);
    int* d_pre_tab_ = nullptr;                     // interpolation tables of the frame loaders, per (kind, h, w, H, W)
    int pre_key_[5] = {0, 0, 0, 0, 0};
    const int* preprocess_table(int kind, int h, int w, int H, int W);
    unsigned long long* d_match_ws_ = nullptr;     // [2][B][cap] best keys of the matcher
    size_t match_ws_elems_ = 0;
    void* d_match_tc_ws_ = nullptr;                // operands + norms of the tensor-core matcher
    size_t match_tc_ws_bytes_ = 0;
    // homography adaptation workspace
    float* d_ha_img_ = nullptr; float* d_ha_prob_ = nullptr; float* d_ha_coeffs_ = nullptr; uint8_t* d_ha_maps_ = nullptr;
    size_t ha_img_elems_ = 0, ha_prob_elems_ = 0, ha_map_bytes_ = 0; int ha_num_ = 0;
    // detect_host staging
    struct HostStage;
    std::unique_ptr<HostStage> stage_;

    // fork / join of the side-by-side groups: three extra streams, one event to fork and one per stream to join
    cudaStream_t side_stream_[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t side_fork_ = nullptr, side_join_[3] = {nullptr, nullptr, nullptr};
    bool use_phases_ = true;      // SPB200_NO_PHASES=1 keeps stride-2 blocks on the per-tap kernel
    bool fused_planes_ = true;    // SPB200_NO_FUSED_PLANES=1: fp32 frames go through the plane pass like 8-bit ones
    bool fused_heat_ = true;      // SPB200_NO_FUSED_HEAT=1: the detector tail writes logits, the NMS computes the softmax values
    bool use_side_ = true;        // SPB200_NO_SIDE=1 runs the phases one after the other on all SMs
    // The workspace (activations, NMS lists, tables, side streams) is shared by every call: a call enqueued on a stream
    // other than the previous call's first waits for the event the previous call recorded when it finished enqueueing.
    struct StreamScope {
        Engine* e; cudaStream_t st;
        StreamScope(Engine* eng, cudaStream_t s);
        ~StreamScope();
    };
    cudaEvent_t ws_done_ = nullptr;
    cudaStream_t ws_stream_ = nullptr;
    bool ws_used_ = false;
    long launches_ = 0;
    bool profiling_ = false;
    std::vector<ProfEntry> prof_;
    cudaEvent_t prof_mark(cudaStream_t st);
    void prof_open(const std::string& name, double flops, double bytes, cudaStream_t st);
    void prof_close(cudaStream_t st);
    double op_flops(const OpSpec& op) const;
};

}  // namespace spb200
