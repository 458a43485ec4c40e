// spb200::Engine implementation: weight folding/packing, workspace, launch sequence.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>

namespace spb200 {

bool pdl_enabled() {
    static const bool on = [] { const char* e = std::getenv("SPB200_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}

namespace {

constexpr float kBnEps = 1e-5f;   // nn.BatchNorm2d default (reference python/src/resnet_blocks.py:8)

template <typename T>
T* dev_alloc(size_t n) {
    void* p = nullptr;
    SPB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    return static_cast<T*>(p);
}

template <typename T>
T* dev_upload(const std::vector<T>& v) {
    T* p = dev_alloc<T>(v.size());
    if (!v.empty()) SPB_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
}

uint16_t f32_to_f16_bits(float f) { return __half_as_ushort(__float2half_rn(f)); }
uint16_t f32_to_bf16_bits(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }

int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

// Host-buffer pipeline state of detect_host: three streams (upload, compute, download) and two SLOTS of device input /
// output buffers, so that two batches can be in flight: while the keypoints and descriptors of batch i cross the bus, the
// GPU computes batch i + 1 (spb200_detect_host_submit / _wait; spb200_detect_host is submit + wait).  Per slot: whole-batch
// buffers cut into chunks, one event pair and one pinned count mirror per chunk, and pinned staging used only when the
// caller's buffers are pageable.
struct Engine::HostStage {
    static constexpr int kMaxChunks = 64;
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    struct Slot {
        cudaEvent_t ev_in[kMaxChunks] = {}, ev_comp[kMaxChunks] = {};
        void* d_img = nullptr;
        size_t d_img_bytes = 0;
        int* d_count = nullptr;
        int* d_xy = nullptr;
        float* d_conf = nullptr;
        void* d_desc = nullptr;
        int* h_count = nullptr;        // pinned [B]
        int d_B = 0, d_cap = 0;
        // pageable callers
        void* h_img = nullptr;         // pinned, whole batch
        size_t h_img_bytes = 0;
        int* h_xy = nullptr;
        float* h_conf = nullptr;
        uint8_t* h_desc = nullptr;
        size_t h_out_cap = 0;          // keypoints the pinned output staging holds
        // the job in flight
        bool busy = false;             // submitted, not yet waited for
        bool done = false;             // its results are in the caller's arrays (or `error` says why not)
        std::string error;
        int B = 0, cap = 0;
        size_t desc_row = 0;           // bytes of one descriptor (128 x 4 or 128 x 2)
        std::vector<int> csize, cstart;
        int* o_count = nullptr; int* o_xy = nullptr; float* o_conf = nullptr; uint8_t* o_desc = nullptr;   // the caller's arrays
    } slot[2];
    int next = 0;
    // The download half runs on a thread of its own: it waits for each chunk's counts, issues the (count-sized) copies
    // on the download stream and signals the slot.  The downloads of consecutive batches then follow each other without
    // waiting for the caller to come back from its own work (PCIe device-to-host is the longest stage of the pipeline).
    int device = 0;
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<int> queue;
    bool quit = false;
    void download(Slot& sl);
    void run() {
        cudaSetDevice(device);
        for (;;) {
            int t;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return quit || !queue.empty(); });
                if (queue.empty()) return;
                t = queue.front();
                queue.pop_front();
            }
            std::string err;
            try {
                download(slot[t]);
            } catch (const std::exception& e) {
                err = e.what();
            } catch (...) {
                err = "unknown error in the download thread";
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                slot[t].error = err;
                slot[t].done = true;
            }
            cv.notify_all();
        }
    }
    void init() {
        SPB_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
        SPB_CUDA(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
        SPB_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
        for (auto& sl : slot)
            for (int i = 0; i < kMaxChunks; ++i) {
                SPB_CUDA(cudaEventCreateWithFlags(&sl.ev_in[i], cudaEventDisableTiming));
                SPB_CUDA(cudaEventCreateWithFlags(&sl.ev_comp[i], cudaEventDisableTiming));
            }
    }
    ~HostStage() {
        if (worker.joinable()) {
            {
                std::lock_guard<std::mutex> lk(mu);
                quit = true;
            }
            cv.notify_all();
            worker.join();
        }
        for (auto& sl : slot) {
            cudaFree(sl.d_img); cudaFree(sl.d_count); cudaFree(sl.d_xy); cudaFree(sl.d_conf); cudaFree(sl.d_desc);
            if (sl.h_count) cudaFreeHost(sl.h_count);
            if (sl.h_img) cudaFreeHost(sl.h_img);
            if (sl.h_xy) cudaFreeHost(sl.h_xy);
            if (sl.h_conf) cudaFreeHost(sl.h_conf);
            if (sl.h_desc) cudaFreeHost(sl.h_desc);
            for (int i = 0; i < kMaxChunks; ++i) {
                if (sl.ev_in[i]) cudaEventDestroy(sl.ev_in[i]);
                if (sl.ev_comp[i]) cudaEventDestroy(sl.ev_comp[i]);
            }
        }
        if (s_in) cudaStreamDestroy(s_in);
        if (s_comp) cudaStreamDestroy(s_comp);
        if (s_out) cudaStreamDestroy(s_out);
    }
};

Engine::Engine(int device) : device_(device) {
    int n = 0;
    SPB_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) throw std::invalid_argument("invalid CUDA device index");
    SPB_CUDA(cudaSetDevice(device_));
    cudaDeviceProp prop{};
    SPB_CUDA(cudaGetDeviceProperties(&prop, device_));
    if (prop.major < 10)
        throw std::runtime_error(std::string("spb200 needs a Blackwell (sm_100a) GPU, found ") + prop.name);
    num_sms_ = prop.multiProcessorCount;
    const char* nf = std::getenv("SPB200_NO_FUSE");
    fuse_blocks_ = !(nf && nf[0] == '1');
    const char* nh = std::getenv("SPB200_NO_HALO");
    use_halo_ = !(nh && nh[0] == '1');
    const char* ns = std::getenv("SPB200_NO_SIDE");
    use_side_ = !(ns && ns[0] == '1');
    const char* nfp = std::getenv("SPB200_NO_FUSED_PLANES");
    fused_planes_ = !(nfp && nfp[0] == '1');
    const char* nfh = std::getenv("SPB200_NO_FUSED_HEAT");
    fused_heat_ = !(nfh && nfh[0] == '1');
    for (int i = 0; i < 3; ++i) {
        SPB_CUDA(cudaStreamCreateWithFlags(&side_stream_[i], cudaStreamNonBlocking));
        SPB_CUDA(cudaEventCreateWithFlags(&side_join_[i], cudaEventDisableTiming));
    }
    SPB_CUDA(cudaEventCreateWithFlags(&side_fork_, cudaEventDisableTiming));
    SPB_CUDA(cudaEventCreateWithFlags(&ws_done_, cudaEventDisableTiming));
    const char* np = std::getenv("SPB200_NO_PHASES");
    use_phases_ = !(np && np[0] == '1');
    const char* ng = std::getenv("SPB200_NO_GRAPH");
    use_graphs_ = !(ng && ng[0] == '1');
    const char* rn = std::getenv("SPB200_ROUND_NEAREST");
    zero_sum_rounding_ = !(rn && rn[0] == '1');
    const char* os = std::getenv("SPB200_OLD_STEM");
    use_planes_ = !(os && os[0] == '1');
    buf_.fill(nullptr);
}

Engine::StreamScope::StreamScope(Engine* eng, cudaStream_t s) : e(eng), st(s) {
    SPB_CUDA(cudaSetDevice(e->device_));
    if (e->ws_used_ && e->ws_stream_ != st) SPB_CUDA(cudaStreamWaitEvent(st, e->ws_done_, 0));
}

Engine::StreamScope::~StreamScope() {
    if (cudaEventRecord(e->ws_done_, st) == cudaSuccess) {
        e->ws_stream_ = st;
        e->ws_used_ = true;
    } else {
        cudaGetLastError();
    }
}

Engine::~Engine() {
    cudaSetDevice(device_);
    clear_graphs();
    if (ws_done_) cudaEventDestroy(ws_done_);
    release_workspace();
    release_weights();
    cudaFree(nms_.keys); cudaFree(nms_.keys_alt); cudaFree(nms_.counters); cudaFree(nms_.mask); cudaFree(nms_.und); cudaFree(nms_.ukey);
    for (int i = 0; i < 3; ++i) {
        if (side_stream_[i]) cudaStreamDestroy(side_stream_[i]);
        if (side_join_[i]) cudaEventDestroy(side_join_[i]);
    }
    if (side_fork_) cudaEventDestroy(side_fork_);
    cudaFree(d_gtab_);
    cudaFree(d_match_ws_);
    cudaFree(d_pre_tab_);
    cudaFree(d_match_tc_ws_);
    cudaFree(d_ha_img_); cudaFree(d_ha_prob_); cudaFree(d_ha_coeffs_); cudaFree(d_ha_maps_);
}

void Engine::destroy_plan_cache() {
    for (auto& kv : plan_cache_)
        for (auto& pl : kv.second) {
            if (pl.plan) tc_plan_destroy(pl.plan);
            if (pl.fused) tc_block_plan_destroy(pl.fused);
            if (pl.halo) tc_halo_plan_destroy(pl.halo);
        }
    plan_cache_.clear();
}

// the current plans go (back) into the cache under the batch size they were built for
void Engine::stash_plans() {
    if (wsB_ <= 0 || precision_ == PREC_FP32) return;
    std::vector<OpPlans>& v = plan_cache_[wsB_];
    v.resize(ops_.size());
    for (size_t i = 0; i < ops_.size(); ++i) {
        v[i] = OpPlans{ops_[i].plan, ops_[i].fused, ops_[i].halo, ops_[i].fused_skip};
        ops_[i].plan = nullptr; ops_[i].fused = nullptr; ops_[i].halo = nullptr;
    }
}

void Engine::release_workspace() {
    clear_graphs();
    stash_plans();
    destroy_plan_cache();
    for (auto& p : buf_) { cudaFree(p); p = nullptr; }
    cudaFree(d_prob_); d_prob_ = nullptr;
    cudaFree(d_inv_); d_inv_ = nullptr;
    cudaFree(d_planes_); d_planes_ = nullptr;
    cudaFree(d_imgf_); d_imgf_ = nullptr;
    wsB_ = wsH_ = wsW_ = capB_ = 0;
}

void Engine::release_weights() {
    stash_plans();
    destroy_plan_cache();
    for (auto& op : ops_) {
        cudaFree(op.d_bias); cudaFree(op.d_w32); cudaFree(op.d_w16);
    }
    ops_.clear();
    convs_.clear();
    cudaFree(d_stem_w_[0]); cudaFree(d_stem_w_[1]); cudaFree(d_stem_b_);
    d_stem_w_[0] = d_stem_w_[1] = d_stem_b_ = nullptr;
    for (int i = 0; i < 2; ++i) {
        if (stem_plan_[i]) stem_tc_plan_destroy(stem_plan_[i]);
        stem_plan_[i] = nullptr;
        cudaFree(d_stem_w16_[i]);
        d_stem_w16_[i] = nullptr;
    }
    if (stem_planes_) {
        stem_planes_plan_destroy(stem_planes_);
        stem_planes_ = nullptr;
    }
}

void Engine::load_checkpoint(const std::string& path) {
    std::string err;
    StateDict sd;
    if (!read_checkpoint(path, sd, err)) throw std::runtime_error("failed to load checkpoint " + path + ": " + err);
    restore_module_prefixes(sd);                       // *_params.pt of InferenceWrapper.trace
    sd_ = std::move(sd);
    finalized_ = false;
}

void Engine::load_tensor(const std::string& key, const float* data, const int64_t* shape, int rank) {
    HostTensor t;
    t.shape.assign(shape, shape + rank);
    t.data.assign(data, data + t.numel());
    sd_[key] = std::move(t);
    finalized_ = false;
}

// conv (no bias unless has_bias) followed by eval-mode BatchNorm: w' = w * s, b' = beta - mean * s (+ bias * s),
// s = gamma / sqrt(var + eps).  A transposed conv's weight is [cin][cout][kh][kw]; it is stored here
// re-indexed as [cout][cin][kh][kw] without flipping (the taps are enumerated explicitly).
const HostConv* Engine::fold(const std::string& conv_key, const std::string& bn_key, bool transposed, bool has_bias) {
    auto need = [&](const std::string& k) -> const HostTensor& {
        auto it = sd_.find(k);
        if (it == sd_.end()) throw std::runtime_error("checkpoint is missing key '" + k + "'");
        return it->second;
    };
    const HostTensor& w = need(conv_key + ".weight");
    if (w.shape.size() != 4) throw std::runtime_error(conv_key + ".weight is not 4-d");
    auto hc = std::make_unique<HostConv>();
    hc->cout = (int)(transposed ? w.shape[1] : w.shape[0]);
    hc->cin = (int)(transposed ? w.shape[0] : w.shape[1]);
    hc->kh = (int)w.shape[2];
    hc->kw = (int)w.shape[3];
    const HostTensor& g = need(bn_key + ".weight");
    const HostTensor& be = need(bn_key + ".bias");
    const HostTensor& mu = need(bn_key + ".running_mean");
    const HostTensor& var = need(bn_key + ".running_var");
    for (const HostTensor* t : {&g, &be, &mu, &var})
        if (t->numel() != hc->cout) throw std::runtime_error(bn_key + " has the wrong number of channels");
    const HostTensor* cb = has_bias ? &need(conv_key + ".bias") : nullptr;
    if (cb && cb->numel() != hc->cout) throw std::runtime_error(conv_key + ".bias has the wrong number of elements");
    if (w.numel() != (int64_t)w.shape[0] * w.shape[1] * w.shape[2] * w.shape[3] || hc->cout <= 0 || hc->cin <= 0 || hc->kh <= 0 || hc->kw <= 0)
        throw std::runtime_error(conv_key + ".weight has an invalid shape");
    hc->w.resize((size_t)hc->cout * hc->cin * hc->kh * hc->kw);
    hc->b.resize(hc->cout);
    for (int co = 0; co < hc->cout; ++co) {
        const float s = g.data[co] / std::sqrt(var.data[co] + kBnEps);
        hc->b[co] = be.data[co] - mu.data[co] * s + (cb ? cb->data[co] * s : 0.f);
        for (int ci = 0; ci < hc->cin; ++ci)
            for (int y = 0; y < hc->kh; ++y)
                for (int x = 0; x < hc->kw; ++x) {
                    const size_t src = transposed ? ((((size_t)ci * hc->cout + co) * hc->kh + y) * hc->kw + x)
                                                  : ((((size_t)co * hc->cin + ci) * hc->kh + y) * hc->kw + x);
                    hc->w[(((size_t)co * hc->cin + ci) * hc->kh + y) * hc->kw + x] = w.data[src] * s;
                }
    }
    convs_.push_back(std::move(hc));
    return convs_.back().get();
}

static std::vector<TapSpec> taps3x3() {
    std::vector<TapSpec> t;
    for (int y = 0; y < 3; ++y)
        for (int x = 0; x < 3; ++x) t.push_back({y - 1, x - 1, y, x});
    return t;
}

// One residual block (reference python/src/resnet_blocks.py:14-27) as two implicit GEMMs:
//   Y   = relu(conv3x3_s(X) + b1)
//   OUT = relu(conv1x1(Y) [+ conv1x1_s(X) as extra K] + b2 [+ bd] [+ X])
// `srcs` lists the source buffers that are concatenated along channels (one, or two for layer_out.0,
// python/src/superpoint.py:59) with the channel count each contributes.
void Engine::add_block(const std::string& p, std::vector<std::pair<int, int>> srcs, int y_buf, int dst_buf,
                       int stride, int cout, bool dst_fp32) {
    const bool has_ds = sd_.count(p + ".identity_downsample.0.weight") > 0;
    const HostConv* c1 = fold(p + ".conv1", p + ".bn1", false, false);
    const HostConv* c2 = fold(p + ".conv2", p + ".bn2", false, false);
    const HostConv* cd_folded = has_ds ? fold(p + ".identity_downsample.0", p + ".identity_downsample.1", false, false) : nullptr;
    int cin_total = 0;
    for (auto& s : srcs) cin_total += s.second;
    // A 64-channel identity block on the tensor-core path adds its shortcut ON the tensor core, as x . I into the second
    // accumulator (exact: 16-bit x times 1.0, fp32 accumulation): four more MMAs per tile, but the second epilogue then
    // neither waits for the shortcut tile to land in its staging box nor reads it (halo_tc.cu keeps ONE box per tile at
    // N = 64, shared memory is full; encoder.layer1.1 121 -> see DESIGN.md section 9).
    const HostConv* cd = cd_folded;
    // split-precision blocks (their source is stored as hi + lo) always do: x.hi . I + x.lo . I is the exact fp32-grade shortcut
    const bool split_in = bufspec_[srcs[0].first].split;
    if (!cd && ((split_in && srcs.size() == 1 && stride == 1 && cin_total == cout) ||
                (precision_ != PREC_FP32 && use_halo_ && fuse_blocks_ && srcs.size() == 1 && stride == 1 && cout == 64 &&
                 cin_total == 64 && bufspec_[y_buf].C == 64 && bufspec_[srcs[0].first].C == 64 && !std::getenv("SPB200_NO_IDENTITY_MMA")))) {
        auto hc = std::make_unique<HostConv>();
        hc->cout = hc->cin = cout; hc->kh = hc->kw = 1;
        hc->synthetic = true;
        hc->w.assign((size_t)cout * cout, 0.f);
        for (int i = 0; i < cout; ++i) hc->w[(size_t)i * cout + i] = 1.f;
        hc->b.assign(cout, 0.f);
        convs_.push_back(std::move(hc));
        cd = convs_.back().get();
    }
    if (c1->cin != cin_total || c1->cout != cout || c1->kh != 3 || c2->cin != cout || c2->cout != cout || c2->kh != 1)
        throw std::runtime_error("unexpected convolution shapes in block " + p);
    if (cd && (cd->cin != cin_total || cd->cout != cout || cd->kh != 1))
        throw std::runtime_error("unexpected downsample shape in block " + p);
    const int cout_pad = cout_pad_of(cout, y_buf);

    // A stride-2 block on the tensor-core path with one source and at most 128 output channels is read as stride-1
    // convolutions over the four pixel phases of its input (row parity a, column parity b): output (oy, ox) takes
    // input row 2 oy + ky - 1, i.e. phase a = 1 at row oy - 1 (ky = 0), a = 0 at oy (ky = 1), a = 1 at oy (ky = 2), and
    // the same along x.  Each phase is then ONE haloed tile for the halo kernel instead of nine per-tap tiles.
    const bool phases = stride == 2 && precision_ != PREC_FP32 && use_halo_ && fuse_blocks_ && srcs.size() == 1 &&
                        cout_pad <= 128 && bufspec_[srcs[0].first].C == 64 && has_ds && use_phases_;
    OpSpec a;
    a.name = p + ".conv1";
    int off = 0;
    if (phases) {
        // phase (0, 0), the one the shortcut convolution reads, comes last: its step is then the last of GEMM 1 and
        // everything before it overlaps the previous tile's final epilogue (halo_tc.cu, lazy wait on the D2 accumulator)
        for (int pa = 1; pa >= 0; --pa)
            for (int pb = 1; pb >= 0; --pb) {
                std::vector<TapSpec> taps;
                for (int ky = 0; ky < 3; ++ky)
                    for (int kx = 0; kx < 3; ++kx)
                        if ((ky != 1) == (pa == 1) && (kx != 1) == (pb == 1)) taps.push_back({ky == 0 ? -1 : 0, kx == 0 ? -1 : 0, ky, kx});
                SegSpec sg{srcs[0].first, c1, 0, srcs[0].second, 1, taps};
                sg.view = 2; sg.vy = pa; sg.vx = pb;
                a.segs.push_back(sg);
            }
    }
    for (auto& s : srcs) {
        if (phases) break;
        a.segs.push_back(SegSpec{s.first, c1, off, s.second, stride, taps3x3()});
        off += s.second;
    }
    a.cout_real = cout; a.cout_pad = cout_pad; a.dst_buf = y_buf; a.relu = true;
    a.bias.assign(cout_pad, 0.f);
    std::copy(c1->b.begin(), c1->b.end(), a.bias.begin());
    ops_.push_back(std::move(a));

    OpSpec b;
    b.name = p + ".conv2";
    b.segs.push_back(SegSpec{y_buf, c2, 0, cout, 1, {TapSpec{0, 0, 0, 0}}});
    b.bias.assign(cout_pad, 0.f);
    std::copy(c2->b.begin(), c2->b.end(), b.bias.begin());
    if (cd) {
        off = 0;
        for (auto& s : srcs) {
            SegSpec sg{s.first, cd, off, s.second, phases ? 1 : stride, {TapSpec{0, 0, 0, 0}}};
            if (phases) sg.view = 2;                 // the 1x1 stride-2 shortcut reads phase (0, 0)
            b.segs.push_back(sg);
            off += s.second;
        }
        for (int i = 0; i < cout; ++i) b.bias[i] += cd->b[i];
    } else {
        if (srcs.size() != 1 || stride != 1) throw std::runtime_error("identity shortcut needs one same-size source: " + p);
        b.res_buf = srcs[0].first;
    }
    b.cout_real = cout; b.cout_pad = cout_pad; b.dst_buf = dst_buf; b.relu = true; b.dst_fp32 = dst_fp32;
    ops_.push_back(std::move(b));
}

void Engine::build_ops() {
    const int dc = det_c_;
    // a buffer is stored in the split layout when the stage that READS it runs split-precision convolutions
    const int L = precision_ == PREC_FP32 ? (int)SPLIT_NONE : split_level_;
    const bool s1 = L >= SPLIT_LAYER1, s2 = L >= SPLIT_LAYER2, s3 = L >= SPLIT_DETECTOR;
    bufspec_[BUF_POOL] = {4, 64, false, s1};
    for (int b : {BUF_L1A_Y, BUF_L1A, BUF_L1B_Y}) bufspec_[b] = {4, 64, false, s1};
    bufspec_[BUF_L1B] = {4, 64, false, s2};
    for (int b : {BUF_L2A_Y, BUF_L2A, BUF_L2B_Y}) bufspec_[b] = {8, 128, false, s2};
    bufspec_[BUF_FEAT] = {8, 128, false, s3};
    // 65 detector channels: 128 per pixel (the kernels' N), or 3 split chunks of 32
    for (int b : {BUF_D0_Y, BUF_D0, BUF_D1_Y}) bufspec_[b] = {8, s3 ? 96 : dc, false, s3};
    bufspec_[BUF_LOGITS] = {8, dc, true, false};
    for (int b : {BUF_I0_Y, BUF_I0, BUF_I1_Y, BUF_I1}) bufspec_[b] = {16, 256, false, false};
    for (int b : {BUF_UP, BUF_O0_Y, BUF_O0, BUF_O1_Y, BUF_DESC}) bufspec_[b] = {8, 128, false, false};
    bufspec_[BUF_FEAT_HI] = {8, s3 ? 128 : 0, false, false};
    const int feat_desc = s3 ? BUF_FEAT_HI : BUF_FEAT;               // what the descriptor head reads

    // encoder (reference python/src/superpoint.py:16-17), detector (:32), descriptor (:43-50)
    add_block("encoder.layer1.0", {{BUF_POOL, 64}}, BUF_L1A_Y, BUF_L1A, 1, 64, false);
    add_block("encoder.layer1.1", {{BUF_L1A, 64}}, BUF_L1B_Y, BUF_L1B, 1, 64, false);
    add_block("encoder.layer2.0", {{BUF_L1B, 64}}, BUF_L2A_Y, BUF_L2A, 2, 128, false);
    add_block("encoder.layer2.1", {{BUF_L2A, 128}}, BUF_L2B_Y, BUF_FEAT, 1, 128, false);
    add_block("detector.layer.0", {{BUF_FEAT, 128}}, BUF_D0_Y, BUF_D0, 1, 65, false);
    add_block("detector.layer.1", {{BUF_D0, 65}}, BUF_D1_Y, BUF_LOGITS, 1, 65, true);
    const size_t first_desc_op = ops_.size();
    add_block("descriptor.layer_in.0", {{feat_desc, 128}}, BUF_I0_Y, BUF_I0, 2, 256, false);
    add_block("descriptor.layer_in.1", {{BUF_I0, 256}}, BUF_I1_Y, BUF_I1, 1, 256, false);
    // ConvTranspose2d(256,128,k3,s2,p1,op1)+bias -> BN -> ReLU (python/src/superpoint.py:45-47,55-57) as four
    // output phases: out[2m+py, 2n+px]; even outputs use kernel index 1 from input m, odd outputs kernel
    // index 2 from input m and kernel index 0 from input m+1 (out index o = 2i - 1 + k).
    const HostConv* up = fold("descriptor.up_sample", "descriptor.bn", true, true);
    if (up->cin != 256 || up->cout != 128 || up->kh != 3) throw std::runtime_error("unexpected up_sample shape");
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
            OpSpec o;
            o.name = "descriptor.up_sample.phase" + std::to_string(py) + std::to_string(px);
            std::vector<TapSpec> taps;
            const int ky[2][2] = {{1, -1}, {2, 0}}, dyv[2] = {0, 1};
            for (int a = 0; a < 2; ++a) {
                if (ky[py][a] < 0) continue;
                for (int c = 0; c < 2; ++c) {
                    if (ky[px][c] < 0) continue;
                    taps.push_back({dyv[a], dyv[c], ky[py][a], ky[px][c]});
                }
            }
            o.segs.push_back(SegSpec{BUF_I1, up, 0, 256, 1, taps});
            o.cout_real = 128; o.cout_pad = 128; o.dst_buf = BUF_UP; o.relu = true;
            o.dst_stride = 2; o.off_y = py; o.off_x = px;
            o.side = py * 2 + px;
            o.bias = up->b;
            ops_.push_back(std::move(o));
        }
    add_block("descriptor.layer_out.0", {{BUF_UP, 128}, {feat_desc, 128}}, BUF_O0_Y, BUF_O0, 1, 128, false);
    add_block("descriptor.layer_out.1", {{BUF_O0, 128}}, BUF_O1_Y, BUF_DESC, 1, 128, false);
    (void)first_desc_op;

    // pack and upload
    auto round16 = [&](float f) {
        return precision_ == PREC_FP16 ? __half2float(__float2half_rn(f)) : __bfloat162float(__float2bfloat16_rn(f));
    };
    for (auto& op : ops_) {
        int K = 0;
        bool any_split = false;
        for (auto& s : op.segs) {
            K += (int)s.taps.size() * bufspec_[s.src_buf].stored();
            any_split = any_split || bufspec_[s.src_buf].split;
        }
        for (auto& s : op.segs)
            if (bufspec_[s.src_buf].split != any_split) throw std::runtime_error("internal: mixed split / plain sources in " + op.name);
        op.K_main = K;
        op.K = any_split ? 2 * K : K;        // split sources: the lo-weight ranges mirror the main ranges
        // packed K tail (common.cuh, SegDev): a full 3x3 segment whose last chunk holds 1..16 real channels
        for (auto& s : op.segs) {
            const int cpad = bufspec_[s.src_buf].stored();
            const int in_last = s.cin_real - (cpad / 64 - 1) * 64;
            bool full3x3 = s.taps.size() == 9;
            for (size_t t = 0; t < s.taps.size() && full3x3; ++t)
                full3x3 = s.taps[t].dy == (int)t / 3 - 1 && s.taps[t].dx == (int)t % 3 - 1;
            s.koff_tail = 0;
            if (precision_ != PREC_FP32 && !any_split && full3x3 && cpad >= 128 && cpad % 64 == 0 && in_last >= 1 && in_last <= 16 &&
                s.view == 1 && s.stride == 1) {
                s.koff_tail = op.K;
                op.K += 3 * 64;
            }
        }
        K = op.K;
        std::vector<float> w32((size_t)K * op.cout_pad, 0.f);
        int koff = 0;
        for (auto& s : op.segs) {
            const int cpad = bufspec_[s.src_buf].stored();
            for (size_t t = 0; t < s.taps.size(); ++t)
                for (int ci = 0; ci < s.cin_real; ++ci)
                    for (int co = 0; co < op.cout_real; ++co) {
                        const float w = s.conv->at(co, s.ci_off + ci, s.taps[t].kh, s.taps[t].kw);
                        if (!any_split) {
                            w32[(size_t)(koff + (int)t * cpad + ci) * op.cout_pad + co] = w;
                            continue;
                        }
                        // split layout (common.cuh, SegDev): channel ci sits at (ci / 32) * 64 + ci % 32 (hi) and + 32 (lo);
                        // w.hi multiplies both, w.lo - kept only where the weight is not exact in 16 bits - the hi half
                        const int j = (ci / 32) * 64 + ci % 32;
                        const size_t row = (size_t)(koff + (int)t * cpad + j);
                        const float hi = round16(w);
                        w32[row * op.cout_pad + co] = hi;
                        w32[(row + 32) * op.cout_pad + co] = hi;
                        if (!s.conv->synthetic) w32[(row + (size_t)op.K_main) * op.cout_pad + co] = w - hi;
                    }
            koff += (int)s.taps.size() * cpad;
        }
        op.d_bias = dev_upload(op.bias);
        if (precision_ != PREC_FP32 && !any_split && zero_sum_rounding_) {
            // Zero-sum rounding: the activations entering a convolution are post-ReLU, i.e. non-negative with a
            // mean of the order of their spread, so the rounding errors d_k of the weights of one output channel contribute
            // mean(a) * sum_k d_k + a fluctuation.  Round-to-nearest leaves sum_k d_k a random walk (sqrt(n) half-ulps);
            // here, per output channel and (source, tap) group of input channels, the weights closest to a rounding midpoint are
            // rounded the other way until the group's errors cancel to within a quarter ulp.  Every weight is still one of its
            // two neighbouring 16-bit values; nothing changes at run time.  Emulated and measured: -20 % heatmap error.
            int k0 = 0;
            std::vector<int> order;
            std::vector<float> err, ulp;
            for (auto& sg : op.segs) {
                const int cpad = bufspec_[sg.src_buf].stored();
                for (size_t t = 0; t < sg.taps.size(); ++t, k0 += cpad) {
                    const int n = sg.cin_real;
                    for (int co = 0; co < op.cout_real; ++co) {
                        order.clear(); err.assign(n, 0.f); ulp.assign(n, 0.f);
                        double tot = 0.0;
                        for (int k = 0; k < n; ++k) {
                            float& v = w32[(size_t)(k0 + k) * op.cout_pad + co];
                            if (v == 0.f) continue;                       // padding and exact zeros stay zero
                            const float r = round16(v);
                            int e = 0;
                            std::frexp(r, &e);                             // |r| in [2^(e-1), 2^e)
                            const int mant = precision_ == PREC_FP16 ? 11 : 8;
                            const int emin = precision_ == PREC_FP16 ? -14 : -126;
                            ulp[k] = std::ldexp(1.f, std::max(e - 1, emin) - (mant - 1));
                            err[k] = r - v;
                            tot += err[k];
                            v = r;
                            if (err[k] != 0.f) order.push_back(k);
                        }
                        std::sort(order.begin(), order.end(), [&](int a, int b) { return std::fabs(err[a]) / ulp[a] > std::fabs(err[b]) / ulp[b]; });
                        for (int k : order) {
                            if (std::fabs(tot) < 0.25 * ulp[k]) break;
                            if (err[k] * tot <= 0) continue;                // flipping this one would make the sum worse
                            const double step = err[k] > 0 ? -(double)ulp[k] : (double)ulp[k];
                            if (std::fabs(tot + step) >= std::fabs(tot)) continue;
                            float& v = w32[(size_t)(k0 + k) * op.cout_pad + co];
                            const float moved = round16(v + (float)step);   // the neighbour on the other side of the exact value
                            tot += (double)moved - (double)v;
                            v = moved;
                        }
                    }
                }
            }
        }
        // the packed K tails repeat the (rounded) weights of the main range
        {
            int k0 = 0;
            for (auto& sg : op.segs) {
                const int cpad = bufspec_[sg.src_buf].stored();
                if (sg.koff_tail > 0)
                    for (int t = 0; t < 9; ++t)
                        for (int i = 0; i < 16; ++i)
                            for (int co = 0; co < op.cout_pad; ++co)
                                w32[(size_t)(sg.koff_tail + (t / 3) * 64 + (t % 3) * 16 + i) * op.cout_pad + co] =
                                    w32[(size_t)(k0 + t * cpad + (cpad - 64) + i) * op.cout_pad + co];
                k0 += (int)sg.taps.size() * cpad;
            }
        }
        if (precision_ == PREC_FP32) {
            op.d_w32 = dev_upload(w32);
        } else {
            std::vector<uint16_t> w16((size_t)op.cout_pad * K);
            for (int k = 0; k < K; ++k)
                for (int co = 0; co < op.cout_pad; ++co) {
                    const float v = w32[(size_t)k * op.cout_pad + co];
                    w16[(size_t)co * K + k] = precision_ == PREC_FP16 ? f32_to_f16_bits(v) : f32_to_bf16_bits(v);
                }
            op.d_w16 = dev_upload(w16);
        }
    }

    // stem: conv7x7 s2 p3 (no bias) + BN (python/src/superpoint.py:12-13); the 1-channel variant sums the
    // three input-channel kernels, exact for a gray image replicated to RGB (python/src/dataset_utils.py:18-20)
    const HostConv* st = fold("encoder.conv1", "encoder.bn1", false, false);
    if (st->cin != 3 || st->cout != 64 || st->kh != 7) throw std::runtime_error("unexpected encoder.conv1 shape");
    std::vector<float> w3((size_t)3 * 49 * 64), w1((size_t)49 * 64, 0.f);
    for (int c = 0; c < 3; ++c)
        for (int y = 0; y < 7; ++y)
            for (int x = 0; x < 7; ++x)
                for (int co = 0; co < 64; ++co) {
                    const float v = st->at(co, c, y, x);
                    w3[((size_t)(c * 49 + y * 7 + x)) * 64 + co] = v;
                    w1[((size_t)(y * 7 + x)) * 64 + co] += v;
                }
    d_stem_w_[0] = dev_upload(w1);
    d_stem_w_[1] = dev_upload(w3);
    d_stem_b_ = dev_upload(st->b);
    if (precision_ != PREC_FP32) {
        // tensor-core stem: [cout][k] with k = c*64 + ky*8 + kx (kx and ky padded 7 -> 8 with zeros: the eight taps
        // of one filter row are one 16-byte piece of the im2col row, stem_tc.cu), followed by the lo parts
        // (w - float(w16)) for the split-precision MMAs
        for (int v = 0; v < 2; ++v) {
            const int cin = v == 0 ? 1 : 3, kpad = 64 * cin, nparts = 2;
            const std::vector<float>& src = v == 0 ? w1 : w3;
            std::vector<uint16_t> w16((size_t)64 * kpad * nparts, 0);
            auto to16 = [&](float f) { return precision_ == PREC_FP16 ? f32_to_f16_bits(f) : f32_to_bf16_bits(f); };
            auto from16 = [&](uint16_t h) {
                return precision_ == PREC_FP16 ? __half2float(__ushort_as_half(h)) : __bfloat162float(__ushort_as_bfloat16(h));
            };
            for (int c = 0; c < cin; ++c)
                for (int ky = 0; ky < 7; ++ky)
                    for (int kx = 0; kx < 7; ++kx)
                        for (int co = 0; co < 64; ++co) {
                            const float f = src[(size_t)(c * 49 + ky * 7 + kx) * 64 + co];
                            const int k = c * 64 + ky * 8 + kx;
                            const uint16_t hi = to16(f);
                            w16[(size_t)co * kpad * nparts + k] = hi;
                            w16[(size_t)co * kpad * nparts + kpad + k] = to16(f - from16(hi));
                        }
            d_stem_w16_[v] = dev_upload(w16);
            stem_plan_[v] = stem_tc_plan_create(d_stem_w16_[v], d_stem_b_, cin, precision_, num_sms_);
            if (v == 0) stem_planes_ = stem_planes_plan_create(d_stem_w16_[0], d_stem_b_, precision_, num_sms_);
        }
    }
}

int Engine::cout_pad_of(int cout, int y_buf) const {
    return precision_ == PREC_FP32 ? bufspec_[y_buf].C : round_up(cout, 64);
}

void Engine::finalize(int precision, int split_level) {
    if (precision != PREC_FP32 && precision != PREC_FP16 && precision != PREC_BF16)
        throw std::invalid_argument("precision must be 0 (fp32 CUDA cores), 1 (fp16 tcgen05) or 2 (bf16 tcgen05)");
    if (split_level < SPLIT_NONE || split_level > SPLIT_DETECTOR)
        throw std::invalid_argument("split level must be 0 (none), 1 (stem + layer1), 2 (+ layer2) or 3 (+ detector)");
    if (split_level != SPLIT_NONE && (precision == PREC_FP32 || !use_halo_ || !fuse_blocks_))
        throw std::invalid_argument("split-precision stages need the fused tensor-core path (precision fp16 / bf16)");
    if (sd_.empty()) throw std::runtime_error("no weights loaded");
    SPB_CUDA(cudaSetDevice(device_));
    SPB_CUDA(cudaDeviceSynchronize());
    release_workspace();
    release_weights();
    precision_ = precision;
    split_level_ = split_level;
    det_c_ = precision == PREC_FP32 ? 80 : 128;
    build_ops();
    finalized_ = true;
}

int Engine::max_keypoints(int H, int W, int nms_dist) {
    const int r = std::max(nms_dist, 0);
    return ((H + r) / (r + 1)) * ((W + r) / (r + 1));
}

void Engine::ensure_workspace(int B, int C, int H, int W, cudaStream_t st) {
    if (!finalized_) throw std::runtime_error("weights are not finalized (call spb200_finalize_weights)");
    if (B <= 0 || H <= 0 || W <= 0 || H % 16 != 0 || W % 16 != 0)
        throw std::invalid_argument("image height and width must be positive multiples of 16");
    if (C != 1 && C != 3) throw std::invalid_argument("images must have 1 or 3 channels");
    if (B == wsB_ && H == wsH_ && W == wsW_) return;
    if (B <= capB_ && H == wsH_ && W == wsW_) {
        // the buffers are large enough: only the plans change (from the cache when this batch size was seen before)
        stash_plans();
        wsB_ = B;
        auto it = plan_cache_.find(B);
        if (it != plan_cache_.end()) {
            for (size_t i = 0; i < ops_.size(); ++i) {
                ops_[i].plan = it->second[i].plan; ops_[i].fused = it->second[i].fused; ops_[i].halo = it->second[i].halo;
                ops_[i].fused_skip = it->second[i].fused_skip;
            }
            plan_cache_.erase(it);
        } else {
            build_plans();
        }
        return;
    }
    SPB_CUDA(cudaStreamSynchronize(st));
    SPB_CUDA(cudaDeviceSynchronize());
    const int cap = (H == wsH_ && W == wsW_) ? std::max(B, capB_) : B;
    release_workspace();
    const size_t esz = precision_ == PREC_FP32 ? 4 : 2;
    for (int i = 0; i < BUF_COUNT; ++i) {
        const BufSpec& bs = bufspec_[i];
        const size_t n = (size_t)cap * (H / bs.div) * (W / bs.div) * bs.stored();
        const size_t bytes = n * (bs.fp32 ? 4 : esz);
        if (bytes == 0) continue;
        SPB_CUDA(cudaMalloc(&buf_[i], bytes));
        SPB_CUDA(cudaMemset(buf_[i], 0, bytes));     // padded channels stay zero forever
    }
    d_prob_ = dev_alloc<float>((size_t)cap * H * W);
    d_inv_ = dev_alloc<float>((size_t)cap * (H / 8) * (W / 8));
    if (precision_ != PREC_FP32) SPB_CUDA(cudaMalloc(&d_planes_, (size_t)cap * H * W * 2));
    wsB_ = B; wsH_ = H; wsW_ = W; capB_ = cap;
    build_plans();
}

// tensor-core plans of every op for the batch size wsB_ (tile counts, tensor maps over the workspace buffers)
void Engine::build_plans() {
    if (precision_ != PREC_FP32) {
        for (size_t i = 0; i < ops_.size(); ++i) {
            OpSpec& op = ops_[i];
            op.fused_skip = false;
            if (!fuse_blocks_) { op.plan = tc_plan_create(make_conv_dev(op), precision_); continue; }
            const bool is_conv1 = op.name.size() > 6 && op.name.compare(op.name.size() - 6, 6, ".conv1") == 0;
            if (is_conv1 && i + 1 < ops_.size()) {
                const ConvDev c1 = make_conv_dev(op), c2 = make_conv_dev(ops_[i + 1]);
                if (use_halo_) op.halo = tc_halo_plan_create(c1, &c2, precision_, op.cout_real, num_sms_);
                if (!op.halo && op.segs[0].view > 1) throw std::runtime_error("internal: no haloed-tile plan for the phase form of " + op.name);
                if (!op.halo && bufspec_[op.segs[0].src_buf].split) throw std::runtime_error("internal: no split-precision plan for " + op.name);
                if (!op.halo) op.fused = tc_block_plan_create(c1, &c2, precision_, op.cout_real, num_sms_);
                ops_[i + 1].fused_skip = true;
                ++i;
            } else {
                const ConvDev c1 = make_conv_dev(op);
                // side-by-side ops share the SMs in proportion to their measured cost (1, 2, 2 and 4 taps of the same
                // tile count: 21 : 23 : 23 : 30 us when each runs alone)
                int sms = num_sms_;
                if (op.side >= 0 && use_side_) {
                    static const int share[4] = {32, 35, 35, 46};
                    sms = std::max(1, num_sms_ * share[op.side] / 148);
                }
                if (use_halo_) op.halo = tc_halo_plan_create(c1, nullptr, precision_, op.cout_real, sms);
                if (!op.halo) op.fused = tc_block_plan_create(c1, nullptr, precision_, op.cout_real, sms);
            }
        }
    }
}

void Engine::ensure_nms(int B, int H, int W) {
    const int r = params_.nms_dist;
    if (B <= nmsB_ && H == nmsH_ && W == nmsW_ && r == nmsR_) return;
    SPB_CUDA(cudaDeviceSynchronize());
    clear_graphs();
    cudaFree(nms_.keys); cudaFree(nms_.keys_alt); cudaFree(nms_.counters); cudaFree(nms_.mask); cudaFree(nms_.und); cudaFree(nms_.ukey);
    nms_.kcap = max_keypoints(H, W, r);
    nms_.mask_w = (W + 31) / 32;
    nms_.keys = dev_alloc<unsigned long long>((size_t)B * nms_.kcap);
    nms_.keys_alt = dev_alloc<unsigned long long>((size_t)B * nms_.kcap);
    nms_.counters = dev_alloc<int>((size_t)B * 8);
    nms_.mask = dev_alloc<unsigned>((size_t)B * H * nms_.mask_w);
    nms_.und = dev_alloc<unsigned>((size_t)B * H * W);
    nms_.ukey = dev_alloc<unsigned>((size_t)B * H * W);
    nmsB_ = B; nmsH_ = H; nmsW_ = W; nmsR_ = r;
    nms_dirty_ = true;
}

// ix = ((gx + 1) / 2) * (Wc - 1) with gx = float(x / (W / 2.) - 1.): the arithmetic of the reference's
// get_descriptors (python/src/netutils.py:110-115) followed by grid_sample's align_corners=True un-normalisation.
const float* Engine::grid_table(int H, int W) {
    if (d_gtab_ && H == gtabH_ && W == gtabW_) return d_gtab_;
    SPB_CUDA(cudaDeviceSynchronize());
    clear_graphs();
    cudaFree(d_gtab_);
    std::vector<float> t((size_t)W + H);
    const int Hc = H / 8, Wc = W / 8;
    for (int x = 0; x < W; ++x) {
        const float gx = (float)((double)x / ((double)W / 2.) - 1.);
        t[x] = ((gx + 1.f) / 2.f) * (float)(Wc - 1);
    }
    for (int y = 0; y < H; ++y) {
        const float gy = (float)((double)y / ((double)H / 2.) - 1.);
        t[(size_t)W + y] = ((gy + 1.f) / 2.f) * (float)(Hc - 1);
    }
    d_gtab_ = dev_upload(t);
    gtabH_ = H; gtabW_ = W;
    return d_gtab_;
}

ConvDev Engine::make_conv_dev(const OpSpec& op) const {
    ConvDev d{};
    const BufSpec& ds = bufspec_[op.dst_buf];
    d.nseg = (int)op.segs.size();
    int koff = 0;
    for (int i = 0; i < d.nseg; ++i) {
        const SegSpec& s = op.segs[i];
        const BufSpec& bs = bufspec_[s.src_buf];
        SegDev& sd = d.seg[i];
        sd.src = buf_[s.src_buf];
        sd.H = wsH_ / bs.div; sd.W = wsW_ / bs.div; sd.C = bs.stored();
        sd.view = s.view; sd.full_H = sd.H; sd.full_W = sd.W;
        if (s.view > 1) {
            const size_t esz = (bs.fp32 || precision_ == PREC_FP32) ? 4 : 2;
            sd.src = static_cast<const char*>(buf_[s.src_buf]) + ((size_t)s.vy * sd.full_W + s.vx) * bs.stored() * esz;
            sd.H /= s.view; sd.W /= s.view;
        }
        sd.cin = bs.stored();
        sd.cin_real = s.cin_real;
        sd.split = bs.split ? 1 : 0;
        sd.koff_lo = (bs.split && !(s.conv && s.conv->synthetic)) ? op.K_main + koff : -1;
        sd.ntaps = (int)s.taps.size();
        sd.stride = s.stride;
        sd.koff = koff;
        sd.koff_tail = s.koff_tail;
        for (int t = 0; t < sd.ntaps; ++t) { sd.dy[t] = (int8_t)s.taps[t].dy; sd.dx[t] = (int8_t)s.taps[t].dx; }
        koff += sd.ntaps * bs.stored();
    }
    d.w = precision_ == PREC_FP32 ? (const void*)op.d_w32 : (const void*)op.d_w16;
    d.bias = op.d_bias;
    d.residual = op.res_buf >= 0 ? buf_[op.res_buf] : nullptr;
    d.res_C = op.res_buf >= 0 ? bufspec_[op.res_buf].stored() : 0;
    d.dst = buf_[op.dst_buf];
    d.B = wsB_;
    d.dst_H = wsH_ / ds.div; d.dst_W = wsW_ / ds.div; d.dst_C = ds.stored();
    d.split_out = ds.split ? 1 : 0;
    d.dst_stride = op.dst_stride; d.dst_off_y = op.off_y; d.dst_off_x = op.off_x;
    d.OH = d.dst_H / op.dst_stride; d.OW = d.dst_W / op.dst_stride;
    d.K = op.K; d.cout_pad = op.cout_pad;
    d.relu = op.relu ? 1 : 0;
    d.dst_fp32 = (op.dst_fp32 || precision_ == PREC_FP32) ? 1 : 0;
    return d;
}

// ---- per-launch profiling -------------------------------------------------------------------------
void Engine::profile_begin() {
    for (auto& e : prof_) { cudaEventDestroy(e.start); cudaEventDestroy(e.stop); }
    prof_.clear();
    profiling_ = true;
}

const std::vector<Engine::ProfEntry>& Engine::profile_end() {
    profiling_ = false;
    SPB_CUDA(cudaDeviceSynchronize());
    for (auto& e : prof_) SPB_CUDA(cudaEventElapsedTime(&e.ms, e.start, e.stop));
    return prof_;
}

void Engine::prof_open(const std::string& name, double flops, double bytes, cudaStream_t st) {
    if (!profiling_) return;
    ProfEntry e{name, nullptr, nullptr, flops, bytes, 0.f};
    SPB_CUDA(cudaEventCreate(&e.start));
    SPB_CUDA(cudaEventCreate(&e.stop));
    SPB_CUDA(cudaEventRecord(e.start, st));
    prof_.push_back(e);
}

void Engine::prof_close(cudaStream_t st) {
    if (!profiling_) return;
    SPB_CUDA(cudaEventRecord(prof_.back().stop, st));
}

// algorithmic FLOPs of one launch: 2 x MACs of the reference convolution, real (unpadded) channels
double Engine::op_flops(const OpSpec& op) const {
    const BufSpec& ds = bufspec_[op.dst_buf];
    const double rows = (double)wsB_ * (wsH_ / ds.div / op.dst_stride) * (wsW_ / ds.div / op.dst_stride);
    double k = 0;
    for (auto& s : op.segs)
        if (!(s.conv && s.conv->synthetic)) k += (double)s.taps.size() * s.cin_real;
    return 2.0 * rows * k * op.cout_real;
}

bool Engine::run_network(const void* img_any, bool img_u8, int B, int C, int H, int W, cudaStream_t st, float* heat_out) {
    bool heat_written = false;
    ensure_workspace(B, C, H, W, st);
    if (img_u8 && C != 1) throw std::invalid_argument("8-bit frames must be single-channel (grayscale)");
    const float* img = static_cast<const float*>(img_any);
    const bool wide_stem = bufspec_[BUF_POOL].split;      // split level >= 1: image hi + lo, three MMAs per product (stem_tc.cu)
    if (img_u8 && !(precision_ != PREC_FP32 && use_planes_ && !wide_stem)) {
        // paths without the plane-fed stem take fp32 images: frame / 255 as the reference's loaders do
        // (python/src/inference.py:78-80, cpp/src/camera.cc:16-18)
        if (!d_imgf_) SPB_CUDA(cudaMalloc((void**)&d_imgf_, sizeof(float) * (size_t)capB_ * H * W));
        launch_u8_to_f32(static_cast<const uint8_t*>(img_any), d_imgf_, (long)B * H * W, st);
        ++launches_;
        img = d_imgf_;
        img_u8 = false;
    }
    // gray-folded stem: 49 MACs per output (the reference's 3-channel stem does 147 on replicated input)
    const bool planes = precision_ != PREC_FP32 && C == 1 && use_planes_ && !wide_stem;
    // fp32 frames feed the plane-fed stem directly (its shifter warps convert on the way); 8-bit frames go through the plane pass
    const bool direct = planes && !img_u8 && fused_planes_ && (reinterpret_cast<uintptr_t>(img_any) & 15) == 0;
    if (planes && !direct) {
        prof_open("image_planes", 0.0, (double)B * H * W * ((img_u8 ? 1 : 4) + 2), st);
        launch_planes(img_any, img_u8 ? 1 : 0, d_planes_, precision_, B, H, W, st);
        prof_close(st);
        ++launches_;
    }
    prof_open("stem_pool", 2.0 * B * (H / 2) * (W / 2) * 64.0 * 49.0 * C,
              (double)B * C * H * W * (planes && !direct ? 2 : 4) + (double)B * (H / 4) * (W / 4) * 64 * (precision_ == PREC_FP32 ? 4 : 2), st);
    if (precision_ == PREC_FP32)
        launch_stem_pool(img, B, C, H, W, d_stem_w_[C == 1 ? 0 : 1], d_stem_b_, buf_[BUF_POOL], precision_, st);
    else if (wide_stem)
        launch_stem_wide(stem_plan_[C == 1 ? 0 : 1], img, buf_[BUF_POOL], B, H, W, st);
    else if (planes)
        launch_stem_planes(stem_planes_, direct ? img_any : d_planes_, direct ? 1 : 0, buf_[BUF_POOL], B, H, W, st);
    else
        launch_stem_tc(stem_plan_[C == 1 ? 0 : 1], img, buf_[BUF_POOL], B, H, W, st);
    prof_close(st);
    ++launches_;
    bool feat_hi_done = !bufspec_[BUF_FEAT].split;
    for (size_t i = 0; i < ops_.size(); ++i) {
        OpSpec& op = ops_[i];
        if (!params_.descriptor_enabled && op.name.compare(0, 11, "descriptor.") == 0) continue;
        if (op.fused_skip) continue;
        if (!feat_hi_done && op.name.compare(0, 11, "descriptor.") == 0) {
            // the descriptor head is not split: it reads the hi halves of the encoder features as a plain tensor
            const long npix = (long)B * (H / 8) * (W / 8);
            prof_open("features_hi", 0.0, (double)npix * 128 * 2 * 3, st);
            launch_split_hi(buf_[BUF_FEAT], buf_[BUF_FEAT_HI], npix, 128, st);
            prof_close(st);
            ++launches_;
            feat_hi_done = true;
        }
        if (precision_ == PREC_FP32) {
            prof_open(op.name, op_flops(op), 0.0, st);
            launch_conv_simt(make_conv_dev(op), st);
        } else if (op.fused || op.halo) {
            const bool block = i + 1 < ops_.size() && ops_[i + 1].fused_skip;
            // a side-by-side group (consecutive ops with side = 0, 1, 2, ...): member 0 runs on the caller's stream, the
            // others on the side streams, between a fork event recorded before member 0 and one join event per side
            // stream; the profile times the group as one entry
            const bool grouped = op.side >= 0 && use_side_;
            cudaStream_t ls = st;
            if (grouped) {
                if (op.side == 0) {
                    double fl = 0.0;
                    for (size_t j = i; j < ops_.size() && ops_[j].side >= 0 && (j == i || ops_[j].side > 0); ++j) fl += op_flops(ops_[j]);
                    prof_open(op.name.substr(0, op.name.rfind('.')), fl, 0.0, st);
                    SPB_CUDA(cudaEventRecord(side_fork_, st));
                } else {
                    ls = side_stream_[op.side - 1];
                    SPB_CUDA(cudaStreamWaitEvent(ls, side_fork_, 0));
                }
            } else {
                prof_open(block ? op.name.substr(0, op.name.size() - 6) : op.name,
                          op_flops(op) + (block ? op_flops(ops_[i + 1]) : 0.0), 0.0, st);
            }
            if (op.halo && heat_out && fused_heat_ && op.name == "detector.layer.1.conv1" && tc_halo_heat_capable(op.halo)) {
                launch_halo_tc_heat(op.halo, heat_out, d_inv_, B, ls);
                heat_written = true;
            } else if (op.halo) {
                launch_halo_tc(op.halo, ls);
            } else {
                launch_block_tc(op.fused, ls);
            }
            if (ls != st) {
                SPB_CUDA(cudaEventRecord(side_join_[op.side - 1], ls));
                SPB_CUDA(cudaStreamWaitEvent(st, side_join_[op.side - 1], 0));
            }
            if (grouped) {
                ++launches_;
                const bool last = !(i + 1 < ops_.size() && ops_[i + 1].side > 0);
                if (last) prof_close(st);
                continue;
            }
        } else {
            prof_open(op.name, op_flops(op), 0.0, st);
            launch_conv_tc(op.plan, st);
        }
        prof_close(st);
        ++launches_;
    }
    return heat_written;
}

void Engine::forward(const float* img, int B, int C, int H, int W, float* prob, float* desc_nchw, float* logits_nchw,
                     cudaStream_t st) {
    StreamScope scope(this, st);
    run_network(img, false, B, C, H, W, st);
    const int Hc = H / 8, Wc = W / 8;
    float* heat = prob ? prob : d_prob_;
    launch_heatmap((const float*)buf_[BUF_LOGITS], (long)Hc * Wc * det_c_, 1, det_c_, B, Hc, Wc, heat, st);
    ++launches_;
    if (logits_nchw) {
        launch_nhwc_to_nchw(buf_[BUF_LOGITS], PREC_FP32, B, Hc * Wc, det_c_, 65, logits_nchw, st);
        ++launches_;
    }
    if (desc_nchw) {
        if (params_.descriptor_enabled) {
            launch_nhwc_to_nchw(buf_[BUF_DESC], precision_, B, Hc * Wc, 128, 128, desc_nchw, st);
            ++launches_;
        } else {
            // MagicPoint: zeros (reference python/src/superpoint.py:105-109)
            SPB_CUDA(cudaMemsetAsync(desc_nchw, 0, sizeof(float) * (size_t)B * 128 * Hc * Wc, st));
        }
    }
}

void Engine::detect(const float* img, int B, int C, int H, int W, int cap, int* count, int* xy, float* conf,
                    float* desc, float* prob, cudaStream_t st) {
    detect_any(img, false, B, C, H, W, cap, count, xy, conf, desc, prob, st);
}

void Engine::detect_u8(const uint8_t* img, int B, int H, int W, int cap, int* count, int* xy, float* conf, float* desc,
                       float* prob, cudaStream_t st) {
    detect_any(img, true, B, 1, H, W, cap, count, xy, conf, desc, prob, st);
}

void Engine::clear_graphs() {
    for (auto& kv : graphs_)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    graphs_.clear();
}

void Engine::detect_any(const void* img, bool img_u8, int B, int C, int H, int W, int cap, int* count, int* xy, float* conf,
                        float* desc, float* prob, cudaStream_t st) {
    StreamScope scope(this, st);
    if (cap <= 0) throw std::invalid_argument("capacity must be positive");
    // a capture cannot start on the legacy default stream; profiling brackets every launch with events of its own
    if (!use_graphs_ || profiling_ || st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread) {
        detect_body(img, img_u8, B, C, H, W, cap, count, xy, conf, desc, prob, st);
        return;
    }
    GraphKey key;
    std::memset(&key, 0, sizeof(key));                         // padding bytes take part in the comparison
    key.img = img; key.u8 = img_u8; key.B = B; key.C = C; key.H = H; key.W = W; key.cap = cap;
    key.count = count; key.xy = xy; key.conf = conf; key.desc = desc; key.prob = prob;
    if (graphs_.size() > 64) clear_graphs();
    GraphEntry& ge = graphs_[key];
    if (ge.exec) {
        SPB_CUDA(cudaGraphLaunch(ge.exec, st));
        launches_ += ge.launches;
        return;
    }
    if (ge.failed || ge.seen++ == 0) {
        // first sight: run it (this also brings workspace, NMS lists and tables to their final size - nothing may be
        // allocated inside a capture)
        detect_body(img, img_u8, B, C, H, W, cap, count, xy, conf, desc, prob, st);
        return;
    }
    const long before = launches_;
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    if (ok) {
        try {
            detect_body(img, img_u8, B, C, H, W, cap, count, xy, conf, desc, prob, st);
        } catch (...) {
            ok = false;
        }
        if (cudaStreamEndCapture(st, &graph) != cudaSuccess || !graph) ok = false;
    }
    if (ok && cudaGraphInstantiate(&ge.exec, graph, 0) != cudaSuccess) { ok = false; ge.exec = nullptr; }
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
        cudaGetLastError();
        launches_ = before;
        ge.failed = true;                                       // this call runs the plain way, now and later
        detect_body(img, img_u8, B, C, H, W, cap, count, xy, conf, desc, prob, st);
        return;
    }
    ge.launches = launches_ - before;
    SPB_CUDA(cudaGraphLaunch(ge.exec, st));
}

void Engine::detect_body(const void* img, bool img_u8, int B, int C, int H, int W, int cap, int* count, int* xy, float* conf,
                         float* desc, float* prob, cudaStream_t st) {
    // The detector's last block leaves the softmax itself (fused into its epilogue, halo_tc.cu): exp(logit) of the 64 pixel channels
    // in depth-to-space order and one normaliser per cell - no logits buffer, no heatmap pass, and round 0 of the NMS reads four
    // bytes per pixel (+ 4 per cell) and multiplies.  A caller who wants the heatmap gets it by one in-place scaling pass.  Where
    // that block does not run on the haloed-tile kernel (fp32 and split-precision modes, A/B switches) round 0 computes the
    // softmax values from the logits and the heatmap is written only when the caller wants it.
    ensure_workspace(B, C, H, W, st);                           // d_prob_ exists before it is handed to the network
    float* heat = prob ? prob : d_prob_;
    const bool heat_done = run_network(img, img_u8, B, C, H, W, st, heat);
    const int Hc = H / 8, Wc = W / 8;
    static const bool force_heat = [] { const char* e = std::getenv("SPB200_NMS_FROM_HEAT"); return e && e[0] == '1'; }();
    const bool from_logits = !heat_done && nms_logits_supported(params_.nms_dist) && !force_heat;
    if (heat_done && prob) {
        prof_open("heatmap_scale", 0.0, (double)B * H * W * 8, st);
        launch_heat_scale(prob, d_inv_, B, H, W, st);
        prof_close(st);
        ++launches_;
    } else if (!heat_done && (prob || !from_logits)) {
        prof_open("heatmap", 0.0, (double)B * (65.0 * Hc * Wc * 4 + (double)H * W * 4), st);
        launch_heatmap((const float*)buf_[BUF_LOGITS], (long)Hc * Wc * det_c_, 1, det_c_, B, Hc, Wc, heat, st);
        prof_close(st);
        ++launches_;
    }
    ensure_nms(B, H, W);
    // algorithmic bytes: the logits (65 channels) or the heatmap in; per survivor 8 B key out, then 8 B in and 12 B out
    // in the finish kernel (added by the caller, who knows the keypoint count)
    prof_open("nms_round0", 0.0, from_logits ? (double)B * 65.0 * Hc * Wc * 4 : (double)B * H * W * 4, st);
    const bool zero_first = nms_dirty_;
    nms_dirty_ = true;
    launch_nms_round0(from_logits ? nullptr : heat, (const float*)buf_[BUF_LOGITS], det_c_, B, H, W,
                      params_.conf_thresh, params_.nms_dist, params_.border_remove, nms_, zero_first, st,
                      heat_done && !prob ? d_inv_ : nullptr);
    prof_close(st);
    prof_open("nms_finish_sort", 0.0, 0.0, st);
    launch_nms_finish(B, H, W, params_.nms_dist, params_.border_remove, params_.top_k, cap, nms_, count, xy, conf, st);
    nms_dirty_ = false;
    prof_close(st);
    launches_ += 2;
    if (desc) {
        if (params_.descriptor_enabled) {
            prof_open("descriptors", 0.0, 0.0, st);              // bytes depend on the keypoint count (caller)
            launch_sample_descriptors(buf_[BUF_DESC], precision_, (long)Hc * Wc * 128, 1, 128, B, 128, Hc, Wc, W,
                                      grid_table(H, W), cap, count, xy, desc, desc_fp16_ ? 1 : 0, st);
            prof_close(st);
            ++launches_;
        } else {
            SPB_CUDA(cudaMemsetAsync(desc, 0, (desc_fp16_ ? sizeof(uint16_t) : sizeof(float)) * (size_t)B * cap * 128, st));
        }
    }
}

void Engine::heatmap_from_logits(const float* logits_nchw, int B, int H, int W, float* prob, cudaStream_t st) {
    StreamScope scope(this, st);
    if (H % 8 || W % 8) throw std::invalid_argument("H and W must be multiples of 8");
    const int Hc = H / 8, Wc = W / 8;
    launch_heatmap(logits_nchw, (long)65 * Hc * Wc, (long)Hc * Wc, 1, B, Hc, Wc, prob, st);
    ++launches_;
}

void Engine::restore_prob_map(const float* softmax_nchw, int B, int H, int W, float* prob, cudaStream_t st) {
    StreamScope scope(this, st);
    if (H % 8 || W % 8) throw std::invalid_argument("H and W must be multiples of 8");
    launch_depth_to_space(softmax_nchw, B, H / 8, W / 8, prob, st);
    ++launches_;
}

// one table set at a time, rebuilt when the geometry changes (a stream of frames keeps its geometry)
const int* Engine::preprocess_table(int kind, int h, int w, int H, int W) {
    const int key[5] = {kind, h, w, H, W};
    if (d_pre_tab_ && std::equal(key, key + 5, pre_key_)) return d_pre_tab_;
    std::vector<int> tab((size_t)4 * (H + W), 0);
    if (kind == 0) build_preprocess_u8_table(h, w, H, W, tab.data());
    else build_preprocess_f32_table(h, w, H, W, tab.data(), reinterpret_cast<float*>(tab.data() + 2 * (W + H)));
    SPB_CUDA(cudaDeviceSynchronize());                       // an earlier launch may still read the old table
    cudaFree(d_pre_tab_);
    d_pre_tab_ = dev_upload(tab);
    std::copy(key, key + 5, pre_key_);
    return d_pre_tab_;
}

void Engine::preprocess_u8(const uint8_t* frames, int B, int h, int w, int C, uint8_t* out, int H, int W, cudaStream_t st) {
    StreamScope scope(this, st);
    if (B <= 0 || h < 2 || w < 2 || H < 1 || W < 1 || (C != 1 && C != 3)) throw std::invalid_argument("preprocess_u8: bad geometry (C must be 1 or 3)");
    launch_preprocess_u8(frames, B, h, w, C, preprocess_table(0, h, w, H, W), out, H, W, st);
    ++launches_;
}

void Engine::preprocess_f32(const float* frames, int B, int h, int w, float* out, int H, int W, cudaStream_t st) {
    StreamScope scope(this, st);
    if (B <= 0 || h < 2 || w < 2 || H < 1 || W < 1) throw std::invalid_argument("preprocess_f32: bad geometry");
    const int* tab = preprocess_table(1, h, w, H, W);
    launch_preprocess_f32(frames, B, h, w, tab, reinterpret_cast<const float*>(tab + 2 * (W + H)), out, H, W, st);
    ++launches_;
}

void Engine::set_descriptor_format(int fmt) {
    clear_graphs();
    if (fmt != 0 && fmt != 1) throw std::invalid_argument("descriptor format must be 0 (fp32) or 1 (fp16)");
    if (stage_)
        for (auto& sl : stage_->slot)
            if (sl.busy) throw std::runtime_error("descriptor format cannot change while a host batch is in flight");

    desc_fp16_ = fmt == 1;
}

void Engine::nms(const float* prob, int B, int H, int W, int cap, int* count, int* xy, float* conf, cudaStream_t st) {
    StreamScope scope(this, st);
    if (cap <= 0) throw std::invalid_argument("capacity must be positive");
    ensure_nms(B, H, W);
    const bool zero_first = nms_dirty_;
    nms_dirty_ = true;
    launch_nms_round0(prob, nullptr, 0, B, H, W, params_.conf_thresh, params_.nms_dist, params_.border_remove, nms_, zero_first, st);
    launch_nms_finish(B, H, W, params_.nms_dist, params_.border_remove, params_.top_k, cap, nms_, count, xy, conf, st);
    nms_dirty_ = false;
    launches_ += 2;
}

void Engine::sample_descriptors(const float* desc_nchw, int B, int D, int H, int W, int cap, const int* count,
                                const int* xy, float* out, cudaStream_t st) {
    StreamScope scope(this, st);
    const int Hc = H / 8, Wc = W / 8;
    launch_sample_descriptors(desc_nchw, PREC_FP32, (long)D * Hc * Wc, (long)Hc * Wc, 1, B, D, Hc, Wc, W, grid_table(H, W), cap,
                              count, xy, out, 0, st);
    ++launches_;
}

void Engine::homography_adaptation(const float* img, int B, int C, int H, int W, const float* homographies, int num, int margin,
                                   int aggregation, float* prob_out, cudaStream_t st) {
    StreamScope scope(this, st);
    if (!img || !homographies || !prob_out) throw std::invalid_argument("homography_adaptation: null argument");
    if (num < 0 || num > 4096) throw std::invalid_argument("homography_adaptation: num must be in [0, 4096]");
    if (aggregation != 0 && aggregation != 1) throw std::invalid_argument("homography_adaptation: aggregation must be 0 (mean) or 1 (max)");
    // views (the image and its num warps) go through the network in groups of G: one batch of G * B images per forward
    // uses the GPU far better than G forwards of B (the workspace grows to G * B once; smaller batches on the same
    // engine keep using it, the plans are cached per batch size)
    const int views = num + 1;
    int G = std::max(1, std::min(views, 256 / std::max(B, 1)));
    if (const char* e = std::getenv("SPB200_HA_GROUP")) G = std::max(1, std::min(views, std::atoi(e)));
    const size_t plane = (size_t)H * W;
    const size_t img_elems = (size_t)G * B * C * plane, prob_elems = (size_t)(num + 1) * B * plane, map_bytes = (size_t)4 * num * plane;
    if (img_elems > ha_img_elems_ || prob_elems > ha_prob_elems_ || map_bytes > ha_map_bytes_ || num > ha_num_) {
        SPB_CUDA(cudaDeviceSynchronize());
        cudaFree(d_ha_img_); cudaFree(d_ha_prob_); cudaFree(d_ha_coeffs_); cudaFree(d_ha_maps_);
        ha_img_elems_ = std::max(img_elems, ha_img_elems_); ha_prob_elems_ = std::max(prob_elems, ha_prob_elems_);
        ha_map_bytes_ = std::max(map_bytes, ha_map_bytes_); ha_num_ = std::max(num, ha_num_);
        d_ha_img_ = dev_alloc<float>(ha_img_elems_);
        d_ha_prob_ = dev_alloc<float>(ha_prob_elems_);
        d_ha_coeffs_ = dev_alloc<float>((size_t)16 * std::max(ha_num_, 1));
        SPB_CUDA(cudaMalloc((void**)&d_ha_maps_, std::max(ha_map_bytes_, (size_t)16)));
    }
    // forward coefficients, then the inverses (invert_homography, homographies.py:199-203; double precision here)
    std::vector<float> coeffs((size_t)16 * std::max(num, 1), 0.f);
    for (int k = 0; k < num; ++k) {
        const float* h = homographies + (size_t)k * 8;
        const double m[9] = {h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], 1.0};
        const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
        if (!(std::fabs(det) > 1e-300)) throw std::invalid_argument("homography_adaptation: singular homography");
        const double inv[9] = {(m[4] * m[8] - m[5] * m[7]) / det, (m[2] * m[7] - m[1] * m[8]) / det, (m[1] * m[5] - m[2] * m[4]) / det,
                               (m[5] * m[6] - m[3] * m[8]) / det, (m[0] * m[8] - m[2] * m[6]) / det, (m[2] * m[3] - m[0] * m[5]) / det,
                               (m[3] * m[7] - m[4] * m[6]) / det, (m[1] * m[6] - m[0] * m[7]) / det, (m[0] * m[4] - m[1] * m[3]) / det};
        for (int i = 0; i < 8; ++i) {
            coeffs[(size_t)k * 8 + i] = h[i];
            coeffs[(size_t)(num + k) * 8 + i] = (float)(inv[i] / inv[8]);
        }
    }
    SPB_CUDA(cudaMemcpyAsync(d_ha_coeffs_, coeffs.data(), sizeof(float) * 16 * std::max(num, 1), cudaMemcpyHostToDevice, st));
    SPB_CUDA(cudaStreamSynchronize(st));                       // `coeffs` is pageable and goes out of scope
    uint8_t* raw = d_ha_maps_;
    uint8_t* eroded = d_ha_maps_ + (size_t)2 * num * plane;
    if (num) {
        launch_ha_valid_maps(d_ha_coeffs_, num, H, W, margin, raw, eroded, st);
        launches_ += margin ? 2 : 1;
    }
    // the detector on the image and on every warp of it; the descriptor head is not needed
    const int desc_was = params_.descriptor_enabled;
    params_.descriptor_enabled = 0;
    const int Hc = H / 8, Wc = W / 8;
    try {
        for (int v0 = 0; v0 < views; v0 += G) {                    // view 0 = the image itself, view k + 1 = homography k
            const int g = std::min(G, views - v0);
            const float* src = img;
            if (!(g == 1 && v0 == 0)) {
                for (int v = v0; v < v0 + g; ++v) {
                    float* slot = d_ha_img_ + (size_t)(v - v0) * B * C * plane;
                    if (v == 0) SPB_CUDA(cudaMemcpyAsync(slot, img, sizeof(float) * (size_t)B * C * plane, cudaMemcpyDeviceToDevice, st));
                    else { launch_ha_warp(img, d_ha_coeffs_ + (size_t)(v - 1) * 8, B, C, H, W, slot, st); ++launches_; }
                }
                src = d_ha_img_;
            }
            run_network(src, false, g * B, C, H, W, st);
            launch_heatmap((const float*)buf_[BUF_LOGITS], (long)Hc * Wc * det_c_, 1, det_c_, g * B, Hc, Wc,
                           d_ha_prob_ + (size_t)v0 * B * plane, st);
            ++launches_;
        }
    } catch (...) {
        params_.descriptor_enabled = desc_was;
        throw;
    }
    params_.descriptor_enabled = desc_was;
    launch_ha_aggregate(d_ha_prob_, eroded, d_ha_coeffs_, num, B, H, W, aggregation, prob_out, st);
    ++launches_;
}

void Engine::match(const float* desc_a, const int* count_a, const float* desc_b, const int* count_b, int B, int cap, int D,
                   float max_dist, int* match_ab, float* dist, cudaStream_t st) {
    StreamScope scope(this, st);
    if (B <= 0 || cap <= 0) throw std::invalid_argument("match: batch and capacity must be positive");
    const size_t need = (size_t)2 * B * cap;
    if (need > match_ws_elems_) {
        SPB_CUDA(cudaDeviceSynchronize());
        cudaFree(d_match_ws_);
        d_match_ws_ = dev_alloc<unsigned long long>(need);
        match_ws_elems_ = need;
    }
    static const bool simt = [] { const char* e = std::getenv("SPB200_MATCH_SIMT"); return e && e[0] == '1'; }();
    prof_open("match", 2.0 * B * (double)cap * cap * D, 0.0, st);
    if (D == 128 && !simt) {
        const size_t wb = match_tc_workspace_bytes(B, cap);
        if (wb > match_tc_ws_bytes_) {
            SPB_CUDA(cudaDeviceSynchronize());
            cudaFree(d_match_tc_ws_);
            SPB_CUDA(cudaMalloc(&d_match_tc_ws_, wb));
            match_tc_ws_bytes_ = wb;
        }
        launch_match_tc(desc_a, count_a, desc_b, count_b, B, cap, max_dist, d_match_tc_ws_, d_match_ws_, d_match_ws_ + (size_t)B * cap,
                        match_ab, dist, num_sms_, st);
        launches_ += 6;
    } else {
        launch_match(desc_a, count_a, desc_b, count_b, B, cap, D, max_dist, d_match_ws_, d_match_ws_ + (size_t)B * cap, match_ab, dist,
                     num_sms_, st);
        launches_ += 3;
    }
    prof_close(st);
}

void Engine::buffer_dims(int id, int* C, int* H, int* W) const {
    if (id < 0 || id >= BUF_COUNT) throw std::invalid_argument("bad buffer id");
    *C = bufspec_[id].C; *H = wsH_ / bufspec_[id].div; *W = wsW_ / bufspec_[id].div;
}

void Engine::export_buffer(int id, float* dst_nchw, int channels, cudaStream_t st) {
    if (id < 0 || id >= BUF_COUNT || !buf_[id]) throw std::invalid_argument("bad buffer id or no workspace");
    StreamScope scope(this, st);
    const BufSpec& bs = bufspec_[id];
    if (bs.split)
        launch_split_to_nchw(buf_[id], precision_, wsB_, (wsH_ / bs.div) * (wsW_ / bs.div), bs.stored(), channels, dst_nchw, st);
    else
        launch_nhwc_to_nchw(buf_[id], bs.fp32 ? PREC_FP32 : precision_, wsB_, (wsH_ / bs.div) * (wsW_ / bs.div), bs.C,
                            channels, dst_nchw, st);
}

// Host-buffer entry point: what ProcessFrame / InferenceWrapper.run do around the network (H2D of the frames, D2H
// of keypoints and descriptors), pipelined: the batch is cut into chunks; the upload of chunk k+1 and the
// download of chunk k-1 run on their own streams under the compute of chunk k.  Only count[b] keypoints of every
// image cross the bus.  Buffers the caller allocated as pinned (cudaHostAlloc / cudaHostRegister / torch
// pin_memory) are used by the DMA engines directly; pageable buffers go through pinned staging.
static bool is_pinned_host(const void* p) {
    if (!p) return true;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

void Engine::detect_host(const float* img, int B, int C, int H, int W, int cap, int* count, int* xy, float* conf,
                         float* desc) {
    detect_host_any(img, false, B, C, H, W, cap, count, xy, conf, desc);
}

void Engine::detect_host_u8(const uint8_t* img, int B, int H, int W, int cap, int* count, int* xy, float* conf, float* desc) {
    detect_host_any(img, true, B, 1, H, W, cap, count, xy, conf, desc);
}

void Engine::detect_host_any(const void* img_any, bool img_u8, int B, int C, int H, int W, int cap, int* count, int* xy,
                             float* conf, float* desc) {
    const int ticket = detect_host_submit(img_any, img_u8, B, C, H, W, cap, count, xy, conf, desc, 16);
    detect_host_wait(ticket);
}

int Engine::detect_host_submit(const void* img_any, bool img_u8, int B, int C, int H, int W, int cap, int* count, int* xy, float* conf,
                               void* desc, int chunk_pref) {
    const uint8_t* img = static_cast<const uint8_t*>(img_any);
    const size_t esz = img_u8 ? 1 : sizeof(float);
    SPB_CUDA(cudaSetDevice(device_));
    if (B <= 0) throw std::invalid_argument("batch must be positive");
    if (cap <= 0) throw std::invalid_argument("capacity must be positive");
    if (img_u8 && C != 1) throw std::invalid_argument("8-bit frames must be single-channel (grayscale)");
    if (!stage_) {
        stage_ = std::make_unique<HostStage>();
        stage_->init();
        stage_->device = device_;
        stage_->worker = std::thread([st = stage_.get()] { st->run(); });
    }
    HostStage& s = *stage_;
    const int ticket = s.next;
    HostStage::Slot& sl = s.slot[ticket];
    if (sl.busy) throw std::runtime_error("detect_host: two batches are already in flight (call spb200_detect_host_wait first)");
    int Bc = B;
    {
        // chunks: a blocking call wants them small (the download of chunk k under the compute of chunk k + 1: 16), a
        // streaming caller large (whole batches already overlap; large chunks use the GPU better: 32)
        int want = chunk_pref > 0 ? chunk_pref : 32;
        if (const char* e = std::getenv("SPB200_HOST_CHUNK")) want = std::max(1, std::atoi(e));
        // the largest divisor of B that is at most `want` (and at least a quarter of it, so that chunks stay efficient)
        if (B > want)
            for (int c = want; c >= std::max(1, want / 4); --c)
                if (B % c == 0 && B / c <= HostStage::kMaxChunks) { Bc = c; break; }
    }
    // chunk sizes: uniform.  SPB200_HOST_SPLIT=1 cuts the first and the last chunk in two (the download, the longest of
    // the three stages, then starts after half a chunk of upload + compute, and the last download, which nothing
    // overlaps, is half as long; the workspace keeps its plans per batch size, so alternating sizes costs nothing) -
    // measured 3 % slower at batch 64: eight images use the GPU too poorly, so it is off by default.
    sl.csize.clear(); sl.cstart.clear();
    {
        const int n_uniform = B / Bc;
        const bool split = n_uniform >= 3 && Bc % 2 == 0 && Bc >= 8 && n_uniform + 2 <= HostStage::kMaxChunks &&
                           std::getenv("SPB200_HOST_SPLIT") != nullptr;
        for (int k = 0; k < n_uniform; ++k) {
            if (split && (k == 0 || k == n_uniform - 1)) { sl.csize.push_back(Bc / 2); sl.csize.push_back(Bc / 2); }
            else sl.csize.push_back(Bc);
        }
        int acc = 0;
        for (int c : sl.csize) { sl.cstart.push_back(acc); acc += c; }
    }
    const int nc = (int)sl.csize.size();
    const size_t img_elem_bytes = esz * (size_t)C * H * W;
    const bool pin_in = is_pinned_host(img);
    const size_t img_bytes = img_elem_bytes * B;
    const size_t desc_row = 128 * (desc_fp16_ ? sizeof(uint16_t) : sizeof(float));

    if (img_bytes > sl.d_img_bytes || B > sl.d_B || cap > sl.d_cap) {
        SPB_CUDA(cudaDeviceSynchronize());
        cudaFree(sl.d_img); cudaFree(sl.d_count); cudaFree(sl.d_xy); cudaFree(sl.d_conf); cudaFree(sl.d_desc);
        if (sl.h_count) cudaFreeHost(sl.h_count);
        const int nb = std::max(B, sl.d_B), ncap = std::max(cap, sl.d_cap);
        const size_t nbytes = std::max(img_bytes, sl.d_img_bytes);
        SPB_CUDA(cudaMalloc(&sl.d_img, nbytes));
        sl.d_count = dev_alloc<int>(nb);
        sl.d_xy = dev_alloc<int>((size_t)nb * ncap * 2);
        sl.d_conf = dev_alloc<float>((size_t)nb * ncap);
        SPB_CUDA(cudaMalloc(&sl.d_desc, (size_t)nb * ncap * 128 * sizeof(float)));      // sized for either descriptor format
        SPB_CUDA(cudaHostAlloc((void**)&sl.h_count, sizeof(int) * nb, cudaHostAllocDefault));
        sl.d_img_bytes = nbytes; sl.d_B = nb; sl.d_cap = ncap;
    }
    if (!pin_in && img_bytes > sl.h_img_bytes) {
        SPB_CUDA(cudaDeviceSynchronize());
        if (sl.h_img) cudaFreeHost(sl.h_img);
        SPB_CUDA(cudaHostAlloc(&sl.h_img, img_bytes, cudaHostAllocDefault));
        sl.h_img_bytes = img_bytes;
    }
    const int dcap = sl.d_cap;
    sl.B = B; sl.cap = cap; sl.desc_row = desc_row;
    sl.o_count = count; sl.o_xy = xy; sl.o_conf = conf; sl.o_desc = static_cast<uint8_t*>(desc);

    // everything the GPU has to do is enqueued here: chunk k's upload on the copy stream, its network + post-processing
    // on the compute stream behind the upload's event, its counts written to the pinned mirror
    for (int k = 0; k < nc; ++k) {
        const int Bk = sl.csize[k];
        const size_t coff = (size_t)sl.cstart[k] * img_elem_bytes, chunk_bytes = (size_t)Bk * img_elem_bytes;
        const uint8_t* src = img + coff;
        uint8_t* d_in = static_cast<uint8_t*>(sl.d_img) + coff;
        if (!pin_in) {
            std::memcpy(static_cast<uint8_t*>(sl.h_img) + coff, src, chunk_bytes);
            src = static_cast<uint8_t*>(sl.h_img) + coff;
        }
        SPB_CUDA(cudaMemcpyAsync(d_in, src, chunk_bytes, cudaMemcpyHostToDevice, s.s_in));
        SPB_CUDA(cudaEventRecord(sl.ev_in[k], s.s_in));
        SPB_CUDA(cudaStreamWaitEvent(s.s_comp, sl.ev_in[k], 0));
        const size_t o = (size_t)sl.cstart[k];
        detect_any(d_in, img_u8, Bk, C, H, W, dcap, sl.d_count + o, sl.d_xy + o * dcap * 2, sl.d_conf + o * dcap,
                   desc ? static_cast<float*>(static_cast<void*>(static_cast<uint8_t*>(sl.d_desc) + o * dcap * desc_row)) : nullptr,
                   nullptr, s.s_comp);
        launch_counts_to_host(sl.d_count + o, sl.h_count + o, Bk, s.s_comp);
        ++launches_;
        SPB_CUDA(cudaEventRecord(sl.ev_comp[k], s.s_comp));
    }
    {
        std::lock_guard<std::mutex> lk(s.mu);
        sl.busy = true;
        sl.done = false;
        sl.error.clear();
        s.queue.push_back(ticket);
    }
    s.cv.notify_all();
    s.next ^= 1;
    return ticket;
}

void Engine::detect_host_wait(int ticket) {
    if (!stage_ || ticket < 0 || ticket > 1) throw std::invalid_argument("detect_host_wait: no batch in flight under this ticket");
    HostStage& s = *stage_;
    HostStage::Slot& sl = s.slot[ticket];
    std::string err;
    {
        std::unique_lock<std::mutex> lk(s.mu);
        if (!sl.busy) throw std::invalid_argument("detect_host_wait: no batch in flight under this ticket");
        s.cv.wait(lk, [&] { return sl.done; });
        sl.busy = false;
        err = sl.error;
    }
    if (!err.empty()) throw std::runtime_error(err);
}

// The download half of a batch (worker thread): as each chunk's counts arrive, its keypoints and descriptors (count[b]
// rows per image, nothing else) are copied on the download stream while later chunks - and the next batch - compute.
void Engine::HostStage::download(Slot& sl) {
    int* count = sl.o_count; int* xy = sl.o_xy; float* conf = sl.o_conf; uint8_t* desc = sl.o_desc;
    const int B = sl.B, cap = sl.cap, dcap = sl.d_cap, nc = (int)sl.csize.size();
    const size_t row = sl.desc_row;
    const bool pin_out = is_pinned_host(xy) && is_pinned_host(conf) && is_pinned_host(desc);
    size_t staged = 0;
    std::vector<size_t> stage_off;
    if (!pin_out) stage_off.assign((size_t)B, 0);
    for (int k = 0; k < nc; ++k) {
        SPB_CUDA(cudaEventSynchronize(sl.ev_comp[k]));
        const int Bk = sl.csize[k];
        const size_t g0 = (size_t)sl.cstart[k];
        size_t total = 0;
        for (int b = 0; b < Bk; ++b) {
            const size_t g = g0 + b;
            count[g] = std::min(std::max(sl.h_count[g], 0), cap);
            total += (size_t)count[g];
        }
        if (!pin_out && staged + total > sl.h_out_cap) {
            // grow the pinned staging; what is already in flight must land first, then it is copied over
            SPB_CUDA(cudaStreamSynchronize(s_out));
            const size_t n = (staged + total) * 2 + 1024;
            int* nxy = nullptr; float* ncf = nullptr; uint8_t* nds = nullptr;
            SPB_CUDA(cudaHostAlloc((void**)&nxy, sizeof(int) * 2 * n, cudaHostAllocDefault));
            SPB_CUDA(cudaHostAlloc((void**)&ncf, sizeof(float) * n, cudaHostAllocDefault));
            SPB_CUDA(cudaHostAlloc((void**)&nds, sizeof(float) * 128 * n, cudaHostAllocDefault));
            if (staged) {
                std::memcpy(nxy, sl.h_xy, sizeof(int) * 2 * staged);
                std::memcpy(ncf, sl.h_conf, sizeof(float) * staged);
                if (desc) std::memcpy(nds, sl.h_desc, row * staged);
            }
            if (sl.h_xy) cudaFreeHost(sl.h_xy);
            if (sl.h_conf) cudaFreeHost(sl.h_conf);
            if (sl.h_desc) cudaFreeHost(sl.h_desc);
            sl.h_xy = nxy; sl.h_conf = ncf; sl.h_desc = nds; sl.h_out_cap = n;
        }
        const uint8_t* d_desc = static_cast<const uint8_t*>(sl.d_desc);
        if (pin_out) {
            // pinned caller arrays: one strided copy per array and chunk (rows = images, width = the largest count of the
            // chunk) instead of three small copies per image; rows past an image's count receive don't-care values
            size_t nmax = 0;
            for (int b = 0; b < Bk; ++b) nmax = std::max(nmax, (size_t)count[g0 + b]);
            if (nmax) {
                SPB_CUDA(cudaMemcpy2DAsync(xy + g0 * cap * 2, sizeof(int) * 2 * cap, sl.d_xy + g0 * dcap * 2, sizeof(int) * 2 * dcap,
                                           sizeof(int) * 2 * nmax, Bk, cudaMemcpyDeviceToHost, s_out));
                SPB_CUDA(cudaMemcpy2DAsync(conf + g0 * cap, sizeof(float) * cap, sl.d_conf + g0 * dcap, sizeof(float) * dcap,
                                           sizeof(float) * nmax, Bk, cudaMemcpyDeviceToHost, s_out));
                if (desc) {
                    // The descriptors are 98 % of the bytes and the link is the bound of the whole path: consecutive images are
                    // grouped so that (images x largest count of the group) rows plus a fixed cost per copy (measured: ~6-8 us of
                    // the copy engine per strided copy = ~384 KB of link time) is minimal - one copy when the counts are even, one per image when they
                    // differ a lot (natural images), exact by dynamic programming over the <= 64 images of a chunk.
                    static const bool one_copy = [] { const char* e = std::getenv("SPB200_HOST_ONE_COPY"); return e && e[0] == '1'; }();
                    const size_t kCopyCost = one_copy ? ((size_t)1 << 40) : (size_t)384 * 1024;
                    std::vector<size_t> best((size_t)Bk + 1, 0);
                    std::vector<int> from((size_t)Bk + 1, 0);
                    for (int j = 1; j <= Bk; ++j) {
                        best[j] = ~(size_t)0;
                        size_t gmax = 0;
                        for (int i = j - 1; i >= 0; --i) {
                            gmax = std::max(gmax, (size_t)count[g0 + i]);
                            const size_t c = best[i] + (size_t)(j - i) * gmax * row + kCopyCost;
                            if (c < best[j]) { best[j] = c; from[j] = i; }
                        }
                    }
                    for (int j = Bk; j > 0; j = from[j]) {
                        const int i = from[j];
                        size_t gmax = 0;
                        for (int b = i; b < j; ++b) gmax = std::max(gmax, (size_t)count[g0 + b]);
                        if (gmax)
                            SPB_CUDA(cudaMemcpy2DAsync(desc + (g0 + i) * cap * row, row * cap, d_desc + (g0 + i) * dcap * row, row * dcap,
                                                       row * gmax, j - i, cudaMemcpyDeviceToHost, s_out));
                    }
                }
            }
            continue;
        }
        for (int b = 0; b < Bk; ++b) {
            const size_t g = g0 + b, n = (size_t)count[g];
            stage_off[g] = staged;
            if (n) {
                SPB_CUDA(cudaMemcpyAsync(sl.h_xy + staged * 2, sl.d_xy + g * dcap * 2, sizeof(int) * 2 * n, cudaMemcpyDeviceToHost, s_out));
                SPB_CUDA(cudaMemcpyAsync(sl.h_conf + staged, sl.d_conf + g * dcap, sizeof(float) * n, cudaMemcpyDeviceToHost, s_out));
                if (desc) SPB_CUDA(cudaMemcpyAsync(sl.h_desc + staged * row, d_desc + g * dcap * row, row * n, cudaMemcpyDeviceToHost, s_out));
            }
            staged += n;
        }
    }
    SPB_CUDA(cudaStreamSynchronize(s_out));
    if (!pin_out) {
        for (size_t g = 0; g < (size_t)B; ++g) {
            const size_t n = (size_t)count[g], off = stage_off[g];
            std::memcpy(xy + g * cap * 2, sl.h_xy + off * 2, sizeof(int) * 2 * n);
            std::memcpy(conf + g * cap, sl.h_conf + off, sizeof(float) * n);
            if (desc) std::memcpy(desc + g * cap * row, sl.h_desc + off * row, row * n);
        }
    }
}

}  // namespace spb200
