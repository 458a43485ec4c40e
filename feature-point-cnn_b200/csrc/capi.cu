// C ABI of the spb200 engine (include/spb200.h): exception firewall around spb200::Engine.
#include <algorithm>
#include <cstdio>
#include <new>
#include <string>

#include "../../include/spb200.h"
#include "engine.h"

struct spb200_engine {
    spb200::Engine impl;
    explicit spb200_engine(int device) : impl(device) {}
};

namespace {
thread_local std::string g_create_error;

template <typename F>
int guarded(spb200_engine* e, F&& f) {
    if (!e) return SPB200_E_INVALID;
    try {
        f(e->impl);
        return SPB200_OK;
    } catch (const std::invalid_argument& ex) {
        e->impl.last_error = ex.what();
        return SPB200_E_INVALID;
    } catch (const std::bad_alloc&) {
        e->impl.last_error = "out of host memory";
        return SPB200_E_NOMEM;
    } catch (const std::exception& ex) {
        e->impl.last_error = ex.what();
        return SPB200_E_RUNTIME;
    } catch (...) {
        e->impl.last_error = "unknown error";
        return SPB200_E_RUNTIME;
    }
}
}  // namespace

extern "C" {

int spb200_create(int device, spb200_engine** out) {
    if (!out) return SPB200_E_INVALID;
    *out = nullptr;
    try {
        *out = new spb200_engine(device);
        return SPB200_OK;
    } catch (const std::invalid_argument& ex) {
        g_create_error = ex.what();
        return SPB200_E_INVALID;
    } catch (const std::exception& ex) {
        g_create_error = ex.what();
        return SPB200_E_RUNTIME;
    } catch (...) {
        g_create_error = "unknown error";
        return SPB200_E_RUNTIME;
    }
}

void spb200_destroy(spb200_engine* e) { delete e; }

const char* spb200_last_error(const spb200_engine* e) { return e ? e->impl.last_error.c_str() : g_create_error.c_str(); }

int spb200_load_checkpoint(spb200_engine* e, const char* path) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!path) throw std::invalid_argument("path is null");
        g.load_checkpoint(path);
    });
}

int spb200_load_tensor(spb200_engine* e, const char* key, const float* host_data, const int64_t* shape, int rank) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!key || !host_data || rank < 0 || (rank > 0 && !shape)) throw std::invalid_argument("bad tensor arguments");
        g.load_tensor(key, host_data, shape, rank);
    });
}

int spb200_finalize_weights(spb200_engine* e, int precision) {
    return guarded(e, [&](spb200::Engine& g) { g.finalize(precision); });
}

int spb200_finalize_weights_split(spb200_engine* e, int precision, int split_level) {
    return guarded(e, [&](spb200::Engine& g) { g.finalize(precision, split_level); });
}

int spb200_set_params(spb200_engine* e, float conf_thresh, int nms_dist, int border_remove, int top_k, int descriptor_enabled) {
    return guarded(e, [&](spb200::Engine& g) {
        if (nms_dist < 0 || nms_dist > 8) throw std::invalid_argument("nms_dist must be in [0, 8]");
        if (border_remove < 0 || top_k < 0) throw std::invalid_argument("border_remove and top_k must be >= 0");
        spb200::Params p;
        p.conf_thresh = conf_thresh; p.nms_dist = nms_dist; p.border_remove = border_remove; p.top_k = top_k;
        p.descriptor_enabled = descriptor_enabled ? 1 : 0;
        g.set_params(p);
    });
}

int spb200_forward(spb200_engine* e, const float* img, int B, int C, int H, int W, float* prob_map, float* desc_map,
                   float* logits, void* stream) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!img || !prob_map) throw std::invalid_argument("img and prob_map must not be null");
        g.forward(img, B, C, H, W, prob_map, desc_map, logits, (cudaStream_t)stream);
    });
}

int spb200_detect(spb200_engine* e, const float* img, int B, int C, int H, int W, int capacity, int* count, int* xy,
                  float* conf, float* desc, float* prob_map, void* stream) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!img || !count || !xy || !conf) throw std::invalid_argument("img, count, xy and conf must not be null");
        g.detect(img, B, C, H, W, capacity, count, xy, conf, desc, prob_map, (cudaStream_t)stream);
    });
}

int spb200_detect_host(spb200_engine* e, const float* img_host, int B, int C, int H, int W, int capacity, int* count_host,
                       int* xy_host, float* conf_host, float* desc_host) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!img_host || !count_host || !xy_host || !conf_host) throw std::invalid_argument("img, count, xy and conf must not be null");
        if (capacity <= 0) throw std::invalid_argument("capacity must be positive");
        g.detect_host(img_host, B, C, H, W, capacity, count_host, xy_host, conf_host, desc_host);
    });
}

int spb200_detect_u8(spb200_engine* e, const uint8_t* img, int B, int H, int W, int capacity, int* count, int* xy, float* conf,
                     float* desc, float* prob_map, void* stream) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!img || !count || !xy || !conf) throw std::invalid_argument("img, count, xy and conf must not be null");
        g.detect_u8(img, B, H, W, capacity, count, xy, conf, desc, prob_map, (cudaStream_t)stream);
    });
}

int spb200_detect_host_u8(spb200_engine* e, const uint8_t* img_host, int B, int H, int W, int capacity, int* count_host,
                          int* xy_host, float* conf_host, float* desc_host) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!img_host || !count_host || !xy_host || !conf_host) throw std::invalid_argument("img, count, xy and conf must not be null");
        if (capacity <= 0) throw std::invalid_argument("capacity must be positive");
        g.detect_host_u8(img_host, B, H, W, capacity, count_host, xy_host, conf_host, desc_host);
    });
}

int spb200_detect_host_submit(spb200_engine* e, const void* img_host, int img_is_u8, int B, int C, int H, int W, int capacity,
                              int* count_host, int* xy_host, float* conf_host, void* desc_host, int* ticket) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!img_host || !ticket || !count_host || !xy_host || !conf_host) throw std::invalid_argument("img, count, xy, conf and ticket must not be null");
        *ticket = g.detect_host_submit(img_host, img_is_u8 != 0, B, C, H, W, capacity, count_host, xy_host, conf_host, desc_host);
    });
}

int spb200_detect_host_wait(spb200_engine* e, int ticket) {
    return guarded(e, [&](spb200::Engine& g) { g.detect_host_wait(ticket); });
}

int spb200_set_descriptor_format(spb200_engine* e, int format) {
    return guarded(e, [&](spb200::Engine& g) { g.set_descriptor_format(format); });
}

int spb200_homography_adaptation(spb200_engine* e, const float* img, int B, int C, int H, int W, const float* homographies_host,
                                 int num, int valid_border_margin, int aggregation, float* prob_map, void* stream) {
    return guarded(e, [&](spb200::Engine& g) {
        g.homography_adaptation(img, B, C, H, W, homographies_host, num, valid_border_margin, aggregation, prob_map,
                                (cudaStream_t)stream);
    });
}

int spb200_match(spb200_engine* e, const float* desc_a, const int* count_a, const float* desc_b, const int* count_b, int B,
                 int capacity, int D, float max_dist, int* match_ab, float* dist, void* stream) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!desc_a || !count_a || !desc_b || !count_b || !match_ab || !dist) throw std::invalid_argument("match: null argument");
        g.match(desc_a, count_a, desc_b, count_b, B, capacity, D, max_dist, match_ab, dist, (cudaStream_t)stream);
    });
}

int spb200_heatmap_from_logits(spb200_engine* e, const float* logits, int B, int H, int W, float* prob_map, void* stream) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!logits || !prob_map || B <= 0 || H <= 0 || W <= 0) throw std::invalid_argument("bad heatmap arguments");
        g.heatmap_from_logits(logits, B, H, W, prob_map, (cudaStream_t)stream);
    });
}

int spb200_restore_prob_map(spb200_engine* e, const float* softmax, int B, int H, int W, float* prob_map, void* stream) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!softmax || !prob_map || B <= 0 || H <= 0 || W <= 0) throw std::invalid_argument("bad restore_prob_map arguments");
        g.restore_prob_map(softmax, B, H, W, prob_map, (cudaStream_t)stream);
    });
}

int spb200_preprocess_u8(spb200_engine* e, const uint8_t* frames, int B, int h, int w, int C, uint8_t* gray, int H, int W, void* stream) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!frames || !gray) throw std::invalid_argument("preprocess_u8: null argument");
        g.preprocess_u8(frames, B, h, w, C, gray, H, W, (cudaStream_t)stream);
    });
}

int spb200_preprocess_f32(spb200_engine* e, const float* frames, int B, int h, int w, float* rgb, int H, int W, void* stream) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!frames || !rgb) throw std::invalid_argument("preprocess_f32: null argument");
        g.preprocess_f32(frames, B, h, w, rgb, H, W, (cudaStream_t)stream);
    });
}

int spb200_nms(spb200_engine* e, const float* prob_map, int B, int H, int W, int capacity, int* count, int* xy, float* conf,
               void* stream) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!prob_map || !count || !xy || !conf || B <= 0 || H <= 0 || W <= 0) throw std::invalid_argument("bad nms arguments");
        g.nms(prob_map, B, H, W, capacity, count, xy, conf, (cudaStream_t)stream);
    });
}

int spb200_sample_descriptors(spb200_engine* e, const float* desc_map, int B, int D, int H, int W, int capacity,
                              const int* count, const int* xy, float* desc, void* stream) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!desc_map || !count || !xy || !desc || B <= 0 || H < 8 || W < 8) throw std::invalid_argument("bad descriptor arguments");
        if (H % 8 || W % 8) throw std::invalid_argument("H and W must be multiples of the 8-pixel cell");
        if (D <= 0 || D % 4) throw std::invalid_argument("descriptor dimension must be a positive multiple of 4");
        if (capacity <= 0) throw std::invalid_argument("capacity must be positive");
        g.sample_descriptors(desc_map, B, D, H, W, capacity, count, xy, desc, (cudaStream_t)stream);
    });
}

int spb200_descriptor_dim(const spb200_engine* e) { return e ? e->impl.descriptor_dim() : 128; }

int spb200_max_keypoints(int H, int W, int nms_dist) { return spb200::Engine::max_keypoints(H, W, nms_dist); }

long spb200_kernel_launches(const spb200_engine* e) { return e ? e->impl.launches() : 0; }

void spb200_reset_kernel_launches(spb200_engine* e) { if (e) e->impl.reset_launches(); }

int spb200_export_activation(spb200_engine* e, int buffer_id, float* dst, int channels, void* stream) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!dst) throw std::invalid_argument("dst is null");
        g.export_buffer(buffer_id, dst, channels, (cudaStream_t)stream);
    });
}

int spb200_activation_dims(const spb200_engine* e, int buffer_id, int* channels, int* height, int* width) {
    if (!e || !channels || !height || !width) return SPB200_E_INVALID;
    try {
        e->impl.buffer_dims(buffer_id, channels, height, width);
        return SPB200_OK;
    } catch (...) {
        return SPB200_E_INVALID;
    }
}

int spb200_profile_begin(spb200_engine* e) {
    return guarded(e, [&](spb200::Engine& g) { g.profile_begin(); });
}

int spb200_profile_end(spb200_engine* e, int max_entries, char* names, float* ms, double* flops, double* bytes, int* n) {
    return guarded(e, [&](spb200::Engine& g) {
        if (!names || !ms || !flops || !bytes || !n || max_entries < 0) throw std::invalid_argument("bad profile arguments");
        const auto& v = g.profile_end();
        *n = (int)std::min<size_t>(v.size(), (size_t)max_entries);
        for (int i = 0; i < *n; ++i) {
            std::snprintf(names + (size_t)i * 64, 64, "%s", v[i].name.c_str());
            ms[i] = v[i].ms; flops[i] = v[i].flops; bytes[i] = v[i].bytes;
        }
    });
}

int spb200_checkpoint_num_tensors(const char* path) {
    if (!path) return -1;
    spb200::StateDict sd;
    std::string err;
    if (!spb200::read_checkpoint(path, sd, err)) { g_create_error = err; return -1; }
    return (int)sd.size();
}

int spb200_checkpoint_tensor(const char* path, const char* key, float* dst, long capacity, int64_t* shape8, int* rank) {
    if (!path || !key) return SPB200_E_INVALID;
    spb200::StateDict sd;
    std::string err;
    if (!spb200::read_checkpoint(path, sd, err)) { g_create_error = err; return SPB200_E_RUNTIME; }
    auto it = sd.find(key);
    if (it == sd.end()) {                               // *_params.pt: keys without their module prefix
        spb200::restore_module_prefixes(sd);
        it = sd.find(key);
    }
    if (it == sd.end()) { g_create_error = std::string("no tensor named ") + key; return SPB200_E_INVALID; }
    const spb200::HostTensor& t = it->second;
    if (t.shape.size() > 8) return SPB200_E_INVALID;
    if (rank) *rank = (int)t.shape.size();
    if (shape8) for (size_t i = 0; i < t.shape.size(); ++i) shape8[i] = t.shape[i];
    if (dst) {
        if (capacity < t.numel()) return SPB200_E_INVALID;
        std::copy(t.data.begin(), t.data.end(), dst);
    }
    return SPB200_OK;
}

static int test_conv_impl(int kernel, int precision, const void* x, const void* w, const float* bias, void* y, int B, int H, int W,
                          int cin, int cout, int taps, int stride, int relu, int out_fp32, void* stream) {
    try {
        if (precision != SPB200_PREC_FP16 && precision != SPB200_PREC_BF16) return SPB200_E_INVALID;
        if ((taps != 1 && taps != 9) || (stride != 1 && stride != 2) || cin % 64 || cout % 16 || H % stride || W % stride)
            return SPB200_E_INVALID;
        spb200::ConvDev d{};
        d.nseg = 1;
        spb200::SegDev& s = d.seg[0];
        s.src = x; s.H = H; s.W = W; s.C = cin; s.cin = cin; s.cin_real = cin; s.ntaps = taps; s.stride = stride; s.koff = 0;
        for (int t = 0; t < taps; ++t) {
            s.dy[t] = (int8_t)(taps == 9 ? t / 3 - 1 : 0);
            s.dx[t] = (int8_t)(taps == 9 ? t % 3 - 1 : 0);
        }
        d.w = w; d.bias = bias; d.residual = nullptr; d.dst = y;
        d.B = B; d.OH = H / stride; d.OW = W / stride; d.K = taps * cin; d.cout_pad = cout;
        d.dst_H = d.OH; d.dst_W = d.OW; d.dst_C = cout; d.dst_stride = 1; d.dst_off_y = 0; d.dst_off_x = 0;
        d.res_C = 0; d.relu = relu; d.dst_fp32 = out_fp32;
        cudaError_t err = cudaSuccess;
        if (kernel == 0) {
            spb200::TcConvPlan* plan = spb200::tc_plan_create(d, precision);
            spb200::launch_conv_tc(plan, (cudaStream_t)stream);
            err = cudaStreamSynchronize((cudaStream_t)stream);
            spb200::tc_plan_destroy(plan);
        } else if (kernel == 1) {
            spb200::TcBlockPlan* plan = spb200::tc_block_plan_create(d, nullptr, precision, cout, 148);
            spb200::launch_block_tc(plan, (cudaStream_t)stream);
            err = cudaStreamSynchronize((cudaStream_t)stream);
            spb200::tc_block_plan_destroy(plan);
        } else if (kernel == 2) {
            spb200::TcHaloPlan* plan = spb200::tc_halo_plan_create(d, nullptr, precision, cout, 148);
            if (!plan) { g_create_error = "the haloed-tile kernel does not take this convolution"; return SPB200_E_INVALID; }
            spb200::launch_halo_tc(plan, (cudaStream_t)stream);
            err = cudaStreamSynchronize((cudaStream_t)stream);
            spb200::tc_halo_plan_destroy(plan);
        } else {
            return SPB200_E_INVALID;
        }
        if (err != cudaSuccess) { g_create_error = cudaGetErrorString(err); return SPB200_E_RUNTIME; }
        return SPB200_OK;
    } catch (const std::exception& ex) {
        g_create_error = ex.what();
        return SPB200_E_RUNTIME;
    }
}

int spb200_test_conv_tc(int precision, const void* x, const void* w, const float* bias, void* y, int B, int H, int W,
                        int cin, int cout, int taps, int stride, int relu, int out_fp32, void* stream) {
    return test_conv_impl(0, precision, x, w, bias, y, B, H, W, cin, cout, taps, stride, relu, out_fp32, stream);
}

int spb200_test_conv_kernel(int kernel, int precision, const void* x, const void* w, const float* bias, void* y, int B, int H,
                            int W, int cin, int cout, int taps, int stride, int relu, int out_fp32, void* stream) {
    return test_conv_impl(kernel, precision, x, w, bias, y, B, H, W, cin, cout, taps, stride, relu, out_fp32, stream);
}

}  // extern "C"
