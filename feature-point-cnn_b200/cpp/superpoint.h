// C++ drop-in for the reference's inference wrapper, over the C ABI of libspb200.so.
//
// Same names and shapes as the reference's C++ boundary:
//   superpoint::SuperPoint(file_name, load_script) / ProcessFrame      cpp/src/superpoint.h:12-36, superpoint.cc:9-96
//   superpoint::FeaturePoint {x, y, confidence, descriptor[256]}       cpp/src/torchutis.h:11-18
//   superpoint::Settings (inference knobs)                             cpp/src/settings.h:27-31
// with the semantics of the current Python model (python/src/inferencewrapper.py:29-46): the reference's
// C++ demo is stale (VGG-style network, NMS result discarded), so the interface is kept and
// the behaviour follows Python.  The descriptor of these checkpoints is 128-d: the first
// descriptor_dim() entries of FeaturePoint::descriptor are filled, the rest are zero.
//
// ProcessFrame takes the frame as a float pointer (CV_32FC1 data, [0,1]); the cv::Mat overload is
// compiled when OpenCV's core header has been included before this file.  Header-only; link with
// -lspb200.  Like the reference class, an instance is not re-entrant (one per thread / GPU).
#ifndef SPB200_SUPERPOINT_H
#define SPB200_SUPERPOINT_H

#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "spb200.h"

namespace superpoint {

using DescriptorType = std::array<float, 256>;

struct FeaturePoint {
  int x = 0;
  int y = 0;
  float confidence = 0;
  DescriptorType descriptor;
};

struct Settings {                       // cpp/src/settings.h:27-31 == python/src/settings.py:3-8
  int nms_dist = 4;
  float confidence_thresh = 0.015f;
  float nn_thresh = 0.7f;               // matcher threshold, unused by the wrapper
  int cell = 8;
  int border_remove = 4;
  int top_k = 0;                        // 0 = every survivor (the reference has no top-k)
  int device = 0;
  int precision = SPB200_PREC_FP16;
};

class SuperPoint {
 public:
  // file_name: snapshots/super_point.pt or magic_point.pt (the torch.save checkpoint of
  // python/src/saveutils.py:54-63, a bare state_dict, or the <name>_params.pt of InferenceWrapper.trace).
  // load_script = true is the reference's "<name>_script.pt" mode (cpp/src/superpoint.cc:11-26): this engine executes no
  // TorchScript, so the weights are taken from the "<name>_params.pt" that the same trace() call wrote next to the script;
  // any other script path fails with a message that says so.
  explicit SuperPoint(const std::string& file_name, bool load_script = false, const Settings& settings = Settings())
      : settings_(settings) {
    std::string weights = file_name;
    if (load_script) {
      const std::string tail = "_script.pt";
      if (weights.size() <= tail.size() || weights.compare(weights.size() - tail.size(), tail.size(), tail) != 0)
        throw std::invalid_argument("load_script: expected a '<name>_script.pt' path; TorchScript modules are not executed, the "
                                    "weights are read from '<name>_params.pt' (InferenceWrapper.trace writes both)");
      weights = weights.substr(0, weights.size() - tail.size()) + "_params.pt";
    }
    if (spb200_create(settings_.device, &engine_) != SPB200_OK) throw std::runtime_error(spb200_last_error(nullptr));
    try {
      Check(spb200_load_checkpoint(engine_, weights.c_str()));
      Check(spb200_finalize_weights(engine_, settings_.precision));
      Check(spb200_set_params(engine_, settings_.confidence_thresh, settings_.nms_dist, settings_.border_remove,
                              settings_.top_k, 1));
    } catch (...) {
      spb200_destroy(engine_);
      throw;
    }
  }
  ~SuperPoint() { spb200_destroy(engine_); }
  SuperPoint(const SuperPoint&) = delete;
  SuperPoint& operator=(const SuperPoint&) = delete;

  int descriptor_dim() const { return spb200_descriptor_dim(engine_); }

  // frame: rows*cols floats in [0,1] (cv::Mat CV_32FC1, continuous); rows and cols multiples of 16.
  std::vector<FeaturePoint> ProcessFrame(const float* frame, int rows, int cols) { return Process(frame, nullptr, rows, cols); }

  // frame: rows*cols bytes, the 8-bit grayscale camera frame before the reference's convertTo(CV_32FC1, 1/255)
  // (cpp/src/camera.cc:12-23); same result as ProcessFrame on frame / 255.f, a quarter of the upload.
  std::vector<FeaturePoint> ProcessFrame8(const uint8_t* frame, int rows, int cols) { return Process(nullptr, frame, rows, cols); }

 private:
  std::vector<FeaturePoint> Process(const float* frame, const uint8_t* frame8, int rows, int cols) {
    const int cap = settings_.top_k > 0 ? settings_.top_k : spb200_max_keypoints(rows, cols, settings_.nms_dist);
    const int dim = descriptor_dim();
    xy_.resize((size_t)cap * 2);
    conf_.resize(cap);
    desc_.resize((size_t)cap * dim);
    int count = 0;
    if (frame8) Check(spb200_detect_host_u8(engine_, frame8, 1, rows, cols, cap, &count, xy_.data(), conf_.data(), desc_.data()));
    else Check(spb200_detect_host(engine_, frame, 1, 1, rows, cols, cap, &count, xy_.data(), conf_.data(), desc_.data()));
    feature_points_.resize(count);
    for (int i = 0; i < count; ++i) {
      FeaturePoint& fp = feature_points_[i];
      fp.x = xy_[2 * i];
      fp.y = xy_[2 * i + 1];
      fp.confidence = conf_[i];
      fp.descriptor.fill(0.f);
      std::memcpy(fp.descriptor.data(), desc_.data() + (size_t)i * dim, sizeof(float) * dim);
    }
    return feature_points_;     // a copy, like the reference (superpoint.cc:95)
  }

 public:
#ifdef OPENCV_CORE_HPP
  std::vector<FeaturePoint> ProcessFrame(const cv::Mat& frame) {
    cv::Mat c = frame.isContinuous() ? frame : frame.clone();
    if (frame.type() == CV_8UC1) return ProcessFrame8(c.ptr<uint8_t>(), c.rows, c.cols);
    if (frame.type() != CV_32FC1) throw std::invalid_argument("ProcessFrame expects CV_32FC1 (torchutis.cc:6) or CV_8UC1");
    return ProcessFrame(c.ptr<float>(), c.rows, c.cols);
  }
#endif

 private:
  void Check(int rc) {
    if (rc != SPB200_OK) throw std::runtime_error(spb200_last_error(engine_));
  }

  Settings settings_;
  spb200_engine* engine_ = nullptr;
  // memory management buffers (the reference keeps these as members too, superpoint.h:31-35)
  std::vector<int> xy_;
  std::vector<float> conf_;
  std::vector<float> desc_;
  std::vector<FeaturePoint> feature_points_;
};

}  // namespace superpoint

#endif  // SPB200_SUPERPOINT_H
