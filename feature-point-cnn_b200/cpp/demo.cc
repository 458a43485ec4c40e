// Headless counterpart of the reference's cpp/src/main.cc loop (no camera, no GUI): loads a checkpoint,
// runs ProcessFrame on a synthetic frame and prints the strongest keypoints.
//   g++ -std=c++17 -I include -I feature-point-cnn_b200/cpp demo.cc -L feature-point-cnn_b200 -lspb200 -o demo
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "superpoint.h"

int main(int argc, char** argv) {
  if (argc < 2) {
    std::fprintf(stderr, "usage: %s super_point.pt [height width]\n", argv[0]);
    return 2;
  }
  const int h = argc > 3 ? std::atoi(argv[2]) : 240, w = argc > 3 ? std::atoi(argv[3]) : 320;
  try {
    superpoint::SuperPoint net(argv[1], false);
    std::vector<float> frame((size_t)h * w);
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) frame[(size_t)y * w + x] = (((x / 20) + (y / 20)) & 1) ? 0.8f : 0.2f;   // checkerboard
    auto pts = net.ProcessFrame(frame.data(), h, w);
    std::printf("%zu keypoints, descriptor dim %d\n", pts.size(), net.descriptor_dim());
    for (size_t i = 0; i < pts.size() && i < 5; ++i)
      std::printf("  (%d, %d) conf %.4f desc[0..2] %.4f %.4f %.4f\n", pts[i].x, pts[i].y, pts[i].confidence,
                  pts[i].descriptor[0], pts[i].descriptor[1], pts[i].descriptor[2]);
    // the same frame as the camera delivers it (8-bit): identical keypoints
    std::vector<uint8_t> frame8((size_t)h * w);
    for (size_t i = 0; i < frame8.size(); ++i) frame8[i] = frame[i] > 0.5f ? 204 : 51;        // 0.8 * 255, 0.2 * 255
    for (size_t i = 0; i < frame.size(); ++i) frame[i] = frame8[i] / 255.f;
    auto ref = net.ProcessFrame(frame.data(), h, w);
    auto pts8 = net.ProcessFrame8(frame8.data(), h, w);
    bool same = ref.size() == pts8.size();
    for (size_t i = 0; same && i < ref.size(); ++i) same = ref[i].x == pts8[i].x && ref[i].y == pts8[i].y && ref[i].confidence == pts8[i].confidence;
    std::printf("8-bit frame: %zu keypoints, %s\n", pts8.size(), same ? "identical to the float frame" : "DIFFERENT from the float frame");
    if (!same) return 3;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
