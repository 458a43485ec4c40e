// Headless counterpart of the reference's cpp/src/main.cc loop (no camera, no GUI): loads a checkpoint,
// runs ProcessFrame on a synthetic frame and prints the strongest keypoints.
//   g++ -std=c++17 -I include -I feature-point-cnn_b200/cpp demo.cc -L feature-point-cnn_b200 -lspb200 -o demo
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
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
