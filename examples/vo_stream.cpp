// examples/vo_stream.cpp -- the reference's apps/vo.cc main loop (apps/vo.cc:57-106) written against the C++ host
// shim (bpvo_b200::VisualOdometry, same interface as bpvo/vo.h:33-107) instead of bpvo::VisualOdometry.
//
//   g++ -std=c++14 -O2 examples/vo_stream.cpp -Ibpvo_b200/csrc/host -Lbpvo_b200 -lbpvo_b200 -Wl,-rpath,$PWD/bpvo_b200 -o vo_stream
//   ./vo_stream frames.bin [poses.txt]
//
// frames.bin (little endian):  int32 rows, cols, nframes, descriptor, numPyramidLevels, lossFunction;  float K[9] (column
// major), baseline;  then per frame rows*cols u8 grey + rows*cols f32 disparity.  (tests/test_gpu_parity.py writes it.)
// Output: one line per frame -- isKeyFrame, keyFramingReason, numFunEvals and the 16 pose floats as %a (exact).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "vo.h"

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s frames.bin [poses.txt]\n", argv[0]); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::perror(argv[1]); return 2; }
  int32_t hdr[6];
  bpvo_b200::Matrix33 K; float baseline = 0;
  if (std::fread(hdr, 4, 6, f) != 6 || std::fread(K.data(), 4, 9, f) != 9 || std::fread(&baseline, 4, 1, f) != 1) return 2;
  const int rows = hdr[0], cols = hdr[1], nframes = hdr[2];

  bpvo_b200::AlgorithmParameters params;
  bpvo_b200_default_params(&params);                       // AlgorithmParameters() defaults (bpvo/types.cc:31-66)
  params.descriptor = hdr[3];
  params.numPyramidLevels = hdr[4];
  params.lossFunction = hdr[5];

  FILE* out = argc > 2 ? std::fopen(argv[2], "w") : stdout;
  try {
    bpvo_b200::VisualOdometry vo(K, baseline, bpvo_b200::ImageSize(rows, cols), params);
    std::vector<uint8_t> image((size_t) rows * cols);
    std::vector<float> disparity((size_t) rows * cols);
    for (int k = 0; k < nframes; ++k) {
      if (std::fread(image.data(), 1, image.size(), f) != image.size()) return 2;
      if (std::fread(disparity.data(), 4, disparity.size(), f) != disparity.size()) return 2;
      bpvo_b200::Result r = vo.addFrame(image.data(), disparity.data());
      std::fprintf(out, "%d %d %d", (int) r.isKeyFrame, r.keyFramingReason, r.numFunEvals);
      for (int i = 0; i < 16; ++i) std::fprintf(out, " %a", r.pose.data()[i]);
      std::fprintf(out, " %zu\n", r.pointCloud ? r.pointCloud->points.size() : (size_t) 0);
    }
    std::fprintf(out, "trajectory %zu points_at_level %d\n", vo.trajectory().size(), vo.numPointsAtLevel());
  } catch (const bpvo_b200::Error& e) {
    std::fprintf(stderr, "bpvo_b200::Error: %s\n", e.what());
    return 1;
  }
  if (out != stdout) std::fclose(out);
  std::fclose(f);
  return 0;
}
