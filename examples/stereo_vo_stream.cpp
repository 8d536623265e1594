// examples/stereo_vo_stream.cpp -- what the reference's apps do with a StereoDataset (utils/dataset.cc:103-134: load the
// pair, StereoAlgorithm::run, hand image + disparity to VisualOdometry::addFrame), written against the C++ host shim:
// bpvo_b200::StereoAlgorithm (host/stereo.h) + bpvo_b200::VisualOdometry (host/vo.h).
//
//   g++ -std=c++14 -O2 examples/stereo_vo_stream.cpp -Ibpvo_b200/csrc/host -Lbpvo_b200 -lbpvo_b200 -Wl,-rpath,$PWD/bpvo_b200 -o stereo_vo_stream
//   ./stereo_vo_stream pairs.bin
//
// pairs.bin (little endian): int32 rows, cols, nframes, numberOfDisparities, SADWindowSize, numPyramidLevels; float K[9]
// (column major), baseline; then per frame rows*cols u8 left + rows*cols u8 right.
// Output: one line per frame -- isKeyFrame, numFunEvals, a checksum of the disparity map, the 16 pose floats as %a (exact).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "stereo.h"

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s pairs.bin\n", argv[0]); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::perror(argv[1]); return 2; }
  int32_t hdr[6];
  bpvo_b200::Matrix33 K; float baseline = 0;
  if (std::fread(hdr, 4, 6, f) != 6 || std::fread(K.data(), 4, 9, f) != 9 || std::fread(&baseline, 4, 1, f) != 1) return 2;
  const int rows = hdr[0], cols = hdr[1], nframes = hdr[2];
  try {
    bpvo_b200::StereoParameters sp;
    bpvo_b200_stereo_default_params(&sp);
    sp.numberOfDisparities = hdr[3]; sp.SADWindowSize = hdr[4];
    bpvo_b200::StereoAlgorithm stereo(bpvo_b200::ImageSize(rows, cols), sp);
    bpvo_b200::AlgorithmParameters params;
    bpvo_b200_default_params(&params);
    params.descriptor = BPVO_B200_BITPLANES; params.numPyramidLevels = hdr[5]; params.minValidDisparity = 1.0f;
    bpvo_b200::VisualOdometry vo(K, baseline, bpvo_b200::ImageSize(rows, cols), params);
    std::vector<uint8_t> left((size_t) rows * cols), right((size_t) rows * cols);
    std::vector<float> dmap((size_t) rows * cols);
    for (int k = 0; k < nframes; ++k) {
      if (std::fread(left.data(), 1, left.size(), f) != left.size() || std::fread(right.data(), 1, right.size(), f) != right.size()) return 2;
      stereo.run(left.data(), right.data(), dmap.data());
      bpvo_b200::Result r = vo.addFrame(left.data(), dmap.data());
      double sum = 0; size_t invalid = 0;
      for (float d : dmap) { if (d == stereo.filteredValue()) ++invalid; else sum += d; }
      std::printf("%d %d %.17g %zu", (int) r.isKeyFrame, r.numFunEvals, sum, invalid);
      for (int i = 0; i < 16; ++i) std::printf(" %a", r.pose.data()[i]);
      std::printf("\n");
    }
  } catch (const bpvo_b200::Error& e) {
    std::fprintf(stderr, "bpvo_b200::Error: %s\n", e.what());
    return 1;
  }
  std::fclose(f);
  return 0;
}
