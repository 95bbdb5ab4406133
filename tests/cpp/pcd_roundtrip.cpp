// Host-only check of the facade's .pcd I/O: load argv[1] (ascii or binary), write it as binary to argv[2] and print the
// point count, dimensions and a checksum of the raw point words.  (tests/test_pcd_io.py drives it.)
#include "../../realsense-pointcloud_b200/host/rspcl.hpp"

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  try {
    rgb_point_cloud c;
    rspcl::io::loadPCDFile(argv[1], c);
    rspcl::io::savePCDFileBinary(argv[2], c);
    unsigned long long sum = 0;
    for (const auto& p : c.points) {
      std::uint32_t w[4];
      std::memcpy(&w[0], &p.x, 4);
      std::memcpy(&w[1], &p.y, 4);
      std::memcpy(&w[2], &p.z, 4);
      w[3] = p.rgba;
      for (int k = 0; k < 4; ++k) sum = sum * 1000003ull + w[k];
    }
    std::printf("%zu %u %u %llu\n", c.size(), c.width, c.height, sum);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
