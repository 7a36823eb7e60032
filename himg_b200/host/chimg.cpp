// chimg -- encode driver (mirrors src/chimg.cpp:36-169: same options and messages).
//   chimg [-q <0..100>] [-rgb] image outfile
// `image` is a binary PGM/PPM/PAM file or "synthetic:WxHxC[:seed[:amp]]" (FreeImage is not used).
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "encoder.h"
#include "pnm.h"

namespace {

const int kDefaultQuality = 50;

void Usage(const char *arg0) {
  std::cout << "Usage: " << arg0 << " [options] image outfile\n";
  std::cout << "Options:\n";
  std::cout << " -q <quality> Set the quality (0-100)\n";
  std::cout << " -rgb         Use RGB color space (instead of YCbCr)\n";
}

}  // namespace

int main(int argc, const char **argv) {
  bool use_ycbcr = true;
  int quality = kDefaultQuality;
  std::vector<std::string> files;
  bool ok = true;
  for (int k = 1; k < argc && ok; ++k) {
    const std::string arg = argv[k];
    if (!arg.empty() && arg[0] == '-' && arg.compare(0, 10, "synthetic:") != 0) {
      if (arg == "-rgb") {
        use_ycbcr = false;
      } else if (arg == "-q") {
        char *end = nullptr;
        if (k + 1 < argc) {
          quality = static_cast<int>(std::strtol(argv[++k], &end, 10));
          if (end == argv[k] || *end != 0) {
            std::cout << "Invalid integer expression: " << argv[k] << "\n";
            ok = false;
          } else if (quality < 0 || quality > 100) {
            std::cout << "Invalid quality level: " << quality << "\n";
            ok = false;
          }
        } else {
          ok = false;
        }
      } else {
        std::cout << "Invalid option: " << arg << "\n";
        ok = false;
      }
    } else {
      files.push_back(arg);
    }
  }
  if (!ok || files.size() != 2) {
    Usage(argv[0]);
    return 0;
  }

  himg::host::Image img;
  if (!himg::host::LoadImage(files[0], &img)) {
    std::cerr << "Unable to load " << files[0] << std::endl;
    return -1;
  }

  himg::Encoder encoder;
  if (!encoder.Encode(img.pixels.data(), img.width, img.height, img.channels, img.channels, quality, use_ycbcr)) {
    std::cerr << "Unable to encode " << files[0] << std::endl;
    return -1;
  }
  std::cout << "Compressed size: " << encoder.packed_size() << std::endl;

  std::ofstream f(files[1].c_str(), std::ofstream::out | std::ofstream::binary);
  f.write(reinterpret_cast<const char *>(encoder.packed_data()), encoder.packed_size());
  return f.good() ? 0 : -1;
}
