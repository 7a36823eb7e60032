// dhimg -- decode driver (mirrors src/dhimg.cpp:17-72).   dhimg image outfile
// Writes binary PGM (1 channel), PPM (3) or PAM (2/4) instead of PNG (FreeImage is not used).
#include <iostream>
#include <vector>

#include "decoder.h"
#include "pnm.h"

int main(int argc, const char **argv) {
  if (argc < 3) {
    std::cout << "Usage: " << argv[0] << " image outfile" << std::endl;
    return 0;
  }
  std::vector<uint8_t> packed;
  if (!himg::host::ReadFile(argv[1], &packed)) {
    std::cout << "Unable to read file " << argv[1] << std::endl;
    return -1;
  }
  std::cout << "File size: " << packed.size() << std::endl;

  himg::Decoder decoder;
  if (!decoder.Decode(packed.data(), static_cast<int>(packed.size()))) {
    std::cout << "Unable to decode image." << std::endl;
    return -1;
  }
  if (!himg::host::WritePnm(argv[2], decoder.unpacked_data(), decoder.width(), decoder.height(),
                            decoder.num_channels())) {
    std::cout << "Unable to write file " << argv[2] << std::endl;
    return -1;
  }
  return 0;
}
