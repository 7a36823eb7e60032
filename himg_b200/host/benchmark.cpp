// benchmark -- 30-iteration wall-clock timer (mirrors src/benchmark.cpp:21-159, same output).
//   benchmark [-d][-e] image
//   -d  decode a .himg file (default)
//   -e  encode: the reference leaves this mode unimplemented (benchmark.cpp:137-139); here it
//       encodes a PGM/PPM/PAM file or "synthetic:WxHxC[:seed[:amp]]" at quality 50.
#include <chrono>
#include <iostream>
#include <string>
#include <vector>

#include <cstring>

#include "../../include/himg_cuda.h"
#include "decoder.h"
#include "encoder.h"
#include "pnm.h"

namespace {

const int kNumIterations = 30;

double NowMs() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

bool IsHimg(const std::vector<uint8_t> &b) {
  return b.size() >= 12 && b[0] == 'R' && b[1] == 'I' && b[2] == 'F' && b[3] == 'F' && b[8] == 'H' &&
         b[9] == 'I' && b[10] == 'M' && b[11] == 'G';
}

void ShowUsage(const char *arg0) {
  std::cout << "Usage: " << arg0 << " [-d][-e] image" << std::endl;
  std::cout << "  -d Decode (default)" << std::endl;
  std::cout << "  -e Encode" << std::endl;
}

}  // namespace

int main(int argc, const char **argv) {
  bool encode = false;
  std::string file_name;
  for (int i = 1; i < argc; ++i) {
    const std::string arg = argv[i];
    if (arg == "-d") {
      encode = false;
    } else if (arg == "-e") {
      encode = true;
    } else if (file_name.empty()) {
      file_name = arg;
    } else {
      ShowUsage(argv[0]);
      return 0;
    }
  }
  if (file_name.empty()) {
    ShowUsage(argv[0]);
    return 0;
  }

  std::vector<uint8_t> buffer;
  himg::host::Image img;
  if (encode) {
    if (!himg::host::LoadImage(file_name, &img)) {
      std::cout << "Unable to read file " << file_name << std::endl;
      return -1;
    }
  } else {
    if (!himg::host::ReadFile(file_name, &buffer)) {
      std::cout << "Unable to read file " << file_name << std::endl;
      return -1;
    }
    std::cout << "File size: " << buffer.size() << std::endl;
    if (!IsHimg(buffer)) {
      std::cout << "Not a RIFF HIMG file." << std::endl;
      return -1;
    }
  }

  // Pixels in page-locked memory: the encoder then copies them to the device at link speed (any other
  // host pointer works too, through the library's staging buffers).
  uint8_t *pixels = nullptr;
  if (encode) {
    pixels = static_cast<uint8_t *>(himgcu_host_alloc(img.pixels.size()));
    if (!pixels) {
      std::cout << "Out of page-locked host memory." << std::endl;
      return -1;
    }
    std::memcpy(pixels, img.pixels.data(), img.pixels.size());
  }

  // Start the device before the clock does (the first CUDA call of a process takes about a second; the
  // reference has no such start-up, and the encode branch has paid it in himgcu_host_alloc already).
  himgcu_host_free(himgcu_host_alloc(64));

  himg::Decoder decoder;
  himg::Encoder encoder;
  encoder.set_verbose(false);
  double min_dt = -1.0, max_dt = -1.0, total_t = 0.0;
  for (int iteration = 1; iteration <= kNumIterations; ++iteration) {
    std::cout << "Iteration " << iteration << "/" << kNumIterations << std::endl;
    const double t0 = NowMs();
    if (encode) {
      if (!encoder.Encode(pixels, img.width, img.height, img.channels, img.channels, 50, true)) {
        std::cout << "Unable to encode image." << std::endl;
        return -1;
      }
    } else if (!decoder.Decode(buffer.data(), static_cast<int>(buffer.size()))) {
      std::cout << "Unable to decode image." << std::endl;
      return -1;
    }
    const double dt = NowMs() - t0;
    if (min_dt < 0.0 || dt < min_dt) min_dt = dt;
    if (max_dt < 0.0 || dt > max_dt) max_dt = dt;
    total_t += dt;
  }
  std::cout << "    Min: " << min_dt << " ms\n";
  std::cout << "    Max: " << max_dt << " ms\n";
  std::cout << "Average: " << total_t / kNumIterations << " ms\n";
  himgcu_host_free(pixels);
  return 0;
}
