// Minimal image I/O for the drivers.  The reference drivers use FreeImage (src/chimg.cpp:95-169,
// src/dhimg.cpp:17-72), which is not available here; these read/write binary PGM (P5), PPM (P6)
// and PAM (P7, for 4 channels), and can synthesise the SURVEY Appendix-B test image from a spec
// "synthetic:WxHxC[:seed[:amp]]".
#ifndef HIMG_B200_HOST_PNM_H_
#define HIMG_B200_HOST_PNM_H_

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace himg {
namespace host {

struct Image {
  int width = 0, height = 0, channels = 0;
  std::vector<uint8_t> pixels;  // [height][width][channels]
};

inline uint32_t Mix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x7feb352dU;
  h ^= h >> 15;
  h *= 0x846ca68bU;
  h ^= h >> 16;
  return h;
}

inline uint32_t Tri(uint32_t t, uint32_t period) {
  const uint32_t m = t % period;
  return m < period / 2 ? m : period - 1 - m;
}

inline void Synthesize(int w, int h, int c, uint32_t seed, uint32_t amp, Image *img) {
  img->width = w;
  img->height = h;
  img->channels = c;
  img->pixels.resize(static_cast<size_t>(w) * h * c);
  for (uint32_t y = 0; y < static_cast<uint32_t>(h); ++y)
    for (uint32_t x = 0; x < static_cast<uint32_t>(w); ++x)
      for (uint32_t k = 0; k < static_cast<uint32_t>(c); ++k) {
        int v = 40 + static_cast<int>(Tri(x * (k + 2) + y, 256)) + static_cast<int>(Tri(y * 3 + k * 40, 128)) +
                ((((x >> 6) + (y >> 6)) & 1) ? 24 : 0);
        if (amp) {
          const uint32_t r = Mix32(seed * 0x9E3779B9U + ((y * static_cast<uint32_t>(w) + x) * 4U + k));
          v += static_cast<int>(r % (2 * amp + 1)) - static_cast<int>(amp);
        }
        img->pixels[(static_cast<size_t>(y) * w + x) * c + k] = static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
      }
}

inline bool ParseSynthetic(const std::string &spec, Image *img) {
  if (spec.compare(0, 10, "synthetic:") != 0) return false;
  int w = 0, h = 0, c = 0;
  unsigned seed = 1, amp = 6;
  const int got = std::sscanf(spec.c_str() + 10, "%dx%dx%d:%u:%u", &w, &h, &c, &seed, &amp);
  if (got < 3 || w < 1 || h < 1 || c < 1 || c > 255) return false;
  Synthesize(w, h, c, seed, amp, img);
  return true;
}

inline bool SkipPnmSpace(std::istream &f) {
  for (;;) {
    const int ch = f.peek();
    if (ch == '#') {
      std::string line;
      std::getline(f, line);
    } else if (ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r') {
      f.get();
    } else {
      return f.good();
    }
  }
}

inline bool ReadPnm(const std::string &path, Image *img) {
  std::ifstream f(path.c_str(), std::ios::in | std::ios::binary);
  if (!f.good()) return false;
  std::string magic;
  f >> magic;
  int maxval = 0;
  if (magic == "P5" || magic == "P6") {
    img->channels = magic == "P5" ? 1 : 3;
    if (!SkipPnmSpace(f)) return false;
    f >> img->width;
    if (!SkipPnmSpace(f)) return false;
    f >> img->height;
    if (!SkipPnmSpace(f)) return false;
    f >> maxval;
    f.get();  // single whitespace before the raster
  } else if (magic == "P7") {
    std::string key;
    while (f >> key) {
      if (key == "ENDHDR") {
        f.get();
        break;
      } else if (key == "WIDTH") {
        f >> img->width;
      } else if (key == "HEIGHT") {
        f >> img->height;
      } else if (key == "DEPTH") {
        f >> img->channels;
      } else if (key == "MAXVAL") {
        f >> maxval;
      } else {
        std::string rest;
        std::getline(f, rest);
      }
    }
  } else {
    return false;
  }
  if (maxval != 255 || img->width < 1 || img->height < 1 || img->channels < 1 || img->channels > 255) return false;
  img->pixels.resize(static_cast<size_t>(img->width) * img->height * img->channels);
  f.read(reinterpret_cast<char *>(img->pixels.data()), static_cast<std::streamsize>(img->pixels.size()));
  return static_cast<size_t>(f.gcount()) == img->pixels.size();
}

inline bool LoadImage(const std::string &spec, Image *img) {
  return ParseSynthetic(spec, img) || ReadPnm(spec, img);
}

inline bool WritePnm(const std::string &path, const uint8_t *pixels, int w, int h, int c) {
  std::ofstream f(path.c_str(), std::ios::out | std::ios::binary);
  if (!f.good()) return false;
  if (c == 1) {
    f << "P5\n" << w << " " << h << "\n255\n";
  } else if (c == 3) {
    f << "P6\n" << w << " " << h << "\n255\n";
  } else {
    f << "P7\nWIDTH " << w << "\nHEIGHT " << h << "\nDEPTH " << c << "\nMAXVAL 255\nTUPLTYPE "
      << (c == 4 ? "RGB_ALPHA" : (c == 2 ? "GRAYSCALE_ALPHA" : "UNKNOWN")) << "\nENDHDR\n";
  }
  f.write(reinterpret_cast<const char *>(pixels), static_cast<std::streamsize>(static_cast<size_t>(w) * h * c));
  return f.good();
}

inline bool ReadFile(const std::string &path, std::vector<uint8_t> *out) {
  std::ifstream f(path.c_str(), std::ios::in | std::ios::binary);
  if (!f.good()) return false;
  f.seekg(0, std::ios::end);
  const std::streamoff n = f.tellg();
  f.seekg(0, std::ios::beg);
  out->resize(static_cast<size_t>(n));
  f.read(reinterpret_cast<char *>(out->data()), n);
  return true;
}

}  // namespace host
}  // namespace himg

#endif  // HIMG_B200_HOST_PNM_H_
