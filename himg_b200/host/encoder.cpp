#include "encoder.h"

#include <cstring>
#include <iostream>

#include "../../include/himg_cuda.h"
#include "device.h"

namespace himg {

Encoder::Encoder() : m_ctx(nullptr), m_verbose(true) {}

Encoder::~Encoder() {
  if (m_ctx) himgcu_destroy(m_ctx);
}

bool Encoder::Encode(const uint8_t *data,
                     int width,
                     int height,
                     int pixel_stride,
                     int num_channels,
                     int quality,
                     bool use_ycbcr) {
  m_packed_data.clear();
  if (!m_ctx && himgcu_create(host::DefaultDevice(), &m_ctx) != HIMGCU_OK) {
    std::cout << "HIMG: no usable CUDA device (there is no CPU fallback).\n";
    return false;
  }
  const size_t bound = himgcu_encode_bound(width, height, num_channels);
  if (bound == 0) {
    std::cout << "HIMG: unsupported image shape.\n";
    return false;
  }
  m_packed_data.resize(bound);
  size_t size = 0;
  const int rc = himgcu_encode(m_ctx, data, width, height, pixel_stride, num_channels, quality,
                               use_ycbcr ? 1 : 0, m_packed_data.data(), bound, &size);
  if (rc != HIMGCU_OK) {
    std::cout << "HIMG: encode failed: " << himgcu_last_error(m_ctx) << "\n";
    m_packed_data.clear();
    return false;
  }
  m_packed_data.resize(size);
  if (m_verbose) {
    // The reference reports the two Huffman chunk sizes while encoding.
    size_t idx = 12;
    while (idx + 8 <= size) {
      const uint8_t *p = m_packed_data.data() + idx;
      const uint32_t sz = p[4] | (p[5] << 8) | (p[6] << 16) | (static_cast<uint32_t>(p[7]) << 24);
      if (!std::memcmp(p, "LRES", 4)) std::cout << "Low resolution data: " << sz << " bytes.\n";
      if (!std::memcmp(p, "FRES", 4)) std::cout << "Full resolution data: " << sz << " bytes.\n";
      idx += 8 + sz;
    }
  }
  return true;
}

}  // namespace himg
