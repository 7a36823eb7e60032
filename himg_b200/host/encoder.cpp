#include "encoder.h"

#include <cstring>
#include <iostream>

#include "../../include/himg_cuda.h"
#include "device.h"

namespace himg {

Encoder::Encoder() : m_ctx(nullptr), m_verbose(true), m_packed_data(nullptr), m_packed_cap(0), m_packed_size(0) {}

Encoder::~Encoder() {
  if (m_packed_data) himgcu_host_free(m_packed_data);
  if (m_ctx) himgcu_destroy(m_ctx);
}

Encoder::Encoder(const Encoder &other)
    : m_ctx(nullptr), m_verbose(other.m_verbose), m_packed_data(nullptr), m_packed_cap(0), m_packed_size(0) {
  if (other.m_packed_size && Reserve(other.m_packed_size)) {
    std::memcpy(m_packed_data, other.m_packed_data, other.m_packed_size);
    m_packed_size = other.m_packed_size;
  }
}

Encoder &Encoder::operator=(const Encoder &other) {
  if (this != &other) {
    m_verbose = other.m_verbose;
    m_packed_size = 0;
    if (other.m_packed_size && Reserve(other.m_packed_size)) {
      std::memcpy(m_packed_data, other.m_packed_data, other.m_packed_size);
      m_packed_size = other.m_packed_size;
    }
  }
  return *this;
}

bool Encoder::Reserve(size_t bytes) {
  if (m_packed_cap >= bytes) return true;
  if (m_packed_data) himgcu_host_free(m_packed_data);
  m_packed_cap = 0;
  m_packed_data = static_cast<uint8_t *>(himgcu_host_alloc(bytes));
  if (!m_packed_data) return false;
  m_packed_cap = bytes;
  return true;
}

bool Encoder::Encode(const uint8_t *data,
                     int width,
                     int height,
                     int pixel_stride,
                     int num_channels,
                     int quality,
                     bool use_ycbcr) {
  m_packed_size = 0;
  if (!m_ctx && himgcu_create(host::DefaultDevice(), &m_ctx) != HIMGCU_OK) {
    std::cout << "HIMG: no usable CUDA device (there is no CPU fallback).\n";
    return false;
  }
  const size_t bound = himgcu_encode_bound(width, height, num_channels);
  if (bound == 0) {
    std::cout << "HIMG: unsupported image shape.\n";
    return false;
  }
  if (!Reserve(bound)) {
    std::cout << "HIMG: out of page-locked host memory.\n";
    return false;
  }
  size_t size = 0;
  const int rc = himgcu_encode(m_ctx, data, width, height, pixel_stride, num_channels, quality,
                               use_ycbcr ? 1 : 0, m_packed_data, m_packed_cap, &size);
  if (rc != HIMGCU_OK) {
    std::cout << "HIMG: encode failed: " << himgcu_last_error(m_ctx) << "\n";
    return false;
  }
  m_packed_size = size;
  if (m_verbose) {
    // The reference reports the two Huffman chunk sizes while encoding.
    size_t idx = 12;
    while (idx + 8 <= size) {
      const uint8_t *p = m_packed_data + idx;
      const uint32_t sz = p[4] | (p[5] << 8) | (p[6] << 16) | (static_cast<uint32_t>(p[7]) << 24);
      if (!std::memcmp(p, "LRES", 4)) std::cout << "Low resolution data: " << sz << " bytes.\n";
      if (!std::memcmp(p, "FRES", 4)) std::cout << "Full resolution data: " << sz << " bytes.\n";
      idx += 8 + sz;
    }
  }
  return true;
}

}  // namespace himg
