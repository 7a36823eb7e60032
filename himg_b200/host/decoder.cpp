#include "decoder.h"

#include <cstring>
#include <iostream>

#include "../../include/himg_cuda.h"
#include "device.h"

namespace himg {

Decoder::Decoder(int max_threads)
    : m_ctx(nullptr), m_max_threads(max_threads), m_unpacked_data(nullptr), m_unpacked_cap(0), m_unpacked_size(0),
      m_width(0), m_height(0), m_num_channels(0) {}

Decoder::~Decoder() {
  if (m_unpacked_data) himgcu_host_free(m_unpacked_data);
  if (m_ctx) himgcu_destroy(m_ctx);
}

Decoder::Decoder(const Decoder &other)
    : m_ctx(nullptr), m_max_threads(other.m_max_threads), m_unpacked_data(nullptr), m_unpacked_cap(0), m_unpacked_size(0),
      m_width(0), m_height(0), m_num_channels(0) {
  CopyFrom(other);
}

Decoder &Decoder::operator=(const Decoder &other) {
  if (this != &other) {
    m_max_threads = other.m_max_threads;
    CopyFrom(other);
  }
  return *this;
}

void Decoder::CopyFrom(const Decoder &other) {
  m_unpacked_size = 0;
  m_width = m_height = m_num_channels = 0;
  if (other.m_unpacked_size && Reserve(other.m_unpacked_size)) {
    std::memcpy(m_unpacked_data, other.m_unpacked_data, other.m_unpacked_size);
    m_unpacked_size = other.m_unpacked_size;
    m_width = other.m_width;
    m_height = other.m_height;
    m_num_channels = other.m_num_channels;
  }
}

bool Decoder::Reserve(size_t bytes) {
  if (m_unpacked_cap >= bytes) return true;
  if (m_unpacked_data) himgcu_host_free(m_unpacked_data);
  m_unpacked_cap = 0;
  m_unpacked_data = static_cast<uint8_t *>(himgcu_host_alloc(bytes));
  if (!m_unpacked_data) return false;
  m_unpacked_cap = bytes;
  return true;
}

bool Decoder::Decode(const uint8_t *packed_data, int packed_size) {
  m_unpacked_size = 0;
  if (packed_size < 0) return false;
  int w = 0, h = 0, n = 0;
  if (himgcu_decode_info(packed_data, static_cast<size_t>(packed_size), &w, &h, &n) != HIMGCU_OK) {
    std::cout << "Not a RIFF HIMG file.\n";  // same diagnostics channel as the reference (stdout)
    return false;
  }
  if (w < 1 || h < 1 || n < 1) {
    std::cout << "Error decoding header.\n";
    return false;
  }
  if (!m_ctx && himgcu_create(host::DefaultDevice(), &m_ctx) != HIMGCU_OK) {
    std::cout << "HIMG: no usable CUDA device (there is no CPU fallback).\n";
    return false;
  }
  const size_t bytes = static_cast<size_t>(w) * h * n;
  if (!Reserve(bytes)) {
    std::cout << "HIMG: out of page-locked host memory.\n";
    return false;
  }
  const int rc = himgcu_decode(m_ctx, packed_data, static_cast<size_t>(packed_size),
                               host::DefaultDecodeFlags(), m_unpacked_data, bytes, &m_width, &m_height,
                               &m_num_channels);
  if (rc != HIMGCU_OK) {
    std::cout << "Error: " << himgcu_last_error(m_ctx) << "\n";
    return false;
  }
  m_unpacked_size = bytes;
  return true;
}

}  // namespace himg
