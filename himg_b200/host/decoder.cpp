#include "decoder.h"

#include <iostream>

#include "../../include/himg_cuda.h"
#include "device.h"

namespace himg {

Decoder::Decoder(int max_threads)
    : m_ctx(nullptr), m_max_threads(max_threads), m_width(0), m_height(0), m_num_channels(0) {}

Decoder::~Decoder() {
  if (m_ctx) himgcu_destroy(m_ctx);
}

bool Decoder::Decode(const uint8_t *packed_data, int packed_size) {
  m_unpacked_data.clear();
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
  m_unpacked_data.resize(static_cast<size_t>(w) * h * n);
  const int rc = himgcu_decode(m_ctx, packed_data, static_cast<size_t>(packed_size),
                               host::DefaultDecodeFlags(), m_unpacked_data.data(),
                               m_unpacked_data.size(), &m_width, &m_height, &m_num_channels);
  if (rc != HIMGCU_OK) {
    std::cout << "Error: " << himgcu_last_error(m_ctx) << "\n";
    m_unpacked_data.clear();
    return false;
  }
  return true;
}

}  // namespace himg
