// himg::Decoder -- drop-in for the reference class (src/lib/decoder.h:22-67).
//
// max_threads is accepted for source compatibility and ignored: block rows are decoded by GPU
// threads, not by a std::thread pool (decoder.cpp:292-326).
#ifndef HIMG_B200_HOST_DECODER_H_
#define HIMG_B200_HOST_DECODER_H_

#include <cstdint>
#include <vector>

struct himgcu_ctx;

namespace himg {

class Decoder {
 public:
  Decoder(int max_threads = 0);
  ~Decoder();
  Decoder(const Decoder &) = delete;
  Decoder &operator=(const Decoder &) = delete;

  bool Decode(const uint8_t *packed_data, int packed_size);

  const uint8_t *unpacked_data() const { return m_unpacked_data.data(); }
  int unpacked_size() const { return static_cast<int>(m_unpacked_data.size()); }

  int width() const { return m_width; }
  int height() const { return m_height; }
  int num_channels() const { return m_num_channels; }

 private:
  himgcu_ctx *m_ctx;
  int m_max_threads;
  std::vector<uint8_t> m_unpacked_data;
  int m_width;
  int m_height;
  int m_num_channels;
};

}  // namespace himg

#endif  // HIMG_B200_HOST_DECODER_H_
