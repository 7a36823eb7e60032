// himg::Decoder -- drop-in for the reference class (src/lib/decoder.h:22-67).
//
// max_threads is accepted for source compatibility and ignored: block rows are decoded by GPU
// threads, not by a std::thread pool (decoder.cpp:292-326).  Like the reference's, objects can be
// copied (a copy owns a copy of the pixels and creates its own device context on first use).
//
// The pixels live in a page-locked buffer that is kept between calls: no per-call allocation or zero
// fill, and the device -> host copy of the image runs at link speed.
#ifndef HIMG_B200_HOST_DECODER_H_
#define HIMG_B200_HOST_DECODER_H_

#include <cstddef>
#include <cstdint>

struct himgcu_ctx;

namespace himg {

class Decoder {
 public:
  Decoder(int max_threads = 0);
  ~Decoder();
  Decoder(const Decoder &other);
  Decoder &operator=(const Decoder &other);

  bool Decode(const uint8_t *packed_data, int packed_size);

  const uint8_t *unpacked_data() const { return m_unpacked_data; }
  int unpacked_size() const { return static_cast<int>(m_unpacked_size); }

  int width() const { return m_width; }
  int height() const { return m_height; }
  int num_channels() const { return m_num_channels; }

 private:
  himgcu_ctx *m_ctx;
  int m_max_threads;
  bool Reserve(size_t bytes);
  void CopyFrom(const Decoder &other);

  uint8_t *m_unpacked_data;  // page-locked (himgcu_host_alloc), m_unpacked_cap bytes
  size_t m_unpacked_cap;
  size_t m_unpacked_size;
  int m_width;
  int m_height;
  int m_num_channels;
};

}  // namespace himg

#endif  // HIMG_B200_HOST_DECODER_H_
