// himg::Encoder -- drop-in for the reference class (src/lib/encoder.h:20-64).
//
// The public interface is the reference's; the body forwards to the sm_100a CUDA path through the
// C ABI (include/himg_cuda.h).  Unlike the reference, an Encoder object can be reused: every
// Encode() behaves like a fresh object (the reference leaks m_downsampled between calls,
// encoder.cpp:198 vs :66).  Like the reference's, objects can be copied: a copy owns
// a copy of the packed data and creates its own device context on first use.
//
// The packed data lives in a page-locked buffer that is kept between calls (no per-call allocation or
// zero fill, device -> host copies at link speed).
#ifndef HIMG_B200_HOST_ENCODER_H_
#define HIMG_B200_HOST_ENCODER_H_

#include <cstddef>
#include <cstdint>

struct himgcu_ctx;

namespace himg {

class Encoder {
 public:
  Encoder();
  ~Encoder();
  Encoder(const Encoder &other);
  Encoder &operator=(const Encoder &other);

  // data: interleaved u8, pixel_stride bytes between pixels, rows contiguous; quality 0..100.
  bool Encode(const uint8_t *data,
              int width,
              int height,
              int pixel_stride,
              int num_channels,
              int quality,
              bool use_ycbcr);

  const uint8_t *packed_data() const { return m_packed_data; }

  int packed_size() const { return static_cast<int>(m_packed_size); }

  // Extension: the reference prints two progress lines on stdout (encoder.cpp:219,:334); they are
  // kept by default for CLI parity and can be silenced.
  void set_verbose(bool verbose) { m_verbose = verbose; }

 private:
  bool Reserve(size_t bytes);

  himgcu_ctx *m_ctx;
  bool m_verbose;
  uint8_t *m_packed_data;  // page-locked (himgcu_host_alloc), m_packed_cap bytes
  size_t m_packed_cap;
  size_t m_packed_size;
};

}  // namespace himg

#endif  // HIMG_B200_HOST_ENCODER_H_
