// himg::Encoder -- drop-in for the reference class (src/lib/encoder.h:20-64).
//
// The public interface is the reference's; the body forwards to the sm_100a CUDA path through the
// C ABI (include/himg_cuda.h).  Unlike the reference, an Encoder object can be reused: every
// Encode() behaves like a fresh object (the reference leaks m_downsampled between calls,
// encoder.cpp:198 vs :66).
#ifndef HIMG_B200_HOST_ENCODER_H_
#define HIMG_B200_HOST_ENCODER_H_

#include <cstdint>
#include <vector>

struct himgcu_ctx;

namespace himg {

class Encoder {
 public:
  Encoder();
  ~Encoder();
  Encoder(const Encoder &) = delete;
  Encoder &operator=(const Encoder &) = delete;

  // data: interleaved u8, pixel_stride bytes between pixels, rows contiguous; quality 0..100.
  bool Encode(const uint8_t *data,
              int width,
              int height,
              int pixel_stride,
              int num_channels,
              int quality,
              bool use_ycbcr);

  const uint8_t *packed_data() const { return m_packed_data.data(); }

  int packed_size() const { return static_cast<int>(m_packed_data.size()); }

  // Extension: the reference prints two progress lines on stdout (encoder.cpp:219,:334); they are
  // kept by default for CLI parity and can be silenced.
  void set_verbose(bool verbose) { m_verbose = verbose; }

 private:
  himgcu_ctx *m_ctx;
  bool m_verbose;
  std::vector<uint8_t> m_packed_data;
};

}  // namespace himg

#endif  // HIMG_B200_HOST_ENCODER_H_
