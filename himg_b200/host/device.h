// Device selection shared by the host classes: HIMG_DEVICE=<ordinal> (default 0).
#ifndef HIMG_B200_HOST_DEVICE_H_
#define HIMG_B200_HOST_DEVICE_H_

#include <cstdlib>

namespace himg {
namespace host {

inline int DefaultDevice() {
  const char *e = std::getenv("HIMG_DEVICE");
  return e ? std::atoi(e) : 0;
}

// HIMG_LENIENT=1 also decodes the reference encoder's streams that its own decoder refuses
// (SURVEY A.4-6); the default mirrors the reference's accept/reject decisions.
inline int DefaultDecodeFlags() {
  const char *e = std::getenv("HIMG_LENIENT");
  return (e && std::atoi(e) != 0) ? 1 : 0;
}

}  // namespace host
}  // namespace himg

#endif  // HIMG_B200_HOST_DEVICE_H_
