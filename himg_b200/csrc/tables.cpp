// Host-side codec tables (see tables.h).
#include "tables.h"

#include <cstring>
#include <mutex>

namespace himgcu {

namespace {

// libjpeg-style base tables the reference scales (quantize.cpp:19-40).
const uint8_t kBaseLuma[64] = {
    16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
    14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
    18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kBaseChroma[64] = {
    17,  18,  24,  47,  100, 110, 115, 120, 18,  21,  26,  66,  100, 110, 118, 121,
    24,  26,  56,  100, 100, 110, 120, 122, 47,  66,  100, 100, 100, 110, 120, 123,
    100, 100, 100, 100, 100, 110, 120, 124, 110, 110, 110, 110, 110, 110, 110, 123,
    120, 120, 120, 120, 120, 110, 100, 122, 124, 124, 126, 126, 125, 123, 122, 105};

struct Knot {
  int q, v;
};
const Knot kShiftKnots[] = {{0, 65535}, {10, 32512}, {20, 13568}, {30, 5120}, {40, 2560},
                            {50, 1024}, {60, 768},   {80, 256},   {100, 0}};       // quantize.cpp:55-65
const Knot kLowKnots[] = {{0, 120}, {5, 90},  {10, 70}, {20, 40},
                          {30, 32}, {40, 26}, {50, 20}, {100, 16}};                // mapper.cpp:38-47

template <int N>
int Interpolate(const Knot (&k)[N], int q) {
  int seg = N - 1;
  for (int i = 0; i + 1 < N; ++i)
    if (k[i + 1].q > q) {
      seg = i;
      break;
    }
  if (seg >= N - 1) return k[N - 1].v;
  const int span = k[seg + 1].q - k[seg].q;
  return k[seg].v + ((k[seg + 1].v - k[seg].v) * (q - k[seg].q) + (span >> 1)) / span;  // truncating
}

// floor(log2 x) plus the bit below the MSB (quantize.cpp:94-102).
int RoundedLog2(unsigned x) {
  int msb = 0;
  for (unsigned t = x; t > 1; t >>= 1) ++msb;
  const int below = msb > 0 ? static_cast<int>((x >> (msb - 1)) & 1u) : 0;
  return msb + below;
}

void ShiftTable(const uint8_t *base, int quality, uint8_t out[64]) {
  const int scale = Interpolate(kShiftKnots, static_cast<int>(static_cast<uint8_t>(quality)));
  for (int i = 0; i < 64; ++i) {
    const unsigned c = static_cast<uint16_t>((static_cast<int>(base[i]) * scale + 512) >> 10);
    const int s = RoundedLog2(c);
    out[i] = static_cast<uint8_t>(s < 15 ? s : 15);
  }
}

// Low-res base curve (mapper.cpp:19-36): identity to 65, then the hand-tuned ramp.
const uint8_t kLowRamp[62] = {67,  68,  70,  71,  73,  74,  76,  78,  79,  81,  83,  85,  87,
                              89,  91,  93,  95,  97,  99,  102, 104, 106, 109, 111, 114, 117,
                              119, 122, 125, 128, 131, 134, 137, 140, 143, 146, 150, 153, 156,
                              160, 164, 167, 171, 175, 178, 182, 186, 190, 195, 199, 203, 207,
                              212, 216, 221, 226, 230, 235, 240, 245, 250, 255};
// Full-res curve (mapper.cpp:54-71): identity to 49, then roughly geometric.
const uint16_t kFullRamp[78] = {
    51,   52,   54,   57,   59,   62,   65,   68,   72,   76,   81,   86,   92,   98,   105,  113,
    121,  130,  140,  151,  163,  176,  190,  205,  221,  239,  259,  280,  303,  327,  354,  382,
    413,  446,  482,  520,  561,  605,  653,  703,  757,  815,  876,  942,  1013, 1087, 1167, 1252,
    1342, 1438, 1540, 1649, 1764, 1885, 2015, 2151, 2296, 2450, 2612, 2783, 2965, 3156, 3358, 3571,
    3796, 4032, 4282, 4545, 4821, 5112, 5418, 5740, 6078, 6433, 6806, 7198, 7608, 8039};

void PutU32(std::vector<uint8_t> *v, uint32_t x) {
  for (int i = 0; i < 4; ++i) v->push_back(static_cast<uint8_t>(x >> (8 * i)));
}
void PutTag(std::vector<uint8_t> *v, const char *tag) { v->insert(v->end(), tag, tag + 4); }

}  // namespace

int MapMagnitude(const uint16_t table[128], int a) {
  if (a == 0) return 0;
  // First m in [1,125] with a < t[m+1]; keep m when strictly closer to t[m], else m+1; falling
  // through (a >= t[126]) yields 127.
  for (int m = 1; m <= 125; ++m) {
    const int hi = static_cast<int16_t>(table[m + 1]);
    if (a < hi) {
      const int lo = static_cast<int16_t>(table[m]);
      return (a - lo < hi - a) ? m : m + 1;
    }
  }
  return 127;
}

void BuildEncodeTables(int quality, bool ycbcr, EncodeTables *t) {
  t->quality = quality;
  t->ycbcr = ycbcr;
  ShiftTable(kBaseLuma, quality, t->shift_luma);
  ShiftTable(kBaseChroma, quality, t->shift_chroma);
  const int ramp = Interpolate(kLowKnots, quality);
  for (int i = 0; i < 128; ++i) {
    int idx = static_cast<int16_t>((i * static_cast<int16_t>(ramp) + 8) >> 4);
    if (idx > 127) idx = 127;
    t->low_table[i] = static_cast<uint16_t>(idx <= 65 ? idx : kLowRamp[idx - 66]);
  }
  for (int i = 0; i < 128; ++i) t->full_table[i] = static_cast<uint16_t>(i <= 49 ? i : kFullRamp[i - 50]);
  for (int a = 0; a < 256; ++a) t->low_map_lut[a] = static_cast<uint8_t>(MapMagnitude(t->low_table, a));
}

const uint8_t *FullMapLut() {
  static uint8_t lut[kFullMapLutSize];
  static std::once_flag once;
  std::call_once(once, [] {
    uint16_t full[128];
    for (int i = 0; i < 128; ++i) full[i] = static_cast<uint16_t>(i <= 49 ? i : kFullRamp[i - 50]);
    for (int a = 0; a < kFullMapLutSize; ++a) lut[a] = static_cast<uint8_t>(MapMagnitude(full, a));
  });
  return lut;
}

std::vector<uint8_t> SerializeMapFun(const uint16_t table[128]) {
  int single = 0;
  while (single < 127 && static_cast<int16_t>(table[single + 1]) < 256) ++single;
  std::vector<uint8_t> out;
  out.push_back(static_cast<uint8_t>(single));
  for (int i = 1; i <= 127; ++i) {
    out.push_back(static_cast<uint8_t>(table[i] & 0xff));
    if (i > single) out.push_back(static_cast<uint8_t>(table[i] >> 8));
  }
  return out;
}

bool ParseMapFun(const uint8_t *in, int size, int16_t unmap[256]) {
  if (size < 1) return false;
  const int single = in[0];
  if (1 + single + 2 * (127 - single) != size) return false;
  const uint8_t *p = in + 1;
  unmap[0] = 0;
  for (int i = 1; i <= 127; ++i) {
    uint16_t v = *p++;
    if (i > single) v = static_cast<uint16_t>(v | (static_cast<uint16_t>(*p++) << 8));
    unmap[i] = static_cast<int16_t>(v);
    unmap[256 - i] = static_cast<int16_t>(-static_cast<int16_t>(v));
  }
  unmap[128] = unmap[129];  // table[-128] = table[-127]
  return true;
}

void BuildContainerTemplate(const EncodeTables &t, int width, int height, int nch,
                            ContainerTemplate *out) {
  std::vector<uint8_t> &h = out->head;
  h.clear();
  PutTag(&h, "RIFF");
  PutU32(&h, 0);  // patched: file size - 8
  PutTag(&h, "HIMG");
  PutTag(&h, "FRMT");
  PutU32(&h, 11);
  h.push_back(1);
  PutU32(&h, static_cast<uint32_t>(width));
  PutU32(&h, static_cast<uint32_t>(height));
  h.push_back(static_cast<uint8_t>(nch));
  h.push_back(t.ycbcr ? 1 : 0);
  const std::vector<uint8_t> lmap = SerializeMapFun(t.low_table);
  PutTag(&h, "LMAP");
  PutU32(&h, static_cast<uint32_t>(lmap.size()));
  h.insert(h.end(), lmap.begin(), lmap.end());
  PutTag(&h, "LRES");
  PutU32(&h, 0);  // patched

  std::vector<uint8_t> &m = out->mid;
  m.clear();
  PutTag(&m, "QCFG");
  PutU32(&m, t.ycbcr ? 64 : 32);
  for (int i = 0; i < 32; ++i) m.push_back(static_cast<uint8_t>((t.shift_luma[2 * i] << 4) | t.shift_luma[2 * i + 1]));
  if (t.ycbcr)
    for (int i = 0; i < 32; ++i)
      m.push_back(static_cast<uint8_t>((t.shift_chroma[2 * i] << 4) | t.shift_chroma[2 * i + 1]));
  const std::vector<uint8_t> fmap = SerializeMapFun(t.full_table);
  PutTag(&m, "FMAP");
  PutU32(&m, static_cast<uint32_t>(fmap.size()));
  m.insert(m.end(), fmap.begin(), fmap.end());
  PutTag(&m, "FRES");
  PutU32(&m, 0);  // patched
}

}  // namespace himgcu
