// Lane-pair arithmetic and asynchronous-copy helpers shared by the fast transform kernels
// (xform_fwd3.cuh, xform_inv4.cuh): two horizontally adjacent 8x8 blocks live in one register as two
// 16-bit lanes (lo = block A, hi = block B) whose values carry a bias so that plain 32-bit adds are
// exact 2-wide SIMD adds; TMA bulk copies + mbarrier, cp.async.
#ifndef HIMG_B200_XFORM_LANE_CUH_
#define HIMG_B200_XFORM_LANE_CUH_

#include <utility>

#include "common.cuh"

namespace himgcu {

constexpr int kLutCenter = 16384;  // signed map LUT: index = m + kLutCenter, m in [-16384, 16383]

// Per-coefficient quantiser record (in SCAN order), read from the kernel parameter block.
struct QuantRec {
  uint32_t c2;     // shift ? round - 1 : 0, replicated in both 16-bit lanes
  uint32_t tmask;  // shift ? 0x00010001 : 0      (negative values round half away from 0)
  uint32_t s16;    // shift + 16
  uint32_t off;    // lut_half - (16384 >> shift): re-centres a shifted lane on the shared LUT
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// packed nine-tap midpoint interpolation on two lanes (values 0..255 per lane)
__device__ __forceinline__ uint32_t mid2(uint32_t a, uint32_t b) { return ((a + b + 0x00010001u) >> 1) & 0x00ff00ffu; }
__device__ __forceinline__ void nine2(uint32_t a, uint32_t b, uint32_t (&t)[9]) {
  t[0] = a;
  t[8] = b;
  t[4] = mid2(t[0], t[8]);
  t[2] = mid2(t[0], t[4]);
  t[6] = mid2(t[4], t[8]);
  t[1] = mid2(t[0], t[2]);
  t[3] = mid2(t[2], t[4]);
  t[5] = mid2(t[4], t[6]);
  t[7] = mid2(t[6], t[8]);
}

// One butterfly level on biased lane pairs: sum doubles the bias, difference gets +2*bias so both
// outputs carry bias 2B.  K = 2B replicated in both lanes.
template <uint32_t K>
__device__ __forceinline__ void bfly(uint32_t a, uint32_t b, uint32_t &s, uint32_t &d) {
  s = a + b;
  d = a - b + K;
}

// 8-point sequency-ordered WHT on lane pairs whose bias is B on entry (8B on exit).
template <uint32_t B>
__device__ __forceinline__ void wht8p(uint32_t &x0, uint32_t &x1, uint32_t &x2, uint32_t &x3, uint32_t &x4,
                                      uint32_t &x5, uint32_t &x6, uint32_t &x7) {
  constexpr uint32_t K1 = (2 * B) * 0x00010001u, K2 = (4 * B) * 0x00010001u, K3 = (8 * B) * 0x00010001u;
  uint32_t a0, a1, a2, a3, a4, a5, a6, a7, b0, b1, b2, b3, b4, b5, b6, b7;
  bfly<K1>(x0, x4, a0, a4);
  bfly<K1>(x1, x5, a1, a5);
  bfly<K1>(x2, x6, a2, a6);
  bfly<K1>(x3, x7, a3, a7);
  bfly<K2>(a0, a2, b0, b2);
  bfly<K2>(a1, a3, b1, b3);
  bfly<K2>(a4, a6, b4, b6);
  bfly<K2>(a5, a7, b5, b7);
  bfly<K3>(b0, b1, x0, x7);
  bfly<K3>(b4, b5, x1, x6);
  bfly<K3>(b6, b7, x2, x5);
  bfly<K3>(b2, b3, x3, x4);
}

// Colour mapping as dp4a weights: sum = dot(bytes ^ xm, weights) + add over the (at most two) words
// that hold the pixel; the 8-bit result is byte 1 (scaled weights) or byte 0 (plain extraction) of
// the sum and bytes 2-3 are zero, so __byte_perm(sum_a, sum_b, sel) is the lane pair.  One code path
// serves Y / Cb / Cr, plain channel extraction and alpha, so the kernel body exists once and the
// channel loop is not unrolled.
struct ColourW {
  uint32_t w0[4], w1[4];  // weights for the pixel's first / second word, indexed by (byte offset & 3)
  uint32_t add, sel;
  uint32_t xm[3];         // XOR masks of the row words (word k uses xm[k % 3]; all equal for 4 channels)
  uint32_t has_xm;
};

__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// Colour-mapped value pair (lo = block A pixel i, hi = block B pixel i).  wa / wb: the 2*NCH raw
// row words of the two blocks (already XORed with the channel's masks).
template <int NCH>
__device__ __forceinline__ uint32_t colour_pair(const uint32_t *wa, const uint32_t *wb, int i, const uint32_t (&cw0)[4],
                                                const uint32_t (&cw1)[4], uint32_t add, uint32_t sel) {
  const int k = i * NCH, w0 = k >> 2, sh = k & 3;
  if (NCH == 1) return __byte_perm(wa[w0], wb[w0], 0x0400 | sh | ((sh + 4) << 8)) & 0x00ff00ffu;
  uint32_t sa = dp4a_uu(wa[w0], cw0[sh], add), sb = dp4a_uu(wb[w0], cw0[sh], add);
  if (sh + NCH > 4) {
    sa = dp4a_uu(wa[w0 + 1], cw1[sh], sa);
    sb = dp4a_uu(wb[w0 + 1], cw1[sh], sb);
  }
  return __byte_perm(sa, sb, sel);
}

// One inverse butterfly level on biased lane pairs (bias B in, 2B out); K = 2B in both lanes.
template <uint32_t K>
__device__ __forceinline__ void ibfly(uint32_t a, uint32_t b, uint32_t &s, uint32_t &d) {
  s = a + b;
  d = a - b + K;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

}  // namespace himgcu

#endif  // HIMG_B200_XFORM_LANE_CUH_
