// K-fwd, fast path (v2): the roofline kernel of the encoder.
//
// Same arithmetic as k_forward (xform_kernels.cuh) -- colour map + low-res subtract + rows/cols WHT
// + sign-magnitude shift quantise + 8-bit map + coefficient-planar scatter (encoder.cpp:275-328) --
// restructured for instruction count and latency:
//
//  * the 8 pixel rows of a 256-block tile are staged in shared memory by ONE thread with
//    cp.async.bulk (TMA bulk copy, completion on an mbarrier): HBM reads are fully asynchronous and
//    perfectly coalesced, 4 CTAs/SM keep ~190 KB in flight per SM;
//  * a thread owns TWO horizontally adjacent 8x8 blocks and keeps them as 16-bit lane pairs in one
//    32-bit register (lo = block A, hi = block B).  Lanes carry a bias (256 after the low-res
//    subtract, doubling per butterfly stage up to 16384) so they stay non-negative: ordinary 32-bit
//    IADD / IADD3 then act as exact 2-wide SIMD adds and subtracts with no cross-lane borrow;
//  * colour mapping uses dp4a straight on the interleaved RGB words (no per-byte unpacking).  The
//    weights are pre-scaled (x64 for Y, x128 for Cb/Cr, with the G bytes complemented by one XOR per
//    row word so that every weight is non-negative): the wanted 8-bit value then sits in byte 1 of
//    the sum and ONE byte-permute builds the lane pair -- no shift, no mask;
//  * the sign-magnitude rounding add is done on both lanes at once; the per-coefficient constants
//    come from the kernel parameter block (constant bank -> uniform registers, no shared-memory
//    traffic); the 8-bit mapping is one lookup per lane in a SIGNED table staged in shared memory;
//  * the two codes of a lane pair are adjacent bytes of a coefficient plane: one 16-bit store per
//    pair, 64 contiguous bytes per warp instruction.
//
// Preconditions (checked on the host, otherwise k_forward handles the image): pixel_stride == nch,
// width % 16 == 0, height % 8 == 0, 16-byte aligned pixel base, all shifts <= 14, nch in {1,3}.
#ifndef HIMG_B200_XFORM_FWD2_CUH_
#define HIMG_B200_XFORM_FWD2_CUH_

#include <utility>

#include "common.cuh"

namespace himgcu {

constexpr int kFwd2Threads = 256;
constexpr int kFwd2Blocks = 2 * kFwd2Threads;  // 8x8 blocks per CTA tile (2 per thread)
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
struct Fwd2Params {
  ColourW cw[4];
  QuantRec rec[2][64];  // [luma | chroma][scan position]
  int lut_half;    // the shared-memory LUT covers m in [-lut_half, lut_half]
  int tile_cols;   // blocks per tile row (even, <= 512)
  int tile_rows;   // block rows per tile (tile_rows * tile_cols <= 512)
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

// Quantise + map + store the 64 lane pairs of one channel, visiting the planes in scan order so
// that the store address is a running pointer (+cols).  `rec` points at the channel's class.
template <int I>
__device__ __forceinline__ void quant_one(const uint32_t (&x)[64], const QuantRec *rec, const uint8_t *slut,
                                          uint8_t *&dst, int cols) {
  constexpr int j = scan_coef(I);
  const QuantRec k = rec[I];  // uniform: constant bank
  // lane = T + 16384.  Sign-magnitude rounding: (T + r - [T < 0]) >> s, and [T >= 0] is bit 14.
  const uint32_t z = x[j];
  const uint32_t tz = (z >> 14) & k.tmask;
  const uint32_t zz = z + k.c2 + tz;
  const uint32_t lo = (zz << 16) >> k.s16, hi = zz >> k.s16;  // = m + (16384 >> s) per lane
  const uint32_t ca = slut[k.off + lo], cb = slut[k.off + hi];
  *reinterpret_cast<uint16_t *>(dst) = (uint16_t)(ca | (cb << 8));
  dst += cols;
}
template <int... Is>
__device__ __forceinline__ void quant_store(const uint32_t (&x)[64], const QuantRec *rec, const uint8_t *slut,
                                            uint8_t *dst, int cols, std::integer_sequence<int, Is...>) {
  (quant_one<Is>(x, rec, slut, dst, cols), ...);
}

// grid (ceil(cols/tile_cols), ceil(rows/tile_rows), n), block 256.
// dynamic smem: LUT window | low-res corners | tile (tile_rows*8 pixel rows of tile_cols*8*NCH bytes).
// The 8 warps of a CTA pass through the phases together (barriers at the phase boundaries): with
// 2 CTAs per SM the live instruction window stays inside the instruction cache.
template <int NCH, bool YCBCR>
__global__ void __launch_bounds__(kFwd2Threads, 2)
    k_forward2(const uint8_t *__restrict__ pixels, const uint8_t *__restrict__ L, Geom g,
               const __grid_constant__ Fwd2Params prm, const uint8_t *__restrict__ slut,
               uint8_t *__restrict__ planes) {
  extern __shared__ __align__(128) uint8_t lut[];  // signed map LUT first: its address is a constant
  __shared__ uint64_t bar;
  const int pitch = prm.tile_cols * 8 * NCH;  // bytes per staged pixel row (multiple of 16)
  const int v0 = blockIdx.y * prm.tile_rows, u0 = blockIdx.x * prm.tile_cols;
  const int nblk = min(prm.tile_cols, g.cols - u0);  // even
  const int nrow = min(prm.tile_rows, g.rows - v0);
  // low-res samples needed by the tile: [NCH][tile_rows + 1][tile_cols + 2], edge clamped
  const int lw = prm.tile_cols + 2, lh = prm.tile_rows + 1;
  uint8_t *sL = lut + ((2 * prm.lut_half + 1 + 15) & ~15);
  uint8_t *tile = sL + ((NCH * lh * lw + 127) & ~127);
  const uint8_t *img = pixels + (size_t)blockIdx.z * g.img_bytes;

  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t row_bytes = (uint32_t)nblk * 8 * NCH;
    mbar_expect_tx(&bar, row_bytes * 8 * nrow);
    for (int y = 0; y < 8 * nrow; ++y)
      bulk_g2s(tile + y * pitch, img + ((size_t)(8 * v0 + y) * g.w + (size_t)u0 * 8) * NCH, row_bytes, &bar);
  }
  {  // the needed window of the signed LUT and the low-res corners, while the tile is in flight
    // lut_half is a multiple of 64 and the global table is padded: aligned 128-bit copies
    const int half = prm.lut_half, nv = (2 * half + 1 + 15) >> 4;
    const uint4 *src = reinterpret_cast<const uint4 *>(slut + (kLutCenter - half));
    for (int i = threadIdx.x; i < nv; i += kFwd2Threads) reinterpret_cast<uint4 *>(lut)[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < NCH * lh * lw; i += kFwd2Threads) {
      const int c = i / (lh * lw), r = (i - c * lh * lw) / lw, q = i - c * lh * lw - r * lw;
      const int vv = min(v0 + r, g.rows - 1), uu = min(u0 + q, g.cols - 1);
      sL[i] = __ldg(L + (((size_t)blockIdx.z * NCH + c) * g.rows + vv) * g.cols + uu);
    }
  }
  __syncthreads();
  const int half_cols = prm.tile_cols >> 1;
  const int rb = threadIdx.x / half_cols;               // block row inside the tile
  const int ub = 2 * (threadIdx.x - rb * half_cols);    // first of this thread's two blocks
  const bool active = rb < nrow && ub < nblk;
  const int v = v0 + rb, u = u0 + ub;
  uint8_t *seg = planes + (size_t)blockIdx.z * g.planes_bytes + (size_t)v * g.seg + u;
  const uint8_t *trow = tile + (size_t)rb * 8 * pitch + ub * 8 * NCH;
  mbar_wait(&bar, 0);

#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    uint32_t x[64];
    __syncthreads();  // keep the CTA's warps in the same phase (instruction-cache locality)
    if (active) {
      uint32_t cw0[4], cw1[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        cw0[k] = prm.cw[c].w0[k];
        cw1[k] = prm.cw[c].w1[k];
      }
      const uint32_t cadd = prm.cw[c].add, csel = prm.cw[c].sel;
      const uint32_t xm0 = prm.cw[c].xm[0], xm1 = prm.cw[c].xm[1], xm2 = prm.cw[c].xm[2];
      const bool has_xm = prm.cw[c].has_xm != 0;
      // ---- low-res corners of both blocks (columns u, u+1, u+2; rows v, v+1), packed per lane
      uint32_t lf[9], rt[9];
      {
        const uint8_t *lp = sL + (c * lh + rb) * lw + ub;  // clamping was applied when staging
        const uint32_t t0 = lp[0], t1 = lp[1], t2 = lp[2];
        const uint32_t b0 = lp[lw], b1 = lp[lw + 1], b2 = lp[lw + 2];
        nine2(t0 | (t1 << 16), b0 | (b1 << 16), lf);  // left columns of A (lo) and B (hi)
        nine2(t1 | (t2 << 16), b1 | (b2 << 16), rt);  // right columns
      }
      // ---- rows: colour map, subtract the interpolated low-res row (bias 256), row WHT
#pragma unroll
      for (int y = 0; y < 8; ++y) {
        const uint4 *rp = reinterpret_cast<const uint4 *>(trow + y * pitch);
        uint32_t w[4 * NCH];  // block A words then block B words
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
          const uint4 q = rp[k];
          w[4 * k] = q.x;
          w[4 * k + 1] = q.y;
          w[4 * k + 2] = q.z;
          w[4 * k + 3] = q.w;
        }
        if (NCH > 1 && has_xm) {  // chroma: complement the G bytes
#pragma unroll
          for (int k = 0; k < 4 * NCH; ++k) w[k] ^= (k % 3 == 0) ? xm0 : (k % 3 == 1) ? xm1 : xm2;
        }
        uint32_t t[9];
        nine2(lf[y], rt[y], t);
        const uint32_t *wa = w, *wb = w + 2 * NCH;
#pragma unroll
        for (int i = 0; i < 8; ++i) x[y * 8 + i] = colour_pair<NCH>(wa, wb, i, cw0, cw1, cadd, csel) - t[i] + 0x01000100u;
        wht8p<256>(x[y * 8 + 0], x[y * 8 + 1], x[y * 8 + 2], x[y * 8 + 3], x[y * 8 + 4], x[y * 8 + 5], x[y * 8 + 6], x[y * 8 + 7]);
      }
      // ---- columns (bias 2048 -> 16384)
#pragma unroll
      for (int q = 0; q < 8; ++q)
        wht8p<2048>(x[q], x[8 + q], x[16 + q], x[24 + q], x[32 + q], x[40 + q], x[48 + q], x[56 + q]);
    }
#ifdef HIMG_FWD2_MID_BARRIER
    __syncthreads();
#endif
    // ---- quantise both lanes, map, store the code pair
    if (active) {
      const int cls = (YCBCR && NCH >= 3 && (c == 1 || c == 2)) ? 1 : 0;
      quant_store(x, prm.rec[cls], lut, seg + (size_t)c * g.cols * 64, g.cols, std::make_integer_sequence<int, 64>{});
    }
  }
}

}  // namespace himgcu

#endif  // HIMG_B200_XFORM_FWD2_CUH_
