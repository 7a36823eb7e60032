// K-inv, fast path (v3): the roofline kernel of the decoder.
//
// Same arithmetic as k_inverse (xform_kernels.cuh): gather + dequantise + inverse WHT with a floor
// >>3 after each pass + low-res add + clamp + inverse colour map (decoder.cpp:366-423,
// quantize.cpp:153-165, hadamard.cpp:90-103, ycbcr.cpp:54-82).  Structure:
//
//  * a thread owns TWO horizontally adjacent blocks as 16-bit lane pairs in one register, like
//    K-fwd.  The inverse sums eight int16 values before each floor shift (19 bits), so the packing
//    is only exact while every dequantised coefficient lies in [-4096, 4095]: lanes carry a bias
//    of 4096, three butterfly stages bring it to 32768, the floor shift back to 4096, and no lane
//    ever leaves 16 bits.  Every image an encoder can produce is far inside that range; a stream
//    may still carry anything (the tables travel in-band), so the dequantiser ORs all lanes
//    together, the warp votes, and a warp with an out-of-range lane redoes the channel with the
//    int32 arithmetic of k_inverse.  Results are bit-exact either way;
//  * dequantisation is one lookup per sample in a table of (int16)(unmap << s) + 4096 for the 16
//    possible shifts (built per CTA: the tables are per image);
//  * low-res add + clamp is one DPX instruction per lane pair (__viaddmin_s16x2_relu), and so is each
//    output of the inverse colour map;
//  * the coefficient planes of a tile (2 block rows x <= 256 blocks for wide images) are staged with
//    cp.async at a compile-time pitch: every gather is an LDS with an immediate offset.  Finished
//    samples overwrite the thread's own (consumed) codes, so no second staging buffer is needed;
//  * pixels leave as 16-byte stores (48 contiguous bytes per thread and pixel row for RGB).
//
// Preconditions (host checked, else k_inverse2 / k_inverse): cols % 16 == 0, height % 8 == 0,
// 16-byte aligned planes / pixels, nch in {1, 3}.
#ifndef HIMG_B200_XFORM_INV3_CUH_
#define HIMG_B200_XFORM_INV3_CUH_

#include "common.cuh"
#include "xform_fwd2.cuh"  // mid2 / nine2
#include "xform_inv2.cuh"  // cp_async16

namespace himgcu {

constexpr int kInv3Threads = 256;
constexpr uint32_t kInv3Bias = 4096;

// One inverse butterfly level on biased lane pairs (bias B in, 2B out); K = 2B in both lanes.
template <uint32_t K>
__device__ __forceinline__ void ibfly(uint32_t a, uint32_t b, uint32_t &s, uint32_t &d) {
  s = a + b;
  d = a - b + K;
}
// 8-point sequency-ordered WHT (same flow graph as wht8 in common.cuh) on lane pairs with bias 4096
// on entry; the floor shift by 3 brings the bias of 32768 back to 4096.
__device__ __forceinline__ void iwht8p(uint32_t &x0, uint32_t &x1, uint32_t &x2, uint32_t &x3, uint32_t &x4,
                                       uint32_t &x5, uint32_t &x6, uint32_t &x7) {
  constexpr uint32_t K1 = 2 * kInv3Bias * 0x00010001u, K2 = 4 * kInv3Bias * 0x00010001u, K3 = 8 * kInv3Bias * 0x00010001u;
  uint32_t a0, a1, a2, a3, a4, a5, a6, a7, b0, b1, b2, b3, b4, b5, b6, b7;
  ibfly<K1>(x0, x4, a0, a4);
  ibfly<K1>(x1, x5, a1, a5);
  ibfly<K1>(x2, x6, a2, a6);
  ibfly<K1>(x3, x7, a3, a7);
  ibfly<K2>(a0, a2, b0, b2);
  ibfly<K2>(a1, a3, b1, b3);
  ibfly<K2>(a4, a6, b4, b6);
  ibfly<K2>(a5, a7, b5, b7);
  ibfly<K3>(b0, b1, x0, x7);
  ibfly<K3>(b4, b5, x1, x6);
  ibfly<K3>(b6, b7, x2, x5);
  ibfly<K3>(b2, b3, x3, x4);
  x0 = (x0 >> 3) & 0x1fff1fffu;
  x1 = (x1 >> 3) & 0x1fff1fffu;
  x2 = (x2 >> 3) & 0x1fff1fffu;
  x3 = (x3 >> 3) & 0x1fff1fffu;
  x4 = (x4 >> 3) & 0x1fff1fffu;
  x5 = (x5 >> 3) & 0x1fff1fffu;
  x6 = (x6 >> 3) & 0x1fff1fffu;
  x7 = (x7 >> 3) & 0x1fff1fffu;
}

// grid (ceil(cols / tile_cols), ceil(rows / TROWS), n), block 256; PITCH = bytes per staged plane row
// (>= tile_cols), TROWS = 512 / PITCH block rows per tile.
// dynamic smem: planes tile [TROWS][NCH * 64][PITCH] | dequantisation tables [16][256] u16
template <int NCH, int PITCH>
__global__ void __launch_bounds__(kInv3Threads, 2)
    k_inverse3(const uint8_t *__restrict__ planes, const uint8_t *__restrict__ R, Geom g,
               const DecTables *__restrict__ tabs, unsigned long long tab_stride, int tile_cols,
               uint8_t *__restrict__ pixels) {
  constexpr int TROWS = 512 / PITCH, HALF = PITCH / 2;
  extern __shared__ __align__(128) uint8_t sPl[];
  __shared__ uint16_t sTabOff[2][64];  // byte offset of the table of coefficient j: shift * 512
  uint16_t *sDq = reinterpret_cast<uint16_t *>(sPl + TROWS * NCH * 64 * PITCH);

  const int t = threadIdx.x, rb = t / HALF, tt = t % HALF;
  const int v0 = blockIdx.y * TROWS, u0 = blockIdx.x * tile_cols;
  const int nblk = min(tile_cols, g.cols - u0);  // multiple of 16
  const int nrow = min(TROWS, g.rows - v0);
  {
    // one warp per plane row, one 16-byte chunk per lane
    const int lane = t & 31, w = t >> 5, nchunk = nblk >> 4;
    if (lane < nchunk)
      for (int rr = 0; rr < nrow; ++rr) {
        const uint8_t *seg = planes + (size_t)blockIdx.z * g.planes_bytes + (size_t)(v0 + rr) * g.seg + u0;
        for (int r = w; r < NCH * 64; r += kInv3Threads / 32)
          cp_async16(sPl + (rr * NCH * 64 + r) * PITCH + lane * 16, seg + (size_t)r * g.cols + lane * 16);
      }
  }
  const DecTables *T = reinterpret_cast<const DecTables *>(reinterpret_cast<const char *>(tabs) + (size_t)blockIdx.z * tab_stride);
  {
    const int un = T->full_unmap[t];
#pragma unroll
    for (int s = 0; s < 16; ++s) sDq[s * 256 + t] = (uint16_t)((int)(short)(un << s) + (int)kInv3Bias);
    if (t < 128) sTabOff[t >> 6][t & 63] = (uint16_t)(T->shift[t >> 6][t & 63] * 512);
  }
  const bool ycbcr = T->ycbcr != 0;
  const bool active = rb < nrow && 2 * tt < nblk;
  const int v = v0 + rb, u = u0 + 2 * tt;
  uint8_t *col = sPl + rb * NCH * 64 * PITCH + 2 * tt;  // this thread's two byte columns of the tile
  const uint8_t *dq = reinterpret_cast<const uint8_t *>(sDq);
  cp_async_wait_all();

#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    __syncthreads();  // staged data visible; warps stay in the same phase
    uint8_t *cc = col + c * 64 * PITCH;
    const uint16_t *toff = sTabOff[(ycbcr && NCH >= 3 && (c == 1 || c == 2)) ? 1 : 0];
    // ---- low-res corners of both blocks (columns u, u+1, u+2; rows v, v+1); loaded early, used late
    uint32_t ctop, cbot, ctop1, cbot1;
    {
      const uint8_t *Rc = R + ((size_t)blockIdx.z * NCH + c) * g.rows * g.cols;
      const int vv = min(v, g.rows - 1), v2 = min(v + 1, g.rows - 1);
      const int ua = min(u, g.cols - 1), ub = min(u + 1, g.cols - 1), uc = min(u + 2, g.cols - 1);
      const uint32_t t0 = __ldg(Rc + vv * g.cols + ua), t1 = __ldg(Rc + vv * g.cols + ub), t2 = __ldg(Rc + vv * g.cols + uc);
      const uint32_t b0 = __ldg(Rc + v2 * g.cols + ua), b1 = __ldg(Rc + v2 * g.cols + ub), b2 = __ldg(Rc + v2 * g.cols + uc);
      ctop = t0 | (t1 << 16);   // top-left corners of A (lo) and B (hi)
      cbot = b0 | (b1 << 16);
      ctop1 = t1 | (t2 << 16);  // top-right corners
      cbot1 = b1 | (b2 << 16);
    }
    // ---- gather + dequantise both blocks (biased lanes), remember every bit that was ever set
    uint32_t x[64], seen = 0;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      const uint32_t ca = cc[scan_pos(j) * PITCH], cb = cc[scan_pos(j) * PITCH + 1];
      const uint8_t *tab = dq + toff[j];
      const uint32_t a = *reinterpret_cast<const uint16_t *>(tab + (ca << 1));
      const uint32_t b = *reinterpret_cast<const uint16_t *>(tab + (cb << 1));
      x[j] = a | (b << 16);
      seen |= x[j];
    }
    const bool narrow = __all_sync(0xffffffffu, !active || (seen & 0xe000e000u) == 0);
    uint32_t lf[9], rt[9];
    nine2(ctop, cbot, lf);    // left columns of A (lo) and B (hi)
    nine2(ctop1, cbot1, rt);  // right columns
    if (narrow) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
        iwht8p(x[r * 8 + 0], x[r * 8 + 1], x[r * 8 + 2], x[r * 8 + 3], x[r * 8 + 4], x[r * 8 + 5], x[r * 8 + 6], x[r * 8 + 7]);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        iwht8p(x[q], x[8 + q], x[16 + q], x[24 + q], x[32 + q], x[40 + q], x[48 + q], x[56 + q]);
#pragma unroll
      for (int y = 0; y < 8; ++y) {
        uint32_t tl[9];
        nine2(lf[y], rt[y], tl);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // lane + (low-res - 4096), min 255, max 0: two samples per instruction
          const uint32_t px = __viaddmin_s16x2_relu(x[y * 8 + i], tl[i] + 0xf000f000u, 0x00ff00ffu);
          // (inactive threads own their two columns of the tile too: no guard needed)
          *reinterpret_cast<uint16_t *>(cc + (y * 8 + i) * PITCH) = (uint16_t)__byte_perm(px, 0u, 0x4420);
        }
      }
    } else {
      // ---- some lane of this warp left [-4096, 4095]: int32 arithmetic, one block at a time
#pragma unroll 1
      for (int blk = 0; blk < 2; ++blk) {
        int y32[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) y32[j] = (int)(short)(((x[j] >> (16 * blk)) & 0xffffu) - kInv3Bias);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          wht8(y32[r * 8 + 0], y32[r * 8 + 1], y32[r * 8 + 2], y32[r * 8 + 3], y32[r * 8 + 4], y32[r * 8 + 5], y32[r * 8 + 6],
               y32[r * 8 + 7]);
#pragma unroll
          for (int i = 0; i < 8; ++i) y32[r * 8 + i] >>= 3;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          wht8(y32[q], y32[8 + q], y32[16 + q], y32[24 + q], y32[32 + q], y32[40 + q], y32[48 + q], y32[56 + q]);
#pragma unroll
          for (int i = 0; i < 8; ++i) y32[i * 8 + q] >>= 3;
        }
#pragma unroll
        for (int y = 0; y < 8; ++y) {
          int tl[9];
          nine((int)((lf[y] >> (16 * blk)) & 0xffu), (int)((rt[y] >> (16 * blk)) & 0xffu), tl);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int px = __vimin_s32_relu((int)(short)(y32[y * 8 + i] + tl[i]), 255);
            cc[(y * 8 + i) * PITCH + blk] = (uint8_t)px;
          }
        }
      }
    }
  }
  if (!active) return;  // (a thread only reads back its own two columns: no barrier needed)

  // ---- inverse colour map + interleave + 16-byte stores
  uint8_t *img = pixels + (size_t)blockIdx.z * g.out_img_bytes;
  const bool do_colour = ycbcr && NCH >= 3;
#pragma unroll
  for (int y = 0; y < 8; ++y) {
    uint32_t s[8 * NCH];  // lane pairs in memory order: pixel-major, channel-minor
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t ch[NCH];
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        ch[c] = __byte_perm((uint32_t) * reinterpret_cast<const uint16_t *>(col + (c * 64 + y * 8 + i) * PITCH), 0u, 0x4140);
      if (NCH >= 3 && do_colour) {
        // cb = 2Cb - 255, cr = 2Cr - 255, G = Y - ((cb + cr + 2) >> 2) = Y + 128 - ((Cb + Cr + 2) >> 1)
        const uint32_t h = ((ch[1] + ch[NCH >= 3 ? 2 : 0] + 0x00020002u) >> 1) & 0x01ff01ffu;
        const uint32_t G = ch[0] + 0x01800180u - h;                      // G + 256 per lane, positive
        const uint32_t cr = (ch[NCH >= 3 ? 2 : 0] << 1) + 0xfe01fe01u;   // cr - 256 per lane (int16)
        const uint32_t cb = (ch[1] << 1) + 0xfe01fe01u;
        ch[0] = __viaddmin_s16x2_relu(G, cr, 0x00ff00ffu);
        ch[1] = __viaddmin_s16x2_relu(G, 0xff00ff00u, 0x00ff00ffu);
        ch[NCH >= 3 ? 2 : 0] = __viaddmin_s16x2_relu(G, cb, 0x00ff00ffu);
      }
#pragma unroll
      for (int c = 0; c < NCH; ++c) s[i * NCH + c] = ch[c];
    }
    // words of block A and of block B: four consecutive lane pairs -> one word each
    uint32_t wa[2 * NCH], wb[2 * NCH];
#pragma unroll
    for (int m = 0; m < 2 * NCH; ++m) {
      const uint32_t p = __byte_perm(s[4 * m], s[4 * m + 1], 0x6240), q = __byte_perm(s[4 * m + 2], s[4 * m + 3], 0x6240);
      wa[m] = __byte_perm(p, q, 0x5410);
      wb[m] = __byte_perm(p, q, 0x7632);
    }
    uint4 *dst = reinterpret_cast<uint4 *>(img + ((size_t)(8 * v + y) * g.w + (size_t)u * 8) * NCH);
    if (NCH == 1) {
      dst[0] = make_uint4(wa[0], wa[1], wb[0], wb[1]);
    } else {
      dst[0] = make_uint4(wa[0], wa[1], wa[2], wa[3]);
      dst[1] = make_uint4(wa[4], wa[5], wb[0], wb[1]);
      dst[2] = make_uint4(wb[2], wb[3], wb[4], wb[5]);
    }
  }
}

}  // namespace himgcu

#endif  // HIMG_B200_XFORM_INV3_CUH_
