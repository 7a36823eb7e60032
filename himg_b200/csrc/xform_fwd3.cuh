// K-fwd, fast path (v3): the roofline kernel of the encoder.
//
// Same arithmetic as k_forward (xform_kernels.cuh) -- colour map + low-res subtract + rows/cols WHT
// + sign-magnitude shift quantise + 8-bit map + coefficient-planar scatter (encoder.cpp:275-328).
// The kernel is bound by instruction issue (the integer pipe of an SM sub-partition accepts one warp
// instruction every other cycle), so everything here is about instructions per sample:
//
//  * FLAT TILING: a CTA of 256 threads owns 256 CONSECUTIVE block pairs of the image in row-major
//    order, whatever the image width -- a tile may start and end in the middle of a block row and
//    span several of them.  No thread idles on 1080p / 4K (240 / 480 blocks per row do not divide
//    into 512-block tiles: the row-aligned tiles of v2 left 1 thread in 16 without work);
//  * the pixel rows of every piece of a block row inside the tile are staged by TMA bulk copies
//    (cp.async.bulk + mbarrier), issued by the lanes of warp 0;
//  * a thread owns TWO horizontally adjacent 8x8 blocks as 16-bit lane pairs in one register
//    (lo = block A, hi = block B); lanes carry a bias so that plain 32-bit adds are exact 2-wide
//    SIMD adds; colour mapping is dp4a on the interleaved words with pre-scaled weights (see v2);
//  * the per-coefficient quantiser constants are one 16-byte shared-memory record per coefficient
//    (one broadcast LDS.128 per lane pair; v2 read them from the parameter block with a run-time
//    class index and paid two LDC.64 per lane pair -- sm_100 has no constant-bank ALU operands);
//  * the low-res corners of a thread's two blocks are loaded by the thread itself while the tile is
//    in flight (v2 staged a window through shared memory with two integer divisions per sample);
//  * the plane stride is a template parameter for the common widths (1080p, 4K, 8K): the 64 stores
//    of a channel then use one base register and immediate offsets instead of a 64-bit add each.
//
// Preconditions (checked on the host, otherwise k_forward handles the image): pixel_stride == nch,
// width % 16 == 0, height % 8 == 0, 16-byte aligned pixel base, 2-byte aligned low-res base, all
// shifts <= 14, nch in {1, 3, 4}.
#ifndef HIMG_B200_XFORM_FWD3_CUH_
#define HIMG_B200_XFORM_FWD3_CUH_

#include <utility>

#include "common.cuh"
#include "xform_lane.cuh"  // lane-pair helpers, TMA / mbarrier wrappers, ColourW, QuantRec

namespace himgcu {

constexpr int kFwd3Threads = 256;  // = block pairs per CTA

struct Fwd3Params {
  ColourW cw[4];
  QuantRec rec[2][64];  // [luma | chroma][scan position]
  int lut_half;         // the shared-memory LUT covers m in [-lut_half, lut_half]
};

// Quantise + map + store the 64 lane pairs of one channel.  The per-coefficient constants of the
// channel's class sit in shared memory as one 16-byte record per scan position (one LDS.128, broadcast):
//   x = (r - 1) in both lanes | y = rounding mask | z = shift + 16 | w = LUT base + re-centring offset
// (sm_100 has no constant-bank operands: reading them from the parameter block costs one LDC per two
// words).  Plane I of the channel starts I * cols bytes after plane 0: an immediate when COLS is known.
template <int COLS, int I>
__device__ __forceinline__ void quant3_one(const uint32_t (&x)[64], const uint4 *rec, uint8_t *dst, uint32_t cols) {
  constexpr int j = scan_coef(I);
  const uint4 k = rec[I];
  // lane = T + 16384.  Sign-magnitude rounding: (T + r - [T < 0]) >> s, and [T >= 0] is bit 14.
  const uint32_t z = x[j];
  const uint32_t tz = (z >> 14) & k.y;
  const uint32_t zz = z + k.x + tz;
  const uint32_t lo = (zz << 16) >> k.z, hi = zz >> k.z;  // = m + (16384 >> s) per lane
  uint32_t ca, cb;  // the record carries the shared-memory ADDRESS of LUT entry 0 of this coefficient
  asm("ld.shared.u8 %0, [%1];" : "=r"(ca) : "r"(k.w + lo));
  asm("ld.shared.u8 %0, [%1];" : "=r"(cb) : "r"(k.w + hi));
  const uint16_t code = (uint16_t)(ca | (cb << 8));
  if (COLS) *reinterpret_cast<uint16_t *>(dst + I * COLS) = code;
  else *reinterpret_cast<uint16_t *>(dst + (size_t)(I * cols)) = code;
}
template <int COLS, int... Is>
__device__ __forceinline__ void quant3_store(const uint32_t (&x)[64], const uint4 *rec, uint8_t *dst, uint32_t cols,
                                             std::integer_sequence<int, Is...>) {
  (quant3_one<COLS, Is>(x, rec, dst, cols), ...);
}

// Rows of one channel of a thread's block pair: colour map, low-res subtract, row WHT, then the
// column WHT (all in registers).  On return x holds T + 16384 per lane.
//   trow / pitch  this thread's 16-pixel row 0 in the staged tile, bytes between pixel rows
//   top / bot     low-res corners of the pair: bytes (u, u+1, u+2) of rows v and v+1 (edge clamped)
template <int NCH>
__device__ __forceinline__ void fwd3_rows(const Fwd3Params &prm, int c, const uint8_t *trow, uint32_t pitch, uint32_t top,
                                          uint32_t bot, uint32_t (&x)[64]) {
  uint32_t cw0[4], cw1[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    cw0[k] = prm.cw[c].w0[k];
    cw1[k] = prm.cw[c].w1[k];
  }
  const uint32_t cadd = prm.cw[c].add, csel = prm.cw[c].sel;
  const uint32_t xm0 = prm.cw[c].xm[0], xm1 = prm.cw[c].xm[1], xm2 = prm.cw[c].xm[2];
  const bool has_xm = prm.cw[c].has_xm != 0;
  // ---- low-res columns of both blocks, packed per lane (lo = A, hi = B)
  uint32_t lf[9], rt[9];
  nine2(__byte_perm(top, 0u, 0x4140), __byte_perm(bot, 0u, 0x4140), lf);  // left columns: corners u | u+1
  nine2(__byte_perm(top, 0u, 0x4241), __byte_perm(bot, 0u, 0x4241), rt);  // right columns: u+1 | u+2
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
}
__device__ __forceinline__ void fwd3_cols(uint32_t (&x)[64]) {  // bias 2048 -> 16384
#pragma unroll
  for (int q = 0; q < 8; ++q)
    wht8p<2048>(x[q], x[8 + q], x[16 + q], x[24 + q], x[32 + q], x[40 + q], x[48 + q], x[56 + q]);
}

// grid (ceil(tiles per image / tiles_per_cta), 1, n), block 256: a CTA walks over tiles_per_cta
// consecutive tiles of one image.  COLS = blocks per row when known at compile time (0: run-time value).
// dynamic smem: signed map LUT window | tile (256 pairs x 8 pixel rows x 16*NCH bytes, piece by piece).
//
// Pipeline: the pixel tile is only read in the row pass of each channel.  As soon as the row pass of a
// tile's LAST channel is over (one extra barrier), the bulk copies of the NEXT tile are issued into the
// same buffer and land while the column pass and the quantiser of that channel run.  LUT window and
// quantiser records are staged once per CTA.
// The 8 warps of a CTA pass through the channels together (one barrier per channel): with 2 CTAs
// per SM the live instruction window stays inside the instruction cache.
template <int NCH, bool YCBCR, int COLS, bool LOCK = false>
__global__ void __launch_bounds__(kFwd3Threads, 2)
    k_forward3(const uint8_t *__restrict__ pixels, const uint8_t *__restrict__ L, Geom g,
               const __grid_constant__ Fwd3Params prm, const uint8_t *__restrict__ slut,
               uint8_t *__restrict__ planes, int tiles_per_cta) {
  extern __shared__ __align__(128) uint8_t lut[];  // signed map LUT first
  __shared__ uint64_t bar;
  __shared__ uint4 sRec[2][64];
  constexpr uint32_t PB = 16 * NCH;  // bytes of one pixel row of a block pair
  uint8_t *tile = lut + ((2 * prm.lut_half + 1 + 127) & ~127);
  const int cols = COLS ? COLS : g.cols;
  const int PR = cols >> 1;
  const int total = g.rows * PR;
  const int tiles = (total + kFwd3Threads - 1) / kFwd3Threads;
  const int tile0 = blockIdx.x * tiles_per_cta, tile1 = min(tile0 + tiles_per_cta, tiles);
  const uint8_t *img = pixels + (size_t)blockIdx.z * g.img_bytes;

  // bulk copies of one tile, issued by the lanes of warp 0: every piece (part of a block row inside
  // the tile) is 8 copies; lanes take pieces in turn
  auto issue_tile = [&](int tl) {
    const int f0 = tl * kFwd3Threads, nact = min(kFwd3Threads, total - f0);
    if (threadIdx.x == 0) mbar_expect_tx(&bar, (uint32_t)nact * 8 * PB);
    __syncwarp();
    const int v_first = f0 / PR, v_last = (f0 + nact - 1) / PR;
    for (int r = v_first + (int)threadIdx.x; r <= v_last; r += 32) {
      const int lo = max(f0, r * PR), hi = min(f0 + nact, (r + 1) * PR);
      const uint32_t bytes = (uint32_t)(hi - lo) * PB;
      uint8_t *dst = tile + (size_t)(lo - f0) * 8 * PB;
      const uint8_t *src = img + ((size_t)(8 * r) * g.w + (size_t)(lo - r * PR) * 16) * NCH;
#pragma unroll 1
      for (int y = 0; y < 8; ++y) bulk_g2s(dst + y * bytes, src + (size_t)y * g.w * NCH, bytes, &bar);
    }
  };

  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x < 32) issue_tile(tile0);
  {  // the needed window of the signed LUT and the quantiser records, while the tile is in flight
    // lut_half is a multiple of 64 and the global table is padded: aligned 128-bit copies
    const int half = prm.lut_half, nv = (2 * half + 1 + 15) >> 4;
    const uint4 *src = reinterpret_cast<const uint4 *>(slut + (kLutCenter - half));
    for (int i = threadIdx.x; i < nv; i += kFwd3Threads) reinterpret_cast<uint4 *>(lut)[i] = __ldg(src + i);
    if (threadIdx.x < 128) {
      const QuantRec &q = prm.rec[threadIdx.x >> 6][threadIdx.x & 63];
      sRec[threadIdx.x >> 6][threadIdx.x & 63] = make_uint4(q.c2, q.tmask, q.s16, smem_u32(lut) + q.off);
    }
  }

#pragma unroll 1
  for (int tl = tile0; tl < tile1; ++tl) {
    // ---- this thread's pair: block row v, blocks u = 2p and u + 1
    const int f0 = tl * kFwd3Threads, nact = min(kFwd3Threads, total - f0);
    const bool active = (int)threadIdx.x < nact;
    const int f = f0 + (active ? (int)threadIdx.x : 0);
    const int v = f / PR, p = f - v * PR, u = 2 * p;
    const int lo_r = max(f0, v * PR), hi_r = min(f0 + nact, (v + 1) * PR);
    const uint32_t pitch = (uint32_t)(hi_r - lo_r) * PB;
    const uint8_t *trow = tile + (size_t)(lo_r - f0) * 8 * PB + (size_t)(f - lo_r) * PB;
    // low-res corners (u, u+1, u+2) x (v, v+1), edge clamped; u is even and cols is even
    uint32_t top[NCH], bot[NCH];
    {
      const int v2 = min(v + 1, g.rows - 1), u2 = min(u + 2, cols - 1);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const uint8_t *Lc = L + ((size_t)blockIdx.z * NCH + c) * g.rows * cols;
        const uint8_t *r0 = Lc + (size_t)v * cols, *r1 = Lc + (size_t)v2 * cols;
        top[c] = (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(r0 + u)) | ((uint32_t)__ldg(r0 + u2) << 16);
        bot[c] = (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(r1 + u)) | ((uint32_t)__ldg(r1 + u2) << 16);
      }
    }
    uint8_t *seg = planes + (size_t)blockIdx.z * g.planes_bytes + (size_t)v * g.seg + u;
    mbar_wait(&bar, (uint32_t)(tl - tile0) & 1u);

#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
      // LUT / records staged; keeps the CTA's warps in the same phase (instruction-cache locality)
      if (LOCK || (c == 0 && tl == tile0)) __syncthreads();
      uint32_t x[64];
      if (active) {
        // (registers cannot be indexed: the corner words of channel c are picked with selects)
        uint32_t tp = top[0], bt = bot[0];
#pragma unroll
        for (int k = 1; k < NCH; ++k) {
          tp = c == k ? top[k] : tp;
          bt = c == k ? bot[k] : bt;
        }
        fwd3_rows<NCH>(prm, c, trow, pitch, tp, bt, x);
      }
      if (c == NCH - 1 && tl + 1 < tile1) {  // the tile has been read for the last time: refill it
        __syncthreads();
        if (threadIdx.x < 32) issue_tile(tl + 1);
      }
      if (active) {
        fwd3_cols(x);
        const int cls = (YCBCR && NCH >= 3 && (c == 1 || c == 2)) ? 1 : 0;
        quant3_store<COLS>(x, sRec[cls], seg + (size_t)c * cols * 64, (uint32_t)cols, std::make_integer_sequence<int, 64>{});
      }
    }
  }
}

}  // namespace himgcu

#endif  // HIMG_B200_XFORM_FWD3_CUH_
