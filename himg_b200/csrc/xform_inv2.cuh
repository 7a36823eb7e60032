// K-inv, one block per thread in int32 with staged planes (v2): the path of 4-channel images and of
// 1- / 3-channel images whose buffers miss the alignment of the lane-pair kernel (xform_inv4.cuh).
// Same arithmetic as k_inverse (xform_kernels.cuh): gather + dequantise +
// inverse WHT with floor >>3 after each pass + low-res add + clamp + inverse colour map
// (decoder.cpp:366-423).  Differences are structural:
//
//  * the coefficient planes of a tile (nch x 64 scan positions x <=256 blocks) are staged in shared
//    memory with cp.async 16-byte copies (fully coalesced, asynchronous); the pitch is a compile
//    time 256 bytes, so every gather is an LDS with an immediate offset and no address arithmetic;
//  * 256-thread CTAs whose warps pass the per-channel phases in lock-step (instruction cache);
//  * per-image dequantisation tables (they travel in-band) live in shared memory.
//
// (The row pass sums eight int16 values before its floor shift: 19 bits.  k_inverse4 packs two blocks
// per register anyway, guarded by a range vote.)
//
// Preconditions (host checked, else k_inverse): width % 128 == 0 or cols % 16 == 0, height % 8 == 0,
// 16-byte aligned planes / pixels, nch in {1,3,4}.
#ifndef HIMG_B200_XFORM_INV2_CUH_
#define HIMG_B200_XFORM_INV2_CUH_

#include "common.cuh"
#include "xform_lane.cuh"  // cp_async16

namespace himgcu {

constexpr int kInv2Threads = 256;
constexpr int kInv2Pitch = 256;  // bytes per staged plane row = max blocks per tile

// grid (ceil(cols/tile_cols), rows, n), block 256.
// dynamic smem: planes tile [NCH*64][256] | sOut [NCH][256][17] words
template <int NCH>
__global__ void __launch_bounds__(kInv2Threads, 2)
    k_inverse2(const uint8_t *__restrict__ planes, const uint8_t *__restrict__ R, Geom g,
               const DecTables *__restrict__ tabs, unsigned long long tab_stride, int tile_cols,
               uint8_t *__restrict__ pixels) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ int16_t sUn[256];
  __shared__ uint8_t sShift[2][64];
  uint8_t *sPl = smem;
  uint32_t(*sOut)[kInv2Threads][17] = reinterpret_cast<uint32_t(*)[kInv2Threads][17]>(smem + NCH * 64 * kInv2Pitch);

  const int v = blockIdx.y, u0 = blockIdx.x * tile_cols, t = threadIdx.x;
  const int nblk = min(tile_cols, g.cols - u0);  // multiple of 16
  const uint8_t *seg = planes + (size_t)blockIdx.z * g.planes_bytes + (size_t)v * g.seg + u0;
  {
    // one warp per plane row, one 16-byte chunk per lane
    const int lane = t & 31, w = t >> 5, nchunk = nblk >> 4;
    if (lane < nchunk)
      for (int r = w; r < NCH * 64; r += kInv2Threads / 32)
        cp_async16(sPl + r * kInv2Pitch + lane * 16, seg + (size_t)r * g.cols + lane * 16);
  }
  const DecTables *T = reinterpret_cast<const DecTables *>(reinterpret_cast<const char *>(tabs) + (size_t)blockIdx.z * tab_stride);
  sUn[t] = T->full_unmap[t];
  if (t < 128) sShift[t >> 6][t & 63] = T->shift[t >> 6][t & 63];
  const bool ycbcr = T->ycbcr != 0;
  const bool active = t < nblk;
  const int u = u0 + t;
  cp_async_wait_all();

#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    __syncthreads();  // staged data visible; warps stay in the same phase
    if (active) {
      int x[64];
      const uint8_t *src = sPl + c * 64 * kInv2Pitch + t;
      const uint8_t *sh = sShift[(ycbcr && NCH >= 3 && (c == 1 || c == 2)) ? 1 : 0];
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        const int val = sUn[src[scan_pos(j) * kInv2Pitch]];
        x[j] = (int)(short)(val << sh[j]);  // (int16)(Unmap << shift), quantize.cpp:163
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        wht8(x[r * 8 + 0], x[r * 8 + 1], x[r * 8 + 2], x[r * 8 + 3], x[r * 8 + 4], x[r * 8 + 5], x[r * 8 + 6], x[r * 8 + 7]);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[r * 8 + i] >>= 3;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        wht8(x[q], x[8 + q], x[16 + q], x[24 + q], x[32 + q], x[40 + q], x[48 + q], x[56 + q]);
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i * 8 + q] >>= 3;
      }
      {
        const uint8_t *Rc = R + ((size_t)blockIdx.z * NCH + c) * g.rows * g.cols;
        const int v2 = min(v + 1, g.rows - 1), u2 = min(u + 1, g.cols - 1);
        int lf[9], rt[9];
        nine(__ldg(Rc + v * g.cols + u), __ldg(Rc + v2 * g.cols + u), lf);
        nine(__ldg(Rc + v * g.cols + u2), __ldg(Rc + v2 * g.cols + u2), rt);
#pragma unroll
        for (int y = 0; y < 8; ++y) {
          int tt[9];
          nine(lf[y], rt[y], tt);
#pragma unroll
          for (int i = 0; i < 8; ++i) x[y * 8 + i] = __vimin_s32_relu((int)(short)(x[y * 8 + i] + tt[i]), 255);
        }
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint32_t lo = __byte_perm((uint32_t)x[4 * k], (uint32_t)x[4 * k + 1], 0x0040);
        const uint32_t hi = __byte_perm((uint32_t)x[4 * k + 2], (uint32_t)x[4 * k + 3], 0x0040);
        sOut[c][t][k] = __byte_perm(lo, hi, 0x5410);
      }
    }
  }
  __syncthreads();
  if (!active) return;

  // ---- inverse colour map + interleave + 8-byte aligned stores
  uint8_t *img = pixels + (size_t)blockIdx.z * g.out_img_bytes;
  const bool do_colour = ycbcr && NCH >= 3;
#pragma unroll
  for (int y = 0; y < 8; ++y) {
    uint32_t by[8 * NCH];  // the row's bytes in memory order
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t wch[NCH];
#pragma unroll
      for (int c = 0; c < NCH; ++c) wch[c] = sOut[c][t][y * 2 + h];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = 4 * h + q;
        int ch[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) ch[c] = (int)__byte_perm(wch[c], 0u, 0x4440u + q);
        if (do_colour) {
          const int Y = ch[0], cb = 2 * ch[1] - 255, cr = 2 * ch[NCH >= 3 ? 2 : 0] - 255;
          const int G = Y - ((cb + cr + 2) >> 2);
          ch[0] = __viaddmin_s32_relu(G, cr, 255);
          ch[1] = __vimin_s32_relu(G, 255);
          ch[NCH >= 3 ? 2 : 0] = __viaddmin_s32_relu(G, cb, 255);
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) by[i * NCH + c] = (uint32_t)ch[c];
      }
    }
    uint2 *q = reinterpret_cast<uint2 *>(img + ((size_t)(8 * v + y) * g.w + (size_t)u * 8) * NCH);
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      uint32_t w2[2];
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int b0 = (2 * k + m) * 4;
        const uint32_t lo = __byte_perm(by[b0], by[b0 + 1], 0x0040), hi = __byte_perm(by[b0 + 2], by[b0 + 3], 0x0040);
        w2[m] = __byte_perm(lo, hi, 0x5410);
      }
      q[k] = make_uint2(w2[0], w2[1]);
    }
  }
}

}  // namespace himgcu

#endif  // HIMG_B200_XFORM_INV2_CUH_
