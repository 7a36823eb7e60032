// K-avg, fast path: window means of the colour-mapped pixels (ycbcr.cpp:24-52 + downsampled.cpp:75-94)
// on lane pairs.  The generic k_lowres_avg maps one byte at a time with scalar integer arithmetic (7.5
// thread instructions per sample, ALU pipe 82 % busy -- compute bound on a kernel that only has to
// stream the pixels once); here a thread owns TWO horizontally adjacent blocks, the colour mapping of a
// pixel pair is the dp4a + byte-permute of K-fwd (xform_lane.cuh: one code path for Y / Cb / Cr / plain
// channels, on the FMA pipe), and the partial sums are 16-bit lanes of one register (64 * 255 fits).
//
// Window of block u = pixel columns 8u-3 .. 8u+4: the thread adds its columns 0..4 (sum A) to the
// columns 5..7 of the block on the left (sum B): own low lane for the right block of the pair, the
// left neighbour thread's high lane (through shared memory) for the left block.
//
// Preconditions (host checked, else k_lowres_avg): width % 16 == 0, tightly packed pixels, 16-byte
// aligned image rows, nch in {1, 3, 4}.
#ifndef HIMG_B200_XFORM_AVG2_CUH_
#define HIMG_B200_XFORM_AVG2_CUH_

#include "common.cuh"
#include "xform_lane.cuh"

namespace himgcu {

constexpr int kAvg2Threads = 128;  // block pairs per CTA

struct Avg2Params {
  ColourW cw[4];
};

// Sums of one 16-pixel row of a block pair: A += mapped pixels 0..4, B += mapped pixels 5..7 (FROM = 5
// skips the A part: the neighbour on the left only contributes its B).
template <int NCH, int FROM>
__device__ __forceinline__ void avg2_row(const Avg2Params &prm, const uint8_t *row, uint32_t (&A)[NCH], uint32_t (&B)[NCH]) {
  const uint4 *rp = reinterpret_cast<const uint4 *>(row);
  uint32_t w[4 * NCH], wx[4 * NCH];
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const uint4 q = __ldg(rp + k);
    w[4 * k] = q.x;
    w[4 * k + 1] = q.y;
    w[4 * k + 2] = q.z;
    w[4 * k + 3] = q.w;
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const ColourW &cw = prm.cw[c];
    const bool has_xm = NCH > 1 && cw.has_xm != 0;
#pragma unroll
    for (int k = 0; k < 4 * NCH; ++k) wx[k] = has_xm ? w[k] ^ ((k % 3 == 0) ? cw.xm[0] : (k % 3 == 1) ? cw.xm[1] : cw.xm[2]) : w[k];
    const uint32_t *wa = wx, *wb = wx + 2 * NCH;
#pragma unroll
    for (int i = FROM; i < 8; ++i) {
      const uint32_t val = colour_pair<NCH>(wa, wb, i, cw.w0, cw.w1, cw.add, cw.sel);
      if (i < 5) A[c] += val;
      else B[c] += val;
    }
  }
}

// grid (ceil(cols / 2 / 128), rows, n), block 128.  avg: [n][nch][rows][cols].
template <int NCH>
__global__ void __launch_bounds__(kAvg2Threads)
    k_lowres_avg2(const uint8_t *__restrict__ pixels, Geom g, const __grid_constant__ Avg2Params prm, uint8_t *__restrict__ avg) {
  __shared__ uint32_t sB[NCH][kAvg2Threads + 1];
  const int v = blockIdx.y, p0 = blockIdx.x * kAvg2Threads, p = p0 + threadIdx.x, PR = g.cols >> 1;
  const uint8_t *img = pixels + (size_t)blockIdx.z * g.img_bytes;
  const int y0 = max(0, 8 * v - 3), y1 = min(g.h - 1, 8 * v + 4);
  const size_t row_bytes = (size_t)g.w * NCH;
  uint32_t A[NCH], B[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) A[c] = B[c] = 0;
  if (p < PR) {
    const uint8_t *row = img + (size_t)y0 * row_bytes + (size_t)p * 16 * NCH;
    for (int y = y0; y <= y1; ++y, row += row_bytes) avg2_row<NCH, 0>(prm, row, A, B);
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) sB[c][threadIdx.x + 1] = B[c];
  if (threadIdx.x == 0) {  // the pair on the left of this tile (its right block's last three columns)
    uint32_t Al[NCH], Bl[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) Al[c] = Bl[c] = 0;
    if (p0 > 0) {
      const uint8_t *row = img + (size_t)y0 * row_bytes + (size_t)(p0 - 1) * 16 * NCH;
      for (int y = y0; y <= y1; ++y, row += row_bytes) avg2_row<NCH, 5>(prm, row, Al, Bl);
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) sB[c][0] = Bl[c];
  }
  __syncthreads();
  if (p < PR) {
    const int ny = y1 - y0 + 1;
    const int cnt_l = (p == 0 ? 5 : 8) * ny, cnt_r = 8 * ny;  // (width % 16 == 0: only block 0 is clipped, on the left)
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      // left block: own A (low lane) + B of the block on its left (high lane of the left neighbour)
      const int sl = (int)(A[c] & 0xffffu) + (int)(sB[c][threadIdx.x] >> 16);
      const int sr = (int)(A[c] >> 16) + (int)(B[c] & 0xffffu);
      const uint32_t out = (uint32_t)((sl + (cnt_l >> 1)) / cnt_l) | ((uint32_t)((sr + (cnt_r >> 1)) / cnt_r) << 8);
      *reinterpret_cast<uint16_t *>(avg + (((size_t)blockIdx.z * NCH + c) * g.rows + v) * g.cols + 2 * p) = (uint16_t)out;
    }
  }
}

}  // namespace himgcu

#endif  // HIMG_B200_XFORM_AVG2_CUH_
