// EXPERIMENT, NOT BUILT (round 2): measured 9.09 ms per 512 1080p images against 3.97 + 1.98 ms for
// k_dec_stream_par + k_inverse4, parity green on the whole GPU suite.  Kept for the
// record.  To rebuild it: include it from csrc/api.cu after xform_inv4.cuh, give inv4_compute a third
// template parameter GUARD that masks the tile stores (and the int32 redo) of threads with !active, and
// launch k_fres_fused<3, 240, 128> / <3, 480, 256> / <1, 1024, 512> with grid (rows, n) and
// fused_smem_bytes<NCH, COLS>() of dynamic shared memory in place of k_dec_stream_fres + stage_inverse.
//
// Fused FRES segment decode + inverse transform (the decoder's batch fast path).
//
// The reference never stores coefficient planes: DecodeFullResBlockRow (decoder.cpp:331-426) decodes
// one Huffman segment = one block row of all channels and immediately gathers / dequantises /
// inverse-transforms it.  Here a CTA does the same with the segment in SHARED memory:
//
//   phase D  the CTA's threads decode the segment's bit stream as ONE team (the subsequence-parallel,
//            self-synchronising scheme of k_dec_stream_par: guessed starts, rounds of "my start = the
//            previous end", prefix sum of the byte counts, second pass that writes) -- but the output
//            is the shared-memory tile: the zero fill is a few STS.128 per thread and a literal is ONE
//            byte store, where the stand-alone kernel needs zero-filled DRAM and line buffers to keep
//            the L2 write transactions down;
//   phase I  the tile is exactly K-inv's tile (rows = channel-major scan positions, COLS bytes each), so
//            the threads run inv4_compute on it in place and store pixels.
//
// The coefficient planes never exist in DRAM (no 2 * nch bytes per pixel of write + read between two
// kernels), and the decoder tables (16 KiB multi-token LUT + tree nodes) share their shared memory
// with the dequantisation tables of phase I.
//
// One instance per (channels, blocks per row): the tile pitch is the segment's own row length, so a
// decoded byte's position in the segment IS its tile address and K-inv's gathers keep immediate offsets.
// Shapes without an instance, single images (their low-res and coefficient branches run on two streams
// and only meet at the inverse) and force_generic take k_dec_stream_par + k_inverse4.
#ifndef HIMG_B200_FRES_FUSED_CUH_
#define HIMG_B200_FRES_FUSED_CUH_

#include "huff_dec_kernels.cuh"
#include "xform_inv4.cuh"

namespace himgcu {

// Decodes one stream of `nbytes` bytes at `src` into out_seg bytes at `o` (SHARED memory, 16-byte
// aligned, out_seg % 16 == 0) with the whole CTA as the team.  lut2 / SN: the image's multi-token LUT
// and packed tree nodes in shared memory.  s_end: blockDim.x words, ws: 33 words, s_flags: 3 ints.
// Returns (uniformly) whether the stream was sound: complete output, read position inside the last
// byte (huffman_dec.cpp:274-418, BitStream::AtTheEnd).  Ends with a barrier.
__device__ __forceinline__ bool dec_stream_smem(const uint8_t *__restrict__ src, uint32_t nbytes, const DecTree *__restrict__ T,
                                                const uint2 *lut2, const uint32_t *SN, uint8_t *o, int out_seg,
                                                uint32_t *s_end, uint32_t *ws, int *s_flags) {
  const int t = threadIdx.x, team = blockDim.x;
  int &s_changed = s_flags[0], &s_bad = s_flags[1], &s_final = s_flags[2];
  const uint32_t *lut = T->lut;      // single-token LUT and second-level tables: rare paths, global memory
  const uint32_t *sub_tab = T->sub;
  if (t == 0) s_bad = 0, s_final = -1;
  for (int i = t; i < (out_seg >> 4); i += team) reinterpret_cast<uint4 *>(o)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const uint32_t total_bits = nbytes * 8u;

  if (T->single) {  // single-leaf tree: every token has the same code, nothing to parallelise
    if (t == 0) {
      PBits br;
      br.seek(src, nbytes, 0);
      int n = 0;
      bool ok = true;
      while (n < out_seg) {
        int lit;
        const int z = decode_token(br, lut, SN, &lit);
        if (z < 0 || br.pos > total_bits || n + z > out_seg) {
          ok = false;
          break;
        }
        if (lit) o[n] = (uint8_t)lit;
        n += z;
      }
      if (ok) ok = br.pos > 8u * (nbytes - 1) && br.pos <= total_bits;
      if (!ok) s_bad = 1;
    }
    __syncthreads();
    return s_bad == 0;
  }

  // ---- phase 1: synchronise.  Thread t owns the tokens that START in [start_t, end_t).
  uint32_t sub = (total_bits + team - 1) / team;
  sub = max((sub + 31u) & ~31u, 128u);
  const uint32_t bound_lo = min((uint32_t)t * sub, total_bits);
  const uint32_t bound_hi = min((uint32_t)(t + 1) * sub, total_bits);
  const bool has_work = bound_lo < total_bits;
  uint32_t start = bound_lo, endpos = kPosInvalid, count = 0;
  bool dirty = has_work;
  constexpr int kCp = 4;  // checkpoints of the previous decode (see k_dec_stream_par)
  uint32_t cpp[kCp], cpc[kCp];
#pragma unroll
  for (int i = 0; i < kCp; ++i) cpp[i] = kPosInvalid, cpc[i] = 0;
  for (int round = 0; round <= team; ++round) {
    if (dirty) {
      PBits br;
      br.seek(src, nbytes, start);
      uint32_t cnt = 0, next_ms = bound_lo + 64u;
      int mi = 0;
      bool ok = true, spliced = false;
      while (br.pos < bound_hi) {
        if (br.pos >= next_ms) {
          uint32_t old_pos = kPosInvalid, old_cnt = 0;
#pragma unroll
          for (int i = 0; i < kCp; ++i)
            if (i == mi) {
              old_pos = cpp[i];
              old_cnt = cpc[i];
              cpp[i] = br.pos;
              cpc[i] = cnt;
            }
          if (old_pos == br.pos) {
            const uint32_t delta = cnt - old_cnt;
#pragma unroll
            for (int i = 0; i < kCp; ++i)
              if (i > mi) cpc[i] += delta;
            count += delta;
            spliced = true;
            break;
          }
          ++mi;
          next_ms = mi < kCp ? bound_lo + (64u << mi) : 0xffffffffu;
        }
        br.refill();
        if (br.pos + kLutBits + 14 <= total_bits) {
          const uint32_t g = lut2[(uint32_t)br.buf & (kLutSize - 1)].x;
          if (g & 15u) {
            const int nb = (int)(g & 15u), nx = (int)((g >> 14) & 15u);
            cnt += ((g >> 4) & 1023u) + ((uint32_t)(br.buf >> nb) & ((1u << nx) - 1u));
            br.consume(nb + nx);
            continue;
          }
          if (g & kLutLong) {
            int lit;
            const int z = decode_long2(br, SN, sub_tab, g, &lit);
            if (z < 0 || br.pos > total_bits) {
              ok = false;
              break;
            }
            cnt += (uint32_t)z;
            continue;
          }
        }
        int lit;
        const int z = decode_token(br, lut, SN, &lit);
        if (z < 0 || br.pos > total_bits) {
          ok = false;
          break;
        }
        cnt += (uint32_t)z;
      }
      if (!spliced) {
        count = cnt;
        endpos = ok ? br.pos : (bound_hi == total_bits ? total_bits : kPosInvalid);
#pragma unroll
        for (int i = 0; i < kCp; ++i)
          if (i >= mi) cpp[i] = kPosInvalid;
      }
    }
    s_end[t] = has_work ? endpos : kPosInvalid;
    if (t == 0) s_changed = 0;
    __syncthreads();
    dirty = false;
    if (has_work && t > 0) {
      const uint32_t prev = s_end[t - 1];
      if (prev != kPosInvalid && prev != start) {
        start = prev;
        dirty = start < bound_hi || bound_hi == total_bits;
        if (!dirty) {
          count = 0;
          endpos = start;
#pragma unroll
          for (int i = 0; i < kCp; ++i) cpp[i] = kPosInvalid;
        }
        s_changed = 1;
      }
    }
    __syncthreads();
    const bool again = s_changed != 0;
    __syncthreads();
    if (!again) break;
  }

  // ---- output offsets
  uint32_t total;
  const uint32_t off = block_exscan_u32(has_work ? count : 0u, ws, &total);
  if (has_work && off < (uint32_t)out_seg && endpos == kPosInvalid) s_bad = 1;
  __syncthreads();
  if (s_bad) return false;  // (uniform)

  // ---- phase 2: decode again; a literal is one byte store into the tile
  if (has_work && off < (uint32_t)out_seg && start < total_bits) {
    PBits br;
    br.seek(src, nbytes, start);
    int n = (int)off;
    bool ok = true;
    while (br.pos < bound_hi && n < out_seg) {
      br.refill();
      uint32_t long_g = 0;
      if (br.pos + kLutBits + 14 <= total_bits) {
        const uint2 g = lut2[(uint32_t)br.buf & (kLutSize - 1)];
        const int nb = (int)(g.x & 15u), nx = (int)((g.x >> 14) & 15u);
        const int gb = (int)((g.x >> 4) & 1023u) + (int)((uint32_t)(br.buf >> nb) & ((1u << nx) - 1u));
        if (nb && n + gb <= out_seg) {
          if (g.y >> 26) {
            o[n + (int)((g.x >> 18) & 1023u)] = (uint8_t)g.y;
            if ((g.y >> 26) > 1) o[n + (int)((g.y >> 8) & 1023u)] = (uint8_t)(g.y >> 18);
          }
          br.consume(nb + nx);
          n += gb;
          if (n == out_seg) s_final = (int)br.pos;
          continue;
        }
        if (!nb && (g.x & kLutLong)) long_g = g.x;
      }
      int lit;
      const int z = long_g ? decode_long2(br, SN, sub_tab, long_g, &lit) : decode_token(br, lut, SN, &lit);
      if (z < 0 || br.pos > total_bits || n + z > out_seg) {
        ok = false;
        break;
      }
      if (lit) o[n] = (uint8_t)lit;
      n += z;
      if (n == out_seg) s_final = (int)br.pos;
    }
    if (!ok) s_bad = 1;
  }
  __syncthreads();
  return !s_bad && s_final >= 0 && (uint32_t)s_final > 8u * (nbytes - 1) && (uint32_t)s_final <= total_bits;
}

constexpr int fused_min_ctas(int tp) { return tp <= 128 ? 3 : tp <= 256 ? 2 : 1; }
template <int NCH, int COLS>
constexpr int fused_smem_bytes() {
  return NCH * 64 * COLS + kLutSize * 8 + (kMaxNodes + 1) * 4;  // (the tables of phase I fit inside the decoder's)
}
static_assert(kLutSize * 8 >= kInvTableBytes, "the dequantisation tables reuse the LUT's shared memory");

// grid (rows, n), block TP >= COLS / 2 (a thread per block pair of the row; all threads decode).
// dynamic smem: tile [NCH * 64][COLS] | multi-token LUT [2048] x 8 | tree nodes   (phase I: dq tables over the LUT)
template <int NCH, int COLS, int TP>
__global__ void __launch_bounds__(TP, fused_min_ctas(TP))
    k_fres_fused(const uint8_t *__restrict__ data, const ChunkDesc *__restrict__ cd, const DecTree *__restrict__ trees,
                 const SegRef *__restrict__ segs, int nseg, const uint8_t *__restrict__ R, Geom g,
                 const InvTables *__restrict__ tabs, unsigned long long tab_stride, uint8_t *__restrict__ pixels,
                 int *__restrict__ status, uint32_t one) {
  extern __shared__ __align__(128) uint8_t sm[];
  constexpr int kTileBytes = NCH * 64 * COLS, PR = COLS / 2;
  static_assert(TP >= PR && TP % 32 == 0 && COLS % 16 == 0, "a thread per block pair");
  uint8_t *tile = sm;
  uint2 *lut2 = reinterpret_cast<uint2 *>(sm + kTileBytes);
  uint32_t *s_nodes = reinterpret_cast<uint32_t *>(sm + kTileBytes + kLutSize * 8);
  uint8_t *sDq = sm + kTileBytes;
  uint32_t *sTab = reinterpret_cast<uint32_t *>(sDq + kInvSlots * 256 * 2);
  __shared__ uint32_t s_end[TP];
  __shared__ uint32_t ws[33];
  __shared__ int s_flags[3];

  const int item = blockIdx.y, b = blockIdx.x, t = threadIdx.x;
  const DecTree *T = trees + item;
  const SegRef sr = segs[(size_t)item * nseg + b];
  bool ok = T->ok && sr.size != 0xffffffffu && sr.size != 0;  // (uniform)
  if (ok) {
    const uint4 *l2 = reinterpret_cast<const uint4 *>(T->lut2);
#pragma unroll 4
    for (int i = t; i < kLutSize / 2; i += TP) reinterpret_cast<uint4 *>(lut2)[i] = __ldg(l2 + i);
    const int nn = T->nnodes;
    for (int i = t; i < nn; i += TP) s_nodes[i] = T->nodes[i];
    __syncthreads();
    ok = dec_stream_smem(data + cd[item].off + sr.off, sr.size, T, lut2, s_nodes, tile, kTileBytes, s_end, ws, s_flags);
  }
  if (!ok && t == 0) atomicMax(&status[item], 1);  // (the pixels of a damaged segment are unspecified)
  __syncthreads();  // the decoder's tables are dead from here on

  // ---- phase I: K-inv on the tile
  const InvTables *IT = reinterpret_cast<const InvTables *>(reinterpret_cast<const char *>(tabs) + (size_t)item * tab_stride);
  {
    const uint4 *tsrc = reinterpret_cast<const uint4 *>(IT);
    for (int i = t; i < kInvTableBytes / 16; i += TP) cp_async16(sDq + 16 * i, tsrc + i);
  }
  const int pre = IT->pre, tab_bias = IT->bias;
  const bool overflow = IT->overflow != 0;
  const bool ycbcr = IT->ycbcr != 0;
  const bool active = t < PR;
  const int v = b, u = active ? 2 * t : 0;
  uint32_t top[NCH], bot[NCH];
  inv4_corners<NCH>(R, item, g, v, u, top, bot);
  cp_async_wait_all();
  __syncthreads();
  uint8_t *img = pixels + (size_t)item * g.out_img_bytes;
  // (threads beyond the row's pairs run along on column 0 -- the warp votes inside -- with their stores masked)
  inv4_compute<NCH, COLS, true>(tile + u, sDq, sTab, pre, tab_bias, overflow, ycbcr, active, top, bot,
                                img + ((size_t)(8 * v) * g.w + (size_t)u * 8) * NCH, (size_t)g.w * NCH, one);
}

}  // namespace himgcu

#endif  // HIMG_B200_FRES_FUSED_CUH_
