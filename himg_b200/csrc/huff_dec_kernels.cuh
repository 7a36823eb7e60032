// RIFF container parsing + RLE/Huffman DECODE kernels (reference: decoder.cpp:144-272, :428-461,
// huffman_dec.cpp:152-418).  Every read is bounds-checked: hostile streams are rejected, never
// over-read (the reference's unchecked fast loop has undefined behaviour there).
//
//   k_dec_parse    one thread per image: RIFF/FRMT/LMAP/LRES/QCFG/FMAP/FRES walk + table recovery
//   k_dec_tree     one CTA per chunk: serialised tree -> node arrays + 10-bit LUT
//   k_dec_segtab   one thread per chunk: walk the 2/4-byte segment headers
//   k_dec_stream   one thread per segment: LUT decode + tree walk for long codes + zero-run
//                  expansion, 32-bit stores (the format's segments are independently decodable)
#ifndef HIMG_B200_HUFF_DEC_KERNELS_CUH_
#define HIMG_B200_HUFF_DEC_KERNELS_CUH_

#include "common.cuh"

namespace himgcu {

constexpr int kLutBits = 10;
constexpr int kLutSize = 1 << kLutBits;

struct ChunkDesc {
  unsigned long long off;  // byte offset of the chunk payload inside the data buffer
  uint32_t size;
  uint32_t ok;
};

struct DecTree {
  uint32_t lut[kLutSize];  // leaf: sym | len << 16;  long code: 0x80000000 | node at depth 10
  short ca[kMaxNodes], cb[kMaxNodes], sym[kMaxNodes];
  short pad;
  int nnodes;
  int data_off;  // first byte after the (byte-aligned) tree
  int ok;
  int single;    // single-leaf tree
};

struct SegRef {
  uint32_t off;   // relative to the chunk payload
  uint32_t size;  // 0xffffffff = invalid
};

__device__ __forceinline__ uint32_t rd_u32(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// Mapping function bytes -> unmap table (mapper.cpp:127-157).
__device__ bool parse_mapfun(const uint8_t *in, int size, int16_t *unmap) {
  if (size < 1) return false;
  const int single = in[0];
  if (1 + single + 2 * (127 - single) != size) return false;
  const uint8_t *p = in + 1;
  unmap[0] = 0;
  for (int i = 1; i <= 127; ++i) {
    uint32_t v = *p++;
    if (i > single) v |= (uint32_t)(*p++) << 8;
    const short sv = (short)(unsigned short)v;
    unmap[i] = sv;
    unmap[256 - i] = (short)(-sv);
  }
  unmap[128] = unmap[129];
  return true;
}

// grid: ceil(n/128) x 128 threads; one thread per image.
__global__ void k_dec_parse(const uint8_t *__restrict__ data, const unsigned long long *__restrict__ offsets,
                            const uint32_t *__restrict__ sizes, int n, int w, int h, int nch,
                            ChunkDesc *__restrict__ lres, ChunkDesc *__restrict__ fres,
                            DecTables *__restrict__ tabs, int *__restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long base = offsets[i];
  const uint8_t *p = data + base;
  const long long size = sizes[i];
  lres[i].ok = fres[i].ok = 0;
  lres[i].off = fres[i].off = base;
  lres[i].size = fres[i].size = 0;
  status[i] = 1;  // reject until proven otherwise
  DecTables *T = tabs + i;
  T->ycbcr = 0;
  if (size < 12 || rd_u32(p) != 0x46464952u /*RIFF*/ || rd_u32(p + 8) != 0x474d4948u /*HIMG*/) return;
  if ((long long)(int)rd_u32(p + 4) + 8 != size) return;
  long long idx = 12;
  const uint32_t want[6] = {0x544d5246u /*FRMT*/, 0x50414d4cu /*LMAP*/, 0x5345524cu /*LRES*/,
                            0x47464351u /*QCFG*/, 0x50414d46u /*FMAP*/, 0x53455246u /*FRES*/};
  bool has_chroma = false;
  for (int k = 0; k < 6; ++k) {
    long long cs = -1;
    for (;;) {  // FindRIFFChunk: unknown chunks are skipped (decoder.cpp:445-461)
      if (idx + 8 > size) return;
      const uint32_t cc = rd_u32(p + idx);
      const int sz = (int)rd_u32(p + idx + 4);
      idx += 8;
      if (sz < 0 || idx + sz > size) return;
      if (cc == want[k]) {
        cs = sz;
        break;
      }
      idx += sz;
    }
    const uint8_t *c = p + idx;
    if (k == 0) {
      if (cs < 11 || c[0] != 1) return;
      if ((int)rd_u32(c + 1) != w || (int)rd_u32(c + 5) != h || (int)c[9] != nch) return;
      T->ycbcr = (c[10] != 0 && nch >= 3) ? 1 : 0;
      has_chroma = T->ycbcr != 0;
    } else if (k == 1) {
      if (!parse_mapfun(c, (int)cs, T->low_unmap)) return;
    } else if (k == 2) {
      lres[i].off = base + idx;
      lres[i].size = (uint32_t)cs;
      lres[i].ok = 1;
    } else if (k == 3) {
      if (cs != (has_chroma ? 64 : 32)) return;
      for (int q = 0; q < 32; ++q) {
        T->shift[0][2 * q] = c[q] >> 4;
        T->shift[0][2 * q + 1] = c[q] & 15;
        T->shift[1][2 * q] = has_chroma ? (c[32 + q] >> 4) : 0;
        T->shift[1][2 * q + 1] = has_chroma ? (c[32 + q] & 15) : 0;
      }
    } else if (k == 4) {
      if (!parse_mapfun(c, (int)cs, T->full_unmap)) return;
    } else {
      fres[i].off = base + idx;
      fres[i].size = (uint32_t)cs;
      fres[i].ok = 1;
    }
    idx += cs;
  }
  status[i] = 0;
}

// Chunk descriptors for the stage-level API (n equally strided chunks).
__global__ void k_dec_make_desc(unsigned long long stride, const uint32_t *__restrict__ sizes, int n,
                                ChunkDesc *__restrict__ cd, int *__restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  cd[i].off = (unsigned long long)i * stride;
  cd[i].size = sizes[i];
  cd[i].ok = 1;
  status[i] = 0;
}

// ---- tree recovery (huffman_dec.cpp:152-213) ---------------------------------------------------
constexpr int kDecTreeThreads = 256;

// grid (n).
__global__ void __launch_bounds__(kDecTreeThreads)
    k_dec_tree(const uint8_t *__restrict__ data, const ChunkDesc *__restrict__ cd, int lenient,
               DecTree *__restrict__ trees, int *__restrict__ status) {
  __shared__ uint8_t raw[kTreeBytesMax + 8];
  __shared__ short ca[kMaxNodes], cb[kMaxNodes], nsym[kMaxNodes];
  __shared__ uint8_t depth[kMaxNodes];
  __shared__ unsigned short pcode[kMaxNodes];
  __shared__ int s_n, s_ok, s_bits;
  __shared__ int stk_slot[kMaxNodes + 2];
  __shared__ uint8_t stk_depth[kMaxNodes + 2];
  __shared__ unsigned short stk_code[kMaxNodes + 2];
  const int item = blockIdx.x, t = threadIdx.x;
  DecTree *out = trees + item;
  const ChunkDesc d = cd[item];
  const int avail = d.ok ? (int)min((uint32_t)kTreeBytesMax, d.size) : 0;
  for (int i = t; i < kTreeBytesMax + 8; i += blockDim.x) raw[i] = i < avail ? data[d.off + i] : 0;
  for (int i = t; i < kLutSize; i += blockDim.x) out->lut[i] = 0;  // len 0 / sym 0: never advances
  __syncthreads();
  if (t == 0) {
    int n = 0, ok = d.ok ? 1 : 0, sp = 0, bit = 0;
    const int nbits = avail * 8;
    stk_slot[0] = -1;
    stk_depth[0] = 0;
    stk_code[0] = 0;
    sp = ok ? 1 : 0;
    while (sp && ok) {
      --sp;
      const int slot = stk_slot[sp], dep = stk_depth[sp];
      const unsigned short code = stk_code[sp];
      if (n >= kMaxNodes) {
        ok = 0;
        break;
      }
      const int k = n++;
      ca[k] = cb[k] = -1;
      nsym[k] = -1;
      depth[k] = (uint8_t)dep;
      pcode[k] = code;
      if (slot >= 0) {
        if (slot & 1) cb[slot >> 1] = (short)k;
        else ca[slot >> 1] = (short)k;
      }
      if (bit + 1 > nbits) {
        ok = 0;
        break;
      }
      const int leaf = (raw[bit >> 3] >> (bit & 7)) & 1;
      ++bit;
      if (leaf) {
        if (bit + 9 > nbits) {
          ok = 0;
          break;
        }
        int s = 0;
        for (int q = 0; q < 9; ++q, ++bit) s |= ((raw[bit >> 3] >> (bit & 7)) & 1) << q;
        nsym[k] = (short)s;
      } else {
        const int nd = min(dep + 1, 255);
        const unsigned short cbit = dep < kLutBits ? (unsigned short)(code | (1u << dep)) : code;
        stk_slot[sp] = k * 2 + 1;
        stk_depth[sp] = (uint8_t)nd;
        stk_code[sp] = cbit;
        ++sp;
        stk_slot[sp] = k * 2;
        stk_depth[sp] = (uint8_t)nd;
        stk_code[sp] = code;
        ++sp;
      }
    }
    s_n = n;
    s_ok = ok;
    s_bits = bit;
  }
  __syncthreads();
  const int n = s_n;
  if (!s_ok) {
    if (t == 0) {
      out->ok = 0;
      out->nnodes = 0;
      out->data_off = 0;
      out->single = 0;
      atomicMax(&status[item], 1);
    }
    return;
  }
  const bool single = n == 1;
  for (int k = t; k < n; k += blockDim.x) {
    out->ca[k] = ca[k];
    out->cb[k] = cb[k];
    out->sym[k] = nsym[k];
    const int dep = depth[k];
    if (nsym[k] >= 0) {
      if (dep <= kLutBits) {
        // The reference's decoder consumes ZERO bits per symbol for a single-leaf tree
        // (huffman_dec.cpp:178-188) although its encoder wrote one; lenient mode consumes it.
        const uint32_t len = (single && lenient) ? 1u : (uint32_t)dep;
        const uint32_t e = (uint32_t)nsym[k] | (len << 16);
        for (uint32_t i = 0; i < (1u << (kLutBits - dep)); ++i) out->lut[(i << dep) | pcode[k]] = e;
      }
    } else if (dep == kLutBits) {
      out->lut[pcode[k]] = 0x80000000u | (uint32_t)k;
    }
  }
  if (t == 0) {
    out->ok = 1;
    out->nnodes = n;
    out->data_off = (s_bits + 7) >> 3;
    out->single = single ? 1 : 0;
  }
}

// ---- segment table (huffman_dec.cpp:215-251) ---------------------------------------------------
// One thread per chunk.  mode: 0 = unframed chunk decoded as ONE stream (LRES, Uncompress);
// 1 = framed chunk of nseg block rows (FRES, UncompressBlock).
__global__ void k_dec_segtab(const uint8_t *__restrict__ data, const ChunkDesc *__restrict__ cd,
                             const DecTree *__restrict__ trees, int n, int nseg, int seg_size, int mode,
                             int lenient, SegRef *__restrict__ segs, int *__restrict__ status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  SegRef *S = segs + (size_t)i * nseg;
  for (int b = 0; b < nseg; ++b) S[b].size = 0xffffffffu, S[b].off = 0;
  const ChunkDesc d = cd[i];
  const DecTree *T = trees + i;
  if (!d.ok || !T->ok) {
    atomicMax(&status[i], 1);
    return;
  }
  const uint32_t start = (uint32_t)T->data_off;
  if (mode == 0) {
    // HuffmanDec(in, size, 0): block size = packed size => never framed; an empty payload can
    // only produce an empty output (huffman_dec.cpp:278-279).
    if (start >= d.size) {
      atomicMax(&status[i], 1);
      return;
    }
    S[0].off = start;
    S[0].size = d.size - start;
    return;
  }
  // The reference decides framing by comparing the UNPACKED block size with the PACKED chunk size
  // (huffman_dec.cpp:217-218); the encoder framed iff there is more than one segment.
  const bool framed = lenient ? (nseg > 1) : ((uint32_t)seg_size < d.size);
  if (!framed) {
    if (!lenient || start >= d.size) {  // UncompressBlock refuses unframed data (:265)
      atomicMax(&status[i], 1);
      return;
    }
    S[0].off = start;
    S[0].size = d.size - start;
    return;
  }
  const uint8_t *p = data + d.off;
  uint32_t pos = start;
  for (int b = 0; b < nseg; ++b) {
    if (pos + 2 > d.size) {
      atomicMax(&status[i], 1);
      return;
    }
    uint32_t sz = (uint32_t)p[pos] | ((uint32_t)p[pos + 1] << 8);
    pos += 2;
    if (sz & 0x8000u) {
      if (pos + 2 > d.size) {
        atomicMax(&status[i], 1);
        return;
      }
      sz = (sz & 0x7fffu) | (((uint32_t)p[pos] | ((uint32_t)p[pos + 1] << 8)) << 15);
      pos += 2;
    }
    if ((unsigned long long)pos + sz > d.size) {
      atomicMax(&status[i], 1);
      return;
    }
    S[b].off = pos;
    S[b].size = sz;
    pos += sz;
  }
}

// ---- stream decode (huffman_dec.cpp:274-418) ---------------------------------------------------
struct BitReader {
  const uint8_t *p;
  uint32_t nbytes, rd;
  uint64_t buf;
  int nb;
  uint32_t pos;  // bits consumed
  __device__ __forceinline__ void init(const uint8_t *ptr, uint32_t n) {
    p = ptr;
    nbytes = n;
    rd = 0;
    buf = 0;
    nb = 0;
    pos = 0;
    while (rd < nbytes && ((reinterpret_cast<uintptr_t>(p + rd)) & 3)) {
      buf |= (uint64_t)p[rd] << nb;
      nb += 8;
      ++rd;
    }
  }
  __device__ __forceinline__ void refill() {
    // aligned 32-bit loads; a word is fetched only if it holds at least one valid byte
    if (nb <= 32 && rd < nbytes) {
      const uint32_t w = *reinterpret_cast<const uint32_t *>(p + rd);
      buf |= (uint64_t)w << nb;
      nb += 32;
      rd += 4;
    }
  }
  __device__ __forceinline__ void consume(int n) {
    buf >>= n;
    nb -= n;
    pos += (uint32_t)n;
  }
};

constexpr int kDecThreads = 32;

// grid (ceil(nseg/32), n), block 32: one thread per segment.  Output segment b of item i goes to
// out + i*out_stride + b*out_seg (out_seg bytes).
__global__ void __launch_bounds__(kDecThreads)
    k_dec_stream(const uint8_t *__restrict__ data, const ChunkDesc *__restrict__ cd,
                 const DecTree *__restrict__ trees, const SegRef *__restrict__ segs, int nseg,
                 int out_seg, uint8_t *__restrict__ out, unsigned long long out_stride,
                 int *__restrict__ status) {
  __shared__ uint32_t lut[kLutSize];
  __shared__ short ca[kMaxNodes], cb[kMaxNodes], nsym[kMaxNodes];
  const int item = blockIdx.y, b = blockIdx.x * kDecThreads + threadIdx.x;
  const DecTree *T = trees + item;
  for (int i = threadIdx.x; i < kLutSize; i += kDecThreads) lut[i] = T->lut[i];
  const int nn = T->ok ? T->nnodes : 0;
  for (int i = threadIdx.x; i < nn; i += kDecThreads) {
    ca[i] = T->ca[i];
    cb[i] = T->cb[i];
    nsym[i] = T->sym[i];
  }
  __syncwarp();
  if (b >= nseg) return;
  const SegRef sr = segs[(size_t)item * nseg + b];
  if (!T->ok || sr.size == 0xffffffffu) {
    atomicMax(&status[item], 1);
    return;
  }
  uint8_t *o = out + (size_t)item * out_stride + (size_t)b * out_seg;
  uint32_t *ow = reinterpret_cast<uint32_t *>(o);  // segment bases are 4-byte aligned
  BitReader br;
  br.init(data + cd[item].off + sr.off, sr.size);
  const uint32_t total_bits = sr.size * 8u;
  int n = 0;
  uint32_t acc = 0;
  bool ok = true;
  while (n < out_seg) {
    br.refill();
    const uint32_t e = lut[(uint32_t)br.buf & (kLutSize - 1)];
    int sym;
    if (!(e & 0x80000000u)) {
      sym = (int)(e & 0xffffu);
      br.consume((int)(e >> 16));
    } else {
      int node = (int)(e & 0xffffu);
      br.consume(kLutBits);
      while (nsym[node] < 0 && br.pos <= total_bits) {
        br.refill();
        const int bit = (int)(br.buf & 1u);
        br.consume(1);
        node = bit ? cb[node] : ca[node];
      }
      sym = nsym[node];
    }
    if (br.pos > total_bits || sym < 0) {
      ok = false;
      break;
    }
    if (sym <= 255) {
      acc |= (uint32_t)sym << (8 * (n & 3));
      ++n;
      if ((n & 3) == 0) {
        ow[(n >> 2) - 1] = acc;
        acc = 0;
      }
    } else {
      int z;
      if (sym == 256) {
        z = 2;
      } else {
        const int nx = sym == 257 ? 2 : (sym == 258 ? 4 : (sym == 259 ? 8 : 14));
        const int addv = sym == 257 ? 3 : (sym == 258 ? 7 : (sym == 259 ? 23 : 279));
        if (sym > 260) {
          ok = false;
          break;
        }
        br.refill();
        z = (int)((uint32_t)br.buf & ((1u << nx) - 1u)) + addv;
        br.consume(nx);
        if (br.pos > total_bits) {
          ok = false;
          break;
        }
      }
      if (n + z > out_seg) {
        ok = false;
        break;
      }
      while (z && (n & 3)) {
        ++n;
        --z;
        if ((n & 3) == 0) {
          ow[(n >> 2) - 1] = acc;
          acc = 0;
        }
      }
      while (z >= 4) {
        ow[n >> 2] = 0;
        n += 4;
        z -= 4;
      }
      n += z;  // acc stays 0, the tail is flushed with the next bytes
    }
  }
  if (ok && (n & 3)) {
    for (int k = 0; k < (n & 3); ++k) o[(n & ~3) + k] = (uint8_t)(acc >> (8 * k));
  }
  // BitStream::AtTheEnd (huffman_dec.cpp:140-145): the read position must be inside the last byte
  // or exactly at the end.
  if (ok) ok = sr.size == 0 ? br.pos == 0 : (br.pos > 8u * (sr.size - 1) && br.pos <= total_bits);
  if (!ok) atomicMax(&status[item], 1);
}

}  // namespace himgcu

#endif  // HIMG_B200_HUFF_DEC_KERNELS_CUH_
