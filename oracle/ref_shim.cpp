// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Thin extern "C" wrapper around the UNMODIFIED reference implementation.  This file is
// compiled together with the reference's own sources *where they lie* under
// /root/reference/src/lib (see oracle/build_oracle.py); only the resulting shared object
// lands in oracle/_ref/.  No reference source is copied into this repository.
//
// Entry points wrapped (reference file:line):
//   himg::Encoder::Encode            src/lib/encoder.cpp:59-109
//   himg::Decoder::Decode            src/lib/decoder.cpp:87-138
//   himg::HuffmanEnc::Compress       src/lib/huffman_enc.cpp:246-363
//   himg::HuffmanDec::{Init,Uncompress,UncompressBlock}  src/lib/huffman_dec.cpp:221-272
//   himg::Downsampled::{SampleImage,GetBlockData,SetBlockData,GetLowresBlock}
//                                    src/lib/downsampled.cpp:67-382
//   himg::Quantize / Mapper / Hadamard / YCbCr            (stage-level checks)
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <vector>

#include "decoder.h"
#include "downsampled.h"
#include "encoder.h"
#include "hadamard.h"
#include "huffman_dec.h"
#include "huffman_enc.h"
#include "mapper.h"
#include "quantize.h"
#include "ycbcr.h"

namespace {
// The reference prints progress text on std::cout; keep test logs quiet.
struct MuteCout {
  std::streambuf* old;
  std::ostringstream sink;
  MuteCout() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~MuteCout() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

// Returns packed size (>0) or -1 if out_cap is too small.  A fresh Encoder per call
// (object reuse is broken in the reference).
int ref_encode(const uint8_t* data, int w, int h, int pixel_stride, int nch, int quality,
               int use_ycbcr, uint8_t* out, int out_cap) {
  MuteCout mute;
  himg::Encoder enc;
  enc.Encode(data, w, h, pixel_stride, nch, quality, use_ycbcr != 0);
  if (enc.packed_size() > out_cap) return -1;
  std::memcpy(out, enc.packed_data(), enc.packed_size());
  return enc.packed_size();
}

// Returns 1 on success, 0 if the reference decoder returned false, -1 if out_cap too small.
int ref_decode(const uint8_t* packed, int size, int max_threads, uint8_t* out, int out_cap,
               int* w, int* h, int* nch) {
  MuteCout mute;
  himg::Decoder dec(max_threads);
  if (!dec.Decode(packed, size)) return 0;
  *w = dec.width();
  *h = dec.height();
  *nch = dec.num_channels();
  if (dec.unpacked_size() > out_cap) return -1;
  std::memcpy(out, dec.unpacked_data(), dec.unpacked_size());
  return 1;
}

int ref_huff_max_size(int n) { return himg::HuffmanEnc::MaxCompressedSize(n); }

int ref_huff_compress(uint8_t* out, const uint8_t* in, int in_size, int block_size) {
  return himg::HuffmanEnc::Compress(out, in, in_size, block_size);
}

// block_no < 0: whole (unframed) stream; else one framed block.
int ref_huff_uncompress(const uint8_t* in, int in_size, int block_size, int block_no,
                        uint8_t* out, int out_size) {
  himg::HuffmanDec dec(in, in_size, block_size);
  if (!dec.Init()) return 0;
  if (block_no < 0) return dec.Uncompress(out, out_size) ? 1 : 0;
  return dec.UncompressBlock(out, out_size, block_no) ? 1 : 0;
}

// Low-res stage: colour-mapped interleaved pixels -> L image [rows*cols] of one channel and
// that channel's LRES unpacked bytes (predictor bytes ++ deltas).
int ref_lowres_channel(const uint8_t* pixels_chan0, int stride, int w, int h, int quality,
                       uint8_t* L_out, uint8_t* blockdata_out) {
  himg::LowResMapper mapper;
  mapper.InitForQuality(quality);
  himg::Downsampled ds;
  ds.SampleImage(pixels_chan0, stride, w, h);
  const int rows = ds.rows(), cols = ds.columns();
  // L is private; recover it through GetLowresBlock (pixel (0,0) of a block == sample).
  int16_t blk[64];
  for (int v = 0; v < rows; ++v)
    for (int u = 0; u < cols; ++u) {
      ds.GetLowresBlock(blk, u, v);
      L_out[v * cols + u] = static_cast<uint8_t>(blk[0]);
    }
  ds.GetBlockData(blockdata_out, mapper);
  return himg::Downsampled::BlockDataSizePerChannel(rows, cols);
}

// Inverse of the above on the decoder side: LRES unpacked bytes of one channel -> R image.
void ref_lowres_restore(const uint8_t* blockdata, int rows, int cols, const uint8_t* lmap,
                        int lmap_size, uint8_t* R_out) {
  himg::Mapper mapper;
  mapper.SetMappingFunction(lmap, lmap_size);
  himg::Downsampled ds;
  ds.SetBlockData(blockdata, rows, cols, mapper);
  int16_t blk[64];
  for (int v = 0; v < rows; ++v)
    for (int u = 0; u < cols; ++u) {
      ds.GetLowresBlock(blk, u, v);
      R_out[v * cols + u] = static_cast<uint8_t>(blk[0]);
    }
}

void ref_lowres_mapfun(int quality, uint8_t* out128) {
  himg::LowResMapper m;
  m.InitForQuality(quality);
  m.GetMappingFunction(out128);
}

int ref_fullres_mapfun(uint8_t* out, int cap) {
  himg::FullResMapper m;
  m.InitForQuality(0);
  if (m.MappingFunctionSize() > cap) return -1;
  m.GetMappingFunction(out);
  return m.MappingFunctionSize();
}

// which: 0 = low-res mapper at `quality`, 1 = full-res mapper.
int ref_map_to_8bit(int which, int quality, int x) {
  if (which == 0) {
    himg::LowResMapper m;
    m.InitForQuality(quality);
    return m.MapTo8Bit(static_cast<int16_t>(x));
  }
  himg::FullResMapper m;
  m.InitForQuality(quality);
  return m.MapTo8Bit(static_cast<int16_t>(x));
}

int ref_quant_config(int quality, int has_chroma, uint8_t* out64) {
  himg::Quantize q;
  q.InitForQuality(static_cast<uint8_t>(quality), has_chroma != 0);
  q.GetConfiguration(out64);
  return q.ConfigurationSize();
}

void ref_quant_pack(int quality, int chroma, const int16_t* in64, uint8_t* out64) {
  himg::Quantize q;
  q.InitForQuality(static_cast<uint8_t>(quality), true);
  himg::FullResMapper m;
  m.InitForQuality(quality);
  q.Pack(out64, in64, chroma != 0, m);
}

void ref_quant_unpack(int quality, int chroma, const uint8_t* in64, int16_t* out64) {
  himg::Quantize q;
  q.InitForQuality(static_cast<uint8_t>(quality), true);
  himg::FullResMapper m;
  m.InitForQuality(quality);
  q.Unpack(out64, in64, chroma != 0, m);
}

void ref_hadamard_forward(const int16_t* in64, int16_t* out64) {
  himg::Hadamard::Forward(out64, in64);
}

void ref_hadamard_inverse(const int16_t* in64, int16_t* out64) {
  // The reference assumes 16-byte aligned buffers.
  alignas(16) int16_t a[64], b[64];
  std::memcpy(a, in64, sizeof(a));
  himg::Hadamard::Inverse(b, a);
  std::memcpy(out64, b, sizeof(b));
}

void ref_rgb_to_ycbcr(const uint8_t* in, uint8_t* out, int w, int h, int stride, int nch) {
  himg::YCbCr::RGBToYCbCr(out, in, w, h, stride, nch);
}

void ref_ycbcr_to_rgb(uint8_t* buf, int w, int h, int nch) {
  himg::YCbCr::YCbCrToRGB(buf, w, h, nch);
}

}  // extern "C"
