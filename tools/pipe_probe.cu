// Issue-rate probe for the instruction mixes of K-fwd / K-inv on sm_100a: warp instructions per cycle and
// SM sub-partition (SMSP) for single opcodes and for ALU + FMA + LSU mixes.  One CTA of 512 threads per
// SM (4 warps per SMSP), every thread runs ILP independent chains; cycles from clock64().
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe tools/pipe_probe.cu && ./pipe_probe
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <vector>

constexpr int ILP = 8, UNROLL = 16, ITERS = 256;

enum Op { LOP3, SHF, PRMT, IADD3, VIADDMNMX, IMAD, IMADHI, IMADWIDE, IDP4A, LDS16, LDS8, STS16, MIX_ALU_FMA, MIX_ALU_FMA_LDS, MIX_2ALU_1FMA, MIX_IDP_ALU, NOPS };
const char *kNames[] = {"LOP3", "SHF", "PRMT", "IADD3 (3-input)", "VIADDMNMX.S16x2", "IMAD", "IMAD.HI.U32", "IMAD.WIDE.U32", "IDP.4A", "LDS.U16",
                        "LDS.U8", "STS.U16", "LOP3 + IMAD (1:1)", "LOP3 + IMAD + LDS.U16 (2:2:1)", "LOP3 + LOP3 + IMAD (2:1)", "IDP.4A + LOP3 (1:1)"};

template <int OP>
__device__ __forceinline__ void body(uint32_t (&x)[ILP], uint32_t a, uint32_t b, uint32_t sbase, int &count) {
#pragma unroll
  for (int k = 0; k < ILP; ++k) {
    uint32_t &v = x[k];
    if (OP == LOP3) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(a), "r"(b)); count += 1; }
    if (OP == SHF) { asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b)); count += 1; }
    if (OP == PRMT) { asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b)); count += 1; }
    if (OP == IADD3) { asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(v) : "r"(a), "r"(b)); count += 1; }
    if (OP == VIADDMNMX) { v = __viaddmin_s16x2_relu(v, a, b); count += 1; }
    if (OP == IMAD) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b)); count += 1; }
    if (OP == IMADHI) { asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b)); count += 1; }
    if (OP == IMADWIDE) {
      unsigned long long w;
      asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w) : "r"(v), "r"(a), "l"((unsigned long long)b << 20 | v));
      v = (uint32_t)(w >> 7) ^ (uint32_t)(w >> 32);
      count += 1;  // (+ the fold-back ops: read the number with care)
    }
    if (OP == IDP4A) { asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b)); count += 1; }
    if (OP == LDS16) { uint32_t t; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(t) : "r"(sbase + 2 * k * 1024)); v ^= t; count += 1; }
    if (OP == LDS8) { uint32_t t; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(t) : "r"(sbase + 2 * k * 1024)); v ^= t; count += 1; }
    if (OP == STS16) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(sbase + 2 * k * 1024), "r"(v)); count += 1; }
    if (OP == MIX_ALU_FMA) {
      if (k & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(a), "r"(b));
      else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b));
      count += 1;
    }
    if (OP == MIX_2ALU_1FMA) {
      if (k % 3 != 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(a), "r"(b));
      else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b));
      count += 1;
    }
    if (OP == MIX_IDP_ALU) {
      if (k & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(a), "r"(b));
      else asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b));
      count += 1;
    }
    if (OP == MIX_ALU_FMA_LDS) {  // per 5 slots: 2 LOP3, 2 IMAD, 1 LDS.U16
      const int m = k % 5;
      if (m == 0 || m == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(a), "r"(b));
      else if (m == 1 || m == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(a), "r"(b));
      else { uint32_t t; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(t) : "r"(sbase + 2 * k * 1024)); v += t; }
      count += 1;
    }
  }
}

template <int OP>
__global__ void __launch_bounds__(512, 1) probe(uint32_t a, uint32_t b, long long *cycles, uint32_t *sink, int *per_iter) {
  __shared__ uint16_t buf[ILP * 1024 + 1024];
  for (int i = threadIdx.x; i < ILP * 1024 + 1024; i += blockDim.x) buf[i] = (uint16_t)i;
  __syncthreads();
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(buf) + 2 * threadIdx.x;
  uint32_t x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) x[k] = threadIdx.x * 17 + k;
  int count = 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int r = 0; r < UNROLL; ++r) {
      int c = 0;
      body<OP>(x, a, b, sbase, c);
      count = c;
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) acc ^= x[k];
  if (acc == 0x12345678u) sink[0] = acc;
  if (threadIdx.x == 0) {
    cycles[blockIdx.x] = t1 - t0;
    per_iter[0] = count * UNROLL;
  }
}

template <int OP>
void run(long long *d_cycles, uint32_t *d_sink, int *d_per, int sms) {
  probe<OP><<<sms, 512>>>(0x9e3779b9u, 0x7f4a7c15u, d_cycles, d_sink, d_per);
  cudaDeviceSynchronize();
  probe<OP><<<sms, 512>>>(0x9e3779b9u, 0x7f4a7c15u, d_cycles, d_sink, d_per);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> cyc(sms);
  int per = 0;
  cudaMemcpy(cyc.data(), d_cycles, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaMemcpy(&per, d_per, sizeof(int), cudaMemcpyDeviceToHost);
  long long worst = 0;
  for (long long c : cyc) worst = c > worst ? c : worst;
  const double instr = 4.0 * per * ITERS;  // warp instructions per SMSP (4 warps each)
  printf("%-34s %8.3f warp-instr/clk/SMSP   (%lld cycles, %s)\n", kNames[OP], instr / worst, worst, cudaGetErrorString(e));
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs; %d chains per thread, 4 warps per SMSP\n", prop.name, sms, ILP);
  long long *d_cycles;
  uint32_t *d_sink;
  int *d_per;
  cudaMalloc(&d_cycles, sms * sizeof(long long));
  cudaMalloc(&d_sink, 4);
  cudaMalloc(&d_per, 4);
  run<LOP3>(d_cycles, d_sink, d_per, sms);
  run<SHF>(d_cycles, d_sink, d_per, sms);
  run<PRMT>(d_cycles, d_sink, d_per, sms);
  run<IADD3>(d_cycles, d_sink, d_per, sms);
  run<VIADDMNMX>(d_cycles, d_sink, d_per, sms);
  run<IMAD>(d_cycles, d_sink, d_per, sms);
  run<IMADHI>(d_cycles, d_sink, d_per, sms);
  run<IMADWIDE>(d_cycles, d_sink, d_per, sms);
  run<IDP4A>(d_cycles, d_sink, d_per, sms);
  run<LDS16>(d_cycles, d_sink, d_per, sms);
  run<LDS8>(d_cycles, d_sink, d_per, sms);
  run<STS16>(d_cycles, d_sink, d_per, sms);
  run<MIX_ALU_FMA>(d_cycles, d_sink, d_per, sms);
  run<MIX_2ALU_1FMA>(d_cycles, d_sink, d_per, sms);
  run<MIX_IDP_ALU>(d_cycles, d_sink, d_per, sms);
  run<MIX_ALU_FMA_LDS>(d_cycles, d_sink, d_per, sms);
  return 0;
}
