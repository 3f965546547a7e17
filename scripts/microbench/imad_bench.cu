// Integer-pipe microbenchmark for sm_100a: which 32-bit multiply-add forms issue at what rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_bench imad_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 16
#define INNER 32

template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t* out, uint32_t iters, uint32_t b, uint32_t c0) {
  uint64_t acc[CHAINS];
  uint32_t x[CHAINS], lo[CHAINS], hi[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; i++) {
    acc[i] = (uint64_t)(threadIdx.x + 1) * (i + 3);
    x[i] = threadIdx.x * 7 + i + c0;
    lo[i] = x[i] ^ 0x5555; hi[i] = x[i] ^ 0xaaaa;
  }
  for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < INNER; j++) {
      if (MODE == 0) {  // IMAD.WIDE.U32, a = low half of accumulator (dependent)
#pragma unroll
        for (int i = 0; i < CHAINS; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((uint32_t)acc[i]), "r"(b));
      } else if (MODE == 1) {  // IMAD.WIDE.U32, separate a register
#pragma unroll
        for (int i = 0; i < CHAINS; i++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((uint32_t)(acc[(i + 5) % CHAINS] >> 32)), "r"(b));
      } else if (MODE == 2) {  // IMAD (lo)
#pragma unroll
        for (int i = 0; i < CHAINS; i++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(lo[(i + 5) % CHAINS]), "r"(b));
      } else if (MODE == 3) {  // IMAD.HI.U32
#pragma unroll
        for (int i = 0; i < CHAINS; i++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[i]) : "r"(hi[(i + 5) % CHAINS]), "r"(b));
      } else if (MODE == 4) {  // lo + hi pair, separate accumulators (one full product = 2 instr)
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
          uint32_t a = lo[(i + 5) % CHAINS];
          asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(a), "r"(b));
          asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[i]) : "r"(a), "r"(b));
        }
      } else if (MODE == 5) {  // carry chain: mad.lo.cc / madc.hi.cc pairs across the 16 accumulators (-> IMAD.WIDE.U32.X)
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(hi[5]), "r"(b));
#pragma unroll
        for (int i = 1; i < CHAINS; i++)
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(hi[(i + 5) % CHAINS]), "r"(b));
      } else if (MODE == 6) {  // mul.wide.u32 (no accumulate)
#pragma unroll
        for (int i = 0; i < CHAINS; i++) { uint64_t t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"((uint32_t)acc[i]), "r"(b)); acc[i] = t; }
      } else if (MODE == 7) {  // DFMA for reference
#pragma unroll
        for (int i = 0; i < CHAINS; i++) { double d = __longlong_as_double(acc[i]); asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d) : "d"(1.0000001)); acc[i] = __double_as_longlong(d); }
      } else if (MODE == 8) {  // IADD3 (alu pipe)
#pragma unroll
        for (int i = 0; i < CHAINS; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[i]) : "r"(lo[(i + 3) % CHAINS]));
      } else if (MODE == 9) {  // IMAD.WIDE + IADD3 mix 1:1 (do they co-issue?)
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(hi[i]), "+r"(x[i]) : "r"(x[(i + 5) % CHAINS]), "r"(b));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[i]) : "r"(lo[(i + 3) % CHAINS]));
        }
      } else if (MODE == 10) {  // IMAD.WIDE + IMAD.lo mix 1:1
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(hi[i]), "+r"(x[i]) : "r"(x[(i + 5) % CHAINS]), "r"(b));
          asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[i]) : "r"(lo[(i + 3) % CHAINS]), "r"(b));
        }
      }
    }
  }
  uint64_t r = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) r ^= acc[i] ^ lo[i] ^ ((uint64_t)hi[i] << 32) ^ x[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, double ops_per_inner, int blocks_per_sm, int sms) {
  uint64_t* out;
  int blocks = sms * blocks_per_sm;
  cudaMalloc(&out, (size_t)blocks * 256 * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  uint32_t iters = 512;
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters, 0x9e3779b9u, rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  double ops = (double)blocks * 256 * iters * INNER * CHAINS * ops_per_inner;
  double rate = ops / (best * 1e-3);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-44s blocks/SM=%d  %8.3f ms  %.3e instr/s  = %.1f /clk/SM at %d MHz (nominal)\n", name, blocks_per_sm, best, rate,
         rate / sms / (clk * 1e3), clk / 1000);
  cudaFree(out);
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs=%d\n", sms);
  for (int bps : {4, 8}) {
    run<0>("IMAD.WIDE.U32 a=lo(acc)", 1, bps, sms);
    run<1>("IMAD.WIDE.U32 separate a", 1, bps, sms);
    run<2>("IMAD lo", 1, bps, sms);
    run<3>("IMAD.HI.U32", 1, bps, sms);
    run<4>("IMAD lo + IMAD.HI pair (2 instr)", 2, bps, sms);
    run<5>("mad.lo.cc/madc.hi.cc chain (pairs=1 op)", 1, bps, sms);
    run<6>("mul.wide.u32", 1, bps, sms);
    run<7>("DFMA", 1, bps, sms);
    run<8>("IADD (alu)", 1, bps, sms);
    run<9>("IMAD.WIDE + IADD 1:1 (2 instr)", 2, bps, sms);
    run<10>("IMAD.WIDE + IMAD lo 1:1 (2 instr)", 2, bps, sms);
  }
  return 0;
}
