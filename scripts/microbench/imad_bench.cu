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
  double ad[CHAINS];
  const double bb = __longlong_as_double(0x4330000000000000ull | (uint64_t)b) - 4503599627370496.0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++)
    ad[i] = __longlong_as_double(0x4330000000000000ull | ((uint64_t)x[i] << 20 | lo[i])) - 4503599627370496.0;  // < 2^52
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
      } else if (MODE == 6) {  // mul.wide.u32 (no accumulate), BOTH halves of the product consumed by the next round
                               // (round 1 fed only the low half back: ptxas dropped the dead high half and emitted plain
                               // IMAD -- the "63 /clk/SM" of r01_imad_bench.log was the IMAD rate, not a wide multiply)
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
          uint64_t t;
          const uint64_t src = acc[(i + 5) % CHAINS];
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"((uint32_t)src ^ (uint32_t)(src >> 32)), "r"(b));
          acc[i] = t;
        }
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
      } else if (MODE == 11 || MODE == 12) {  // DFMA co-issued with the IMAD.WIDE carry chain, 1:1 and 2:1 per thread
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(hi[i]), "+r"(x[i]) : "r"(x[(i + 5) % CHAINS]), "r"(b));
          double d = __longlong_as_double(acc[i]);
          asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d) : "d"(1.0000001));
          if (MODE == 12) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d) : "d"(0.9999999));
          acc[i] = __double_as_longlong(d);
        }
      } else if (MODE == 13 || MODE == 14) {
        // One 52 x 52 -> 104-bit limb product the double-precision way (Emmart & al., ARITH 2018): hi = fma_rz(a, b, 2^104),
        // lo = fma_rz(a, b, (2^104 + 2^52) - hi) [one DADD], both accumulated as 64-bit INTEGERS (the exponent bits are a known
        // constant per term).  3 FP64-pipe + 2 x (IADD3 + IADD3.X) per limb product.  MODE 14 adds two IMAD.WIDE.X of an
        // independent 32-bit-limb chain per limb product (the hybrid: both multiplier pipes busy in one thread).
        const double c1 = 20282409603651670423947251286016.0;               // 2^104
        const double c2 = 20282409603651670423947251286016.0 + 4503599627370496.0;  // 2^104 + 2^52
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
          const double a = ad[(i + 5) % CHAINS];  // operand limbs are converted once per field element, not per product
          double h, l, sub;
          asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(h) : "d"(a), "d"(bb), "d"(c1));
          asm volatile("sub.rn.f64 %0, %1, %2;" : "=d"(sub) : "d"(c2), "d"(h));
          asm volatile("fma.rz.f64 %0, %1, %2, %3;" : "=d"(l) : "d"(a), "d"(bb), "d"(sub));
          acc[i] += (uint64_t)__double_as_longlong(h);
          acc[(i + 1) % CHAINS] += (uint64_t)__double_as_longlong(l);
          ad[i] = l;  // keeps the products loop-variant (timing only: l is in [2^52, 2^53))
          if (MODE == 14) {
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(hi[i]), "+r"(x[i]) : "r"(x[(i + 5) % CHAINS]), "r"(b));
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(x[(i + 3) % CHAINS]) : "r"(hi[(i + 7) % CHAINS]), "r"(b));
          }
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
    run<6>("mul.wide.u32 (both halves live)", 1, bps, sms);
    run<7>("DFMA", 1, bps, sms);
    run<8>("IADD (alu)", 1, bps, sms);
    run<9>("IMAD.WIDE + IADD 1:1 (2 instr)", 2, bps, sms);
    run<10>("IMAD.WIDE + IMAD lo 1:1 (2 instr)", 2, bps, sms);
    run<11>("IMAD.WIDE + DFMA 1:1 (counted: 2 instr)", 2, bps, sms);
    run<12>("IMAD.WIDE + 2 DFMA (counted: 3 instr)", 3, bps, sms);
    run<13>("52-bit limb product via 2 DFMA + DADD + 2 IADD64 (counted: 1)", 1, bps, sms);
    run<14>("52-bit limb product + 2 IMAD.WIDE hybrid (counted: 1)", 1, bps, sms);
  }
  return 0;
}
