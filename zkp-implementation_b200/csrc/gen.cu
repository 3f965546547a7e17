// Setup-time generators and the integer-pipe microbenchmark.
//
//  * gen_srs_dev      `Srs::new_from_secret` (kzg/src/srs.rs:48-69): [secret^i * G]_{i<n}, affine.
//                     The reference runs n serial scalar multiplications (255 doublings + ~128
//                     additions and one Fq inversion each).  Here the generator is a FIXED base: a
//                     32 x 256 table of d * 2^(8w) * G turns every scalar multiplication into 32 mixed
//                     additions, every thread owns a short run of consecutive powers, and the
//                     normalisation uses Montgomery's batch-inversion trick.
//  * gen_bases_dev    n distinct pseudo-random points (a0 + i*delta)*G for benchmark configs 2/5
//                     ("random points, not an SRS with known structure" -- SURVEY.md 8d).
//  * bench_imad       independent 32x32->64 multiply-add chains on every SM: the measured
//                     integer-pipe peak that the MSM roofline is quoted against.
#include <string.h>

#include "engine.h"
#include "memops.cuh"

namespace zkp {

static constexpr uint32_t GEN_THREADS = 128;
static constexpr uint32_t CHAIN_LEN = 64;   // points per thread in gen_bases / batch normalise
static constexpr uint32_t SRS_RUN = 8;      // consecutive powers per thread in gen_srs

// tmp[i] = start + i * step  (XYZZ), thread t owns [t*CHAIN_LEN, (t+1)*CHAIN_LEN)
__global__ void __launch_bounds__(GEN_THREADS) gen_chain_kernel(const G1Xyzz* start_p, const G1Affine* step_p, size_t n, G1Xyzz* tmp) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t first = t * CHAIN_LEN;
  if (first >= n) return;
  const G1Affine step = ld_affine(step_p);
  const G1Xyzz start = ld_xyzz(start_p);
  G1Xyzz acc = xyzz_mul_u32(G1Xyzz::from_affine(step), (uint32_t)first);
  xyzz_add(acc, start);
  for (uint32_t i = 0; i < CHAIN_LEN && first + i < n; i++) {
    st_xyzz(tmp + first + i, acc);
    xyzz_madd(acc, step);
  }
}

// Fixed-base table of the generator: gtab[w * 256 + d] = d * 2^(8w) * G (d = 0 is the point at infinity).
// Row w is built by one thread walking d = 1 .. 255 with mixed additions of 2^(8w) G.
static constexpr uint32_t GTAB_WINDOWS = 32, GTAB_DIGITS = 256;

__global__ void gen_gtab_kernel(G1Xyzz* tab) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= GTAB_WINDOWS) return;
  G1Xyzz base = G1Xyzz::from_affine(g1_generator());
  for (uint32_t k = 0; k < 8 * w; k++) base = xyzz_dbl(base);
  G1Xyzz acc = G1Xyzz::infinity();
  for (uint32_t d = 0; d < GTAB_DIGITS; d++) {
    st_xyzz(tab + (size_t)w * GTAB_DIGITS + d, acc);
    xyzz_add(acc, base);
  }
}

// tmp[i] = secret^(start + i) * G (XYZZ) = sum over the 32 bytes of the canonical power of a table entry
// (`start`: index of the first power, for a rank that owns a point range of a sharded SRS)
__global__ void __launch_bounds__(GEN_THREADS) gen_srs_kernel(Fr secret, size_t start, size_t n,
                                                              const G1Affine* __restrict__ gtab, G1Xyzz* tmp) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t first = t * SRS_RUN;
  if (first >= n) return;
  Fr cur = fp_pow_u64(secret, (uint64_t)(start + first));
  for (uint32_t i = 0; i < SRS_RUN && first + i < n; i++) {
    const Fr k = fp_from_mont(cur);
    G1Xyzz acc = G1Xyzz::infinity();
    for (uint32_t w = 0; w < GTAB_WINDOWS; w++) {
      const uint32_t d = (k.v[w >> 2] >> ((w & 3) * 8)) & 0xffu;
      if (d) xyzz_madd(acc, ld_affine(gtab + (size_t)w * GTAB_DIGITS + d));
    }
    st_xyzz(tmp + first + i, acc);
    cur = fp_mul(cur, secret);
  }
}

// XYZZ -> affine for runs of CHAIN_LEN points with one inversion per run (Montgomery's trick).
// out[i].x is used as scratch for the running prefix products before it receives the result.
__global__ void __launch_bounds__(GEN_THREADS) batch_normalise_kernel(const G1Xyzz* tmp, size_t n, G1Affine* out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t first = t * CHAIN_LEN;
  if (first >= n) return;
  const uint32_t cnt = (uint32_t)((n - first < CHAIN_LEN) ? (n - first) : CHAIN_LEN);
  Fq prod = Fq::one();
  for (uint32_t i = 0; i < cnt; i++) {
    const G1Xyzz p = ld_xyzz(tmp + first + i);
    st_fq(&out[first + i].x, prod);
    if (!p.is_inf()) prod = prod * (p.zz * p.zzz);
  }
  Fq inv = fq_inv_gcd(prod);
  for (int i = (int)cnt - 1; i >= 0; i--) {
    const G1Xyzz p = ld_xyzz(tmp + first + i);
    G1Affine a = G1Affine::infinity();
    if (!p.is_inf()) {
      const Fq pre = ld_fq(&out[first + i].x);
      const Fq ti = inv * pre;          // (zz * zzz)^-1
      inv = inv * (p.zz * p.zzz);
      a.x = p.x * (ti * p.zzz);         // X / ZZ
      a.y = p.y * (ti * p.zz);          // Y / ZZZ
    }
    st_fq(&out[first + i].x, a.x);
    st_fq(&out[first + i].y, a.y);
  }
}

static uint64_t splitmix64(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

int normalise_dev(Ctx* ctx, const G1Xyzz* tmp, size_t n, G1Affine* out);
static int normalise(Ctx* ctx, const G1Xyzz* tmp, size_t n, G1Affine* out) { return normalise_dev(ctx, tmp, n, out); }
int normalise_dev(Ctx* ctx, const G1Xyzz* tmp, size_t n, G1Affine* out) {
  const size_t threads = (n + CHAIN_LEN - 1) / CHAIN_LEN;
  ZKP_LAUNCH_NOSYNC(batch_normalise_kernel, dim3((unsigned)((threads + GEN_THREADS - 1) / GEN_THREADS)), dim3(GEN_THREADS), 0,
             ctx->stream, tmp, n, out);
  return rt::check_last();
}

int gen_bases_range_dev(Ctx* ctx, uint64_t seed, size_t first, size_t n, G1Affine* out);
int gen_bases_dev(Ctx* ctx, uint64_t seed, size_t n, G1Affine* out) { return gen_bases_range_dev(ctx, seed, 0, n, out); }

// points [first, first + n) of the progression: a shard of a larger synthetic point set (multi-GPU runs generate
// the SAME global set whatever the number of ranks)
int gen_bases_range_dev(Ctx* ctx, uint64_t seed, size_t first, size_t n, G1Affine* out) {
  if (n == 0) return ZKP_OK;
  if (n >= ((size_t)1 << 32) || first >= ((size_t)1 << 32)) return ZKP_ERR_INVALID_ARG;
  uint64_t s = seed;
  uint32_t a0[8], dl[8];
  for (int i = 0; i < 4; i++) {
    uint64_t x = splitmix64(s), y = splitmix64(s);
    a0[2 * i] = (uint32_t)x; a0[2 * i + 1] = (uint32_t)(x >> 32);
    dl[2 * i] = (uint32_t)y; dl[2 * i + 1] = (uint32_t)(y >> 32);
  }
  a0[7] &= 0x3fffffffu;  // < 2^254 < r
  dl[7] &= 0x3fffffffu;
  dl[0] |= 1;
  const G1Xyzz g = G1Xyzz::from_affine(g1_generator());
  G1Xyzz start = xyzz_mul_limbs(g, a0, 8);
  const G1Affine step = xyzz_to_affine(xyzz_mul_limbs(g, dl, 8));
  if (first) {
    const G1Xyzz skip = xyzz_mul_u32(G1Xyzz::from_affine(step), (uint32_t)first);
    xyzz_add(start, skip);
  }
  DevBuf tmp;
  ZKP_TRY(tmp.reserve(n * sizeof(G1Xyzz) + sizeof(G1Xyzz) + sizeof(G1Affine)));
  // chain parameters live in device memory (large by-value kernel parameters are avoided)
  G1Xyzz* start_d = tmp.as<G1Xyzz>() + n;
  G1Affine* step_d = reinterpret_cast<G1Affine*>(start_d + 1);
  ZKP_TRY(rt::h2d(start_d, &start, sizeof(start), ctx->stream));
  ZKP_TRY(rt::h2d(step_d, &step, sizeof(step), ctx->stream));
  ZKP_TRY(rt::sync(ctx->stream));
  const size_t threads = (n + CHAIN_LEN - 1) / CHAIN_LEN;
  ZKP_LAUNCH_NOSYNC(gen_chain_kernel, dim3((unsigned)((threads + GEN_THREADS - 1) / GEN_THREADS)), dim3(GEN_THREADS), 0,
             ctx->stream, start_d, step_d, n, tmp.as<G1Xyzz>());
  int st = normalise(ctx, tmp.as<G1Xyzz>(), n, out);
  if (st == ZKP_OK) st = rt::sync(ctx->stream);
  tmp.release();
  return st;
}

int gen_srs_dev(Ctx* ctx, const Fr& secret, size_t start, size_t n, G1Affine* out) {
  if (n == 0) return ZKP_OK;
  const size_t tab_n = (size_t)GTAB_WINDOWS * GTAB_DIGITS;
  DevBuf tmp, tab;
  const size_t tmp_n = n > tab_n ? n : tab_n;
  ZKP_TRY(tmp.reserve(tmp_n * sizeof(G1Xyzz)));
  int st0 = tab.reserve(tab_n * sizeof(G1Affine));
  if (st0 != ZKP_OK) { tmp.release(); return st0; }
  // the generator's window table (XYZZ in tmp, normalised into tab)
  ZKP_LAUNCH_NOSYNC(gen_gtab_kernel, dim3(1), dim3(GTAB_WINDOWS), 0, ctx->stream, tmp.as<G1Xyzz>());
  st0 = normalise(ctx, tmp.as<G1Xyzz>(), tab_n, tab.as<G1Affine>());
  if (st0 != ZKP_OK) { tmp.release(); tab.release(); return st0; }
  const size_t threads = (n + SRS_RUN - 1) / SRS_RUN;
  ZKP_LAUNCH_NOSYNC(gen_srs_kernel, dim3((unsigned)((threads + GEN_THREADS - 1) / GEN_THREADS)), dim3(GEN_THREADS), 0,
             ctx->stream, secret, start, n, (const G1Affine*)tab.as<G1Affine>(), tmp.as<G1Xyzz>());
  int st = normalise(ctx, tmp.as<G1Xyzz>(), n, out);
  if (st == ZKP_OK) st = rt::sync(ctx->stream);
  tmp.release();
  tab.release();
  return st;
}

// ------------------------------------------------------------------------------------------------
// Integer-pipe peak: 16 independent multiply-add chains per thread, all SMs busy.
// ------------------------------------------------------------------------------------------------
#ifndef ZKP_EMU
static constexpr int IMAD_CHAINS = 16;
static constexpr int IMAD_INNER = 64;

// The multiplier the field code actually issues: mad.lo.cc / madc.hi.cc pairs on one carry chain,
// which ptxas fuses into IMAD.WIDE.U32.X (32x32+64 with carry in/out).  The multiplicand of every
// link comes from another chain's previous result so nothing is loop-invariant.
__global__ void __launch_bounds__(256) imad_wide_kernel(uint64_t* out, uint32_t iters, uint32_t b) {
  uint32_t lo[IMAD_CHAINS], hi[IMAD_CHAINS];
#pragma unroll
  for (int k = 0; k < IMAD_CHAINS; k++) { lo[k] = (threadIdx.x + 1) * (k + 3); hi[k] = lo[k] ^ 0x55555555u; }
  for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < IMAD_INNER; j++) {
      asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                   : "+r"(lo[0]), "+r"(hi[0]) : "r"(hi[5]), "r"(b));
#pragma unroll
      for (int k = 1; k < IMAD_CHAINS; k++)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;"
                     : "+r"(lo[k]), "+r"(hi[k]) : "r"(hi[(k + 5) % IMAD_CHAINS]), "r"(b));
    }
  }
  uint64_t x = 0;
#pragma unroll
  for (int k = 0; k < IMAD_CHAINS; k++) x ^= lo[k] ^ ((uint64_t)hi[k] << 32);
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x;
}

__global__ void __launch_bounds__(256) imad_lo_kernel(uint64_t* out, uint32_t iters, uint32_t b) {
  uint32_t acc[IMAD_CHAINS];
#pragma unroll
  for (int k = 0; k < IMAD_CHAINS; k++) acc[k] = (threadIdx.x + 1) * (k + 3);
  for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
    for (int j = 0; j < IMAD_INNER; j++) {
#pragma unroll
      for (int k = 0; k < IMAD_CHAINS; k++)
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(acc[(k + 5) % IMAD_CHAINS]), "r"(b));
    }
  }
  uint32_t x = 0;
#pragma unroll
  for (int k = 0; k < IMAD_CHAINS; k++) x ^= acc[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x;
}

int bench_imad(Ctx* ctx, double* wide, double* lo) {
  const unsigned blocks = (unsigned)ctx->sm_count * 8, threads = 256;
  const uint32_t iters = 256;
  DevBuf out;
  ZKP_TRY(out.reserve((size_t)blocks * threads * 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best[2] = {0, 0};
  for (int which = 0; which < 2; which++) {
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0, ctx->stream);
      if (which == 0) imad_wide_kernel<<<blocks, threads, 0, ctx->stream>>>(out.as<uint64_t>(), iters, 0x9e3779b9u);
      else imad_lo_kernel<<<blocks, threads, 0, ctx->stream>>>(out.as<uint64_t>(), iters, 0x9e3779b9u);
      cudaEventRecord(e1, ctx->stream);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double ops = (double)blocks * threads * iters * IMAD_INNER * IMAD_CHAINS;
      const double rate = ops / (ms * 1e-3);
      if (rep > 0 && rate > best[which]) best[which] = rate;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  out.release();
  *wide = best[0];
  *lo = best[1];
  return rt::check_last();
}
#else
int bench_imad(Ctx*, double* wide, double* lo) {
  *wide = 0;
  *lo = 0;
  return ZKP_OK;
}
#endif

}  // namespace zkp
