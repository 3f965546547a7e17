// Batched-affine bucket accumulation for the Pippenger MSM (msm.cu): the first R "tree rounds" of the bucket sums.
//
// Replaces, for the bulk of the additions, the XYZZ mixed-add loop of `msm_accumulate_kernel` -- same sum
// (kzg/src/scheme.rs:84-96: sum_i c_i * P_i), fewer field products per addition:
//   XYZZ mixed add           8 M + 2 S                       = 10 products
//   affine add, shared inv   lambda, lambda^2, lambda * dx   =  3 products  + 3 for Montgomery's trick
// An affine addition needs 1 / (x2 - x1).  Inside one bucket the additions are serial, but the additions of one TREE
// LEVEL are all independent: round r replaces every bucket's run of L points by ceil(L / 2) points (neighbours added
// in pairs, an odd last point carried over), so the whole round is ONE batch of M_out independent pair-additions whose
// denominators are inverted together:
//   A. denominators  thread t takes K consecutive outputs, multiplies their denominators into a running product,
//                    stores the exclusive prefixes (48 B each) and its total;
//   B. inversion     the per-thread totals are inverted by the same up-sweep / down-sweep, 16-fold per level, until
//                    <= 4096 values are left for one inversion each (safegcd division steps, inv_gcd.cuh: the only
//                    inversions of the round);
//   C. additions     thread t walks its K outputs backwards: 1/den_k = inv * prefix_k, inv *= den_k, then
//                    lambda = num / den, x3 = lambda^2 - x1 - x2, y3 = lambda (x1 - x3) - y1.
// Prefixes and totals travel through HBM (about 0.5 KB per addition over a round, against ~1750 multiply-adds on
// the integer pipe that bounds the kernel): products per addition = 6 - 2/K (element 0 of a thread: prefix 1, no running-inverse
// update; -1/K more in A) + 3/K (level B), K = 16.
// After R rounds every run is ~2^R times shorter; what is left goes through the XYZZ task kernel (msm.cu), which
// also owns the splitting of heavily loaded buckets.
//
// All special cases are decided from the operands (not from the schedule), so they are identical in A and C:
// P + O, O + P, P + P (tangent slope 3x^2 / 2y), P + (-P) = O.  Field arithmetic is exact: the round's output points
// are the unique affine representatives, whatever K or R.
#include "engine.h"
#include "inv_gcd.cuh"
#include "memops.cuh"

namespace zkp {

#ifndef ZKP_AFF_K
#define ZKP_AFF_K 16
#endif
static constexpr uint32_t AFF_K = ZKP_AFF_K;  // outputs per thread in A / C
static constexpr uint32_t AFF_INV_K = 16;     // fan-in of the inversion tree (level B)
static constexpr uint32_t AFF_INV_DIRECT = 4096;  // at most this many values go to the Fermat kernel
static constexpr uint32_t AFF_THREADS = 128;
static constexpr uint32_t AFF_SIGN = 0x80000000u;
static constexpr uint32_t AFF_NONE = 0xffffffffu;

struct AffRound {
  const uint32_t* off_in;   // [nb + 1] run starts of the round's input (exclusive scan of the run lengths)
  const uint32_t* off_out;  // [nb + 1] run starts of the output; off_out[nb] = M_out
  uint32_t nb;
  const uint32_t* svals;    // first round: sorted (point index | sign << 31); later rounds: null
  const G1Affine* in;       // first round: the bases / fixed-base table; later rounds: the previous round's output
  uint32_t in_stride16;     // first round: record stride of `in` in 16-byte words (6 packed, 8 padded table)
  G1Affine* out;
  Fq* pre;                  // [M_out] exclusive prefix products of the denominators (per thread)
  Fq* tot;                  // [ceil(M_out / K)] per-thread products
  uint32_t* stats;          // profiling: stats[0] = M_in, stats[1] = M_out (null when profiling is off)
  const uint32_t* tb;       // [ceil(M_out / K)] bucket holding output t * K (aff_start_bucket_kernel): no search per thread
};

// len_out[b] = ceil(len_in[b] / 2); entry nb = 0 so that the exclusive scan leaves M_out there
__global__ void __launch_bounds__(256) aff_halve_kernel(const uint32_t* __restrict__ off_in, uint32_t nb,
                                                        uint32_t* __restrict__ len_out) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nb) return;
  len_out[b] = (b < nb) ? ((off_in[b + 1] - off_in[b] + 1) >> 1) : 0u;
}

// len0[b] = bend[b] - bstart[b] (msm_bounds_kernel leaves both 0 for an empty bucket)
__global__ void __launch_bounds__(256) aff_len0_kernel(const uint32_t* __restrict__ bstart, const uint32_t* __restrict__ bend,
                                                       uint32_t nb, uint32_t* __restrict__ len) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nb) return;
  len[b] = (b < nb) ? (bend[b] - bstart[b]) : 0u;
}

// tb[t] = the bucket holding output t * K, for every thread t of A / C.  A binary search per thread costs ~21 dependent
// L2 reads before the first useful load (a third of kernel A's per-thread latency); the inverse map is a scatter: bucket
// b owns the threads ceil(off_out[b] / K) .. ceil(off_out[b + 1] / K) - 1 (3 on average).  A bucket owning more than
// AFF_TB_SERIAL threads (skewed scalars: few, heavily loaded buckets) goes to a list filled by whole blocks.
static constexpr uint32_t AFF_TB_SERIAL = 64;
__global__ void __launch_bounds__(256) aff_start_bucket_kernel(const uint32_t* __restrict__ off_out, uint32_t nb,
                                                               uint32_t* __restrict__ tb, uint32_t* __restrict__ heavy,
                                                               uint32_t heavy_cap) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  const uint32_t ob = off_out[b], oe = off_out[b + 1];
  if (oe == ob) return;
  const uint32_t t0 = (ob + AFF_K - 1) / AFF_K, t1 = (oe + AFF_K - 1) / AFF_K;
  if (t1 - t0 > AFF_TB_SERIAL) {
    const uint32_t slot = atomicAdd(heavy, 1u);
    if (slot < heavy_cap) { heavy[1 + slot] = b; return; }
  }
  for (uint32_t t = t0; t < t1; t++) tb[t] = b;
}

__global__ void __launch_bounds__(256) aff_start_bucket_heavy_kernel(const uint32_t* __restrict__ off_out,
                                                                     uint32_t* __restrict__ tb,
                                                                     const uint32_t* __restrict__ heavy, uint32_t heavy_cap) {
  uint32_t nh = heavy[0];
  if (nh > heavy_cap) nh = heavy_cap;
  for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
    const uint32_t b = heavy[1 + h];
    const uint32_t t0 = (off_out[b] + AFF_K - 1) / AFF_K, t1 = (off_out[b + 1] + AFF_K - 1) / AFF_K;
    for (uint32_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) tb[t] = b;
  }
}

// idx[k] = position of the first operand of output o0 + k in the input order, bit 31 set when the output is a lone
// carried-over point; AFF_NONE past the end
__device__ __forceinline__ void aff_walk(const AffRound& a, uint32_t t, uint32_t o0, uint32_t mout, uint32_t (&idx)[AFF_K]) {
  uint32_t b = a.tb[t];
  uint32_t ob = a.off_out[b], oe = a.off_out[b + 1], ib = a.off_in[b], ie = a.off_in[b + 1];
#pragma unroll
  for (uint32_t k = 0; k < AFF_K; k++) {
    const uint32_t o = o0 + k;
    if (o >= mout) { idx[k] = AFF_NONE; continue; }
    while (o >= oe) {
      b++;
      ob = oe; oe = a.off_out[b + 1];
      ib = ie; ie = a.off_in[b + 1];
    }
    const uint32_t in0 = ib + 2 * (o - ob);
    idx[k] = in0 | ((in0 + 1 >= ie) ? AFF_SIGN : 0u);
  }
}

enum { AFF_GENERAL = 0, AFF_DOUBLE = 1, AFF_IS_P1 = 2, AFF_IS_P2 = 3, AFF_IS_INF = 4 };

__device__ __forceinline__ int aff_classify(const G1Affine& p1, const G1Affine& p2) {
  if (p1.is_inf()) return AFF_IS_P2;
  if (p2.is_inf()) return AFF_IS_P1;
  if (p1.x == p2.x) return (p1.y == p2.y && !p1.y.is_zero()) ? AFF_DOUBLE : AFF_IS_INF;
  return AFF_GENERAL;
}

// First round: resolve the (point index | sign) indirection once, so that each element costs one long-latency hop
// (the gather) instead of two.  ref[k] = table index | sign of the first operand, ref2[k] of the second.
template <bool FIRST>
__device__ __forceinline__ void aff_refs(const AffRound& a, const uint32_t (&idx)[AFF_K], uint32_t (&r1)[AFF_K],
                                         uint32_t (&r2)[AFF_K]) {
#pragma unroll
  for (uint32_t k = 0; k < AFF_K; k++) {
    const uint32_t id = idx[k];
    if (id == AFF_NONE) { r1[k] = 0; r2[k] = 0; continue; }
    const uint32_t p = id & ~AFF_SIGN;
    if (FIRST) {
      r1[k] = a.svals[p];
      r2[k] = (id & AFF_SIGN) ? 0u : a.svals[p + 1];
    } else {
      r1[k] = p;
      r2[k] = (id & AFF_SIGN) ? p : p + 1;  // a carried-over point has no partner: never read past the run
    }
  }
}

template <bool FIRST>
__device__ __forceinline__ G1Affine aff_ld_ref(const AffRound& a, uint32_t ref) {
  if (!FIRST) return ld_affine(a.in + ref);
  G1Affine p = ld_affine_s(a.in, ref & ~AFF_SIGN, a.in_stride16);
  if (ref & AFF_SIGN) p = g1_neg(p);
  return p;
}

template <bool FIRST>
__device__ __forceinline__ Fq aff_ld_ref_x(const AffRound& a, uint32_t ref) {
  if (!FIRST) return ld_fq(&a.in[ref].x);
  return ld_affine_x_s(a.in, ref & ~AFF_SIGN, a.in_stride16);
}

// ---- A: denominators ------------------------------------------------------------------------------
#ifndef ZKP_AFF_A_BLOCKS
#define ZKP_AFF_A_BLOCKS 4
#endif
#ifndef ZKP_AFF_A_PREFETCH
#define ZKP_AFF_A_PREFETCH 1
#endif
template <bool FIRST>
__global__ void __launch_bounds__(AFF_THREADS, ZKP_AFF_A_BLOCKS) aff_denominators_kernel(AffRound a) {
  const uint32_t mout = a.off_out[a.nb];
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t o0_64 = (uint64_t)t * AFF_K;
  if (t == 0 && a.stats) { a.stats[0] = a.off_in[a.nb]; a.stats[1] = mout; }
  if (o0_64 >= mout) return;
  const uint32_t o0 = (uint32_t)o0_64;
  uint32_t idx[AFF_K], r1[AFF_K], r2[AFF_K];
  aff_walk(a, t, o0, mout, idx);
  aff_refs<FIRST>(a, idx, r1, r2);
  Fq run = Fq::one();
  // the x coordinates of element k + 1 are requested before element k is multiplied in
#if ZKP_AFF_A_PREFETCH
  Fq nx1 = aff_ld_ref_x<FIRST>(a, r1[0]), nx2 = aff_ld_ref_x<FIRST>(a, r2[0]);
#endif
#pragma unroll 1
  for (uint32_t k = 0; k < AFF_K; k++) {
#if ZKP_AFF_A_PREFETCH
    const Fq x1 = nx1, x2 = nx2;
    if (k + 1 < AFF_K) {
      nx1 = aff_ld_ref_x<FIRST>(a, r1[k + 1]);
      nx2 = aff_ld_ref_x<FIRST>(a, r2[k + 1]);
    }
#else
    const Fq x1 = aff_ld_ref_x<FIRST>(a, r1[k]), x2 = aff_ld_ref_x<FIRST>(a, r2[k]);
#endif
    const uint32_t id = idx[k];
    if (id == AFF_NONE || (id & AFF_SIGN)) continue;  // nothing / a carried-over point: no denominator
    Fq den = fp_sub(x2, x1);
    if (x1.is_zero() || x2.is_zero() || den.is_zero()) {  // rare: infinity operand, doubling or cancellation
      const G1Affine p1 = aff_ld_ref<FIRST>(a, r1[k]), p2 = aff_ld_ref<FIRST>(a, r2[k]);
      const int kind = aff_classify(p1, p2);
      if (kind == AFF_DOUBLE) den = fp_dbl(p1.y);
      else if (kind != AFF_GENERAL) continue;
    }
    // element 0 always sees run = 1: its prefix is never read (C uses inv itself) and the product by one is skipped
    if (k == 0) {
      run = den;
    } else {
      st_fq(a.pre + o0 + k, run);
      run = run * den;
    }
  }
  st_fq(a.tot + t, run);
}

// ---- C: additions ----------------------------------------------------------------------------------
#ifndef ZKP_AFF_C_BLOCKS
#define ZKP_AFF_C_BLOCKS 4
#endif
#ifndef ZKP_AFF_C_PREFETCH
#define ZKP_AFF_C_PREFETCH 0
#endif
template <bool FIRST>
__global__ void __launch_bounds__(AFF_THREADS, ZKP_AFF_C_BLOCKS) aff_add_kernel(AffRound a) {
  const uint32_t mout = a.off_out[a.nb];
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t o0_64 = (uint64_t)t * AFF_K;
  if (o0_64 >= mout) return;
  const uint32_t o0 = (uint32_t)o0_64;
  uint32_t idx[AFF_K], r1[AFF_K], r2[AFF_K];
  aff_walk(a, t, o0, mout, idx);
  aff_refs<FIRST>(a, idx, r1, r2);
  Fq inv = ld_fq(a.tot + t);  // 1 / (product of this thread's denominators)
  // the operands of element k - 1 are requested before element k is computed (same pattern as msm_accumulate_kernel)
#if ZKP_AFF_C_PREFETCH
  G1Affine n1 = aff_ld_ref<FIRST>(a, r1[AFF_K - 1]), n2 = aff_ld_ref<FIRST>(a, r2[AFF_K - 1]);
#endif
#pragma unroll 1
  for (int k = AFF_K - 1; k >= 0; k--) {
#if ZKP_AFF_C_PREFETCH
    const G1Affine p1 = n1, p2 = n2;
    if (k > 0) {
      n1 = aff_ld_ref<FIRST>(a, r1[k - 1]);
      n2 = aff_ld_ref<FIRST>(a, r2[k - 1]);
    }
#else
    const G1Affine p1 = aff_ld_ref<FIRST>(a, r1[k]), p2 = aff_ld_ref<FIRST>(a, r2[k]);
#endif
    const uint32_t id = idx[k];
    if (id == AFF_NONE) continue;
    G1Affine* dst = a.out + o0 + k;
    if (id & AFF_SIGN) { st_affine(dst, p1); continue; }
    Fq num = fp_sub(p2.y, p1.y), den = fp_sub(p2.x, p1.x);
    if (p1.x.is_zero() || p2.x.is_zero() || den.is_zero()) {
      const int kind = aff_classify(p1, p2);
      if (kind == AFF_DOUBLE) {
        const Fq xx = fp_sqr(p1.x);
        num = fp_add(fp_dbl(xx), xx);
        den = fp_dbl(p1.y);
      } else if (kind != AFF_GENERAL) {
        st_affine(dst, kind == AFF_IS_P1 ? p1 : (kind == AFF_IS_P2 ? p2 : G1Affine::infinity()));
        continue;
      }
    }
    // the last element of the backward walk: its prefix is 1 and nobody needs the updated running inverse
    Fq dinv = inv;
    if (k > 0) {
      dinv = inv * ld_fq(a.pre + o0 + k);
      inv = inv * den;
    }
    const Fq lam = num * dinv;
    G1Affine r;
    r.x = fp_sub(fp_sub(fp_sqr(lam), p1.x), p2.x);
    r.y = fp_sub(lam * fp_sub(p1.x, r.x), p1.y);
    st_affine(dst, r);
  }
}

// ---- B: inversion tree over the per-thread totals ---------------------------------------------------
// level sizes follow from M_out on the device: n(div) = ceil(M_out / div)
__device__ __forceinline__ uint32_t aff_level_count(const uint32_t* __restrict__ mout_ptr, uint32_t div) {
  const uint32_t m = *mout_ptr;
  return (uint32_t)(((uint64_t)m + div - 1) / div);
}

__global__ void __launch_bounds__(AFF_THREADS) aff_inv_up_kernel(const Fq* __restrict__ v, Fq* __restrict__ pre,
                                                                 Fq* __restrict__ tot, const uint32_t* __restrict__ mout_ptr,
                                                                 uint32_t div) {
  const uint32_t n = aff_level_count(mout_ptr, div);
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t i0 = (uint64_t)u * AFF_INV_K;
  if (i0 >= n) return;
  const uint32_t cnt = (n - i0 < AFF_INV_K) ? (uint32_t)(n - i0) : AFF_INV_K;
  Fq run = ld_fq(v + i0);
  for (uint32_t i = 1; i < cnt; i++) {
    st_fq(pre + i0 + i, run);
    run = run * ld_fq(v + i0 + i);
  }
  st_fq(tot + u, run);
}

__global__ void __launch_bounds__(AFF_THREADS) aff_inv_down_kernel(Fq* __restrict__ v, const Fq* __restrict__ pre,
                                                                   const Fq* __restrict__ tot_inv,
                                                                   const uint32_t* __restrict__ mout_ptr, uint32_t div) {
  const uint32_t n = aff_level_count(mout_ptr, div);
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t i0 = (uint64_t)u * AFF_INV_K;
  if (i0 >= n) return;
  const uint32_t cnt = (n - i0 < AFF_INV_K) ? (uint32_t)(n - i0) : AFF_INV_K;
  Fq inv = ld_fq(tot_inv + u);
  for (uint32_t i = cnt; i-- > 1;) {
    const Fq x = ld_fq(v + i0 + i);
    st_fq(v + i0 + i, inv * ld_fq(pre + i0 + i));
    inv = inv * x;
  }
  st_fq(v + i0, inv);
}

__global__ void __launch_bounds__(AFF_THREADS) aff_inv_direct_kernel(Fq* __restrict__ v, const uint32_t* __restrict__ mout_ptr,
                                                                     uint32_t div) {
  const uint32_t n = aff_level_count(mout_ptr, div);
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  st_fq(v + u, fq_inv_gcd(ld_fq(v + u)));  // products of non-zero denominators: never zero
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
// Upper bound of the number of points after one more round: sum_b ceil(L_b / 2) <= (M + #non-empty runs) / 2
static size_t aff_next_bound(size_t m, size_t nb) { return (m + (m < nb ? m : nb) + 1) / 2; }

// Measured on B200 (profiles/r02_msm_affine_ab.txt): a round costs ~0.22 ns per addition against 0.36 ns in XYZZ plus
// ~0.35 ms of fixed latency (scans, inversion tree, one division-step inversion -- 0.9 ms while that inversion was a
// Fermat ladder), and round j holds total / 2^j additions: it pays while that is >= ~3 M.  2^20 points (13.6 M pairs):
// 2 rounds, 2^22: 3-4, 2^24: 5; never more rounds than the mean run is deep.
#ifndef ZKP_AFF_MAX_ROUNDS
#define ZKP_AFF_MAX_ROUNDS 5
#endif
uint32_t msm_affine_choose_rounds(size_t total, size_t total_buckets) {
  if (total_buckets == 0) return 0;
  uint32_t lg = 0;
  for (size_t mean = total / total_buckets; mean > 1; mean >>= 1) lg++;
  uint32_t r = 0;
  while (r < ZKP_AFF_MAX_ROUNDS && r < lg && (total >> (r + 1)) >= ((size_t)3 << 20)) r++;
  return r;
}

size_t msm_affine_bound(size_t total, size_t nb, uint32_t rounds) {
  for (uint32_t r = 0; r < rounds; r++) total = aff_next_bound(total, nb);
  return total;
}

static unsigned aff_blocks(size_t items) { return (unsigned)((items + AFF_THREADS - 1) / AFF_THREADS); }

// Runs `rounds` tree rounds over the bucket-sorted (point | sign) list.  On return *pts is the device array holding the
// shortened runs and *off its run starts ([nb + 1] entries, off[nb] = number of points), *bound an upper bound of that
// number.  Returns ZKP_ERR_OOM without side effects when the buffers do not fit (the caller falls back to rounds = 0).
int msm_affine_rounds_dev(Ctx* ctx, uint32_t rounds, const uint32_t* svals, const G1Affine* bases, uint32_t base_stride16,
                          const uint32_t* bstart,
                          const uint32_t* bend, uint32_t nb, size_t total, const G1Affine** pts, const uint32_t** off,
                          size_t* bound) {
  MsmScratch& m = ctx->msm;
  cudaStream_t st = ctx->stream;
  size_t m1 = aff_next_bound(total, nb), m2 = aff_next_bound(m1, nb);
  const size_t threads1 = (m1 + AFF_K - 1) / AFF_K;
  // inversion tree: level l holds ceil(threads1 / 16^l) values + as many prefixes
  size_t inv_elems = 0;
  for (size_t n = threads1;; n = (n + AFF_INV_K - 1) / AFF_INV_K) {
    inv_elems += 2 * n + 2;
    if (n <= AFF_INV_DIRECT) break;
  }
#ifndef ZKP_EMU
  {  // all-or-nothing: never leave half of the round buffers allocated next to a large table
    const size_t need[5] = {m1 * sizeof(G1Affine), rounds > 1 ? m2 * sizeof(G1Affine) : 0, m1 * sizeof(Fq),
                            inv_elems * sizeof(Fq), 2 * ((size_t)nb + 1) * sizeof(uint32_t)};
    const DevBuf* have[5] = {&m.aff_a, &m.aff_b, &m.aff_pre, &m.aff_inv, &m.aff_off};
    size_t grow = 0;
    for (int i = 0; i < 5; i++)
      if (need[i] > have[i]->cap) grow += need[i];  // reserve() frees the old block before allocating the new one
    size_t free_b = 0, total_b = 0;
    if (grow && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
      size_t reclaim = 0;
      for (int i = 0; i < 5; i++)
        if (need[i] > have[i]->cap) reclaim += have[i]->cap;
      // keep room for the task list / partial sums / reduction levels that are sized after the rounds
      if (grow + ((size_t)6 << 30) > free_b + reclaim) return ZKP_ERR_OOM;
    }
  }
#endif
  ZKP_TRY(m.aff_a.reserve(m1 * sizeof(G1Affine)));
  if (rounds > 1) ZKP_TRY(m.aff_b.reserve(m2 * sizeof(G1Affine)));
  ZKP_TRY(m.aff_pre.reserve(m1 * sizeof(Fq)));
  ZKP_TRY(m.aff_inv.reserve(inv_elems * sizeof(Fq)));
  ZKP_TRY(m.aff_off.reserve(2 * ((size_t)nb + 1) * sizeof(uint32_t)));
  // start-bucket table of the A / C threads + the list of buckets that own more than AFF_TB_SERIAL of them
  const size_t heavy_cap = threads1 / AFF_TB_SERIAL + 1;
  ZKP_TRY(m.aff_tb.reserve((threads1 + 1 + heavy_cap + 1) * sizeof(uint32_t)));
  uint32_t* tb = m.aff_tb.as<uint32_t>();
  uint32_t* tb_heavy = tb + threads1 + 1;
  uint32_t* off_cur = m.aff_off.as<uint32_t>();
  uint32_t* off_nxt = off_cur + (nb + 1);
  ZKP_LAUNCH_NOSYNC(aff_len0_kernel, dim3((nb + 256) / 256), dim3(256), 0, st, bstart, bend, nb, off_cur);
  ZKP_TRY(scan_exclusive_u32_dev(ctx, off_cur, off_cur, nb + 1));
  ctx->msm_launches += 1;
  const G1Affine* in = bases;
  size_t mcur = total;
  for (uint32_t r = 0; r < rounds; r++) {
    const size_t mout = aff_next_bound(mcur, nb);
    G1Affine* out = (r & 1) ? m.aff_b.as<G1Affine>() : m.aff_a.as<G1Affine>();
    ZKP_LAUNCH_NOSYNC(aff_halve_kernel, dim3((nb + 256) / 256), dim3(256), 0, st, (const uint32_t*)off_cur, nb, off_nxt);
    ZKP_TRY(scan_exclusive_u32_dev(ctx, off_nxt, off_nxt, nb + 1));
    AffRound a;
    a.off_in = off_cur;
    a.off_out = off_nxt;
    a.nb = nb;
    a.svals = (r == 0) ? svals : nullptr;
    a.in = in;
    a.in_stride16 = (r == 0) ? base_stride16 : 6u;
    a.out = out;
    a.pre = m.aff_pre.as<Fq>();
    a.stats = nullptr;
    a.tb = tb;
    ZKP_TRY(rt::dev_memset(tb_heavy, 0, sizeof(uint32_t), st));
    ZKP_LAUNCH_NOSYNC(aff_start_bucket_kernel, dim3((nb + 255) / 256), dim3(256), 0, st, (const uint32_t*)off_nxt, nb, tb,
                      tb_heavy, (uint32_t)heavy_cap);
    ZKP_LAUNCH_NOSYNC(aff_start_bucket_heavy_kernel, dim3(256), dim3(256), 0, st, (const uint32_t*)off_nxt, tb,
                      (const uint32_t*)tb_heavy, (uint32_t)heavy_cap);
#ifndef ZKP_EMU
    if (ctx->profiling && r + 1 < (uint32_t)Ctx::AFF_STATS) {
      if (!ctx->aff_stats_dev) ZKP_TRY(rt::dev_malloc((void**)&ctx->aff_stats_dev, Ctx::AFF_STATS * sizeof(uint32_t)));
      a.stats = ctx->aff_stats_dev + r;
    }
#endif
    Fq* level = m.aff_inv.as<Fq>();
    a.tot = level;
    const size_t nthreads = (mout + AFF_K - 1) / AFF_K;
    const uint32_t* mout_ptr = off_nxt + nb;
    if (r == 0) ZKP_LAUNCH_NOSYNC(aff_denominators_kernel<true>, dim3(aff_blocks(nthreads)), dim3(AFF_THREADS), 0, st, a);
    else ZKP_LAUNCH_NOSYNC(aff_denominators_kernel<false>, dim3(aff_blocks(nthreads)), dim3(AFF_THREADS), 0, st, a);
    // B: up-sweeps, one direct inversion, down-sweeps
    struct Lvl { Fq* v; Fq* pre; size_t n; uint32_t div; };
    Lvl lv[8];
    int nl = 0;
    {
      size_t n = nthreads;
      uint32_t div = AFF_K;
      Fq* p = level;
      for (;;) {
        lv[nl].v = p; lv[nl].pre = p + n + 1; lv[nl].n = n; lv[nl].div = div;
        p += 2 * n + 2;
        nl++;
        if (n <= AFF_INV_DIRECT || nl == 8) break;
        n = (n + AFF_INV_K - 1) / AFF_INV_K;
        div *= AFF_INV_K;
      }
    }
    for (int l = 0; l + 1 < nl; l++) {
      const size_t th = (lv[l].n + AFF_INV_K - 1) / AFF_INV_K;
      ZKP_LAUNCH_NOSYNC(aff_inv_up_kernel, dim3(aff_blocks(th)), dim3(AFF_THREADS), 0, st, (const Fq*)lv[l].v, lv[l].pre, lv[l + 1].v,
                        mout_ptr, lv[l].div);
    }
    ZKP_LAUNCH_NOSYNC(aff_inv_direct_kernel, dim3(aff_blocks(lv[nl - 1].n)), dim3(AFF_THREADS), 0, st, lv[nl - 1].v, mout_ptr,
                      lv[nl - 1].div);
    for (int l = nl - 2; l >= 0; l--) {
      const size_t th = (lv[l].n + AFF_INV_K - 1) / AFF_INV_K;
      ZKP_LAUNCH_NOSYNC(aff_inv_down_kernel, dim3(aff_blocks(th)), dim3(AFF_THREADS), 0, st, lv[l].v, (const Fq*)lv[l].pre,
                        (const Fq*)lv[l + 1].v, mout_ptr, lv[l].div);
    }
    if (r == 0) {
#ifndef ZKP_EMU
      if (ctx->profiling) {
        for (int e = 0; e < 2; e++)
          if (!ctx->aff_ev[e]) { cudaEvent_t ev; if (cudaEventCreate(&ev) == cudaSuccess) ctx->aff_ev[e] = (void*)ev; }
        if (ctx->aff_ev[0]) cudaEventRecord((cudaEvent_t)ctx->aff_ev[0], st);
      }
#endif
      ZKP_LAUNCH_NOSYNC(aff_add_kernel<true>, dim3(aff_blocks(nthreads)), dim3(AFF_THREADS), 0, st, a);
#ifndef ZKP_EMU
      if (ctx->profiling && ctx->aff_ev[1]) cudaEventRecord((cudaEvent_t)ctx->aff_ev[1], st);
#endif
    } else {
      ZKP_LAUNCH_NOSYNC(aff_add_kernel<false>, dim3(aff_blocks(nthreads)), dim3(AFF_THREADS), 0, st, a);
    }
    ctx->msm_launches += 5 + 2 * (uint32_t)(nl - 1) + 1;
    in = out;
    mcur = mout;
    uint32_t* t = off_cur; off_cur = off_nxt; off_nxt = t;
  }
  *pts = in;
  *off = off_cur;
  *bound = mcur;
  return rt::check_last();
}

}  // namespace zkp
