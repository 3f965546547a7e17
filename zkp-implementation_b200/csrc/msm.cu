// BLS12-381 G1 multi-scalar multiplication: signed-digit Pippenger on the device.
//
// Replaces the hot loops of `KzgScheme::evaluate_in_s` (kzg/src/scheme.rs:84-96):
//     coeffs.zip(points).map(|(c, s)| s.mul(c).into_affine()).reduce(|a, e| a.add(e).into_affine())
// i.e. sum_i c_i * P_i returned as the normalised affine point.  A group element has exactly one
// normalised affine representative, so the result is bit-identical to the reference's per-term
// double-and-add no matter how the sum is scheduled.
//
// Pipeline (all on one stream):
//   1. recode      scalars: Montgomery -> canonical -> W signed c-bit digits; one (bucket, point|sign)
//                  pair per digit
//   2. sort        ONE radix sort of all n*W pairs by global bucket id (sort.cu: stable LSD, 8-bit digits)
//   3. boundaries  start/end of every bucket's run in the sorted order
//   4. tasks       buckets longer than S_max are split so no thread owns an unbounded run; the task
//                  list is sorted by run length (longest first) so the 32 lanes of a warp finish together
//   5. accumulate  one thread per task: XYZZ accumulator += affine base (mixed add, 8M+2S),
//                  next base prefetched while the current add runs
//   6. reduce      sum_b (b+1) B_b per bucket set by bit planes of the bucket index: level l halves the array
//                  (one independent addition per thread) and A_l = sum of the odd entries of level l, summed by
//                  pairwise trees that advance in the same steps, so the serial depth is c - 1 additions + the
//                  combine however many buckets there are
//   7. host        windowed mode only: Horner over the W window sums; the caller normalises (one inversion)
//
// Two modes.  WINDOWED (ad-hoc bases): window w has its own 2^(c-1) buckets.  FIXED-BASE (the resident SRS
// after zkp_srs_precompute): the table holds 2^(c w) P_i for every window, so digit w of scalar i selects
// table entry (w, i) and ALL windows share one bucket set -- W times fewer buckets to reduce, no Horner, and
// the cheaper reduction lets c grow by a few bits (fewer windows, fewer additions).
#include <string.h>

// This translation unit's Fq products are calls of one out-of-line multiplier (field.cuh): its kernels that matter for
// small MSMs are short launches whose straight-line code was fetched cold (2^16: 1.69 -> 1.44 ms; 2^20 / 2^24 unchanged).
#ifndef ZKP_FQ_INLINE
#define ZKP_FQ_CALL 1
#endif

#include "engine.h"
#include "memops.cuh"

#if defined(ZKP_USE_CUB) && !defined(ZKP_EMU)
#include <cub/cub.cuh>
#endif
#ifdef ZKP_EMU
#include <algorithm>
#include <numeric>
#endif

namespace zkp {

static constexpr uint32_t ACC_THREADS = 128;
#ifdef ZKP_EMU
static constexpr uint32_t RED_THREADS = 32;   // emulated build: one OS thread per CUDA thread, keep the blocks small
#else
static constexpr uint32_t RED_THREADS = 128;
#endif
#ifndef ZKP_RED_MIN_BLOCKS
#define ZKP_RED_MIN_BLOCKS 1
#endif
#ifndef ZKP_ACC_MIN_BLOCKS
#define ZKP_ACC_MIN_BLOCKS 1
#endif
static constexpr uint32_t SIGN_BIT = 0x80000000u;

struct MsmTask {
  uint32_t start;  // index into the window-major sorted value array
  uint32_t len;
};

// ---- 1. recode ---------------------------------------------------------------------------------
// keys[w*n + i] = global bucket id of digit w of scalar i (weight |digit| = id within the set + 1), or the
//                 sentinel `total_buckets` for a zero digit
// vals[w*n + i] = base index | sign << 31; fixed-base mode: index (w * table_stride + i) into the table
// `set`: fixed-base batches run several MSMs in one pipeline, MSM j owning bucket set j; `sentinel` = total
// number of buckets of the whole batch.
__global__ void __launch_bounds__(256) msm_recode_kernel(const Fr* __restrict__ scalars, uint32_t n, uint32_t c,
                                                         uint32_t nwin, uint32_t table_stride, uint32_t fixed,
                                                         uint32_t set, uint32_t sentinel, uint32_t* __restrict__ keys,
                                                         uint32_t* __restrict__ vals) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* sp = reinterpret_cast<const uint4*>(scalars + i);
  uint4 a = sp[0], b = sp[1];
  Fr s;
  s.v[0] = a.x; s.v[1] = a.y; s.v[2] = a.z; s.v[3] = a.w;
  s.v[4] = b.x; s.v[5] = b.y; s.v[6] = b.z; s.v[7] = b.w;
  s = fp_from_mont(s);  // `cof.into_bigint()` inside ark-ec's scalar mul (scheme.rs:92)
  const uint32_t nbuckets = 1u << (c - 1);
  const uint32_t cmask = (1u << c) - 1;
  uint32_t carry = 0;
  for (uint32_t w = 0; w < nwin; w++) {
    const uint32_t bit = w * c;
    const uint32_t limb = bit >> 5, off = bit & 31;
    uint32_t raw = 0;
    if (limb < 8) {
      uint64_t two = s.v[limb];
      if (limb + 1 < 8) two |= (uint64_t)s.v[limb + 1] << 32;
      raw = (uint32_t)(two >> off) & cmask;
    }
    uint32_t d = raw + carry;
    uint32_t neg = 0;
    if (d > nbuckets) {  // digit in (2^(c-1), 2^c] -> d - 2^c in (-2^(c-1), 0]
      d = (1u << c) - d;
      neg = 1;
      carry = 1;
    } else {
      carry = 0;
    }
    const size_t o = (size_t)w * n + i;
    keys[o] = d ? ((fixed ? set * nbuckets : w * nbuckets) + d - 1) : sentinel;
    vals[o] = ((fixed ? w * table_stride : 0u) + i) | (neg && d ? SIGN_BIT : 0u);
  }
}

// ---- 3. bucket boundaries ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) msm_bounds_kernel(const uint32_t* __restrict__ skeys, size_t total,
                                                         uint32_t total_buckets, uint32_t* __restrict__ bstart,
                                                         uint32_t* __restrict__ bend) {
  // four keys per thread and iteration from one 128-bit load (the scalar form kept one 4-byte load per thread in flight and
  // ran at a fifth of the HBM rate); the neighbours across the group's edges are two extra, mostly cached, loads
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t nvec = total >> 2;
  for (size_t v = gid; v < nvec; v += stride) {
    const uint4 q = reinterpret_cast<const uint4*>(skeys)[v];
    const size_t i0 = v << 2;
    const uint32_t k[4] = {q.x, q.y, q.z, q.w};
    const uint32_t prev = i0 ? skeys[i0 - 1] : ~k[0];
    const uint32_t next = (i0 + 4 < total) ? skeys[i0 + 4] : ~k[3];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t key = k[j];
      if (key >= total_buckets) continue;
      if ((j ? k[j - 1] : prev) != key) bstart[key] = (uint32_t)(i0 + j);
      if ((j < 3 ? k[j + 1] : next) != key) bend[key] = (uint32_t)(i0 + j) + 1;
    }
  }
  for (size_t idx = (nvec << 2) + gid; idx < total; idx += stride) {  // the last total mod 4 keys
    const uint32_t key = skeys[idx];
    if (key >= total_buckets) continue;
    if (idx == 0 || skeys[idx - 1] != key) bstart[key] = (uint32_t)idx;
    if (idx == total - 1 || skeys[idx + 1] != key) bend[key] = (uint32_t)idx + 1;
  }
}

// ---- 4. tasks ------------------------------------------------------------------------------------
// heavy[0] = number of buckets split into more than HEAVY_TASKS tasks, heavy[1 + k] = their ids: their
// partial sums are folded by a whole block each (msm_heavy_fold_kernel) before the reduction reads them
static constexpr uint32_t HEAVY_TASKS = 24;
static constexpr uint32_t HEAVY_CAP = 1u << 16;

__global__ void __launch_bounds__(256) msm_task_count_kernel(const uint32_t* __restrict__ bstart,
                                                             const uint32_t* __restrict__ bend, uint32_t total_buckets,
                                                             uint32_t smax, uint32_t* __restrict__ ntask,
                                                             uint32_t* __restrict__ heavy) {
  const uint32_t gb = blockIdx.x * blockDim.x + threadIdx.x;
  if (gb >= total_buckets) return;
  const uint32_t len = bend[gb] - bstart[gb];
  const uint32_t nt = (len + smax - 1) / smax;
  ntask[gb] = nt;
  if (nt > HEAVY_TASKS) {
    const uint32_t slot = atomicAdd(heavy, 1u);
    if (slot < HEAVY_CAP) heavy[1 + slot] = gb;
  }
}

// One block per heavy bucket: partials[off] = sum of its nt partials (the reduction then reads one entry).
__global__ void __launch_bounds__(RED_THREADS) msm_heavy_fold_kernel(const uint32_t* __restrict__ heavy,
                                                             const uint32_t* __restrict__ task_off,
                                                             const uint32_t* __restrict__ ntask,
                                                             G1Xyzz* __restrict__ partials, uint32_t* __restrict__ nfold) {
  __shared__ G1Xyzz sh[RED_THREADS];
  const uint32_t tid = threadIdx.x;
  uint32_t count = heavy[0];
  if (count > HEAVY_CAP) count = HEAVY_CAP;
  for (uint32_t h = blockIdx.x; h < count; h += gridDim.x) {
    const uint32_t gb = heavy[1 + h];
    const uint32_t nt = ntask[gb], off = task_off[gb];
    G1Xyzz acc = G1Xyzz::infinity();
    for (uint32_t k = tid; k < nt; k += blockDim.x) {
      G1Xyzz p = ld_xyzz(partials + off + k);
      xyzz_add(acc, p);
    }
    st_xyzz(&sh[tid], acc);
    __syncthreads();
    for (uint32_t s = blockDim.x / 2; s > 0; s >>= 1) {
      if (tid < s) {
        G1Xyzz a = ld_xyzz(&sh[tid]), b = ld_xyzz(&sh[tid + s]);
        xyzz_add(a, b);
        st_xyzz(&sh[tid], a);
      }
      __syncthreads();
    }
    if (tid == 0) {
      st_xyzz(partials + off, ld_xyzz(&sh[0]));
      nfold[gb] = 1;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) msm_task_build_kernel(const uint32_t* __restrict__ bstart,
                                                             const uint32_t* __restrict__ bend,
                                                             const uint32_t* __restrict__ task_off, uint32_t total_buckets,
                                                             uint32_t smax, MsmTask* __restrict__ tasks,
                                                             uint32_t* __restrict__ task_len, uint32_t* __restrict__ task_id) {
  const uint32_t gb = blockIdx.x * blockDim.x + threadIdx.x;
  if (gb >= total_buckets) return;
  uint32_t start = bstart[gb];
  const uint32_t end = bend[gb];
  uint32_t t = task_off[gb];
  while (start < end) {
    const uint32_t len = (end - start < smax) ? (end - start) : smax;
    tasks[t].start = start;
    tasks[t].len = len;
    task_len[t] = len;  // sort key: threads of a warp get runs of (nearly) equal length
    task_id[t] = t;
    t++;
    start += len;
  }
}

// ---- 5. accumulate -------------------------------------------------------------------------------
// DIRECT = false: the run is a slice of the sorted (point index | sign) list and the points are gathered from `bases`;
// DIRECT = true: the run is a slice of `bases` itself (the output of the batched-affine tree rounds, msm_affine.cu).
template <bool DIRECT>
__global__ void __launch_bounds__(ACC_THREADS, ZKP_ACC_MIN_BLOCKS) msm_accumulate_kernel(const MsmTask* __restrict__ tasks,
                                                                     const uint32_t* __restrict__ order,
                                                                     const uint32_t* __restrict__ ntasks_ptr,
                                                                     const uint32_t* __restrict__ svals,
                                                                     const G1Affine* __restrict__ bases, uint32_t stride16,
                                                                     G1Xyzz* __restrict__ partials) {
  // The grid covers the host's upper bound of the task count; the real count stays on the device (no mid-pipeline
  // read-back): the length-sorted order puts the zero-length padding slots last.
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= *ntasks_ptr) return;
  const uint32_t t = order[slot];  // longest runs first; partials stay in bucket order
  const MsmTask tk = tasks[t];
  const uint32_t* v = DIRECT ? nullptr : svals + tk.start;
  const G1Affine* run = bases + tk.start;
  G1Xyzz acc = G1Xyzz::infinity();
  uint32_t pv = DIRECT ? 0u : v[0];
  G1Affine nxt = DIRECT ? ld_affine(run) : ld_affine_s(bases, pv & ~SIGN_BIT, stride16);
  for (uint32_t i = 0; i < tk.len; i++) {
    G1Affine cur = nxt;
    const uint32_t cv = pv;
    if (i + 1 < tk.len) {
      if (DIRECT) {
        nxt = ld_affine(run + i + 1);
      } else {
        pv = v[i + 1];
        nxt = ld_affine_s(bases, pv & ~SIGN_BIT, stride16);
      }
    }
    if (cv & SIGN_BIT) cur = g1_neg(cur);
    xyzz_madd(acc, cur);
  }
  st_xyzz(partials + t, acc);
}

// ---- 6. reduce -----------------------------------------------------------------------------------
// F = sum_b (b + 1) B_b over one bucket set of m = 2^(c-1) buckets, by bit planes of the bucket index:
//   F = G + sum_l 2^l A_l,   G = sum_b B_b,   A_l = sum of the buckets whose index has bit l set.
// Level l halves the array: X^(l+1)[g] = X^l[2g] + X^l[2g+1] (X^0 = buckets), and A_l is the sum of the odd
// entries of X^l.  Every level is one independent addition per thread, so the serial depth is c - 1 additions
// however many buckets there are -- the reduction stays short for the small bucket sets of small MSMs and is
// throughput-bound (2m additions in all) for the large ones.
static constexpr uint32_t MAX_RED_LEVELS = 26;

// X^0[gb] = sum of the bucket's partial sums (tasks of a split bucket; heavily split ones were folded first)
__global__ void __launch_bounds__(RED_THREADS, ZKP_RED_MIN_BLOCKS) msm_bucket_gather_kernel(const G1Xyzz* __restrict__ partials,
                                                                        const uint32_t* __restrict__ task_off,
                                                                        const uint32_t* __restrict__ nfold,
                                                                        uint32_t total_buckets, G1Xyzz* __restrict__ x0) {
  const uint32_t gb = blockIdx.x * blockDim.x + threadIdx.x;
  if (gb >= total_buckets) return;
  const uint32_t nt = nfold[gb], off = task_off[gb];
  G1Xyzz acc = G1Xyzz::infinity();
  for (uint32_t k = 0; k < nt; k++) {
    G1Xyzz p = ld_xyzz(partials + off + k);
    xyzz_add(acc, p);
  }
  st_xyzz(x0 + gb, acc);
}

// The tree.  Step L (L = 0 .. levels - 1) holds (L + 1) * (m >> (L + 1)) independent additions per set, all ONE addition deep:
//   main     X^(L+1)[j] = X^L[2j] + X^L[2j+1]
//   plane p  (p < L) P_p^(L+1)[j] = P_p^L[2j] + P_p^L[2j+1], where P_p^(p+1)[j] = X^p[2j+1] is read in place -- the pairwise
//            tree over the odd entries of X^p, one level per step in lockstep with the main tree.
// After the last step X^levels[0] = G and P_p^levels[0] = A_p (A_(levels-1) = X^(levels-1)[1] is never added to anything),
// so the serial depth of the whole reduction is `levels` additions + the combine, 2m additions in all (as before: the
// plane sums used to be separate chunked sums + a second stage + a combine launch, ~3x the depth).
// Wide steps are one launch each; once a step has <= TAIL_ITEMS additions per set the rest runs in ONE launch, one block
// per set with a block barrier per step, and the same block finishes F = G + sum_l 2^l A_l (thread l doubles A_l l times,
// then a tree over the planes).
struct RedTree {
  uint32_t off[MAX_RED_LEVELS + 1];   // element offset of X^l, layout [set][m >> l]; off[levels] = X^levels (one entry per set)
  uint32_t poff[MAX_RED_LEVELS + 1];  // plane p's ping-pong buffer, layout [2][set][h_p], h_p = max(m >> (p + 2), 1)
  uint32_t m, levels, nsets;
};
#ifdef ZKP_EMU
static constexpr uint32_t TAIL_THREADS = 32;
#else
static constexpr uint32_t TAIL_THREADS = 256;
#endif
static constexpr uint32_t TAIL_ITEMS = 2 * TAIL_THREADS;

__device__ __forceinline__ uint32_t tree_plane_len(const RedTree& t, uint32_t p) {
  const uint32_t h = t.m >> (p + 2);
  return h ? h : 1u;
}

// item (p, j) of set w at step L: p == L is the main tree, p < L plane p
__device__ __forceinline__ void tree_item(G1Xyzz* __restrict__ buf, const RedTree& t, uint32_t L, uint32_t w, uint32_t p,
                                          uint32_t j) {
  const G1Xyzz *a, *b;
  G1Xyzz* dst;
  if (p == L) {
    const G1Xyzz* in = buf + t.off[L] + (size_t)w * (t.m >> L);
    a = in + 2 * (size_t)j;
    b = a + 1;
    dst = buf + t.off[L + 1] + (size_t)w * (t.m >> (L + 1)) + j;
  } else {
    const uint32_t hp = tree_plane_len(t, p);
    G1Xyzz* pb = buf + t.poff[p];
    const uint32_t oh = (L - p - 1) & 1u;
    dst = pb + ((size_t)oh * t.nsets + w) * hp + j;
    if (L == p + 1) {
      const G1Xyzz* in = buf + t.off[p] + (size_t)w * (t.m >> p);
      a = in + 4 * (size_t)j + 1;
      b = a + 2;
    } else {
      const G1Xyzz* in = pb + ((size_t)(oh ^ 1u) * t.nsets + w) * hp;
      a = in + 2 * (size_t)j;
      b = a + 1;
    }
  }
  G1Xyzz x = ld_xyzz(a);
  const G1Xyzz y = ld_xyzz(b);
  xyzz_add(x, y);
  st_xyzz(dst, x);
}

__global__ void __launch_bounds__(RED_THREADS, ZKP_RED_MIN_BLOCKS) msm_tree_step_kernel(G1Xyzz* __restrict__ buf, RedTree t,
                                                                                       uint32_t L) {
  const uint32_t half_log = t.levels - L - 1;
  const uint32_t per_set = (L + 1) << half_log;
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (uint64_t)per_set * t.nsets) return;
  const uint32_t w = (uint32_t)(idx / per_set);
  const uint32_t rem = (uint32_t)(idx - (uint64_t)w * per_set);
  tree_item(buf, t, L, w, rem >> half_log, rem & ((1u << half_log) - 1));
}

__global__ void __launch_bounds__(TAIL_THREADS) msm_tree_tail_kernel(G1Xyzz* __restrict__ buf, RedTree t, uint32_t L0,
                                                                     G1Xyzz* __restrict__ out) {
  __shared__ G1Xyzz sh[32];
  const uint32_t w = blockIdx.x, tid = threadIdx.x;
  for (uint32_t L = L0; L < t.levels; L++) {
    const uint32_t half_log = t.levels - L - 1;
    const uint32_t per_set = (L + 1) << half_log;
    for (uint32_t i = tid; i < per_set; i += blockDim.x) tree_item(buf, t, L, w, i >> half_log, i & ((1u << half_log) - 1));
    __syncthreads();
  }
  // F[w] = G[w] + sum_l 2^l A_l[w]
  if (tid < 32) {
    G1Xyzz acc = G1Xyzz::infinity();
    if (tid + 1 < t.levels) {  // A_l = P_l^levels[0]: written by the last step into half (levels - l) & 1
      const uint32_t oh = (t.levels - tid) & 1u;
      acc = ld_xyzz(buf + t.poff[tid] + ((size_t)oh * t.nsets + w) * tree_plane_len(t, tid));
    } else if (tid + 1 == t.levels) {
      acc = ld_xyzz(buf + t.off[tid] + (size_t)w * 2 + 1);
    } else if (tid == t.levels) {
      acc = ld_xyzz(buf + t.off[t.levels] + w);
    }
    if (tid < t.levels)
      for (uint32_t k = 0; k < tid; k++) acc = xyzz_dbl(acc);
    st_xyzz(&sh[tid], acc);
  }
  __syncthreads();
  for (uint32_t s = 16; s > 0; s >>= 1) {
    if (tid < s) {
      G1Xyzz a = ld_xyzz(&sh[tid]), b = ld_xyzz(&sh[tid + s]);
      xyzz_add(a, b);
      st_xyzz(&sh[tid], a);
    }
    __syncthreads();
  }
  if (tid == 0) st_xyzz(out + w, ld_xyzz(&sh[0]));
}

// 2^c * P for every point of one table window (fixed-base precomputation)
// table window (records stride16 x 16 bytes apart) <- packed points
__global__ void __launch_bounds__(256) msm_table_store_kernel(const G1Affine* __restrict__ in, size_t n, G1Affine* __restrict__ tab,
                                                              size_t first, uint32_t stride16) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st_affine_s(tab, first + i, stride16, ld_affine(in + i));
}

__global__ void __launch_bounds__(128) msm_shift_window_kernel(const G1Affine* __restrict__ in, size_t n, uint32_t c,
                                                               G1Xyzz* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const G1Affine p = ld_affine(in + i);
  G1Xyzz acc = G1Xyzz::infinity();
  if (!p.is_inf()) {
    acc = xyzz_dbl_affine(p);
    for (uint32_t k = 1; k < c; k++) acc = xyzz_dbl(acc);
  }
  st_xyzz(out + i, acc);
}

// ------------------------------------------------------------------------------------------------
// Host orchestration
// ------------------------------------------------------------------------------------------------
// cost model in mixed-add units: n*W additions + the reduction's cost per bucket.  Fixed-base mode: measured 1.27 ns per
// bucket on top of a size-independent part (reduce phase 1.5 ms at 2^17 buckets, 4.0 ms at 2^21; 1.3 / 3.5 ms with the fused
// tree, scripts/window_sweep.py) = 3 additions of
// 0.36 ns; windowed mode keeps the round-1 figure (every window has its own bucket set and the Horner tail).
static uint32_t choose_window_bits(size_t n, bool fixed) {
  uint32_t best = 4;
  double best_cost = 1e300;
  const uint32_t cmax = fixed ? 24 : 20;
  const double per_bucket = fixed ? 3.0 : 8.0;
  for (uint32_t c = 4; c <= cmax; c++) {
    const uint32_t W = 255 / c + 1;
    const double nb = (double)(1u << (c - 1)) * (fixed ? 1 : W);
    double cost = (double)n * W + per_bucket * nb;
    // a top window of only a few bits (c = 14: 3, c = 15: 0 + the carry) piles all n points into a handful of buckets: long
    // runs to split and fold (2^14..2^16: c = 14 / 15 measured 10-20 % slower than c = 13 / 16)
    if (fixed && 255 - (int)(c * (W - 1)) < 6) cost += (double)n;
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}

// The pair sort and the scan are the engine's own kernels (sort.cu).  -DZKP_USE_CUB swaps in cub::DeviceRadixSort /
// cub::DeviceScan for A/B measurements (profiles/r01_sort_ab.txt).
#if defined(ZKP_USE_CUB) && !defined(ZKP_EMU)
static int sort_pairs(Ctx* ctx, uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, uint32_t n, uint32_t key_bits,
                      bool descending, uint32_t** kres, uint32_t** vres) {
  size_t tmp = 0;
  cudaError_t e = descending ? cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, k0, k1, v0, v1, (int)n, 0, (int)key_bits, ctx->stream)
                             : cub::DeviceRadixSort::SortPairs(nullptr, tmp, k0, k1, v0, v1, (int)n, 0, (int)key_bits, ctx->stream);
  if (e != cudaSuccess) return rt::wrap(e);
  ZKP_TRY(ctx->msm.sort_tmp.reserve(tmp));
  tmp = ctx->msm.sort_tmp.cap;
  e = descending ? cub::DeviceRadixSort::SortPairsDescending(ctx->msm.sort_tmp.p, tmp, k0, k1, v0, v1, (int)n, 0, (int)key_bits, ctx->stream)
                 : cub::DeviceRadixSort::SortPairs(ctx->msm.sort_tmp.p, tmp, k0, k1, v0, v1, (int)n, 0, (int)key_bits, ctx->stream);
  *kres = k1;
  *vres = v1;
  return rt::wrap(e);
}
static int exclusive_scan_u32(Ctx* ctx, const uint32_t* in, uint32_t* out, uint32_t n) {
  size_t tmp = 0;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int)n, ctx->stream);
  if (e != cudaSuccess) return rt::wrap(e);
  ZKP_TRY(ctx->msm.sort_tmp.reserve(tmp));
  tmp = ctx->msm.sort_tmp.cap;
  e = cub::DeviceScan::ExclusiveSum(ctx->msm.sort_tmp.p, tmp, in, out, (int)n, ctx->stream);
  return rt::wrap(e);
}
#elif defined(ZKP_EMU)
// Emulated build: the radix-sort kernels are exercised directly (tests/test_sort.py); inside the MSM pipeline the
// emulator groups with std::stable_sort -- the same stable order -- because a ballot costs two OS-thread barriers there.
static int sort_pairs(Ctx* ctx, uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, uint32_t n, uint32_t key_bits,
                      bool descending, uint32_t** kres, uint32_t** vres) {
  (void)ctx;
  const uint32_t mask = key_bits >= 32 ? 0xffffffffu : ((1u << key_bits) - 1u);
  std::vector<uint32_t> idx(n);
  std::iota(idx.begin(), idx.end(), 0u);
  std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {
    return descending ? (k0[a] & mask) > (k0[b] & mask) : (k0[a] & mask) < (k0[b] & mask);
  });
  for (uint32_t i = 0; i < n; i++) { k1[i] = k0[idx[i]]; v1[i] = v0[idx[i]]; }
  *kres = k1;
  *vres = v1;
  return ZKP_OK;
}
#else
static int sort_pairs(Ctx* ctx, uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, uint32_t n, uint32_t key_bits,
                      bool descending, uint32_t** kres, uint32_t** vres) {
  return radix_sort_pairs_dev(ctx, k0, v0, k1, v1, n, key_bits, descending, kres, vres);
}
#endif
#if !(defined(ZKP_USE_CUB) && !defined(ZKP_EMU))
static int exclusive_scan_u32(Ctx* ctx, const uint32_t* in, uint32_t* out, uint32_t n) {
  return scan_exclusive_u32_dev(ctx, in, out, n);
}
#endif

static void phase_mark(Ctx* ctx, int i) {
#ifndef ZKP_EMU
  if (!ctx->profiling) return;
  if (!ctx->phase_ev[i]) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    ctx->phase_ev[i] = (void*)e;
  }
  cudaEventRecord((cudaEvent_t)ctx->phase_ev[i], ctx->stream);
#else
  (void)ctx; (void)i;
#endif
}

static void phase_collect(Ctx* ctx) {
#ifndef ZKP_EMU
  if (!ctx->profiling) return;
  for (int i = 0; i < Ctx::NPHASE; i++) {
    float ms = -1;
    if (ctx->phase_ev[i] && ctx->phase_ev[i + 1] &&
        cudaEventElapsedTime(&ms, (cudaEvent_t)ctx->phase_ev[i], (cudaEvent_t)ctx->phase_ev[i + 1]) != cudaSuccess)
      ms = -1;
    ctx->phase_ms[i] = ms;
  }
  ctx->aff_add1_ms = -1;
  memset(ctx->aff_stats, 0, sizeof(ctx->aff_stats));
  if (ctx->last_affine_rounds && ctx->aff_ev[0] && ctx->aff_ev[1]) {
    float ms = -1;
    if (cudaEventElapsedTime(&ms, (cudaEvent_t)ctx->aff_ev[0], (cudaEvent_t)ctx->aff_ev[1]) == cudaSuccess) ctx->aff_add1_ms = ms;
    if (ctx->aff_stats_dev)  // the stream has been synchronised by the caller
      cudaMemcpy(ctx->aff_stats, ctx->aff_stats_dev, sizeof(ctx->aff_stats), cudaMemcpyDeviceToHost);
  }
#else
  (void)ctx;
#endif
}

int msm_run_dev(Ctx* ctx, const Fr* scalars, const G1Affine* bases, size_t n, G1Xyzz* out_host, uint32_t fixed_c,
                size_t table_stride) {
  return msm_run_multi_dev(ctx, &scalars, &n, 1, bases, out_host, fixed_c, table_stride);
}

// `fixed_c != 0`: bases is a precomputed table of nwin windows x table_stride points (msm_precompute_dev) and
// `count` MSMs over it run as one pipeline (MSM j = bucket set j).  Windowed mode takes count == 1.
int msm_run_multi_dev(Ctx* ctx, const Fr* const* scalars_list, const size_t* n_list, uint32_t count, const G1Affine* bases,
                      G1Xyzz* out_host, uint32_t fixed_c, size_t table_stride) {
  ctx->msm_launches = 0;
  ctx->sort_launches = 0;
  const bool fixed = fixed_c != 0;
  const uint32_t base_stride16 = fixed ? ctx->srs_tab_stride16 : 6u;  // only the fixed-base table has padded records
  if (count == 0) return ZKP_OK;
  if (!fixed && count != 1) return ZKP_ERR_INVALID_ARG;
  size_t n_sum = 0, n_max = 0;
  for (uint32_t j = 0; j < count; j++) {
    out_host[j] = G1Xyzz::infinity();
    if (n_list[j] >= ((size_t)1 << 28)) return ZKP_ERR_INVALID_ARG;
    n_sum += n_list[j];
    if (n_list[j] > n_max) n_max = n_list[j];
  }
  if (n_sum == 0) return ZKP_OK;  // scheme.rs:94 unwrap_or(G1Point::zero())
  const uint32_t c = fixed ? fixed_c : (ctx->msm_window_bits ? ctx->msm_window_bits : choose_window_bits(n_max, false));
  const uint32_t nwin = 255 / c + 1;
  const uint32_t nbuckets = 1u << (c - 1);
  const uint32_t nsets = fixed ? count : nwin;
  if ((size_t)nbuckets * nsets >= ((size_t)1 << 31)) return ZKP_ERR_INVALID_ARG;
  const uint32_t total_buckets = nbuckets * nsets;
  const size_t total = n_sum * nwin;
  if (total >= ((size_t)1 << 31)) return ZKP_ERR_INVALID_ARG;
  if (fixed && (size_t)nwin * table_stride >= ((size_t)1 << 31)) return ZKP_ERR_INVALID_ARG;
  // reduction levels: X^l has nbuckets >> l entries per set, l = 0 .. c - 1 (the last one is G)
  const uint32_t levels = c - 1;
  if (levels + 1 > 32 || levels > MAX_RED_LEVELS) return ZKP_ERR_INVALID_ARG;
  const size_t lvl_elems = 2 * (size_t)nbuckets;  // sum over l of nbuckets >> l, per set

  MsmScratch& m = ctx->msm;
  unsigned key_bits = 1;
  while ((1ull << key_bits) <= total_buckets) key_bits++;  // sentinel = total_buckets must be representable
  ZKP_TRY(m.keys_a.reserve(total * 4));
  ZKP_TRY(m.keys_b.reserve(total * 4));
  ZKP_TRY(m.vals_a.reserve(total * 4));
  ZKP_TRY(m.vals_b.reserve(total * 4));
  ZKP_TRY(m.bucket_start.reserve((size_t)total_buckets * 4));
  ZKP_TRY(m.bucket_end.reserve((size_t)total_buckets * 4));
  ZKP_TRY(m.misc.reserve((size_t)(total_buckets + 1) * 12 + (HEAVY_CAP + 1) * 4));
  // plane ping-pong buffers of the reduction tree: 2 * max(nbuckets >> (p + 2), 1) entries per set and plane
  size_t plane_elems = 0;
  for (uint32_t p = 0; p + 1 < levels; p++) plane_elems += 2 * (size_t)((nbuckets >> (p + 2)) ? (nbuckets >> (p + 2)) : 1);
  if ((lvl_elems + plane_elems) * nsets >= ((size_t)1 << 32)) return ZKP_ERR_INVALID_ARG;
  ZKP_TRY(m.seg_out.reserve((lvl_elems + plane_elems) * nsets * sizeof(G1Xyzz)));
  ZKP_TRY(m.win_out.reserve((size_t)(nsets + 1) * sizeof(G1Xyzz)));
  uint32_t* keys_a = m.keys_a.as<uint32_t>();
  uint32_t* keys_b = m.keys_b.as<uint32_t>();
  uint32_t* vals_a = m.vals_a.as<uint32_t>();
  uint32_t* vals_b = m.vals_b.as<uint32_t>();
  uint32_t* bstart = m.bucket_start.as<uint32_t>();
  uint32_t* bend = m.bucket_end.as<uint32_t>();
  uint32_t* ntask = m.misc.as<uint32_t>();
  uint32_t* task_off = ntask + (total_buckets + 1);
  uint32_t* nfold = task_off + (total_buckets + 1);  // ntask after heavy buckets were folded to one partial
  uint32_t* heavy = nfold + (total_buckets + 1);
  G1Xyzz* lvl_buf = m.seg_out.as<G1Xyzz>();
  G1Xyzz* win_out = m.win_out.as<G1Xyzz>();  // [nsets]
  cudaStream_t st = ctx->stream;

  ctx->last_window_bits = c;
  ctx->last_windows = nwin;
  // 1. recode
  phase_mark(ctx, 0);
  {
    size_t off = 0;
    for (uint32_t j = 0; j < count; j++) {
      const uint32_t n = (uint32_t)n_list[j];
      if (!n) continue;
      ZKP_LAUNCH_NOSYNC(msm_recode_kernel, dim3((n + 255) / 256), dim3(256), 0, st, scalars_list[j], n, c, nwin,
                 (uint32_t)table_stride, fixed ? 1u : 0u, j, total_buckets, keys_a + off, vals_a + off);
      ctx->msm_launches++;
      off += (size_t)n * nwin;
    }
  }
  phase_mark(ctx, 1);
  // 2. one sort of every (bucket, point) pair by global bucket id
  uint32_t *skeys = nullptr, *svals = nullptr;
  ZKP_TRY(sort_pairs(ctx, keys_a, vals_a, keys_b, vals_b, (uint32_t)total, key_bits, false, &skeys, &svals));

  phase_mark(ctx, 2);
  // 3. boundaries
  ZKP_TRY(rt::dev_memset(bstart, 0, (size_t)total_buckets * 4, st));
  ZKP_TRY(rt::dev_memset(bend, 0, (size_t)total_buckets * 4, st));
  {
    size_t blocks = (total / 4 + 256) / 256;
    const size_t cap = (size_t)ctx->sm_count * 32;
    if (blocks > cap) blocks = cap;
    ZKP_LAUNCH_NOSYNC(msm_bounds_kernel, dim3((unsigned)blocks), dim3(256), 0, st, (const uint32_t*)skeys, total, total_buckets,
                      bstart, bend);
    ctx->msm_launches++;
  }
  phase_mark(ctx, 3);
  // 3b. batched-affine tree rounds (msm_affine.cu): every run of L points becomes ~L / 2^R points at ~6 field products
  // per addition instead of 10; what is left is accumulated in XYZZ below
  uint32_t rounds = ctx->msm_affine_rounds >= 0 ? (uint32_t)ctx->msm_affine_rounds
                                                : msm_affine_choose_rounds(total, total_buckets);
  const uint32_t* run_start = bstart;
  const uint32_t* run_end = bend;
  const G1Affine* run_pts = nullptr;
  size_t total_fin = total;
  if (rounds) {
    const uint32_t* off = nullptr;
    const int as = msm_affine_rounds_dev(ctx, rounds, svals, bases, base_stride16, bstart, bend, total_buckets, total, &run_pts,
                                         &off, &total_fin);
    if (as == ZKP_ERR_OOM) {  // the round buffers do not fit next to the table: plain XYZZ accumulation
      rounds = 0;
      run_pts = nullptr;
      total_fin = total;
#ifndef ZKP_EMU
      cudaGetLastError();
#endif
    } else {
      ZKP_TRY(as);
      run_start = off;
      run_end = off + 1;
    }
  }
  ctx->last_affine_rounds = rounds;
  // bound the longest run one thread owns: 4x the mean bucket load, but never so long that fewer than ~128 K
  // tasks exist (few, heavily loaded buckets), and at least 32
  uint32_t smax = (uint32_t)((4 * total_fin) / total_buckets);
  if (smax > total_fin / 131072) smax = (uint32_t)(total_fin / 131072);
  if (smax < 32) smax = 32;
  const size_t max_tasks = (size_t)total_buckets + total_fin / smax + 1;
  ZKP_TRY(m.task_meta.reserve(max_tasks * (sizeof(MsmTask) + 4 * sizeof(uint32_t))));
  ZKP_TRY(m.partials.reserve(max_tasks * sizeof(G1Xyzz)));
  MsmTask* tasks = m.task_meta.as<MsmTask>();
  uint32_t* task_len = reinterpret_cast<uint32_t*>(tasks + max_tasks);
  uint32_t* task_id = task_len + max_tasks;
  uint32_t* task_len_sorted = task_id + max_tasks;
  uint32_t* task_order = task_len_sorted + max_tasks;
  G1Xyzz* partials = m.partials.as<G1Xyzz>();
  // 4. tasks
  ZKP_TRY(rt::dev_memset(task_len, 0, max_tasks * 2 * sizeof(uint32_t), st));  // lengths and ids of the padding slots
  ZKP_TRY(rt::dev_memset(ntask + total_buckets, 0, 4, st));
  ZKP_TRY(rt::dev_memset(heavy, 0, 4, st));
  ZKP_LAUNCH_NOSYNC(msm_task_count_kernel, dim3((total_buckets + 255) / 256), dim3(256), 0, st, run_start, run_end,
             total_buckets, smax, ntask, heavy);
  ZKP_TRY(exclusive_scan_u32(ctx, ntask, task_off, total_buckets + 1));
  ZKP_LAUNCH_NOSYNC(msm_task_build_kernel, dim3((total_buckets + 255) / 256), dim3(256), 0, st, run_start, run_end, task_off,
             total_buckets, smax, tasks, task_len, task_id);
  ctx->msm_launches += 2;
  // 5. accumulate (XYZZ finish).  The number of tasks stays on the device (task_off[total_buckets]); the sort and the
  // launch are sized by the bound max_tasks, with zero-length padding entries sorted to the end.
  {
    const uint32_t* ntasks_dev = task_off + total_buckets;
    const uint32_t nslots = (uint32_t)max_tasks;
    uint32_t len_bits = 1;
    while ((1u << len_bits) <= smax) len_bits++;
    uint32_t *len_sorted = nullptr, *order = nullptr;
    ZKP_TRY(sort_pairs(ctx, task_len, task_id, task_len_sorted, task_order, nslots, len_bits, true, &len_sorted, &order));
    if (run_pts)
      ZKP_LAUNCH_NOSYNC(msm_accumulate_kernel<true>, dim3((nslots + ACC_THREADS - 1) / ACC_THREADS), dim3(ACC_THREADS), 0, st,
                        (const MsmTask*)tasks, (const uint32_t*)order, ntasks_dev, (const uint32_t*)nullptr, run_pts, 6u, partials);
    else
      ZKP_LAUNCH_NOSYNC(msm_accumulate_kernel<false>, dim3((nslots + ACC_THREADS - 1) / ACC_THREADS), dim3(ACC_THREADS), 0, st,
                        (const MsmTask*)tasks, (const uint32_t*)order, ntasks_dev, (const uint32_t*)svals, bases, base_stride16,
                        partials);
    ctx->msm_launches += 1;
  }
  // 6. reduce: fold the partials of heavily split buckets, gather one value per bucket, then the bit-plane levels
  phase_mark(ctx, 4);
  ZKP_TRY(rt::d2d(nfold, ntask, (size_t)total_buckets * 4, st));
  ZKP_LAUNCH(msm_heavy_fold_kernel, dim3(ctx->sm_count * 4), dim3(RED_THREADS), 0, st, (const uint32_t*)heavy,
             (const uint32_t*)task_off, (const uint32_t*)ntask, partials, nfold);
  ZKP_LAUNCH_NOSYNC(msm_bucket_gather_kernel, dim3((total_buckets + RED_THREADS - 1) / RED_THREADS), dim3(RED_THREADS), 0, st,
             (const G1Xyzz*)partials, (const uint32_t*)task_off, (const uint32_t*)nfold, total_buckets, lvl_buf);
  ctx->msm_launches += 2;
  {
    RedTree t;
    memset(&t, 0, sizeof(t));
    t.m = nbuckets;
    t.levels = levels;
    t.nsets = nsets;
    size_t off = 0;
    for (uint32_t l = 0; l <= levels; l++) {
      t.off[l] = (uint32_t)off;
      off += (size_t)(nbuckets >> l) * nsets;
    }
    off = lvl_elems * nsets;
    for (uint32_t p = 0; p + 1 < levels; p++) {
      t.poff[p] = (uint32_t)off;
      off += 2 * (size_t)((nbuckets >> (p + 2)) ? (nbuckets >> (p + 2)) : 1) * nsets;
    }
    uint32_t L = 0;
    for (; L < levels; L++) {  // wide steps: one launch each
      const size_t per_set = (size_t)(L + 1) * (nbuckets >> (L + 1));
      if (L > 0 && per_set <= TAIL_ITEMS) break;
      const size_t items = per_set * nsets;
      ZKP_LAUNCH_NOSYNC(msm_tree_step_kernel, dim3((unsigned)((items + RED_THREADS - 1) / RED_THREADS)), dim3(RED_THREADS), 0, st,
                        lvl_buf, t, L);
      ctx->msm_launches++;
    }
    // the narrow steps and the combine in one launch, one block per set
    ZKP_LAUNCH(msm_tree_tail_kernel, dim3(nsets), dim3(TAIL_THREADS), 0, st, lvl_buf, t, L, win_out);
    ctx->msm_launches++;
  }
  phase_mark(ctx, 5);
  ctx->msm_launches += ctx->sort_launches;  // pair sort, task sort, task-offset scan
  ZKP_TRY(rt::check_last());
  // 7. host: Horner over the window sums (windowed mode); in fixed-base mode set j IS the result of MSM j
  std::vector<G1Xyzz> wins(nsets);
  ZKP_TRY(rt::d2h(wins.data(), win_out, (size_t)nsets * sizeof(G1Xyzz), st));
  ZKP_TRY(rt::sync(st));
  phase_collect(ctx);
  if (fixed) {
    for (uint32_t j = 0; j < count; j++) out_host[j] = wins[j];
    return ZKP_OK;
  }
  G1Xyzz acc = wins[nsets - 1];
  for (int w = (int)nsets - 2; w >= 0; w--) {
    for (uint32_t k = 0; k < c; k++) acc = xyzz_dbl(acc);
    xyzz_add(acc, wins[w]);
  }
  out_host[0] = acc;
  return ZKP_OK;
}

// Fixed-base table for the resident SRS: window w holds 2^(c w) * srs[i].  One pass per window: c doublings
// per point in XYZZ, then batch normalisation back to affine (gen.cu).
int normalise_dev(Ctx* ctx, const G1Xyzz* tmp, size_t n, G1Affine* out);

int msm_precompute_dev(Ctx* ctx, uint32_t c) {
  rt::dev_free(ctx->srs_tab);
  ctx->srs_tab = nullptr;
  ctx->srs_tab_c = 0;
  const size_t n = ctx->srs_len;
  if (n == 0) return ZKP_OK;
  if (c == 0) c = choose_window_bits(n, true);
  if (c < 2 || c > 26) return ZKP_ERR_INVALID_ARG;
  const uint32_t nwin = 255 / c + 1;
  if ((size_t)nwin * n >= ((size_t)1 << 31)) return ZKP_ERR_INVALID_ARG;
  // Record layout: packed 96-byte records by default.  -DZKP_TABLE_PADDED switches to 128-byte records (a gathered point
  // = exactly one 128-byte DRAM line instead of 1.5 on average).  Measured on B200 at 2^24 (profiles/r02_msm_affine_ab.txt):
  // DRAM bytes of the first-round kernels drop as predicted but their time does not (73.7 vs 73.5 ms per MSM) -- the
  // gathers are bound by the number of random line fetches over an 18 GiB table, not by bytes -- so the 6 GiB are not spent.
  uint32_t stride16 = 6;
#ifdef ZKP_TABLE_PADDED
  stride16 = 8;
#ifndef ZKP_EMU
  {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && (size_t)nwin * n * 128 > total_b / 3) stride16 = 6;
  }
#endif
#endif
  ZKP_TRY(rt::dev_malloc((void**)&ctx->srs_tab, (size_t)nwin * n * stride16 * 16));
  DevBuf tmp, win;
  int st = tmp.reserve(n * sizeof(G1Xyzz));
  if (st == ZKP_OK) st = win.reserve(2 * n * sizeof(G1Affine));  // previous / current window, packed
  const unsigned sblocks = (unsigned)((n + 255) / 256);
  const G1Affine* prev = ctx->srs;
  if (st == ZKP_OK) {
    if (stride16 == 8) st = rt::dev_memset(ctx->srs_tab, 0, (size_t)nwin * n * 128, ctx->stream);  // defined padding
    ZKP_LAUNCH_NOSYNC(msm_table_store_kernel, dim3(sblocks), dim3(256), 0, ctx->stream, prev, n, ctx->srs_tab, (size_t)0, stride16);
  }
  for (uint32_t w = 1; w < nwin && st == ZKP_OK; w++) {
    G1Affine* cur = win.as<G1Affine>() + (size_t)(w & 1) * n;
    ZKP_LAUNCH_NOSYNC(msm_shift_window_kernel, dim3((unsigned)((n + 127) / 128)), dim3(128), 0, ctx->stream, prev, n, c,
               tmp.as<G1Xyzz>());
    st = normalise_dev(ctx, tmp.as<G1Xyzz>(), n, cur);
    ZKP_LAUNCH_NOSYNC(msm_table_store_kernel, dim3(sblocks), dim3(256), 0, ctx->stream, (const G1Affine*)cur, n, ctx->srs_tab,
                      (size_t)w * n, stride16);
    prev = cur;
  }
  if (st == ZKP_OK) st = rt::sync(ctx->stream);
  tmp.release();
  win.release();
  if (st != ZKP_OK) {
    rt::dev_free(ctx->srs_tab);
    ctx->srs_tab = nullptr;
    return st;
  }
  ctx->srs_tab_c = c;
  ctx->srs_tab_windows = nwin;
  ctx->srs_tab_stride16 = stride16;
  return ZKP_OK;
}

void msm_destroy(Ctx* ctx) {
#ifndef ZKP_EMU
  for (int i = 0; i <= Ctx::NPHASE; i++)
    if (ctx->phase_ev[i]) { cudaEventDestroy((cudaEvent_t)ctx->phase_ev[i]); ctx->phase_ev[i] = nullptr; }
  for (int i = 0; i < 2; i++)
    if (ctx->aff_ev[i]) { cudaEventDestroy((cudaEvent_t)ctx->aff_ev[i]); ctx->aff_ev[i] = nullptr; }
  rt::dev_free(ctx->aff_stats_dev);
  ctx->aff_stats_dev = nullptr;
#endif
  MsmScratch& m = ctx->msm;
  ctx->sort_scan.release();
  ctx->sort_hist.release();
  DevBuf* all[] = {&m.scalars, &m.bases, &m.keys_a, &m.keys_b, &m.vals_a, &m.vals_b, &m.sort_tmp, &m.bucket_start,
                   &m.bucket_end, &m.task_meta, &m.partials, &m.seg_out, &m.win_out, &m.misc, &m.aff_a, &m.aff_b, &m.aff_pre,
                   &m.aff_inv, &m.aff_off, &m.aff_tb};
  for (DevBuf* b : all) b->release();
}

}  // namespace zkp
