// Radix-2 NTT / iNTT / coset-NTT over BLS12-381 Fr, natural order in and out.
//
// Replaces ark-poly 0.4 `Radix2EvaluationDomain::{fft_in_place, ifft_in_place}` (and the coset
// variants) that the reference reaches through `Evaluations::interpolate()`
// (plonk/src/prover.rs:374-375,463; plonk/src/circuit.rs:175,230-232) and through
// `&DensePolynomial * &DensePolynomial` (plonk/src/prover.rs:396-437,516-548).
// Definition reproduced: omega_N = 7^((r-1)/N),
//   forward  X[k] = sum_j x[j] (h omega^k)^j          (h = coset offset, 1 by default)
//   inverse  x[j] = h^-j N^-1 sum_k X[k] omega^(-jk)
// Field arithmetic is exact, so any correct schedule is element-for-element identical.
//
// Schedule: log N = r1 + r2 + r3 (<= 3 passes of <= 10 levels).  The passes are consecutive groups of levels of ONE
// decimation-in-frequency transform of size N: pass i runs its r_i levels on a shared-memory tile that holds 2^r_i
// rows (stride 2^s apart) of 2^q neighbouring columns (so every global access is a run of 2^q * 32 contiguous
// bytes).  A butterfly of the level of order 2^k multiplies by omega_{2^k}^(index mod 2^(k-1)), read from the
// per-level table T_k -- the FULL twiddle, so there is no separate inter-pass product (the four-step formulation
// spends one or two extra products per element and pass on omega_N^(col * digit)): a transform costs
// (log N - 1) * N / 2 products and nothing else.  Only the first pass reads tables larger than L2 (T_logN ... :
// N * 32 bytes in all, streamed once); later passes hit L1 / L2.  The last pass writes transposed, which puts the
// digit-reversed result in natural order without a separate permutation sweep.  Pass 1 reads `data` and writes
// `scratch`, the last pass writes back to `data`: 64 * N bytes of HBM traffic per pass.
#include "engine.h"
#include "memops.cuh"
#include "poly.h"

namespace zkp {

#ifndef NTT_WLOG
#define NTT_WLOG 11
#endif
#ifndef NTT_RMAX
#define NTT_RMAX 10
#endif
static constexpr uint32_t WLOG = NTT_WLOG;  // the per-level twiddle tables always cover orders up to 2^11 (one tile)
static constexpr uint32_t RMAX = NTT_RMAX;  // largest digit of a multi-pass schedule
// Tile / CTA shape (measured on B200, profiles/r02_ntt_variants.txt): radix-4 register blocks (96 registers) with
// 32 KB tiles and 128-thread CTAs keep 5 CTAs = 20 warps and five independent barrier domains per SM; the radix-8 /
// 64 KB / 256-thread shape of round 1 (128 registers, 2 CTAs) was 15 % slower on the same schedule.
#ifndef NTT_TILE_LOG
#define NTT_TILE_LOG 10
#endif
#ifndef NTT_THREADS_PER_CTA
#define NTT_THREADS_PER_CTA 128
#endif
static constexpr uint32_t TILE_LOG = NTT_TILE_LOG;   // elements per shared-memory tile (2^10 = 32 KB)
static constexpr uint32_t SINGLE_MAX = NTT_TILE_LOG; // largest transform done in one tile
static constexpr uint32_t NTT_THREADS = NTT_THREADS_PER_CTA;
#ifndef NTT_MIN_BLOCKS
#define NTT_MIN_BLOCKS 5
#endif
#ifndef NTT_KMAX
#define NTT_KMAX 2  // levels per register block (2: radix-4, 3: radix-8)
#endif
#ifndef NTT_TW_PRELOAD
#define NTT_TW_PRELOAD 1
#endif
#ifndef NTT_TW_PREFETCH2
#define NTT_TW_PREFETCH2 0
#endif
#ifndef NTT_LOAD_BATCH
#define NTT_LOAD_BATCH 4  // tile elements a thread requests before it stores the first one (1 = the plain loop)
#endif

struct NttPassArgs {
  const Fr* in;
  Fr* out;
  uint32_t log_n, r, q, s;
  uint32_t mode;      // 0: strided in-place digit, 1: last digit (transposing write)
  uint32_t log_n1, log_mid;
  const Fr* w;        // per-level butterfly twiddles: T_k at [2^(k-1), 2^k)
  uint32_t full_tw;   // 1: butterflies use the full order-2^(s + level) twiddle (single-GPU passes: no inter-pass product)
  const Fr* tw_lo;    // two-level inter-stage twiddles (distributed stage only)
  const Fr* tw_hi;
  uint32_t tw_lb, tw_shift;
  const Fr* pre_lo;   // load-time scale by table[addr] (coset forward), or null
  const Fr* pre_hi;
  uint32_t pre_lb;
  const Fr* post_lo;  // store-time scale by table[addr] (coset inverse), or null
  const Fr* post_hi;
  uint32_t post_lb;
  const Fr* scale;    // store-time constant (N^-1 on the single-pass inverse), or null
  size_t batch_stride;
};

// Distributed four-step transform: the first stage of the forward transform (last stage of
// the inverse) is this same pass kernel run on the rank's column block of the 2^r x 2^s_glob matrix.
// The twiddle / coset exponents use GLOBAL column indices; rows can be pushed to (forward) or pulled from
// (inverse) the exchange buffers of the peer GPUs over NVLink, which fuses the all-to-all transpose into
// the kernel's store / load loop.  Exchange-buffer layout on every rank: [row'][global column].
struct NttDistArgs {
  uint32_t s_glob;      // log_n - r
  uint32_t col_off;     // global index of local column 0
  uint32_t tw_on_load;  // inverse stage: multiply by omega^-(col * k) while loading row k
  uint32_t peer_log;    // rows owned by one peer = 2^peer_log
  uint32_t push, pull;
  Fr* peer[8];
};

// Shared-memory tile: two planes of 16-byte halves so that consecutive lanes touch consecutive
// 16-byte words (conflict-free for unit-stride element access).  The slot of element i = (row << q) + c is
// i ^ ((i >> sw) & 7), sw = max(q, 3): the low three column bits are XORed with the low row bits, so the eight
// lanes of a quarter-warp hit eight different 16-byte bank groups both when they walk along a row (the butterfly
// steps, the coalesced loads / stores of the strided passes) and when they walk down a column (the transposing
// load of the last pass, which used to be an 8-way conflict).
__device__ __forceinline__ uint32_t tile_slot(uint32_t i, uint32_t sw) { return i ^ ((i >> sw) & 7u); }
__device__ __forceinline__ Fr ld_tile(const uint4* lo, const uint4* hi, uint32_t i, uint32_t sw) {
  i = tile_slot(i, sw);
  uint4 a = lo[i], b = hi[i];
  Fr r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ void st_tile(uint4* lo, uint4* hi, uint32_t i, uint32_t sw, const Fr& r) {
  i = tile_slot(i, sw);
  lo[i] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
  hi[i] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// K decimation-in-frequency levels (l0 .. l0+K-1 of an r-level transform along the tile rows) held
// in registers: each work item owns the 2^K rows that differ in bits [lo_bit, lo_bit + K).
// Twiddles: the level of local order 2^order_log belongs to the global level of order 2^(order_log + s_eff); element
// (row j of the butterfly group, column col) reads T[(j << s_eff) + col] of that order (s_eff = 0, col = 0: plain
// size-2^r transform along the rows).
// The 2^K - 1 twiddles of one work item, level by level (level t contributes 2^(K-1-t) of them).
template <int K>
struct TileTw {
  Fr t[(1 << K) - 1];
};

struct TileGeom {
  uint32_t r, q, sw, s_eff, colbase;
  const Fr* __restrict__ w;
};

template <int K>
__device__ __forceinline__ void tile_tw_load(TileTw<K>& tw, const TileGeom& g, uint32_t l0, uint32_t it) {
  const uint32_t lo_bit = g.r - l0 - K;
  const uint32_t c = it & ((1u << g.q) - 1);
  const uint32_t dlow = (it >> g.q) & ((1u << lo_bit) - 1);
  const uint32_t col = g.s_eff ? g.colbase + c : 0u;
  int idx = 0;
#pragma unroll
  for (int t = 0; t < K; t++) {
    const int hm = 1 << (K - 1 - t);
    const uint32_t order_log = lo_bit + K - t;  // butterflies of this level use omega_{2^(order_log + s_eff)}
    const Fr* __restrict__ tk = g.w + ((size_t)1 << (order_log + g.s_eff - 1));
#pragma unroll
    for (int mm = 0; mm < hm; mm++) {
      const uint32_t j = dlow + ((uint32_t)mm << lo_bit);
      // omega_2^0 = 1 on the very last level (slot 1 of the table array): never multiplied, see tile_block
      tw.t[idx++] = ldg_fr(tk + (((size_t)j << g.s_eff) + col));
    }
  }
}

// Same addresses as tile_tw_load, requested into L1 without binding registers (the thread's SECOND work item of a step:
// its twiddles would otherwise be fetched on demand after the first item's butterflies).
template <int K>
__device__ __forceinline__ void tile_tw_prefetch(const TileGeom& g, uint32_t l0, uint32_t it) {
#ifndef ZKP_EMU
  const uint32_t lo_bit = g.r - l0 - K;
  const uint32_t c = it & ((1u << g.q) - 1);
  const uint32_t dlow = (it >> g.q) & ((1u << lo_bit) - 1);
  const uint32_t col = g.s_eff ? g.colbase + c : 0u;
#pragma unroll
  for (int t = 0; t < K; t++) {
    const int hm = 1 << (K - 1 - t);
    const uint32_t order_log = lo_bit + K - t;
    const Fr* __restrict__ tk = g.w + ((size_t)1 << (order_log + g.s_eff - 1));
#pragma unroll
    for (int mm = 0; mm < hm; mm++) {
      const uint32_t j = dlow + ((uint32_t)mm << lo_bit);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(tk + (((size_t)j << g.s_eff) + col)));
    }
  }
#else
  (void)g; (void)l0; (void)it;
#endif
}

template <int K>
__device__ __forceinline__ void tile_block(uint4* lo, uint4* hi, const TileGeom& g, uint32_t l0, uint32_t it,
                                           const TileTw<K>& tw) {
  const uint32_t lo_bit = g.r - l0 - K;
  const uint32_t q = g.q, sw = g.sw;
  const uint32_t c = it & ((1u << q) - 1);
  const uint32_t jr = it >> q;
  const uint32_t dlow = jr & ((1u << lo_bit) - 1);
  const uint32_t base_d = ((jr >> lo_bit) << (lo_bit + K)) | dlow;
  Fr e[1 << K];
#pragma unroll
  for (int m = 0; m < (1 << K); m++) e[m] = ld_tile(lo, hi, ((base_d + ((uint32_t)m << lo_bit)) << q) + c, sw);
  int off = 0;
#pragma unroll
  for (int t = 0; t < K; t++) {
    const int hm = 1 << (K - 1 - t);
    const uint32_t order_log = lo_bit + K - t;
#pragma unroll
    for (int m = 0; m < (1 << K); m++) {
      if (m & hm) continue;
      Fr u = e[m], v = e[m + hm];
      e[m] = fp_add(u, v);
      Fr d = fp_sub(u, v);
      if (order_log + g.s_eff == 1) {
        e[m + hm] = d;  // omega_2^0 = 1 on the very last level
      } else {
        e[m + hm] = fp_mul(d, tw.t[off + (m & (hm - 1))]);
      }
    }
    off += hm;
  }
#pragma unroll
  for (int m = 0; m < (1 << K); m++) st_tile(lo, hi, ((base_d + ((uint32_t)m << lo_bit)) << q) + c, sw, e[m]);
}

// One step = barrier + K levels.  The twiddles of the thread's first work item are requested BEFORE the barrier
// (they do not depend on the tile), so their L2 / HBM latency overlaps the wait for the other warps.
template <int K>
__device__ __forceinline__ void tile_step(uint4* lo, uint4* hi, const TileGeom& g, uint32_t l0, uint32_t tid,
                                          uint32_t nthreads) {
  const uint32_t items = 1u << (g.r - K + g.q);
  TileTw<K> tw;
#if NTT_TW_PRELOAD
  if (tid < items) tile_tw_load<K>(tw, g, l0, tid);
#if NTT_TW_PREFETCH2
  if (tid + nthreads < items) tile_tw_prefetch<K>(g, l0, tid + nthreads);
#endif
  __syncthreads();
  for (uint32_t it = tid; it < items; it += nthreads) {
    if (it != tid) tile_tw_load<K>(tw, g, l0, it);
    tile_block<K>(lo, hi, g, l0, it, tw);
  }
#else
  __syncthreads();
  for (uint32_t it = tid; it < items; it += nthreads) {
    tile_tw_load<K>(tw, g, l0, it);
    tile_block<K>(lo, hi, g, l0, it, tw);
  }
#endif
}

template <bool DIST>
__device__ __forceinline__ void ntt_pass_body(const NttPassArgs& a, const NttDistArgs* dd) {
  ZKP_DYN_SMEM(uint4, sm);
  const uint32_t r = a.r, q = a.q, s = a.s;
  const uint32_t E = 1u << (r + q);
  uint4* lo = sm;
  uint4* hi = sm + E;
  const uint32_t tid = threadIdx.x, nthreads = blockDim.x;
  const uint32_t tile = blockIdx.x;
  const Fr* in = a.in + (size_t)blockIdx.y * a.batch_stride;
  Fr* out = a.out + (size_t)blockIdx.y * a.batch_stride;
  const uint32_t qmask = (1u << q) - 1, rmask = (1u << r) - 1;
  const uint32_t sw = q > 3 ? q : 3;

  // ---- tile geometry ----
  uint32_t base = 0, cb = 0, mid = 0, ab = 0;
  if (a.mode == 0) {
    const uint32_t ncb_log = s - q;
    const uint32_t outer = tile >> ncb_log;
    cb = tile & ((1u << ncb_log) - 1);
    base = (outer << (s + r)) + (cb << q);
  } else {
    const uint32_t nab_log = a.log_n1 - q;
    mid = tile >> nab_log;
    ab = tile & ((1u << nab_log) - 1);
  }

  // ---- load (optionally scaled by the coset powers h^addr) ----
  // Single-GPU passes: a thread's elements are requested NTT_LOAD_BATCH at a time before the first one is stored (the plain
  // loop compiles to load -> wait -> store per element).  Measured on B200 (profiles/r02_ntt_variants.txt): 3.69 / 3.61 /
  // 3.66 ms at 2^24 for batches of 1 / 4 / 8 -- the five resident CTAs already hide the load phase, so this is worth 1-2 %.
  if constexpr (!DIST) {
    // global address / tile slot of tile element idx
    auto locate = [&](uint32_t idx, uint32_t& addr, uint32_t& sidx) {
      if (a.mode == 0) {
        const uint32_t c = idx & qmask, d = idx >> q;
        addr = base + (d << s) + c;
        sidx = idx;
      } else {
        const uint32_t d = idx & rmask, c = idx >> r;
        const uint32_t k1 = (ab << q) + c;
        addr = (((k1 << a.log_mid) + mid) << r) + d;
        sidx = (d << q) + c;
      }
    };
    for (uint32_t idx0 = tid; idx0 < E; idx0 += nthreads * NTT_LOAD_BATCH) {
      Fr v[NTT_LOAD_BATCH];
#pragma unroll
      for (uint32_t u = 0; u < NTT_LOAD_BATCH; u++) {
        const uint32_t idx = idx0 + u * nthreads;
        if (idx < E) {
          uint32_t addr, sidx;
          locate(idx, addr, sidx);
          v[u] = ld_fr(in + addr);
        }
      }
#pragma unroll
      for (uint32_t u = 0; u < NTT_LOAD_BATCH; u++) {
        const uint32_t idx = idx0 + u * nthreads;
        if (idx < E) {
          uint32_t addr, sidx;
          locate(idx, addr, sidx);
          if (a.pre_lo) {
            v[u] = fp_mul(v[u], ldg_fr(a.pre_lo + (addr & ((1u << a.pre_lb) - 1))));
            v[u] = fp_mul(v[u], ldg_fr(a.pre_hi + (addr >> a.pre_lb)));
          }
          st_tile(lo, hi, sidx, sw, v[u]);
        }
      }
    }
  } else {
  for (uint32_t idx = tid; idx < E; idx += nthreads) {
    uint32_t addr, sidx;
    const Fr* src = in;
    uint32_t tw_k = 0, tw_col = 0;
    if (a.mode == 0) {
      const uint32_t c = idx & qmask, d = idx >> q;
      addr = base + (d << s) + c;
      sidx = idx;
      if constexpr (DIST) {
        const uint32_t col = (cb << q) + c;
        tw_k = d;
        tw_col = dd->col_off + col;
        const uint32_t gaddr = (d << dd->s_glob) + tw_col;
        if (dd->pull) {
          src = dd->peer[d >> dd->peer_log];
          addr = ((d & ((1u << dd->peer_log) - 1)) << dd->s_glob) + tw_col;
        }
        Fr v = ld_fr(src + addr);
        if (dd->tw_on_load) {
          const uint64_t e = ((uint64_t)tw_col * tw_k) << a.tw_shift;
          v = fp_mul(v, ldg_fr(a.tw_hi + (size_t)(e >> a.tw_lb)));
          v = fp_mul(v, ldg_fr(a.tw_lo + (size_t)(e & ((1u << a.tw_lb) - 1))));
        }
        if (a.pre_lo) {
          v = fp_mul(v, ldg_fr(a.pre_lo + (gaddr & ((1u << a.pre_lb) - 1))));
          v = fp_mul(v, ldg_fr(a.pre_hi + (gaddr >> a.pre_lb)));
        }
        st_tile(lo, hi, sidx, sw, v);
        continue;
      }
    } else {
      const uint32_t d = idx & rmask, c = idx >> r;
      const uint32_t k1 = (ab << q) + c;
      addr = (((k1 << a.log_mid) + mid) << r) + d;
      sidx = (d << q) + c;
    }
    Fr v = ld_fr(in + addr);
    if (a.pre_lo) {
      v = fp_mul(v, ldg_fr(a.pre_lo + (addr & ((1u << a.pre_lb) - 1))));
      v = fp_mul(v, ldg_fr(a.pre_hi + (addr >> a.pre_lb)));
    }
    st_tile(lo, hi, sidx, sw, v);
  }

  }  // DIST
  // ---- r DIF levels on the tile rows (result row d holds output digit bitrev_r(d)) ----
  TileGeom g;
  g.r = r; g.q = q; g.sw = sw;
  g.s_eff = (!DIST && a.mode == 0 && a.full_tw) ? s : 0u;
  g.colbase = cb << q;
  g.w = a.w;
  uint32_t l0 = 0;
#if NTT_KMAX == 1  // experiment knob: radix-2 steps only
  while (l0 < r) {
    tile_step<1>(lo, hi, g, l0, tid, nthreads);
    l0 += 1;
  }
#else
  while (r - l0 >= 3) {
#if NTT_KMAX >= 3
    tile_step<3>(lo, hi, g, l0, tid, nthreads);
    l0 += 3;
#else  // radix-4 register blocks only (fewer live registers -> more resident warps, one more shared-memory round trip)
    tile_step<2>(lo, hi, g, l0, tid, nthreads);
    l0 += 2;
#endif
  }
  if (r - l0 == 2) {
    tile_step<2>(lo, hi, g, l0, tid, nthreads);
  } else if (r - l0 == 1) {
    tile_step<1>(lo, hi, g, l0, tid, nthreads);
  }
#endif
  __syncthreads();

  // ---- store: un-bit-reverse the digit, apply inter-pass twiddle / coset / N^-1 ----
  for (uint32_t idx = tid; idx < E; idx += nthreads) {
    const uint32_t c = idx & qmask, k = idx >> q;
    const uint32_t d = r ? (__brev(k) >> (32 - r)) : 0;
    Fr v = ld_tile(lo, hi, (d << q) + c, sw);
    uint32_t addr;
    if (a.mode == 0) {
      const uint32_t col = (cb << q) + c;
      if constexpr (DIST) {
        const uint32_t colg = dd->col_off + col;
        if (!dd->tw_on_load) {
          const uint64_t e = ((uint64_t)colg * k) << a.tw_shift;
          v = fp_mul(v, ldg_fr(a.tw_hi + (size_t)(e >> a.tw_lb)));
          v = fp_mul(v, ldg_fr(a.tw_lo + (size_t)(e & ((1u << a.tw_lb) - 1))));
        }
        if (a.post_lo) {
          const uint32_t gaddr = (k << dd->s_glob) + colg;
          v = fp_mul(v, ldg_fr(a.post_lo + (gaddr & ((1u << a.post_lb) - 1))));
          v = fp_mul(v, ldg_fr(a.post_hi + (gaddr >> a.post_lb)));
        }
        if (dd->push) {
          Fr* dst = dd->peer[k >> dd->peer_log];
          st_fr(dst + (((k & ((1u << dd->peer_log) - 1)) << dd->s_glob) + colg), v);
        } else {
          st_fr(out + (base + (k << s) + c), v);
        }
        continue;
      }
      if (!a.full_tw) {  // four-step formulation (kept for A/B measurements): omega_M^(col * k), two-level table
        const uint64_t e = ((uint64_t)col * k) << a.tw_shift;
        v = fp_mul(v, ldg_fr(a.tw_hi + (size_t)(e >> a.tw_lb)));
        v = fp_mul(v, ldg_fr(a.tw_lo + (size_t)(e & ((1u << a.tw_lb) - 1))));
      }
      addr = base + (k << s) + c;
    } else {
      const uint32_t k1 = (ab << q) + c;
      addr = k1 + (mid << a.log_n1) + (k << (a.log_n - r));
      if (a.post_lo) {
        v = fp_mul(v, ldg_fr(a.post_lo + (addr & ((1u << a.post_lb) - 1))));
        v = fp_mul(v, ldg_fr(a.post_hi + (addr >> a.post_lb)));
      }
      if (a.scale) v = fp_mul(v, ldg_fr(a.scale));
    }
    st_fr(out + addr, v);
  }
}

__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) ntt_pass_kernel(NttPassArgs a) {
  ntt_pass_body<false>(a, nullptr);
}

// Same pass with global column indexing and (optionally) peer loads / stores: the distributed stage.
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) ntt_dist_pass_kernel(NttPassArgs a, NttDistArgs d) {
  ntt_pass_body<true>(a, &d);
}

// All-to-all companion for the NCCL path: forward  recv[g][row'][col'] -> out[row'][g][col'],
// inverse  in[row'][g][col'] -> send[g][row'][col']   (row' < 2^rows_log, g < 2^wl, col' < 2^cols_log).
__global__ void __launch_bounds__(256) ntt_dist_permute_kernel(const Fr* in, Fr* out, uint32_t rows_log, uint32_t wl,
                                                               uint32_t cols_log, uint32_t inverse) {
  const size_t total = (size_t)1 << (rows_log + wl + cols_log);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const size_t c = i & (((size_t)1 << cols_log) - 1);
    const size_t g = (i >> cols_log) & (((size_t)1 << wl) - 1);
    const size_t rw = i >> (cols_log + wl);
    const size_t j = (((g << rows_log) + rw) << cols_log) + c;  // index in the [g][row'][col'] layout
    if (inverse) st_fr(out + j, ld_fr(in + i));
    else st_fr(out + i, ld_fr(in + j));
  }
}

__global__ void __launch_bounds__(256) fr_pointwise_mul_kernel(Fr* a, const Fr* b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) st_fr(a + i, fp_mul(ld_fr(a + i), ld_fr(b + i)));
}

// ------------------------------------------------------------------------------------------------
// Host side: schedules and twiddle tables
// ------------------------------------------------------------------------------------------------
static Fr fr_root_of_unity_2_32() {
  // 7^((r-1)/2^32): the exponent is r-1 shifted down one 32-bit limb (low limb of r-1 is zero)
  Fr seven = Fr::zero();
  seven.v[0] = 7;
  seven = fp_to_mont(seven);
  Fr acc = Fr::one();
  for (int i = 7 * 32 - 1; i >= 0; i--) {
    acc = fp_sqr(acc);
    const uint32_t limb = FrParams::mod(1 + (i >> 5));
    if ((limb >> (i & 31)) & 1) acc = fp_mul(acc, seven);
  }
  return acc;
}

static Fr fr_omega(uint32_t log_n) {  // group_gen of the size-2^log_n domain
  Fr w = fr_root_of_unity_2_32();
  for (uint32_t i = log_n; i < 32; i++) w = fp_sqr(w);
  return w;
}

static Fr fr_from_u64(uint64_t x) {
  Fr a = Fr::zero();
  a.v[0] = (uint32_t)x;
  a.v[1] = (uint32_t)(x >> 32);
  return fp_to_mont(a);
}

static int upload_powers(Ctx* ctx, Fr** dev, const Fr& base, const Fr& first, size_t count) {
  std::vector<Fr> h(count);
  Fr cur = first;
  for (size_t i = 0; i < count; i++) {
    h[i] = cur;
    cur = fp_mul(cur, base);
  }
  ZKP_TRY(rt::dev_malloc((void**)dev, count * sizeof(Fr)));
  ZKP_TRY(rt::h2d(*dev, h.data(), count * sizeof(Fr), ctx->stream));
  return rt::sync(ctx->stream);
}

// Per-level twiddle tables T_1 .. T_log (both directions), built on the device; grown when a larger transform
// shows up (the old arrays are released after the stream has drained).
static int ensure_twiddles(Ctx* ctx, uint32_t log_n) {
  const uint32_t need = log_n > WLOG ? log_n : WLOG;
  if (ctx->tw_log >= need) return ZKP_OK;
  Fr *f = nullptr, *b = nullptr;
  const size_t count = (size_t)1 << need;
  ZKP_TRY(rt::dev_malloc((void**)&f, count * sizeof(Fr)));
  int st = rt::dev_malloc((void**)&b, count * sizeof(Fr));
  const Fr one = Fr::one();
  if (st == ZKP_OK) st = rt::h2d(f, &one, sizeof(Fr), ctx->stream);  // slot 0 is unused
  if (st == ZKP_OK) st = rt::h2d(b, &one, sizeof(Fr), ctx->stream);
  for (uint32_t k = 1; k <= need && st == ZKP_OK; k++) {
    const Fr w = fr_omega(k);
    const size_t half = (size_t)1 << (k - 1);
    st = fr_powers_dev(ctx, f + half, w, one, half);
    if (st == ZKP_OK) st = fr_powers_dev(ctx, b + half, fp_inv(w), one, half);
  }
  if (st == ZKP_OK) st = rt::sync(ctx->stream);
  if (st != ZKP_OK) {
    rt::dev_free(f);
    rt::dev_free(b);
    return st;
  }
  rt::dev_free(ctx->tw_fwd);
  rt::dev_free(ctx->tw_inv);
  ctx->tw_fwd = f;
  ctx->tw_inv = b;
  ctx->tw_log = need;
  return ZKP_OK;
}

int ntt_init(Ctx* ctx) {
  ZKP_TRY(ensure_twiddles(ctx, WLOG));
  ZKP_TRY(rt::allow_smem((const void*)ntt_pass_kernel, ((size_t)1 << TILE_LOG) * sizeof(Fr)));
  ZKP_TRY(rt::allow_smem((const void*)ntt_dist_pass_kernel, ((size_t)1 << TILE_LOG) * sizeof(Fr)));
  ZKP_TRY(rt::prefer_smem_carveout((const void*)ntt_dist_pass_kernel));
  return rt::prefer_smem_carveout((const void*)ntt_pass_kernel);  // several tiles resident per SM
}

void ntt_destroy(Ctx* ctx) {
  rt::dev_free(ctx->tw_fwd);
  rt::dev_free(ctx->tw_inv);
  ctx->tw_fwd = ctx->tw_inv = nullptr;
  ctx->tw_log = 0;
  for (auto& kv : ctx->ntt_tables) {
    rt::dev_free(kv.second.tw_lo);
    rt::dev_free(kv.second.tw_hi);
    rt::dev_free(kv.second.scale);
  }
  for (auto& c : ctx->coset_tables) {
    rt::dev_free(c.lo);
    rt::dev_free(c.hi);
  }
  ctx->ntt_tables.clear();
  ctx->coset_tables.clear();
  ctx->ntt_scratch.release();
  ctx->ntt_io.release();
  ctx->ntt_io2.release();
  ctx->poly_scratch.release();
  ctx->eval_out.release();
  ctx->eval_partials.release();
}

static int get_tables(Ctx* ctx, uint32_t log_n, bool inverse, NttTables** out) {
  const uint32_t key = log_n * 2 + (inverse ? 1 : 0);
  auto it = ctx->ntt_tables.find(key);
  if (it != ctx->ntt_tables.end()) {
    *out = &it->second;
    return ZKP_OK;
  }
  NttTables t;
  t.log_n = log_n;
  t.inverse = inverse;
  if (log_n <= SINGLE_MAX) {
    t.npass = 1;
    t.digits[0] = log_n;
  } else {
    t.npass = (log_n + RMAX - 1) / RMAX;
    if (t.npass > 3) return ZKP_ERR_DOMAIN_TOO_LARGE;
    const uint32_t b = log_n / t.npass, rem = log_n % t.npass;
    for (uint32_t i = 0; i < t.npass; i++) t.digits[i] = b + (i < rem ? 1 : 0);
  }
  if (inverse) {
    const Fr n_inv = fp_inv(fr_from_u64((uint64_t)1 << log_n));
    ZKP_TRY(upload_powers(ctx, &t.scale, Fr::one(), n_inv, 1));
  }
#ifdef NTT_FOUR_STEP_TWIDDLES  // A/B knob: the four-step formulation with its inter-pass products
  if (t.npass > 1) {
    t.lb = t.digits[0];
    Fr w = fr_omega(log_n);
    if (inverse) w = fp_inv(w);
    Fr whi = w;
    for (uint32_t i = 0; i < t.lb; i++) whi = fp_sqr(whi);
    ZKP_TRY(upload_powers(ctx, &t.tw_lo, w, Fr::one(), (size_t)1 << t.lb));
    ZKP_TRY(upload_powers(ctx, &t.tw_hi, whi, Fr::one(), (size_t)1 << (log_n - t.lb)));
  }
#endif
  auto ins = ctx->ntt_tables.emplace(key, t);
  *out = &ins.first->second;
  return ZKP_OK;
}

// `scaled`: fold N^-1 into the low table (inverse transforms of the single-GPU path)
static int get_coset(Ctx* ctx, uint32_t log_n, bool inverse, const Fr& h, bool scaled, CosetTables** out) {
  for (auto& c : ctx->coset_tables)
    if (c.log_n == log_n && c.inverse == inverse && c.scaled == scaled && c.offset == h) {
      *out = &c;
      return ZKP_OK;
    }
  if (ctx->coset_tables.size() >= 8) {  // small LRU-less cache: drop the oldest entry
    rt::dev_free(ctx->coset_tables.front().lo);
    rt::dev_free(ctx->coset_tables.front().hi);
    ctx->coset_tables.erase(ctx->coset_tables.begin());
  }
  CosetTables c;
  c.log_n = log_n;
  c.inverse = inverse;
  c.offset = h;
  c.scaled = scaled;
  c.lb = log_n / 2;
  Fr g = inverse ? fp_inv(h) : h;
  Fr ghi = g;
  for (uint32_t i = 0; i < c.lb; i++) ghi = fp_sqr(ghi);
  const Fr first = scaled ? fp_inv(fr_from_u64((uint64_t)1 << log_n)) : Fr::one();
  ZKP_TRY(upload_powers(ctx, &c.lo, g, first, (size_t)1 << c.lb));
  ZKP_TRY(upload_powers(ctx, &c.hi, ghi, Fr::one(), (size_t)1 << (log_n - c.lb)));
  ctx->coset_tables.push_back(c);
  *out = &ctx->coset_tables.back();
  return ZKP_OK;
}

int ntt_run_dev(Ctx* ctx, Fr* data, uint32_t log_n, size_t batch, bool inverse, const Fr* coset_host) {
  ctx->ntt_launches = 0;
  if (batch == 0) return ZKP_OK;
  if (log_n > 27) return ZKP_ERR_DOMAIN_TOO_LARGE;
  if (batch > 65535) return ZKP_ERR_INVALID_ARG;
  if (log_n == 0) return ZKP_OK;  // size-1 domain: identity in both directions (h^0 = 1, 1^-1 = 1)
  NttTables* t = nullptr;
  ZKP_TRY(get_tables(ctx, log_n, inverse, &t));
  ZKP_TRY(ensure_twiddles(ctx, log_n));
  CosetTables* cs = nullptr;
  if (coset_host) {
    const Fr one = Fr::one();
    if (!(*coset_host == one)) ZKP_TRY(get_coset(ctx, log_n, inverse, *coset_host, inverse, &cs));
  }
  const size_t N = (size_t)1 << log_n;
  Fr* scratch = nullptr;
  if (t->npass > 1) {
    ZKP_TRY(ctx->ntt_scratch.reserve(N * batch * sizeof(Fr)));
    scratch = ctx->ntt_scratch.as<Fr>();
  }
  uint32_t below = log_n;  // bits below the current digit + the digit itself
  for (uint32_t i = 0; i < t->npass; i++) {
    const uint32_t r = t->digits[i];
    below -= r;
    const bool last = (i + 1 == t->npass);
    NttPassArgs a;
    memset(&a, 0, sizeof(a));
    a.in = (i == 0) ? data : scratch;
    a.out = last ? data : scratch;
    a.log_n = log_n;
    a.r = r;
    a.s = below;
    a.w = inverse ? ctx->tw_inv : ctx->tw_fwd;
    a.batch_stride = N;
    a.full_tw = 1;
    if (!last) {
      a.mode = 0;
      a.q = (TILE_LOG - r < a.s) ? (TILE_LOG - r) : a.s;
#ifdef NTT_FOUR_STEP_TWIDDLES
      a.full_tw = 0;
      a.tw_lo = t->tw_lo;
      a.tw_hi = t->tw_hi;
      a.tw_lb = t->lb;
      a.tw_shift = log_n - (a.s + r);
#endif
    } else {
      a.mode = 1;
      if (t->npass == 1) {
        a.log_n1 = 0;
        a.log_mid = 0;
        a.q = 0;
      } else {
        a.log_n1 = t->digits[0];
        a.log_mid = (t->npass == 3) ? t->digits[1] : 0;
        a.q = (TILE_LOG - r < a.log_n1) ? (TILE_LOG - r) : a.log_n1;
      }
    }
    if (i == 0 && cs && !inverse) { a.pre_lo = cs->lo; a.pre_hi = cs->hi; a.pre_lb = cs->lb; }
    if (last && cs && inverse) { a.post_lo = cs->lo; a.post_hi = cs->hi; a.post_lb = cs->lb; }  // carries N^-1
    if (last && inverse && !cs) a.scale = t->scale;
    const uint32_t E = 1u << (r + a.q);
    const uint32_t tiles = (uint32_t)(N >> (r + a.q));
    uint32_t threads = NTT_THREADS;
    while (threads > 32 && threads > E / 2) threads >>= 1;
    dim3 grid(tiles, (unsigned)batch, 1);
    ZKP_LAUNCH(ntt_pass_kernel, grid, dim3(threads), (size_t)E * sizeof(Fr), ctx->stream, a);
    ctx->ntt_launches++;
  }
  return rt::check_last();
}

// ------------------------------------------------------------------------------------------------
// Distributed four-step transform (one process per GPU; the exchange itself is done by the caller with
// NCCL, or fused into the stage kernel through peer pointers)
// ------------------------------------------------------------------------------------------------
uint32_t ntt_dist_rows_log(uint32_t log_n, uint32_t world_log) {
  // N = 2^r rows x 2^(log_n - r) columns; every rank needs >= 1 row block and >= 1 column block
  if (log_n < 2 * world_log || log_n < 2) return 0;
  uint32_t r = log_n / 2;
  if (r > RMAX) r = RMAX;
  if (r < world_log) r = world_log;
  if (log_n - r < world_log) return 0;
  return r;
}

static int get_dist_tables(Ctx* ctx, uint32_t log_n, uint32_t r, bool inverse, NttTables** out) {
  const uint32_t key = 0x10000u + log_n * 64 + r * 2 + (inverse ? 1 : 0);
  auto it = ctx->ntt_tables.find(key);
  if (it != ctx->ntt_tables.end()) {
    *out = &it->second;
    return ZKP_OK;
  }
  NttTables t;
  t.log_n = log_n;
  t.inverse = inverse;
  t.npass = 1;
  t.digits[0] = r;
  t.lb = log_n / 2;
  Fr w = fr_omega(log_n);
  if (inverse) w = fp_inv(w);
  Fr whi = w;
  for (uint32_t i = 0; i < t.lb; i++) whi = fp_sqr(whi);
  // inverse: the stage contributes (2^r)^-1; the local transforms of size 2^(log_n - r) carry their own
  const Fr first = inverse ? fp_inv(fr_from_u64((uint64_t)1 << r)) : Fr::one();
  ZKP_TRY(upload_powers(ctx, &t.tw_lo, w, first, (size_t)1 << t.lb));
  ZKP_TRY(upload_powers(ctx, &t.tw_hi, whi, Fr::one(), (size_t)1 << (log_n - t.lb)));
  auto ins = ctx->ntt_tables.emplace(key, t);
  *out = &ins.first->second;
  return ZKP_OK;
}

int ntt_dist_stage_dev(Ctx* ctx, Fr* data, uint32_t log_n, uint32_t rank, uint32_t world_log, bool inverse,
                       const Fr* coset_host, Fr* const* peers) {
  ctx->ntt_launches = 0;
  const uint32_t r = ntt_dist_rows_log(log_n, world_log);
  if (r == 0 || log_n > 31 || world_log > 3 || rank >= (1u << world_log)) return ZKP_ERR_INVALID_ARG;
  NttTables* t = nullptr;
  ZKP_TRY(get_dist_tables(ctx, log_n, r, inverse, &t));
  ZKP_TRY(ensure_twiddles(ctx, r));
  CosetTables* cs = nullptr;
  if (coset_host) {
    const Fr one = Fr::one();
    if (!(*coset_host == one)) ZKP_TRY(get_coset(ctx, log_n, inverse, *coset_host, false, &cs));
  }
  const uint32_t s_glob = log_n - r, s_loc = s_glob - world_log;
  NttPassArgs a;
  memset(&a, 0, sizeof(a));
  a.in = data;
  a.out = data;
  a.log_n = log_n;
  a.r = r;
  a.s = s_loc;
  a.q = (TILE_LOG - r < s_loc) ? (TILE_LOG - r) : s_loc;
  a.mode = 0;
  a.w = inverse ? ctx->tw_inv : ctx->tw_fwd;
  a.full_tw = 0;  // the stage keeps the four-step twiddle omega_N^(col * k): its output crosses GPUs in that form
  a.tw_lo = t->tw_lo;
  a.tw_hi = t->tw_hi;
  a.tw_lb = t->lb;
  a.tw_shift = 0;
  a.batch_stride = 0;
  if (cs && !inverse) { a.pre_lo = cs->lo; a.pre_hi = cs->hi; a.pre_lb = cs->lb; }
  if (cs && inverse) { a.post_lo = cs->lo; a.post_hi = cs->hi; a.post_lb = cs->lb; }
  NttDistArgs d;
  memset(&d, 0, sizeof(d));
  d.s_glob = s_glob;
  d.col_off = rank << s_loc;
  d.tw_on_load = inverse ? 1 : 0;
  d.peer_log = r - world_log;
  if (peers) {
    for (uint32_t g = 0; g < (1u << world_log); g++) {
      if (!peers[g]) return ZKP_ERR_INVALID_ARG;
      d.peer[g] = peers[g];
    }
    d.push = inverse ? 0 : 1;
    d.pull = inverse ? 1 : 0;
  }
  const uint32_t E = 1u << (r + a.q);
  const uint32_t tiles = 1u << (s_loc - a.q);
  uint32_t threads = NTT_THREADS;
  while (threads > 32 && threads > E / 2) threads >>= 1;
  ZKP_LAUNCH(ntt_dist_pass_kernel, dim3(tiles, 1, 1), dim3(threads), (size_t)E * sizeof(Fr), ctx->stream, a, d);
  ctx->ntt_launches = 1;
  return rt::check_last();
}

int ntt_dist_permute_dev(Ctx* ctx, const Fr* in, Fr* out, uint32_t log_n, uint32_t world_log, bool inverse) {
  const uint32_t r = ntt_dist_rows_log(log_n, world_log);
  if (r == 0 || in == out) return ZKP_ERR_INVALID_ARG;
  const uint32_t rows_log = r - world_log, cols_log = log_n - r - world_log;
  const size_t total = (size_t)1 << (log_n - world_log);
  size_t blocks = (total + 255) / 256;
  const size_t cap = (size_t)ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  ZKP_LAUNCH_NOSYNC(ntt_dist_permute_kernel, dim3((unsigned)blocks), dim3(256), 0, ctx->stream, in, out, rows_log, world_log,
             cols_log, inverse ? 1u : 0u);
  return rt::check_last();
}

int fr_pointwise_mul_dev(Ctx* ctx, Fr* a, const Fr* b, size_t n) {
  if (n == 0) return ZKP_OK;
  size_t blocks = (n + 255) / 256;
  const size_t cap = (size_t)ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  ZKP_LAUNCH_NOSYNC(fr_pointwise_mul_kernel, dim3((unsigned)blocks), dim3(256), 0, ctx->stream, a, b, n);
  return rt::check_last();
}

}  // namespace zkp
