// BLS12-381 G1 multi-scalar multiplication: signed-digit Pippenger on the device.
//
// Replaces the hot loops of `KzgScheme::evaluate_in_s` (kzg/src/scheme.rs:84-96):
//     coeffs.zip(points).map(|(c, s)| s.mul(c).into_affine()).reduce(|a, e| a.add(e).into_affine())
// i.e. sum_i c_i * P_i returned as the normalised affine point.  A group element has exactly one
// normalised affine representative, so the result is bit-identical to the reference's per-term
// double-and-add no matter how the sum is scheduled.
//
// Pipeline (all on one stream, no host synchronisation until the W window sums come back):
//   1. recode      scalars: Montgomery -> canonical -> W signed c-bit digits; one (bucket, point|sign)
//                  pair per digit, laid out window-major
//   2. sort        per window radix sort of the pairs by bucket (cub::DeviceRadixSort)
//   3. boundaries  start/end of every bucket's run in the sorted order
//   4. tasks       buckets longer than S_max are split so no thread owns an unbounded run; the task
//                  list is sorted by run length (longest first) so the 32 lanes of a warp finish together
//   5. accumulate  one thread per task: XYZZ accumulator += affine base (mixed add, 8M+2S),
//                  next base prefetched while the current add runs
//   6. reduce      per window: running-sum trick over segments of the bucket array, then a tree
//   7. host        Horner over the W window sums (c doublings + 1 add each) -- ~270 group
//                  operations out of ~n*W -- and the caller normalises with one Fq inversion
#include <string.h>

#include "engine.h"
#include "memops.cuh"

#ifndef ZKP_EMU
#include <cub/cub.cuh>
#else
#include <algorithm>
#include <numeric>
#endif

namespace zkp {

static constexpr uint32_t ACC_THREADS = 128;
static constexpr uint32_t RED_THREADS = 128;
static constexpr uint32_t SEG_LOG = 5;       // buckets per reduce thread = 32
static constexpr uint32_t SIGN_BIT = 0x80000000u;

struct MsmTask {
  uint32_t start;  // index into the window-major sorted value array
  uint32_t len;
};

// ---- 1. recode ---------------------------------------------------------------------------------
// keys[w*n + i] = |digit| - 1 (bucket index, weight |digit|) or nbuckets for a zero digit
// vals[w*n + i] = i | sign << 31
__global__ void __launch_bounds__(256) msm_recode_kernel(const Fr* __restrict__ scalars, uint32_t n, uint32_t c,
                                                         uint32_t nwin, uint32_t* __restrict__ keys,
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
    keys[o] = d ? (d - 1) : nbuckets;
    vals[o] = i | (neg && d ? SIGN_BIT : 0u);
  }
}

// ---- 3. bucket boundaries ------------------------------------------------------------------------
__global__ void __launch_bounds__(256) msm_bounds_kernel(const uint32_t* __restrict__ skeys, uint32_t n, uint32_t nwin,
                                                         uint32_t nbuckets, uint32_t* __restrict__ bstart,
                                                         uint32_t* __restrict__ bend) {
  const size_t total = (size_t)n * nwin;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; idx < total; idx += stride) {
    const uint32_t w = (uint32_t)(idx / n);
    const uint32_t i = (uint32_t)(idx - (size_t)w * n);
    const uint32_t key = skeys[idx];
    if (key >= nbuckets) continue;
    const size_t gb = (size_t)w * nbuckets + key;
    if (i == 0 || skeys[idx - 1] != key) bstart[gb] = (uint32_t)idx;
    if (i == n - 1 || skeys[idx + 1] != key) bend[gb] = (uint32_t)idx + 1;
  }
}

// ---- 4. tasks ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) msm_task_count_kernel(const uint32_t* __restrict__ bstart,
                                                             const uint32_t* __restrict__ bend, uint32_t total_buckets,
                                                             uint32_t smax, uint32_t* __restrict__ ntask) {
  const uint32_t gb = blockIdx.x * blockDim.x + threadIdx.x;
  if (gb >= total_buckets) return;
  const uint32_t len = bend[gb] - bstart[gb];
  ntask[gb] = (len + smax - 1) / smax;
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
__global__ void __launch_bounds__(ACC_THREADS) msm_accumulate_kernel(const MsmTask* __restrict__ tasks,
                                                                     const uint32_t* __restrict__ order, uint32_t ntasks,
                                                                     const uint32_t* __restrict__ svals,
                                                                     const G1Affine* __restrict__ bases,
                                                                     G1Xyzz* __restrict__ partials) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= ntasks) return;
  const uint32_t t = order[slot];  // longest runs first; partials stay in bucket order
  const MsmTask tk = tasks[t];
  const uint32_t* v = svals + tk.start;
  G1Xyzz acc = G1Xyzz::infinity();
  uint32_t pv = v[0];
  G1Affine nxt = ld_affine(bases + (pv & ~SIGN_BIT));
  for (uint32_t i = 0; i < tk.len; i++) {
    G1Affine cur = nxt;
    const uint32_t cv = pv;
    if (i + 1 < tk.len) {
      pv = v[i + 1];
      nxt = ld_affine(bases + (pv & ~SIGN_BIT));
    }
    if (cv & SIGN_BIT) cur = g1_neg(cur);
    xyzz_madd(acc, cur);
  }
  st_xyzz(partials + t, acc);
}

// ---- 6. reduce -----------------------------------------------------------------------------------
// One thread per segment of 2^SEG_LOG consecutive buckets of one window:
//   seg_out = sum_{b in segment} (b + 1) * bucket[b]
__global__ void __launch_bounds__(RED_THREADS) msm_segment_reduce_kernel(const G1Xyzz* __restrict__ partials,
                                                                         const uint32_t* __restrict__ task_off,
                                                                         const uint32_t* __restrict__ ntask,
                                                                         uint32_t nbuckets, uint32_t nwin, uint32_t seg_log,
                                                                         G1Xyzz* __restrict__ seg_out) {
  const uint32_t nseg = nbuckets >> seg_log;  // per window
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nseg * nwin) return;
  const uint32_t w = t / nseg, seg = t - w * nseg;
  const uint32_t lo = seg << seg_log;
  G1Xyzz run = G1Xyzz::infinity(), tot = G1Xyzz::infinity();
  for (int j = (1 << seg_log) - 1; j >= 0; j--) {
    const uint32_t gb = w * nbuckets + lo + (uint32_t)j;
    const uint32_t nt = ntask[gb], off = task_off[gb];
    for (uint32_t k = 0; k < nt; k++) {
      G1Xyzz p = ld_xyzz(partials + off + k);
      xyzz_add(run, p);
    }
    xyzz_add(tot, run);
  }
  if (lo) {
    G1Xyzz shifted = xyzz_mul_u32(run, lo);
    xyzz_add(tot, shifted);
  }
  st_xyzz(seg_out + t, tot);
}

// One block per window: sum the window's segment results.
__global__ void __launch_bounds__(RED_THREADS) msm_window_reduce_kernel(const G1Xyzz* __restrict__ seg_out, uint32_t nseg,
                                                                        G1Xyzz* __restrict__ win_out) {
  __shared__ G1Xyzz sh[RED_THREADS];
  const uint32_t w = blockIdx.x, tid = threadIdx.x;
  G1Xyzz acc = G1Xyzz::infinity();
  for (uint32_t i = tid; i < nseg; i += blockDim.x) {
    G1Xyzz p = ld_xyzz(seg_out + (size_t)w * nseg + i);
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
  if (tid == 0) st_xyzz(win_out + w, ld_xyzz(&sh[0]));
}

// ------------------------------------------------------------------------------------------------
// Host orchestration
// ------------------------------------------------------------------------------------------------
static uint32_t choose_window_bits(size_t n) {
  // cost model: 10 * n * W mixed adds + ~2.7 * 14 * 2^(c-1) * W reduction adds
  uint32_t best = 4;
  double best_cost = 1e300;
  for (uint32_t c = 4; c <= 20; c++) {
    const uint32_t W = 255 / c + 1;
    const double cost = 10.0 * (double)n * W + 2.7 * 14.0 * (double)(1u << (c - 1)) * W;
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}

static int sort_window(Ctx* ctx, const uint32_t* kin, uint32_t* kout, const uint32_t* vin, uint32_t* vout, uint32_t n,
                       uint32_t key_bits) {
#ifdef ZKP_EMU
  (void)ctx; (void)key_bits;
  std::vector<uint32_t> idx(n);
  std::iota(idx.begin(), idx.end(), 0u);
  std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return kin[a] < kin[b]; });
  for (uint32_t i = 0; i < n; i++) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
  return ZKP_OK;
#else
  size_t tmp = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin, kout, vin, vout, (int)n, 0, (int)key_bits, ctx->stream);
  if (e != cudaSuccess) return rt::wrap(e);
  ZKP_TRY(ctx->msm.sort_tmp.reserve(tmp));
  tmp = ctx->msm.sort_tmp.cap;
  e = cub::DeviceRadixSort::SortPairs(ctx->msm.sort_tmp.p, tmp, kin, kout, vin, vout, (int)n, 0, (int)key_bits, ctx->stream);
  return rt::wrap(e);
#endif
}

// order[] = task ids sorted by run length, longest first
static int sort_tasks_desc(Ctx* ctx, const uint32_t* len_in, uint32_t* len_out, const uint32_t* id_in, uint32_t* id_out,
                           uint32_t n, uint32_t key_bits) {
#ifdef ZKP_EMU
  (void)ctx; (void)key_bits;
  std::vector<uint32_t> idx(n);
  std::iota(idx.begin(), idx.end(), 0u);
  std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return len_in[a] > len_in[b]; });
  for (uint32_t i = 0; i < n; i++) { len_out[i] = len_in[idx[i]]; id_out[i] = id_in[idx[i]]; }
  return ZKP_OK;
#else
  size_t tmp = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp, len_in, len_out, id_in, id_out, (int)n, 0,
                                                            (int)key_bits, ctx->stream);
  if (e != cudaSuccess) return rt::wrap(e);
  ZKP_TRY(ctx->msm.sort_tmp.reserve(tmp));
  tmp = ctx->msm.sort_tmp.cap;
  e = cub::DeviceRadixSort::SortPairsDescending(ctx->msm.sort_tmp.p, tmp, len_in, len_out, id_in, id_out, (int)n, 0,
                                                (int)key_bits, ctx->stream);
  return rt::wrap(e);
#endif
}

static int exclusive_scan_u32(Ctx* ctx, const uint32_t* in, uint32_t* out, uint32_t n) {
#ifdef ZKP_EMU
  (void)ctx;
  uint32_t acc = 0;
  for (uint32_t i = 0; i < n; i++) { uint32_t v = in[i]; out[i] = acc; acc += v; }
  return ZKP_OK;
#else
  size_t tmp = 0;
  cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int)n, ctx->stream);
  if (e != cudaSuccess) return rt::wrap(e);
  ZKP_TRY(ctx->msm.sort_tmp.reserve(tmp));
  tmp = ctx->msm.sort_tmp.cap;
  e = cub::DeviceScan::ExclusiveSum(ctx->msm.sort_tmp.p, tmp, in, out, (int)n, ctx->stream);
  return rt::wrap(e);
#endif
}

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
#else
  (void)ctx;
#endif
}

int msm_run_dev(Ctx* ctx, const Fr* scalars, const G1Affine* bases, size_t n_, G1Xyzz* out_host) {
  ctx->msm_launches = 0;
  *out_host = G1Xyzz::infinity();
  if (n_ == 0) return ZKP_OK;  // scheme.rs:94 unwrap_or(G1Point::zero())
  if (n_ >= ((size_t)1 << 28)) return ZKP_ERR_INVALID_ARG;
  const uint32_t n = (uint32_t)n_;
  const uint32_t c = ctx->msm_window_bits ? ctx->msm_window_bits : choose_window_bits(n);
  const uint32_t nwin = 255 / c + 1;
  const uint32_t nbuckets = 1u << (c - 1);
  const uint32_t total_buckets = nbuckets * nwin;
  const size_t total = (size_t)n * nwin;
  if (total >= ((size_t)1 << 32)) return ZKP_ERR_INVALID_ARG;
  const uint32_t seg_log = (c - 1 < SEG_LOG) ? (c - 1) : SEG_LOG;
  const uint32_t nseg = nbuckets >> seg_log;
  // bound the longest run one thread owns: 4x the mean bucket load, at least 64
  uint32_t smax = (uint32_t)((4 * (size_t)n) / nbuckets);
  if (smax < 64) smax = 64;
  const size_t max_tasks = (size_t)total_buckets + total / smax + 1;

  MsmScratch& m = ctx->msm;
  ZKP_TRY(m.keys_a.reserve(total * 4));
  ZKP_TRY(m.keys_b.reserve(total * 4));
  ZKP_TRY(m.vals_a.reserve(total * 4));
  ZKP_TRY(m.vals_b.reserve(total * 4));
  ZKP_TRY(m.bucket_start.reserve((size_t)total_buckets * 4));
  ZKP_TRY(m.bucket_end.reserve((size_t)total_buckets * 4));
  ZKP_TRY(m.misc.reserve((size_t)(total_buckets + 1) * 8));
  ZKP_TRY(m.task_meta.reserve(max_tasks * (sizeof(MsmTask) + 4 * sizeof(uint32_t))));
  ZKP_TRY(m.partials.reserve(max_tasks * sizeof(G1Xyzz)));
  ZKP_TRY(m.seg_out.reserve((size_t)nseg * nwin * sizeof(G1Xyzz)));
  ZKP_TRY(m.win_out.reserve((size_t)(nwin + 1) * sizeof(G1Xyzz)));
  uint32_t* keys_a = m.keys_a.as<uint32_t>();
  uint32_t* keys_b = m.keys_b.as<uint32_t>();
  uint32_t* vals_a = m.vals_a.as<uint32_t>();
  uint32_t* vals_b = m.vals_b.as<uint32_t>();
  uint32_t* bstart = m.bucket_start.as<uint32_t>();
  uint32_t* bend = m.bucket_end.as<uint32_t>();
  uint32_t* ntask = m.misc.as<uint32_t>();
  uint32_t* task_off = ntask + (total_buckets + 1);
  MsmTask* tasks = m.task_meta.as<MsmTask>();
  uint32_t* task_len = reinterpret_cast<uint32_t*>(tasks + max_tasks);
  uint32_t* task_id = task_len + max_tasks;
  uint32_t* task_len_sorted = task_id + max_tasks;
  uint32_t* task_order = task_len_sorted + max_tasks;
  G1Xyzz* partials = m.partials.as<G1Xyzz>();
  G1Xyzz* seg_out = m.seg_out.as<G1Xyzz>();
  G1Xyzz* win_out = m.win_out.as<G1Xyzz>();
  cudaStream_t st = ctx->stream;

  ctx->last_window_bits = c;
  ctx->last_windows = nwin;
  // 1. recode
  phase_mark(ctx, 0);
  ZKP_LAUNCH(msm_recode_kernel, dim3((n + 255) / 256), dim3(256), 0, st, scalars, n, c, nwin, keys_a, vals_a);
  ctx->msm_launches++;
  phase_mark(ctx, 1);
  // 2. sort each window by bucket (c bits: bucket index plus the zero-digit sentinel)
  for (uint32_t w = 0; w < nwin; w++) {
    ZKP_TRY(sort_window(ctx, keys_a + (size_t)w * n, keys_b + (size_t)w * n, vals_a + (size_t)w * n,
                        vals_b + (size_t)w * n, n, c));
    ctx->msm_launches += 3;
  }
  phase_mark(ctx, 2);
  // 3. boundaries
  ZKP_TRY(rt::dev_memset(bstart, 0, (size_t)total_buckets * 4, st));
  ZKP_TRY(rt::dev_memset(bend, 0, (size_t)total_buckets * 4, st));
  {
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 32;
    if (blocks > cap) blocks = cap;
    ZKP_LAUNCH(msm_bounds_kernel, dim3((unsigned)blocks), dim3(256), 0, st, keys_b, n, nwin, nbuckets, bstart, bend);
    ctx->msm_launches++;
  }
  // 4. tasks
  ZKP_TRY(rt::dev_memset(ntask + total_buckets, 0, 4, st));
  ZKP_LAUNCH(msm_task_count_kernel, dim3((total_buckets + 255) / 256), dim3(256), 0, st, bstart, bend, total_buckets, smax,
             ntask);
  ZKP_TRY(exclusive_scan_u32(ctx, ntask, task_off, total_buckets + 1));
  ZKP_LAUNCH(msm_task_build_kernel, dim3((total_buckets + 255) / 256), dim3(256), 0, st, bstart, bend, task_off,
             total_buckets, smax, tasks, task_len, task_id);
  ctx->msm_launches += 3;
  uint32_t ntasks = 0;
  ZKP_TRY(rt::d2h(&ntasks, task_off + total_buckets, 4, st));
  ZKP_TRY(rt::sync(st));
  // 5. accumulate
  phase_mark(ctx, 3);
  if (ntasks) {
    uint32_t len_bits = 1;
    while ((1u << len_bits) <= smax) len_bits++;
    ZKP_TRY(sort_tasks_desc(ctx, task_len, task_len_sorted, task_id, task_order, ntasks, len_bits));
    ZKP_LAUNCH(msm_accumulate_kernel, dim3((ntasks + ACC_THREADS - 1) / ACC_THREADS), dim3(ACC_THREADS), 0, st, tasks,
               task_order, ntasks, vals_b, bases, partials);
    ctx->msm_launches += 3;
  }
  // 6. reduce
  phase_mark(ctx, 4);
  ZKP_LAUNCH(msm_segment_reduce_kernel, dim3((nseg * nwin + RED_THREADS - 1) / RED_THREADS), dim3(RED_THREADS), 0, st,
             partials, task_off, ntask, nbuckets, nwin, seg_log, seg_out);
  ZKP_LAUNCH(msm_window_reduce_kernel, dim3(nwin), dim3(RED_THREADS), 0, st, seg_out, nseg, win_out);
  ctx->msm_launches += 2;
  phase_mark(ctx, 5);
  ZKP_TRY(rt::check_last());
  // 7. host: Horner over windows
  std::vector<G1Xyzz> wins(nwin);
  ZKP_TRY(rt::d2h(wins.data(), win_out, (size_t)nwin * sizeof(G1Xyzz), st));
  ZKP_TRY(rt::sync(st));
  phase_collect(ctx);
  G1Xyzz acc = wins[nwin - 1];
  for (int w = (int)nwin - 2; w >= 0; w--) {
    for (uint32_t k = 0; k < c; k++) acc = xyzz_dbl(acc);
    xyzz_add(acc, wins[w]);
  }
  *out_host = acc;
  return ZKP_OK;
}

void msm_destroy(Ctx* ctx) {
#ifndef ZKP_EMU
  for (int i = 0; i <= Ctx::NPHASE; i++)
    if (ctx->phase_ev[i]) { cudaEventDestroy((cudaEvent_t)ctx->phase_ev[i]); ctx->phase_ev[i] = nullptr; }
#endif
  MsmScratch& m = ctx->msm;
  DevBuf* all[] = {&m.scalars, &m.bases, &m.keys_a, &m.keys_b, &m.vals_a, &m.vals_b, &m.sort_tmp, &m.bucket_start,
                   &m.bucket_end, &m.task_meta, &m.partials, &m.seg_out, &m.win_out, &m.misc};
  for (DevBuf* b : all) b->release();
}

}  // namespace zkp
