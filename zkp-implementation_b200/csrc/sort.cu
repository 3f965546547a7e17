// Device radix sort of (key, value) u32 pairs and an exclusive u32 scan: the grouping step of the MSM
// ("a GPU radix sort of (bucket, point) pairs").  Least-significant-digit first, 8-bit digits, stable, so
// ceil(key_bits / 8) passes.  Every pass is three launches:
//
//   radix_hist_kernel      digit histogram of each tile of RS_TILE pairs          -> hist[digit][tile]
//   scan_u32_*             exclusive scan of hist (digit-major)                   -> first output slot of (digit, tile)
//   radix_scatter_kernel   per tile: stable rank of every pair inside its digit (warp match by ballots + per-warp
//                          counters), the tile is re-ordered by digit in shared memory and written out in runs
//
// The ranking is order-preserving (pairs with equal digits keep their input order), which is what makes the LSD
// passes compose; it also makes the whole MSM pipeline run-to-run deterministic.
#include <stdint.h>
#include <string.h>

#include "engine.h"

namespace zkp {

#ifdef ZKP_EMU
static constexpr uint32_t RS_THREADS = 64;   // emulated build: one OS thread per CUDA thread, keep the blocks small
#else
static constexpr uint32_t RS_THREADS = 256;
#endif
#ifndef ZKP_RS_MIN_BLOCKS
#define ZKP_RS_MIN_BLOCKS 4
#endif
static constexpr uint32_t RS_MIN_BLOCKS = ZKP_RS_MIN_BLOCKS;  // resident scatter blocks per SM (64 registers per thread)
static constexpr uint32_t RS_WARPS = RS_THREADS / 32;
#ifndef ZKP_RS_ITEMS
#define ZKP_RS_ITEMS 16
#endif
static constexpr uint32_t RS_ITEMS = ZKP_RS_ITEMS;             // pairs per thread
static constexpr uint32_t RS_TILE = RS_THREADS * RS_ITEMS;     // pairs per block
static constexpr uint32_t RS_DIGITS = 256;
static constexpr uint32_t SCAN_THREADS = RS_THREADS;
static constexpr uint32_t SCAN_ITEMS = 16;
static constexpr uint32_t SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// ---- exclusive scan of u32 --------------------------------------------------------------------------
// level kernel: inclusive scan of one tile in place (made exclusive by the caller's shift), tile sums out
__global__ void __launch_bounds__(SCAN_THREADS) scan_u32_tile_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                     uint32_t n, uint32_t* __restrict__ tile_sum) {
  __shared__ uint32_t sh[2][SCAN_THREADS];
  const uint32_t tid = threadIdx.x;
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)tid * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  uint32_t run = 0;
#pragma unroll
  for (uint32_t i = 0; i < SCAN_ITEMS; i++) {
    const uint32_t x = (base + i < n) ? in[base + i] : 0u;
    v[i] = run;  // exclusive within the thread
    run += x;
  }
  sh[0][tid] = run;
  __syncthreads();
  uint32_t cur = 0;
  for (uint32_t d = 1; d < SCAN_THREADS; d <<= 1) {
    uint32_t x = sh[cur][tid];
    if (tid >= d) x += sh[cur][tid - d];
    sh[cur ^ 1][tid] = x;
    cur ^= 1;
    __syncthreads();
  }
  const uint32_t pre = tid ? sh[cur][tid - 1] : 0u;
#pragma unroll
  for (uint32_t i = 0; i < SCAN_ITEMS; i++)
    if (base + i < n) out[base + i] = v[i] + pre;
  if (tid == SCAN_THREADS - 1 && tile_sum) tile_sum[blockIdx.x] = sh[cur][tid];
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_u32_add_kernel(uint32_t* __restrict__ data, uint32_t n,
                                                                    const uint32_t* __restrict__ tile_pre) {
  const uint32_t pre = tile_pre[blockIdx.x];
  const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (uint32_t i = 0; i < SCAN_ITEMS; i++)
    if (base + i < n) data[base + i] += pre;
}

static size_t scan_scratch_elems(size_t n) {
  size_t total = 0;
  while (n > SCAN_TILE) {
    n = (n + SCAN_TILE - 1) / SCAN_TILE;
    total += n;
  }
  return total + 1;
}

// out = exclusive scan of in (n entries; in == out allowed); scratch holds the per-tile sums of every level
static int scan_u32_rec(Ctx* ctx, const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* scratch) {
  const uint32_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  ZKP_LAUNCH(scan_u32_tile_kernel, dim3(tiles), dim3(SCAN_THREADS), 0, ctx->stream, in, out, n,
             tiles > 1 ? scratch : (uint32_t*)nullptr);
  ctx->sort_launches++;
  if (tiles > 1) {
    ZKP_TRY(scan_u32_rec(ctx, scratch, scratch, tiles, scratch + tiles));
    ZKP_LAUNCH_NOSYNC(scan_u32_add_kernel, dim3(tiles), dim3(SCAN_THREADS), 0, ctx->stream, out, n, (const uint32_t*)scratch);
    ctx->sort_launches++;
  }
  return ZKP_OK;
}

int scan_exclusive_u32_dev(Ctx* ctx, const uint32_t* in, uint32_t* out, uint32_t n) {
  if (!n) return ZKP_OK;
  ZKP_TRY(ctx->sort_scan.reserve(scan_scratch_elems(n) * sizeof(uint32_t)));
  ZKP_TRY(scan_u32_rec(ctx, in, out, n, ctx->sort_scan.as<uint32_t>()));
  return rt::check_last();
}

// ---- radix sort -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rs_digit(uint32_t key, uint32_t shift, uint32_t mask, uint32_t flip) {
  return ((key ^ flip) >> shift) & mask;
}

__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t shift,
                                                                uint32_t mask, uint32_t flip, uint32_t ntiles,
                                                                uint32_t* __restrict__ hist) {
  __shared__ uint32_t cnt[RS_DIGITS];
  const uint32_t tid = threadIdx.x;
  for (uint32_t d = tid; d < RS_DIGITS; d += RS_THREADS) cnt[d] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * RS_TILE;
  const uint32_t count = (n - base < RS_TILE) ? (uint32_t)(n - base) : RS_TILE;
  // four keys per 128-bit load when the tile is 16-byte aligned (the counts do not depend on the order): the scalar loop
  // kept one 4-byte load per thread in flight
  const uint32_t nvec = ((reinterpret_cast<uintptr_t>(keys + base) & 15u) == 0) ? (count >> 2) : 0u;
  const uint4* k4 = reinterpret_cast<const uint4*>(keys + base);
  for (uint32_t v = tid; v < nvec; v += RS_THREADS) {
    const uint4 q = k4[v];
    atomicAdd(&cnt[rs_digit(q.x, shift, mask, flip)], 1u);
    atomicAdd(&cnt[rs_digit(q.y, shift, mask, flip)], 1u);
    atomicAdd(&cnt[rs_digit(q.z, shift, mask, flip)], 1u);
    atomicAdd(&cnt[rs_digit(q.w, shift, mask, flip)], 1u);
  }
  for (uint32_t i = (nvec << 2) + tid; i < count; i += RS_THREADS)
    atomicAdd(&cnt[rs_digit(keys[base + i], shift, mask, flip)], 1u);
  __syncthreads();
  for (uint32_t d = tid; d <= mask; d += RS_THREADS) hist[(size_t)d * ntiles + blockIdx.x] = cnt[d];
}

__global__ void __launch_bounds__(RS_THREADS, RS_MIN_BLOCKS) radix_scatter_kernel(const uint32_t* __restrict__ kin, const uint32_t* __restrict__ vin,
                                                                   uint32_t* __restrict__ kout, uint32_t* __restrict__ vout,
                                                                   uint32_t n, uint32_t shift, uint32_t mask, uint32_t flip,
                                                                   uint32_t ntiles, const uint32_t* __restrict__ offs) {
  __shared__ uint32_t wcnt[RS_WARPS][RS_DIGITS];  // per-warp digit counts, then the warp's first slot inside the digit
  __shared__ uint32_t dstart[RS_DIGITS];          // first tile slot of every digit
  __shared__ uint32_t scan_tmp[2][RS_THREADS];
  ZKP_DYN_SMEM(uint32_t, skey);  // RS_TILE keys, then RS_TILE values
  uint32_t* sval = skey + RS_TILE;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t base = (size_t)blockIdx.x * RS_TILE;
  const uint32_t count = (n - base < RS_TILE) ? (uint32_t)(n - base) : RS_TILE;
  for (uint32_t d = tid; d < RS_WARPS * RS_DIGITS; d += RS_THREADS) (&wcnt[0][0])[d] = 0;
  __syncthreads();

  // phase A: item (warp, round, lane) -> rank among the warp's earlier items with the same digit.  Tile order is
  // warp-major, then round, then lane, so the rank order is the input order.
  uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
  const uint32_t wbase = warp * (32 * RS_ITEMS);
#pragma unroll
  for (uint32_t r = 0; r < RS_ITEMS; r++) {
    const uint32_t i = wbase + r * 32 + lane;
    const bool live = i < count;
    key[r] = live ? kin[base + i] : 0xffffffffu;
    val[r] = live ? vin[base + i] : 0u;
    const uint32_t d = live ? rs_digit(key[r], shift, mask, flip) : 0u;
    // lanes holding the same digit (MATCH.ANY); a dead lane gets a value no digit can take, so it matches nobody live
    const uint32_t peers = __match_any_sync(0xffffffffu, live ? d : (RS_DIGITS + lane));
    const uint32_t before = peers & ((1u << lane) - 1u);
    uint32_t prev = 0;
    if (live) prev = wcnt[warp][d];
    __syncwarp();
    if (live && before == 0) wcnt[warp][d] = prev + __popc(peers);  // the first peer books the whole group
    __syncwarp();
    rank[r] = prev + __popc(before);
  }
  __syncthreads();

  // phase B: per digit, exclusive prefix over the warps; digit totals -> exclusive scan over the digits
  uint32_t tot_local = 0;  // this thread's digits: tid, tid + RS_THREADS, ... (RS_DIGITS / RS_THREADS of them, or 0)
  for (uint32_t d = tid; d < RS_DIGITS; d += RS_THREADS) {
    uint32_t run = 0;
    for (uint32_t w = 0; w < RS_WARPS; w++) {
      const uint32_t c = wcnt[w][d];
      wcnt[w][d] = run;
      run += c;
    }
    dstart[d] = run;  // digit total for now
    tot_local += run;
  }
  __syncthreads();
  // scan of the 256 digit totals: thread t owns the contiguous digits [t * per, (t + 1) * per)
  constexpr uint32_t PER = (RS_DIGITS + RS_THREADS - 1) / RS_THREADS;
  uint32_t mine = 0;
  for (uint32_t k = 0; k < PER; k++) {
    const uint32_t d = tid * PER + k;
    if (d < RS_DIGITS) mine += dstart[d];
  }
  scan_tmp[0][tid] = mine;
  __syncthreads();
  uint32_t cur = 0;
  for (uint32_t s = 1; s < RS_THREADS; s <<= 1) {
    uint32_t x = scan_tmp[cur][tid];
    if (tid >= s) x += scan_tmp[cur][tid - s];
    scan_tmp[cur ^ 1][tid] = x;
    cur ^= 1;
    __syncthreads();
  }
  {
    uint32_t run = tid ? scan_tmp[cur][tid - 1] : 0u;
    for (uint32_t k = 0; k < PER; k++) {
      const uint32_t d = tid * PER + k;
      if (d < RS_DIGITS) {
        const uint32_t c = dstart[d];
        dstart[d] = run;
        run += c;
      }
    }
  }
  __syncthreads();

  // phase C: re-order the tile by digit in shared memory
#pragma unroll
  for (uint32_t r = 0; r < RS_ITEMS; r++) {
    const uint32_t i = wbase + r * 32 + lane;
    if (i < count) {
      const uint32_t d = rs_digit(key[r], shift, mask, flip);
      const uint32_t slot = dstart[d] + wcnt[warp][d] + rank[r];
      skey[slot] = key[r];
      sval[slot] = val[r];
    }
  }
  __syncthreads();

  // phase D: runs of equal digits go out to consecutive addresses; goff[d] = global slot of the digit's first
  // pair of this tile minus its tile slot, so the output position of tile slot j is goff[d] + j
  uint32_t* goff = &wcnt[0][0];  // the per-warp counters are dead after phase C
  for (uint32_t d = tid; d <= mask; d += RS_THREADS) goff[d] = offs[(size_t)d * ntiles + blockIdx.x] - dstart[d];
  __syncthreads();
  for (uint32_t j = tid; j < count; j += RS_THREADS) {
    const uint32_t k = skey[j];
    const uint32_t pos = goff[rs_digit(k, shift, mask, flip)] + j;
    kout[pos] = k;
    vout[pos] = sval[j];
  }
  (void)tot_local;
}

// Stable sort of n pairs by the low key_bits of the key (descending: by the complement).  The result lands in
// (*kres, *vres), which is either the (k0, v0) or the (k1, v1) pair of buffers; both pairs are clobbered.
int radix_sort_pairs_dev(Ctx* ctx, uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, uint32_t n, uint32_t key_bits,
                         bool descending, uint32_t** kres, uint32_t** vres) {
  *kres = k0;
  *vres = v0;
  if (n == 0 || key_bits == 0) return ZKP_OK;
  const uint32_t ntiles = (n + RS_TILE - 1) / RS_TILE;
  const size_t hist_n = (size_t)RS_DIGITS * ntiles;
  if (hist_n >= ((size_t)1 << 32)) return ZKP_ERR_INVALID_ARG;
  ZKP_TRY(ctx->sort_hist.reserve(hist_n * sizeof(uint32_t)));
  uint32_t* hist = ctx->sort_hist.as<uint32_t>();
  if (!ctx->sort_ready) {
    ZKP_TRY(rt::allow_smem((const void*)radix_scatter_kernel, 2 * RS_TILE * sizeof(uint32_t)));
    ctx->sort_ready = true;
  }
  const uint32_t flip = descending ? 0xffffffffu : 0u;
  uint32_t *kin = k0, *vin = v0, *kout = k1, *vout = v1;
  for (uint32_t shift = 0; shift < key_bits; shift += 8) {
    const uint32_t bits = (key_bits - shift < 8) ? (key_bits - shift) : 8;
    const uint32_t mask = (1u << bits) - 1;
    const uint32_t digits = mask + 1;
    ZKP_LAUNCH(radix_hist_kernel, dim3(ntiles), dim3(RS_THREADS), 0, ctx->stream, (const uint32_t*)kin, n, shift, mask, flip,
               ntiles, hist);
    ZKP_TRY(scan_exclusive_u32_dev(ctx, hist, hist, digits * ntiles));
    ZKP_LAUNCH(radix_scatter_kernel, dim3(ntiles), dim3(RS_THREADS), 2 * RS_TILE * sizeof(uint32_t), ctx->stream, (const uint32_t*)kin,
               (const uint32_t*)vin, kout, vout, n, shift, mask, flip, ntiles, (const uint32_t*)hist);
    ctx->sort_launches += 2;
    uint32_t* t = kin; kin = kout; kout = t;
    t = vin; vin = vout; vout = t;
  }
  *kres = kin;
  *vres = vin;
  return rt::check_last();
}

}  // namespace zkp
