// Engine context shared by the NTT, MSM and C-ABI translation units.
#pragma once
#include <map>
#include <mutex>
#include <vector>

#include "curve.cuh"
#include "runtime.h"

namespace zkp {

// Grow-only device allocation (scratch areas are sized by the largest call seen so far; 180 GB of
// HBM per B200 makes "keep everything resident" the right default for a prover session).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return ZKP_OK;
    rt::dev_free(p);
    p = nullptr;
    cap = 0;
    ZKP_TRY(rt::dev_malloc(&p, bytes));
    cap = bytes;
    return ZKP_OK;
  }
  void release() { rt::dev_free(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Per (log_n, direction) twiddle tables, device resident.
struct NttTables {
  uint32_t log_n = 0;
  bool inverse = false;
  uint32_t npass = 0;
  uint32_t digits[3] = {0, 0, 0};  // bits per pass, first pass = most significant input digit
  uint32_t lb = 0;                 // boundary tables: lo has 2^lb entries, hi has 2^(log_n - lb)
  // two-level inter-stage twiddles: only the distributed four-step stage uses them (the single-GPU passes read the
  // per-level tables Ctx::tw_fwd / tw_inv and have no inter-pass product)
  Fr* tw_lo = nullptr;             // omega_N^i            (inverse: omega_N^-i * (2^r)^-1)
  Fr* tw_hi = nullptr;             // omega_N^(i * 2^lb)   (inverse: omega_N^-(i * 2^lb))
  Fr* scale = nullptr;             // N^-1 (inverse transforms), else null
};

struct CosetTables {
  uint32_t log_n = 0;
  bool inverse = false;
  Fr offset;            // h (Montgomery)
  uint32_t lb = 0;
  bool scaled = false;  // inverse tables of the single-GPU path carry N^-1 in `lo` (saves the separate scaling product)
  Fr* lo = nullptr;     // forward: h^i ; inverse: h^-i (* N^-1 when `scaled`)
  Fr* hi = nullptr;     // forward: h^(i * 2^lb); inverse: h^-(i * 2^lb)
};

struct MsmScratch {
  DevBuf scalars;       // staged host scalars
  DevBuf bases;         // staged ad-hoc bases
  DevBuf keys_a, keys_b, vals_a, vals_b, sort_tmp;
  DevBuf bucket_start, bucket_end, task_meta, partials, seg_out, win_out;
  DevBuf misc;
  // batched-affine tree rounds (msm_affine.cu): ping-pong point buffers, denominator prefixes, inversion tree, run offsets
  DevBuf aff_a, aff_b, aff_pre, aff_inv, aff_off, aff_tb;
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = 0;
  bool own_stream = false;
  int sm_count = 148;
  std::mutex mu;        // the reference's KzgScheme is Send + Sync (plain data): calls are serialised

  // SRS bases resident on the device (kzg/src/srs.rs:14-21 `g1_points`)
  G1Affine* srs = nullptr;
  size_t srs_len = 0;
  // fixed-base table over the SRS (zkp_srs_precompute): window w = 2^(c w) * srs[i], w < srs_tab_windows
  // Records are srs_tab_stride16 x 16 bytes apart: 6 (packed, the default) or 8 (128-byte records, -DZKP_TABLE_PADDED).
  G1Affine* srs_tab = nullptr;
  uint32_t srs_tab_c = 0, srs_tab_windows = 0, srs_tab_stride16 = 6;
  G1Affine* srs0_tab = nullptr;  // 32 x 256 byte-window table of srs[0] for commit_para (built on first use)

  // NTT state
  // per-level butterfly twiddles, all orders in one array: T_k = [omega_{2^k}^i, i < 2^(k-1)] lives at
  // [2^(k-1), 2^k), k = 1 .. tw_log (grown to the largest transform seen; 2^tw_log * 32 bytes per direction)
  Fr* tw_fwd = nullptr;
  Fr* tw_inv = nullptr;
  uint32_t tw_log = 0;
  std::map<uint32_t, NttTables> ntt_tables;  // key = log_n * 2 + inverse
  std::vector<CosetTables> coset_tables;
  DevBuf ntt_scratch, ntt_io, ntt_io2;

  // polynomial layer (poly.cu): scan block totals, queued evaluation results and their partial sums
  static constexpr uint32_t EVAL_SLOTS = 16;
  static constexpr uint32_t EVAL_PARTIALS = 4096;  // blocks of 256 x 32 coefficients: up to 2^25 coefficients
  DevBuf poly_scratch, eval_out, eval_partials;
  uint32_t sort_launches = 0;   // kernels launched by sort.cu since the caller last cleared it
  bool sort_ready = false;      // dynamic shared memory of the scatter kernel opted in
  DevBuf sort_scan, sort_hist;  // sort.cu: tile sums of the u32 scan; digit histograms / offsets [digit][tile]

  // MSM state
  MsmScratch msm;
  uint32_t msm_window_bits = 0;  // 0 = choose from n
  int msm_affine_rounds = -1;    // batched-affine tree rounds before the XYZZ finish; -1 = choose from the mean bucket load
  uint32_t last_affine_rounds = 0;
  uint32_t msm_launches = 0;     // kernels launched by the last MSM (bench.py's gpu_launches)
  uint32_t ntt_launches = 0;

  // per-phase timing of the last MSM (events on `stream`; only when profiling is on)
  static constexpr int NPHASE = 5;
  bool profiling = false;
  void* phase_ev[NPHASE + 1] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float phase_ms[NPHASE] = {-1, -1, -1, -1, -1};
  // profiling of the batched-affine rounds: events around the first round's addition kernel (the longest launch of an
  // MSM) and the number of points entering / leaving every round (device counters read back after the final sync)
  void* aff_ev[2] = {nullptr, nullptr};
  float aff_add1_ms = -1;
  static constexpr int AFF_STATS = 8;
  uint32_t* aff_stats_dev = nullptr;      // [AFF_STATS]: [0] points before round 1, [r] points after round r
  uint32_t aff_stats[AFF_STATS] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t last_window_bits = 0, last_windows = 0;
};

// ---- NTT (ntt.cu) ----
int ntt_init(Ctx* ctx);
void ntt_destroy(Ctx* ctx);
// In-place batched transform on device memory.  coset == nullptr -> plain domain.
int ntt_run_dev(Ctx* ctx, Fr* data, uint32_t log_n, size_t batch, bool inverse, const Fr* coset_host);
int fr_pointwise_mul_dev(Ctx* ctx, Fr* a, const Fr* b, size_t n);  // a[i] *= b[i]
// Distributed four-step transform, one stage per call (the all-to-all sits between them; see ntt.cu)
uint32_t ntt_dist_rows_log(uint32_t log_n, uint32_t world_log);
int ntt_dist_stage_dev(Ctx* ctx, Fr* data, uint32_t log_n, uint32_t rank, uint32_t world_log, bool inverse,
                       const Fr* coset_host, Fr* const* peers);
int ntt_dist_permute_dev(Ctx* ctx, const Fr* in, Fr* out, uint32_t log_n, uint32_t world_log, bool inverse);

// ---- MSM (msm.cu) ----
// result: W window sums are reduced on the device; the final Horner over windows and the single
// Fq inversion run on the host (272 group operations out of ~n*W).
// fixed_c != 0: `bases_dev` is a fixed-base table built for window width fixed_c (windows table_stride apart)
int msm_run_dev(Ctx* ctx, const Fr* scalars_dev, const G1Affine* bases_dev, size_t n, G1Xyzz* out_host,
                uint32_t fixed_c = 0, size_t table_stride = 0);
// several MSMs over the same fixed-base table as ONE pipeline (one sort, one accumulate, one reduction)
int msm_run_multi_dev(Ctx* ctx, const Fr* const* scalars_list, const size_t* n_list, uint32_t count, const G1Affine* bases,
                      G1Xyzz* out_host, uint32_t fixed_c, size_t table_stride);
int msm_precompute_dev(Ctx* ctx, uint32_t window_bits);
// batched-affine tree rounds over the bucket-sorted point list (msm_affine.cu)
uint32_t msm_affine_choose_rounds(size_t total, size_t total_buckets);
size_t msm_affine_bound(size_t total, size_t nb, uint32_t rounds);
int msm_affine_rounds_dev(Ctx* ctx, uint32_t rounds, const uint32_t* svals, const G1Affine* bases, uint32_t base_stride16,
                          const uint32_t* bstart,
                          const uint32_t* bend, uint32_t nb, size_t total, const G1Affine** pts, const uint32_t** off,
                          size_t* bound);

// ---- sort / scan (sort.cu) ----
// Stable LSD radix sort of n (key, value) pairs by the low key_bits of the key; the result is in (*kres, *vres), one
// of the two buffer pairs (both are clobbered).
int radix_sort_pairs_dev(Ctx* ctx, uint32_t* k0, uint32_t* v0, uint32_t* k1, uint32_t* v1, uint32_t n, uint32_t key_bits,
                         bool descending, uint32_t** kres, uint32_t** vres);
int scan_exclusive_u32_dev(Ctx* ctx, const uint32_t* in, uint32_t* out, uint32_t n);
void msm_destroy(Ctx* ctx);

}  // namespace zkp
