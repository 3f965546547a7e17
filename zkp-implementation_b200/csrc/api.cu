// extern "C" boundary (include/zkp_b200.h): context management, host-buffer staging, result
// normalisation.  No torch types, plain pointers and sizes.
#include <string.h>

#include <new>

#include "../../include/zkp_b200.h"
#include "../../include/zkp_plonk.h"
#include <thread>

#include "engine.h"
#include "poly.h"

using namespace zkp;

struct zkp_ctx {
  Ctx c;
  // Single-process multi-GPU context (zkp_ctx_create_multi): this object is shard 0, `peers` are the contexts of the
  // other devices.  The resident SRS is split by point range (shard g holds [shard_lo[g], shard_lo[g + 1])); every
  // commitment against it runs on all shards at once and the partial sums are folded on the host.
  std::vector<zkp_ctx*> peers;
  std::vector<size_t> shard_lo;
  size_t srs_total = 0;
  bool multi() const { return !peers.empty(); }
  size_t nshards() const { return peers.size() + 1; }
  zkp_ctx* shard(size_t g) { return g == 0 ? this : peers[g - 1]; }
};

namespace zkp {
int gen_bases_dev(Ctx* ctx, uint64_t seed, size_t n, G1Affine* out);
int gen_bases_range_dev(Ctx* ctx, uint64_t seed, size_t first, size_t n, G1Affine* out);
int gen_srs_dev(Ctx* ctx, const Fr& secret, size_t start, size_t n, G1Affine* out);
int bench_imad(Ctx* ctx, double* wide, double* lo);
}  // namespace zkp

static void write_affine(const G1Xyzz& p, uint64_t out_xy[12], uint8_t* out_inf) {
  G1Affine a = xyzz_to_affine(p);
  memcpy(out_xy, &a, sizeof(a));
  if (out_inf) *out_inf = a.is_inf() ? 1 : 0;
}

// count results normalised with ONE Fq inversion (Montgomery's trick over the ZZ * ZZZ of the finite entries)
static void write_affine_batch(const G1Xyzz* p, uint32_t count, uint64_t* out_xy, uint8_t* out_inf) {
  std::vector<Fq> pre(count);
  Fq prod = Fq::one();
  for (uint32_t i = 0; i < count; i++) {
    pre[i] = prod;
    if (!p[i].is_inf()) prod = prod * (p[i].zz * p[i].zzz);
  }
  Fq inv = fq_inv_gcd(prod);
  for (uint32_t i = count; i-- > 0;) {
    G1Affine a = G1Affine::infinity();
    if (!p[i].is_inf()) {
      const Fq ti = inv * pre[i];  // (zz * zzz)^-1
      inv = inv * (p[i].zz * p[i].zzz);
      a.x = p[i].x * (ti * p[i].zzz);  // X / ZZ
      a.y = p[i].y * (ti * p[i].zz);   // Y / ZZZ
    }
    memcpy(out_xy + 12 * (size_t)i, &a, sizeof(a));
    if (out_inf) out_inf[i] = a.is_inf() ? 1 : 0;
  }
}

extern "C" {

const char* zkp_strerror(int status) {
  switch (status) {
    case ZKP_OK: return "ok";
    case ZKP_ERR_INVALID_ARG: return "invalid argument";
    case ZKP_ERR_CUDA: return "CUDA runtime error";
    case ZKP_ERR_OOM: return "out of device memory";
    case ZKP_ERR_SRS_TOO_SMALL: return "SRS shorter than the polynomial (g1_points.len() > polynomial.degree() violated)";
    case ZKP_ERR_DOMAIN_TOO_LARGE: return "evaluation domain too large";
    case ZKP_ERR_NO_DEVICE: return "no usable CUDA device (this engine has no CPU fallback)";
    case ZKP_ERR_EMPTY_POLY: return "empty polynomial (at least 1 coefficient expected)";
    default: return "unknown status";
  }
}

int zkp_ctx_create(zkp_ctx** out, int device) {
  if (!out) return ZKP_ERR_INVALID_ARG;
  *out = nullptr;
  if (device < 0 || device >= rt::device_count()) return ZKP_ERR_NO_DEVICE;
  ZKP_TRY(rt::set_device(device));
  zkp_ctx* h = new (std::nothrow) zkp_ctx();
  if (!h) return ZKP_ERR_OOM;
  h->c.device = device;
  h->c.sm_count = rt::sm_count(device);
  if (h->c.sm_count <= 0) h->c.sm_count = 148;
#ifndef ZKP_EMU
  if (cudaStreamCreateWithFlags(&h->c.stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return ZKP_ERR_CUDA;
  }
  h->c.own_stream = true;
#endif
  int st = ntt_init(&h->c);
  if (st != ZKP_OK) {
    zkp_ctx_destroy(h);
    return st;
  }
  *out = h;
  return ZKP_OK;
}

int zkp_ctx_create_multi(zkp_ctx** out, const int* device_ids, int n_devices) {
  if (!out || !device_ids || n_devices < 1 || n_devices > 16) return ZKP_ERR_INVALID_ARG;
  *out = nullptr;
  zkp_ctx* h = nullptr;
  ZKP_TRY(zkp_ctx_create(&h, device_ids[0]));
  for (int g = 1; g < n_devices; g++) {
    zkp_ctx* p = nullptr;
    const int st = zkp_ctx_create(&p, device_ids[g]);
    if (st != ZKP_OK) {
      zkp_ctx_destroy(h);
      return st;
    }
    h->peers.push_back(p);
  }
#ifndef ZKP_EMU
  // peer access where the topology offers it (NVLink / NVSwitch); copies fall back to staging otherwise
  for (int a = 0; a < n_devices; a++)
    for (int b = 0; b < n_devices; b++) {
      int can = 0;
      if (device_ids[a] == device_ids[b]) continue;
      if (cudaDeviceCanAccessPeer(&can, device_ids[a], device_ids[b]) == cudaSuccess && can) {
        cudaSetDevice(device_ids[a]);
        if (cudaDeviceEnablePeerAccess(device_ids[b], 0) != cudaSuccess) cudaGetLastError();  // already enabled
      }
    }
  cudaSetDevice(device_ids[0]);
#endif
  h->shard_lo.assign((size_t)n_devices + 1, 0);
  *out = h;
  return ZKP_OK;
}

int zkp_ctx_shards(const zkp_ctx* h) { return h ? (int)(h->peers.size() + 1) : 0; }

void zkp_ctx_destroy(zkp_ctx* h) {
  if (!h) return;
  for (zkp_ctx* p : h->peers) zkp_ctx_destroy(p);
  h->peers.clear();
  rt::set_device(h->c.device);
  rt::sync(h->c.stream);
  ntt_destroy(&h->c);
  msm_destroy(&h->c);
  rt::dev_free(h->c.srs);
  rt::dev_free(h->c.srs_tab);
  rt::dev_free(h->c.srs0_tab);
#ifndef ZKP_EMU
  if (h->c.own_stream) cudaStreamDestroy(h->c.stream);
#endif
  delete h;
}

int zkp_ctx_set_stream(zkp_ctx* h, void* stream) {
  if (!h) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(h->c.mu);
#ifndef ZKP_EMU
  rt::sync(h->c.stream);
  if (h->c.own_stream) cudaStreamDestroy(h->c.stream);
  h->c.stream = (cudaStream_t)stream;
  h->c.own_stream = false;
#else
  (void)stream;
#endif
  return ZKP_OK;
}

int zkp_ctx_synchronize(zkp_ctx* h) {
  if (!h) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(h->c.mu);
  ZKP_TRY(rt::set_device(h->c.device));
  return rt::sync(h->c.stream);
}

int zkp_ctx_set_msm_window(zkp_ctx* h, uint32_t bits) {
  if (!h || (bits != 0 && (bits < 2 || bits > 22))) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->c.mu);
  h->c.msm_window_bits = bits;
  return ZKP_OK;
}

int zkp_ctx_set_msm_affine(zkp_ctx* h, int rounds) {
  if (!h || rounds < -1 || rounds > 30) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->c.mu);
  h->c.msm_affine_rounds = rounds;
  return ZKP_OK;
}

int zkp_ctx_set_profiling(zkp_ctx* h, int on) {
  if (!h) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->c.mu);
  h->c.profiling = on != 0;
  return ZKP_OK;
}

double zkp_ctx_last_phase_ms(zkp_ctx* h, int phase) {
  if (h && phase == 100) return (double)h->c.aff_add1_ms;  // first-round batched-affine addition kernel
  if (h && phase >= 200 && phase < 200 + Ctx::AFF_STATS) return (double)h->c.aff_stats[phase - 200];  // points per round
  if (!h || phase < 0 || phase >= Ctx::NPHASE) return -1.0;
  return (double)h->c.phase_ms[phase];
}

int zkp_ctx_last_launches(zkp_ctx* h, int kind) {
  if (h && kind == 2) return (int)h->c.last_window_bits;
  if (h && kind == 3) return (int)h->c.last_windows;
  if (h && kind == 4) return (int)h->c.last_affine_rounds;
  if (!h) return 0;
  return kind == 0 ? (int)h->c.msm_launches : (int)h->c.ntt_launches;
}

// ---- SRS ----------------------------------------------------------------------------------------
static int stage_affine(Ctx* c, const uint64_t* xy, const uint8_t* inf, size_t n, G1Affine* dev) {
  if (!inf) return rt::h2d(dev, xy, n * sizeof(G1Affine), c->stream);
  // honour ark-ec's `infinity: bool`: flagged entries become the (0, 0) sentinel
  std::vector<G1Affine> tmp(n);
  memcpy(tmp.data(), xy, n * sizeof(G1Affine));
  for (size_t i = 0; i < n; i++)
    if (inf[i]) tmp[i] = G1Affine::infinity();
  ZKP_TRY(rt::h2d(dev, tmp.data(), n * sizeof(G1Affine), c->stream));
  return rt::sync(c->stream);
}

static int srs_alloc(Ctx* c, size_t n) {
  rt::dev_free(c->srs);
  rt::dev_free(c->srs_tab);
  rt::dev_free(c->srs0_tab);
  c->srs = nullptr;
  c->srs_tab = nullptr;
  c->srs0_tab = nullptr;
  c->srs_tab_c = 0;
  c->srs_len = 0;
  ZKP_TRY(rt::dev_malloc((void**)&c->srs, n * sizeof(G1Affine)));
  c->srs_len = n;
  return ZKP_OK;
}

// point range of every shard: contiguous, the first n % G shards hold one extra point (dist.shard_range)
static void split_srs(zkp_ctx* h, size_t n) {
  const size_t G = h->nshards();
  h->shard_lo.assign(G + 1, 0);
  for (size_t g = 0; g < G; g++) h->shard_lo[g + 1] = h->shard_lo[g] + n / G + (g < n % G ? 1 : 0);
  h->srs_total = n;
}

}  // extern "C" (templates need C++ linkage)
// run fn(g) for every shard, peers on their own host threads (each sets its device); returns the first error
template <class F>
static int for_each_shard(zkp_ctx* h, F fn) {
  const size_t G = h->nshards();
  std::vector<int> st(G, ZKP_OK);
#ifdef ZKP_EMU
  for (size_t g = 0; g < G; g++) st[g] = fn(g);
#else
  std::vector<std::thread> th;
  for (size_t g = 1; g < G; g++) th.emplace_back([&, g] { st[g] = fn(g); });
  st[0] = fn(0);
  for (auto& t : th) t.join();
#endif
  for (size_t g = 0; g < G; g++)
    if (st[g] != ZKP_OK) return st[g];
  return ZKP_OK;
}
extern "C" {

int zkp_srs_upload(zkp_ctx* h, const uint64_t* xy, const uint8_t* infinity, size_t n) {
  if (!h || (n && !xy)) return ZKP_ERR_INVALID_ARG;
  if (h->multi()) {
    std::lock_guard<std::mutex> g0(h->c.mu);
    split_srs(h, n);
    return for_each_shard(h, [&](size_t g) {
      zkp_ctx* s = h->shard(g);
      std::unique_lock<std::mutex> lk(s->c.mu, std::defer_lock);
      if (g) lk.lock();
      const size_t lo = h->shard_lo[g], m = h->shard_lo[g + 1] - lo;
      ZKP_TRY(rt::set_device(s->c.device));
      ZKP_TRY(srs_alloc(&s->c, m));
      ZKP_TRY(stage_affine(&s->c, xy + 12 * lo, infinity ? infinity + lo : nullptr, m, s->c.srs));
      return rt::sync(s->c.stream);
    });
  }
  std::lock_guard<std::mutex> g(h->c.mu);
  ZKP_TRY(rt::set_device(h->c.device));
  ZKP_TRY(srs_alloc(&h->c, n));
  ZKP_TRY(stage_affine(&h->c, xy, infinity, n, h->c.srs));
  return rt::sync(h->c.stream);
}

int zkp_srs_upload_dev(zkp_ctx* h, const void* xy_dev, size_t n) {
  if (!h || (n && !xy_dev)) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(h->c.mu);
  ZKP_TRY(rt::set_device(h->c.device));
  ZKP_TRY(srs_alloc(&h->c, n));
  ZKP_TRY(rt::d2d(h->c.srs, xy_dev, n * sizeof(G1Affine), h->c.stream));
  return rt::sync(h->c.stream);
}

size_t zkp_srs_len(const zkp_ctx* h) { return h ? (h->multi() ? h->srs_total : h->c.srs_len) : 0; }

int zkp_srs_precompute(zkp_ctx* h, uint32_t window_bits) {
  if (!h) return ZKP_ERR_INVALID_ARG;
  if (h->multi()) {
    std::lock_guard<std::mutex> g0(h->c.mu);
    return for_each_shard(h, [&](size_t g) {
      zkp_ctx* s = h->shard(g);
      std::unique_lock<std::mutex> lk(s->c.mu, std::defer_lock);
      if (g) lk.lock();
      ZKP_TRY(rt::set_device(s->c.device));
      return msm_precompute_dev(&s->c, window_bits);
    });
  }
  std::lock_guard<std::mutex> g(h->c.mu);
  ZKP_TRY(rt::set_device(h->c.device));
  return msm_precompute_dev(&h->c, window_bits);
}

int zkp_srs_generate(zkp_ctx* h, const uint64_t secret[4], size_t n, uint64_t* xy_out) {
  return zkp_srs_generate_range(h, secret, 0, n, xy_out);
}

int zkp_srs_generate_range(zkp_ctx* h, const uint64_t secret[4], size_t first, size_t n, uint64_t* xy_out) {
  if (!h || !secret) return ZKP_ERR_INVALID_ARG;
  if (h->multi()) {
    std::lock_guard<std::mutex> g0(h->c.mu);
    split_srs(h, n);
    Fr sec;
    memcpy(sec.v, secret, 32);
    return for_each_shard(h, [&](size_t g) {
      zkp_ctx* s = h->shard(g);
      std::unique_lock<std::mutex> lk(s->c.mu, std::defer_lock);
      if (g) lk.lock();
      const size_t lo = h->shard_lo[g], m = h->shard_lo[g + 1] - lo;
      ZKP_TRY(rt::set_device(s->c.device));
      ZKP_TRY(srs_alloc(&s->c, m));
      ZKP_TRY(gen_srs_dev(&s->c, sec, first + lo, m, s->c.srs));
      if (xy_out) ZKP_TRY(rt::d2h(xy_out + 12 * lo, s->c.srs, m * sizeof(G1Affine), s->c.stream));
      return rt::sync(s->c.stream);
    });
  }
  std::lock_guard<std::mutex> g(h->c.mu);
  ZKP_TRY(rt::set_device(h->c.device));
  ZKP_TRY(srs_alloc(&h->c, n));
  Fr s;
  memcpy(s.v, secret, 32);
  ZKP_TRY(gen_srs_dev(&h->c, s, first, n, h->c.srs));
  if (xy_out) ZKP_TRY(rt::d2h(xy_out, h->c.srs, n * sizeof(G1Affine), h->c.stream));
  return rt::sync(h->c.stream);
}

// ---- MSM ----------------------------------------------------------------------------------------
static int msm_nolock(Ctx* c, const void* scalars_dev, const void* bases_dev, size_t n, G1Xyzz* acc) {
  const G1Affine* bases = (const G1Affine*)bases_dev;
  if (!bases) {
    if (n > c->srs_len) return ZKP_ERR_SRS_TOO_SMALL;
    // fixed-base table: its window width was tuned for the whole SRS, so short prefixes stay windowed
    if (c->srs_tab && n >= c->srs_len / 4)
      return msm_run_dev(c, (const Fr*)scalars_dev, c->srs_tab, n, acc, c->srs_tab_c, c->srs_len);
    bases = c->srs;
  }
  return msm_run_dev(c, (const Fr*)scalars_dev, bases, n, acc);
}

// Multi-context commitment against the sharded resident SRS: shard g runs its Pippenger on scalars [lo_g, min(hi_g, n))
// (copied from the host buffer, or across devices from the primary device's buffer), all shards at once; the XYZZ
// partials are folded here (G - 1 additions).  The caller holds the primary context's mutex.
static int msm_sharded_nolock(zkp_ctx* h, const void* scalars, bool on_host, size_t n, G1Xyzz* acc) {
  if (n > h->srs_total) return ZKP_ERR_SRS_TOO_SMALL;
  const size_t G = h->nshards();
  std::vector<G1Xyzz> part(G, G1Xyzz::infinity());
  if (!on_host) ZKP_TRY(rt::sync(h->c.stream));  // the primary stream may still be producing the scalars
  ZKP_TRY(for_each_shard(h, [&](size_t g) {
    zkp_ctx* s = h->shard(g);
    const size_t lo = h->shard_lo[g], hi = h->shard_lo[g + 1] < n ? h->shard_lo[g + 1] : n;
    if (lo >= hi) return (int)ZKP_OK;
    std::unique_lock<std::mutex> lk(s->c.mu, std::defer_lock);
    if (g) lk.lock();
    Ctx* c = &s->c;
    const size_t m = hi - lo;
    ZKP_TRY(rt::set_device(c->device));
    const uint8_t* src = (const uint8_t*)scalars + lo * sizeof(Fr);
    const void* dev_scalars = src;
    if (on_host || g) {  // shard 0 reads device scalars in place
      ZKP_TRY(c->msm.scalars.reserve(m * sizeof(Fr)));
      ZKP_TRY(on_host ? rt::h2d(c->msm.scalars.p, src, m * sizeof(Fr), c->stream)
                      : rt::copy_any(c->msm.scalars.p, src, m * sizeof(Fr), c->stream));
      dev_scalars = c->msm.scalars.p;
    }
    return msm_nolock(c, dev_scalars, nullptr, m, &part[g]);
  }));
  rt::set_device(h->c.device);
  *acc = part[0];
  for (size_t g = 1; g < G; g++) xyzz_add(*acc, part[g]);
  uint32_t launches = 0;
  for (size_t g = 0; g < G; g++) launches += h->shard(g)->c.msm_launches;
  h->c.msm_launches = launches;
  return ZKP_OK;
}

int zkp_msm_g1_partial_dev(zkp_ctx* h, const void* scalars_dev, const void* bases_dev, size_t n, uint64_t out_xyzz[24]) {
  if (!h || !out_xyzz || (n && !scalars_dev)) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(h->c.mu);
  ZKP_TRY(rt::set_device(h->c.device));
  G1Xyzz acc;
  ZKP_TRY(msm_nolock(&h->c, scalars_dev, bases_dev, n, &acc));
  memcpy(out_xyzz, &acc, sizeof(acc));
  return ZKP_OK;
}

int zkp_msm_g1_dev(zkp_ctx* h, const void* scalars_dev, const void* bases_dev, size_t n, uint64_t out_xy[12],
                   uint8_t* out_infinity) {
  if (!h || !out_xy || (n && !scalars_dev)) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(h->c.mu);
  ZKP_TRY(rt::set_device(h->c.device));
  G1Xyzz acc;
  if (h->multi() && !bases_dev) ZKP_TRY(msm_sharded_nolock(h, scalars_dev, false, n, &acc));
  else ZKP_TRY(msm_nolock(&h->c, scalars_dev, bases_dev, n, &acc));
  write_affine(acc, out_xy, out_infinity);
  return ZKP_OK;
}

int zkp_msm_g1_multi_dev(zkp_ctx* h, uint32_t count, const void* const* scalars_dev, const size_t* lens, uint64_t* out_xy,
                         uint8_t* out_infinity) {
  if (!h || (count && (!scalars_dev || !lens || !out_xy)) || count > 16) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(h->c.mu);
  Ctx* c = &h->c;
  ZKP_TRY(rt::set_device(c->device));
  G1Xyzz acc[16];
  if (h->multi()) {  // one sharded commitment after the other (the batch pipeline is per device)
    uint32_t launches = 0;
    for (uint32_t j = 0; j < count; j++) {
      if (lens[j] && !scalars_dev[j]) return ZKP_ERR_INVALID_ARG;
      ZKP_TRY(msm_sharded_nolock(h, scalars_dev[j], false, lens[j], &acc[j]));
      launches += c->msm_launches;
    }
    c->msm_launches = launches;
    write_affine_batch(acc, count, out_xy, out_infinity);
    return ZKP_OK;
  }
  size_t shortest = (size_t)-1;
  for (uint32_t j = 0; j < count; j++) {
    if (lens[j] > c->srs_len) return ZKP_ERR_SRS_TOO_SMALL;
    if (lens[j] && !scalars_dev[j]) return ZKP_ERR_INVALID_ARG;
    if (lens[j] < shortest) shortest = lens[j];
  }
  if (count > 1 && c->srs_tab && shortest >= c->srs_len / 4) {
    ZKP_TRY(msm_run_multi_dev(c, (const Fr* const*)scalars_dev, lens, count, c->srs_tab, acc, c->srs_tab_c, c->srs_len));
  } else {
    uint32_t launches = 0;
    for (uint32_t j = 0; j < count; j++) {
      ZKP_TRY(msm_nolock(c, scalars_dev[j], nullptr, lens[j], &acc[j]));
      launches += c->msm_launches;
    }
    c->msm_launches = launches;
  }
  write_affine_batch(acc, count, out_xy, out_infinity);
  return ZKP_OK;
}

int zkp_msm_g1_multi_partial_dev(zkp_ctx* h, uint32_t count, const void* const* scalars_dev, const size_t* lens,
                                 uint64_t* out_xyzz) {
  if (!h || (count && (!scalars_dev || !lens || !out_xyzz)) || count > 16) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(h->c.mu);
  Ctx* c = &h->c;
  ZKP_TRY(rt::set_device(c->device));
  size_t shortest = (size_t)-1;
  for (uint32_t j = 0; j < count; j++) {
    if (lens[j] > c->srs_len) return ZKP_ERR_SRS_TOO_SMALL;
    if (lens[j] && !scalars_dev[j]) return ZKP_ERR_INVALID_ARG;
    if (lens[j] < shortest) shortest = lens[j];
  }
  G1Xyzz acc[16];
  if (count > 1 && c->srs_tab && shortest >= c->srs_len / 4) {
    ZKP_TRY(msm_run_multi_dev(c, (const Fr* const*)scalars_dev, lens, count, c->srs_tab, acc, c->srs_tab_c, c->srs_len));
  } else {
    uint32_t launches = 0;
    for (uint32_t j = 0; j < count; j++) {
      ZKP_TRY(msm_nolock(c, scalars_dev[j], nullptr, lens[j], &acc[j]));
      launches += c->msm_launches;
    }
    c->msm_launches = launches;
  }
  memcpy(out_xyzz, acc, (size_t)count * sizeof(G1Xyzz));
  return ZKP_OK;
}

int zkp_g1_fold_partials(const uint64_t* partials, size_t count, uint64_t out_xy[12], uint8_t* out_infinity) {
  if (!out_xy || (count && !partials)) return ZKP_ERR_INVALID_ARG;
  G1Xyzz acc = G1Xyzz::infinity();
  for (size_t i = 0; i < count; i++) {
    G1Xyzz p;
    memcpy(&p, partials + 24 * i, sizeof(p));
    xyzz_add(acc, p);
  }
  write_affine(acc, out_xy, out_infinity);
  return ZKP_OK;
}

int zkp_msm_g1(zkp_ctx* h, const uint64_t* scalars, size_t n, uint64_t out_xy[12], uint8_t* out_infinity) {
  if (!h || !out_xy || (n && !scalars)) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(h->c.mu);
  Ctx* c = &h->c;
  if (h->multi()) {
    G1Xyzz acc;
    ZKP_TRY(msm_sharded_nolock(h, scalars, true, n, &acc));
    write_affine(acc, out_xy, out_infinity);
    return ZKP_OK;
  }
  if (n > c->srs_len) return ZKP_ERR_SRS_TOO_SMALL;
  ZKP_TRY(rt::set_device(c->device));
  ZKP_TRY(c->msm.scalars.reserve(n * sizeof(Fr)));
  ZKP_TRY(rt::h2d(c->msm.scalars.p, scalars, n * sizeof(Fr), c->stream));
  G1Xyzz acc;
  ZKP_TRY(msm_nolock(c, c->msm.scalars.p, nullptr, n, &acc));
  write_affine(acc, out_xy, out_infinity);
  return ZKP_OK;
}

int zkp_msm_g1_bases(zkp_ctx* h, const uint64_t* scalars, const uint64_t* xy, const uint8_t* infinity, size_t n,
                     uint64_t out_xy[12], uint8_t* out_infinity) {
  if (!h || !out_xy || (n && (!scalars || !xy))) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(h->c.mu);
  Ctx* c = &h->c;
  ZKP_TRY(rt::set_device(c->device));
  G1Xyzz acc = G1Xyzz::infinity();
  if (n) {
    ZKP_TRY(c->msm.scalars.reserve(n * sizeof(Fr)));
    ZKP_TRY(c->msm.bases.reserve(n * sizeof(G1Affine)));
    ZKP_TRY(rt::h2d(c->msm.scalars.p, scalars, n * sizeof(Fr), c->stream));
    ZKP_TRY(stage_affine(c, xy, infinity, n, c->msm.bases.as<G1Affine>()));
    ZKP_TRY(msm_nolock(c, c->msm.scalars.p, c->msm.bases.p, n, &acc));
  }
  write_affine(acc, out_xy, out_infinity);
  return ZKP_OK;
}

// ---- NTT ----------------------------------------------------------------------------------------
static int ntt_nolock(Ctx* c, void* data_dev, uint32_t log_n, size_t batch, int inverse, const uint64_t* coset) {
  Fr hh;
  if (coset) memcpy(hh.v, coset, 32);
  return ntt_run_dev(c, (Fr*)data_dev, log_n, batch, inverse != 0, coset ? &hh : nullptr);
}

int zkp_ntt_fr_dev(zkp_ctx* h, void* data_dev, uint32_t log_n, size_t batch, int inverse, const uint64_t* coset) {
  if (!h || (batch && !data_dev)) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(h->c.mu);
  ZKP_TRY(rt::set_device(h->c.device));
  return ntt_nolock(&h->c, data_dev, log_n, batch, inverse, coset);
}

int zkp_ntt_fr(zkp_ctx* h, uint64_t* data, uint32_t log_n, size_t batch, int inverse, const uint64_t* coset) {
  if (!h || (batch && !data)) return ZKP_ERR_INVALID_ARG;
  if (log_n > 27) return ZKP_ERR_DOMAIN_TOO_LARGE;
  const size_t bytes = ((size_t)batch << log_n) * sizeof(Fr);
  std::lock_guard<std::mutex> g(h->c.mu);
  ZKP_TRY(rt::set_device(h->c.device));
  ZKP_TRY(h->c.ntt_io.reserve(bytes));
  void* dev = h->c.ntt_io.p;
  ZKP_TRY(rt::h2d(dev, data, bytes, h->c.stream));
  ZKP_TRY(ntt_nolock(&h->c, dev, log_n, batch, inverse, coset));
  ZKP_TRY(rt::d2h(data, dev, bytes, h->c.stream));
  return rt::sync(h->c.stream);
}

int zkp_fr_mul_pointwise_dev(zkp_ctx* h, void* a_dev, const void* b_dev, size_t n) {
  if (!h || (n && (!a_dev || !b_dev))) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(h->c.mu);
  ZKP_TRY(rt::set_device(h->c.device));
  return fr_pointwise_mul_dev(&h->c, (Fr*)a_dev, (const Fr*)b_dev, n);
}

int zkp_poly_mul_fr(zkp_ctx* h, const uint64_t* a, size_t la, const uint64_t* b, size_t lb, uint64_t* out) {
  if (!h) return ZKP_ERR_INVALID_ARG;
  if (la == 0 || lb == 0) return ZKP_OK;  // ark-poly: zero * anything = zero polynomial
  if (!a || !b || !out) return ZKP_ERR_INVALID_ARG;
  const size_t lo = la + lb - 1;
  uint32_t log_n = 0;
  while (((size_t)1 << log_n) < lo) log_n++;
  if (log_n > 27) return ZKP_ERR_DOMAIN_TOO_LARGE;
  const size_t N = (size_t)1 << log_n;
  std::lock_guard<std::mutex> g(h->c.mu);
  Ctx* c = &h->c;
  ZKP_TRY(rt::set_device(c->device));
  ZKP_TRY(c->ntt_io.reserve(2 * N * sizeof(Fr)));
  Fr* da = c->ntt_io.as<Fr>();
  Fr* db = da + N;
  ZKP_TRY(rt::dev_memset(da, 0, 2 * N * sizeof(Fr), c->stream));
  ZKP_TRY(rt::h2d(da, a, la * sizeof(Fr), c->stream));
  ZKP_TRY(rt::h2d(db, b, lb * sizeof(Fr), c->stream));
  ZKP_TRY(ntt_run_dev(c, da, log_n, 2, false, nullptr));  // both operands as one batch of two
  uint32_t launches = c->ntt_launches;
  ZKP_TRY(fr_pointwise_mul_dev(c, da, db, N));
  ZKP_TRY(ntt_run_dev(c, da, log_n, 1, true, nullptr));
  c->ntt_launches += launches + 1;
  ZKP_TRY(rt::d2h(out, da, lo * sizeof(Fr), c->stream));
  return rt::sync(c->stream);
}

// ---- multi-GPU four-step NTT ---------------------------------------------------------------------
static int world_log_of(uint32_t world, uint32_t* wl) {
  uint32_t l = 0;
  while ((1u << l) < world) l++;
  if (world == 0 || (1u << l) != world || l > 3) return ZKP_ERR_INVALID_ARG;
  *wl = l;
  return ZKP_OK;
}

uint32_t zkp_ntt_dist_rows_log(uint32_t log_n, uint32_t world) {
  uint32_t wl;
  if (world_log_of(world, &wl) != ZKP_OK) return 0;
  return ntt_dist_rows_log(log_n, wl);
}

int zkp_ntt_dist_stage_dev(zkp_ctx* h, void* data_dev, uint32_t log_n, uint32_t rank, uint32_t world, int inverse,
                           const uint64_t* coset, void* const* peer_bufs) {
  if (!h || !data_dev) return ZKP_ERR_INVALID_ARG;
  uint32_t wl;
  ZKP_TRY(world_log_of(world, &wl));
  std::lock_guard<std::mutex> g(h->c.mu);
  ZKP_TRY(rt::set_device(h->c.device));
  Fr hh;
  if (coset) memcpy(hh.v, coset, 32);
  return ntt_dist_stage_dev(&h->c, (Fr*)data_dev, log_n, rank, wl, inverse != 0, coset ? &hh : nullptr,
                            (Fr* const*)peer_bufs);
}

int zkp_ntt_dist_permute_dev(zkp_ctx* h, const void* in_dev, void* out_dev, uint32_t log_n, uint32_t world, int inverse) {
  if (!h || !in_dev || !out_dev) return ZKP_ERR_INVALID_ARG;
  uint32_t wl;
  ZKP_TRY(world_log_of(world, &wl));
  std::lock_guard<std::mutex> g(h->c.mu);
  ZKP_TRY(rt::set_device(h->c.device));
  return ntt_dist_permute_dev(&h->c, (const Fr*)in_dev, (Fr*)out_dev, log_n, wl, inverse != 0);
}

int zkp_dev_alloc(zkp_ctx* h, size_t bytes, void** out_dev) {
  if (!h || !out_dev) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->c.mu);  // reads the context's stream / device: serialised with zkp_ctx_set_stream
  ZKP_TRY(rt::set_device(h->c.device));
  return rt::dev_malloc(out_dev, bytes);
}

int zkp_dev_free(zkp_ctx* h, void* dev) {
  if (!h) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->c.mu);  // reads the context's stream / device: serialised with zkp_ctx_set_stream
  ZKP_TRY(rt::set_device(h->c.device));
  rt::sync(h->c.stream);
  rt::dev_free(dev);
  return ZKP_OK;
}

int zkp_dev_copy(zkp_ctx* h, void* dst_dev, const void* src_dev, size_t bytes) {
  if (!h || (bytes && (!dst_dev || !src_dev))) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lk(h->c.mu);  // reads the context's stream / device: serialised with zkp_ctx_set_stream
  ZKP_TRY(rt::set_device(h->c.device));
  return rt::d2d(dst_dev, src_dev, bytes, h->c.stream);
}

int zkp_ipc_export(zkp_ctx* h, const void* dev, uint8_t handle[64]) {
  if (!h || !dev || !handle) return ZKP_ERR_INVALID_ARG;
  ZKP_TRY(rt::set_device(h->c.device));
  return rt::ipc_export(dev, handle);
}

int zkp_ipc_open(zkp_ctx* h, const uint8_t handle[64], void** out_dev) {
  if (!h || !handle || !out_dev) return ZKP_ERR_INVALID_ARG;
  ZKP_TRY(rt::set_device(h->c.device));
  return rt::ipc_open(handle, out_dev);
}

int zkp_ipc_close(zkp_ctx* h, void* dev) {
  if (!h) return ZKP_ERR_INVALID_ARG;
  ZKP_TRY(rt::set_device(h->c.device));
  return rt::ipc_close(dev);
}

// ---- device-resident Fr vectors ------------------------------------------------------------------
static Fr fr_of(const uint64_t* p) {
  Fr r;
  memcpy(r.v, p, 32);
  return r;
}
#define ZKP_ENTER(h)                              \
  if (!(h)) return ZKP_ERR_INVALID_ARG;           \
  std::lock_guard<std::mutex> g((h)->c.mu);       \
  ZKP_TRY(rt::set_device((h)->c.device));         \
  Ctx* c = &(h)->c

int zkp_dev_upload(zkp_ctx* h, void* dst_dev, const void* src_host, size_t bytes) {
  ZKP_ENTER(h);
  if (bytes && (!dst_dev || !src_host)) return ZKP_ERR_INVALID_ARG;
  ZKP_TRY(rt::h2d(dst_dev, src_host, bytes, c->stream));
  return rt::sync(c->stream);
}

int zkp_dev_download(zkp_ctx* h, void* dst_host, const void* src_dev, size_t bytes) {
  ZKP_ENTER(h);
  if (bytes && (!dst_host || !src_dev)) return ZKP_ERR_INVALID_ARG;
  ZKP_TRY(rt::d2h(dst_host, src_dev, bytes, c->stream));
  return rt::sync(c->stream);
}

int zkp_dev_zero(zkp_ctx* h, void* dst_dev, size_t bytes) {
  ZKP_ENTER(h);
  if (bytes && !dst_dev) return ZKP_ERR_INVALID_ARG;
  return rt::dev_memset(dst_dev, 0, bytes, c->stream);
}

int zkp_fr_powers_dev(zkp_ctx* h, void* out_dev, const uint64_t base[4], const uint64_t first[4], size_t n) {
  ZKP_ENTER(h);
  if (!base || !first || (n && !out_dev)) return ZKP_ERR_INVALID_ARG;
  return fr_powers_dev(c, (Fr*)out_dev, fr_of(base), fr_of(first), n);
}

int zkp_fr_batch_inverse_dev(zkp_ctx* h, void* data_dev, size_t n) {
  ZKP_ENTER(h);
  if (n && !data_dev) return ZKP_ERR_INVALID_ARG;
  return fr_batch_inverse_dev(c, (Fr*)data_dev, n);
}

int zkp_fr_scan_dev(zkp_ctx* h, void* data_dev, size_t n, int op, int reverse) {
  ZKP_ENTER(h);
  if ((n && !data_dev) || (op != 0 && op != 1)) return ZKP_ERR_INVALID_ARG;
  return fr_scan_dev(c, (Fr*)data_dev, n, op, reverse != 0);
}

int zkp_fr_lincomb_dev(zkp_ctx* h, void* out_dev, size_t out_len, uint32_t count, const void* const* polys_dev,
                       const size_t* lens, const uint64_t* coefs, const uint64_t* c0) {
  ZKP_ENTER(h);
  if ((out_len && !out_dev) || count > LincombArgs::MAX_TERMS || (count && (!polys_dev || !lens || !coefs)))
    return ZKP_ERR_INVALID_ARG;
  LincombArgs a;
  memset(&a, 0, sizeof(a));
  a.count = count;
  for (uint32_t k = 0; k < count; k++) {
    if (lens[k] && !polys_dev[k]) return ZKP_ERR_INVALID_ARG;
    a.p[k] = (const Fr*)polys_dev[k];
    a.len[k] = lens[k];
    a.coef[k] = fr_of(coefs + 4 * k);
  }
  if (c0) {
    a.has_c0 = 1;
    a.c0 = fr_of(c0);
  }
  return fr_lincomb_dev(c, (Fr*)out_dev, out_len, a);
}

int zkp_fr_add_at_dev(zkp_ctx* h, void* data_dev, uint32_t count, const size_t* idx, const uint64_t* vals) {
  ZKP_ENTER(h);
  if (count > SparseAddArgs::MAX_TERMS || (count && (!data_dev || !idx || !vals))) return ZKP_ERR_INVALID_ARG;
  SparseAddArgs a;
  memset(&a, 0, sizeof(a));
  a.count = count;
  for (uint32_t k = 0; k < count; k++) {
    a.idx[k] = idx[k];
    a.val[k] = fr_of(vals + 4 * k);
  }
  return fr_add_at_dev(c, (Fr*)data_dev, a);
}

int zkp_fr_eval_dev(zkp_ctx* h, uint32_t count, const void* const* polys_dev, const size_t* lens, const uint64_t* xs,
                    uint64_t* out) {
  ZKP_ENTER(h);
  if (count >= Ctx::EVAL_SLOTS || (count && (!polys_dev || !lens || !xs || !out))) return ZKP_ERR_INVALID_ARG;
  Fr pts[Ctx::EVAL_SLOTS];
  for (uint32_t k = 0; k < count; k++) {
    if (lens[k] && !polys_dev[k]) return ZKP_ERR_INVALID_ARG;
    pts[k] = fr_of(xs + 4 * k);
  }
  ZKP_TRY(fr_eval_batch_dev(c, count, reinterpret_cast<const Fr* const*>(polys_dev), lens, pts));
  return fr_eval_fetch(c, (Fr*)out, count);
}

int zkp_fr_trimmed_len_dev(zkp_ctx* h, const void* coeffs_dev, size_t n, size_t* out_len) {
  ZKP_ENTER(h);
  if (!out_len || (n && !coeffs_dev)) return ZKP_ERR_INVALID_ARG;
  return fr_trimmed_len_dev(c, (const Fr*)coeffs_dev, n, out_len);
}

int zkp_g1_mul_srs0(zkp_ctx* h, const uint64_t* scalars, uint32_t count, uint64_t* out_xy) {
  ZKP_ENTER(h);
  if (count > 64 || (count && (!scalars || !out_xy))) return ZKP_ERR_INVALID_ARG;
  if (count && c->srs_len == 0) return ZKP_ERR_SRS_TOO_SMALL;
  G1Xyzz tmp[64];
  ZKP_TRY(g1_scalar_mul_dev(c, (const Fr*)scalars, count, tmp));
  write_affine_batch(tmp, count, out_xy, nullptr);
  return ZKP_OK;
}

// ---- PLONK pointwise kernels (include/zkp_plonk.h) ---------------------------------------------------
int zkp_plonk_numden_dev(zkp_ctx* h, const zkp_plonk_numden_args* a) {
  ZKP_ENTER(h);
  if (!a) return ZKP_ERR_INVALID_ARG;
  PlonkNumDenArgs p;
  p.a = (const Fr*)a->a_dev; p.b = (const Fr*)a->b_dev; p.c = (const Fr*)a->c_dev;
  p.s1 = (const Fr*)a->s1_dev; p.s2 = (const Fr*)a->s2_dev; p.s3 = (const Fr*)a->s3_dev;
  p.roots = (const Fr*)a->roots_dev;
  p.beta = fr_of(a->beta);
  p.gamma = fr_of(a->gamma);
  p.beta_k1 = fp_mul(p.beta, fr_of(a->k1));
  p.beta_k2 = fp_mul(p.beta, fr_of(a->k2));
  p.n = a->n;
  p.num = (Fr*)a->num_dev;
  p.den = (Fr*)a->den_dev;
  if (p.n && (!p.a || !p.b || !p.c || !p.s1 || !p.s2 || !p.s3 || !p.roots || !p.num || !p.den)) return ZKP_ERR_INVALID_ARG;
  return plonk_numden_dev(c, p);
}

int zkp_plonk_quotient_dev(zkp_ctx* h, const zkp_plonk_quotient_args* a) {
  ZKP_ENTER(h);
  if (!a) return ZKP_ERR_INVALID_ARG;
  PlonkQuotientArgs p;
  p.a = (const Fr*)a->a_dev; p.b = (const Fr*)a->b_dev; p.c = (const Fr*)a->c_dev; p.z = (const Fr*)a->z_dev;
  p.ql = (const Fr*)a->ql_dev; p.qr = (const Fr*)a->qr_dev; p.qo = (const Fr*)a->qo_dev; p.qm = (const Fr*)a->qm_dev;
  p.qc = (const Fr*)a->qc_dev; p.pi = (const Fr*)a->pi_dev;
  p.s1 = (const Fr*)a->s1_dev; p.s2 = (const Fr*)a->s2_dev; p.s3 = (const Fr*)a->s3_dev;
  p.l1 = (const Fr*)a->l1_dev; p.x = (const Fr*)a->x_dev;
  p.beta = fr_of(a->beta);
  p.gamma = fr_of(a->gamma);
  p.alpha = fr_of(a->alpha);
  p.alpha2 = fp_mul(p.alpha, p.alpha);
  p.beta_k1 = fp_mul(p.beta, fr_of(a->k1));
  p.beta_k2 = fp_mul(p.beta, fr_of(a->k2));
  for (int i = 0; i < 8; i++) p.zh_inv[i] = fr_of(a->zh_inv[i]);
  p.d = a->d;
  p.rho = a->rho;
  p.t = (Fr*)a->t_dev;
  if (p.d && (!p.a || !p.b || !p.c || !p.z || !p.ql || !p.qr || !p.qo || !p.qm || !p.qc || !p.pi || !p.s1 || !p.s2 ||
              !p.s3 || !p.l1 || !p.x || !p.t))
    return ZKP_ERR_INVALID_ARG;
  return plonk_quotient_dev(c, p);
}

int zkp_plonk_gate_check_dev(zkp_ctx* h, const void* const cols_dev[9], size_t n, int* ok) {
  ZKP_ENTER(h);
  if (!cols_dev || !ok) return ZKP_ERR_INVALID_ARG;
  const Fr* cols[9];
  for (int i = 0; i < 9; i++) {
    if (n && !cols_dev[i]) return ZKP_ERR_INVALID_ARG;
    cols[i] = (const Fr*)cols_dev[i];
  }
  bool good = true;
  ZKP_TRY(plonk_gate_check_dev(c, cols, n, &good));
  *ok = good ? 1 : 0;
  return ZKP_OK;
}

// ---- radix sort / scan --------------------------------------------------------------------------------
int zkp_sort_pairs_dev(zkp_ctx* h, void* keys_dev, void* vals_dev, size_t n, uint32_t key_bits, int descending) {
  ZKP_ENTER(h);
  if (n >= ((size_t)1 << 31) || key_bits > 32 || (n && (!keys_dev || !vals_dev))) return ZKP_ERR_INVALID_ARG;
  if (!n) return ZKP_OK;
  ZKP_TRY(c->msm.sort_tmp.reserve(2 * n * sizeof(uint32_t)));
  uint32_t* k1 = c->msm.sort_tmp.as<uint32_t>();
  uint32_t* v1 = k1 + n;
  uint32_t *kres = nullptr, *vres = nullptr;
  ZKP_TRY(radix_sort_pairs_dev(c, (uint32_t*)keys_dev, (uint32_t*)vals_dev, k1, v1, (uint32_t)n, key_bits, descending != 0,
                               &kres, &vres));
  if (kres != keys_dev) {
    ZKP_TRY(rt::d2d(keys_dev, kres, n * sizeof(uint32_t), c->stream));
    ZKP_TRY(rt::d2d(vals_dev, vres, n * sizeof(uint32_t), c->stream));
  }
  return rt::check_last();
}

int zkp_scan_exclusive_u32_dev(zkp_ctx* h, const void* in_dev, void* out_dev, size_t n) {
  ZKP_ENTER(h);
  if (n >= ((size_t)1 << 32) || (n && (!in_dev || !out_dev))) return ZKP_ERR_INVALID_ARG;
  return scan_exclusive_u32_dev(c, (const uint32_t*)in_dev, (uint32_t*)out_dev, (uint32_t)n);
}

// ---- synthetic workloads / microbenchmarks ------------------------------------------------------
int zkp_g1_generate_bases_dev(zkp_ctx* h, uint64_t seed, size_t n, void* bases_dev) {
  if (!h || (n && !bases_dev)) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(h->c.mu);
  ZKP_TRY(rt::set_device(h->c.device));
  return gen_bases_dev(&h->c, seed, n, (G1Affine*)bases_dev);
}

int zkp_g1_generate_bases_range_dev(zkp_ctx* h, uint64_t seed, size_t first, size_t n, void* bases_dev) {
  if (!h || (n && !bases_dev)) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(h->c.mu);
  ZKP_TRY(rt::set_device(h->c.device));
  return gen_bases_range_dev(&h->c, seed, first, n, (G1Affine*)bases_dev);
}

int zkp_bench_imad_peak(zkp_ctx* h, double* wide, double* lo) {
  if (!h || !wide || !lo) return ZKP_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> g(h->c.mu);
  ZKP_TRY(rt::set_device(h->c.device));
  return bench_imad(&h->c, wide, lo);
}

}  // extern "C"
