// CPU backend of the C ABI (include/zkp_b200.h) -- TEST / BASELINE INFRASTRUCTURE, never loaded by the product.
//
// Purpose: time the reference's OWN prover algorithm on the GPU box's host cores (bench.py `extra.plonk_cpu_*`) and
// cross-check the host orchestration against the golden proofs without a GPU.  zkp-implementation_b200/host/plonk.cpp
// (the restatement of plonk/src/circuit.rs:166-197 and plonk/src/prover.rs:61-293 above the C ABI) is linked, unchanged,
// against this file instead of the CUDA engine; every entry point it needs is implemented with the oracle's CPU
// kernels (oracle/zkp_oracle.c):
//   zkp_msm_g1        -> orc_msm_naive      kzg/src/scheme.rs:84-96 literally (per-term double-and-add + into_affine),
//                        or orc_msm_pippenger on all cores when the context was created with pippenger = 1
//   zkp_ntt_fr(_dev)  -> orc_ntt            ark-poly radix-2 in-order transform
//   zkp_poly_mul_fr   -> orc_poly_mul       `&a * &b`
// "Device" memory is host memory here.  Entry points the reference-shaped provers never call return an error.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../include/zkp_plonk.h"
#include "../zkp-implementation_b200/host/mont_host.hpp"

extern "C" {
void orc_msm_naive(const uint64_t* scalars_mont, const uint64_t* bases_xy, size_t n, uint64_t* out_xy);
void orc_msm_pippenger(const uint64_t* scalars_mont, const uint64_t* bases_xy, size_t n, uint64_t* out_xy, int threads);
void orc_srs(const uint64_t secret_mont[4], size_t count, uint64_t* out_xy);
void orc_ntt(uint64_t* data, int log_n, int inverse, const uint64_t* coset_mont, int threads);
void orc_poly_mul(const uint64_t* a, size_t la, const uint64_t* b, size_t lb, uint64_t* out, int threads);
}

using zkp_host::Fr;

struct zkp_ctx {
  std::vector<uint64_t> srs;  // n x 12
  size_t srs_len = 0;
  int threads = 1;    // 1: single-threaded like the reference (no rayon, no `parallel` feature)
  int pippenger = 0;  // 0: the reference's per-term algorithm; 1: bucket method (best-effort CPU line)
};

enum { ERR_ARG = 1, ERR_UNSUPPORTED = 2, ERR_SRS_TOO_SMALL = 4 };

extern "C" {

int zkp_cpu_ctx_create(zkp_ctx** out, int threads, int pippenger) {
  if (!out) return ERR_ARG;
  zkp_ctx* c = new zkp_ctx();
  c->threads = threads;
  c->pippenger = pippenger;
  *out = c;
  return 0;
}
void zkp_cpu_ctx_destroy(zkp_ctx* c) { delete c; }

int zkp_srs_upload(zkp_ctx* c, const uint64_t* xy, const uint8_t* infinity, size_t n) {
  if (!c || (n && !xy)) return ERR_ARG;
  c->srs.assign(xy, xy + 12 * n);
  if (infinity)
    for (size_t i = 0; i < n; i++)
      if (infinity[i]) memset(&c->srs[12 * i], 0, 96);
  c->srs_len = n;
  return 0;
}
int zkp_srs_generate(zkp_ctx* c, const uint64_t secret[4], size_t n, uint64_t* xy_out) {  // srs.rs:48-69
  if (!c) return ERR_ARG;
  c->srs.assign(12 * n, 0);
  orc_srs(secret, n, c->srs.data());
  c->srs_len = n;
  if (xy_out) memcpy(xy_out, c->srs.data(), 96 * n);
  return 0;
}
size_t zkp_srs_len(const zkp_ctx* c) { return c ? c->srs_len : 0; }
const char* zkp_strerror(int status) {
  switch (status) {
    case 0: return "ok";
    case ERR_SRS_TOO_SMALL: return "assertion failed: g1_points.len() > polynomial.degree() (kzg/src/scheme.rs:86)";
    case ERR_UNSUPPORTED: return "entry point not provided by the CPU reference backend";
    default: return "error";
  }
}
int zkp_ctx_synchronize(zkp_ctx*) { return 0; }

int zkp_msm_g1(zkp_ctx* c, const uint64_t* scalars, size_t n, uint64_t out_xy[12], uint8_t* out_infinity) {
  if (!c || !out_xy) return ERR_ARG;
  if (n > c->srs_len) return ERR_SRS_TOO_SMALL;  // scheme.rs:86
  if (c->pippenger) orc_msm_pippenger(scalars, c->srs.data(), n, out_xy, c->threads);
  else orc_msm_naive(scalars, c->srs.data(), n, out_xy);
  if (out_infinity) {
    uint64_t o = 0;
    for (int i = 0; i < 12; i++) o |= out_xy[i];
    *out_infinity = o == 0;
  }
  return 0;
}

int zkp_msm_g1_multi_dev(zkp_ctx* c, uint32_t count, const void* const* scalars, const size_t* lens, uint64_t* out_xy,
                         uint8_t* out_infinity) {
  for (uint32_t j = 0; j < count; j++) {
    const int st = zkp_msm_g1(c, (const uint64_t*)scalars[j], lens[j], out_xy + 12 * j, out_infinity ? out_infinity + j : nullptr);
    if (st) return st;
  }
  return 0;
}

int zkp_g1_mul_srs0(zkp_ctx* c, const uint64_t* scalars, uint32_t count, uint64_t* out_xy) {  // scheme.rs:78-82
  if (!c || c->srs_len == 0) return ERR_SRS_TOO_SMALL;
  for (uint32_t k = 0; k < count; k++) orc_msm_naive(scalars + 4 * k, c->srs.data(), 1, out_xy + 12 * k);
  return 0;
}

int zkp_ntt_fr(zkp_ctx* c, uint64_t* data, uint32_t log_n, size_t batch, int inverse, const uint64_t* coset) {
  if (!c) return ERR_ARG;
  for (size_t b = 0; b < batch; b++) orc_ntt(data + ((b << log_n) * 4), (int)log_n, inverse, coset, c->threads);
  return 0;
}
int zkp_ntt_fr_dev(zkp_ctx* c, void* data, uint32_t log_n, size_t batch, int inverse, const uint64_t* coset) {
  return zkp_ntt_fr(c, (uint64_t*)data, log_n, batch, inverse, coset);
}
int zkp_poly_mul_fr(zkp_ctx* c, const uint64_t* a, size_t la, const uint64_t* b, size_t lb, uint64_t* out) {
  if (!c) return ERR_ARG;
  if (la && lb) orc_poly_mul(a, la, b, lb, out, c->threads);
  return 0;
}

int zkp_dev_alloc(zkp_ctx*, size_t bytes, void** out) { *out = malloc(bytes ? bytes : 1); return *out ? 0 : 3; }
int zkp_dev_free(zkp_ctx*, void* p) { free(p); return 0; }
int zkp_dev_copy(zkp_ctx*, void* d, const void* s, size_t n) { memmove(d, s, n); return 0; }
int zkp_dev_upload(zkp_ctx*, void* d, const void* s, size_t n) { memcpy(d, s, n); return 0; }
int zkp_dev_download(zkp_ctx*, void* d, const void* s, size_t n) { memcpy(d, s, n); return 0; }
int zkp_dev_zero(zkp_ctx*, void* d, size_t n) { memset(d, 0, n); return 0; }

int zkp_fr_powers_dev(zkp_ctx*, void* out, const uint64_t base[4], const uint64_t first[4], size_t n) {
  Fr b, cur;
  memcpy(b.v, base, 32);
  memcpy(cur.v, first, 32);
  Fr* o = (Fr*)out;
  for (size_t i = 0; i < n; i++) { o[i] = cur; cur = cur * b; }
  return 0;
}

int zkp_plonk_gate_check_dev(zkp_ctx*, const void* const cols[9], size_t n, int* ok) {
  const Fr *a = (const Fr*)cols[0], *b = (const Fr*)cols[1], *c = (const Fr*)cols[2], *ql = (const Fr*)cols[3],
           *qr = (const Fr*)cols[4], *qo = (const Fr*)cols[5], *qm = (const Fr*)cols[6], *qc = (const Fr*)cols[7],
           *pi = (const Fr*)cols[8];
  *ok = 1;
  for (size_t i = 0; i < n; i++) {
    const Fr v = ql[i] * a[i] + qr[i] * b[i] + qo[i] * c[i] + qm[i] * a[i] * b[i] + qc[i] + pi[i];
    if (!v.is_zero()) { *ok = 0; break; }
  }
  return 0;
}

// ---- never reached by zkp_plonk_compile / zkp_plonk_prove_products / zkp_plonk_prove_reference ----
int zkp_fr_add_at_dev(zkp_ctx*, void*, uint32_t, const size_t*, const uint64_t*) { return ERR_UNSUPPORTED; }
int zkp_fr_batch_inverse_dev(zkp_ctx*, void*, size_t) { return ERR_UNSUPPORTED; }
int zkp_fr_eval_dev(zkp_ctx*, uint32_t, const void* const*, const size_t*, const uint64_t*, uint64_t*) { return ERR_UNSUPPORTED; }
int zkp_fr_lincomb_dev(zkp_ctx*, void*, size_t, uint32_t, const void* const*, const size_t*, const uint64_t*, const uint64_t*) {
  return ERR_UNSUPPORTED;
}
int zkp_fr_mul_pointwise_dev(zkp_ctx*, void*, const void*, size_t) { return ERR_UNSUPPORTED; }
int zkp_fr_scan_dev(zkp_ctx*, void*, size_t, int, int) { return ERR_UNSUPPORTED; }
int zkp_fr_trimmed_len_dev(zkp_ctx*, const void*, size_t, size_t*) { return ERR_UNSUPPORTED; }
int zkp_g1_fold_partials(const uint64_t*, size_t, uint64_t*, uint8_t*) { return ERR_UNSUPPORTED; }
int zkp_msm_g1_multi_partial_dev(zkp_ctx*, uint32_t, const void* const*, const size_t*, uint64_t*) { return ERR_UNSUPPORTED; }
int zkp_plonk_numden_dev(zkp_ctx*, const zkp_plonk_numden_args*) { return ERR_UNSUPPORTED; }
int zkp_plonk_quotient_dev(zkp_ctx*, const zkp_plonk_quotient_args*) { return ERR_UNSUPPORTED; }

}  // extern "C"
