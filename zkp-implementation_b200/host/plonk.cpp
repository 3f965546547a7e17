// Host orchestration of the reference's PLONK prover over the GPU engine (include/zkp_plonk.h).
// Round structure and every formula follow plonk/src/prover.rs:61-293 and its helpers; the circuit
// builder follows plonk/src/circuit.rs and gate.rs.  This file is a pure client of the C ABI
// (include/zkp_b200.h), exactly as the Rust shim would be.
//
//   zkp_plonk_prove            device-resident: polynomials live in HBM from compile to the last opening;
//                              the host only runs the Fiat-Shamir transcript and a few scalar formulas.
//                              The quotient t(X) is computed from coset evaluations (5 transforms of size
//                              4n) instead of the reference's 18 polynomial products (54 transforms at
//                              4n / 8n) -- t is the unique polynomial with t * Z_H = numerator, so its
//                              coefficients, commitments and everything hashed after them are identical.
//   zkp_plonk_prove_products   the same proof computed the way prover.rs is written, one GPU polynomial
//                              product per `&a * &b`, polynomials on the host between calls (cross-check).
#include "../../include/zkp_plonk.h"

#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <new>
#include <vector>

#include "mont_host.hpp"
#include "transcript.hpp"

using namespace zkp_host;
typedef std::vector<Fr> Poly;

namespace {

struct Timers {
  double msm = 0, ntt = 0;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  static double since(std::chrono::steady_clock::time_point t) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count();
  }
};

// ---- ark-poly DensePolynomial semantics ---------------------------------------------------------
void trim(Poly& p) {
  while (!p.empty() && p.back().is_zero()) p.pop_back();
}
Poly add(const Poly& a, const Poly& b) {
  const Poly& lo = a.size() < b.size() ? a : b;
  Poly r = a.size() < b.size() ? b : a;
#pragma omp parallel for schedule(static) if (lo.size() > 4096)
  for (long i = 0; i < (long)lo.size(); i++) r[i] += lo[i];
  trim(r);
  return r;
}
Poly neg(const Poly& a) {
  Poly r(a.size());
#pragma omp parallel for schedule(static) if (a.size() > 4096)
  for (long i = 0; i < (long)a.size(); i++) r[i] = -a[i];
  return r;
}
Poly sub(const Poly& a, const Poly& b) { return add(a, neg(b)); }
Poly scale(const Poly& a, const Fr& k) {
  if (a.empty() || k.is_zero()) return Poly();
  Poly r(a.size());
#pragma omp parallel for schedule(static) if (a.size() > 4096)
  for (long i = 0; i < (long)a.size(); i++) r[i] = a[i] * k;
  return r;
}
// p + c (constant term)
Poly add_const(Poly p, const Fr& c) {
  if (p.empty()) p.push_back(Fr::zero());
  p[0] += c;
  trim(p);
  return p;
}
// p * (X^n - 1)
Poly mul_by_vanishing(const Poly& p, size_t n) {
  if (p.empty()) return Poly();
  Poly r(n + p.size(), Fr::zero());
  for (size_t i = 0; i < p.size(); i++) r[n + i] = p[i];
  for (size_t i = 0; i < p.size(); i++) r[i] -= p[i];
  trim(r);
  return r;
}
// ark-poly divide_by_vanishing_poly; returns false when the remainder is non-zero
bool divide_by_vanishing(const Poly& a, size_t n, Poly& q) {
  if (a.size() < n) {
    q.clear();
    return a.empty();
  }
  q.assign(a.begin() + n, a.end());
  for (size_t i = 1; i < a.size() / n; i++) {
    const size_t off = n * (i + 1);
    for (size_t j = 0; off + j < a.size(); j++) q[j] += a[off + j];
  }
  bool zero = true;
  for (size_t j = 0; j < n; j++) {
    Fr r = a[j];
    if (j < q.size()) r += q[j];
    if (!r.is_zero()) zero = false;
  }
  trim(q);
  return zero;
}
// Horner, chunked so that the host cores share one evaluation
Fr eval(const Poly& p, const Fr& x) {
  const size_t n = p.size();
  if (n == 0) return Fr::zero();
  const size_t chunk = 1 << 14;
  if (n <= chunk) {
    Fr acc = Fr::zero();
    for (size_t i = n; i-- > 0;) acc = acc * x + p[i];
    return acc;
  }
  const size_t nchunks = (n + chunk - 1) / chunk;
  std::vector<Fr> part(nchunks);
#pragma omp parallel for schedule(static)
  for (long c = 0; c < (long)nchunks; c++) {
    const size_t lo = (size_t)c * chunk, hi = std::min(n, lo + chunk);
    Fr acc = Fr::zero();
    for (size_t i = hi; i-- > lo;) acc = acc * x + p[i];
    part[c] = acc;
  }
  const Fr xc = fr_pow(x, chunk);
  Fr acc = Fr::zero();
  for (size_t c = nchunks; c-- > 0;) acc = acc * xc + part[c];
  return acc;
}
// a / (X - root): quotient, remainder
Fr divide_linear(const Poly& a, const Fr& root, Poly& q) {
  if (a.empty()) { q.clear(); return Fr::zero(); }
  q.assign(a.size() - 1, Fr::zero());
  Fr carry = Fr::zero();
  for (size_t i = a.size() - 1; i >= 1; i--) {
    carry = a[i] + carry * root;
    q[i - 1] = carry;
  }
  Fr rem = a[0] + carry * root;
  trim(q);
  return rem;
}

// ---- engine calls -------------------------------------------------------------------------------
int commit(zkp_ctx* ctx, const Poly& p, G1& out, Timers& tm) {
  auto t = std::chrono::steady_clock::now();
  uint8_t inf = 0;
  int st = zkp_msm_g1(ctx, p.empty() ? nullptr : p[0].v, p.size(), out.xy, &inf);
  tm.msm += Timers::since(t);
  return st;
}
int commit_para(zkp_ctx* ctx, const Fr& e, G1& out, Timers& tm) {  // kzg scheme.rs:78-82: e * g1_points[0]
  auto t = std::chrono::steady_clock::now();
  uint8_t inf = 0;
  int st = zkp_msm_g1(ctx, e.v, 1, out.xy, &inf);
  tm.msm += Timers::since(t);
  return st;
}
int mul(zkp_ctx* ctx, const Poly& a, const Poly& b, Poly& out, Timers& tm) {  // &a * &b
  if (a.empty() || b.empty()) { out.clear(); return 0; }
  auto t = std::chrono::steady_clock::now();
  out.assign(a.size() + b.size() - 1, Fr::zero());
  int st = zkp_poly_mul_fr(ctx, a[0].v, a.size(), b[0].v, b.size(), out[0].v);
  tm.ntt += Timers::since(t);
  trim(out);
  return st;
}
int interpolate_batch(zkp_ctx* ctx, std::vector<Fr>& cols, uint32_t log_n, size_t batch, Timers& tm) {
  auto t = std::chrono::steady_clock::now();
  int st = zkp_ntt_fr(ctx, cols[0].v, log_n, batch, 1, nullptr);
  tm.ntt += Timers::since(t);
  return st;
}

}  // namespace

// ---- circuit ------------------------------------------------------------------------------------
struct zkp_plonk_circuit {
  std::vector<uint8_t> kinds;
  std::vector<uint64_t> pos;  // 6 per gate
  std::vector<Fr> vals;       // 3 per gate
  std::vector<Fr> pis;        // 1 per gate (as given; the gate stores -pi, gate.rs:47)
};

struct zkp_plonk_compiled {
  size_t size = 0;
  uint32_t log_n = 0;
  Poly poly[12];               // f_a f_b f_c q_l q_r q_o q_m q_c pi s1 s2 s3 (trimmed coefficients)
  std::vector<Fr> ev_abc[3];   // wire values on the domain (zero on padding rows)
  std::vector<Fr> ev_sigma[3]; // sigma evaluations on the domain
  Fr k1, k2, omega;
  // ---- device-resident form (zkp_plonk_prove) ----
  zkp_ctx* ctx = nullptr;
  uint32_t rho = 4, log_d = 0;  // quotient coset: d = rho * n points x_i = h * omega_d^i, h = 7
  size_t d = 0;
  Fr coset_h;
  bool gate_ok = true;          // gate equation holds on every row (prover.rs:404)
  Fr* d_vals = nullptr;         // 12 x n: the columns above as VALUES on the domain
  Fr* d_coef = nullptr;         // 12 x n: their interpolations, zero padded
  Fr* d_roots = nullptr;        // n: omega^i
  Fr* d_cos = nullptr;          // 11 x d: coset evaluations of q_l q_r q_o q_m q_c pi s1 s2 s3 L1, then the points x_i
  Fr* d_work = nullptr;         // prover workspace (see Work below)
  size_t work_elems = 0;
  std::mutex prove_mu;          // one proof at a time per compiled circuit: the workspace is shared
  // verifier-side preprocessed commitments (q_m q_l q_r q_o q_c s1 s2 s3), cached after the first zkp_plonk_preprocess
  bool pre_valid = false;
  size_t pre_srs_len = 0;
  uint64_t pre_xy[8][12];
  ~zkp_plonk_compiled() {
    if (ctx) {
      zkp_dev_free(ctx, d_vals);
      zkp_dev_free(ctx, d_coef);
      zkp_dev_free(ctx, d_roots);
      zkp_dev_free(ctx, d_cos);
      zkp_dev_free(ctx, d_work);
    }
  }
};

#define PLONK_TRY(e)                \
  do {                              \
    int st_ = (e);                  \
    if (st_ != 0) return st_;       \
  } while (0)

extern "C" {

zkp_plonk_circuit* zkp_plonk_circuit_new(void) { return new (std::nothrow) zkp_plonk_circuit(); }
void zkp_plonk_circuit_free(zkp_plonk_circuit* c) { delete c; }
size_t zkp_plonk_circuit_len(const zkp_plonk_circuit* c) { return c ? c->kinds.size() : 0; }

int zkp_plonk_circuit_add_gates(zkp_plonk_circuit* c, size_t count, const uint8_t* kinds, const uint64_t* positions,
                                const uint64_t* values, const uint64_t* pis) {
  if (!c || (count && (!kinds || !positions || !values || !pis))) return ZKP_B200_ERR_INVALID_ARG;
  for (size_t i = 0; i < count; i++)
    if (kinds[i] > ZKP_PLONK_GATE_CONST) return ZKP_B200_ERR_INVALID_ARG;
  c->kinds.insert(c->kinds.end(), kinds, kinds + count);
  c->pos.insert(c->pos.end(), positions, positions + 6 * count);
  const size_t v0 = c->vals.size(), p0 = c->pis.size();
  c->vals.resize(v0 + 3 * count);
  c->pis.resize(p0 + count);
  memcpy(c->vals[v0].v, values, 3 * count * 32);
  memcpy(c->pis[p0].v, pis, count * 32);
  return 0;
}

int zkp_plonk_compile(zkp_ctx* ctx, const zkp_plonk_circuit* c, zkp_plonk_compiled** out) {
  if (!ctx || !c || !out) return ZKP_B200_ERR_INVALID_ARG;
  *out = nullptr;
  const size_t ln = c->kinds.size();
  if (ln < 2) return ZKP_PLONK_ERR_TOO_FEW_GATES;  // circuit.rs:148-157
  uint32_t log_n = 0;
  while (((size_t)1 << log_n) < ln) log_n++;
  const size_t n = (size_t)1 << log_n;
  zkp_plonk_compiled* cc = new (std::nothrow) zkp_plonk_compiled();
  if (!cc) return ZKP_B200_ERR_OOM;
  cc->size = n;
  cc->log_n = log_n;
  cc->omega = fr_omega(log_n);
  cc->k1 = Fr::from_u64(2);  // find_cosets: roots[0] + 1, + 1 (circuit.rs:238-245)
  cc->k2 = Fr::from_u64(3);
  // columns 0..11 = a b c ql qr qo qm qc pi s1 s2 s3, each n long, dummy rows zero (get_assignment skips them
  // and `interpolate` zero-pads: same vector)
  std::vector<Fr> cols(12 * n, Fr::zero());
  const Fr one = Fr::one(), mone = -Fr::one();
  for (size_t i = 0; i < ln; i++) {
    cols[0 * n + i] = c->vals[3 * i];
    cols[1 * n + i] = c->vals[3 * i + 1];
    cols[2 * n + i] = c->vals[3 * i + 2];
    switch (c->kinds[i]) {
      case ZKP_PLONK_GATE_ADD:   cols[3 * n + i] = one; cols[4 * n + i] = one; cols[5 * n + i] = mone; break;  // gate.rs:38-56
      case ZKP_PLONK_GATE_MUL:   cols[6 * n + i] = one; cols[5 * n + i] = mone; break;                         // gate.rs:58-76
      default:                   cols[3 * n + i] = one; cols[7 * n + i] = -c->vals[3 * i]; break;              // gate.rs:78-97
    }
    cols[8 * n + i] = -c->pis[i];
  }
  // cal_permutation (circuit.rs:200-235)
  std::vector<Fr> roots(n);
  roots[0] = Fr::one();
  for (size_t i = 1; i < n; i++) roots[i] = roots[i - 1] * cc->omega;
  const Fr ks[3] = {Fr::one(), cc->k1, cc->k2};
  for (int col = 0; col < 3; col++)
    for (size_t i = 0; i < n; i++) cols[(9 + col) * n + i] = roots[i] * ks[col];
  for (size_t i = 0; i < ln; i++)
    for (int wire = 0; wire < 3; wire++) {
      const uint64_t pc = c->pos[6 * i + 2 * wire], pr = c->pos[6 * i + 2 * wire + 1];
      if (pc > 2 || pr >= n) { delete cc; return ZKP_PLONK_ERR_INVALID_POSITION; }
      cols[(9 + wire) * n + i] = roots[pr] * ks[pc];
    }
  for (int k = 0; k < 3; k++) {
    cc->ev_abc[k].assign(cols.begin() + k * n, cols.begin() + (k + 1) * n);
    cc->ev_sigma[k].assign(cols.begin() + (9 + k) * n, cols.begin() + (10 + k) * n);
  }
  // ---- device-resident form: values, interpolations (circuit.rs:173-176, 230-232 as one batched iNTT),
  //      domain elements, and the coset evaluations of everything the quotient needs that does not depend
  //      on the blinding / challenges ----
  cc->ctx = ctx;
  cc->rho = (n < 8) ? 8 : 4;  // t has 3n + 6 coefficients: needs d >= 3n + 6
  cc->d = cc->rho * n;
  cc->log_d = log_n + (cc->rho == 8 ? 3 : 2);
  cc->coset_h = Fr::from_u64(7);
  const size_t d = cc->d;
  int st = 0;
  auto fail = [&](int code) { delete cc; return code; };
  if ((st = zkp_dev_alloc(ctx, 12 * n * 32, (void**)&cc->d_vals))) return fail(st);
  if ((st = zkp_dev_alloc(ctx, 12 * n * 32, (void**)&cc->d_coef))) return fail(st);
  if ((st = zkp_dev_alloc(ctx, n * 32, (void**)&cc->d_roots))) return fail(st);
  if ((st = zkp_dev_alloc(ctx, 11 * d * 32, (void**)&cc->d_cos))) return fail(st);
  if ((st = zkp_dev_upload(ctx, cc->d_vals, cols[0].v, 12 * n * 32))) return fail(st);
  if ((st = zkp_dev_copy(ctx, cc->d_coef, cc->d_vals, 12 * n * 32))) return fail(st);
  if ((st = zkp_ntt_fr_dev(ctx, cc->d_coef, log_n, 12, 1, nullptr))) return fail(st);
  if ((st = zkp_dev_download(ctx, cols[0].v, cc->d_coef, 12 * n * 32))) return fail(st);
  for (int k = 0; k < 12; k++) {
    cc->poly[k].assign(cols.begin() + k * n, cols.begin() + (k + 1) * n);
    trim(cc->poly[k]);
  }
  if ((st = zkp_fr_powers_dev(ctx, cc->d_roots, cc->omega.v, one.v, n))) return fail(st);
  {
    const void* chk[9];
    for (int k = 0; k < 9; k++) chk[k] = cc->d_vals + (size_t)k * n;
    int ok = 1;
    if ((st = zkp_plonk_gate_check_dev(ctx, chk, n, &ok))) return fail(st);
    cc->gate_ok = ok != 0;
  }
  // coset cache rows: 0..8 = q_l q_r q_o q_m q_c pi s1 s2 s3 (compiled polys 3..11), 9 = L1, 10 = x_i
  if ((st = zkp_dev_zero(ctx, cc->d_cos, 10 * d * 32))) return fail(st);
  for (int k = 0; k < 9; k++)
    if ((st = zkp_dev_copy(ctx, cc->d_cos + (size_t)k * d, cc->d_coef + (size_t)(3 + k) * n, n * 32))) return fail(st);
  {
    const Fr n_inv = fr_inv(Fr::from_u64(n));  // l1_poly: interpolation of (1, 0, ..., 0) = (1/n) sum X^i (prover.rs:459-464)
    if ((st = zkp_fr_powers_dev(ctx, cc->d_cos + 9 * d, one.v, n_inv.v, n))) return fail(st);
  }
  if ((st = zkp_ntt_fr_dev(ctx, cc->d_cos, cc->log_d, 10, 0, cc->coset_h.v))) return fail(st);
  {
    const Fr eta = fr_omega(cc->log_d);
    if ((st = zkp_fr_powers_dev(ctx, cc->d_cos + 10 * d, eta.v, cc->coset_h.v, d))) return fail(st);
  }
  *out = cc;
  return 0;
}

void zkp_plonk_compiled_free(zkp_plonk_compiled* cc) { delete cc; }

// verifier.rs:160-185 `get_circuit_commitment` / cpi_parser.rs:76-106 `CommonPreprocessedInput::new`: the eight
// commitments every `verify` of the reference recomputes (8 MSMs of size n).  Here: ONE batched pipeline over the
// coefficient vectors already resident in HBM, cached on the compiled circuit.
int zkp_plonk_preprocess(zkp_ctx* ctx, zkp_plonk_compiled* cc, uint64_t out_xy[8][12], int refresh) {
  if (!ctx || !cc || !out_xy) return ZKP_B200_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(cc->prove_mu);
  const size_t srs_len = zkp_srs_len(ctx);
  if (refresh || !cc->pre_valid || cc->pre_srs_len != srs_len) {
    static const int which[8] = {6, 3, 4, 5, 7, 9, 10, 11};  // compiled polynomial indices of q_m q_l q_r q_o q_c s1 s2 s3
    const void* ptrs[8];
    size_t lens[8];
    for (int k = 0; k < 8; k++) {
      ptrs[k] = cc->d_coef + (size_t)which[k] * cc->size;
      lens[k] = cc->poly[which[k]].size();  // DensePolynomial: trimmed coefficient count (scheme.rs:84-96 zips over it)
      if (lens[k] > srs_len) return ZKP_B200_ERR_SRS_TOO_SMALL;
    }
    uint8_t inf[8];
    PLONK_TRY(zkp_msm_g1_multi_dev(ctx, 8, ptrs, lens, &cc->pre_xy[0][0], inf));
    cc->pre_valid = true;
    cc->pre_srs_len = srs_len;
  }
  memcpy(out_xy, cc->pre_xy, sizeof(cc->pre_xy));
  return 0;
}
size_t zkp_plonk_compiled_size(const zkp_plonk_compiled* cc) { return cc ? cc->size : 0; }

int zkp_plonk_compiled_poly(const zkp_plonk_compiled* cc, int which, uint64_t* out) {
  if (!cc || !out || which < 0 || which > 11) return ZKP_B200_ERR_INVALID_ARG;
  memset(out, 0, cc->size * 32);
  if (!cc->poly[which].empty()) memcpy(out, cc->poly[which][0].v, cc->poly[which].size() * 32);
  return 0;
}

static int prove_products_impl(zkp_ctx* ctx, const zkp_plonk_compiled* cc, const uint64_t* blinding, zkp_plonk_proof* out,
                               double* timings_ms, bool literal_acc);

int zkp_plonk_prove_products(zkp_ctx* ctx, const zkp_plonk_compiled* cc, const uint64_t* blinding, zkp_plonk_proof* out,
                             double* timings_ms) {
  return prove_products_impl(ctx, cc, blinding, out, timings_ms, false);
}

int zkp_plonk_prove_reference(zkp_ctx* ctx, const zkp_plonk_compiled* cc, const uint64_t* blinding, zkp_plonk_proof* out,
                              double* timings_ms) {
  return prove_products_impl(ctx, cc, blinding, out, timings_ms, true);
}

static int prove_products_impl(zkp_ctx* ctx, const zkp_plonk_compiled* cc, const uint64_t* blinding, zkp_plonk_proof* out,
                               double* timings_ms, bool literal_acc) {
  if (!ctx || !cc || !blinding || !out) return ZKP_B200_ERR_INVALID_ARG;
  Timers tm;
  const size_t n = cc->size;
  const Fr w = cc->omega;
  Fr b[10];
  for (int i = 1; i <= 9; i++) memcpy(b[i].v, blinding + 4 * (i - 1), 32);
  const Poly &f_a = cc->poly[0], &f_b = cc->poly[1], &f_c = cc->poly[2], &q_l = cc->poly[3], &q_r = cc->poly[4],
             &q_o = cc->poly[5], &q_m = cc->poly[6], &q_c = cc->poly[7], &pi = cc->poly[8], &s1 = cc->poly[9],
             &s2 = cc->poly[10], &s3 = cc->poly[11];
  G1 cm[9];

  // ---- Round 1 (prover.rs:68-92) ----
  auto blind2 = [&](const Fr& hi, const Fr& lo) { Poly p = {lo, hi}; trim(p); return mul_by_vanishing(p, n); };
  const Poly ax = add(f_a, blind2(b[1], b[2]));
  const Poly bx = add(f_b, blind2(b[3], b[4]));
  const Poly cx = add(f_c, blind2(b[5], b[6]));
  PLONK_TRY(commit(ctx, ax, cm[0], tm));
  PLONK_TRY(commit(ctx, bx, cm[1], tm));
  PLONK_TRY(commit(ctx, cx, cm[2], tm));

  // ---- Round 2 (prover.rs:98-123) ----
  ChallengeGenerator ch;
  ch.feed(cm[0]); ch.feed(cm[1]); ch.feed(cm[2]);
  Fr bg[2];
  if (!ch.generate(2, bg)) return ZKP_PLONK_ERR_TRANSCRIPT;
  const Fr beta = bg[0], gamma = bg[1];
  Poly pre4 = {b[9], b[8], b[7]};
  trim(pre4);
  pre4 = mul_by_vanishing(pre4, n);
  Poly pre4w = {b[9], b[8] * w, b[7] * w * w};
  trim(pre4w);
  pre4w = mul_by_vanishing(pre4w, n);
  // compute_acc (prover.rs:302-377) from the domain evaluations kept at compile time
  std::vector<Fr> acc2(2 * n);
  if (literal_acc) {
    // prover.rs:302-377 as written: nine Horner evaluations of degree-n polynomials and one field division per row
    // (O(n^2); the host-CPU baseline of bench.py and a third cross-check of the proof bytes at small n)
    std::vector<Fr> roots(n);
    roots[0] = Fr::one();
    for (size_t i = 1; i < n; i++) roots[i] = roots[i - 1] * w;
    Fr pre_acc = Fr::one();
    acc2[0] = Fr::one();
    for (size_t i = 1; i < n; i++) {
      const Fr& x = roots[i - 1];
      const Fr numerator = (eval(f_a, x) + beta * x + gamma) * (eval(f_b, x) + beta * cc->k1 * x + gamma) *
                           (eval(f_c, x) + beta * cc->k2 * x + gamma);
      const Fr denominator = (eval(f_a, x) + beta * eval(s1, x) + gamma) * (eval(f_b, x) + beta * eval(s2, x) + gamma) *
                             (eval(f_c, x) + beta * eval(s3, x) + gamma);
      pre_acc = pre_acc * numerator * fr_inv(denominator);
      acc2[i] = pre_acc;
    }
    for (size_t i = 0; i < n; i++) acc2[n + i] = acc2[(i + 1) % n];  // rotate_left(1)
  } else {
    std::vector<Fr> num(n), den(n), roots(n);
    roots[0] = Fr::one();
    for (size_t i = 1; i < n; i++) roots[i] = roots[i - 1] * w;
    const Fr bk1 = beta * cc->k1, bk2 = beta * cc->k2;
#pragma omp parallel for schedule(static) if (n > 4096)
    for (long i = 0; i < (long)n; i++) {
      const Fr a = cc->ev_abc[0][i] + gamma, bb = cc->ev_abc[1][i] + gamma, c = cc->ev_abc[2][i] + gamma;
      num[i] = (a + beta * roots[i]) * (bb + bk1 * roots[i]) * (c + bk2 * roots[i]);
      den[i] = (a + beta * cc->ev_sigma[0][i]) * (bb + beta * cc->ev_sigma[1][i]) * (c + beta * cc->ev_sigma[2][i]);
    }
    // batch inversion of den[0..n-2] (Montgomery's trick), then the running product
    std::vector<Fr> pre(n);
    Fr run = Fr::one();
    for (size_t i = 0; i + 1 < n; i++) { pre[i] = run; run = run * den[i]; }
    Fr inv = fr_inv(run);  // a zero denominator makes the reference's `/` panic; here it yields garbage -> remainder error
    for (size_t i = n - 1; i-- > 0;) { const Fr di = inv * pre[i]; inv = inv * den[i]; den[i] = di; }
    acc2[0] = Fr::one();
    for (size_t i = 1; i < n; i++) acc2[i] = acc2[i - 1] * num[i - 1] * den[i - 1];
    for (size_t i = 0; i < n; i++) acc2[n + i] = acc2[(i + 1) % n];  // rotate_left(1)
  }
  PLONK_TRY(interpolate_batch(ctx, acc2, cc->log_n, 2, tm));  // prover.rs:374-375
  Poly acc_x(acc2.begin(), acc2.begin() + n), acc_wx(acc2.begin() + n, acc2.end());
  trim(acc_x);
  trim(acc_wx);
  const Poly z_x = add(pre4, acc_x);
  const Poly z_wx = add(pre4w, acc_wx);
  PLONK_TRY(commit(ctx, z_x, cm[3], tm));

  // ---- Round 3 (prover.rs:136-150, 381-444) ----
  ch.feed(cm[3]);
  Fr alpha;
  if (!ch.generate(1, &alpha)) return ZKP_PLONK_ERR_TRANSCRIPT;
  Poly t1, t2, t3, quotient1, quotient23, quotient4;
  {
    PLONK_TRY(mul(ctx, ax, bx, t1, tm));
    PLONK_TRY(mul(ctx, t1, q_m, t2, tm));
    Poly line1 = t2;
    PLONK_TRY(mul(ctx, ax, q_l, t1, tm)); line1 = add(line1, t1);
    PLONK_TRY(mul(ctx, bx, q_r, t1, tm)); line1 = add(line1, t1);
    PLONK_TRY(mul(ctx, cx, q_o, t1, tm)); line1 = add(line1, t1);
    line1 = add(add(line1, pi), q_c);
    if (!divide_by_vanishing(line1, n, quotient1)) return ZKP_PLONK_ERR_REMAINDER;  // "No remainder 1"
  }
  {
    auto lin = [&](const Poly& p, const Fr& c0, const Fr& c1) { Poly l = {c0, c1}; trim(l); return add(p, l); };
    PLONK_TRY(mul(ctx, lin(ax, gamma, beta), lin(bx, gamma, beta * cc->k1), t1, tm));
    PLONK_TRY(mul(ctx, t1, lin(cx, gamma, beta * cc->k2), t2, tm));
    PLONK_TRY(mul(ctx, t2, z_x, t1, tm));
    const Poly line2 = scale(t1, alpha);
    auto perm = [&](const Poly& p, const Poly& s) { return add_const(add(p, scale(s, beta)), gamma); };
    PLONK_TRY(mul(ctx, perm(ax, s1), perm(bx, s2), t1, tm));
    PLONK_TRY(mul(ctx, t1, perm(cx, s3), t2, tm));
    PLONK_TRY(mul(ctx, t2, z_wx, t3, tm));
    const Poly line3 = scale(t3, alpha);
    if (!divide_by_vanishing(sub(line2, line3), n, quotient23)) return ZKP_PLONK_ERR_REMAINDER;
  }
  // l1_poly: interpolation of (1, 0, ..., 0) = (1/n) * sum X^i  (prover.rs:459-464)
  Poly l1(n, fr_inv(Fr::from_u64(n)));
  Poly zx2 = z_x;
  if (zx2.empty()) zx2.push_back(Fr::zero());
  zx2[0] -= Fr::one();
  {
    Poly zx2t = zx2;
    trim(zx2t);
    PLONK_TRY(mul(ctx, zx2t, l1, t1, tm));
    if (!divide_by_vanishing(scale(t1, alpha * alpha), n, quotient4)) return ZKP_PLONK_ERR_REMAINDER;
  }
  const Poly tx = add(add(quotient1, quotient23), quotient4);
  // SlicePoly::new (slice_polynomial.rs:22-43)
  size_t tmp = tx.size() / 3;
  if (tmp * 3 < tx.size()) tmp++;
  Poly slices[3];
  for (int i = 0; i < 3 && tmp; i++) {
    const size_t lo = std::min(tx.size(), (size_t)i * tmp), hi = std::min(tx.size(), (size_t)(i + 1) * tmp);
    slices[i].assign(tx.begin() + lo, tx.begin() + hi);
    trim(slices[i]);
  }
  const uint64_t degree = (uint64_t)tmp - 1;  // wraps like the reference's usize arithmetic would panic; tmp >= 1 for valid circuits
  for (int i = 0; i < 3; i++) PLONK_TRY(commit(ctx, slices[i], cm[4 + i], tm));

  // ---- Round 4 (prover.rs:156-178) ----
  ch.feed(cm[4]); ch.feed(cm[5]); ch.feed(cm[6]);
  Fr zeta;
  if (!ch.generate(1, &zeta)) return ZKP_PLONK_ERR_TRANSCRIPT;
  const Fr bar_a = eval(ax, zeta), bar_b = eval(bx, zeta), bar_c = eval(cx, zeta);
  const Fr bar_s1 = eval(s1, zeta), bar_s2 = eval(s2, zeta);
  const Fr bar_z_w = eval(z_x, zeta * w);
  const Fr pi_e = eval(pi, zeta);
  Poly tx_compact;
  for (int i = 0; i < 3; i++) tx_compact = add(tx_compact, scale(slices[i], fr_pow(zeta, (degree + 1) * (uint64_t)i)));

  // ---- Round 5 (prover.rs:183-272) ----
  const Fr bars[6] = {bar_a, bar_b, bar_c, bar_s1, bar_s2, bar_z_w};
  for (int i = 0; i < 6; i++) {
    G1 e;
    PLONK_TRY(commit_para(ctx, bars[i], e, tm));
    ch.feed(e);
  }
  Fr v;
  if (!ch.generate(1, &v)) return ZKP_PLONK_ERR_TRANSCRIPT;
  // compute_linearisation_polynomial (prover.rs:469-568; the self-check products are not recomputed)
  Poly line1 = add(add(add(add(scale(q_m, bar_a * bar_b), scale(q_l, bar_a)), scale(q_r, bar_b)), scale(q_o, bar_c)), q_c);
  line1 = add_const(line1, pi_e);
  const Fr sc2 = (bar_a + beta * zeta + gamma) * (bar_b + beta * cc->k1 * zeta + gamma) * (bar_c + beta * cc->k2 * zeta + gamma) * alpha;
  const Poly line2 = scale(z_x, sc2);
  const Fr sc3 = (bar_a + beta * bar_s1 + gamma) * (bar_b + beta * bar_s2 + gamma) * bar_z_w * alpha;
  const Poly line3 = scale(add_const(scale(s3, beta), bar_c + gamma), sc3);
  const Fr z_h_e = fr_pow(zeta, n) - Fr::one();
  const Fr l1_e = eval(l1, zeta);
  Poly zx2t = zx2;
  trim(zx2t);
  const Poly line4 = scale(zx2t, l1_e * alpha * alpha);
  const Poly line5 = scale(tx_compact, z_h_e);
  const Poly r_x = add(add(add(add(line1, line2), neg(line3)), line4), neg(line5));
  const Fr bar_r = eval(r_x, zeta);
  auto sub_para = [&](const Poly& p, const Fr& k) { return add_const(p, -k); };
  const Fr v2 = v * v, v3 = v2 * v, v4 = v3 * v, v5 = v4 * v;
  Poly wx = sub_para(r_x, bar_r);
  wx = add(wx, scale(sub_para(ax, bar_a), v));
  wx = add(wx, scale(sub_para(bx, bar_b), v2));
  wx = add(wx, scale(sub_para(cx, bar_c), v3));
  wx = add(wx, scale(sub_para(s1, bar_s1), v4));
  wx = add(wx, scale(sub_para(s2, bar_s2), v5));
  Poly w_ev_x, w_ev_wx;
  if (!divide_linear(wx, zeta, w_ev_x).is_zero()) return ZKP_PLONK_ERR_REMAINDER;            // prover.rs:228-240
  if (!divide_linear(sub_para(z_x, bar_z_w), zeta * w, w_ev_wx).is_zero()) return ZKP_PLONK_ERR_REMAINDER;  // :248-260
  PLONK_TRY(commit(ctx, w_ev_x, cm[7], tm));
  PLONK_TRY(commit(ctx, w_ev_wx, cm[8], tm));
  ch.feed(cm[7]); ch.feed(cm[8]);
  Fr u;
  if (!ch.generate(1, &u)) return ZKP_PLONK_ERR_TRANSCRIPT;

  for (int i = 0; i < 9; i++) memcpy(out->commitments[i], cm[i].xy, 96);
  for (int i = 0; i < 6; i++) memcpy(out->evaluations[i], bars[i].v, 32);
  memcpy(out->u, u.v, 32);
  out->degree = degree;
  if (timings_ms) {
    timings_ms[0] = Timers::since(tm.t0);
    timings_ms[1] = tm.msm;
    timings_ms[2] = tm.ntt;
    timings_ms[3] = timings_ms[0] - tm.msm - tm.ntt;
  }
  return 0;
}

// ---- device-resident prover -------------------------------------------------------------------------
namespace {

struct DevTimers {
  double msm = 0, ntt = 0;
  bool precise = false;  // the caller asked for a breakdown: synchronise at the phase boundaries
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
};

// Point-range sharding of the commitments over several GPUs (one process each): the SRS resident in this rank's
// context is [first, first + zkp_srs_len) of the global SRS; every rank runs the whole prover (the transforms and
// pointwise work are a small part of a proof) but only its range of each commitment, and the 192-byte partial sums
// are exchanged through `allgather` (NCCL in production, gloo in the CPU tests) and folded.
struct Shard {
  uint32_t rank = 0, world = 1;
  size_t first = 0, total = 0;  // this rank's first global SRS index; global SRS length
  zkp_allgather_fn allgather = nullptr;
  void* user = nullptr;
  bool on() const { return world > 1; }
};

// Workspace layout inside cc->d_work (element offsets); L = n + 4 slots per coefficient vector.
struct Work {
  size_t n, d, L;
  Fr* base;
  Fr* coef(int k) const { return base + (size_t)k * L; }           // 0..3 = a, b, c, z coefficients
  Fr* r() const { return base + 4 * L; }                           // linearisation polynomial
  Fr* wx() const { return base + 5 * L; }                          // opening numerator / quotient at zeta
  Fr* wwx() const { return base + 6 * L; }                         // ... at zeta * omega
  Fr* pw() const { return base + 7 * L; }                          // powers of the opening point
  Fr* pwi() const { return base + 8 * L; }                         // powers of its inverse
  Fr* acc() const { return base + 9 * L; }                         // n + 1: grand product
  Fr* den() const { return base + 10 * L; }                        // n
  Fr* cos(int k) const { return base + 11 * L + (size_t)k * d; }   // 0..3 = a, b, c, z on the coset
  Fr* t() const { return base + 11 * L + 4 * d; }                  // d: quotient evaluations -> coefficients
  static size_t elems(size_t n, size_t d) { return 11 * (n + 4) + 5 * d; }
};

// several commitments against the resident SRS as one MSM pipeline (zkp_msm_g1_multi_dev); sharded: this rank's
// range of every polynomial, all-gather of the partial sums, fold
int dev_commit_multi(zkp_ctx* ctx, const Shard& sh, uint32_t count, const Fr* const* scalars_dev, const size_t* lens, G1* out,
                     DevTimers& tm) {
  if (tm.precise) zkp_ctx_synchronize(ctx);  // queued transform / pointwise work is not charged to the commitments
  auto t = std::chrono::steady_clock::now();
  int st = 0;
  if (!sh.on()) {
    st = zkp_msm_g1_multi_dev(ctx, count, (const void* const*)scalars_dev, lens, out[0].xy, nullptr);
  } else {
    const size_t mine = zkp_srs_len(ctx);
    const void* ptrs[16];
    size_t part_len[16];
    for (uint32_t j = 0; j < count; j++) {
      const size_t hi = std::min(lens[j], sh.first + mine);
      part_len[j] = hi > sh.first ? hi - sh.first : 0;
      ptrs[j] = scalars_dev[j] + sh.first;
    }
    std::vector<uint64_t> part((size_t)count * 24), all((size_t)count * 24 * sh.world);
    st = zkp_msm_g1_multi_partial_dev(ctx, count, ptrs, part_len, part.data());
    if (!st) st = sh.allgather(sh.user, part.data(), part.size() * 8, all.data());
    std::vector<uint64_t> mine_of(24 * (size_t)sh.world);
    for (uint32_t j = 0; j < count && !st; j++) {
      for (uint32_t g = 0; g < sh.world; g++)
        memcpy(&mine_of[24 * g], &all[((size_t)g * count + j) * 24], 192);
      uint8_t inf = 0;
      st = zkp_g1_fold_partials(mine_of.data(), sh.world, out[j].xy, &inf);
    }
  }
  tm.msm += Timers::since(t);
  return st;
}
int dev_commit(zkp_ctx* ctx, const Shard& sh, const Fr* scalars_dev, size_t len, G1& out, DevTimers& tm) {
  return dev_commit_multi(ctx, sh, 1, &scalars_dev, &len, &out, tm);
}

// (poly - poly(root)) / (X - root) in place on the device: weight by root^i, suffix sums, unweight.
// On return the quotient (len - 1 coefficients) starts at data + 1; *rem_zero tells whether the remainder
// vanished (prover.rs:228-240 / 248-260 assert it).
int dev_divide_linear(zkp_ctx* ctx, const Work& w, Fr* data, size_t len, const Fr& root, bool* rem_zero) {
  if (root.is_zero()) return ZKP_B200_ERR_INVALID_ARG;  // probability 2^-255 for a Fiat-Shamir challenge
  const Fr one = Fr::one(), rinv = fr_inv(root);
  PLONK_TRY(zkp_fr_powers_dev(ctx, w.pw(), root.v, one.v, len));
  PLONK_TRY(zkp_fr_powers_dev(ctx, w.pwi(), rinv.v, one.v, len));
  PLONK_TRY(zkp_fr_mul_pointwise_dev(ctx, data, w.pw(), len));
  PLONK_TRY(zkp_fr_scan_dev(ctx, data, len, 1, 1));
  Fr rem;
  PLONK_TRY(zkp_dev_download(ctx, rem.v, data, 32));
  *rem_zero = rem.is_zero();
  return zkp_fr_mul_pointwise_dev(ctx, data, w.pwi(), len);
}

}  // namespace

static int prove_device(zkp_ctx* ctx, const zkp_plonk_compiled* cc_in, const uint64_t* blinding, zkp_plonk_proof* out,
                        double* timings_ms, const Shard& sh);

int zkp_plonk_prove(zkp_ctx* ctx, const zkp_plonk_compiled* cc_in, const uint64_t* blinding, zkp_plonk_proof* out,
                    double* timings_ms) {
  return prove_device(ctx, cc_in, blinding, out, timings_ms, Shard());
}

int zkp_plonk_prove_sharded(zkp_ctx* ctx, const zkp_plonk_compiled* cc, const uint64_t* blinding, zkp_plonk_proof* out,
                            double* timings_ms, uint32_t rank, uint32_t world, size_t srs_first, size_t srs_total,
                            zkp_allgather_fn allgather, void* user) {
  if (world == 0 || rank >= world || (world > 1 && !allgather)) return ZKP_B200_ERR_INVALID_ARG;
  Shard sh;
  sh.rank = rank;
  sh.world = world;
  sh.first = srs_first;
  sh.total = srs_total;
  sh.allgather = allgather;
  sh.user = user;
  return prove_device(ctx, cc, blinding, out, timings_ms, sh);
}

static int prove_device(zkp_ctx* ctx, const zkp_plonk_compiled* cc_in, const uint64_t* blinding, zkp_plonk_proof* out,
                        double* timings_ms, const Shard& sh) {
  if (!ctx || !cc_in || !blinding || !out || cc_in->ctx != ctx) return ZKP_B200_ERR_INVALID_ARG;
  zkp_plonk_compiled* cc = const_cast<zkp_plonk_compiled*>(cc_in);  // the workspace is a cache, not circuit state
  std::lock_guard<std::mutex> prove_lock(cc->prove_mu);
  DevTimers tm;
  tm.precise = timings_ms != nullptr;
  const size_t n = cc->size, d = cc->d;
  const uint32_t log_n = cc->log_n;
  const Fr w = cc->omega, one = Fr::one();
  if ((sh.on() ? sh.total : zkp_srs_len(ctx)) < n + 3) return ZKP_B200_ERR_SRS_TOO_SMALL;  // scheme.rs:86 for the degree n + 2 commitment
  if (sh.on() && sh.first + zkp_srs_len(ctx) > sh.total) return ZKP_B200_ERR_INVALID_ARG;
  if (!cc->d_work) {
    cc->work_elems = Work::elems(n, d);
    PLONK_TRY(zkp_dev_alloc(ctx, cc->work_elems * 32, (void**)&cc->d_work));
  }
  Work wk{n, d, n + 4, cc->d_work};
  Fr b[10];
  for (int i = 1; i <= 9; i++) memcpy(b[i].v, blinding + 4 * (i - 1), 32);
  G1 cm[9];
  auto coef_of = [&](int which) { return cc->d_coef + (size_t)which * n; };  // compiled polynomial, n coefficients

  // ---- Round 1 (prover.rs:68-92): a = f_a + (b1 X + b2)(X^n - 1), ... ----
  PLONK_TRY(zkp_dev_zero(ctx, wk.base, 4 * wk.L * 32));
  for (int k = 0; k < 3; k++) {
    PLONK_TRY(zkp_dev_copy(ctx, wk.coef(k), coef_of(k), n * 32));
    const Fr hi = b[2 * k + 1], lo = b[2 * k + 2];
    const size_t idx[4] = {0, 1, n, n + 1};
    const Fr vals[4] = {-lo, -hi, lo, hi};
    PLONK_TRY(zkp_fr_add_at_dev(ctx, wk.coef(k), 4, idx, vals[0].v));
  }
  {
    const Fr* polys[3] = {wk.coef(0), wk.coef(1), wk.coef(2)};
    const size_t lens[3] = {n + 2, n + 2, n + 2};
    PLONK_TRY(dev_commit_multi(ctx, sh, 3, polys, lens, &cm[0], tm));  // commit_round1 (prover.rs:571-581)
  }

  // ---- Round 2 (prover.rs:98-123, 302-377) ----
  ChallengeGenerator ch;
  ch.feed(cm[0]); ch.feed(cm[1]); ch.feed(cm[2]);
  Fr bg[2];
  if (!ch.generate(2, bg)) return ZKP_PLONK_ERR_TRANSCRIPT;
  const Fr beta = bg[0], gamma = bg[1];
  bool grand_product_ok = true;
  {
    zkp_plonk_numden_args a;
    memset(&a, 0, sizeof(a));
    a.a_dev = cc->d_vals; a.b_dev = cc->d_vals + n; a.c_dev = cc->d_vals + 2 * n;
    a.s1_dev = cc->d_vals + 9 * n; a.s2_dev = cc->d_vals + 10 * n; a.s3_dev = cc->d_vals + 11 * n;
    a.roots_dev = cc->d_roots;
    memcpy(a.beta, beta.v, 32); memcpy(a.gamma, gamma.v, 32); memcpy(a.k1, cc->k1.v, 32); memcpy(a.k2, cc->k2.v, 32);
    a.n = n;
    a.num_dev = wk.acc() + 1;  // acc[0] = 1, acc[i + 1] = acc[i] * num[i] / den[i]
    a.den_dev = wk.den();
    PLONK_TRY(zkp_plonk_numden_dev(ctx, &a));
    PLONK_TRY(zkp_fr_batch_inverse_dev(ctx, wk.den(), n));
    PLONK_TRY(zkp_fr_mul_pointwise_dev(ctx, wk.acc() + 1, wk.den(), n));
    PLONK_TRY(zkp_dev_upload(ctx, wk.acc(), one.v, 32));
    PLONK_TRY(zkp_fr_scan_dev(ctx, wk.acc(), n + 1, 0, 0));
    // acc[n] is the product over the whole domain: it is 1 exactly when the copy constraints hold, which is
    // what makes line2 - line3 divisible by Z_H (prover.rs:431 expect("No remainder here"))
    Fr total;
    PLONK_TRY(zkp_dev_download(ctx, total.v, wk.acc() + n, 32));
    grand_product_ok = (total == one);
    auto t = std::chrono::steady_clock::now();
    if (tm.precise) { zkp_ctx_synchronize(ctx); t = std::chrono::steady_clock::now(); }
    PLONK_TRY(zkp_ntt_fr_dev(ctx, wk.acc(), log_n, 1, 1, nullptr));  // prover.rs:374: acc evaluations -> coefficients
    if (tm.precise) zkp_ctx_synchronize(ctx);
    tm.ntt += Timers::since(t);
    PLONK_TRY(zkp_dev_copy(ctx, wk.coef(3), wk.acc(), n * 32));
    const size_t idx[6] = {0, 1, 2, n, n + 1, n + 2};  // + (b7 X^2 + b8 X + b9)(X^n - 1)
    const Fr vals[6] = {-b[9], -b[8], -b[7], b[9], b[8], b[7]};
    PLONK_TRY(zkp_fr_add_at_dev(ctx, wk.coef(3), 6, idx, vals[0].v));
  }
  PLONK_TRY(dev_commit(ctx, sh, wk.coef(3), n + 3, cm[3], tm));

  // ---- Round 3 (prover.rs:136-150, 381-444) ----
  ch.feed(cm[3]);
  Fr alpha;
  if (!ch.generate(1, &alpha)) return ZKP_PLONK_ERR_TRANSCRIPT;
  if (!cc->gate_ok) return ZKP_PLONK_ERR_REMAINDER;       // "No remainder 1"
  if (!grand_product_ok) return ZKP_PLONK_ERR_REMAINDER;  // "No remainder here" (line 2 - line 3)
  {
    PLONK_TRY(zkp_dev_zero(ctx, wk.cos(0), 4 * d * 32));
    for (int k = 0; k < 4; k++) PLONK_TRY(zkp_dev_copy(ctx, wk.cos(k), wk.coef(k), (n + 3) * 32));
    if (tm.precise) zkp_ctx_synchronize(ctx);
    auto t = std::chrono::steady_clock::now();
    PLONK_TRY(zkp_ntt_fr_dev(ctx, wk.cos(0), cc->log_d, 4, 0, cc->coset_h.v));
    if (tm.precise) zkp_ctx_synchronize(ctx);
    tm.ntt += Timers::since(t);
    zkp_plonk_quotient_args q;
    memset(&q, 0, sizeof(q));
    q.a_dev = wk.cos(0); q.b_dev = wk.cos(1); q.c_dev = wk.cos(2); q.z_dev = wk.cos(3);
    const Fr* cs = cc->d_cos;
    q.ql_dev = cs; q.qr_dev = cs + d; q.qo_dev = cs + 2 * d; q.qm_dev = cs + 3 * d; q.qc_dev = cs + 4 * d;
    q.pi_dev = cs + 5 * d; q.s1_dev = cs + 6 * d; q.s2_dev = cs + 7 * d; q.s3_dev = cs + 8 * d;
    q.l1_dev = cs + 9 * d; q.x_dev = cs + 10 * d;
    memcpy(q.beta, beta.v, 32); memcpy(q.gamma, gamma.v, 32); memcpy(q.alpha, alpha.v, 32);
    memcpy(q.k1, cc->k1.v, 32); memcpy(q.k2, cc->k2.v, 32);
    // Z_H(x_i) = h^n (omega_d^n)^i - 1 with omega_d^n a primitive rho-th root of unity
    const Fr hn = fr_pow(cc->coset_h, n), wr = fr_omega(cc->rho == 8 ? 3 : 2);
    Fr cur = hn;
    for (uint32_t i = 0; i < cc->rho; i++) {
      const Fr zi = fr_inv(cur - one);
      memcpy(q.zh_inv[i], zi.v, 32);
      cur = cur * wr;
    }
    q.d = d;
    q.rho = cc->rho;
    q.t_dev = wk.t();
    PLONK_TRY(zkp_plonk_quotient_dev(ctx, &q));
    if (tm.precise) zkp_ctx_synchronize(ctx);
    t = std::chrono::steady_clock::now();
    PLONK_TRY(zkp_ntt_fr_dev(ctx, wk.t(), cc->log_d, 1, 1, cc->coset_h.v));
    if (tm.precise) zkp_ctx_synchronize(ctx);
    tm.ntt += Timers::since(t);
  }
  // SlicePoly::new (slice_polynomial.rs:22-43) on the trimmed quotient
  size_t t_len = 0;
  PLONK_TRY(zkp_fr_trimmed_len_dev(ctx, wk.t(), d, &t_len));
  size_t tmp = t_len / 3;
  if (tmp * 3 < t_len) tmp++;
  if (tmp == 0) return ZKP_B200_ERR_EMPTY_POLY;  // `chunks(0)` panics in the reference
  size_t slice_len[3];
  {
    const Fr* polys[3];
    for (int i = 0; i < 3; i++) {
      const size_t lo = std::min(t_len, (size_t)i * tmp), hi = std::min(t_len, (size_t)(i + 1) * tmp);
      slice_len[i] = hi - lo;
      polys[i] = wk.t() + lo;
    }
    PLONK_TRY(dev_commit_multi(ctx, sh, 3, polys, slice_len, &cm[4], tm));  // SlicePoly::commit (slice_polynomial.rs:51-53)
  }
  const uint64_t degree = (uint64_t)tmp - 1;

  // ---- Round 4 (prover.rs:156-178) ----
  ch.feed(cm[4]); ch.feed(cm[5]); ch.feed(cm[6]);
  Fr zeta;
  if (!ch.generate(1, &zeta)) return ZKP_PLONK_ERR_TRANSCRIPT;
  const Fr zeta_w = zeta * w;
  Fr ev[7];
  {
    const void* polys[7] = {wk.coef(0), wk.coef(1), wk.coef(2), coef_of(9), coef_of(10), wk.coef(3), coef_of(8)};
    const size_t lens[7] = {n + 2, n + 2, n + 2, n, n, n + 3, n};
    Fr xs[7] = {zeta, zeta, zeta, zeta, zeta, zeta_w, zeta};
    PLONK_TRY(zkp_fr_eval_dev(ctx, 7, polys, lens, xs[0].v, ev[0].v));
  }
  const Fr bar_a = ev[0], bar_b = ev[1], bar_c = ev[2], bar_s1 = ev[3], bar_s2 = ev[4], bar_z_w = ev[5], pi_e = ev[6];

  // ---- Round 5 (prover.rs:183-272) ----
  const Fr bars[6] = {bar_a, bar_b, bar_c, bar_s1, bar_s2, bar_z_w};
  {
    G1 para[6];
    auto t = std::chrono::steady_clock::now();
    if (!sh.on()) {
      PLONK_TRY(zkp_g1_mul_srs0(ctx, bars[0].v, 6, para[0].xy));  // scheme.commit_para x 6
    } else {  // g1_points[0] lives on the rank whose shard starts at 0; its six points are what everybody hashes
      std::vector<uint64_t> mine(6 * 12, 0), all((size_t)6 * 12 * sh.world);
      if (sh.first == 0) PLONK_TRY(zkp_g1_mul_srs0(ctx, bars[0].v, 6, mine.data()));
      std::vector<uint64_t> owner_flag_and_pts(1 + 6 * 12, 0);
      owner_flag_and_pts[0] = sh.first == 0 ? 1 : 0;
      memcpy(&owner_flag_and_pts[1], mine.data(), 6 * 96);
      std::vector<uint64_t> gathered((size_t)(1 + 6 * 12) * sh.world);
      PLONK_TRY(sh.allgather(sh.user, owner_flag_and_pts.data(), owner_flag_and_pts.size() * 8, gathered.data()));
      bool found = false;
      for (uint32_t g = 0; g < sh.world && !found; g++)
        if (gathered[(size_t)g * (1 + 6 * 12)] == 1) {
          memcpy(para[0].xy, &gathered[(size_t)g * (1 + 6 * 12) + 1], 6 * 96);
          found = true;
        }
      if (!found) return ZKP_B200_ERR_INVALID_ARG;
    }
    tm.msm += Timers::since(t);
    for (int i = 0; i < 6; i++) ch.feed(para[i]);
  }
  Fr v;
  if (!ch.generate(1, &v)) return ZKP_PLONK_ERR_TRANSCRIPT;
  // compute_linearisation_polynomial (prover.rs:469-568) as one linear combination of resident vectors
  const Fr alpha2 = alpha * alpha;
  const Fr sc2 = (bar_a + beta * zeta + gamma) * (bar_b + beta * cc->k1 * zeta + gamma) * (bar_c + beta * cc->k2 * zeta + gamma) * alpha;
  const Fr sc3 = (bar_a + beta * bar_s1 + gamma) * (bar_b + beta * bar_s2 + gamma) * bar_z_w * alpha;
  const Fr z_h_e = fr_pow(zeta, n) - one;
  // l1_poly(domain).evaluate(zeta) = (1/n) sum zeta^i = (zeta^n - 1) / (n (zeta - 1))
  const Fr l1_e = (zeta == one) ? one : z_h_e * fr_inv(Fr::from_u64(n) * (zeta - one));
  const Fr zm = fr_pow(zeta, degree + 1);
  const size_t r_len = std::max(n + 3, tmp);
  {
    const void* polys[10] = {coef_of(6), coef_of(3), coef_of(4), coef_of(5), coef_of(7), wk.coef(3), coef_of(11),
                             wk.t(), wk.t() + std::min(t_len, tmp), wk.t() + std::min(t_len, 2 * tmp)};
    const size_t lens[10] = {n, n, n, n, n, n + 3, n, slice_len[0], slice_len[1], slice_len[2]};
    const Fr coefs[10] = {bar_a * bar_b, bar_a, bar_b, bar_c, one, sc2 + l1_e * alpha2, -(beta * sc3),
                          -z_h_e, -(z_h_e * zm), -(z_h_e * zm * zm)};
    const Fr c0 = pi_e - (bar_c + gamma) * sc3 - l1_e * alpha2;
    PLONK_TRY(zkp_fr_lincomb_dev(ctx, wk.r(), r_len, 10, polys, lens, coefs[0].v, c0.v));
  }
  Fr bar_r;
  {
    const void* polys[1] = {wk.r()};
    const size_t lens[1] = {r_len};
    PLONK_TRY(zkp_fr_eval_dev(ctx, 1, polys, lens, zeta.v, bar_r.v));
  }
  const Fr v2 = v * v, v3 = v2 * v, v4 = v3 * v, v5 = v4 * v;
  bool rem_ok = false;
  {
    const void* polys[6] = {wk.r(), wk.coef(0), wk.coef(1), wk.coef(2), coef_of(9), coef_of(10)};
    const size_t lens[6] = {r_len, n + 2, n + 2, n + 2, n, n};
    const Fr coefs[6] = {one, v, v2, v3, v4, v5};
    const Fr c0 = -(bar_r + v * bar_a + v2 * bar_b + v3 * bar_c + v4 * bar_s1 + v5 * bar_s2);
    PLONK_TRY(zkp_fr_lincomb_dev(ctx, wk.wx(), r_len, 6, polys, lens, coefs[0].v, c0.v));
    PLONK_TRY(dev_divide_linear(ctx, wk, wk.wx(), r_len, zeta, &rem_ok));
    if (!rem_ok) return ZKP_PLONK_ERR_REMAINDER;  // "w_ev_x was computed incorrectly"
  }
  {
    PLONK_TRY(zkp_dev_copy(ctx, wk.wwx(), wk.coef(3), (n + 3) * 32));
    const size_t idx[1] = {0};
    const Fr vals[1] = {-bar_z_w};
    PLONK_TRY(zkp_fr_add_at_dev(ctx, wk.wwx(), 1, idx, vals[0].v));
    PLONK_TRY(dev_divide_linear(ctx, wk, wk.wwx(), n + 3, zeta_w, &rem_ok));
    if (!rem_ok) return ZKP_PLONK_ERR_REMAINDER;  // "w_ev_wx was computed incorrectly"
    const Fr* polys[2] = {wk.wx() + 1, wk.wwx() + 1};
    const size_t lens[2] = {r_len - 1, n + 2};
    PLONK_TRY(dev_commit_multi(ctx, sh, 2, polys, lens, &cm[7], tm));
  }
  ch.feed(cm[7]); ch.feed(cm[8]);
  Fr u;
  if (!ch.generate(1, &u)) return ZKP_PLONK_ERR_TRANSCRIPT;

  for (int i = 0; i < 9; i++) memcpy(out->commitments[i], cm[i].xy, 96);
  for (int i = 0; i < 6; i++) memcpy(out->evaluations[i], bars[i].v, 32);
  memcpy(out->u, u.v, 32);
  out->degree = degree;
  if (timings_ms) {
    timings_ms[0] = Timers::since(tm.t0);
    timings_ms[1] = tm.msm;
    timings_ms[2] = tm.ntt;
    timings_ms[3] = timings_ms[0] - tm.msm - tm.ntt;
  }
  return 0;
}

}  // extern "C"
