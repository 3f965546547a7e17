// Argument blocks and host entry points of the device-resident polynomial layer (poly.cu).
#pragma once
#include "engine.h"

namespace zkp {

struct LincombArgs {
  static constexpr uint32_t MAX_TERMS = 12;
  const Fr* p[MAX_TERMS];
  size_t len[MAX_TERMS];
  Fr coef[MAX_TERMS];
  uint32_t count;
  uint32_t has_c0;
  Fr c0;  // added to out[0]
};

struct SparseAddArgs {
  static constexpr uint32_t MAX_TERMS = 8;
  size_t idx[MAX_TERMS];
  Fr val[MAX_TERMS];
  uint32_t count;
};

struct PlonkNumDenArgs {
  const Fr *a, *b, *c, *s1, *s2, *s3, *roots;
  Fr beta, gamma, beta_k1, beta_k2;
  size_t n;
  Fr *num, *den;
};

struct PlonkQuotientArgs {
  const Fr *a, *b, *c, *z;                             // coset evaluations, d each
  const Fr *ql, *qr, *qo, *qm, *qc, *pi, *s1, *s2, *s3, *l1, *x;  // cached coset evaluations + the coset points
  Fr beta, gamma, alpha, alpha2, beta_k1, beta_k2;
  Fr zh_inv[8];                                        // 1 / Z_H(x_i), periodic in i with period rho
  size_t d;                                            // coset size (power of two)
  uint32_t rho;                                        // d / n
  Fr* t;
};

int fr_powers_dev(Ctx* ctx, Fr* out, const Fr& base, const Fr& first, size_t n);
int fr_batch_inverse_dev(Ctx* ctx, Fr* data, size_t n);
int fr_scan_dev(Ctx* ctx, Fr* data, size_t n, int op /* 0 mul, 1 add */, bool reverse);
int fr_lincomb_dev(Ctx* ctx, Fr* out, size_t out_len, const LincombArgs& a);
int fr_add_at_dev(Ctx* ctx, Fr* data, const SparseAddArgs& a);
int fr_eval_batch_dev(Ctx* ctx, uint32_t count, const Fr* const* coeffs, const size_t* lens, const Fr* xs);
int fr_eval_fetch(Ctx* ctx, Fr* out_host, uint32_t count);
int fr_trimmed_len_dev(Ctx* ctx, const Fr* coeffs, size_t n, size_t* out_len);
int plonk_numden_dev(Ctx* ctx, const PlonkNumDenArgs& p);
int plonk_quotient_dev(Ctx* ctx, const PlonkQuotientArgs& p);
int plonk_gate_check_dev(Ctx* ctx, const Fr* const cols[9], size_t n, bool* ok);
int g1_scalar_mul_dev(Ctx* ctx, const Fr* scalars_host, uint32_t count, G1Xyzz* out_host);  // scalars[k] * srs[0]

}  // namespace zkp
