// `KzgScheme::open` (kzg/src/scheme.rs:108-120) over the GPU engine: y = p(z) (Horner in the reference, chunked
// Horner + tree on the device here), q = (p - y) / (X - z) (ark-poly long division in the reference; weight by
// z^i, suffix sums, unweight on the device here), commitment of q against the resident SRS.  Pure client of the C
// ABI in include/zkp_b200.h; values are exact field / group elements, so (witness, y) equal the reference's.
#include <string.h>

#include "../../include/zkp_b200.h"
#include "mont_host.hpp"

using namespace zkp_host;

#define KZG_TRY(e)            \
  do {                        \
    int st_ = (e);            \
    if (st_ != 0) { zkp_dev_free(ctx, buf); return st_; } \
  } while (0)

extern "C" int zkp_kzg_open(zkp_ctx* ctx, const uint64_t* coeffs, size_t n, const uint64_t z_in[4], uint64_t out_xy[12],
                            uint8_t* out_infinity, uint64_t out_y[4]) {
  if (!ctx || !z_in || !out_xy || !out_y || (n && !coeffs)) return ZKP_B200_ERR_INVALID_ARG;
  // DensePolynomial::from_coefficients_* trims trailing zeros; an empty polynomial hits expect("at least 1")
  while (n && (coeffs[4 * (n - 1)] | coeffs[4 * (n - 1) + 1] | coeffs[4 * (n - 1) + 2] | coeffs[4 * (n - 1) + 3]) == 0) n--;
  if (n == 0) return ZKP_B200_ERR_EMPTY_POLY;
  if (n - 1 > zkp_srs_len(ctx)) return ZKP_B200_ERR_SRS_TOO_SMALL;  // the quotient has n - 1 coefficients
  Fr z;
  memcpy(z.v, z_in, 32);
  Fr* buf = nullptr;  // [0, n): p, then p - y, then the quotient at buf + 1;  [n, 2n): z^i;  [2n, 3n): z^-i
  int st = zkp_dev_alloc(ctx, 3 * n * 32, (void**)&buf);
  if (st) return st;
  KZG_TRY(zkp_dev_upload(ctx, buf, coeffs, n * 32));
  Fr y;
  {
    const void* polys[1] = {buf};
    const size_t lens[1] = {n};
    KZG_TRY(zkp_fr_eval_dev(ctx, 1, polys, lens, z.v, y.v));
  }
  memcpy(out_y, y.v, 32);
  uint8_t inf = 1;
  memset(out_xy, 0, 96);
  if (n > 1) {
    const Fr one = Fr::one();
    if (z.is_zero()) {
      // (p - p(0)) / X: the quotient is the coefficient vector shifted down by one
      KZG_TRY(zkp_msm_g1_dev(ctx, buf + 1, nullptr, n - 1, out_xy, &inf));
    } else {
      const size_t idx[1] = {0};
      const Fr neg_y = -y, zinv = fr_inv(z);
      KZG_TRY(zkp_fr_add_at_dev(ctx, buf, 1, idx, neg_y.v));
      KZG_TRY(zkp_fr_powers_dev(ctx, buf + n, z.v, one.v, n));
      KZG_TRY(zkp_fr_powers_dev(ctx, buf + 2 * n, zinv.v, one.v, n));
      KZG_TRY(zkp_fr_mul_pointwise_dev(ctx, buf, buf + n, n));
      KZG_TRY(zkp_fr_scan_dev(ctx, buf, n, 1, 1));          // suffix sums of c_i z^i
      KZG_TRY(zkp_fr_mul_pointwise_dev(ctx, buf, buf + 2 * n, n));  // q_j = S_(j+1) z^-(j+1)
      KZG_TRY(zkp_msm_g1_dev(ctx, buf + 1, nullptr, n - 1, out_xy, &inf));
    }
  }
  if (out_infinity) *out_infinity = inf;
  zkp_dev_free(ctx, buf);
  return 0;
}
