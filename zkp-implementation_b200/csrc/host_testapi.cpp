// Host-only build of the field / curve headers for the CPU test-suite (`pytest -m "not gpu"`).
// Compiled by g++ with -DZKP_FIELD_CHAIN_ON_HOST toggled per call so both multipliers (portable
// CIOS and the even/odd carry-chain algorithm that the GPU runs, here on the emulated carry flag)
// are checked against the big-integer oracle.  Not part of the product library.
#include <string.h>

#include "curve.cuh"
#include "inv_gcd.cuh"

using namespace zkp;

extern "C" {

// op: 0 add, 1 sub, 2 mul (portable), 3 mul (carry-chain algorithm), 4 inv, 5 to_mont, 6 from_mont, 7 mul (64-bit host CIOS),
//     8 sqr (dedicated carry-chain squaring), 9 inv by division steps (inv_gcd.cuh),
//     10 mul (Karatsuba product + separated reduction, the -DZKP_KARATSUBA_* experiment)
int zkp_t_fr_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Fr x, y, r;
  memcpy(x.v, a, 32);
  if (b) memcpy(y.v, b, 32);
  switch (op) {
    case 0: r = fp_add(x, y); break;
    case 1: r = fp_sub(x, y); break;
    case 2: r = fp_mul_portable(x, y); break;
    case 3: r = fp_mul_chain(x, y); break;
    case 4: r = fp_inv(x); break;
    case 5: r = fp_to_mont(x); break;
    case 6: r = fp_from_mont(x); break;
    case 7: r = fp_mul_host64(x, y); break;
    case 8: r = fp_sqr_chain(x); break;
    case 9: r = fr_inv_gcd(x); break;
    case 10: r = fp_mul_kara(x, y); break;
    default: return 1;
  }
  memcpy(out, r.v, 32);
  return 0;
}

int zkp_t_fq_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Fq x, y, r;
  memcpy(x.v, a, 48);
  if (b) memcpy(y.v, b, 48);
  switch (op) {
    case 0: r = fp_add(x, y); break;
    case 1: r = fp_sub(x, y); break;
    case 2: r = fp_mul_portable(x, y); break;
    case 3: r = fp_mul_chain(x, y); break;
    case 4: r = fp_inv(x); break;
    case 5: r = fp_to_mont(x); break;
    case 6: r = fp_from_mont(x); break;
    case 7: r = fp_mul_host64(x, y); break;
    case 8: r = fp_sqr_chain(x); break;
    case 9: r = fq_inv_gcd(x); break;
    case 10: r = fp_mul_kara(x, y); break;
    default: return 1;
  }
  memcpy(out, r.v, 48);
  return 0;
}

// op: 0 madd (acc xyzz += q affine), 1 add (xyzz += xyzz), 2 dbl, 3 to_affine (out: 24 words), 4 mul_u32 (k in b[0])
int zkp_t_g1_op(int op, const uint32_t* a /*48 words xyzz*/, const uint32_t* b, uint32_t* out) {
  G1Xyzz acc;
  memcpy(&acc, a, sizeof(acc));
  switch (op) {
    case 0: { G1Affine q; memcpy(&q, b, sizeof(q)); xyzz_madd(acc, q); break; }
    case 1: { G1Xyzz q; memcpy(&q, b, sizeof(q)); xyzz_add(acc, q); break; }
    case 2: acc = xyzz_dbl(acc); break;
    case 3: { G1Affine r = xyzz_to_affine(acc); memcpy(out, &r, sizeof(r)); return 0; }
    case 4: acc = xyzz_mul_u32(acc, b[0]); break;
    default: return 1;
  }
  memcpy(out, &acc, sizeof(acc));
  return 0;
}

}  // extern "C"
