// Fq / Fr inversion by division steps ("safegcd", D. J. Bernstein and B.-Y. Yang, "Fast constant-time gcd computation and
// modular inversion", TCHES 2019) instead of Fermat's a^(p-2).
//
// Why: every batched-affine round of the MSM (msm_affine.cu) ends in ONE inversion per <= 4096 running products, with one
// warp per SM -- pure latency.  a^(p-2) is 381 squarings + 189 products = 570 dependent Fq products (~0.51 ms on B200);
// the division steps are ~1100 shift/add steps on the low 30 bits plus 37 updates of four 390-bit numbers by a 2 x 2
// matrix of 31-bit entries: ~10x fewer dependent instructions.  The result is the same field element (the inverse is
// unique), so nothing downstream changes; `tests/test_host_arith.py` checks it against pow(a, p - 2, p) and the Fermat ladder.
//
// Formulation (the published algorithm, restated for 32-bit registers; 64-bit sums are what the compiler makes of
// IMAD.WIDE + IADD3.X):
//   divstep(delta, f, g) = (1 - delta, g, (g - f) / 2)            if delta > 0 and g odd
//                          (1 + delta, f, (g + (g mod 2) f) / 2)  otherwise
//   start (1, p, a); after m = floor((49 * 381 + 57) / 17) = 1101 steps g = 0 and f = +-gcd = +-1 (Theorem 11.2 of the
//   paper, d = 381 >= 46).  30 steps depend only on the low 30 bits of f, g and give a matrix T with |entries| <= 2^30,
//   2^30 (f', g') = T (f, g); the same T (with a multiple of p added to make the division by 2^30 exact) keeps
//   f = d * a, g = e * a (mod p) for (d, e) started at (0, 1), so that at the end d = +-1/a.
//   Numbers are 13 limbs of 30 bits, limbs 0..11 in [0, 2^30), limb 12 signed.
// The value passed in is the Montgomery residue aR; 1/(aR) is multiplied by R^2 twice to give (1/a) R.
#pragma once
#include "field.cuh"

namespace zkp {
namespace sgcd {

static constexpr int32_t M30 = 0x3fffffff;
// Per field: L limbs of 30 bits (values live in (-2p, p): bits of p + 1 + sign), p^-1 mod 2^30, and the number of
// 30-step chunks that covers floor((49 d + 57) / 17) steps for d = bits of p.
template <class P> struct Cfg;
template <> struct Cfg<FqParams> {
  static constexpr int L = 13;
  static constexpr uint32_t P_INV30 = 0x00030003u;
  static constexpr int CHUNKS = 37;  // d = 381: 1101 steps <= 37 * 30
};
template <> struct Cfg<FrParams> {
  static constexpr int L = 9;
  static constexpr uint32_t P_INV30 = 0x00000001u;  // r = 1 mod 2^32
  static constexpr int CHUNKS = 25;  // d = 255: 738 steps <= 25 * 30
};

template <class P>
ZKP_HD constexpr uint32_t p30(int j) {
  const int bit = 30 * j, w = bit >> 5, s = bit & 31;
  const uint64_t lo = (w < P::N) ? P::mod(w) : 0, hi = (w + 1 < P::N) ? P::mod(w + 1) : 0;
  return (uint32_t)(((lo | (hi << 32)) >> s) & 0x3fffffffu);
}

// 30 division steps on the low bits; t = (u, v, q, r) with 2^30 f' = u f + v g, 2^30 g' = q f + r g
ZKP_HD void divsteps30(int32_t& delta, uint32_t f, uint32_t g, int32_t (&t)[4]) {
  uint32_t u = 1, v = 0, q = 0, r = 1;
  int32_t d = delta;
#pragma unroll 6
  for (int i = 0; i < 30; i++) {
    const uint32_t odd = 0u - (g & 1u);
    const uint32_t sw = odd & (uint32_t)((-d) >> 31);  // delta > 0 and g odd
    // (f, g) <- (g, -f), (u, v, q, r) <- (q, r, -u, -v), delta <- -delta when sw
    uint32_t x;
    x = (f ^ g) & sw; f ^= x; g ^= x;
    x = (u ^ q) & sw; u ^= x; q ^= x;
    x = (v ^ r) & sw; v ^= x; r ^= x;
    g = (g ^ sw) - sw;
    q = (q ^ sw) - sw;
    r = (r ^ sw) - sw;
    d = (int32_t)(((uint32_t)d ^ sw) - sw) + 1;
    g += f & odd;
    q += u & odd;
    r += v & odd;
    g >>= 1;  // one valid high bit is lost per step: 32 - 30 are left when the last parity is read
    u <<= 1;
    v <<= 1;
  }
  delta = d;
  t[0] = (int32_t)u; t[1] = (int32_t)v; t[2] = (int32_t)q; t[3] = (int32_t)r;
}

// (f, g) <- T (f, g) / 2^30, exact
template <int L>
ZKP_HD void update_fg(int32_t (&f)[L], int32_t (&g)[L], const int32_t (&t)[4]) {
  const int64_t u = t[0], v = t[1], q = t[2], r = t[3];
  int64_t cf = u * f[0] + v * g[0], cg = q * f[0] + r * g[0];
  cf >>= 30;
  cg >>= 30;
#pragma unroll
  for (int i = 1; i < L; i++) {
    const int64_t fi = f[i], gi = g[i];
    cf += u * fi + v * gi;
    cg += q * fi + r * gi;
    f[i - 1] = (int32_t)cf & M30;
    g[i - 1] = (int32_t)cg & M30;
    cf >>= 30;
    cg >>= 30;
  }
  f[L - 1] = (int32_t)cf;
  g[L - 1] = (int32_t)cg;
}

// (d, e) <- T (d, e) / 2^30 mod p; d, e stay in (-2p, p)
template <class P, int L>
ZKP_HD void update_de(int32_t (&d)[L], int32_t (&e)[L], const int32_t (&t)[4]) {
  constexpr uint32_t P_INV30 = Cfg<P>::P_INV30;
  const int32_t u = t[0], v = t[1], q = t[2], r = t[3];
  const int32_t sd = d[L - 1] >> 31, se = e[L - 1] >> 31;
  // one p per negative operand keeps the sums from drifting down ...
  int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
  int64_t cd = (int64_t)u * d[0] + (int64_t)v * e[0], ce = (int64_t)q * d[0] + (int64_t)r * e[0];
  // ... and the multiple of p that clears the low 30 bits
  md -= (int32_t)((P_INV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
  me -= (int32_t)((P_INV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
  cd += (int64_t)p30<P>(0) * md;
  ce += (int64_t)p30<P>(0) * me;
  cd >>= 30;
  ce >>= 30;
#pragma unroll
  for (int i = 1; i < L; i++) {
    const int64_t di = d[i], ei = e[i];
    cd += (int64_t)u * di + (int64_t)v * ei + (int64_t)p30<P>(i) * md;
    ce += (int64_t)q * di + (int64_t)r * ei + (int64_t)p30<P>(i) * me;
    d[i - 1] = (int32_t)cd & M30;
    e[i - 1] = (int32_t)ce & M30;
    cd >>= 30;
    ce >>= 30;
  }
  d[L - 1] = (int32_t)cd;
  e[L - 1] = (int32_t)ce;
}

// low limbs back into [0, 2^30), the sign stays in the top limb
template <int L>
ZKP_HD void carry(int32_t (&x)[L]) {
#pragma unroll
  for (int i = 0; i < L - 1; i++) {
    x[i + 1] += x[i] >> 30;
    x[i] &= M30;
  }
}

// x in (-2p, p), negated when neg is all ones, brought to [0, p)
template <class P, int L>
ZKP_HD void normalize(int32_t (&x)[L], int32_t neg) {
  int32_t add = x[L - 1] >> 31;
#pragma unroll
  for (int i = 0; i < L; i++) x[i] += (int32_t)p30<P>(i) & add;
  carry(x);
#pragma unroll
  for (int i = 0; i < L; i++) x[i] = (x[i] ^ neg) - neg;
  carry(x);
  add = x[L - 1] >> 31;
#pragma unroll
  for (int i = 0; i < L; i++) x[i] += (int32_t)p30<P>(i) & add;
  carry(x);
}

}  // namespace sgcd

// 1 / a for a Montgomery residue a (0 -> 0, like the Fermat ladder)
template <class P>
ZKP_HD Fp<P> fp_inv_gcd_impl(const Fp<P>& a) {
  using namespace sgcd;
  constexpr int L = Cfg<P>::L, N = P::N;
  int32_t f[L], g[L], d[L], e[L];
#pragma unroll
  for (int j = 0; j < L; j++) {
    const int bit = 30 * j, w = bit >> 5, s = bit & 31;
    const uint32_t lo = (w < N) ? a.v[w < N ? w : 0] : 0u, hi = (w + 1 < N) ? a.v[w + 1 < N ? w + 1 : 0] : 0u;
    const uint32_t x = s ? ((lo >> s) | (hi << (32 - s))) : lo;
    g[j] = (int32_t)(x & (uint32_t)M30);
    f[j] = (int32_t)p30<P>(j);
    d[j] = 0;
    e[j] = 0;
  }
  e[0] = 1;
  int32_t delta = 1;
#pragma unroll 1
  for (int c = 0; c < Cfg<P>::CHUNKS; c++) {
    uint32_t nz = 0;
#pragma unroll
    for (int j = 0; j < L; j++) nz |= (uint32_t)g[j];
    if (nz == 0) break;  // g = 0: further steps change neither f nor d
    int32_t t[4];
    divsteps30(delta, (uint32_t)f[0] | ((uint32_t)f[1] << 30), (uint32_t)g[0] | ((uint32_t)g[1] << 30), t);
    update_de<P, L>(d, e, t);
    update_fg<L>(f, g, t);
  }
  normalize<P, L>(d, f[L - 1] >> 31);  // f = -1: d = -1/a
  // L x 30 bits -> N x 32 bits
  Fp<P> r;
  {
    uint64_t acc = 0;
    int bits = 0, w = 0;
#pragma unroll
    for (int j = 0; j < L; j++) {
      acc |= (uint64_t)(uint32_t)d[j] << bits;
      bits += 30;
      if (bits >= 32 && w < N) {
        r.v[w++] = (uint32_t)acc;
        acc >>= 32;
        bits -= 32;
      }
    }
  }
  const Fp<P> r2 = Fp<P>::r2();
  return fp_mul(fp_mul(r, r2), r2);
}

ZKP_HD_NOINLINE Fq fq_inv_gcd(const Fq& a) { return fp_inv_gcd_impl<FqParams>(a); }
ZKP_HD_NOINLINE Fr fr_inv_gcd(const Fr& a) { return fp_inv_gcd_impl<FrParams>(a); }

}  // namespace zkp
