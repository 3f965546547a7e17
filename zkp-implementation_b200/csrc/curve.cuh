// BLS12-381 G1 group law in extended Jacobian ("XYZZ") coordinates: x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2.
//
// Replaces ark-ec `short_weierstrass::{Affine, Projective}` as used by kzg/src/scheme.rs:92-93
// (`s.mul(cof).into_affine()` / `acc.add(e).into_affine()`).  The reference normalises to affine
// after every term; a group element has one normalised affine representative, so the engine keeps
// XYZZ accumulators and normalises once at the very end (`xyzz_to_affine`).
//
// Affine points cross the C ABI as x || y (2 x 48 bytes, Montgomery limbs); the all-zero encoding
// (0, 0) is not on y^2 = x^3 + 4 and serves as the point at infinity (ark's `infinity: bool`).
#pragma once
#include "field.cuh"
#include "inv_gcd.cuh"

namespace zkp {

struct alignas(16) G1Affine {
  Fq x, y;
  ZKP_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  ZKP_HD static G1Affine infinity() { G1Affine r; r.x = Fq::zero(); r.y = Fq::zero(); return r; }
};

struct alignas(16) G1Xyzz {
  Fq x, y, zz, zzz;
  ZKP_HD bool is_inf() const { return zz.is_zero(); }
  ZKP_HD static G1Xyzz infinity() {
    G1Xyzz r; r.x = Fq::zero(); r.y = Fq::zero(); r.zz = Fq::zero(); r.zzz = Fq::zero(); return r; }
  ZKP_HD static G1Xyzz from_affine(const G1Affine& a) {
    if (a.is_inf()) return infinity();
    G1Xyzz r; r.x = a.x; r.y = a.y; r.zz = Fq::one(); r.zzz = Fq::one(); return r; }
};

// dbl-2008-s-1 (a = 0): 6M + 3S... written with mul only (sqr == mul for now)
ZKP_HD G1Xyzz xyzz_dbl(const G1Xyzz& p) {
  if (p.is_inf()) return p;
  Fq u = fp_dbl(p.y);
  Fq v = fq_sqr(u);
  Fq w = u * v;
  Fq s = p.x * v;
  Fq xx = fq_sqr(p.x);
  Fq m = fp_add(fp_dbl(xx), xx);
  G1Xyzz r;
  r.x = fp_sub(fq_sqr(m), fp_dbl(s));
  r.y = fp_sub(m * fp_sub(s, r.x), w * p.y);
  r.zz = v * p.zz;
  r.zzz = w * p.zzz;
  return r;
}

// Doubling of an affine point (ZZ = ZZZ = 1): mdbl-2008-s-1
ZKP_HD G1Xyzz xyzz_dbl_affine(const G1Affine& p) {
  Fq u = fp_dbl(p.y);
  Fq v = fq_sqr(u);
  Fq w = u * v;
  Fq s = p.x * v;
  Fq xx = fq_sqr(p.x);
  Fq m = fp_add(fp_dbl(xx), xx);
  G1Xyzz r;
  r.x = fp_sub(fq_sqr(m), fp_dbl(s));
  r.y = fp_sub(m * fp_sub(s, r.x), w * p.y);
  r.zz = v;
  r.zzz = w;
  return r;
}

// Mixed addition acc += q (q affine, not infinity unless flagged): madd-2008-s, 8M + 2S.
// Handles acc = O, q = O, acc = q (doubling) and acc = -q (cancellation).
ZKP_HD void xyzz_madd(G1Xyzz& acc, const G1Affine& q) {
  if (q.is_inf()) return;
  if (acc.is_inf()) { acc.x = q.x; acc.y = q.y; acc.zz = Fq::one(); acc.zzz = Fq::one(); return; }
  Fq u2 = q.x * acc.zz;
  Fq s2 = q.y * acc.zzz;
  Fq p = fp_sub(u2, acc.x);
  Fq r = fp_sub(s2, acc.y);
  if (p.is_zero()) {
    if (r.is_zero()) { acc = xyzz_dbl_affine(q); } else { acc = G1Xyzz::infinity(); }
    return;
  }
  Fq pp = fq_sqr(p);
  Fq ppp = p * pp;
  Fq qq = acc.x * pp;
  Fq x3 = fp_sub(fp_sub(fq_sqr(r), ppp), fp_dbl(qq));
  Fq y3 = fp_sub(r * fp_sub(qq, x3), acc.y * ppp);
  acc.x = x3;
  acc.y = y3;
  acc.zz = acc.zz * pp;
  acc.zzz = acc.zzz * ppp;
}

// Full addition a += b: add-2008-s, 12M + 2S, all corner cases.
ZKP_HD void xyzz_add(G1Xyzz& a, const G1Xyzz& b) {
  if (b.is_inf()) return;
  if (a.is_inf()) { a = b; return; }
  Fq u1 = a.x * b.zz;
  Fq u2 = b.x * a.zz;
  Fq s1 = a.y * b.zzz;
  Fq s2 = b.y * a.zzz;
  Fq p = fp_sub(u2, u1);
  Fq r = fp_sub(s2, s1);
  if (p.is_zero()) {
    if (r.is_zero()) { a = xyzz_dbl(a); } else { a = G1Xyzz::infinity(); }
    return;
  }
  Fq pp = fq_sqr(p);
  Fq ppp = p * pp;
  Fq qq = u1 * pp;
  Fq x3 = fp_sub(fp_sub(fq_sqr(r), ppp), fp_dbl(qq));
  Fq y3 = fp_sub(r * fp_sub(qq, x3), s1 * ppp);
  a.x = x3;
  a.y = y3;
  a.zz = a.zz * b.zz * pp;
  a.zzz = a.zzz * b.zzz * ppp;
}

ZKP_HD G1Affine g1_neg(const G1Affine& p) {
  G1Affine r;
  r.x = p.x;
  r.y = p.y.is_zero() ? p.y : fp_sub(Fq::modulus(), p.y);  // keeps (0,0) = infinity fixed
  return r;
}

// Normalise (host-side finishing: one Fq inversion per MSM).  Matches ark-ec `into_affine`:
// identity -> (0, 0, infinity = true); otherwise the unique (x, y).
ZKP_HD_NOINLINE G1Affine xyzz_to_affine(const G1Xyzz& p) {
  if (p.is_inf()) return G1Affine::infinity();
  // 1/ZZZ, then 1/ZZ = ZZZ^-2 * ZZ^2 ... cheaper: one inversion of ZZ*ZZZ
  Fq t = p.zz * p.zzz;
  Fq ti = fq_inv_gcd(t);
  Fq zz_inv = ti * p.zzz;
  Fq zzz_inv = ti * p.zz;
  G1Affine r;
  r.x = p.x * zz_inv;
  r.y = p.y * zzz_inv;
  return r;
}

// k * p for a small unsigned k (bucket-segment offsets), MSB-first double-and-add.
ZKP_HD_NOINLINE G1Xyzz xyzz_mul_u32(const G1Xyzz& p, uint32_t k) {
  G1Xyzz acc = G1Xyzz::infinity();
  for (int i = 31; i >= 0; i--) {
    acc = xyzz_dbl(acc);
    if ((k >> i) & 1) xyzz_add(acc, p);
  }
  return acc;
}

// k * p for a canonical (non-Montgomery) little-endian multi-limb scalar, MSB-first double-and-add
// (the schedule of ark-ec's `mul_bigint`, used here for setup-time work only).
ZKP_HD_NOINLINE G1Xyzz xyzz_mul_limbs(const G1Xyzz& p, const uint32_t* k, int nlimbs) {
  G1Xyzz acc = G1Xyzz::infinity();
  for (int i = nlimbs * 32 - 1; i >= 0; i--) {
    acc = xyzz_dbl(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) xyzz_add(acc, p);
  }
  return acc;
}

// BLS12-381 G1 generator (Montgomery limbs), ark-bls12-381 `G1Affine::generator()`.
ZKP_HD G1Affine g1_generator() {
  constexpr uint32_t gx[12] = {0xfd530c16u, 0x5cb38790u, 0x9976fff5u, 0x7817fc67u, 0x143ba1c1u, 0x154f95c7u,
                               0xf3d0e747u, 0xf0ae6acdu, 0x21dbf440u, 0xedce6eccu, 0x9e0bfb75u, 0x12017741u};
  constexpr uint32_t gy[12] = {0x0ce72271u, 0xbaac93d5u, 0x7918fd8eu, 0x8c22631au, 0x570725ceu, 0xdd595f13u,
                               0x50405194u, 0x51ac5829u, 0xad0059c0u, 0x0e1c8c3fu, 0x5008a26au, 0x0bbc3efcu};
  G1Affine g;
#pragma unroll
  for (int i = 0; i < 12; i++) { g.x.v[i] = gx[i]; g.y.v[i] = gy[i]; }
  return g;
}

}  // namespace zkp
