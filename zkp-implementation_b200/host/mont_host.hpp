// Host-side Montgomery arithmetic on 64-bit limbs (the layout arkworks uses in memory), for the O(n)
// scalar work the reference's prover also does on the CPU between its MSMs and FFTs
// (plonk/src/prover.rs: polynomial add / scale / Horner / division by a linear factor).
// The GPU engine's own field code is csrc/field.cuh; this header is independent of it.
#pragma once
#include <stdint.h>
#include <string.h>

namespace zkp_host {

typedef unsigned __int128 u128;

template <int N>
struct MontField {
  uint64_t p[N];
  uint64_t inv;     // -p^-1 mod 2^64
  uint64_t one[N];  // R mod p
  uint64_t r2[N];   // R^2 mod p
};

inline const MontField<4>& FR() {
  static const MontField<4> f = {
      {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull},
      0xfffffffeffffffffull,
      {0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full},
      {0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull}};
  return f;
}

inline const MontField<6>& FQ() {
  static const MontField<6> f = {
      {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull,
       0x1a0111ea397fe69aull},
      0x89f3fffcfffcfffdull,
      {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull, 0x77ce585370525745ull, 0x5c071a97a256ec6dull,
       0x15f65ec3fa80e493ull},
      {0xf4df1f341c341746ull, 0x0a76e6a609d104f1ull, 0x8de5476c4c95b6d5ull, 0x67eb88a9939d83c0ull, 0x9a793e85b519952dull,
       0x11988fe592cae3aaull}};
  return f;
}

template <int N>
inline bool geq(const uint64_t* a, const uint64_t* b) {
  for (int i = N - 1; i >= 0; i--) {
    if (a[i] > b[i]) return true;
    if (a[i] < b[i]) return false;
  }
  return true;
}
template <int N>
inline void sub_limbs(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t borrow = 0;
  for (int i = 0; i < N; i++) {
    u128 t = (u128)a[i] - b[i] - borrow;
    r[i] = (uint64_t)t;
    borrow = (uint64_t)(t >> 64) & 1;
  }
}
template <int N>
inline void mont_mul(const MontField<N>& F, uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t t[N + 2] = {0};
  for (int i = 0; i < N; i++) {
    uint64_t c = 0;
    for (int j = 0; j < N; j++) {
      u128 uv = (u128)a[j] * b[i] + t[j] + c;
      t[j] = (uint64_t)uv;
      c = (uint64_t)(uv >> 64);
    }
    u128 s = (u128)t[N] + c;
    t[N] = (uint64_t)s;
    t[N + 1] = (uint64_t)(s >> 64);
    const uint64_t m = t[0] * F.inv;
    u128 uv = (u128)m * F.p[0] + t[0];
    c = (uint64_t)(uv >> 64);
    for (int j = 1; j < N; j++) {
      uv = (u128)m * F.p[j] + t[j] + c;
      t[j - 1] = (uint64_t)uv;
      c = (uint64_t)(uv >> 64);
    }
    s = (u128)t[N] + c;
    t[N - 1] = (uint64_t)s;
    t[N] = t[N + 1] + (uint64_t)(s >> 64);
  }
  if (t[N] || geq<N>(t, F.p)) sub_limbs<N>(t, t, F.p);
  memcpy(r, t, 8 * N);
}

// ---- Fr value type -------------------------------------------------------------------------------
struct Fr {
  uint64_t v[4];
  static Fr zero() { Fr r; memset(r.v, 0, 32); return r; }
  static Fr one() { Fr r; memcpy(r.v, FR().one, 32); return r; }
  static Fr from_u64(uint64_t x) {
    Fr a = zero();
    a.v[0] = x;
    Fr r;
    mont_mul<4>(FR(), r.v, a.v, FR().r2);
    return r;
  }
  bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
  bool operator==(const Fr& o) const { return memcmp(v, o.v, 32) == 0; }
  bool operator!=(const Fr& o) const { return !(*this == o); }
};

inline Fr operator+(const Fr& a, const Fr& b) {
  Fr r;
  uint64_t c = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a.v[i] + b.v[i] + c;
    r.v[i] = (uint64_t)t;
    c = (uint64_t)(t >> 64);
  }
  if (geq<4>(r.v, FR().p)) sub_limbs<4>(r.v, r.v, FR().p);
  return r;
}
inline Fr operator-(const Fr& a, const Fr& b) {
  Fr r;
  if (geq<4>(a.v, b.v)) {
    sub_limbs<4>(r.v, a.v, b.v);
  } else {
    uint64_t t[4];
    sub_limbs<4>(t, FR().p, b.v);
    uint64_t c = 0;
    for (int i = 0; i < 4; i++) {
      u128 s = (u128)a.v[i] + t[i] + c;
      r.v[i] = (uint64_t)s;
      c = (uint64_t)(s >> 64);
    }
  }
  return r;
}
inline Fr operator-(const Fr& a) { return Fr::zero() - a; }
inline Fr operator*(const Fr& a, const Fr& b) {
  Fr r;
  mont_mul<4>(FR(), r.v, a.v, b.v);
  return r;
}
inline Fr& operator+=(Fr& a, const Fr& b) { a = a + b; return a; }
inline Fr& operator-=(Fr& a, const Fr& b) { a = a - b; return a; }
inline Fr& operator*=(Fr& a, const Fr& b) { a = a * b; return a; }

inline Fr fr_pow(Fr base, uint64_t e) {
  Fr acc = Fr::one();
  while (e) {
    if (e & 1) acc = acc * base;
    base = base * base;
    e >>= 1;
  }
  return acc;
}
inline Fr fr_inv(const Fr& a) {  // a^(r-2)
  uint64_t e[4], two[4] = {2, 0, 0, 0};
  sub_limbs<4>(e, FR().p, two);
  Fr acc = Fr::one();
  for (int i = 255; i >= 0; i--) {
    acc = acc * acc;
    if ((e[i >> 6] >> (i & 63)) & 1) acc = acc * a;
  }
  return acc;
}
// group_gen of the size-2^log_n domain: 7^((r-1)/2^32) squared 32 - log_n times (ark-ff FrConfig)
inline Fr fr_omega(uint32_t log_n) {
  Fr seven = Fr::from_u64(7);
  uint64_t rm1[4];
  memcpy(rm1, FR().p, 32);
  rm1[0] -= 1;
  uint64_t e[4];
  for (int i = 0; i < 4; i++) e[i] = (rm1[i] >> 32) | (i + 1 < 4 ? rm1[i + 1] << 32 : 0);
  Fr acc = Fr::one();
  for (int i = 255; i >= 0; i--) {
    acc = acc * acc;
    if ((e[i >> 6] >> (i & 63)) & 1) acc = acc * seven;
  }
  for (uint32_t i = log_n; i < 32; i++) acc = acc * acc;
  return acc;
}

}  // namespace zkp_host
