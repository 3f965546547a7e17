/* zkp_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; the product path (zkp-implementation_b200/) never links, imports or calls it.
 *
 * The reference (sota-zk-labs/zkp-implementation) is Rust on un-vendored arkworks crates
 * (ark-ff/ark-ec/ark-poly 0.4.2, ark-bls12-381 0.4.0 -- kzg/Cargo.toml:12-18, no Cargo.lock) and
 * there is no Rust toolchain in this image, so the reference itself cannot be compiled here
 * (no oracle/_ref).  This file restates, in plain C with 64-bit limbs (arkworks' own limb size,
 * deliberately different from the engine's 32-bit limbs) and Jacobian coordinates (ark-ec's
 * `Projective`, deliberately different from the engine's XYZZ):
 *
 *   orc_msm_naive     kzg/src/scheme.rs:84-96  `evaluate_in_s`: per term MSB-first double-and-add
 *                     (ark-ec `mul_bigint`), `into_affine` (one Fq inversion), then an affine fold
 *                     with `into_affine` after every addition; zip-truncation; empty -> identity.
 *   orc_msm_pippenger the same sum by bucket method, OpenMP over windows -- the "best-effort CPU"
 *                     baseline of BASELINE.md section 3 and the checker for sizes where the naive loop
 *                     would take minutes.
 *   orc_srs           kzg/src/srs.rs:48-69     `Srs::new_from_secret`.
 *   orc_ntt           ark-poly Radix2EvaluationDomain fft/ifft (+ coset), natural order in/out,
 *                     reached from plonk/src/prover.rs:374-375,396-437 and plonk/src/circuit.rs:175.
 *   orc_poly_mul      ark-poly `&DensePolynomial * &DensePolynomial` (2 NTT + pointwise + iNTT).
 *   orc_open_quotient kzg/src/scheme.rs:108-117 Horner + division by (X - z).
 *
 * Parity status: pinned against the only deterministic known-answer test the reference holds
 * (kzg/src/commitment.rs:36-54: secret = 2, commit(1+2X+3X^2) = 17*G), the public BLS12-381
 * constants, and the independent big-integer oracle oracle/pyref.py (tests/test_oracle.py).  The
 * reference has no literal golden vectors, so absolute parity with a Rust run is otherwise unpinned.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;

/* ------------------------------------------------------------------ generic Montgomery (CIOS) */
typedef struct {
  int n;
  u64 p[6];
  u64 inv; /* -p^-1 mod 2^64 */
  u64 one[6];
  u64 r2[6];
} field_t;

static const field_t FR = {4,
                           {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull},
                           0xfffffffeffffffffull,
                           {0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full},
                           {0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull}};

static const field_t FQ = {6,
                           {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull,
                            0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull},
                           0x89f3fffcfffcfffdull,
                           {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull, 0x77ce585370525745ull,
                            0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull},
                           {0xf4df1f341c341746ull, 0x0a76e6a609d104f1ull, 0x8de5476c4c95b6d5ull, 0x67eb88a9939d83c0ull,
                            0x9a793e85b519952dull, 0x11988fe592cae3aaull}};

static inline int ge(const u64* a, const u64* b, int n) {
  for (int i = n - 1; i >= 0; i--) {
    if (a[i] > b[i]) return 1;
    if (a[i] < b[i]) return 0;
  }
  return 1;
}
static inline void sub_n(u64* r, const u64* a, const u64* b, int n) {
  u64 borrow = 0;
  for (int i = 0; i < n; i++) {
    u128 t = (u128)a[i] - b[i] - borrow;
    r[i] = (u64)t;
    borrow = (u64)(t >> 64) & 1;
  }
}
static inline void f_add(const field_t* F, u64* r, const u64* a, const u64* b) {
  u64 c = 0;
  for (int i = 0; i < F->n; i++) {
    u128 t = (u128)a[i] + b[i] + c;
    r[i] = (u64)t;
    c = (u64)(t >> 64);
  }
  if (ge(r, F->p, F->n)) sub_n(r, r, F->p, F->n);
}
static inline void f_sub(const field_t* F, u64* r, const u64* a, const u64* b) {
  u64 t[6];
  if (ge(a, b, F->n)) {
    sub_n(r, a, b, F->n);
  } else {
    sub_n(t, F->p, b, F->n);
    u64 c = 0;
    for (int i = 0; i < F->n; i++) {
      u128 s = (u128)a[i] + t[i] + c;
      r[i] = (u64)s;
      c = (u64)(s >> 64);
    }
  }
}
static inline void f_mul(const field_t* F, u64* r, const u64* a, const u64* b) {
  const int n = F->n;
  u64 t[8] = {0};
  for (int i = 0; i < n; i++) {
    u64 c = 0;
    for (int j = 0; j < n; j++) {
      u128 uv = (u128)a[j] * b[i] + t[j] + c;
      t[j] = (u64)uv;
      c = (u64)(uv >> 64);
    }
    u128 s = (u128)t[n] + c;
    t[n] = (u64)s;
    t[n + 1] = (u64)(s >> 64);
    u64 m = t[0] * F->inv;
    u128 uv = (u128)m * F->p[0] + t[0];
    c = (u64)(uv >> 64);
    for (int j = 1; j < n; j++) {
      uv = (u128)m * F->p[j] + t[j] + c;
      t[j - 1] = (u64)uv;
      c = (u64)(uv >> 64);
    }
    s = (u128)t[n] + c;
    t[n - 1] = (u64)s;
    t[n] = t[n + 1] + (u64)(s >> 64);
  }
  if (t[n] || ge(t, F->p, n)) sub_n(t, t, F->p, n);
  memcpy(r, t, 8 * n);
}
static inline int f_is_zero(const field_t* F, const u64* a) {
  u64 o = 0;
  for (int i = 0; i < F->n; i++) o |= a[i];
  return o == 0;
}
static inline int f_eq(const field_t* F, const u64* a, const u64* b) { return memcmp(a, b, 8 * F->n) == 0; }
static void f_from_mont(const field_t* F, u64* r, const u64* a) {
  u64 one[6] = {1, 0, 0, 0, 0, 0};
  f_mul(F, r, a, one);
}
static void f_to_mont(const field_t* F, u64* r, const u64* a) { f_mul(F, r, a, F->r2); }
/* a^(p-2) */
static void f_inv(const field_t* F, u64* r, const u64* a) {
  u64 e[6], two[6] = {2, 0, 0, 0, 0, 0}, acc[6], base[6];
  sub_n(e, F->p, two, F->n);
  memcpy(acc, F->one, 8 * F->n);
  memcpy(base, a, 8 * F->n);
  for (int i = F->n * 64 - 1; i >= 0; i--) {
    f_mul(F, acc, acc, acc);
    if ((e[i >> 6] >> (i & 63)) & 1) f_mul(F, acc, acc, base);
  }
  memcpy(r, acc, 8 * F->n);
}
__attribute__((unused)) static void f_pow_u64(const field_t* F, u64* r, const u64* a, u64 e) {
  u64 acc[6], base[6];
  memcpy(acc, F->one, 8 * F->n);
  memcpy(base, a, 8 * F->n);
  for (int i = 63; i >= 0; i--) {
    f_mul(F, acc, acc, acc);
    if ((e >> i) & 1) f_mul(F, acc, acc, base);
  }
  memcpy(r, acc, 8 * F->n);
}

/* ------------------------------------------------------------------ G1: Jacobian (ark-ec Projective) */
typedef struct { u64 x[6], y[6]; } aff_t;           /* (0,0) = infinity */
typedef struct { u64 x[6], y[6], z[6]; } jac_t;      /* z = 0 = infinity */

static int aff_is_inf(const aff_t* a) { return f_is_zero(&FQ, a->x) && f_is_zero(&FQ, a->y); }
static void jac_set_inf(jac_t* r) { memset(r, 0, sizeof(*r)); memcpy(r->x, FQ.one, 48); memcpy(r->y, FQ.one, 48); }
static int jac_is_inf(const jac_t* a) { return f_is_zero(&FQ, a->z); }

/* dbl-2009-l (a = 0) */
static void jac_dbl(jac_t* r, const jac_t* p) {
  if (jac_is_inf(p)) { *r = *p; return; }
  u64 A[6], B[6], C[6], D[6], E[6], F_[6], t[6];
  f_mul(&FQ, A, p->x, p->x);
  f_mul(&FQ, B, p->y, p->y);
  f_mul(&FQ, C, B, B);
  f_add(&FQ, t, p->x, B);
  f_mul(&FQ, t, t, t);
  f_sub(&FQ, t, t, A);
  f_sub(&FQ, t, t, C);
  f_add(&FQ, D, t, t);
  f_add(&FQ, E, A, A);
  f_add(&FQ, E, E, A);
  f_mul(&FQ, F_, E, E);
  u64 z3[6];
  f_mul(&FQ, z3, p->y, p->z);
  f_add(&FQ, z3, z3, z3);
  f_sub(&FQ, t, F_, D);
  f_sub(&FQ, r->x, t, D);
  f_sub(&FQ, t, D, r->x);
  f_mul(&FQ, t, E, t);
  u64 c8[6];
  f_add(&FQ, c8, C, C);
  f_add(&FQ, c8, c8, c8);
  f_add(&FQ, c8, c8, c8);
  f_sub(&FQ, r->y, t, c8);
  memcpy(r->z, z3, 48);
}

/* madd-2007-bl with full corner cases (ark-ec `add_assign(&Affine)`) */
static void jac_madd(jac_t* r, const jac_t* p, const aff_t* q) {
  if (aff_is_inf(q)) { *r = *p; return; }
  if (jac_is_inf(p)) {
    memcpy(r->x, q->x, 48); memcpy(r->y, q->y, 48); memcpy(r->z, FQ.one, 48);
    return;
  }
  u64 z1z1[6], u2[6], s2[6], h[6], hh[6], i[6], j[6], rr[6], v[6], t[6];
  f_mul(&FQ, z1z1, p->z, p->z);
  f_mul(&FQ, u2, q->x, z1z1);
  f_mul(&FQ, s2, q->y, p->z);
  f_mul(&FQ, s2, s2, z1z1);
  if (f_eq(&FQ, u2, p->x)) {
    if (f_eq(&FQ, s2, p->y)) { jac_dbl(r, p); return; }
    jac_set_inf(r);
    return;
  }
  f_sub(&FQ, h, u2, p->x);
  f_mul(&FQ, hh, h, h);
  f_add(&FQ, i, hh, hh);
  f_add(&FQ, i, i, i);
  f_mul(&FQ, j, h, i);
  f_sub(&FQ, rr, s2, p->y);
  f_add(&FQ, rr, rr, rr);
  f_mul(&FQ, v, p->x, i);
  jac_t o;
  f_mul(&FQ, o.x, rr, rr);
  f_sub(&FQ, o.x, o.x, j);
  f_sub(&FQ, o.x, o.x, v);
  f_sub(&FQ, o.x, o.x, v);
  f_sub(&FQ, t, v, o.x);
  f_mul(&FQ, t, rr, t);
  u64 yj[6];
  f_mul(&FQ, yj, p->y, j);
  f_add(&FQ, yj, yj, yj);
  f_sub(&FQ, o.y, t, yj);
  f_add(&FQ, o.z, p->z, h);
  f_mul(&FQ, o.z, o.z, o.z);
  f_sub(&FQ, o.z, o.z, z1z1);
  f_sub(&FQ, o.z, o.z, hh);
  *r = o;
}

/* add-2007-bl, all corner cases */
static void jac_add(jac_t* r, const jac_t* p, const jac_t* q) {
  if (jac_is_inf(p)) { *r = *q; return; }
  if (jac_is_inf(q)) { *r = *p; return; }
  u64 z1z1[6], z2z2[6], u1[6], u2[6], s1[6], s2[6], h[6], i[6], j[6], rr[6], v[6], t[6];
  f_mul(&FQ, z1z1, p->z, p->z);
  f_mul(&FQ, z2z2, q->z, q->z);
  f_mul(&FQ, u1, p->x, z2z2);
  f_mul(&FQ, u2, q->x, z1z1);
  f_mul(&FQ, s1, p->y, q->z);
  f_mul(&FQ, s1, s1, z2z2);
  f_mul(&FQ, s2, q->y, p->z);
  f_mul(&FQ, s2, s2, z1z1);
  if (f_eq(&FQ, u1, u2)) {
    if (f_eq(&FQ, s1, s2)) { jac_dbl(r, p); return; }
    jac_set_inf(r);
    return;
  }
  f_sub(&FQ, h, u2, u1);
  f_add(&FQ, i, h, h);
  f_mul(&FQ, i, i, i);
  f_mul(&FQ, j, h, i);
  f_sub(&FQ, rr, s2, s1);
  f_add(&FQ, rr, rr, rr);
  f_mul(&FQ, v, u1, i);
  jac_t o;
  f_mul(&FQ, o.x, rr, rr);
  f_sub(&FQ, o.x, o.x, j);
  f_sub(&FQ, o.x, o.x, v);
  f_sub(&FQ, o.x, o.x, v);
  f_sub(&FQ, t, v, o.x);
  f_mul(&FQ, t, rr, t);
  u64 sj[6];
  f_mul(&FQ, sj, s1, j);
  f_add(&FQ, sj, sj, sj);
  f_sub(&FQ, o.y, t, sj);
  f_add(&FQ, o.z, p->z, q->z);
  f_mul(&FQ, o.z, o.z, o.z);
  f_sub(&FQ, o.z, o.z, z1z1);
  f_sub(&FQ, o.z, o.z, z2z2);
  f_mul(&FQ, o.z, o.z, h);
  *r = o;
}

/* ark-ec `into_affine`: z = 0 -> identity, else (X/Z^2, Y/Z^3) */
static void jac_to_affine(aff_t* r, const jac_t* p) {
  if (jac_is_inf(p)) { memset(r, 0, sizeof(*r)); return; }
  u64 zi[6], zi2[6], zi3[6];
  f_inv(&FQ, zi, p->z);
  f_mul(&FQ, zi2, zi, zi);
  f_mul(&FQ, zi3, zi2, zi);
  f_mul(&FQ, r->x, p->x, zi2);
  f_mul(&FQ, r->y, p->y, zi3);
}

/* ark-ec `mul_bigint`: MSB-first double-and-add over the canonical scalar */
static void aff_mul(jac_t* r, const aff_t* p, const u64 k_canon[4]) {
  jac_t acc;
  jac_set_inf(&acc);
  int started = 0;
  for (int i = 255; i >= 0; i--) {
    int bit = (int)((k_canon[i >> 6] >> (i & 63)) & 1);
    if (started) jac_dbl(&acc, &acc);
    if (bit) { jac_madd(&acc, &acc, p); started = 1; }
  }
  *r = acc;
}

static void aff_neg(aff_t* r, const aff_t* p) {
  *r = *p;
  if (!f_is_zero(&FQ, p->y)) sub_n(r->y, FQ.p, p->y, 6);
}

static const aff_t* generator(void) {
  static aff_t g;
  static int init = 0;
  if (!init) {
    const u64 gx[6] = {0xfb3af00adb22c6bbull, 0x6c55e83ff97a1aefull, 0xa14e3a3f171bac58ull, 0xc3688c4f9774b905ull,
                       0x2695638c4fa9ac0full, 0x17f1d3a73197d794ull};
    const u64 gy[6] = {0x0caa232946c5e7e1ull, 0xd03cc744a2888ae4ull, 0x00db18cb2c04b3edull, 0xfcf5e095d5d00af6ull,
                       0xa09e30ed741d8ae4ull, 0x08b3f481e3aaa0f1ull};
    f_to_mont(&FQ, g.x, gx);
    f_to_mont(&FQ, g.y, gy);
    init = 1;
  }
  return &g;
}

/* ------------------------------------------------------------------ exported: KZG */
/* kzg/src/scheme.rs:84-96, literally: per-term scalar mul + into_affine, affine fold + into_affine */
void orc_msm_naive(const u64* scalars_mont, const u64* bases_xy, size_t n, u64* out_xy) {
  aff_t acc;
  memset(&acc, 0, sizeof(acc));
  int have = 0;
  for (size_t i = 0; i < n; i++) {
    u64 k[4];
    f_from_mont(&FR, k, scalars_mont + 4 * i); /* cof.into_bigint() */
    jac_t t;
    aff_t term;
    aff_mul(&t, (const aff_t*)(bases_xy + 12 * i), k);
    jac_to_affine(&term, &t); /* .into_affine() at :92 */
    if (!have) {
      acc = term;
      have = 1;
    } else {
      jac_t a;
      memcpy(a.x, acc.x, 48); memcpy(a.y, acc.y, 48);
      if (aff_is_inf(&acc)) jac_set_inf(&a); else memcpy(a.z, FQ.one, 48);
      jac_madd(&a, &a, &term);   /* acc.add(e) */
      jac_to_affine(&acc, &a);   /* .into_affine() at :93 */
    }
  }
  memcpy(out_xy, &acc, sizeof(acc)); /* unwrap_or(zero) -> (0,0) */
}

static int pick_window(size_t n) {
  int best = 4;
  double bc = 1e300;
  for (int c = 4; c <= 18; c++) {
    int W = 255 / c + 1;
    double cost = (double)n * W + 2.0 * (double)(1u << (c - 1)) * W;
    if (cost < bc) { bc = cost; best = c; }
  }
  return best;
}

/* Bucket method, signed digits; windows distributed over OpenMP threads. */
void orc_msm_pippenger(const u64* scalars_mont, const u64* bases_xy, size_t n, u64* out_xy, int threads) {
  aff_t res;
  memset(&res, 0, sizeof(res));
  if (n == 0) { memcpy(out_xy, &res, sizeof(res)); return; }
  const int c = pick_window(n);
  const int W = 255 / c + 1;
  const size_t nb = (size_t)1 << (c - 1);
  int32_t* digits = (int32_t*)malloc(sizeof(int32_t) * n * W);
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(static)
#endif
  for (long i = 0; i < (long)n; i++) {
    u64 k[5] = {0, 0, 0, 0, 0};
    f_from_mont(&FR, k, scalars_mont + 4 * i);
    int carry = 0;
    for (int w = 0; w < W; w++) {
      int bit = w * c, limb = bit >> 6, off = bit & 63;
      u64 raw = k[limb] >> off;
      if (off + c > 64 && limb + 1 < 5) raw |= k[limb + 1] << (64 - off);
      int d = (int)(raw & (((u64)1 << c) - 1)) + carry;
      if (d > (int)nb) { d -= (1 << c); carry = 1; } else carry = 0;
      digits[(size_t)w * n + i] = d;
    }
  }
  jac_t* wsum = (jac_t*)malloc(sizeof(jac_t) * W);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int w = 0; w < W; w++) {
    jac_t* buckets = (jac_t*)malloc(sizeof(jac_t) * nb);
    for (size_t b = 0; b < nb; b++) jac_set_inf(&buckets[b]);
    const int32_t* dg = digits + (size_t)w * n;
    for (size_t i = 0; i < n; i++) {
      int d = dg[i];
      if (!d) continue;
      const aff_t* p = (const aff_t*)(bases_xy + 12 * i);
      if (d > 0) {
        jac_madd(&buckets[d - 1], &buckets[d - 1], p);
      } else {
        aff_t np;
        aff_neg(&np, p);
        jac_madd(&buckets[-d - 1], &buckets[-d - 1], &np);
      }
    }
    jac_t run, tot;
    jac_set_inf(&run);
    jac_set_inf(&tot);
    for (long b = (long)nb - 1; b >= 0; b--) {
      jac_add(&run, &run, &buckets[b]);
      jac_add(&tot, &tot, &run);
    }
    wsum[w] = tot;
    free(buckets);
  }
  jac_t acc = wsum[W - 1];
  for (int w = W - 2; w >= 0; w--) {
    for (int k = 0; k < c; k++) jac_dbl(&acc, &acc);
    jac_add(&acc, &acc, &wsum[w]);
  }
  jac_to_affine(&res, &acc);
  memcpy(out_xy, &res, sizeof(res));
  free(wsum);
  free(digits);
}

/* kzg/src/srs.rs:48-69: size+3 is applied by the caller; here `count` points [secret^i * G] */
void orc_srs(const u64 secret_mont[4], size_t count, u64* out_xy) {
  u64 cur[4];
  memcpy(cur, FR.one, 32);
  for (size_t i = 0; i < count; i++) {
    u64 k[4];
    jac_t t;
    f_from_mont(&FR, k, cur);
    aff_mul(&t, generator(), k);
    jac_to_affine((aff_t*)(out_xy + 12 * i), &t);
    f_mul(&FR, cur, cur, secret_mont);
  }
}

int orc_g1_on_curve(const u64* xy) {
  const aff_t* a = (const aff_t*)xy;
  if (aff_is_inf(a)) return 1;
  u64 l[6], r[6], four[6] = {4, 0, 0, 0, 0, 0}, b[6];
  f_to_mont(&FQ, b, four);
  f_mul(&FQ, l, a->y, a->y);
  f_mul(&FQ, r, a->x, a->x);
  f_mul(&FQ, r, r, a->x);
  f_add(&FQ, r, r, b);
  return f_eq(&FQ, l, r);
}

/* kzg/src/scheme.rs:108-117: y = p(z) (Horner), q = (p - y)/(X - z); returns y, writes n-1 coeffs */
void orc_open_quotient(const u64* coeffs_mont, size_t n, const u64 z_mont[4], u64* q_out, u64 y_out[4]) {
  u64 carry[4] = {0, 0, 0, 0}, t[4];
  for (size_t i = n; i-- > 1;) {
    f_mul(&FR, t, carry, z_mont);
    f_add(&FR, carry, t, coeffs_mont + 4 * i);
    memcpy(q_out + 4 * (i - 1), carry, 32);
  }
  f_mul(&FR, t, carry, z_mont);
  f_add(&FR, y_out, t, coeffs_mont);
}

/* ------------------------------------------------------------------ exported: NTT (ark-poly Radix2) */
static void fr_root(u64* w, int log_n) {
  /* TWO_ADIC_ROOT_OF_UNITY = 7^((r-1)/2^32); group_gen = root^(2^(32-log_n)) */
  u64 seven[4] = {7, 0, 0, 0}, g[4], acc[4];
  f_to_mont(&FR, g, seven);
  /* exponent (r-1) >> 32 */
  u64 e[4];
  u64 rm1[4];
  memcpy(rm1, FR.p, 32);
  rm1[0] -= 1;
  for (int i = 0; i < 4; i++) e[i] = (rm1[i] >> 32) | (i + 1 < 4 ? rm1[i + 1] << 32 : 0);
  memcpy(acc, FR.one, 32);
  for (int i = 255; i >= 0; i--) {
    f_mul(&FR, acc, acc, acc);
    if ((e[i >> 6] >> (i & 63)) & 1) f_mul(&FR, acc, acc, g);
  }
  for (int i = log_n; i < 32; i++) f_mul(&FR, acc, acc, acc);
  memcpy(w, acc, 32);
}

static size_t bitrev(size_t x, int bits) {
  size_t r = 0;
  for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
}

/* In-order in, in-order out; iterative DIT after a bit-reversal permutation.
 * inverse: omega^-1 and a final multiplication by N^-1.  coset (may be NULL): forward multiplies
 * x[j] by h^j first; inverse multiplies the result by h^-j last (ark-poly coset fft/ifft). */
void orc_ntt(u64* data, int log_n, int inverse, const u64* coset_mont, int threads) {
  const size_t n = (size_t)1 << log_n;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#else
  (void)threads;
#endif
  if (coset_mont && !inverse) {
    u64 hp[4];
    memcpy(hp, FR.one, 32);
    for (size_t j = 0; j < n; j++) {
      f_mul(&FR, data + 4 * j, data + 4 * j, hp);
      f_mul(&FR, hp, hp, coset_mont);
    }
  }
  u64 w[4];
  fr_root(w, log_n);
  if (inverse) f_inv(&FR, w, w);
  for (size_t i = 0; i < n; i++) {
    size_t j = bitrev(i, log_n);
    if (i < j) {
      u64 t[4];
      memcpy(t, data + 4 * i, 32);
      memcpy(data + 4 * i, data + 4 * j, 32);
      memcpy(data + 4 * j, t, 32);
    }
  }
  /* twiddles w^k, k < n/2 */
  u64* tw = (u64*)malloc(32 * (n / 2 + 1));
  memcpy(tw, FR.one, 32);
  for (size_t k = 1; k < n / 2; k++) f_mul(&FR, tw + 4 * k, tw + 4 * (k - 1), w);
  for (int s = 1; s <= log_n; s++) {
    const size_t m = (size_t)1 << s, half = m >> 1, stride = n / m;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long blk = 0; blk < (long)(n / m); blk++) {
      u64* base = data + 4 * (size_t)blk * m;
      for (size_t k = 0; k < half; k++) {
        u64 t[4], u[4];
        f_mul(&FR, t, base + 4 * (k + half), tw + 4 * (k * stride));
        memcpy(u, base + 4 * k, 32);
        f_add(&FR, base + 4 * k, u, t);
        f_sub(&FR, base + 4 * (k + half), u, t);
      }
    }
  }
  free(tw);
  if (inverse) {
    u64 nn[4] = {(u64)n, 0, 0, 0}, ninv[4];
    f_to_mont(&FR, nn, nn);
    f_inv(&FR, ninv, nn);
    for (size_t j = 0; j < n; j++) f_mul(&FR, data + 4 * j, data + 4 * j, ninv);
    if (coset_mont) {
      u64 hinv[4], hp[4];
      f_inv(&FR, hinv, coset_mont);
      memcpy(hp, FR.one, 32);
      for (size_t j = 0; j < n; j++) {
        f_mul(&FR, data + 4 * j, data + 4 * j, hp);
        f_mul(&FR, hp, hp, hinv);
      }
    }
  }
}

/* ark-poly `&a * &b`: D = next_pow2(la + lb - 1); fft, fft, pointwise, ifft.  out: la+lb-1 coeffs */
void orc_poly_mul(const u64* a, size_t la, const u64* b, size_t lb, u64* out, int threads) {
  if (!la || !lb) return;
  size_t lo = la + lb - 1;
  int log_n = 0;
  while (((size_t)1 << log_n) < lo) log_n++;
  size_t n = (size_t)1 << log_n;
  u64* fa = (u64*)calloc(n, 32);
  u64* fb = (u64*)calloc(n, 32);
  memcpy(fa, a, 32 * la);
  memcpy(fb, b, 32 * lb);
  orc_ntt(fa, log_n, 0, NULL, threads);
  orc_ntt(fb, log_n, 0, NULL, threads);
  for (size_t i = 0; i < n; i++) f_mul(&FR, fa + 4 * i, fa + 4 * i, fb + 4 * i);
  orc_ntt(fa, log_n, 1, NULL, threads);
  memcpy(out, fa, 32 * lo);
  free(fa);
  free(fb);
}

/* field helpers exposed for the test-suite */
void orc_fr_mul(const u64* a, const u64* b, u64* r) { f_mul(&FR, r, a, b); }
void orc_fq_mul(const u64* a, const u64* b, u64* r) { f_mul(&FQ, r, a, b); }
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
