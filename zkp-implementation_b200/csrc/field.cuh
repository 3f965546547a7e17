// Montgomery arithmetic for BLS12-381 Fr (8 x u32) and Fq (12 x u32).
//
// Replaces ark-ff `Fp<MontBackend<_, 4>>` / `Fp<MontBackend<_, 6>>` as used by the reference through
// kzg/src/types.rs:6-10 (ScalarField / BaseField).  The in-memory representation is the same as
// arkworks': little-endian limbs of a*R mod p with R = 2^256 (Fr) or 2^384 (Fq), so a `[u64; 4]` /
// `[u64; 6]` coming over the C ABI is reinterpreted as 8 / 12 u32 limbs with no conversion.
//
// Two implementations of the multiplier live here:
//   * a portable CIOS one (64-bit temporaries) -- host code and the validation build
//     (-DZKP_FIELD_PORTABLE) use it;
//   * the device one: carry chains written in PTX (`mad.lo.cc/madc.hi.cc`), arranged as two
//     interleaved accumulators (even/odd limb products) so that every lo/hi pair lands on an aligned
//     register pair and ptxas fuses it into one IMAD.WIDE.U32 with carry-in/out.
// The PTX primitives have a host emulation (explicit carry flag) so the very same algorithm text is
// unit-tested on the CPU build before it ever reaches a GPU.
//
// Credit: the even/odd two-accumulator carry-chain schedule of the device multiplier (detail::mul_n, cmad_n,
// madc_n_rshift, mad_n_redc below) follows the public technique of supranational/sppark's `ff/mont_t.cuh`
// (Apache-2.0); it is re-stated here for this engine's limb containers.  The Fr low-limb reduction
// (r = 1 mod 2^32), the dedicated squaring, and the host carry-flag emulation are this repository's own.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZKP_HD __host__ __device__ __forceinline__
#define ZKP_HD_NOINLINE inline __host__ __device__ __noinline__
#else
#define ZKP_HD inline __attribute__((always_inline))
#define ZKP_HD_NOINLINE inline __attribute__((noinline))
#endif

#if defined(__CUDA_ARCH__) && !defined(ZKP_FIELD_PORTABLE)
#define ZKP_PTX_DEVICE 1  // device build: the even/odd carry-chain multiplier
#else
#define ZKP_PTX_DEVICE 0  // host pass (or validation build): portable CIOS multiplier
#endif

namespace zkp {

// ------------------------------------------------------------------------------------------------
// PTX carry-chain primitives (device) and their host emulation.
// ------------------------------------------------------------------------------------------------
namespace ptx {
#if !defined(__CUDA_ARCH__)
// Host emulation of the PTX condition-code register: one explicit carry flag per host thread.
inline uint32_t& cf_ref() { static thread_local uint32_t cf = 0; return cf; }
#endif

#if defined(__CUDA_ARCH__)
#define ZKP_PTX2(name, ins, emu)                                                        \
  ZKP_HD uint32_t name(uint32_t a, uint32_t b) {                                        \
    uint32_t r; asm volatile(ins " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
#define ZKP_PTX3(name, ins, emu)                                                        \
  ZKP_HD uint32_t name(uint32_t a, uint32_t b, uint32_t c) {                            \
    uint32_t r; asm volatile(ins " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
#define ZKP_PTX2(name, ins, emu) ZKP_HD uint32_t name(uint32_t a, uint32_t b) { emu }
#define ZKP_PTX3(name, ins, emu) ZKP_HD uint32_t name(uint32_t a, uint32_t b, uint32_t c) { emu }
#endif

ZKP_PTX2(add_cc, "add.cc.u32", uint64_t t = (uint64_t)a + b; cf_ref() = (uint32_t)(t >> 32); return (uint32_t)t;)
ZKP_PTX2(addc_cc, "addc.cc.u32", uint64_t t = (uint64_t)a + b + cf_ref(); cf_ref() = (uint32_t)(t >> 32); return (uint32_t)t;)
ZKP_PTX2(addc, "addc.u32", return a + b + cf_ref();)
ZKP_PTX2(sub_cc, "sub.cc.u32", uint64_t t = (uint64_t)a - b; cf_ref() = (uint32_t)(t >> 63); return (uint32_t)t;)
ZKP_PTX2(subc_cc, "subc.cc.u32", uint64_t t = (uint64_t)a - b - cf_ref(); cf_ref() = (uint32_t)(t >> 63); return (uint32_t)t;)
ZKP_PTX2(subc, "subc.u32", return a - b - cf_ref();)
ZKP_PTX2(mul_lo, "mul.lo.u32", return a * b;)
ZKP_PTX2(mul_hi, "mul.hi.u32", return (uint32_t)(((uint64_t)a * b) >> 32);)
ZKP_PTX3(mad_lo_cc, "mad.lo.cc.u32", uint64_t t = (uint64_t)(uint32_t)(a * b) + c; cf_ref() = (uint32_t)(t >> 32); return (uint32_t)t;)
ZKP_PTX3(madc_lo_cc, "madc.lo.cc.u32", uint64_t t = (uint64_t)(uint32_t)(a * b) + c + cf_ref(); cf_ref() = (uint32_t)(t >> 32); return (uint32_t)t;)
ZKP_PTX3(madc_hi_cc, "madc.hi.cc.u32", uint64_t t = (((uint64_t)a * b) >> 32) + c + cf_ref(); cf_ref() = (uint32_t)(t >> 32); return (uint32_t)t;)
ZKP_PTX3(madc_hi, "madc.hi.u32", return (uint32_t)(((uint64_t)a * b) >> 32) + c + cf_ref();)
#undef ZKP_PTX2
#undef ZKP_PTX3
}  // namespace ptx

// ------------------------------------------------------------------------------------------------
// Field parameters.  Limb tables are constexpr-function locals so that after full unrolling every
// use folds to an immediate operand (no constant-bank or local-memory traffic).
// ------------------------------------------------------------------------------------------------
struct FrParams {
  static constexpr int N = 8;
  static constexpr bool DEDICATED_SQR = false;  // 3r > 2^256: the squaring rows' partial sums would not fit 8 limbs
  // r = ...ffffffff00000001: limb 0 is 1, limb 1 is 2^32 - 1 and -r^-1 = -1 (mod 2^32), so in every
  // Montgomery reduction step m = -t0, m * r[0] is an addition and m * r[1] = (m << 32) - m is two ALU
  // ops: 6 instead of 8 multiply-adds per step (112 instead of 128 per product).
#ifdef ZKP_FR_GENERIC_REDC
  static constexpr bool LOW_LIMBS_SPECIAL = false;  // experiment knob: plain CIOS reduction rows
#else
  static constexpr bool LOW_LIMBS_SPECIAL = true;
#endif
  static constexpr uint32_t M0 = 0xffffffffu;  // -r^-1 mod 2^32
  ZKP_HD static constexpr uint32_t mod(int i) {
    constexpr uint32_t t[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                               0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
    return t[i];
  }
  ZKP_HD static constexpr uint32_t one(int i) {  // R mod r
    constexpr uint32_t t[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                               0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
    return t[i];
  }
  ZKP_HD static constexpr uint32_t r2(int i) {  // R^2 mod r
    constexpr uint32_t t[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                               0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
    return t[i];
  }
};

struct FqParams {
  static constexpr int N = 12;
  static constexpr bool DEDICATED_SQR = true;   // 3p < 2^383: see fp_sqr_chain
  static constexpr bool LOW_LIMBS_SPECIAL = false;
  static constexpr uint32_t M0 = 0xfffcfffdu;  // -p^-1 mod 2^32
  ZKP_HD static constexpr uint32_t mod(int i) {
    constexpr uint32_t t[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
    return t[i];
  }
  ZKP_HD static constexpr uint32_t one(int i) {  // R mod p
    constexpr uint32_t t[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
    return t[i];
  }
  ZKP_HD static constexpr uint32_t r2(int i) {  // R^2 mod p
    constexpr uint32_t t[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                                0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
    return t[i];
  }
};

// ------------------------------------------------------------------------------------------------
// Fp<P>: value in Montgomery form, fully reduced ([0, p)) between operations.
// ------------------------------------------------------------------------------------------------
template <class P>
struct Fp {
  static constexpr int N = P::N;
  uint32_t v[N];

  ZKP_HD static Fp zero() { Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = 0;
    return r; }
  ZKP_HD static Fp one() { Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::one(i);
    return r; }
  ZKP_HD static Fp r2() { Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::r2(i);
    return r; }
  ZKP_HD static Fp modulus() { Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::mod(i);
    return r; }

  ZKP_HD bool is_zero() const { uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= v[i];
    return o == 0; }
  ZKP_HD bool operator==(const Fp& b) const { uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < N; i++) o |= v[i] ^ b.v[i];
    return o == 0; }
  ZKP_HD bool operator!=(const Fp& b) const { return !(*this == b); }
};

// r = a - p if a >= p else a        (a < 2p, no overflow out of N limbs)
template <class P>
ZKP_HD void fp_final_sub(Fp<P>& a) {
  constexpr int N = P::N;
  uint32_t t[N];
  t[0] = ptx::sub_cc(a.v[0], P::mod(0));
#pragma unroll
  for (int i = 1; i < N; i++) t[i] = ptx::subc_cc(a.v[i], P::mod(i));
  uint32_t borrow = ptx::subc(0u, 0u);  // 0 or 0xffffffff
#pragma unroll
  for (int i = 0; i < N; i++) a.v[i] = borrow ? a.v[i] : t[i];
}

template <class P>
ZKP_HD Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  Fp<P> r;
  r.v[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
  r.v[N - 1] = ptx::addc(a.v[N - 1], b.v[N - 1]);  // both moduli leave a spare top bit
  fp_final_sub(r);
  return r;
}

template <class P>
ZKP_HD Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  Fp<P> r;
  r.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < N; i++) r.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
  uint32_t borrow = ptx::subc(0u, 0u);  // 0 or 0xffffffff
  r.v[0] = ptx::add_cc(r.v[0], P::mod(0) & borrow);
#pragma unroll
  for (int i = 1; i < N - 1; i++) r.v[i] = ptx::addc_cc(r.v[i], P::mod(i) & borrow);
  r.v[N - 1] = ptx::addc(r.v[N - 1], P::mod(N - 1) & borrow);
  return r;
}

template <class P>
ZKP_HD Fp<P> fp_neg(const Fp<P>& a) {
  return fp_sub(Fp<P>::zero(), a);
}

template <class P>
ZKP_HD Fp<P> fp_dbl(const Fp<P>& a) {
  return fp_add(a, a);
}

// ---- portable CIOS Montgomery product (host path, and the -DZKP_FIELD_PORTABLE validation build) --
template <class P>
ZKP_HD Fp<P> fp_mul_portable(const Fp<P>& a, const Fp<P>& b) {
  constexpr int N = P::N;
  uint32_t t[N + 1];
#pragma unroll
  for (int i = 0; i <= N; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < N; j++) {
      uint64_t uv = (uint64_t)a.v[j] * b.v[i] + t[j] + c;
      t[j] = (uint32_t)uv;
      c = uv >> 32;
    }
    uint64_t top = (uint64_t)t[N] + c;  // fits 33 bits; both moduli have >= 1 spare bit so top < 2^32
    uint32_t m = t[0] * P::M0;
    uint64_t uv = (uint64_t)m * P::mod(0) + t[0];
    c = uv >> 32;
#pragma unroll
    for (int j = 1; j < N; j++) {
      uv = (uint64_t)m * P::mod(j) + t[j] + c;
      t[j - 1] = (uint32_t)uv;
      c = uv >> 32;
    }
    uv = top + c;
    t[N - 1] = (uint32_t)uv;
    t[N] = (uint32_t)(uv >> 32);
  }
  Fp<P> r;
#pragma unroll
  for (int i = 0; i < N; i++) r.v[i] = t[i];
  fp_final_sub(r);
  return r;
}

// ---- 64-bit-limb CIOS for the host pass (finishing arithmetic of every MSM, the CPU kernel emulator) ---------
#if !defined(__CUDA_ARCH__) && defined(__SIZEOF_INT128__)
#define ZKP_HOST_MUL64 1
template <class P>
inline Fp<P> fp_mul_host64(const Fp<P>& a, const Fp<P>& b) {
  constexpr int M = P::N / 2;
  static_assert(P::N % 2 == 0, "even limb count");
  typedef unsigned __int128 u128;
  uint64_t x[M], y[M], p[M], t[M + 2];
  for (int i = 0; i < M; i++) {
    x[i] = (uint64_t)a.v[2 * i] | ((uint64_t)a.v[2 * i + 1] << 32);
    y[i] = (uint64_t)b.v[2 * i] | ((uint64_t)b.v[2 * i + 1] << 32);
    p[i] = (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32);
    t[i] = 0;
  }
  t[M] = t[M + 1] = 0;
  // -p^-1 mod 2^64 from the 32-bit constant by one Newton step
  uint64_t pinv = (uint64_t)(0u - P::M0);      // p^-1 mod 2^32
  pinv = pinv * (2 - p[0] * pinv);             // mod 2^64
  const uint64_t ninv = 0 - pinv;
  for (int i = 0; i < M; i++) {
    uint64_t c = 0;
    for (int j = 0; j < M; j++) {
      const u128 uv = (u128)x[j] * y[i] + t[j] + c;
      t[j] = (uint64_t)uv;
      c = (uint64_t)(uv >> 64);
    }
    u128 s2 = (u128)t[M] + c;
    t[M] = (uint64_t)s2;
    t[M + 1] = (uint64_t)(s2 >> 64);
    const uint64_t m = t[0] * ninv;
    u128 uv = (u128)m * p[0] + t[0];
    c = (uint64_t)(uv >> 64);
    for (int j = 1; j < M; j++) {
      uv = (u128)m * p[j] + t[j] + c;
      t[j - 1] = (uint64_t)uv;
      c = (uint64_t)(uv >> 64);
    }
    s2 = (u128)t[M] + c;
    t[M - 1] = (uint64_t)s2;
    t[M] = t[M + 1] + (uint64_t)(s2 >> 64);
  }
  Fp<P> r;
  for (int i = 0; i < M; i++) {
    r.v[2 * i] = (uint32_t)t[i];
    r.v[2 * i + 1] = (uint32_t)(t[i] >> 32);
  }
  fp_final_sub(r);  // t < 2p and both moduli leave spare top bits, so t[M] == 0 here
  return r;
}
#endif

// ---- even/odd carry-chain Montgomery product (device path) ---------------------------------------
#if defined(__CUDACC__)
// 0xffffffff read from the constant bank at run time: keeps ptxas from treating m = -t0 as a negation
// (it then rewrites the m * r[j] products and loses the IMAD.WIDE fusion of the reduction chain).
static __device__ __constant__ uint32_t zkp_opaque_minus_one = 0xffffffffu;
#endif
namespace detail {
// acc[j], acc[j+1] = lo, hi of a[j]*bi for even j  (n products, n even)
template <int n>
ZKP_HD void mul_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
  for (int j = 0; j < n; j += 2) {
    acc[j] = ptx::mul_lo(a[j], bi);
    acc[j + 1] = ptx::mul_hi(a[j], bi);
  }
}
// acc += a[even j]*bi along one carry chain; carry-out is left in CC.CF
template <int n>
ZKP_HD void cmad_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
  acc[0] = ptx::mad_lo_cc(a[0], bi, acc[0]);
  acc[1] = ptx::madc_hi_cc(a[0], bi, acc[1]);
#pragma unroll
  for (int j = 2; j < n; j += 2) {
    acc[j] = ptx::madc_lo_cc(a[j], bi, acc[j]);
    acc[j + 1] = ptx::madc_hi_cc(a[j], bi, acc[j + 1]);
  }
}
// odd[j] = a[j]*bi + odd[j+2] (shift right by two limbs while accumulating); carry-in from CC.CF
template <int n>
ZKP_HD void madc_n_rshift(uint32_t* odd, const uint32_t* a, uint32_t bi) {
#pragma unroll
  for (int j = 0; j < n - 2; j += 2) {
    odd[j] = ptx::madc_lo_cc(a[j], bi, odd[j + 2]);
    odd[j + 1] = ptx::madc_hi_cc(a[j], bi, odd[j + 3]);
  }
  odd[n - 2] = ptx::madc_lo_cc(a[n - 2], bi, 0u);
  odd[n - 1] = ptx::madc_hi(a[n - 2], bi, 0u);
}

template <class P>
struct ModLimbs {
  uint32_t m[P::N + 1];
  ZKP_HD ModLimbs() {
#pragma unroll
    for (int i = 0; i < P::N; i++) m[i] = P::mod(i);
    m[P::N] = 0;
  }
};

template <class P>
ZKP_HD void mad_n_redc(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi, const uint32_t* mod,
                       bool first) {
  constexpr int n = P::N;
  if (first) {
    mul_n<n>(odd, a + 1, bi);
    mul_n<n>(even, a, bi);
  } else {
    even[0] = ptx::add_cc(even[0], odd[1]);
    madc_n_rshift<n>(odd, a + 1, bi);
    cmad_n<n>(even, a, bi);
    odd[n - 1] = ptx::addc(odd[n - 1], 0u);
  }
  if constexpr (P::LOW_LIMBS_SPECIAL) {
#if defined(__CUDA_ARCH__)
    const uint32_t mi = even[0] * zkp_opaque_minus_one;
#else
    const uint32_t mi = 0u - even[0];
#endif
    // odd accumulator += mi * mod[1, 3, 5, ...]; mi * (2^32 - 1) = (hi, lo) = (mi - [mi != 0], -mi)
    const uint32_t lo1 = 0u - mi;
    const uint32_t hi1 = mi - (mi != 0u ? 1u : 0u);
    odd[0] = ptx::add_cc(odd[0], lo1);
    odd[1] = ptx::addc_cc(odd[1], hi1);
#pragma unroll
    for (int j = 2; j < n; j += 2) {
      odd[j] = ptx::madc_lo_cc(mod[j + 1], mi, odd[j]);
      odd[j + 1] = ptx::madc_hi_cc(mod[j + 1], mi, odd[j + 1]);
    }
    // even accumulator += mi * mod[0, 2, 4, ...]; mi * 1 clears limb 0
    even[0] = ptx::add_cc(even[0], mi);
    even[1] = ptx::addc_cc(even[1], 0u);
#pragma unroll
    for (int j = 2; j < n; j += 2) {
      even[j] = ptx::madc_lo_cc(mod[j], mi, even[j]);
      even[j + 1] = ptx::madc_hi_cc(mod[j], mi, even[j + 1]);
    }
    odd[n - 1] = ptx::addc(odd[n - 1], 0u);
  } else {
    const uint32_t mi = even[0] * P::M0;
    cmad_n<n>(odd, mod + 1, mi);
    cmad_n<n>(even, mod, mi);
    odd[n - 1] = ptx::addc(odd[n - 1], 0u);
  }
}
}  // namespace detail

template <class P>
ZKP_HD Fp<P> fp_mul_chain(const Fp<P>& a, const Fp<P>& b) {
  constexpr int n = P::N;
  static_assert(n % 2 == 0, "even limb count");
  // a padded with one zero limb: the odd accumulator multiplies a[1..n] where a[n] = 0 is never read
  // for products, but keeps the index arithmetic uniform.
  uint32_t even[n], odd[n];
  detail::ModLimbs<P> M;
  uint32_t aa[n + 1];
#pragma unroll
  for (int i = 0; i < n; i++) aa[i] = a.v[i];
  aa[n] = 0;
#pragma unroll
  for (int i = 0; i < n; i += 2) {
    detail::mad_n_redc<P>(even, odd, aa, b.v[i], M.m, i == 0);
    detail::mad_n_redc<P>(odd, even, aa, b.v[i + 1], M.m, false);
  }
  // merge: result limb k = even[k] + odd[k+1]
  Fp<P> r;
  r.v[0] = ptx::add_cc(even[0], odd[1]);
#pragma unroll
  for (int i = 1; i < n - 1; i++) r.v[i] = ptx::addc_cc(even[i], odd[i + 1]);
  r.v[n - 1] = ptx::addc(even[n - 1], 0u);
  fp_final_sub(r);
  return r;
}

// ---- Karatsuba product + separated reduction (experiment knob -DZKP_KARATSUBA) --------------------------------
// The interleaved multiplier above spends n^2 limb products on a*b and n^2 (Fr: n (n - 2)) on the reduction rows.  The
// a*b part splits: with a = aL + aH 2^(16 n), b likewise,
//   a b = z0 + (z0 + z2 + (aL - aH)(bH - bL)) 2^(16 n) + z2 2^(32 n),  z0 = aL bL,  z2 = aH bH
// -- three half-size products instead of four (Fq: 108 instead of 144 limb products, Fr: 48 instead of 64), paid for
// with add-with-carry instructions on the ALU pipe, whose issue slots the multiplier-bound kernels leave idle.  The
// price is that the product must exist as 2n plain limbs before the reduction, so the reduction rows run on their own:
// accumulators U (pairs at even limbs) and V (pairs at odd limbs) indexed by ABSOLUTE limb -- the unrolled code needs no
// shifting -- with the upper half of the product fed in one limb per row so that every carry lands in a limb that holds
// only earlier carries (no long ripples), and the carry out of the cleared limb kept in a small separate word.
namespace detail {
// out[2h] = x[h] * y[h], h even: E holds the pairs at even limbs, O the pairs at odd limbs (O[k] is limb k + 1)
template <int h>
ZKP_HD void mul_half(uint32_t* out, const uint32_t* x, const uint32_t* y) {
  static_assert(h % 2 == 0, "even half size");
  uint32_t E[2 * h], O[2 * h];
#pragma unroll
  for (int k = 0; k < 2 * h; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
  for (int i = 0; i < h; i++) {
    const uint32_t yi = y[i];
    if ((i & 1) == 0) {
      E[i] = ptx::mad_lo_cc(x[0], yi, E[i]);
      E[i + 1] = ptx::madc_hi_cc(x[0], yi, E[i + 1]);
#pragma unroll
      for (int j = 2; j < h; j += 2) {
        E[i + j] = ptx::madc_lo_cc(x[j], yi, E[i + j]);
        E[i + j + 1] = ptx::madc_hi_cc(x[j], yi, E[i + j + 1]);
      }
      E[i + h] = ptx::addc(E[i + h], 0u);
      O[i] = ptx::mad_lo_cc(x[1], yi, O[i]);
      O[i + 1] = ptx::madc_hi_cc(x[1], yi, O[i + 1]);
#pragma unroll
      for (int j = 3; j < h; j += 2) {
        O[i + j - 1] = ptx::madc_lo_cc(x[j], yi, O[i + j - 1]);
        O[i + j] = ptx::madc_hi_cc(x[j], yi, O[i + j]);
      }
      O[i + h] = ptx::addc(O[i + h], 0u);
    } else {
      O[i - 1] = ptx::mad_lo_cc(x[0], yi, O[i - 1]);
      O[i] = ptx::madc_hi_cc(x[0], yi, O[i]);
#pragma unroll
      for (int j = 2; j < h; j += 2) {
        O[i + j - 1] = ptx::madc_lo_cc(x[j], yi, O[i + j - 1]);
        O[i + j] = ptx::madc_hi_cc(x[j], yi, O[i + j]);
      }
      O[i + h - 1] = ptx::addc(O[i + h - 1], 0u);
      E[i + 1] = ptx::mad_lo_cc(x[1], yi, E[i + 1]);
      E[i + 2] = ptx::madc_hi_cc(x[1], yi, E[i + 2]);
#pragma unroll
      for (int j = 3; j < h; j += 2) {
        E[i + j] = ptx::madc_lo_cc(x[j], yi, E[i + j]);
        E[i + j + 1] = ptx::madc_hi_cc(x[j], yi, E[i + j + 1]);
      }
      if (i + h + 1 < 2 * h) E[i + h + 1] = ptx::addc(E[i + h + 1], 0u);  // the last row ends at the top limb: no carry
    }
  }
  out[0] = E[0];
  out[1] = ptx::add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 2; k < 2 * h - 1; k++) out[k] = ptx::addc_cc(E[k], O[k - 1]);
  out[2 * h - 1] = ptx::addc(E[2 * h - 1], O[2 * h - 2]);
}

// d = |x - y| over h limbs; returns the mask of x < y (0 or 0xffffffff)
template <int h>
ZKP_HD uint32_t abs_diff(uint32_t* d, const uint32_t* x, const uint32_t* y) {
  d[0] = ptx::sub_cc(x[0], y[0]);
#pragma unroll
  for (int k = 1; k < h; k++) d[k] = ptx::subc_cc(x[k], y[k]);
  const uint32_t m = ptx::subc(0u, 0u);
  d[0] = ptx::add_cc(d[0] ^ m, m & 1u);
#pragma unroll
  for (int k = 1; k < h - 1; k++) d[k] = ptx::addc_cc(d[k] ^ m, 0u);
  d[h - 1] = ptx::addc(d[h - 1] ^ m, 0u);
  return m;
}

// r = T / 2^(32 n) mod p for a 2n-limb T < p 2^(32 n); result < 2p in n limbs
template <class P>
ZKP_HD void redc_wide(uint32_t* r, const uint32_t* T) {
  constexpr int n = P::N;
  uint32_t U[2 * n + 2], V[2 * n + 2];
#pragma unroll
  for (int k = 0; k < 2 * n + 2; k++) { U[k] = (k < n) ? T[k] : 0u; V[k] = 0; }
  uint32_t cy = 0;
#pragma unroll
  for (int i = 0; i < n; i++) {
    // feed the next limb of the upper half
    U[n + i] = ptx::add_cc(U[n + i], T[n + i]);
    U[n + i + 1] = ptx::addc(U[n + i + 1], 0u);
    const uint32_t s = U[i] + V[i] + cy;
    uint32_t* X = (i & 1) ? V : U;  // pairs aligned with limb i
    uint32_t* Y = (i & 1) ? U : V;  // pairs aligned with limb i + 1
    if constexpr (P::LOW_LIMBS_SPECIAL) {
#if defined(__CUDA_ARCH__)
      const uint32_t mi = s * zkp_opaque_minus_one;
#else
      const uint32_t mi = 0u - s;
#endif
      // mod[0] = 1: an addition; mod[1] = 2^32 - 1: (hi, lo) = (mi - [mi != 0], -mi)
      X[i] = ptx::add_cc(X[i], mi);
      X[i + 1] = ptx::addc_cc(X[i + 1], 0u);
#pragma unroll
      for (int j = 2; j < n; j += 2) {
        X[i + j] = ptx::madc_lo_cc(P::mod(j), mi, X[i + j]);
        X[i + j + 1] = ptx::madc_hi_cc(P::mod(j), mi, X[i + j + 1]);
      }
      X[i + n] = ptx::addc_cc(X[i + n], 0u);
      X[i + n + 1] = ptx::addc(X[i + n + 1], 0u);
      const uint32_t lo1 = 0u - mi;
      const uint32_t hi1 = mi - (mi != 0u ? 1u : 0u);
      Y[i + 1] = ptx::add_cc(Y[i + 1], lo1);
      Y[i + 2] = ptx::addc_cc(Y[i + 2], hi1);
#pragma unroll
      for (int j = 3; j < n; j += 2) {
        Y[i + j] = ptx::madc_lo_cc(P::mod(j), mi, Y[i + j]);
        Y[i + j + 1] = ptx::madc_hi_cc(P::mod(j), mi, Y[i + j + 1]);
      }
      Y[i + n + 1] = ptx::addc(Y[i + n + 1], 0u);
    } else {
      const uint32_t mi = s * P::M0;
      X[i] = ptx::mad_lo_cc(P::mod(0), mi, X[i]);
      X[i + 1] = ptx::madc_hi_cc(P::mod(0), mi, X[i + 1]);
#pragma unroll
      for (int j = 2; j < n; j += 2) {
        X[i + j] = ptx::madc_lo_cc(P::mod(j), mi, X[i + j]);
        X[i + j + 1] = ptx::madc_hi_cc(P::mod(j), mi, X[i + j + 1]);
      }
      X[i + n] = ptx::addc_cc(X[i + n], 0u);
      X[i + n + 1] = ptx::addc(X[i + n + 1], 0u);
      Y[i + 1] = ptx::mad_lo_cc(P::mod(1), mi, Y[i + 1]);
      Y[i + 2] = ptx::madc_hi_cc(P::mod(1), mi, Y[i + 2]);
#pragma unroll
      for (int j = 3; j < n; j += 2) {
        Y[i + j] = ptx::madc_lo_cc(P::mod(j), mi, Y[i + j]);
        Y[i + j + 1] = ptx::madc_hi_cc(P::mod(j), mi, Y[i + j + 1]);
      }
      Y[i + n + 1] = ptx::addc(Y[i + n + 1], 0u);
    }
    // limb i is now 0 mod 2^32: keep its carry (0, 1 or 2)
    uint32_t t = ptx::add_cc(U[i], V[i]);
    const uint32_t c1 = ptx::addc(0u, 0u);
    t = ptx::add_cc(t, cy);
    cy = ptx::addc(c1, 0u);
    (void)t;
  }
  r[0] = ptx::add_cc(U[n], V[n]);
#pragma unroll
  for (int k = 1; k < n - 1; k++) r[k] = ptx::addc_cc(U[n + k], V[n + k]);
  r[n - 1] = ptx::addc(U[2 * n - 1], V[2 * n - 1]);
  r[0] = ptx::add_cc(r[0], cy);
#pragma unroll
  for (int k = 1; k < n - 1; k++) r[k] = ptx::addc_cc(r[k], 0u);
  r[n - 1] = ptx::addc(r[n - 1], 0u);
}
}  // namespace detail

template <class P>
ZKP_HD Fp<P> fp_mul_kara(const Fp<P>& a, const Fp<P>& b) {
  constexpr int n = P::N, h = n / 2;
  uint32_t T[2 * n], z1[n], mid[n], da[h], db[h];
  detail::mul_half<h>(T, a.v, b.v);              // z0
  detail::mul_half<h>(T + n, a.v + h, b.v + h);  // z2
  const uint32_t ma = detail::abs_diff<h>(da, a.v, a.v + h);      // |aL - aH|
  const uint32_t mb = detail::abs_diff<h>(db, b.v + h, b.v);      // |bH - bL|
  detail::mul_half<h>(z1, da, db);
  const uint32_t sgn = ma ^ mb;  // all ones: (aL - aH)(bH - bL) = -z1
  // mid = z0 + z2 + sgn z1  (n limbs + a small top word; never negative: it is aL bH + aH bL)
  mid[0] = ptx::add_cc(T[0], T[n]);
#pragma unroll
  for (int k = 1; k < n; k++) mid[k] = ptx::addc_cc(T[k], T[n + k]);
  uint32_t midc = ptx::addc(0u, 0u);
  (void)ptx::add_cc(sgn, sgn);  // carry = 1 when subtracting: ~z1 + 1
#pragma unroll
  for (int k = 0; k < n; k++) mid[k] = ptx::addc_cc(mid[k], z1[k] ^ sgn);
  midc = ptx::addc(midc, sgn);
  // T += mid 2^(32 h)
  T[h] = ptx::add_cc(T[h], mid[0]);
#pragma unroll
  for (int k = 1; k < n; k++) T[h + k] = ptx::addc_cc(T[h + k], mid[k]);
  T[h + n] = ptx::addc_cc(T[h + n], midc);
#pragma unroll
  for (int k = h + n + 1; k < 2 * n - 1; k++) T[k] = ptx::addc_cc(T[k], 0u);
  T[2 * n - 1] = ptx::addc(T[2 * n - 1], 0u);
  Fp<P> r;
  detail::redc_wide<P>(r.v, T);
  fp_final_sub(r);
  return r;
}

// ---- dedicated squaring in the same carry-chain form ---------------------------------------------------
// a^2 = sum_i a_i^2 2^(64 i) + 2 sum_{i<j} a_i a_j 2^(32 (i+j)).  Row i of the interleaved (CIOS) schedule multiplies
// the scalar a_i by the vector  u_i = [a_i, 2 a_(i+1), 2 a_(i+2), ...]  -- the doubled higher limbs, with the bit
// shifted across limb boundaries -- and skips the limbs below i (their products were added, doubled, by rows j < i).
// Everything at limb position k has been added by row k, which is what the interleaved reduction needs.  Row i costs
// n - i products instead of n: n (n + 1) / 2 + n^2 limb products per squaring instead of 2 n^2 (222 instead of 288 for
// Fq); the skipped products become add-with-carry instructions that keep the carry chain and the per-row shift alive.
// The partial sums run ahead of the plain schedule (row i adds the DOUBLED cross terms of a_i at once), so the accumulator
// is bounded by 2a + p < 3p instead of a + p: the modulus must satisfy 3p < 2^(32 n) -- true for Fq (p < 2^381), not for Fr
// (r ~ 0.45 * 2^256), which keeps fp_mul(a, a) (P::DEDICATED_SQR; the NTT has no squarings anyway).
namespace detail {
// odd[j] = u[j]*bi + odd[j+2] for the products at or above index `skip` (index into the ODD-limb vector `a`, i.e. limb
// 1 + j of the row vector), plain carry propagation below; carry-in from CC.CF
template <int n>
ZKP_HD void madc_n_rshift_from(uint32_t* odd, const uint32_t* a, uint32_t bi, int skip) {
#pragma unroll
  for (int j = 0; j < n - 2; j += 2) {
    if (j + 1 >= skip) {
      odd[j] = ptx::madc_lo_cc(a[j], bi, odd[j + 2]);
      odd[j + 1] = ptx::madc_hi_cc(a[j], bi, odd[j + 3]);
    } else {
      odd[j] = ptx::addc_cc(odd[j + 2], 0u);
      odd[j + 1] = ptx::addc_cc(odd[j + 3], 0u);
    }
  }
  if (n - 1 >= skip) {
    odd[n - 2] = ptx::madc_lo_cc(a[n - 2], bi, 0u);
    odd[n - 1] = ptx::madc_hi(a[n - 2], bi, 0u);
  } else {
    odd[n - 2] = ptx::addc_cc(0u, 0u);
    odd[n - 1] = 0u;
  }
}
// acc += a[even j]*bi for j >= skip along one carry chain (lower pairs untouched); carry-out left in CC.CF
template <int n>
ZKP_HD void cmad_n_from(uint32_t* acc, const uint32_t* a, uint32_t bi, int skip) {
  bool started = false;
#pragma unroll
  for (int j = 0; j < n; j += 2) {
    if (j < skip) continue;
    if (!started) {
      acc[j] = ptx::mad_lo_cc(a[j], bi, acc[j]);
      started = true;
    } else {
      acc[j] = ptx::madc_lo_cc(a[j], bi, acc[j]);
    }
    acc[j + 1] = ptx::madc_hi_cc(a[j], bi, acc[j + 1]);
  }
  if (!started) ptx::add_cc(0u, 0u);  // clears CC.CF: the caller's closing addc must see no carry
}

// one row of the squaring: `row` is this row's index i (limbs below i are skipped), u the row vector (n + 1 entries)
template <class P>
ZKP_HD void sqr_row_redc(uint32_t* even, uint32_t* odd, const uint32_t* u, uint32_t bi, const uint32_t* mod, int row) {
  constexpr int n = P::N;
  if (row == 0) {
    mul_n<n>(odd, u + 1, bi);
    mul_n<n>(even, u, bi);
  } else {
    even[0] = ptx::add_cc(even[0], odd[1]);
    madc_n_rshift_from<n>(odd, u + 1, bi, row);
    cmad_n_from<n>(even, u, bi, row);
    odd[n - 1] = ptx::addc(odd[n - 1], 0u);
  }
  if constexpr (P::LOW_LIMBS_SPECIAL) {
#if defined(__CUDA_ARCH__)
    const uint32_t mi = even[0] * zkp_opaque_minus_one;
#else
    const uint32_t mi = 0u - even[0];
#endif
    const uint32_t lo1 = 0u - mi;
    const uint32_t hi1 = mi - (mi != 0u ? 1u : 0u);
    odd[0] = ptx::add_cc(odd[0], lo1);
    odd[1] = ptx::addc_cc(odd[1], hi1);
#pragma unroll
    for (int j = 2; j < n; j += 2) {
      odd[j] = ptx::madc_lo_cc(mod[j + 1], mi, odd[j]);
      odd[j + 1] = ptx::madc_hi_cc(mod[j + 1], mi, odd[j + 1]);
    }
    even[0] = ptx::add_cc(even[0], mi);
    even[1] = ptx::addc_cc(even[1], 0u);
#pragma unroll
    for (int j = 2; j < n; j += 2) {
      even[j] = ptx::madc_lo_cc(mod[j], mi, even[j]);
      even[j + 1] = ptx::madc_hi_cc(mod[j], mi, even[j + 1]);
    }
    odd[n - 1] = ptx::addc(odd[n - 1], 0u);
  } else {
    const uint32_t mi = even[0] * P::M0;
    cmad_n<n>(odd, mod + 1, mi);
    cmad_n<n>(even, mod, mi);
    odd[n - 1] = ptx::addc(odd[n - 1], 0u);
  }
}
}  // namespace detail

template <class P>
ZKP_HD Fp<P> fp_sqr_chain(const Fp<P>& a) {
  constexpr int n = P::N;
  uint32_t even[n], odd[n];
  detail::ModLimbs<P> M;
  // 2a, limb by limb (the top limb has spare bits: nothing is shifted out)
  uint32_t a2[n];
  a2[0] = a.v[0] << 1;
#pragma unroll
  for (int i = 1; i < n; i++) a2[i] = (a.v[i] << 1) | (a.v[i - 1] >> 31);
#pragma unroll
  for (int i = 0; i < n; i++) {
    // row vector u_i: limb i = a_i, limb i + 1 = a_(i+1) << 1 (a_i's top bit belongs to the skipped part), above = a2
    uint32_t u[n + 1];
#pragma unroll
    for (int j = 0; j < n; j++) u[j] = (j < i) ? 0u : (j == i ? a.v[j] : (j == i + 1 ? (a.v[j] << 1) : a2[j]));
    u[n] = 0;
    if ((i & 1) == 0) detail::sqr_row_redc<P>(even, odd, u, a.v[i], M.m, i);
    else detail::sqr_row_redc<P>(odd, even, u, a.v[i], M.m, i);
  }
  Fp<P> r;
  r.v[0] = ptx::add_cc(even[0], odd[1]);
#pragma unroll
  for (int i = 1; i < n - 1; i++) r.v[i] = ptx::addc_cc(even[i], odd[i + 1]);
  r.v[n - 1] = ptx::addc(even[n - 1], 0u);
  fp_final_sub(r);
  return r;
}

template <class P>
ZKP_HD Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
#if ZKP_PTX_DEVICE || defined(ZKP_FIELD_CHAIN_ON_HOST)
#if defined(ZKP_KARATSUBA_FQ)
  if constexpr (P::N == 12) return fp_mul_kara(a, b);
#endif
#if defined(ZKP_KARATSUBA_FR)
  if constexpr (P::N == 8) return fp_mul_kara(a, b);
#endif
  return fp_mul_chain(a, b);
#elif defined(ZKP_HOST_MUL64) && !defined(ZKP_FIELD_PORTABLE)
  return fp_mul_host64(a, b);
#else
  return fp_mul_portable(a, b);
#endif
}

template <class P>
ZKP_HD Fp<P> fp_sqr(const Fp<P>& a) {
#if (ZKP_PTX_DEVICE || defined(ZKP_FIELD_CHAIN_ON_HOST)) && !defined(ZKP_NO_DEDICATED_SQR)
  if constexpr (P::DEDICATED_SQR) return fp_sqr_chain(a);
#endif
  return fp_mul(a, a);
}

template <class P>
ZKP_HD Fp<P> fp_to_mont(const Fp<P>& a) { return fp_mul(a, Fp<P>::r2()); }

template <class P>
ZKP_HD Fp<P> fp_from_mont(const Fp<P>& a) {
  Fp<P> o = Fp<P>::zero();
  o.v[0] = 1;
  return fp_mul(a, o);
}

// a^(p-2); host-side finishing only (one call per MSM), so a plain square-and-multiply ladder.
template <class P>
ZKP_HD_NOINLINE Fp<P> fp_inv(const Fp<P>& a) {
  constexpr int N = P::N;
  uint32_t e[N];
  // p - 2 with borrow propagation (r ends in ...00000001)
  uint32_t borrow = 2;
  for (int i = 0; i < N; i++) {
    uint32_t m = P::mod(i);
    e[i] = m - borrow;
    borrow = (m < borrow) ? 1u : 0u;
  }
  Fp<P> r = Fp<P>::one();
  for (int i = N * 32 - 1; i >= 0; i--) {
    r = fp_sqr(r);
    if ((e[i >> 5] >> (i & 31)) & 1) r = fp_mul(r, a);
  }
  return r;
}

template <class P>
ZKP_HD_NOINLINE Fp<P> fp_pow_u64(const Fp<P>& a, uint64_t e) {
  Fp<P> r = Fp<P>::one();
  int top = 63;
  while (top >= 0 && !((e >> top) & 1)) top--;  // nothing to square above the leading bit
  for (int i = top; i >= 0; i--) {
    r = fp_sqr(r);
    if ((e >> i) & 1) r = fp_mul(r, a);
  }
  return r;
}

using Fr = Fp<FrParams>;
using Fq = Fp<FqParams>;

ZKP_HD Fr operator+(const Fr& a, const Fr& b) { return fp_add(a, b); }
ZKP_HD Fr operator-(const Fr& a, const Fr& b) { return fp_sub(a, b); }
ZKP_HD Fr operator*(const Fr& a, const Fr& b) { return fp_mul(a, b); }
ZKP_HD Fq operator+(const Fq& a, const Fq& b) { return fp_add(a, b); }
ZKP_HD Fq operator-(const Fq& a, const Fq& b) { return fp_sub(a, b); }
// -DZKP_FQ_CALL (per translation unit): Fq products of the curve formulas become calls of ONE out-of-line multiplier /
// squarer instead of ~4.5 KB of inlined SASS each.  A point addition is then ~1 KB of code around a 10 KB callee that
// stays in the instruction cache, instead of 50-100 KB of straight-line code that a small, latency-bound launch fetches
// cold (the reduction levels and short accumulations of 2^16..2^20-point MSMs).  Arguments travel in registers.
#if defined(ZKP_FQ_CALL) && ZKP_PTX_DEVICE && defined(__CUDACC__)
static __device__ __noinline__ Fq fq_mul_call(Fq a, Fq b) { return fp_mul(a, b); }
static __device__ __noinline__ Fq fq_sqr_call(Fq a) { return fp_sqr(a); }
ZKP_HD Fq operator*(const Fq& a, const Fq& b) { return fq_mul_call(a, b); }
ZKP_HD Fq fq_sqr(const Fq& a) { return fq_sqr_call(a); }
#else
ZKP_HD Fq operator*(const Fq& a, const Fq& b) { return fp_mul(a, b); }
ZKP_HD Fq fq_sqr(const Fq& a) { return fp_sqr(a); }
#endif

}  // namespace zkp
