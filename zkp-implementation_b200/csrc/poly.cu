// Device-resident polynomial arithmetic over Fr: the O(n) work the reference's PLONK prover does on the CPU
// between its MSMs and FFTs (plonk/src/prover.rs: DensePolynomial add / scale / evaluate / divide by a
// linear factor, the grand-product loop of compute_acc :302-377, the quotient of
// compute_quotient_polynomial :381-444), kept in HBM so a proof never round-trips polynomials through the
// host.  Everything here is exact field arithmetic: results are the same field elements the reference
// computes, whatever the schedule.
//
//   fr_powers        out[i] = first * base^i
//   fr_batch_inverse Montgomery's trick, one inversion per thread-owned run
//   fr_scan          inclusive prefix / suffix scan under * or +   (grand product; synthetic division)
//   fr_lincomb       out = sum_k coef_k * poly_k (+ constant)       (linearisation polynomial, opening numerators)
//   fr_eval          p(x) by chunked Horner + tree                  (bar_a ... bar_z_w, r(zeta))
//   fr_trimmed_len   DensePolynomial::from_coefficients_vec's trailing-zero trim
//   plonk_*          the two pointwise kernels of rounds 2 and 3
//   g1_scalar_mul    k * P for a handful of scalars                 (KzgScheme::commit_para, scheme.rs:78-82)
#include <string.h>

#include "engine.h"
#include "memops.cuh"
#include "poly.h"

namespace zkp {

#ifdef ZKP_EMU
static constexpr uint32_t PT = 64;   // emulated build: one OS thread per CUDA thread, keep the blocks small
#else
static constexpr uint32_t PT = 256;
#endif
static constexpr uint32_t POW_CHUNK = 64;
static constexpr uint32_t INV_CHUNK = 16;
static constexpr uint32_t SCAN_PER_THREAD = 8;
static constexpr uint32_t SCAN_BLOCK = PT * SCAN_PER_THREAD;
static constexpr uint32_t EV_CHUNK = 32;

static unsigned blocks_for(size_t items, unsigned per_block) { return (unsigned)((items + per_block - 1) / per_block); }

// ---- powers ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT) fr_powers_kernel(Fr* out, Fr base, Fr first, size_t n) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t start = t * POW_CHUNK;
  if (start >= n) return;
  Fr cur = fp_mul(first, fp_pow_u64(base, (uint64_t)start));
  for (uint32_t i = 0; i < POW_CHUNK && start + i < n; i++) {
    st_fr(out + start + i, cur);
    cur = fp_mul(cur, base);
  }
}

int fr_powers_dev(Ctx* ctx, Fr* out, const Fr& base, const Fr& first, size_t n) {
  if (!n) return ZKP_OK;
  ZKP_LAUNCH_NOSYNC(fr_powers_kernel, dim3(blocks_for(n, PT * POW_CHUNK)), dim3(PT), 0, ctx->stream, out, base, first, n);
  return rt::check_last();
}

// ---- batch inversion ------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT) fr_batch_inverse_kernel(Fr* data, size_t n) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t start = t * INV_CHUNK;
  if (start >= n) return;
  const uint32_t cnt = (uint32_t)((n - start < INV_CHUNK) ? (n - start) : INV_CHUNK);
  Fr pre[INV_CHUNK];
  Fr prod = Fr::one();
  for (uint32_t i = 0; i < cnt; i++) {
    pre[i] = prod;
    prod = fp_mul(prod, ld_fr(data + start + i));
  }
  Fr inv = fr_inv_gcd(prod);  // division steps (inv_gcd.cuh); a zero entry zeroes its whole run (the reference panics on 1/0)
  for (uint32_t i = cnt; i-- > 0;) {
    const Fr v = ld_fr(data + start + i);
    st_fr(data + start + i, fp_mul(inv, pre[i]));
    inv = fp_mul(inv, v);
  }
}

int fr_batch_inverse_dev(Ctx* ctx, Fr* data, size_t n) {
  if (!n) return ZKP_OK;
  ZKP_LAUNCH_NOSYNC(fr_batch_inverse_kernel, dim3(blocks_for(n, PT * INV_CHUNK)), dim3(PT), 0, ctx->stream, data, n);
  return rt::check_last();
}

// ---- scans ----------------------------------------------------------------------------------------
template <int OP>
__device__ __forceinline__ Fr scan_op(const Fr& a, const Fr& b) {
  if (OP == 0) return fp_mul(a, b);
  return fp_add(a, b);
}
template <int OP>
__device__ __forceinline__ Fr scan_identity() {
  if (OP == 0) return Fr::one();
  return Fr::zero();
}

// Inclusive scan of one block of SCAN_BLOCK elements (in scan order); block totals go to `tot`.
template <int OP>
__global__ void __launch_bounds__(PT) fr_scan_block_kernel(Fr* data, size_t n, Fr* tot, uint32_t reverse) {
  __shared__ __align__(16) Fr sh[2][PT];
  const uint32_t tid = threadIdx.x;
  const size_t first = ((size_t)blockIdx.x * PT + tid) * SCAN_PER_THREAD;  // position in scan order
  Fr v[SCAN_PER_THREAD];
  Fr run = scan_identity<OP>();
#pragma unroll
  for (uint32_t i = 0; i < SCAN_PER_THREAD; i++) {
    const size_t pos = first + i;
    if (pos < n) {
      run = scan_op<OP>(run, ld_fr(data + (reverse ? n - 1 - pos : pos)));
    }
    v[i] = run;
  }
  st_fr(&sh[0][tid], run);
  __syncthreads();
  uint32_t cur = 0;
  for (uint32_t d = 1; d < PT; d <<= 1) {
    Fr x = ld_fr(&sh[cur][tid]);
    if (tid >= d) x = scan_op<OP>(ld_fr(&sh[cur][tid - d]), x);
    st_fr(&sh[cur ^ 1][tid], x);
    cur ^= 1;
    __syncthreads();
  }
  // sh[cur][tid] = inclusive scan of the thread totals
  if (tid > 0) {
    const Fr pre = ld_fr(&sh[cur][tid - 1]);
#pragma unroll
    for (uint32_t i = 0; i < SCAN_PER_THREAD; i++) v[i] = scan_op<OP>(pre, v[i]);
  }
#pragma unroll
  for (uint32_t i = 0; i < SCAN_PER_THREAD; i++) {
    const size_t pos = first + i;
    if (pos < n) st_fr(data + (reverse ? n - 1 - pos : pos), v[i]);
  }
  if (tid == PT - 1 && tot) st_fr(tot + blockIdx.x, ld_fr(&sh[cur][PT - 1]));
}

// data[block b >= 1] = op(scanned_tot[b - 1], data)
template <int OP>
__global__ void __launch_bounds__(PT) fr_scan_apply_kernel(Fr* data, size_t n, const Fr* tot_scanned, uint32_t reverse) {
  const uint32_t b = blockIdx.x + 1;
  const Fr pre = ld_fr(tot_scanned + (b - 1));
  const size_t first = ((size_t)b * PT + threadIdx.x) * SCAN_PER_THREAD;
#pragma unroll
  for (uint32_t i = 0; i < SCAN_PER_THREAD; i++) {
    const size_t pos = first + i;
    if (pos < n) {
      Fr* p = data + (reverse ? n - 1 - pos : pos);
      st_fr(p, scan_op<OP>(pre, ld_fr(p)));
    }
  }
}

template <int OP>
static int fr_scan_rec(Ctx* ctx, Fr* data, size_t n, bool reverse, Fr* scratch) {
  const unsigned nb = blocks_for(n, SCAN_BLOCK);
  ZKP_LAUNCH(fr_scan_block_kernel<OP>, dim3(nb), dim3(PT), 0, ctx->stream, data, n, nb > 1 ? scratch : (Fr*)nullptr,
             reverse ? 1u : 0u);
  if (nb > 1) {
    ZKP_TRY(fr_scan_rec<OP>(ctx, scratch, nb, false, scratch + nb));
    ZKP_LAUNCH_NOSYNC(fr_scan_apply_kernel<OP>, dim3(nb - 1), dim3(PT), 0, ctx->stream, data, n, scratch, reverse ? 1u : 0u);
  }
  return rt::check_last();
}

int fr_scan_dev(Ctx* ctx, Fr* data, size_t n, int op, bool reverse) {
  if (!n) return ZKP_OK;
  // scratch for the block totals of every level: n / 2048 + n / 2048^2 + ... + slack
  ZKP_TRY(ctx->poly_scratch.reserve((n / SCAN_BLOCK + 64) * 2 * sizeof(Fr)));
  Fr* scratch = ctx->poly_scratch.as<Fr>();
  return op == 0 ? fr_scan_rec<0>(ctx, data, n, reverse, scratch) : fr_scan_rec<1>(ctx, data, n, reverse, scratch);
}

// ---- linear combination ---------------------------------------------------------------------------
__global__ void __launch_bounds__(PT) fr_lincomb_kernel(Fr* out, size_t out_len, LincombArgs a) {
  size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; j < out_len; j += stride) {
    Fr acc = (j == 0 && a.has_c0) ? a.c0 : Fr::zero();
    for (uint32_t k = 0; k < a.count; k++)
      if (j < a.len[k]) acc = fp_add(acc, fp_mul(a.coef[k], ld_fr(a.p[k] + j)));
    st_fr(out + j, acc);
  }
}

int fr_lincomb_dev(Ctx* ctx, Fr* out, size_t out_len, const LincombArgs& a) {
  if (!out_len) return ZKP_OK;
  if (a.count > LincombArgs::MAX_TERMS) return ZKP_ERR_INVALID_ARG;
  unsigned blocks = blocks_for(out_len, PT);
  const unsigned cap = (unsigned)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  ZKP_LAUNCH_NOSYNC(fr_lincomb_kernel, dim3(blocks), dim3(PT), 0, ctx->stream, out, out_len, a);
  return rt::check_last();
}

// data[idx[k]] += val[k], sequentially (a handful of blinding terms)
__global__ void fr_add_at_kernel(Fr* data, SparseAddArgs a) {
  if (blockIdx.x || threadIdx.x) return;
  for (uint32_t k = 0; k < a.count; k++) st_fr(data + a.idx[k], fp_add(ld_fr(data + a.idx[k]), a.val[k]));
}

int fr_add_at_dev(Ctx* ctx, Fr* data, const SparseAddArgs& a) {
  if (!a.count) return ZKP_OK;
  if (a.count > SparseAddArgs::MAX_TERMS) return ZKP_ERR_INVALID_ARG;
  ZKP_LAUNCH_NOSYNC(fr_add_at_kernel, dim3(1), dim3(32), 0, ctx->stream, data, a);
  return rt::check_last();
}

// ---- evaluation -----------------------------------------------------------------------------------
// All evaluations of one call (the prover's seven openings at zeta / zeta omega) run as ONE launch pair: block (bx, k)
// handles chunk bx of polynomial k.  partial[k][bx] = sum over the block's threads of (Horner of 32 coefficients) *
// x^(first index); thread-level cost = 32 products + one exponentiation over the bits of the index.
struct EvalBatch {
  const Fr* c[Ctx::EVAL_SLOTS];
  size_t n[Ctx::EVAL_SLOTS];
  Fr x[Ctx::EVAL_SLOTS];
};

__global__ void __launch_bounds__(PT) fr_eval_partial_kernel(EvalBatch b, Fr* partial) {
  __shared__ __align__(16) Fr sh[PT];
  const uint32_t tid = threadIdx.x, k = blockIdx.y;
  const size_t n = b.n[k];
  if ((size_t)blockIdx.x * (PT * EV_CHUNK) >= n) return;  // block-uniform: shorter polynomial of the batch
  const Fr* c = b.c[k];
  const Fr x = b.x[k];
  const size_t t = (size_t)blockIdx.x * blockDim.x + tid;
  const size_t start = t * EV_CHUNK;
  Fr acc = Fr::zero();
  if (start < n) {
    const uint32_t cnt = (uint32_t)((n - start < EV_CHUNK) ? (n - start) : EV_CHUNK);
    for (uint32_t i = cnt; i-- > 0;) acc = fp_add(fp_mul(acc, x), ld_fr(c + start + i));
    acc = fp_mul(acc, fp_pow_u64(x, (uint64_t)start));
  }
  st_fr(&sh[tid], acc);
  __syncthreads();
  for (uint32_t s = PT / 2; s > 0; s >>= 1) {
    if (tid < s) st_fr(&sh[tid], fp_add(ld_fr(&sh[tid]), ld_fr(&sh[tid + s])));
    __syncthreads();
  }
  if (tid == 0) st_fr(partial + (size_t)k * Ctx::EVAL_PARTIALS + blockIdx.x, ld_fr(&sh[0]));
}

// out[k] = sum of polynomial k's block partials (0 for an empty polynomial)
__global__ void __launch_bounds__(PT) fr_sum_kernel(const Fr* partial, EvalBatch b, Fr* out) {
  __shared__ __align__(16) Fr sh[PT];
  const uint32_t tid = threadIdx.x, k = blockIdx.x;
  const size_t nb = (b.n[k] + (size_t)PT * EV_CHUNK - 1) / ((size_t)PT * EV_CHUNK);
  const Fr* in = partial + (size_t)k * Ctx::EVAL_PARTIALS;
  Fr acc = Fr::zero();
  for (size_t i = tid; i < nb; i += PT) acc = fp_add(acc, ld_fr(in + i));
  st_fr(&sh[tid], acc);
  __syncthreads();
  for (uint32_t s = PT / 2; s > 0; s >>= 1) {
    if (tid < s) st_fr(&sh[tid], fp_add(ld_fr(&sh[tid]), ld_fr(&sh[tid + s])));
    __syncthreads();
  }
  if (tid == 0) st_fr(out + k, ld_fr(&sh[0]));
}

// Evaluate `count` polynomials (device coefficients, natural order) at their points; the values land in the context's
// result area, slot k, and are read back in bulk by fr_eval_fetch.
int fr_eval_batch_dev(Ctx* ctx, uint32_t count, const Fr* const* coeffs, const size_t* lens, const Fr* xs) {
  if (count > Ctx::EVAL_SLOTS) return ZKP_ERR_INVALID_ARG;
  if (!count) return ZKP_OK;
  ZKP_TRY(ctx->eval_out.reserve(Ctx::EVAL_SLOTS * sizeof(Fr)));
  ZKP_TRY(ctx->eval_partials.reserve(Ctx::EVAL_SLOTS * Ctx::EVAL_PARTIALS * sizeof(Fr)));
  EvalBatch b;
  memset(&b, 0, sizeof(b));
  unsigned nb_max = 1;
  for (uint32_t k = 0; k < count; k++) {
    const unsigned nb = lens[k] ? blocks_for(lens[k], PT * EV_CHUNK) : 0u;
    if (nb > Ctx::EVAL_PARTIALS) return ZKP_ERR_INVALID_ARG;
    if (nb > nb_max) nb_max = nb;
    b.c[k] = coeffs[k];
    b.n[k] = lens[k];
    b.x[k] = xs[k];
  }
  ZKP_LAUNCH(fr_eval_partial_kernel, dim3(nb_max, count), dim3(PT), 0, ctx->stream, b, ctx->eval_partials.as<Fr>());
  ZKP_LAUNCH(fr_sum_kernel, dim3(count), dim3(PT), 0, ctx->stream, (const Fr*)ctx->eval_partials.as<Fr>(), b, ctx->eval_out.as<Fr>());
  return rt::check_last();
}

int fr_eval_fetch(Ctx* ctx, Fr* out_host, uint32_t count) {
  if (count > Ctx::EVAL_SLOTS) return ZKP_ERR_INVALID_ARG;
  ZKP_TRY(ctx->eval_out.reserve(Ctx::EVAL_SLOTS * sizeof(Fr)));
  ZKP_TRY(rt::d2h(out_host, ctx->eval_out.p, count * sizeof(Fr), ctx->stream));
  return rt::sync(ctx->stream);
}

// ---- trimmed length -------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT) fr_trimmed_len_kernel(const Fr* c, size_t n, unsigned* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned best = 0;
  for (; i < n; i += stride)
    if (!ld_fr(c + i).is_zero()) best = (unsigned)i + 1;
  if (best) atomicMax(out, best);
}

int fr_trimmed_len_dev(Ctx* ctx, const Fr* coeffs, size_t n, size_t* out_len) {
  *out_len = 0;
  if (!n) return ZKP_OK;
  if (n >= ((size_t)1 << 32)) return ZKP_ERR_INVALID_ARG;
  ZKP_TRY(ctx->eval_out.reserve(Ctx::EVAL_SLOTS * sizeof(Fr)));
  unsigned* d = reinterpret_cast<unsigned*>(ctx->eval_out.as<Fr>() + (Ctx::EVAL_SLOTS - 1));  // last slot doubles as the counter
  ZKP_TRY(rt::dev_memset(d, 0, sizeof(unsigned), ctx->stream));
  unsigned blocks = blocks_for(n, PT);
  const unsigned cap = (unsigned)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  ZKP_LAUNCH_NOSYNC(fr_trimmed_len_kernel, dim3(blocks), dim3(PT), 0, ctx->stream, coeffs, n, d);
  unsigned h = 0;
  ZKP_TRY(rt::d2h(&h, d, sizeof(unsigned), ctx->stream));
  ZKP_TRY(rt::sync(ctx->stream));
  *out_len = h;
  return rt::check_last();
}

// ---- PLONK round 2: factors of the grand product (prover.rs:314-369) -----------------------------
//   num[i] = (a_i + beta w^i + gamma)(b_i + beta k1 w^i + gamma)(c_i + beta k2 w^i + gamma)
//   den[i] = (a_i + beta s1_i + gamma)(b_i + beta s2_i + gamma)(c_i + beta s3_i + gamma)
// a, b, c are the wire VALUES on the domain (what f_a, f_b, f_c evaluate to at w^i), s1..s3 the sigma values.
__global__ void __launch_bounds__(PT) plonk_numden_kernel(PlonkNumDenArgs p) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < p.n; i += stride) {
    const Fr a = fp_add(ld_fr(p.a + i), p.gamma), b = fp_add(ld_fr(p.b + i), p.gamma), c = fp_add(ld_fr(p.c + i), p.gamma);
    const Fr w = ld_fr(p.roots + i);
    Fr num = fp_add(a, fp_mul(p.beta, w));
    num = fp_mul(num, fp_add(b, fp_mul(p.beta_k1, w)));
    num = fp_mul(num, fp_add(c, fp_mul(p.beta_k2, w)));
    Fr den = fp_add(a, fp_mul(p.beta, ld_fr(p.s1 + i)));
    den = fp_mul(den, fp_add(b, fp_mul(p.beta, ld_fr(p.s2 + i))));
    den = fp_mul(den, fp_add(c, fp_mul(p.beta, ld_fr(p.s3 + i))));
    st_fr(p.num + i, num);
    st_fr(p.den + i, den);
  }
}

int plonk_numden_dev(Ctx* ctx, const PlonkNumDenArgs& p) {
  if (!p.n) return ZKP_OK;
  unsigned blocks = blocks_for(p.n, PT);
  const unsigned cap = (unsigned)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  ZKP_LAUNCH_NOSYNC(plonk_numden_kernel, dim3(blocks), dim3(PT), 0, ctx->stream, p);
  return rt::check_last();
}

// ---- PLONK round 3: the quotient on a coset (prover.rs:381-444) ------------------------------------
// t(x) = [ a b q_m + a q_l + b q_r + c q_o + pi + q_c
//          + alpha ( (a + beta x + gamma)(b + beta k1 x + gamma)(c + beta k2 x + gamma) z(x)
//                  - (a + beta s1 + gamma)(b + beta s2 + gamma)(c + beta s3 + gamma) z(w x) )
//          + alpha^2 (z(x) - 1) L1(x) ] / Z_H(x)
// at the D = rho * n points x_i = h eta^i (eta = omega_D): z(w x_i) is z at index i + rho, and Z_H(x_i)
// takes rho distinct values.  The three lines are each divisible by Z_H for a satisfied circuit; the
// host checks that on the domain itself (gate equation, grand product) as the reference's `expect`s do.
__global__ void __launch_bounds__(PT) plonk_quotient_kernel(PlonkQuotientArgs p) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t dmask = p.d - 1;
  for (; i < p.d; i += stride) {
    const Fr a = ld_fr(p.a + i), b = ld_fr(p.b + i), c = ld_fr(p.c + i), z = ld_fr(p.z + i);
    const Fr zw = ld_fr(p.z + ((i + p.rho) & dmask));
    const Fr x = ld_fr(p.x + i);
    Fr gate = fp_mul(fp_mul(a, b), ld_fr(p.qm + i));
    gate = fp_add(gate, fp_mul(a, ld_fr(p.ql + i)));
    gate = fp_add(gate, fp_mul(b, ld_fr(p.qr + i)));
    gate = fp_add(gate, fp_mul(c, ld_fr(p.qo + i)));
    gate = fp_add(gate, fp_add(ld_fr(p.pi + i), ld_fr(p.qc + i)));
    const Fr ag = fp_add(a, p.gamma), bg = fp_add(b, p.gamma), cg = fp_add(c, p.gamma);
    Fr p2 = fp_add(ag, fp_mul(p.beta, x));
    p2 = fp_mul(p2, fp_add(bg, fp_mul(p.beta_k1, x)));
    p2 = fp_mul(p2, fp_add(cg, fp_mul(p.beta_k2, x)));
    p2 = fp_mul(p2, z);
    Fr p3 = fp_add(ag, fp_mul(p.beta, ld_fr(p.s1 + i)));
    p3 = fp_mul(p3, fp_add(bg, fp_mul(p.beta, ld_fr(p.s2 + i))));
    p3 = fp_mul(p3, fp_add(cg, fp_mul(p.beta, ld_fr(p.s3 + i))));
    p3 = fp_mul(p3, zw);
    const Fr l4 = fp_mul(fp_sub(z, Fr::one()), ld_fr(p.l1 + i));
    Fr t = fp_add(gate, fp_mul(p.alpha, fp_sub(p2, p3)));
    t = fp_add(t, fp_mul(p.alpha2, l4));
    st_fr(p.t + i, fp_mul(t, p.zh_inv[i & (p.rho - 1)]));
  }
}

int plonk_quotient_dev(Ctx* ctx, const PlonkQuotientArgs& p) {
  if (!p.d) return ZKP_OK;
  if (p.rho > 8 || (p.rho & (p.rho - 1)) || (p.d & (p.d - 1))) return ZKP_ERR_INVALID_ARG;
  unsigned blocks = blocks_for(p.d, PT);
  const unsigned cap = (unsigned)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  ZKP_LAUNCH_NOSYNC(plonk_quotient_kernel, dim3(blocks), dim3(PT), 0, ctx->stream, p);
  return rt::check_last();
}

// Gate equation on the domain rows: flag[0] |= (q_l a + q_r b + q_o c + q_m a b + q_c + pi != 0)
__global__ void __launch_bounds__(PT) plonk_gate_check_kernel(const Fr* a, const Fr* b, const Fr* c, const Fr* ql,
                                                              const Fr* qr, const Fr* qo, const Fr* qm, const Fr* qc,
                                                              const Fr* pi, size_t n, unsigned* flag) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned bad = 0;
  for (; i < n; i += stride) {
    const Fr av = ld_fr(a + i), bv = ld_fr(b + i);
    Fr g = fp_mul(fp_mul(av, bv), ld_fr(qm + i));
    g = fp_add(g, fp_mul(av, ld_fr(ql + i)));
    g = fp_add(g, fp_mul(bv, ld_fr(qr + i)));
    g = fp_add(g, fp_mul(ld_fr(c + i), ld_fr(qo + i)));
    g = fp_add(g, fp_add(ld_fr(qc + i), ld_fr(pi + i)));
    if (!g.is_zero()) bad = 1;
  }
  if (bad) atomicMax(flag, 1u);
}

int plonk_gate_check_dev(Ctx* ctx, const Fr* const cols[9], size_t n, bool* ok) {
  *ok = true;
  if (!n) return ZKP_OK;
  ZKP_TRY(ctx->eval_out.reserve(Ctx::EVAL_SLOTS * sizeof(Fr)));
  unsigned* d = reinterpret_cast<unsigned*>(ctx->eval_out.as<Fr>() + (Ctx::EVAL_SLOTS - 1));
  ZKP_TRY(rt::dev_memset(d, 0, sizeof(unsigned), ctx->stream));
  unsigned blocks = blocks_for(n, PT);
  const unsigned cap = (unsigned)ctx->sm_count * 8;
  if (blocks > cap) blocks = cap;
  ZKP_LAUNCH_NOSYNC(plonk_gate_check_kernel, dim3(blocks), dim3(PT), 0, ctx->stream, cols[0], cols[1], cols[2], cols[3], cols[4],
             cols[5], cols[6], cols[7], cols[8], n, d);
  unsigned h = 0;
  ZKP_TRY(rt::d2h(&h, d, sizeof(unsigned), ctx->stream));
  ZKP_TRY(rt::sync(ctx->stream));
  *ok = (h == 0);
  return rt::check_last();
}

// ---- k * P for a few scalars (commit_para) --------------------------------------------------------
// The base is always srs[0]: a 32 x 256 window table d * 2^(8w) * P (built once per SRS) turns k * P into 32
// table entries, one per byte of the canonical scalar, summed by a warp.
static constexpr uint32_t PTAB_WINDOWS = 32, PTAB_DIGITS = 256;

__global__ void g1_base_table_kernel(const G1Affine* base, G1Xyzz* tab) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= PTAB_WINDOWS) return;
  G1Xyzz b = G1Xyzz::from_affine(ld_affine(base));
  for (uint32_t k = 0; k < 8 * w; k++) b = xyzz_dbl(b);
  G1Xyzz acc = G1Xyzz::infinity();
  for (uint32_t d = 0; d < PTAB_DIGITS; d++) {
    st_xyzz(tab + (size_t)w * PTAB_DIGITS + d, acc);
    xyzz_add(acc, b);
  }
}

// one block of 32 threads per scalar: lane w contributes table[w][byte w of k]
__global__ void __launch_bounds__(32) g1_scalar_mul_kernel(const G1Affine* __restrict__ tab, const Fr* scalars, uint32_t count,
                                                           G1Xyzz* out) {
  __shared__ G1Xyzz sh[32];
  const uint32_t t = blockIdx.x, w = threadIdx.x;
  if (t >= count) return;
  const Fr k = fp_from_mont(ld_fr(scalars + t));  // `into_bigint()` inside ark-ec's scalar mul
  const uint32_t d = (k.v[w >> 2] >> ((w & 3) * 8)) & 0xffu;
  G1Xyzz acc = G1Xyzz::infinity();
  if (d) acc = G1Xyzz::from_affine(ld_affine(tab + (size_t)w * PTAB_DIGITS + d));
  st_xyzz(&sh[w], acc);
  __syncthreads();
  for (uint32_t s = 16; s > 0; s >>= 1) {
    if (w < s) {
      G1Xyzz a = ld_xyzz(&sh[w]), b = ld_xyzz(&sh[w + s]);
      xyzz_add(a, b);
      st_xyzz(&sh[w], a);
    }
    __syncthreads();
  }
  if (w == 0) st_xyzz(out + t, ld_xyzz(&sh[0]));
}

int normalise_dev(Ctx* ctx, const G1Xyzz* tmp, size_t n, G1Affine* out);  // gen.cu

// scalars[k] * srs[0] for count <= 64 scalars (host in, host XYZZ out)
int g1_scalar_mul_dev(Ctx* ctx, const Fr* scalars_host, uint32_t count, G1Xyzz* out_host) {
  if (!count) return ZKP_OK;
  if (count > 64 || ctx->srs_len == 0) return ZKP_ERR_INVALID_ARG;
  const size_t tab_n = (size_t)PTAB_WINDOWS * PTAB_DIGITS;
  if (!ctx->srs0_tab) {  // first commit_para against this SRS
    DevBuf tmp;
    ZKP_TRY(tmp.reserve(tab_n * sizeof(G1Xyzz)));
    int st = rt::dev_malloc((void**)&ctx->srs0_tab, tab_n * sizeof(G1Affine));
    if (st == ZKP_OK) {
      ZKP_LAUNCH_NOSYNC(g1_base_table_kernel, dim3(1), dim3(PTAB_WINDOWS), 0, ctx->stream, (const G1Affine*)ctx->srs,
                        tmp.as<G1Xyzz>());
      st = normalise_dev(ctx, tmp.as<G1Xyzz>(), tab_n, ctx->srs0_tab);
    }
    if (st == ZKP_OK) st = rt::sync(ctx->stream);
    tmp.release();
    if (st != ZKP_OK) {
      rt::dev_free(ctx->srs0_tab);
      ctx->srs0_tab = nullptr;
      return st;
    }
  }
  ZKP_TRY(ctx->eval_partials.reserve(Ctx::EVAL_SLOTS * Ctx::EVAL_PARTIALS * sizeof(Fr)));
  Fr* ds = ctx->eval_partials.as<Fr>();
  G1Xyzz* dout = reinterpret_cast<G1Xyzz*>(ds + 64);
  ZKP_TRY(rt::h2d(ds, scalars_host, count * sizeof(Fr), ctx->stream));
  ZKP_LAUNCH(g1_scalar_mul_kernel, dim3(count), dim3(32), 0, ctx->stream, (const G1Affine*)ctx->srs0_tab, (const Fr*)ds, count,
             dout);
  ZKP_TRY(rt::d2h(out_host, dout, count * sizeof(G1Xyzz), ctx->stream));
  ZKP_TRY(rt::sync(ctx->stream));
  return rt::check_last();
}

}  // namespace zkp
