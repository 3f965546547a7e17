/* zkp_plonk.h -- host orchestration of the reference's PLONK prover over the GPU engine.
 *
 * Mirrors the `plonk` crate's builder + prover for the rows of SURVEY.md section 8 that CALL the hot path
 * (a7-a11): `Circuit::{add_*_gate, compile}` (plonk/src/circuit.rs:85-197) and
 * `prover::generate_proof` (plonk/src/prover.rs:61-293).  Every G1 sum goes through `zkp_msm_g1`,
 * every interpolation / polynomial product through `zkp_ntt_fr` / `zkp_poly_mul_fr`
 * (include/zkp_b200.h); the O(n) scalar work between them (add, scale, Horner, division by a
 * linear factor, the vanishing-polynomial fold) runs on the host as it does in the reference.
 *
 * Differences from the reference, all output-preserving (a proof is a function of circuit, SRS and
 * blinding scalars only -- field and group arithmetic are exact):
 *   - the nine blinding scalars b1..b9 (`StdRng::from_entropy()`, prover.rs:68) are an input;
 *   - `compute_acc` (prover.rs:302-377) reads the wire / sigma evaluations kept from `compile`
 *     instead of Horner-evaluating nine degree-n polynomials at each of n points (O(n^2));
 *   - the always-on self-check products at prover.rs:515-553 are not recomputed.
 * Where the reference panics the functions return a status (the Rust shim would panic! on it).
 */
#ifndef ZKP_PLONK_H
#define ZKP_PLONK_H

#include <stddef.h>
#include <stdint.h>

#include "zkp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
  ZKP_PLONK_ERR_REMAINDER = 20,        /* prover.rs:404/431/441 expect("No remainder ..."): unsatisfied circuit */
  ZKP_PLONK_ERR_INVALID_POSITION = 21, /* circuit.rs:221 panic!("Invalid position") */
  ZKP_PLONK_ERR_TOO_FEW_GATES = 22,    /* circuit.rs:151 (len - 1).ilog2() on 0 or 1 gates */
  ZKP_PLONK_ERR_TRANSCRIPT = 23        /* challenge.rs:61-63 */
};

enum { ZKP_PLONK_GATE_ADD = 0, ZKP_PLONK_GATE_MUL = 1, ZKP_PLONK_GATE_CONST = 2 };

typedef struct zkp_plonk_circuit zkp_plonk_circuit;   /* plonk::circuit::Circuit */
typedef struct zkp_plonk_compiled zkp_plonk_compiled; /* plonk::compiled_circuit::CompiledCircuit */

/* plonk::prover::Proof (prover.rs:24-58): commitments in the order a, b, c, z, t_lo, t_mid, t_hi,
 * w_ev_x, w_ev_wx (affine x || y Montgomery, (0,0) = identity); evaluations bar_a, bar_b, bar_c,
 * bar_s_sigma_1, bar_s_sigma_2, bar_z_w; u; degree. */
typedef struct zkp_plonk_proof {
  uint64_t commitments[9][12];
  uint64_t evaluations[6][4];
  uint64_t u[4];
  uint64_t degree;
} zkp_plonk_proof;

zkp_plonk_circuit* zkp_plonk_circuit_new(void);               /* Circuit::default() */
void zkp_plonk_circuit_free(zkp_plonk_circuit* c);
/* add_addition_gate / add_multiplication_gate / add_constant_gate in bulk (circuit.rs:85-115).
 * positions: count x 6 = (column, row) of wires a, b, c; values: count x 3 Fr (Montgomery); pis: count Fr. */
int zkp_plonk_circuit_add_gates(zkp_plonk_circuit* c, size_t count, const uint8_t* kinds, const uint64_t* positions,
                                const uint64_t* values, const uint64_t* pis);
size_t zkp_plonk_circuit_len(const zkp_plonk_circuit* c);

/* Circuit::compile (circuit.rs:166-197): pad to a power of two, 12 interpolations (one batched iNTT). */
int zkp_plonk_compile(zkp_ctx* ctx, const zkp_plonk_circuit* c, zkp_plonk_compiled** out);
/* Teardown order: a compiled circuit owns device buffers allocated through `ctx`; free it BEFORE zkp_ctx_destroy(ctx)
 * (freeing it afterwards would touch a destroyed context). */
void zkp_plonk_compiled_free(zkp_plonk_compiled* cc);
size_t zkp_plonk_compiled_size(const zkp_plonk_compiled* cc);
/* Coefficients of one compiled polynomial, zero-padded to `size` (which: 0..8 = f_a f_b f_c q_l q_r q_o q_m q_c pi,
 * 9..11 = s_sigma_1..3). */
int zkp_plonk_compiled_poly(const zkp_plonk_compiled* cc, int which, uint64_t* out /* size x 4 */);

/* The verifier's preprocessed commitments -- `get_circuit_commitment` (verifier.rs:160-185, recomputed by every
 * `verify`: 8 MSMs of size n) and `CommonPreprocessedInput::new` (common_preprocessed_input/cpi_parser.rs:76-106):
 * commit(q_m), commit(q_l), commit(q_r), commit(q_o), commit(q_c), commit(s_sigma_1..3), in that order, as ONE batched
 * MSM pipeline over the coefficient vectors that `zkp_plonk_compile` left in HBM.  The result is cached on the compiled
 * circuit; refresh != 0 (or a resident SRS of a different length) recomputes it -- pass refresh after replacing the SRS. */
int zkp_plonk_preprocess(zkp_ctx* ctx, zkp_plonk_compiled* cc, uint64_t out_xy[8][12], int refresh);

/* prover::generate_proof (prover.rs:61-293) against the SRS resident in `ctx` (>= size + 3 points).
 * blinding: b1..b9 (9 x 4 u64, Montgomery).  timings_ms (may be NULL; when given, the stream is synchronised at the
 * phase boundaries so the split is exact): [0] total, [1] commitments (MSM), [2] transforms (NTT / products),
 * [3] the rest (pointwise kernels, scans, evaluations, transcript). */
int zkp_plonk_prove(zkp_ctx* ctx, const zkp_plonk_compiled* cc, const uint64_t* blinding, zkp_plonk_proof* out,
                    double* timings_ms);

/* The same proof with the nine commitments point-range-sharded over `world` GPUs, one process each:
 * the SRS resident in `ctx` is [srs_first, srs_first + zkp_srs_len) of the global SRS of `srs_total` points
 * (zkp_srs_generate_range / zkp_srs_upload of the shard, then zkp_srs_precompute); every rank runs the whole prover
 * -- the transforms and pointwise kernels are a small part of a proof -- but only its range of each MSM, and the
 * 192-byte partial sums go through `allgather` (recv = world x bytes, rank-major; NCCL / gloo behind it) and are
 * folded on every rank.  All ranks return the same proof, byte-identical to zkp_plonk_prove's. */
typedef int (*zkp_allgather_fn)(void* user, const void* send, size_t bytes, void* recv);
int zkp_plonk_prove_sharded(zkp_ctx* ctx, const zkp_plonk_compiled* cc, const uint64_t* blinding, zkp_plonk_proof* out,
                            double* timings_ms, uint32_t rank, uint32_t world, size_t srs_first, size_t srs_total,
                            zkp_allgather_fn allgather, void* user);

/* Same proof, computed the way prover.rs is written: every `&a * &b` of compute_quotient_polynomial as its own
 * GPU product (zkp_poly_mul_fr: 2 NTT + pointwise + iNTT at 4n / 8n), polynomials held on the host between
 * calls.  Kept as an independent cross-check of zkp_plonk_prove (the proofs must be byte-identical). */
int zkp_plonk_prove_products(zkp_ctx* ctx, const zkp_plonk_compiled* cc, const uint64_t* blinding, zkp_plonk_proof* out,
                             double* timings_ms);

/* zkp_plonk_prove_products with `compute_acc` exactly as prover.rs:302-377 writes it: nine Horner evaluations of
 * degree-n polynomials and one field division per row (O(n^2)).  Same bytes again; it exists so that the reference's
 * own algorithm can be timed step for step (bench.py's host-CPU PLONK baseline links this file against a CPU backend
 * of the C ABI, oracle/cpu_backend.cpp) and as a third cross-check at small n. */
int zkp_plonk_prove_reference(zkp_ctx* ctx, const zkp_plonk_compiled* cc, const uint64_t* blinding, zkp_plonk_proof* out,
                              double* timings_ms);

/* ---- the two pointwise kernels of the device-resident prover (csrc/poly.cu) -------------------------- */
/* prover.rs:314-369: num[i], den[i] of the grand-product factor at row i, from the wire / sigma VALUES on the
 * domain and roots[i] = omega^i. */
typedef struct zkp_plonk_numden_args {
  const void *a_dev, *b_dev, *c_dev, *s1_dev, *s2_dev, *s3_dev, *roots_dev;
  uint64_t beta[4], gamma[4], k1[4], k2[4];
  size_t n;
  void *num_dev, *den_dev;
} zkp_plonk_numden_args;
int zkp_plonk_numden_dev(zkp_ctx* ctx, const zkp_plonk_numden_args* args);

/* prover.rs:381-444: t = (line1 + line2 - line3 + line4) / Z_H evaluated pointwise on the coset
 * x_i = h * omega_d^i, d = rho * n (rho = 4, or 8 when 3n + 6 > 4n); every input is the coset evaluation of
 * the named polynomial, z(omega x_i) is read at index i + rho, zh_inv[i mod rho] = 1 / (x_i^n - 1). */
typedef struct zkp_plonk_quotient_args {
  const void *a_dev, *b_dev, *c_dev, *z_dev;
  const void *ql_dev, *qr_dev, *qo_dev, *qm_dev, *qc_dev, *pi_dev, *s1_dev, *s2_dev, *s3_dev, *l1_dev, *x_dev;
  uint64_t beta[4], gamma[4], alpha[4], k1[4], k2[4];
  uint64_t zh_inv[8][4];
  size_t d;
  uint32_t rho;
  void* t_dev;
} zkp_plonk_quotient_args;
int zkp_plonk_quotient_dev(zkp_ctx* ctx, const zkp_plonk_quotient_args* args);

/* q_l a + q_r b + q_o c + q_m a b + q_c + pi == 0 on every row (cols: a b c q_l q_r q_o q_m q_c pi VALUES):
 * what makes prover.rs:404 `expect("No remainder 1")` hold. */
int zkp_plonk_gate_check_dev(zkp_ctx* ctx, const void* const cols_dev[9], size_t n, int* ok);

/* ---- Fiat-Shamir transcript pieces (plonk/src/challenge.rs:49-89), exposed so that known-answer tests can pin each
 * third-party semantic the reference relies on against PUBLIC vectors (tests/test_transcript_kat.py): FIPS 180-4
 * SHA-256, the PCG32 output function + rand_core 0.6 `seed_from_u64`, the ChaCha block function / word order / 64-bit
 * counter of rand 0.8's StdRng (double_rounds = 6; 10 = ChaCha20 for RFC 8439 vectors), ark-bls12-381's uncompressed
 * G1 encoding, and the whole `feed* -> generate_challenges` chain. ---- */
void zkp_transcript_sha256(const uint8_t* data, size_t n, uint8_t out[32]);
uint32_t zkp_transcript_pcg32_output(uint64_t state);
void zkp_transcript_seed_from_u64(uint64_t seed, uint32_t key_out[8]);
/* the first `count` 32-bit words of StdRng::from_seed(key) (key = eight little-endian words of the 32-byte seed) */
void zkp_transcript_chacha_words(const uint32_t key[8], int double_rounds, size_t count, uint32_t* out);
void zkp_transcript_g1_serialize(const uint64_t xy[12], uint8_t out[96]);
/* ChallengeGenerator: feed(points[0..k)) then generate_challenges::<n>() -> n Fr (Montgomery); != 0 where it panics */
int zkp_transcript_challenges(const uint64_t* points, size_t k, size_t n, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* ZKP_PLONK_H */
