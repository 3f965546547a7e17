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
void zkp_plonk_compiled_free(zkp_plonk_compiled* cc);
size_t zkp_plonk_compiled_size(const zkp_plonk_compiled* cc);
/* Coefficients of one compiled polynomial, zero-padded to `size` (which: 0..8 = f_a f_b f_c q_l q_r q_o q_m q_c pi,
 * 9..11 = s_sigma_1..3). */
int zkp_plonk_compiled_poly(const zkp_plonk_compiled* cc, int which, uint64_t* out /* size x 4 */);

/* prover::generate_proof (prover.rs:61-293) against the SRS resident in `ctx` (>= size + 3 points).
 * blinding: b1..b9 (9 x 4 u64, Montgomery).  timings_ms (may be NULL): [0] total, [1] MSM calls,
 * [2] NTT / poly-product calls, [3] host arithmetic. */
int zkp_plonk_prove(zkp_ctx* ctx, const zkp_plonk_compiled* cc, const uint64_t* blinding, zkp_plonk_proof* out,
                    double* timings_ms);

#ifdef __cplusplus
}
#endif
#endif /* ZKP_PLONK_H */
