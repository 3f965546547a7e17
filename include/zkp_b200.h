/* zkp_b200.h -- C ABI of the B200-native polynomial-commitment engine.
 *
 * The reference (sota-zk-labs/zkp-implementation) is pure Rust on arkworks and has no FFI; the
 * drop-in seam is created at two existing function boundaries and every entry
 * point below names the reference code it replaces.  INTEGRATION.md shows the Rust `extern "C"`
 * block and the shim a maintainer adds to the `kzg` / `plonk` crates.
 *
 * Conventions
 *   - Field elements cross the boundary exactly as arkworks stores them: little-endian u64 limbs
 *     of the MONTGOMERY representation (Fr: 4 limbs, R = 2^256; Fq: 6 limbs, R = 2^384), so the
 *     shim passes `fr.0.0` / `pt.x.0.0` without conversion.
 *   - A G1 affine point is x || y (12 u64).  The point at infinity is the all-zero encoding
 *     (0, 0) -- not on the curve -- optionally accompanied by an `infinity` byte array mirroring
 *     ark-ec's `Affine { x, y, infinity: bool }`.
 *   - Every function returns 0 on success or a ZKP_ERR_* code; nothing throws or aborts.  The Rust
 *     shim maps non-zero to `panic!`, which is the reference's own error convention
 *     (kzg/src/scheme.rs:86 `assert!`, :112 `expect`).
 *   - Pointers are HOST memory unless the name ends in `_dev`.  The caller owns its buffers for the
 *     duration of the call; outputs are written only on success.
 *   - A context is bound to one GPU and one CUDA stream; calls on one context are serialised
 *     (KzgScheme is Send + Sync in the reference).  Use one context per GPU / per process.
 *   - There is NO CPU fallback: without a usable sm_100 device `zkp_ctx_create` fails.
 */
#ifndef ZKP_B200_H
#define ZKP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zkp_ctx zkp_ctx;

enum {
  ZKP_B200_OK = 0,
  ZKP_B200_ERR_INVALID_ARG = 1,
  ZKP_B200_ERR_CUDA = 2,
  ZKP_B200_ERR_OOM = 3,
  ZKP_B200_ERR_SRS_TOO_SMALL = 4,    /* kzg/src/scheme.rs:86  assert!(g1_points.len() > polynomial.degree()) */
  ZKP_B200_ERR_DOMAIN_TOO_LARGE = 5, /* ark-poly GeneralEvaluationDomain::new(..) == None (plonk/src/prover.rs:70) */
  ZKP_B200_ERR_NO_DEVICE = 6,
  ZKP_B200_ERR_EMPTY_POLY = 7        /* kzg/src/scheme.rs:112 expect("at least 1") */
};

/* ---- context ------------------------------------------------------------------------------- */
/* Created once next to `KzgScheme::new(srs)` (kzg/src/scheme.rs:34). */
int zkp_ctx_create(zkp_ctx** out, int device);
/* Single-process multi-GPU context: the one `KzgScheme` value of the reference (kzg/src/scheme.rs:34, `&self` at :49-52)
 * backed by `n_devices` GPUs, no torch / MPI in between.  The context runs everything on device_ids[0] except the
 * hot path of the kzg crate: the resident SRS (zkp_srs_upload / zkp_srs_generate[_range]) is split by point range over
 * the devices, zkp_srs_precompute builds one window table per device, and every commitment (zkp_msm_g1, zkp_msm_g1_dev
 * and zkp_msm_g1_multi_dev with bases = resident SRS) runs one Pippenger per device on its point range, concurrently
 * (one host thread per device), with the 192-byte partial sums folded on the host.  Device-resident scalars live on
 * device_ids[0] and are copied to the peers over NVLink.  A device id may repeat (several shards on one GPU). */
int zkp_ctx_create_multi(zkp_ctx** out, const int* device_ids, int n_devices);
int zkp_ctx_shards(const zkp_ctx* ctx); /* 1 for a plain context */
void zkp_ctx_destroy(zkp_ctx* ctx);
/* Run all work of this context on an existing CUDA stream (cudaStream_t passed as void*). */
int zkp_ctx_set_stream(zkp_ctx* ctx, void* cuda_stream);
/* Block until everything queued on the context's stream has finished. */
int zkp_ctx_synchronize(zkp_ctx* ctx);
/* Pippenger window width in bits (0 = pick from n). */
int zkp_ctx_set_msm_window(zkp_ctx* ctx, uint32_t bits);
/* Batched-affine tree rounds of the bucket accumulation (csrc/msm_affine.cu) before the XYZZ finish: -1 = pick from the
 * mean bucket load (default), 0 = XYZZ only, R > 0 = exactly R rounds.  Any setting returns the same point. */
int zkp_ctx_set_msm_affine(zkp_ctx* ctx, int rounds);
/* Kernel launches issued by the last call of the given kind (0 = MSM, 1 = NTT); kind 2 / 3 return
 * the window width c / window count W the last MSM used, kind 4 its number of batched-affine rounds. */
int zkp_ctx_last_launches(zkp_ctx* ctx, int kind);
/* Per-phase device timing of the last MSM (CUDA events on the context's stream, recorded only when
 * profiling is on).  phase: 0 recode, 1 sort, 2 bucket boundaries, 3 accumulate (batched-affine rounds + task list +
 * XYZZ finish), 4 bucket/window reduction.  Returns milliseconds, or a negative value if nothing was recorded.
 * phase 100: the first batched-affine round's addition kernel alone (the longest launch of a large MSM);
 * phase 200 + r: points entering round 1 (r = 0) / leaving round r (counts, not times). */
int zkp_ctx_set_profiling(zkp_ctx* ctx, int on);
double zkp_ctx_last_phase_ms(zkp_ctx* ctx, int phase);
const char* zkp_strerror(int status);

/* ---- SRS: replaces the per-call `self.0.g1_points()` clone (kzg/src/srs.rs:78-80 at
 *      kzg/src/scheme.rs:85) with bases uploaded once and kept resident in HBM. ---------------- */
int zkp_srs_upload(zkp_ctx* ctx, const uint64_t* xy /* n x 12 */, const uint8_t* infinity /* n or NULL */, size_t n);
/* Adopt `n` affine points already in device memory (copied device-to-device). */
int zkp_srs_upload_dev(zkp_ctx* ctx, const void* xy_dev, size_t n);
size_t zkp_srs_len(const zkp_ctx* ctx);
/* Fixed-base table over the resident SRS (called once next to `KzgScheme::new`): for every window w of the
 * signed-digit recoding the table holds 2^(c w) * srs[i], so all windows of an MSM share one bucket set.
 * window_bits = 0 picks c from the SRS length.  Costs (255 / c + 1) x the SRS in HBM (16.5 GiB at 2^24, c = 24)
 * and one pass of c doublings per entry; zkp_msm_g1 / zkp_msm_g1_dev with the SRS then use it automatically.
 * Uploading or generating a new SRS drops the table. */
int zkp_srs_precompute(zkp_ctx* ctx, uint32_t window_bits);
/* `Srs::new_from_secret` (kzg/src/srs.rs:48-69): fill the resident SRS with [secret^i * G], i < n,
 * computed on the GPU; optionally copy the points back to `xy_out` (n x 12 u64, may be NULL). */
int zkp_srs_generate(zkp_ctx* ctx, const uint64_t secret[4], size_t n, uint64_t* xy_out);
/* The point range [first, first + n) of that SRS: the shard of one rank of a point-range-sharded SRS. */
int zkp_srs_generate_range(zkp_ctx* ctx, const uint64_t secret[4], size_t first, size_t n, uint64_t* xy_out);

/* ---- MSM: replaces `KzgScheme::evaluate_in_s` (kzg/src/scheme.rs:84-96) ---------------------
 * out = sum_{i<n} scalars[i] * srs[i], normalised affine; n == 0 -> infinity (scheme.rs:94).
 * n > zkp_srs_len -> ZKP_B200_ERR_SRS_TOO_SMALL. */
int zkp_msm_g1(zkp_ctx* ctx, const uint64_t* scalars /* n x 4 */, size_t n, uint64_t out_xy[12], uint8_t* out_infinity);
/* Same with ad-hoc bases (benchmark config 2: random points, not an SRS). */
int zkp_msm_g1_bases(zkp_ctx* ctx, const uint64_t* scalars, const uint64_t* xy, const uint8_t* infinity, size_t n,
                     uint64_t out_xy[12], uint8_t* out_infinity);
/* Device-resident operands: scalars (n x 32 B) and, if non-NULL, bases (n x 96 B) already in HBM;
 * bases_dev == NULL uses the resident SRS. */
int zkp_msm_g1_dev(zkp_ctx* ctx, const void* scalars_dev, const void* bases_dev, size_t n, uint64_t out_xy[12],
                   uint8_t* out_infinity);
/* `count` <= 16 commitments against the resident SRS in ONE pipeline (the prover's a/b/c, t_lo/t_mid/t_hi and the
 * two opening commitments, plonk/src/prover.rs:577-579, slice_polynomial.rs:52): with the fixed-base table every
 * MSM of the batch is one more bucket set of the same sort / accumulate / reduce launches.  out_xy: count x 12. */
int zkp_msm_g1_multi_dev(zkp_ctx* ctx, uint32_t count, const void* const* scalars_dev, const size_t* lens, uint64_t* out_xy,
                         uint8_t* out_infinity /* count, or NULL */);
/* The same batch as un-normalised partial sums (count x 24 u64, XYZZ) over this rank's SRS shard: the sharded prover
 * all-gathers them and folds per commitment with zkp_g1_fold_partials. */
int zkp_msm_g1_multi_partial_dev(zkp_ctx* ctx, uint32_t count, const void* const* scalars_dev, const size_t* lens,
                                 uint64_t* out_xyzz);
/* Multi-GPU point-range sharding: each rank computes the un-normalised partial sum of its shard
 * (XYZZ coordinates, 4 x 6 u64) ...                                                              */
int zkp_msm_g1_partial_dev(zkp_ctx* ctx, const void* scalars_dev, const void* bases_dev, size_t n, uint64_t out_xyzz[24]);
/* ... the partials are exchanged (NCCL all-gather of 192-byte records) and folded on the host. */
int zkp_g1_fold_partials(const uint64_t* partials_xyzz /* count x 24 */, size_t count, uint64_t out_xy[12],
                         uint8_t* out_infinity);

/* ---- NTT: replaces ark-poly `EvaluationDomain::<Fr>::{fft_in_place, ifft_in_place}` and their
 *      coset forms, reached via `Evaluations::interpolate()` (plonk/src/prover.rs:374-375,463;
 *      plonk/src/circuit.rs:175,230-232) and `&DensePolynomial * &DensePolynomial`
 *      (plonk/src/prover.rs:396-437,516-548).  Natural order in and out; `inverse` includes N^-1;
 *      coset_offset (4 u64, Montgomery) may be NULL for the plain domain. -------------------- */
int zkp_ntt_fr(zkp_ctx* ctx, uint64_t* data /* batch x 2^log_n x 4, in place */, uint32_t log_n, size_t batch, int inverse,
               const uint64_t* coset_offset);
int zkp_ntt_fr_dev(zkp_ctx* ctx, void* data_dev, uint32_t log_n, size_t batch, int inverse, const uint64_t* coset_offset);
/* `&a * &b` for DensePolynomial (2 NTT + pointwise + iNTT on the device, no host round trip).
 * out must hold la + lb - 1 coefficients; la == 0 or lb == 0 writes nothing (zero polynomial). */
int zkp_poly_mul_fr(zkp_ctx* ctx, const uint64_t* a, size_t la, const uint64_t* b, size_t lb, uint64_t* out);
/* a[i] *= b[i] on device vectors (the pointwise step of a product kept resident). */
int zkp_fr_mul_pointwise_dev(zkp_ctx* ctx, void* a_dev, const void* b_dev, size_t n);

/* ---- `KzgScheme::open` / `open_vector` (kzg/src/scheme.rs:108-120, 132-142): y = p(z), witness = commit((p - y) / (X - z))
 *      against the resident SRS.  coeffs: n x 4 u64 (trailing zeros are trimmed as DensePolynomial does); an empty
 *      polynomial returns ZKP_B200_ERR_EMPTY_POLY (scheme.rs:112 expect("at least 1")).  Evaluation, division and the
 *      commitment all run on the device. ------------------------------------------------------------------------- */
int zkp_kzg_open(zkp_ctx* ctx, const uint64_t* coeffs, size_t n, const uint64_t z[4], uint64_t out_xy[12],
                 uint8_t* out_infinity, uint64_t out_y[4]);

/* ---- multi-GPU four-step NTT (one process per GPU).  N = 2^r x 2^(log_n - r),
 *      r = zkp_ntt_dist_rows_log.  Rank g of G holds
 *        layout A:  a[j1][c]  = x[j1 * 2^(log_n-r) + g * 2^(log_n-r)/G + c]   (its column block, row-major)
 *        layout B:  b[k'][k2] = X[(g * 2^r/G + k') + 2^r * k2]                (its rows of the result)
 *      forward: A -> stage (column transforms + omega_N^(col*k) twiddles) -> all-to-all -> local batched
 *      transforms of size 2^(log_n-r) (zkp_ntt_fr_dev, batch 2^r/G) -> B.   inverse: B -> A, the exact
 *      reverse.  The all-to-all is either NCCL (stage in place, then zkp_ntt_dist_permute_dev around the
 *      collective) or fused into the stage kernel: with `peer_bufs` (G device pointers, one exchange buffer
 *      of N/G elements per rank, opened with the zkp_ipc_* calls) the forward stage stores each row
 *      straight into its owner's buffer over NVLink and the inverse stage loads from them. ------------- */
uint32_t zkp_ntt_dist_rows_log(uint32_t log_n, uint32_t world);
int zkp_ntt_dist_stage_dev(zkp_ctx* ctx, void* data_dev, uint32_t log_n, uint32_t rank, uint32_t world, int inverse,
                           const uint64_t* coset_offset, void* const* peer_bufs);
int zkp_ntt_dist_permute_dev(zkp_ctx* ctx, const void* in_dev, void* out_dev, uint32_t log_n, uint32_t world, int inverse);
/* Exchange buffers: plain device allocations whose CUDA IPC handle (64 bytes) is handed to the peer processes. */
int zkp_dev_alloc(zkp_ctx* ctx, size_t bytes, void** out_dev);
int zkp_dev_free(zkp_ctx* ctx, void* dev);
int zkp_dev_copy(zkp_ctx* ctx, void* dst_dev, const void* src_dev, size_t bytes); /* device-to-device, on the context's stream */
int zkp_ipc_export(zkp_ctx* ctx, const void* dev, uint8_t handle[64]);
int zkp_ipc_open(zkp_ctx* ctx, const uint8_t handle[64], void** out_dev);
int zkp_ipc_close(zkp_ctx* ctx, void* dev);

/* ---- device-resident Fr vectors: the O(n) polynomial work the reference's prover does on the CPU between
 *      its MSMs and FFTs (plonk/src/prover.rs: DensePolynomial add / scale / evaluate / divide), kept in HBM.
 *      Elements are 32-byte Montgomery Fr; `*_dev` pointers come from zkp_dev_alloc. --------------------- */
int zkp_dev_upload(zkp_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);   /* returns after the copy */
int zkp_dev_download(zkp_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes); /* returns after the copy */
int zkp_dev_zero(zkp_ctx* ctx, void* dst_dev, size_t bytes);
/* out[i] = first * base^i   (`domain.elements()`, coset points, powers of an opening point) */
int zkp_fr_powers_dev(zkp_ctx* ctx, void* out_dev, const uint64_t base[4], const uint64_t first[4], size_t n);
/* data[i] = 1 / data[i] (Montgomery's trick; a zero entry poisons its run of 16 -- the reference panics on 1/0) */
int zkp_fr_batch_inverse_dev(zkp_ctx* ctx, void* data_dev, size_t n);
/* in-place inclusive scan; op 0 = product, 1 = sum; reverse != 0 scans from the top index down */
int zkp_fr_scan_dev(zkp_ctx* ctx, void* data_dev, size_t n, int op, int reverse);
/* out[j] = sum_k coefs[k] * polys[k][j] (j < lens[k])  + c0 at j == 0; count <= 12; out may alias no input */
int zkp_fr_lincomb_dev(zkp_ctx* ctx, void* out_dev, size_t out_len, uint32_t count, const void* const* polys_dev,
                       const size_t* lens, const uint64_t* coefs /* count x 4 */, const uint64_t* c0 /* 4 or NULL */);
/* data[idx[k]] += vals[k], k < count <= 8, applied in order (blinding terms, constant-term edits) */
int zkp_fr_add_at_dev(zkp_ctx* ctx, void* data_dev, uint32_t count, const size_t* idx, const uint64_t* vals);
/* out[k] = polys[k](xs[k])  (`Polynomial::evaluate`), count <= 15, one read-back for the whole batch */
int zkp_fr_eval_dev(zkp_ctx* ctx, uint32_t count, const void* const* polys_dev, const size_t* lens, const uint64_t* xs,
                    uint64_t* out);
/* number of coefficients left after DensePolynomial's trailing-zero trim */
int zkp_fr_trimmed_len_dev(zkp_ctx* ctx, const void* coeffs_dev, size_t n, size_t* out_len);
/* `KzgScheme::commit_para` (kzg/src/scheme.rs:78-82) for `count` <= 64 scalars at once: scalars[k] * srs[0] */
int zkp_g1_mul_srs0(zkp_ctx* ctx, const uint64_t* scalars, uint32_t count, uint64_t* out_xy /* count x 12 */);

/* ---- the MSM's grouping step on its own: stable LSD radix sort of n (key, value) u32 pairs by the low key_bits of
 *      the key (descending != 0: largest first), in place on device arrays; and the exclusive u32 scan it uses. ---- */
int zkp_sort_pairs_dev(zkp_ctx* ctx, void* keys_dev, void* vals_dev, size_t n, uint32_t key_bits, int descending);
int zkp_scan_exclusive_u32_dev(zkp_ctx* ctx, const void* in_dev, void* out_dev, size_t n);

/* ---- synthetic workloads (bench configs 2/5): n distinct pseudo-random G1 points generated on
 *      the device from a seed (a0 + i*delta) * G, affine, written to bases_dev (n x 96 B). ----- */
int zkp_g1_generate_bases_dev(zkp_ctx* ctx, uint64_t seed, size_t n, void* bases_dev);
/* points [first, first + n) of the same progression: a rank's shard of a point set that does not depend on the number
 * of ranks (config 5: the sharded 2^26 MSM computes the same point at 1, 2, 4 and 8 GPUs). */
int zkp_g1_generate_bases_range_dev(zkp_ctx* ctx, uint64_t seed, size_t first, size_t n, void* bases_dev);

/* ---- integer-pipe microbenchmark: the MSM roofline denominator.
 *      Runs independent 32x32->64 multiply-add chains on every SM and returns multiply-adds/s. */
int zkp_bench_imad_peak(zkp_ctx* ctx, double* wide_madds_per_s, double* lo_madds_per_s);

#ifdef __cplusplus
}
#endif
#endif /* ZKP_B200_H */
