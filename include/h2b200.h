/*
 * h2b200 -- B200-native (sm_100a) replacement for the data-parallel hot path of the Halo2 KZG
 * prover as DCMMC/halo2-scaffold exercises it: BN254 G1 multi-scalar multiplication and the
 * radix-2 NTT over BN254 Fr.  Plain C ABI: pointers and sizes only, no C++ types, no exceptions.
 *
 * What each entry point replaces (the arithmetic lives in un-vendored git dependencies of the
 * reference, marked [UP]; the reference's own call sites are given as file:line under
 * /root/reference):
 *
 *   h2b_msm_bn254_g1   <- [UP] halo2_proofs::arithmetic::best_multiexp::<G1Affine>
 *                         (halo2_proofs/src/arithmetic.rs, PSE tag v2023_02_02 -- Cargo.toml:13 --
 *                          and the Axiom fork `axiom/dev` -- Cargo.toml:16), reached through
 *                         ParamsKZG::commit / commit_lagrange from keygen_vk / keygen_pk
 *                         (src/scaffold.rs:132,135,146,149,284,287,298,301), create_proof
 *                         (src/scaffold.rs:191-199,207-214,322-330,338-346;
 *                          examples/standard_plonk.rs:41-49) and verify_proof
 *                         (src/scaffold.rs:223-230,354-361; examples/standard_plonk.rs:57-64).
 *   h2b_ntt_bn254_fr   <- [UP] halo2_proofs::arithmetic::best_fft::<Fr>, reached through
 *                         EvaluationDomain::{lagrange_to_coeff, coeff_to_extended,
 *                         extended_to_coeff} from keygen_pk (src/scaffold.rs:135,149,287,301) and
 *                         create_proof (same lines as above).
 *
 * Data layouts are exactly halo2curves 0.3.x's in-memory layouts, so Rust slices can be passed
 * by pointer cast:
 *   Fr / Fq   : 4 x u64 little-endian limbs, Montgomery form (R = 2^256), fully reduced.
 *   G1Affine  : x | y (8 x u64); the identity is (0, 0).
 *   G1        : x | y | z Jacobian (12 x u64); z == 0 is the identity.
 *
 * Errors: every function returns 0 on success and a negative H2B_ERR_* code otherwise;
 * h2b_last_error() returns a thread-local description.  The upstream Rust functions are
 * infallible (they assert/panic), so the Rust shim turns a non-zero return into a panic.
 * Threading: all entry points may be called concurrently from several host threads; host-pointer
 * calls are synchronous (they return after the result is in caller memory).
 * There is no CPU fallback: without a usable CUDA device h2b_init fails.
 */
#ifndef H2B200_H
#define H2B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define H2B_OK 0
#define H2B_ERR_NOT_INITIALIZED (-1)
#define H2B_ERR_BAD_ARGUMENT (-2)
#define H2B_ERR_CUDA (-3)
#define H2B_ERR_OOM (-4)
#define H2B_ERR_NO_DEVICE (-5)
#define H2B_ERR_BAD_HANDLE (-6)

/* ---- lifecycle ------------------------------------------------------------------------------- */
/* Use CUDA devices 0..n_devices-1 (0 = every visible device).  Idempotent. */
int h2b_init(int n_devices);
/* Use exactly one device, by CUDA ordinal (one-process-per-GPU launches: pass LOCAL_RANK). */
int h2b_init_device(int device);
void h2b_shutdown(void);
int h2b_device_count(void);
const char* h2b_last_error(void);
const char* h2b_version(void);
/* 1 if this binary is the CPU kernel-logic emulator build used by unit tests, 0 for the CUDA build. */
int h2b_is_emulator(void);

/* ---- drop-in entry points (host pointers, synchronous) ------------------------------------------- */
/* best_multiexp(coeffs, bases): out_jac = sum_i scalars[i] * bases[i].  n may be any value >= 0
 * (n == 0 gives the identity).  A pure function of its arguments, like upstream's:
 *  - arrays of fewer than 4096 points, or whose length is not a multiple of 1024 points (verifier MSMs,
 *    odd prefixes), are uploaded for the call and never cached;
 *  - longer arrays -- in a prover only the SRS vectors `g` and `g_lagrange` -- are kept
 *    resident (window tables from the second use on) and found again by CONTENT: a 128-bit digest of
 *    every 64 KiB block is taken at upload time, and a resident copy is used for a call only if every block
 *    of the caller's array still matches (host threads verify while the GPU runs the MSM; a mismatch discards
 *    the result, re-uploads and recomputes).  In-place edits and address reuse can therefore not return stale
 *    results, and the same vector re-loaded at another address (src/scaffold.rs:174) is recognised.
 * With more than one device the resident copy is sharded by point range (device d holds rows
 * [d n/D, (d+1) n/D)), every device runs a complete Pippenger on its slice and the partial sums are
 * added on device 0.  H2B_IMPLICIT_CACHE=0 disables the cache. */
int h2b_msm_bn254_g1(const uint64_t* scalars, const uint64_t* bases, size_t n, uint64_t out_jac[12]);

/* best_fft(a, omega, log_n) for G = Fr: in place, natural order in and out, no scaling;
 * a has 2^log_n elements, log_n <= 28.
 * With D = 2^j > 1 devices and log_n >= 22 (H2B_NTT_MULTI_MIN_LOG; H2B_NTT_MULTI=0 turns it off) the ONE transform is split over
 * the devices as a four-step NTT (SURVEY.md 8e "one NTT across GPUs"): device d uploads the column block [d C/D, (d+1) C/D) of the
 * R x C matrix straight from `a` (D PCIe links in parallel), runs its column transforms and the twiddle step, the devices exchange
 * n/D^2-element blocks by 2-D peer copies over NVLink (the only exchange; no collective), and every device runs its row transforms
 * and writes its part of the result into `a`. */
int h2b_ntt_bn254_fr(uint64_t* a, const uint64_t omega[4], uint32_t log_n);

/* Explicit device residency for an SRS vector (copied; the host array may be freed afterwards).  Registration
 * also precomputes the window tables T_j[i] = 2^(spacing*j) * bases[i] on the device (one-off cost of a few MSMs;
 * memory n_tables x n x 64 B, see h2b_base_set_info) so that later MSMs over the set need no doublings and a
 * single shared bucket set.  ParamsKZG's `g` and `g_lagrange` ([UP] halo2_proofs/src/poly/kzg/commitment.rs) are
 * the two vectors a prover registers.  Implicitly cached arrays (h2b_msm_bn254_g1) get their tables on second use. */
int h2b_register_bases(const uint64_t* bases, size_t n, uint64_t* handle);
/* The same with the rows split over the devices of h2b_init: device d keeps rows [d n/D, (d+1) n/D) and their tables
 * (1/D of the memory and of the registration time; table spacing chosen for n/D points).  MSMs over such a set are split
 * by point range over the devices that hold the rows; with one device it equals
 * h2b_register_bases. */
int h2b_register_bases_sharded(const uint64_t* bases, size_t n, uint64_t* handle);
/* Drops one reference; calls still running on the set finish on it (device memory is released afterwards). */
int h2b_unregister_bases(uint64_t handle);
/* MSM over bases[offset .. offset+n) of a registered set. */
int h2b_msm_bn254_g1_registered(const uint64_t* scalars, uint64_t handle, size_t offset, size_t n, uint64_t out_jac[12]);

/* ---- batched drop-ins: independent columns / polynomials, distributed round-robin over the devices ------------------
 * The advice, lookup and permutation columns a proof phase commits are independent MSMs over the same SRS vector, and
 * its per-polynomial (i)NTTs are independent too: with D devices, D columns are in flight at
 * once, each on a device that holds its own copy of the SRS tables.  Results are identical to `count` single calls. */
int h2b_msm_bn254_g1_batch_registered(const uint64_t* const* scalars, const size_t* lens, size_t count, uint64_t handle,
                                      uint64_t* out_jac /* count x 12 */);
/* Columns that land on the same device and are short enough to be latency bound (<= 2^21 scalars) share ONE
 * decompose / sort / accumulate / reduce sequence (column id in the bucket index; up to 32 columns and 2^23 scalars per
 * group): the examples the reference ships run k = 16..20, where a single MSM is dominated by dependent steps. */
int h2b_ntt_bn254_fr_batch(uint64_t* const* a, size_t count, const uint64_t omega[4], uint32_t log_n);

/* ---- device-resident entry points (device pointers on `device`, caller's CUDA stream) ------------- */
/* Asynchronous with respect to the host: work is enqueued on `stream` (a cudaStream_t; NULL = the
 * legacy default stream).  `device` is an index into the devices given to h2b_init.
 * Stream contract: the *_dev entry points of one device share that device's grow-only scratch buffers (sort lists, NTT work
 * buffer, scan / lookup scratch), so calls on the SAME device must be ordered with respect to each other -- issue them on one
 * stream (what a prover thread does), or order the streams with events.  Different devices are independent.  (The compiled
 * programs of evaluate_h are the exception: they rotate through a small ring guarded by events.) */
int h2b_ntt_bn254_fr_dev(int device, void* d_a, const uint64_t omega[4], uint32_t log_n, void* stream);
/* `count` device-resident polynomials of 2^log_n elements each (d_polys: host array of device pointers), same omega: groups of up
 * to 16 polynomials share every pass launch (polynomial index in blockIdx.y) -- the per-column (i)NTTs of a proof phase. */
int h2b_ntt_bn254_fr_dev_batch(int device, void* const* d_polys, size_t count, const uint64_t omega[4], uint32_t log_n, void* stream);
int h2b_msm_bn254_g1_dev(int device, const void* d_scalars, const void* d_bases, size_t n, void* d_out_jac /* 96 B */, void* stream);
/* Point-range sharding across processes (one process per GPU): each rank computes a partial
 * result block (224 B: Jacobian x|y|z followed by the XYZZ form) for its slice, the blocks are gathered by the
 * caller (e.g. torch.distributed.all_gather) and folded on one device. */
int h2b_msm_bn254_g1_dev_partial(int device, const void* d_scalars, const void* d_bases, size_t n, void* d_out_block /* 224 B */, void* stream);
/* same over rows [offset, offset + n) of a registered base set (uses its window tables); d_scalars on `device` */
int h2b_msm_bn254_g1_dev_registered(int device, const void* d_scalars, uint64_t handle, size_t offset, size_t n, void* d_out_block /* 224 B */, void* stream);
/* `count` device-resident columns (d_scalars: host array of device pointers, lens[j] scalars each) over rows [0, lens[j]) of a
 * registered set in one batched kernel sequence; d_out_blocks: count x 224 B.  Asynchronous on `stream`. */
int h2b_msm_bn254_g1_dev_batch_registered(int device, const void* const* d_scalars, const size_t* lens, size_t count, uint64_t handle,
                                          void* d_out_blocks, void* stream);
int h2b_msm_fold_partials(int device, const uint64_t* host_blocks /* count x 28 u64 */, size_t count, uint64_t out_jac[12]);
/* same, blocks and result in device memory, asynchronous on `stream` */
int h2b_msm_fold_partials_dev(int device, const void* d_blocks, size_t count, void* d_out_jac /* 96 B */, void* stream);
/* a[i] *= factors[i % count], 1 <= count <= 4096 (more than 8 factors go through a device table and synchronise the stream once): the 1/n scaling of lagrange_to_coeff / extended_to_coeff (count 1), the
 * (1, zeta, zeta^2) coset pattern of coeff_to_extended (count 3) and EvaluationDomain::divide_by_vanishing_poly, whose
 * t_evaluations repeat with period 2^(extended_k - k) ([UP] halo2_proofs/src/poly/domain.rs). */
int h2b_fr_scale_dev(int device, void* d_a, size_t n, const uint64_t* factors, int count, void* stream);

/* EvaluationDomain's three conversions on a device-resident column ([UP] halo2_proofs/src/poly/domain.rs, SURVEY.md row
 * a6); the domain constants are the caller's (computed once per domain on the host, as EvaluationDomain::new does):
 *   lagrange_to_coeff : best_fft(a, omega_inv, k); a[i] *= ifft_divisor                         (2^k elements, in place)
 *   coeff_to_extended : a[i] *= zeta_powers[i mod 3] for i < 2^k; zero-pad to 2^extended_k; best_fft(a, extended_omega, extended_k)
 *                       (d_a must hold 2^extended_k elements; zeta_powers = 1, zeta, zeta^2)
 *   extended_to_coeff : best_fft(a, extended_omega_inv, extended_k); a[i] *= factors[i mod 3]
 *                       (factors = extended_ifft_divisor * (1, zeta^2, zeta); the caller truncates to n * (j - 1)) */
int h2b_lagrange_to_coeff_dev(int device, void* d_a, uint32_t k, const uint64_t omega_inv[4], const uint64_t ifft_divisor[4], void* stream);
int h2b_coeff_to_extended_dev(int device, void* d_a, uint32_t k, uint32_t extended_k, const uint64_t extended_omega[4], const uint64_t zeta_powers[12],
                              void* stream);
int h2b_extended_to_coeff_dev(int device, void* d_a, uint32_t extended_k, const uint64_t extended_omega_inv[4], const uint64_t factors[12], void* stream);
/* lagrange_to_coeff / coeff_to_extended for `count` columns at once (d_cols: host array of device pointers; every column as in the
 * single-column call): the transform passes and the scalings are launched once per batch of up to 16 columns */
int h2b_lagrange_to_coeff_dev_batch(int device, void* const* d_cols, size_t count, uint32_t k, const uint64_t omega_inv[4], const uint64_t ifft_divisor[4],
                                    void* stream);
int h2b_coeff_to_extended_dev_batch(int device, void* const* d_cols, size_t count, uint32_t k, uint32_t extended_k, const uint64_t extended_omega[4],
                                    const uint64_t zeta_powers[12], void* stream);

/* ONE column through commit_lagrange -> lagrange_to_coeff -> coeff_to_extended with a single upload ([UP]
 * plonk/prover.rs commits every advice / permuted / product column in Lagrange form and then converts it twice): `lagrange` is the
 * host column (2^k x 4 words), handle_g_lagrange a registered copy of ParamsKZG::g_lagrange resident on `device`.  Writes the
 * commitment (Jacobian), and -- each optional, NULL to skip -- the coefficient form (host, 2^k x 4), the extended-coset form to the
 * host (2^extended_k x 4) and/or leaves it on the device (*d_extended, to be released with h2b_dev_free) for h2b_evaluate_*_dev.
 * Synchronous.  Against three host-pointer drop-in calls it saves two uploads and one download of the column. */
int h2b_column_pipeline(int device, const uint64_t* lagrange, uint64_t handle_g_lagrange, uint32_t k, uint32_t extended_k, const uint64_t omega_inv[4],
                        const uint64_t ifft_divisor[4], const uint64_t extended_omega[4], const uint64_t zeta_powers[12], uint64_t out_commitment_jac[12],
                        uint64_t* out_coeff, uint64_t* out_extended, void** d_extended);

/* Grand-product building blocks on a device-resident Fr column ([UP] halo2_proofs
 * plonk/permutation/prover.rs, plonk/lookup/prover.rs): z(omega^i) is the exclusive running product of
 * numerator[i] / denominator[i], the denominators inverted with ff::BatchInvert.
 *   batch_invert   : a[i] <- 1 / a[i] in place; zeros stay zero
 *   prefix_product : out[0] = 1, out[i] = in[0] * ... * in[i-1]   (n outputs; out may alias in) */
int h2b_fr_batch_invert_dev(int device, void* d_a, size_t n, void* stream);
int h2b_fr_prefix_product_dev(int device, const void* d_in, void* d_out, size_t n, void* stream);
/* [UP] halo2_proofs::arithmetic::eval_polynomial(poly, point) = sum_i poly[i] * point^i  -> one Fr (32 B) at d_out;
 * [UP] halo2_proofs::arithmetic::kate_division(a, b): the n - 1 coefficients of a(X) / (X - b) -> d_q (must not alias d_a) */
int h2b_fr_eval_polynomial_dev(int device, const void* d_coeffs, size_t n, const uint64_t x[4], void* d_out, void* stream);
int h2b_fr_kate_division_dev(int device, const void* d_a, size_t n, const uint64_t b[4], void* d_q, void* stream);

/* out[i] = sum_j coeffs[j] * cols[j][i] (m columns of n elements; out may alias none of them): the y- / v-weighted sums of
 * the multi-open argument ([UP] poly/kzg/multiopen/shplonk/prover.rs) and the theta-compression of lookup expressions
 * ([UP] plonk/lookup/prover.rs compress_expressions) */
int h2b_fr_lincomb_dev(int device, const void* const* d_cols, const uint64_t* coeffs /* m x 4 */, uint32_t m, size_t n, void* d_out, void* stream);
/* out[c * rows + r] = in[r * cols + c] for a rows x cols matrix of Fr elements (out must not alias in): the re-layout step of the
 * four-step NTT that splits one transform over several devices (h2b_ntt_bn254_fr), usable on its own for row / column blocks of
 * device-resident columns.  Whole 32 x 32 tiles are moved by the TMA unit (cp.async.bulk + mbarrier). */
int h2b_fr_transpose_dev(int device, const void* d_in, void* d_out, uint32_t rows, uint32_t cols, void* stream);
/* The grand products themselves, on device-resident Lagrange-basis columns of n = 2^k rows:
 *   [UP] plonk/permutation/prover.rs Argument::commit, one call per set (chunk of cs.degree() - 2 columns; any number, 16 per launch):
 *        z[0] = last_z,  z[i+1] = z[i] * prod_j (v_j[i] + deltaomega * delta^j * omega^i * beta + gamma)
 *                                      / prod_j (v_j[i] + beta * s_j[i] + gamma)
 *        d_values[j] / d_permutations[j]: the j-th column of the set and its permutation column s_j; deltaomega = delta^(index of
 *        the set's first column in the whole argument); last_z = one for the first set, z[n - (blinding_factors + 1)] of the
 *        previous set afterwards.  All n entries of d_z are written; the caller overwrites the blinding rows.
 *   [UP] plonk/lookup/prover.rs Permuted::commit_product:
 *        z[0] = 1,  z[i+1] = z[i] * (a[i] + beta) (s[i] + gamma) / ((a'[i] + beta) (s'[i] + gamma))
 *        a, s: theta-compressed input / table expressions; a', s': their permuted versions. */
int h2b_permutation_product_dev(int device, const void* const* d_values, const void* const* d_permutations, uint32_t n_columns, size_t n,
                                const uint64_t beta[4], const uint64_t gamma[4], const uint64_t delta[4], const uint64_t deltaomega[4], const uint64_t omega[4],
                                const uint64_t last_z[4], void* d_z, void* stream);
int h2b_lookup_product_dev(int device, const void* d_compressed_input, const void* d_compressed_table, const void* d_permuted_input,
                           const void* d_permuted_table, size_t n, const uint64_t beta[4], const uint64_t gamma[4], void* d_z, void* stream);

/* [UP] plonk/lookup/prover.rs permute_expression_pair on the first usable_rows = n - (blinding_factors + 1) rows of the
 * theta-compressed input and table columns (device, Montgomery): permuted_input = the input sorted by canonical value;
 * permuted_table: every first occurrence of a value in permuted_input has that value at the same row, the remaining rows
 * receive the leftover table values exactly as upstream assigns them (ascending leftovers onto the repeated rows taken from
 * the end).  The outputs may alias the inputs; rows from usable_rows on (the blinding rows) are not touched.  The call
 * synchronises the stream: an input value that the table does not hold fails with H2B_ERR_BAD_ARGUMENT, as upstream returns
 * Error::ConstraintSystemFailure. */
int h2b_lookup_permute_dev(int device, const void* d_input, const void* d_table, uint32_t usable_rows, void* d_permuted_input, void* d_permuted_table,
                           void* stream);
/* The same without the synchronisation: the verdict is written to the 32-bit device word d_status on `stream` (0 = every input value
 * is in the table, otherwise 1 + the sorted row of one that is not -- the outputs are then meaningless); the caller reads it when it
 * next synchronises.  Lets several devices permute their lookups concurrently. */
int h2b_lookup_permute_async_dev(int device, const void* d_input, const void* d_table, uint32_t usable_rows, void* d_permuted_input,
                                 void* d_permuted_table, void* d_status, void* stream);

/* ---- quotient evaluation: evaluate_h on device-resident extended-coset columns ----------
 * [UP] halo2_proofs/src/plonk/evaluation.rs.  A GraphEvaluator is passed in the vocabulary upstream builds it in
 * (ValueSource / Calculation / CalculationInfo), flattened into plain arrays; the Rust shim fills these structs from
 * `Evaluator { custom_gates, lookups }` once per proving key.  Every column is a device pointer to `size` Fr elements
 * (the 2^extended_k evaluations over the zeta coset that coeff_to_extended produced). */
enum h2b_value_kind {          /* [UP] evaluation.rs `enum ValueSource`; index / rotation as upstream's tuple fields */
    H2B_VS_CONSTANT = 0,       /* constants[index] */
    H2B_VS_INTERMEDIATE = 1,   /* intermediates[index] */
    H2B_VS_FIXED = 2,          /* fixed[index][rotations[rotation]] */
    H2B_VS_ADVICE = 3,         /* advice[index][rotations[rotation]] */
    H2B_VS_INSTANCE = 4,       /* instance[index][rotations[rotation]] */
    H2B_VS_CHALLENGE = 5,      /* challenges[index] */
    H2B_VS_BETA = 6, H2B_VS_GAMMA = 7, H2B_VS_THETA = 8, H2B_VS_Y = 9,
    H2B_VS_PREVIOUS_VALUE = 10
};
enum h2b_calc_op {             /* [UP] evaluation.rs `enum Calculation` */
    H2B_CALC_ADD = 0, H2B_CALC_SUB = 1, H2B_CALC_MUL = 2, H2B_CALC_SQUARE = 3, H2B_CALC_DOUBLE = 4, H2B_CALC_NEGATE = 5,
    H2B_CALC_HORNER = 6,       /* value = a; for part in parts: value = value * b + part */
    H2B_CALC_STORE = 7
};
typedef struct h2b_value_source { uint32_t kind, index, rotation; } h2b_value_source;
typedef struct h2b_calculation {
    uint32_t op, target;                 /* intermediates[target] = op(...) */
    h2b_value_source a, b;               /* Horner: a = start value, b = factor */
    uint32_t parts_offset, parts_len;    /* Horner only: parts[parts_offset .. parts_offset + parts_len) */
} h2b_calculation;
typedef struct h2b_graph {               /* [UP] evaluation.rs `struct GraphEvaluator` */
    const uint64_t* constants; uint32_t n_constants;     /* n x 4, Montgomery */
    const int32_t* rotations; uint32_t n_rotations;
    const h2b_calculation* calculations; uint32_t n_calculations;
    const h2b_value_source* parts; uint32_t n_parts;
    uint32_t n_intermediates;
} h2b_graph;
typedef struct h2b_eval_columns {        /* host arrays of device pointers + the per-proof scalars of evaluate_h */
    const void* const* fixed; uint32_t n_fixed;
    const void* const* advice; uint32_t n_advice;
    const void* const* instance; uint32_t n_instance;
    const uint64_t* challenges; uint32_t n_challenges;    /* n x 4 */
    const uint64_t* beta; const uint64_t* gamma; const uint64_t* theta; const uint64_t* y;   /* 4 words each */
} h2b_eval_columns;
/* values[idx] = graph.evaluate(.., previous_value = values[idx], idx, rot_scale, isize = size) for every idx < size: the
 * "Custom gates" loop of evaluate_h.  The same call on Lagrange-basis columns (size = 2^k, rot_scale = 1) is the lookup argument's
 * compress_expressions ([UP] plonk/lookup/prover.rs).  A graph without calculations writes zero, as upstream does.  The rotated row is
 * get_rotation_idx(idx, rot, rot_scale, isize) = (idx + rot * rot_scale).rem_euclid(isize). */
int h2b_evaluate_graph_dev(int device, const h2b_graph* graph, const h2b_eval_columns* cols, void* d_values, uint32_t size, int32_t rot_scale,
                           void* stream);
/* The "Permutations" loop of evaluate_h: values[idx] is advanced by the l_0 / l_last / set-chaining / product terms.
 *   d_product_cosets[n_sets]   : permutation_product_coset of every set (z_i)
 *   d_columns[n_columns]       : the value coset of every column of the argument, in p.columns order (the caller resolves
 *                                Any::{Advice, Fixed, Instance} to a pointer); d_perm_cosets[n_columns] = pk.permutation.cosets
 *   chunk_len = cs.degree() - 2; last_rotation = -(blinding_factors + 1); delta = Fr::DELTA; zeta = Fr::ZETA */
int h2b_evaluate_h_permutation_dev(int device, void* d_values, uint32_t size, int32_t rot_scale, const void* const* d_product_cosets, uint32_t n_sets,
                                   const void* const* d_columns, const void* const* d_perm_cosets, uint32_t n_columns, uint32_t chunk_len,
                                   int32_t last_rotation, const void* d_l0, const void* d_l_last, const void* d_l_active_row, const uint64_t beta[4],
                                   const uint64_t gamma[4], const uint64_t y[4], const uint64_t delta[4], const uint64_t zeta[4],
                                   const uint64_t extended_omega[4], void* stream);
/* One iteration of the "Lookups" loop of evaluate_h: table_value = graph.evaluate(.., previous_value = 0, ..) and the five
 * lookup terms on product_coset (z), permuted_input_coset (a') and permuted_table_coset (s'). */
int h2b_evaluate_h_lookup_dev(int device, const h2b_graph* graph, const h2b_eval_columns* cols, void* d_values, uint32_t size, int32_t rot_scale,
                              const void* d_product_coset, const void* d_permuted_input_coset, const void* d_permuted_table_coset, const void* d_l0,
                              const void* d_l_last, const void* d_l_active_row, void* stream);
/* Row-sharded variants (evaluate_h across devices): this call produces rows [row0, row0 + rows) of the domain.  Every column pointer
 * (fixed / advice / instance, product / permuted / sigma cosets, l_0, l_last, l_active) is then a SLICE of rows + 2 * halo rows that
 * starts at global row (row0 - halo) mod size -- wrap-around included -- and d_values is the halo-free slice of `rows` rows.  halo must
 * cover the largest rotated read: max |rotation| * rot_scale of the graph, rot_scale for lookups, max(1, |last_rotation|) * rot_scale
 * for the permutation argument; otherwise the call is refused.  Row values do not depend on the sharding. */
typedef struct h2b_eval_shard { uint32_t row0, rows, halo; } h2b_eval_shard;
int h2b_evaluate_graph_shard_dev(int device, const h2b_graph* graph, const h2b_eval_columns* cols, void* d_values, uint32_t size, int32_t rot_scale,
                                 const h2b_eval_shard* shard, void* stream);
int h2b_evaluate_h_permutation_shard_dev(int device, void* d_values, uint32_t size, int32_t rot_scale, const void* const* d_product_cosets, uint32_t n_sets,
                                         const void* const* d_columns, const void* const* d_perm_cosets, uint32_t n_columns, uint32_t chunk_len,
                                         int32_t last_rotation, const void* d_l0, const void* d_l_last, const void* d_l_active_row, const uint64_t beta[4],
                                         const uint64_t gamma[4], const uint64_t y[4], const uint64_t delta[4], const uint64_t zeta[4],
                                         const uint64_t extended_omega[4], const h2b_eval_shard* shard, void* stream);
int h2b_evaluate_h_lookup_shard_dev(int device, const h2b_graph* graph, const h2b_eval_columns* cols, void* d_values, uint32_t size, int32_t rot_scale,
                                    const void* d_product_coset, const void* d_permuted_input_coset, const void* d_permuted_table_coset, const void* d_l0,
                                    const void* d_l_last, const void* d_l_active_row, const h2b_eval_shard* shard, void* stream);
/* slots (live values per row) and micro-operations the last compiled graph of this thread needed -- diagnostics / tests */
int h2b_evaluate_graph_info(uint32_t* slots, uint32_t* micro_ops);

/* ---- SRS on-disk format ------------------------------------------------------------------
 * [UP] halo2_proofs/src/poly/kzg/commitment.rs ParamsKZG::{read_custom, write_custom} / halo2_proofs::SerdeFormat and
 * [UP] halo2curves 0.3.x GroupEncoding / SerdeObject for G1Affine; the reference loads params/kzg_bn254_{k}.srs at
 * src/scaffold.rs:119,174,271.  File: k (u32 LE) | g[2^k] | g_lagrange[2^k] | g2 | s_g2. */
enum h2b_serde_format {
    H2B_SERDE_PROCESSED = 0,            /* compressed: 32 B = canonical LE x, (y & 1) << 7 in byte 31, all-zero = identity */
    H2B_SERDE_RAW_BYTES = 1,            /* x | y Montgomery limbs (64 B), range- and curve-checked */
    H2B_SERDE_RAW_BYTES_UNCHECKED = 2   /* the same bytes, unchecked */
};
/* n encoded points (device) -> n affine points x | y Montgomery (device; may alias the input for the raw formats).
 * With first_invalid != NULL the call synchronises the stream, stores the index of the first invalid encoding (or
 * UINT64_MAX) and fails with H2B_ERR_BAD_ARGUMENT if there is one (upstream unwraps / returns io::Error); with NULL it
 * stays asynchronous and invalid points decode to the identity. */
int h2b_g1_decode_dev(int device, const void* d_bytes, size_t n, int format, void* d_out_affine, uint64_t* first_invalid, void* stream);
/* n affine points (device) -> n x 32 compressed bytes (device): G1Affine::to_bytes */
int h2b_g1_encode_dev(int device, const void* d_affine, size_t n, void* d_out_bytes, void* stream);
/* ParamsKZG::read_custom: reads the file, decodes g and g_lagrange on the device and leaves both resident as registered
 * base sets (window tables built, every device of h2b_init) -> *handle_g, *handle_g_lagrange for
 * h2b_msm_bn254_g1_registered (commit / commit_lagrange).  out_g / out_g_lagrange (2^k x 8 words each, may be NULL) receive
 * the decoded points for the host-side ParamsKZG; the G2 section (g2 | s_g2, 2 x 64 B compressed or 2 x 128 B raw) is
 * returned as bytes for the host to parse.  Either handle pointer may be NULL (that vector is then only decoded). */
int h2b_srs_read(const char* path, int format, uint32_t* k, uint64_t* out_g, uint64_t* out_g_lagrange, uint8_t* g2_bytes, size_t g2_cap,
                 size_t* g2_len, uint64_t* handle_g, uint64_t* handle_g_lagrange);
/* A file that was read before and has not changed since (same path, format, size and mtime) is NOT decoded again: the call
 * hands out the resident base sets (reference counted -- every handle still has to be given back with
 * h2b_unregister_bases) and, if asked, downloads the points.  The reference re-reads the SRS file for every proof
 * (src/scaffold.rs:174).  H2B_SRS_CACHE=0 in the environment disables this; h2b_srs_cache_clear drops the cache's references. */
int h2b_srs_cache_clear(void);
/* ParamsKZG::write_custom for host arrays (Processed: compressed on the device) */
int h2b_srs_write(const char* path, int format, uint32_t k, const uint64_t* g, const uint64_t* g_lagrange, const uint8_t* g2_bytes, size_t g2_len);

/* ---- raw device memory helpers (so that non-CUDA hosts -- ctypes, Rust -- can hold device buffers) --- */
int h2b_dev_alloc(int device, size_t bytes, void** out);
int h2b_dev_free(int device, void* p);
int h2b_memcpy_h2d(int device, void* d_dst, const void* h_src, size_t bytes);
int h2b_memcpy_d2h(int device, void* h_dst, const void* d_src, size_t bytes);
/* Upload a fresh column while the device keeps working: nothing in flight may touch d_dst; work queued on `stream` after the
 * call sees the data.  Pageable sources are consumed before the call returns (pinned staging threads), pinned sources must
 * stay valid until the stream reaches the copy. */
int h2b_memcpy_h2d_async(int device, void* d_dst, const void* h_src, size_t bytes, void* stream);
/* column copies / zero padding on the caller's stream (the padding of coeff_to_extended, copies that keep the Lagrange form) */
int h2b_memcpy_d2d_async(int device, void* d_dst, const void* d_src, size_t bytes, void* stream);
int h2b_memset_zero_async(int device, void* d_dst, size_t bytes, void* stream);
int h2b_dev_sync(int device);

/* ---- synthetic workload + diagnostics ---------------------------------------------------------------- */
/* P_i = [z_i] G with z_i the i-th SplitMix64 value of stream `seed` (valid, distinct G1 points). */
int h2b_gen_points_dev(int device, uint64_t seed, size_t n, void* d_out_affine, void* stream);
/* kind 0: uniform in [0, r);  kind 1: witness-like (50% 0, 20% 1, 20% < 2^19, 10% r - small). */
int h2b_gen_scalars_dev(int device, uint64_t seed, size_t n, int kind, void* d_out, void* stream);
/* O(n) checksum of an MSM over the synthetic points: for P_i = [z_i] G (h2b_gen_points_dev, stream `seed`, indices first + i),
 * sum_i s_i P_i = [sum_i s_i z_i mod r] G.  Writes the CANONICAL (non-Montgomery) value of sum_i s_i z_i mod r as 32 bytes at d_out;
 * field arithmetic only, so it shares no code with the MSM it checks (bench.py `verified`). */
int h2b_msm_checksum_dev(int device, const void* d_scalars /* n x 32 B Montgomery */, uint64_t seed, uint64_t first, size_t n, void* d_out, void* stream);
/* element-wise field op on host arrays.  field: 0 Fr, 1 Fq.  op: 0 add, 1 sub, 2 mul, 3 sqr(a), 4 inv(a),
 * 5 from_mont(a), 6 to_mont(a) */
int h2b_field_op(int field, int op, const uint64_t* a, const uint64_t* b, size_t n, uint64_t* out);
/* element-wise group op on host arrays of affine points; op: 0 p+q, 1 p-q, 2 4(p+q) via the general
 * XYZZ adder and doubling, 3 p+q through the Jacobian conversion.  out: affine. */
int h2b_ec_op(int op, const uint64_t* p, const uint64_t* q, size_t n, uint64_t* out);
/* integer-pipe micro-benchmark; kind 0 IMAD, 1 IMAD.WIDE, 2 dependent Fq multiplications, 3 dependent
 * XYZZ mixed additions, 4 dependent Fq squarings, 5 dependent a*b + c*d with one reduction.  Returns elapsed milliseconds and the number of operations executed. */
int h2b_imad_bench(int device, int kind, int iters, float* ms_out, double* ops_out);
/* number of CUDA kernels this library has launched since it was loaded */
unsigned long long h2b_launch_count(void);
/* per-kernel timing with CUDA events recorded on the launching stream. After h2b_profile_enable(device, 1) every
 * MSM / NTT call records one event per phase; h2b_profile_read synchronises the device and returns (tag, ms) pairs
 * in launch order.  MSM tags: 1 decompose, 2 scan, 3 scatter, 4 plan, 5 accumulate, 6 combine, 7 bucket reduction,
 * 8 final;  NTT tags: 15 twiddle tables, 16 + p = pass p. */
int h2b_profile_enable(int device, int on);
int h2b_profile_read(int device, int* tags, float* ms, int cap, int* count);
/* force the MSM window size (0 = automatic) -- tuning / tests only */
int h2b_set_msm_window(int c);
/* window-table policy for base sets registered from now on: -1 no tables, 0 automatic spacing, 2..24 forced
 * spacing (tuning / tests; the environment variable H2B_MSM_PRECOMP sets the initial value) */
int h2b_set_msm_precomp(int spacing);
/* tables, spacing (bits) and device bytes per device (the largest share) of a registered base set */
int h2b_base_set_info(uint64_t handle, uint32_t* n_tables, uint32_t* spacing, uint64_t* device_bytes);
/* implicit cache of h2b_msm_bn254_g1: out = {uploads, verified hits, stale copies discarded, uncached (direct) calls,
 * resident implicit sets, of which with tables} -- diagnostics / tests */
int h2b_implicit_cache_stats(uint64_t out[6]);

#ifdef __cplusplus
}
#endif
#endif /* H2B200_H */
