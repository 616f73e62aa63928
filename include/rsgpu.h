/*
 * rsgpu.h -- C ABI of the B200-native prover backend for zkFHE/ringSNARK (librsgpu.so).
 *
 * The reference has no FFI: its boundary is the C++ template concept RingT / EncT that
 * ringsnark/zk_proof_systems, ringsnark/reductions and ringsnark/util require.
 * This header is what a backend header pair (ringsnark_b200/cpp/ringsnark/seal_gpu/seal_ring.hpp) binds to;
 * every entry point names the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain C, opaque handles, plain pointers and sizes; no exceptions cross the boundary;
 *   - every function returns RSG_OK (0) or a negative status; rsg_last_error() gives the thread-local text;
 *   - all data are uint64 words holding canonical residues;
 *   - ring element  : [L_R][N_R] words, limb-major            (depends/SEAL-Polytools/include/poly_arith.h:81-102)
 *   - encoding      : [L_R][2][L_E][N_E] words, NTT form        (ringsnark/seal/seal_ring.hpp:225,
 *                                                                depends/SEAL/native/src/seal/ciphertext.h:337-349)
 *   - "h_" pointers are HOST memory, "d_" pointers are DEVICE memory of the context's GPU;
 *   - thread-safe: concurrent calls on one context are serialised per stream slot (rinocchio.tcc:106-163
 *     calls inner_product from 10 OpenMP sections);
 *   - there is NO CPU fallback: without a usable CUDA device rsg_context_create fails with RSG_ERR_CUDA.
 */
#ifndef RSGPU_H
#define RSGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSG_OK 0
#define RSG_ERR_ARG (-1)       /* invalid argument (reference: std::invalid_argument) */
#define RSG_ERR_CUDA (-2)      /* CUDA runtime failure / no device */
#define RSG_ERR_STATE (-3)     /* context not set / handle misuse (reference: "context not set") */
#define RSG_ERR_NOTINV (-4)    /* "element is not invertible in ring" (seal_ring.tcc:87-103) */
#define RSG_ERR_UNSUPPORTED (-5)
#define RSG_ERR_NOISE (-6)     /* decoding_error: a ciphertext has no noise budget left (seal_ring.tcc:445-453) */
#define RSG_ERR_TRANSPARENT (-7) /* a fused prover met a transparent-ciphertext candidate (seal_ring.tcc:493-504): the outputs are
                                  * NOT valid; redo the proof with the per-inner-product entry points, which resolve it exactly */

/* Per-term dispatch of EncodingElem::operator*= (seal_ring.tcc:509-548) decided by the host shim exactly as the
 * reference does: is_zero() (incl. the SealPoly::is_zero prefix quirk) -> SKIP; scalar 1 -> ONE (ciphertext taken
 * unchanged); anything else -> GENERAL (batch-encode, lift, NTT, dyadic product). */
#define RSG_TERM_SKIP 0
#define RSG_TERM_ONE 1
#define RSG_TERM_GENERAL 2

typedef struct rsg_context rsg_context;   /* replaces the statics of RingElem / EncodingElem (seal_ring.hpp:25,218-223) */
typedef struct rsg_crs rsg_crs;           /* a vector<EncodingElem> laid out contiguously in HBM (groth16.hpp:14-20) */
typedef struct rsg_ringvec rsg_ringvec;   /* a vector<RingElem> resident in HBM */

const char *rsg_last_error(void);
int rsg_device_count(void);

/* RingElem::set_context + EncodingElem::set_context (seal_ring.hpp:52-58, 266-320).
 * q[L_R] ring primes (= plaintext moduli t_j), Q[L_E] first-level ciphertext primes; every prime < 2^61 and
 * == 1 mod 2*N_E.  NTT tables use SEAL's minimal primitive 2N-th root (util/numth.cpp:386-412). */
int rsg_context_create(rsg_context **out, size_t N_R, size_t L_R, const uint64_t *q, size_t N_E, size_t L_E,
                       const uint64_t *Q, int device);
void rsg_context_destroy(rsg_context *ctx);
int rsg_context_sync(rsg_context *ctx);
/* Run every subsequent launch of this context on a caller-owned CUDA stream (cudaStream_t as void*). */
int rsg_context_set_stream(rsg_context *ctx, void *cuda_stream);
/* Kernels launched by this context since creation (the bench's "gpu_launches" claim). */
uint64_t rsg_context_launch_count(const rsg_context *ctx);
/* Work counters since context creation (measurement only; bench.py derives the algorithmic bytes of k_crs_lincomb from
 * them): "lincomb_terms" (CRS elements streamed), "lincomb_plain_terms" (of which multiplied by a plaintext),
 * "lincomb_launches", "ntt_forward_polys", "ntt_inverse_polys" (N_E-point transforms), "merged_lincombs",
 * "exact_fallbacks" (transparent-prefix resolutions, seal_ring.tcc:493-504).  Unknown name: 0. */
uint64_t rsg_context_stat(const rsg_context *ctx, const char *name);

/* ---- CRS: vector<EncodingElem> produced by EncodingElem::encode (seal_ring.tcc:324-359) ---- */
int rsg_crs_create(rsg_context *ctx, size_t n_elems, rsg_crs **out);                         /* uninitialised arena */
int rsg_crs_upload(rsg_crs *crs, size_t first, size_t count, const uint64_t *h_words);       /* count encodings */
int rsg_crs_download(const rsg_crs *crs, size_t first, size_t count, uint64_t *h_words);
int rsg_crs_fill_uniform(rsg_crs *crs, uint64_t seed);   /* synthetic CRS: uniform residues, generated on device */
/* The same stream of words for a SHARD: elements [first, first + count) of this arena receive what elements
 * [virtual_first, virtual_first + count) of an arena filled by rsg_crs_fill_uniform(seed) hold -- every rank of a multi-GPU
 * run then proves over the same CRS as the single-GPU run (bench.py checks the proofs word for word). */
int rsg_crs_fill_uniform_at(rsg_crs *crs, size_t first, size_t count, uint64_t virtual_first, uint64_t seed);
uint64_t *rsg_crs_device_ptr(rsg_crs *crs);
void rsg_crs_destroy(rsg_crs *crs);

/* ---- vectors of ring elements ---- */
int rsg_ringvec_create(rsg_context *ctx, size_t n_elems, rsg_ringvec **out);
int rsg_ringvec_upload(rsg_ringvec *v, size_t first, size_t count, const uint64_t *h_words);
int rsg_ringvec_download(const rsg_ringvec *v, size_t first, size_t count, uint64_t *h_words);
int rsg_ringvec_fill_uniform(rsg_ringvec *v, uint64_t seed);
uint64_t *rsg_ringvec_device_ptr(rsg_ringvec *v);
size_t rsg_ringvec_size(const rsg_ringvec *v);
void rsg_ringvec_destroy(rsg_ringvec *v);
/* SealPoly::is_zero with its byte/word confusion (poly_arith.cpp:147-153): h_flags[i] = 1 iff bytes
 * [0, L_R*N_R + 7) of element first+i are zero.  One kernel + one small copy for the whole range. */
int rsg_ringvec_is_zero_prefix(const rsg_ringvec *v, size_t first, size_t count, uint8_t *h_flags);

/* ---- RingElem operators on device-resident vectors: the element-wise fall-backs of the RingT concept (seal_ring.tcc:
 * 105-263 over poly_arith.cpp:164-350).  `count` elements starting at the given indices; out may alias a. ---- */
#define RSG_OP_ADD 0
#define RSG_OP_SUB 1
#define RSG_OP_MUL 2
int rsg_ring_binop(rsg_context *ctx, int op, const rsg_ringvec *a, size_t a_first, const rsg_ringvec *b, size_t b_first,
                   rsg_ringvec *out, size_t out_first, size_t count);
/* poly (op) scalar with SEAL's scalar semantics: add/sub take the scalar as is (one correction), mul reduces it first. */
int rsg_ring_scalar_op(rsg_context *ctx, int op, const rsg_ringvec *a, size_t a_first, uint64_t scalar, rsg_ringvec *out,
                       size_t out_first, size_t count);
int rsg_ring_negate(rsg_context *ctx, const rsg_ringvec *a, size_t a_first, rsg_ringvec *out, size_t out_first, size_t count);
/* Per-slot inverses; RSG_ERR_NOTINV if some element has a zero slot ("element is not invertible in ring",
 * seal_ring.tcc:87-103); h_ok (nullable) receives 1 per invertible element. */
int rsg_ring_invert(rsg_context *ctx, const rsg_ringvec *a, size_t a_first, rsg_ringvec *out, size_t out_first, size_t count,
                    uint8_t *h_ok);

/* ---- hot path (b): EncodingElem::inner_product (seal_ring.tcc:361-433) ----
 * out = sum over i in [0, count) with tag[i] != SKIP of crs[crs_first+i] (*) coeffs[coeff_first+i].
 * h_out (host, may be NULL) and/or d_out (device, may be NULL) receive one encoding.  *n_used = number of summed
 * terms; 0 means the reference returns an EMPTY EncodingElem and the outputs are all-zero words. */
int rsg_inner_product(rsg_context *ctx, const rsg_crs *crs, size_t crs_first, const rsg_ringvec *coeffs,
                      size_t coeff_first, size_t count, const uint8_t *h_tags, uint64_t *h_out, uint64_t *d_out,
                      size_t *n_used);
/* Same with explicit term lists (iterator ranges that are not contiguous in their arenas): term i multiplies
 * crs[h_crs_idx[i]] by coeffs[h_coeff_idx[i]]. */
int rsg_inner_product_idx(rsg_context *ctx, const rsg_crs *crs, const uint32_t *h_crs_idx, const rsg_ringvec *coeffs,
                          const uint32_t *h_coeff_idx, size_t count, const uint8_t *h_tags, uint64_t *h_out,
                          uint64_t *d_out, size_t *n_used);
/* EncodingElem::operator+= for non-empty operands (seal_ring.tcc:479-507 -> evaluator.cpp:217-231): acc += other. */
int rsg_enc_add(rsg_context *ctx, uint64_t *d_acc, const uint64_t *d_other);
/* Copy-construct encodings inside / between arenas (EncodingElem's value semantics, seal_ring.hpp:245-247). */
int rsg_crs_copy(rsg_crs *dst, size_t dst_first, const rsg_crs *src, size_t src_first, size_t count);
/* out = sum of `parts` blocks stored back to back, each block = n_enc encodings (n_enc = 3: a whole proof):
 * the modular-add kernel that follows the NCCL all-gather (modular addition is not an NCCL reduction). */
int rsg_enc_sum(rsg_context *ctx, const uint64_t *d_parts, size_t parts, size_t n_enc, uint64_t *d_out);
/* The same over parts that lie part_stride_words apart (0 = back to back): the all-gathered [partial proof | probe block]
 * records of rsg_groth16_lincombs_shard. */
int rsg_enc_sum_strided(rsg_context *ctx, const uint64_t *d_parts, size_t parts, size_t n_enc, size_t part_stride_words,
                        uint64_t *d_out);

/* ---- hot path (a): r1cs_to_qrp_witness_map (reductions/r1cs_to_qrp/r1cs_to_qrp.tcc:148-259) ----
 * evals: 9*n ring elements, order A_mid,B_mid,C_mid,A_io,B_io,C_io,A_full,B_full,C_full (the outputs of
 * linear_combination::evaluate, :175-219), each block n elements.
 * coeffs: receives 6*n elements A_io,B_io,C_io,A_mid,B_mid,C_mid (qrp_witness order, qrp.hpp:171-181).
 * H: receives n+1 elements: (A*B - C)/Z in [0, n-1), zeros at n-1 and n (non-ZK call of groth16.tcc:82-84).
 * Interpolation on the domain {0..n-1} (util/polynomials.tcc:9-43), product and exact division by the monic
 * Z (util/polynomials.tcc:61-81, util/evaluation_domain.tcc:53-84); all slot-parallel on the GPU.
 * Two implementations with identical (canonical) results: dense constant-matrix products for small n, and for
 * n >= 320 the quasi-linear path of csrc/witness_fast.cuh (Newton coefficients by one negacyclic product, Newton -> monomial
 * on the subproduct tree, quotient by two products): one transform per product while its size stays <= min(N_E, 32768)
 * (n <= 8208 at N_E = 2^14, <= 16400 at 2^15), products assembled from blocks of N_E/2 coefficients beyond that, up to
 * n = 2*N_E (n = 2^16 at N_E = 2^15); larger n takes the dense path.  RSG_WITNESS=dense|fast overrides the choice
 * (RSG_WF_TS=<power of two> caps the transform size: tests reach the blocked mode at small n with it);
 * rsg_context_stat("witness_fast_launches" / "witness_dense_launches") says which one ran. */
int rsg_witness_map(rsg_context *ctx, size_t n, const rsg_ringvec *evals, rsg_ringvec *coeffs, rsg_ringvec *H);
/* The zero-knowledge variant rinocchio::prover calls (rinocchio.tcc:88-93, r1cs_to_qrp.tcc:225-235):
 * h_d = 3 ring elements d1, d2, d3 (host words, [3][L_R][N_R]); H[i] += d2*A[i] + d1*B[i] (i < n), H[0] -= d3,
 * H[i] += d1*d2*Z[i] (i <= n) with A, B the interpolants of the FULL assignment.  h_d == NULL is rsg_witness_map. */
int rsg_witness_map_zk(rsg_context *ctx, size_t n, const rsg_ringvec *evals, const uint64_t *h_d, rsg_ringvec *coeffs,
                       rsg_ringvec *H);
/* The same when `evals` IS the output of rsg_r1cs_evaluate(r1cs, .): then full = mid + io - (constant wire), interpolation
 * is linear, and the interpolants of the full assignment are formed from the other six instead of being computed
 * (6 instead of 8 interpolations per proof; identical residues). */
typedef struct rsg_r1cs rsg_r1cs;
int rsg_witness_map_r1cs(rsg_context *ctx, rsg_r1cs *r1cs, const rsg_ringvec *evals, const uint64_t *h_d,
                         rsg_ringvec *coeffs, rsg_ringvec *H);
/* The witness map as groth16::prover consumes it (groth16.tcc:82-112, non-ZK): coefficients_for_C_io / C_mid are never
 * read by that prover and C does not reach the quotient H (deg C < n = deg Z), so only A and B are interpolated
 * (4 interpolations).  The C_io / C_mid blocks of `coeffs` are left untouched. */
int rsg_witness_map_groth16(rsg_context *ctx, rsg_r1cs *r1cs, const rsg_ringvec *evals, rsg_ringvec *coeffs, rsg_ringvec *H);
/* util/polynomials.tcc:9-43 on its own: vectors of n ring elements; `batch` vectors back to back. */
int rsg_interpolate(rsg_context *ctx, size_t n, size_t batch, const rsg_ringvec *y, size_t y_first, rsg_ringvec *out,
                    size_t out_first);
/* util/evaluation_domain.tcc:53-60: coefficients of Z(x) = prod_{i<n}(x-i) mod each ring prime: h_Z[L_R][n+1]. */
int rsg_vanishing(rsg_context *ctx, size_t n, uint64_t *h_Z);

/* ---- the step before the hot path: linear_combination::evaluate (relations/variable.tcc:246-254) ----
 * R1CS in CSR form, rows r = m*n + i for matrix m in {A, B, C} and constraint i; col 0 is the constant wire,
 * col v >= 1 is variable v (primary inputs first); coeff are the uint64 scalars of the linear terms
 * (relations/variable.hpp:29: negative integers wrap through uint64, as in the reference). */
int rsg_r1cs_create(rsg_context *ctx, size_t n, size_t n_io, size_t n_aux, const uint32_t *h_row_ptr /* 3n+1 */,
                    const uint32_t *h_col, const uint64_t *h_coeff, rsg_r1cs **out);
void rsg_r1cs_destroy(rsg_r1cs *r);
/* assignment: n_io + n_aux ring elements; evals: 9n elements in rsg_witness_map's input order. */
int rsg_r1cs_evaluate(rsg_context *ctx, const rsg_r1cs *r, const rsg_ringvec *assignment, rsg_ringvec *evals);

/* ---- instance map with evaluation: r1cs_to_qrp_instance_map_with_evaluation (reductions/r1cs_to_qrp/r1cs_to_qrp.tcc:
 * 75-116) with evaluate_all_lagrange_polynomials / compute_vanishing_polynomial (util/evaluation_domain.tcc:20-51) --
 * the O(m^2) step of generator (groth16.tcc:11-12, rinocchio.tcc:12-13) and of every verifier call (groth16.tcc:127-128,
 * rinocchio.tcc:220-221).  t = element t_first of `t` (must not hit a domain point in any slot: the caller keeps the
 * reference's std::find check).  ABCt: 3 * (n_io + n_aux + 1) elements At | Bt | Ct indexed by variable (0 = constant
 * wire); Ht: n + 1 elements 1, t, .., t^n; Zt: 1 element Z(t). ---- */
int rsg_instance_map(rsg_context *ctx, rsg_r1cs *r1cs, const rsg_ringvec *t, size_t t_first, rsg_ringvec *ABCt, rsg_ringvec *Ht,
                     rsg_ringvec *Zt);

/* ---- EncodingElem::decode (ringsnark/seal/seal_ring.tcc:435-477): the verifier's front half (groth16.tcc:121-123,
 * rinocchio.tcc:205-217) for first-level BGV ciphertexts in NTT form with correction factor 1.  Per ring limb j:
 * invariant noise budget (decryptor.cpp:383-461), c0 + c1 s, inverse NTT, exact base conversion q -> t
 * (util/rns.cpp:466-539), BatchEncoder::decode (batchencoder.cpp:278-315), first N_R slots.
 * h_sk: [L_R][L_E][N_E] words of the secret keys in NTT form (SecretKey::data() of limb j's context, first L_E limbs);
 * encodings from HBM (d_enc) or from the host (h_enc), exactly one non-null; h_ring: [count][L_R][N_R];
 * h_budget (nullable): [count][L_R] noise budgets in bits.  Returns RSG_ERR_NOISE (after filling both outputs) when some
 * budget is <= 0, the reference's decoding_error.  An all-zero encoding (SEAL's empty ciphertext) decodes to zero. ---- */
int rsg_decode(rsg_context *ctx, const uint64_t *h_sk, const uint64_t *d_enc, const uint64_t *h_enc, size_t count, uint64_t *h_ring,
               int32_t *h_budget);

/* ---- EncodingElem::encode (ringsnark/seal/seal_ring.tcc:324-359): BatchEncoder::encode + Encryptor::encrypt_symmetric (BGV,
 * depends/SEAL/native/src/seal/encryptor.cpp:242-312, util/rlwe.cpp:276-390) of ring elements [first, first + count) of
 * `elems` straight into encodings [out_first, ...) of the arena `out` -- the CRS never exists on the host.
 * h_sk: [L_R][L_E][N_E] secret keys in NTT form (as rsg_decode).  h_seeds: [count][L_R][8] words -- per (element, ring limb)
 * the 64-byte seed SEAL's random generator factory would hand to that encryption's bootstrap PRNG (randomgen.h:440-448:
 * fresh system randomness per call, or the factory's fixed seed for a seeded context).  Randomness is SEAL's Blake2xbPRNG
 * and SEAL's samplers (uniform with rejection, centred binomial), so a seeded context gives SEAL's ciphertext words bit
 * for bit.  Scalar ring elements must be uploaded as polynomials with every slot set (RingElem::to_poly, as the reference
 * does at seal_ring.tcc:343-344). ---- */
int rsg_encode(rsg_context *ctx, const uint64_t *h_sk, const rsg_ringvec *elems, size_t first, size_t count, const uint64_t *h_seeds,
               rsg_crs *out, size_t out_first);

/* ---- (de)serialisation of encodings: a CRS / proving-key range or a proof.  The reference declares the stream
 * operators (zk_proof_systems/r1cs_ppzksnark.hpp:43-47,142-146) and never defines them; the container is documented in
 * ringsnark_b200/csrc/serialize.inl (magic, parameters and primes, payload in HBM layout, checksum, end mark).
 * rsg_enc_file_* are host-only (no GPU needed); kind 1 = CRS range, 2 = proof.  Readers fail with RSG_ERR_ARG on foreign
 * parameters, non-canonical words, checksum mismatch or truncation. ---- */
#define RSG_FILE_CRS 1
#define RSG_FILE_PROOF 2
int rsg_enc_file_info(const char *path, uint64_t *info /* kind, N_R, L_R, N_E, L_E, n_elems */, uint64_t *q /* nullable, 8 */,
                      uint64_t *Q /* nullable, 16 */);
int rsg_enc_file_write(const char *path, uint64_t kind, size_t N_R, size_t L_R, const uint64_t *q, size_t N_E, size_t L_E,
                       const uint64_t *Q, size_t n_elems, const uint64_t *h_words);
int rsg_enc_file_read(const char *path, size_t N_R, size_t L_R, const uint64_t *q, size_t N_E, size_t L_E, const uint64_t *Q,
                      size_t cap_elems, uint64_t *h_words, size_t *n_elems, uint64_t *kind);
int rsg_crs_save(const rsg_crs *crs, size_t first, size_t count, const char *path);   /* HBM -> file, pinned staging */
int rsg_crs_load(rsg_crs *crs, size_t first, const char *path, size_t *count);          /* file -> HBM */

/* ---- groth16::prover (zk_proof_systems/groth16/groth16.tcc:69-115), whole prover in one call ----
 * The proving key's CRS vectors live in ONE arena; each vector may be a shard [lo, hi) of its terms (multi-GPU:
 * the partial proofs of all ranks are all-gathered and summed with rsg_enc_sum). */
typedef struct {
  size_t s_pows_off, s_pows_lo, s_pows_hi;          /* s_pows[0..n]  (groth16.hpp:17); the prover uses [0, n) */
  size_t delta_ts_off, delta_ts_lo, delta_ts_hi;    /* delta_ts[0..n] */
  size_t delta_mid_off, delta_mid_lo, delta_mid_hi; /* delta_mid[0..n_aux) */
  size_t alpha_idx, beta_idx;                       /* arena index, or (size_t)-1 if this shard does not add them */
} rsg_groth16_layout;
#define RSG_AUX_POLY 0xFF   /* h_aux_kind[i]: element is a polynomial -> SealPoly::is_zero prefix test decides */
/* h_assignment (host, nullable): if given it is first copied into `assignment` (n_io + n_aux elements).
 * h_aux_kind (nullable = all RSG_AUX_POLY): RSG_TERM_* for auxiliary inputs the caller holds as scalars.
 * Outputs (either may be NULL): 3 encodings A, B, C.  n_used[3] (nullable): summed terms per proof element. */
int rsg_groth16_prove(rsg_context *ctx, const rsg_r1cs *r1cs, const rsg_crs *crs, const rsg_groth16_layout *layout,
                      rsg_ringvec *assignment, const uint64_t *h_assignment, const uint8_t *h_aux_kind,
                      uint64_t *h_proof, uint64_t *d_proof, size_t *n_used);

/* The same prover over CRS vectors that live in DIFFERENT arenas -- what the generator templates produce: one
 * EncodingElem::encode call, hence one arena, per vector (groth16.tcc:36-55).  refs[0..4] = s_pows[0..n], delta_ts[0..n],
 * delta_mid[0..n_aux), alpha, beta: the arena and the index of the FIRST encoding of each.  Returns RSG_ERR_TRANSPARENT when a
 * probe sum vanished and RSG_ERR_UNSUPPORTED when the term set exceeds the NTT-plaintext budget; the caller then forms the proof
 * from rsg_inner_product / rsg_enc_add as the template does. */
typedef struct {
  const rsg_crs *crs;
  size_t first;
} rsg_crs_ref;
int rsg_groth16_prove_refs(rsg_context *ctx, const rsg_r1cs *r1cs, const rsg_crs_ref refs[5], rsg_ringvec *assignment,
                           const uint64_t *h_assignment, const uint8_t *h_aux_kind, uint64_t *h_proof, uint64_t *d_proof,
                           size_t *n_used);
/* rinocchio::prover (zk_proof_systems/rinocchio/rinocchio.tcc:74-190) in one call.  refs[0..5] = s_pows[0..n],
 * alpha_s_pows[0..n], beta_prods[0..n_aux), beta_rv_ts, beta_rw_ts, beta_ry_ts (the last three only read in zero-knowledge mode
 * with auxiliary inputs).  h_d: the prover's d1, d2, d3 ([3][L_R][N_R] host words, RingElem::random_invertible_element drawn by
 * the caller in the reference's order) or NULL for the non-zero-knowledge proof (rinocchio.tcc:81-90).  Outputs: 9 encodings
 * A, alpha_A, B, alpha_B, C, alpha_C, D, alpha_D, F (proof order, rinocchio.hpp); F is all-zero words (the reference's empty
 * encoding) without auxiliary inputs.  Each coefficient is batch-encoded and transformed ONCE and multiplied into both the
 * s_pows and the alpha_s_pows stream.  Error returns as rsg_groth16_prove_refs. */
int rsg_rinocchio_prove(rsg_context *ctx, const rsg_r1cs *r1cs, const rsg_crs_ref refs[6], rsg_ringvec *assignment,
                        const uint64_t *h_assignment, const uint8_t *h_aux_kind, const uint64_t *h_d, uint64_t *h_proof,
                        uint64_t *d_proof, size_t *n_used);

/* The second half of rsg_groth16_prove on its own (the multi-GPU driver runs the witness map sharded by SLOT, exchanges
 * the coefficients with one all-to-all, and then calls this on every rank's TERM shard): the six inner products and the
 * operator+= chain of groth16.tcc:89-112 from device-resident coefficient vectors.  d_vec[k], k = A_io, A_mid, B_io, B_mid
 * (terms [s_pows_lo, min(s_pows_hi, n))), H (terms [delta_ts_lo, min(delta_ts_hi, n+1))), aux (terms [delta_mid_lo,
 * min(delta_mid_hi, n_aux))), each points at the ring element of the FIRST term of its range. h_aux_kind is indexed by the
 * absolute auxiliary index as in rsg_groth16_prove. */
int rsg_groth16_lincombs(rsg_context *ctx, const rsg_crs *crs, const rsg_groth16_layout *layout, size_t n, size_t n_aux,
                         const uint64_t *const d_vec[6], const uint8_t *h_aux_kind, uint64_t *h_proof, uint64_t *d_proof,
                         size_t *n_used);
/* ---- the same on G term shards with the reference's ORDER-DEPENDENT rule kept exact (seal_ring.tcc:493-504: a running sum
 * whose c1 vanishes is dropped) ----
 * rsg_groth16_lincombs_shard: this rank's partial proof (3 encodings, the plain modular sums over its term ranges) into d_part
 * and, behind it, a PROBE BLOCK of rsg_groth16_shard_block_words(L_R, pstride) words: the running sums of every inner product
 * at one fixed NTT slot (c1, limb 0, x = 0) over this rank's live terms, their totals, the live-term counts and rank 0's
 * alpha / beta words.  pstride >= the largest term range of any rank.  No local fallback: the ranks all-gather
 * [d_part | block] and EVERY rank calls rsg_groth16_shard_check on the gathered blocks (host words): it shifts each rank's
 * running sums by the totals of the ranks before it -- the global prefix sums -- and replays the operator+= chains of
 * groth16.tcc:89-112 at that slot.  *verdict = 0: no prefix vanished anywhere, the modular sum of the G partial proofs IS the
 * reference's proof.  *verdict = 1 (structured CRS only, e.g. the tiny_transp golden case): the ranks run the exact CHAIN instead:
 *   rank 0 .. G-1 in order:  rsg_groth16_lincombs_chain(carry from the rank before) -> carry for the next rank
 *   rank 0 (holds alpha, beta):  rsg_groth16_chain_finish(last carry) -> the proof
 * carry = the six inner products <s_pows,A_io>, <s_pows,A_mid>, <s_pows,B_io>, <s_pows,B_mid>, <delta_ts,H>, <delta_mid,aux> over
 * all terms so far (6 encodings, device) + present[6] (0 = still the empty EncodingElem).  A rank enters its carry as the first
 * term of each inner product (coefficient 1, seal_ring.tcc:525-528), so prefixes and drops are the global ones.  The arena
 * must hold 6 spare encodings from index carry_first on (the carry is staged there). */
#define RSG_SHARD_BLOCK_HEADER 8   /* words: [0] flags (bit 0: static plan not applicable -> chain), [1..6] live terms per inner product, [7] pstride */
size_t rsg_groth16_shard_block_words(size_t L_R, size_t pstride);
int rsg_groth16_lincombs_shard(rsg_context *ctx, const rsg_crs *crs, const rsg_groth16_layout *layout, size_t n, size_t n_aux,
                               const uint64_t *const d_vec[6], const uint8_t *h_aux_kind, uint64_t *d_part, size_t pstride,
                               size_t *n_used);
/* Pure host arithmetic (no device, no context): h_blocks = world blocks back to back, Q0 = first encoding prime. */
int rsg_groth16_shard_check(const uint64_t *h_blocks, size_t world, size_t L_R, size_t pstride, uint64_t Q0, int *verdict);
int rsg_groth16_lincombs_chain(rsg_context *ctx, rsg_crs *crs, size_t carry_first, const rsg_groth16_layout *layout, size_t n,
                               size_t n_aux, const uint64_t *const d_vec[6], const uint8_t *h_aux_kind, uint64_t *d_carry,
                               uint8_t *h_present);
int rsg_groth16_chain_finish(rsg_context *ctx, const rsg_crs *crs, const rsg_groth16_layout *layout, const uint64_t *d_carry,
                             const uint8_t *h_present, uint64_t *d_proof);
/* ---- the N-GPU driver's two exchange steps over NVLink peer memory (csrc/p2p.cuh): the host driver maps every rank's buffers
 * into every process (CUDA IPC / symmetric memory) and orders the launches with device-side barriers; h_peer_* are `world` device
 * pointers (rank order).
 * rsg_exchange_p2p (on the rank's WITNESS context, N_R = slots per rank): this rank's slot block of the witness rows d_wit
 *   ([7n+2][L_R*S]: A_io|B_io|C_io|A_mid|B_mid|C_mid (n each) | H (n+1) | a zero row) goes straight into the term owners'
 *   coefficient buffers h_peer_full[d] = [5 (A_io, A_mid, B_io, B_mid, H)][per][L_R][world*S] -- the slots<->terms all-to-all.
 * rsg_enc_sum_p2p: rank `rank` sums slice `rank` of the world records [n_enc encodings | block_words] at h_peer_parts and stores it
 *   into every rank's h_peer_final (reduce-scatter + all-gather of the modular sum in one launch); the probe blocks are copied
 *   to d_blocks ([world][block_words]) for rsg_groth16_shard_check. */
int rsg_exchange_p2p(rsg_context *ctx, const uint64_t *d_wit, size_t n, size_t world, size_t rank, size_t per,
                     uint64_t *const *h_peer_full);
int rsg_enc_sum_p2p(rsg_context *ctx, uint64_t *const *h_peer_parts, uint64_t *const *h_peer_final, size_t world, size_t rank,
                    size_t n_enc, size_t block_words, uint64_t *d_blocks);
/* A non-owning rsg_ringvec over caller-owned device memory (e.g. a torch tensor): n_elems ring elements at d_words. */
int rsg_ringvec_wrap(rsg_context *ctx, uint64_t *d_words, size_t n_elems, rsg_ringvec **out);

/* ---- low-level entry points used by tests and by the multi-GPU driver (device pointers) ---- */
/* BatchEncoder::encode (batchencoder.cpp:110-149): count elements [L_R][N_R] -> plaintext coeffs [count][L_R][N_E] */
int rsg_batch_encode(rsg_context *ctx, const uint64_t *d_ring, size_t count, uint64_t *d_plain);
/* Evaluator::transform_to_ntt_inplace (evaluator.cpp:2174-2265): [count][L_R][N_E] -> [count][L_R][L_E][N_E] */
int rsg_plain_to_ntt(rsg_context *ctx, const uint64_t *d_plain, size_t count, uint64_t *d_plain_ntt);
/* util::ntt_negacyclic_harvey / inverse_ntt_negacyclic_harvey (util/ntt.cpp:407-474) over one prime of the context:
 * which = 0 -> Q[idx], which = 1 -> q[idx]; batch polynomials of N_E words, in place. */
int rsg_ntt(rsg_context *ctx, uint64_t *d_data, size_t batch, int which, size_t idx, int inverse);
/* The streaming multiply-accumulate alone: d_plain_ntt indexed by h_pidx[i] (or ~0u for RSG_TERM_ONE terms). */
int rsg_crs_lincomb(rsg_context *ctx, const uint64_t *d_crs, const uint32_t *h_term, const uint32_t *h_pidx,
                    size_t n_terms, const uint64_t *d_plain_ntt, uint64_t *d_out);
/* Timing hook: device milliseconds of the most recent kernels, by name, measured with CUDA events on the
 * context's stream when profiling is enabled. */
/* Host-side API trace (environment RSG_TRACE=1): wall time and call count per entry point, process-wide.  Writes a text
 * table into buf (nullable) and returns the size needed; reset != 0 clears the counters. */
size_t rsg_trace_report(char *buf, size_t cap, int reset);
int rsg_context_enable_timing(rsg_context *ctx, int on);
int rsg_context_last_timing(rsg_context *ctx, const char *kernel, float *ms, uint64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* RSGPU_H */
