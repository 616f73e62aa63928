/*
 * rs_oracle.c -- CPU restatement of the reference's prover hot path, in plain C.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; the product (ringsnark_b200/librsgpu.so) never links, loads or calls it.
 *
 * PARITY PINNED: every function here is checked in tests/test_oracle.py against outputs of the unmodified
 * reference (SEAL 4.1.1 + SEAL-Polytools + ringSNARK headers compiled from /root/reference by
 * oracle/Makefile.ref), committed as .rsgv files under tests/golden, and against SEAL's own NTT known-answer test.
 *
 * Each function cites the reference file:line it restates (paths relative to /root/reference).
 * Arithmetic is deliberately naive (unsigned __int128 and %): the reference always returns canonical
 * residues in [0, p), so only the mathematical function matters.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

static inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t p) { return (uint64_t)(((u128)a * b) % p); }
static inline uint64_t addmod(uint64_t a, uint64_t b, uint64_t p) { uint64_t s = a + b; return s >= p ? s - p : s; }
static inline uint64_t submod(uint64_t a, uint64_t b, uint64_t p) { return a >= b ? a - b : a + p - b; }
static uint64_t powmod(uint64_t a, uint64_t e, uint64_t p) {
  uint64_t r = 1 % p;
  a %= p;
  while (e) { if (e & 1) r = mulmod(r, a, p); a = mulmod(a, a, p); e >>= 1; }
  return r;
}
/* depends/SEAL/native/src/seal/util/uintarithsmallmod.h (try_invert_uint_mod): inverse or failure. p prime here. */
int ro_try_invert(uint64_t a, uint64_t p, uint64_t *inv) {
  a %= p;
  if (a == 0) return 0;
  *inv = powmod(a, p - 2, p);
  return 1;
}
uint64_t ro_mulmod(uint64_t a, uint64_t b, uint64_t p) { return mulmod(a % p, b % p, p); }

static uint32_t reverse_bits(uint32_t x, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
}
static int ilog2(size_t n) { int l = 0; while (((size_t)1 << l) < n) l++; return l; }

/*
 * Minimal primitive 2N-th root of unity mod p.
 * depends/SEAL/native/src/seal/util/numth.cpp:386-412 (try_minimal_primitive_root): take any primitive root r of
 * order `degree`, then walk r, r^3, r^5, ... (all primitive degree-th roots) and keep the smallest.
 * numth.cpp:353-384 finds the starting root by random sampling; the minimum does not depend on that choice.
 */
uint64_t ro_minimal_primitive_root(uint64_t degree, uint64_t p) {
  if ((p - 1) % degree != 0) return 0;
  uint64_t root = 0;
  for (uint64_t g = 2; g < p; g++) {
    uint64_t r = powmod(g, (p - 1) / degree, p);
    if (powmod(r, degree / 2, p) == p - 1) { root = r; break; }
  }
  uint64_t gen_sq = mulmod(root, root, p), cur = root, best = root;
  for (uint64_t i = 0; i < degree / 2; i++) {
    if (cur < best) best = cur;
    cur = mulmod(cur, gen_sq, p);
  }
  return best;
}

/* depends/SEAL/native/src/seal/util/ntt.cpp:240-299 (NTTTables::initialize): psi powers in bit-reversed slots;
 * inverse powers in the "reverse_bits(i-1)+1" order consumed sequentially by transform_from_rev. */
static void build_tables(int logn, uint64_t p, uint64_t *roots, uint64_t *inv_roots, uint64_t *inv_n) {
  size_t n = (size_t)1 << logn;
  uint64_t psi = ro_minimal_primitive_root(2 * n, p), ipsi = 0;
  ro_try_invert(psi, p, &ipsi);
  uint64_t pw = psi;
  for (size_t i = 1; i < n; i++) { roots[reverse_bits((uint32_t)i, logn)] = pw; pw = mulmod(pw, psi, p); }
  roots[0] = 1;
  pw = ipsi;
  for (size_t i = 1; i < n; i++) { inv_roots[reverse_bits((uint32_t)(i - 1), logn) + 1] = pw; pw = mulmod(pw, ipsi, p); }
  inv_roots[0] = 1;
  ro_try_invert((uint64_t)n % p, p, inv_n);
}

/* Forward negacyclic NTT, natural order in -> bit-reversed order out, canonical residues.
 * depends/SEAL/native/src/seal/util/ntt.cpp:407-436 + util/dwthandler.h:94-190 (transform_to_rev, Cooley-Tukey,
 * roots consumed sequentially from index 1). */
void ro_ntt_forward(uint64_t *a, size_t n, uint64_t p) {
  int logn = ilog2(n);
  uint64_t *roots = malloc(n * 8), *iroots = malloc(n * 8), inv_n;
  build_tables(logn, p, roots, iroots, &inv_n);
  size_t gap = n >> 1, ri = 0;
  for (size_t m = 1; m < n; m <<= 1, gap >>= 1) {
    size_t off = 0;
    for (size_t i = 0; i < m; i++, off += gap << 1) {
      uint64_t r = roots[++ri];
      for (size_t j = 0; j < gap; j++) {
        uint64_t u = a[off + j], v = mulmod(a[off + j + gap], r, p);
        a[off + j] = addmod(u, v, p);
        a[off + j + gap] = submod(u, v, p);
      }
    }
  }
  free(roots); free(iroots);
}

/* Inverse negacyclic NTT, bit-reversed in -> natural out, scaled by n^-1.
 * depends/SEAL/native/src/seal/util/ntt.cpp:452-474 + util/dwthandler.h:202-330 (transform_from_rev,
 * Gentleman-Sande; the n^-1 factor is merged into the last stage, same product). */
void ro_ntt_inverse(uint64_t *a, size_t n, uint64_t p) {
  int logn = ilog2(n);
  uint64_t *roots = malloc(n * 8), *iroots = malloc(n * 8), inv_n;
  build_tables(logn, p, roots, iroots, &inv_n);
  size_t gap = 1, ri = 0;
  for (size_t m = n >> 1; m >= 1; m >>= 1, gap <<= 1) {
    size_t off = 0;
    for (size_t i = 0; i < m; i++, off += gap << 1) {
      uint64_t r = iroots[++ri];
      for (size_t j = 0; j < gap; j++) {
        uint64_t u = a[off + j], v = a[off + j + gap];
        a[off + j] = addmod(u, v, p);
        a[off + j + gap] = mulmod(submod(u, v, p), r, p);
      }
    }
    if (m == 1) break;
  }
  for (size_t i = 0; i < n; i++) a[i] = mulmod(a[i], inv_n, p);
  free(roots); free(iroots);
}

/* depends/SEAL/native/src/seal/batchencoder.cpp:64-88 (populate_matrix_reps_index_map): slot k of the batched
 * vector lives at coefficient index map[k] of the NTT-domain plaintext; generator 3 of (Z/2N)^*. */
void ro_batch_index_map(size_t N, uint64_t *map) {
  int logn = ilog2(N);
  size_t row = N >> 1, m = N << 1;
  uint64_t gen = 3, pos = 1;
  for (size_t i = 0; i < row; i++) {
    uint64_t i1 = (pos - 1) >> 1, i2 = (m - pos - 1) >> 1;
    map[i] = reverse_bits((uint32_t)i1, logn);
    map[row | i] = reverse_bits((uint32_t)i2, logn);
    pos = (pos * gen) & (m - 1);
  }
}

/* depends/SEAL/native/src/seal/batchencoder.cpp:110-149 (BatchEncoder::encode(vector<uint64_t>)):
 * values[k] -> coefficient map[k], the other N_E - nvals slots zero, then inverse NTT mod t. */
void ro_batch_encode(const uint64_t *vals, size_t nvals, size_t N_E, uint64_t t, uint64_t *out) {
  uint64_t *map = malloc(N_E * 8);
  ro_batch_index_map(N_E, map);
  memset(out, 0, N_E * 8);
  for (size_t k = 0; k < nvals; k++) out[map[k]] = vals[k];
  ro_ntt_inverse(out, N_E, t);
  free(map);
}

/* depends/SEAL/native/src/seal/evaluator.cpp:2174-2265 (transform_to_ntt_inplace(Plaintext)):
 * centred lift v -> v (v < ceil(t/2), context.cpp:329) or v + (Q - t), reduced into every Q_l
 * (fast path evaluator.cpp:2243-2259, slow multi-word path :2220-2242; both equal the expression below),
 * then one forward NTT per Q_l.  out is [L_E][N_E]. */
void ro_plain_lift_ntt(const uint64_t *plain, size_t N_E, uint64_t t, const uint64_t *Q, size_t L_E, uint64_t *out) {
  uint64_t thr = (t + 1) >> 1;
  for (size_t l = 0; l < L_E; l++) {
    uint64_t q = Q[l], tm = t % q;
    uint64_t *o = out + l * N_E;
    for (size_t i = 0; i < N_E; i++) {
      uint64_t v = plain[i], r = v % q;
      if (v >= thr) r = submod(r, tm, q); /* v + (Q - t) == v - t (mod Q_l) */
      o[i] = r;
    }
    ro_ntt_forward(o, N_E, q);
  }
}

/* depends/SEAL-Polytools/src/poly_arith.cpp:147-153 (SealPoly::is_zero), bug included:
 * `*mm == 0 && !memcmp(mm, mm + 1, size - 1)` compares BYTES although size counts words, so it proves only
 * that bytes [0, size + 7) are zero. */
int ro_is_zero_quirk(const uint64_t *w, size_t size) {
  if (size == 0) return 1;
  if (w[0] != 0) return 0;
  return memcmp(w, w + 1, size - 1) == 0;
}

/* depends/SEAL-Polytools/src/poly_arith.cpp:155-162 (SealPoly::is_equal): memcmp over size BYTES. */
int ro_is_equal_quirk(const uint64_t *a, const uint64_t *b, size_t size) { return memcmp(a, b, size) == 0; }

/*
 * EncodingElem::inner_product, ringsnark/seal/seal_ring.tcc:361-433 with operator*= :509-548 and operator+= :479-507,
 * on flat words.  crs: [T][L_R][2][L_E][N_E]; coeff: [T][L_R][N_R] (dense words of to_poly());
 * tag[i]: 0 = skip (is_zero() true), 1 = scalar one (ciphertext used unchanged, :525-528), 2 = general.
 * out: [L_R][2][L_E][N_E].  Returns the number of summed terms (0 => the reference returns an EMPTY element).
 * Per general term and ring limb j: batch-encode limb j mod q_j (a11), lift + NTT (a12/a13), dyadic product with
 * both ciphertext polynomials (a14, util/polyarithsmallmod.cpp:226-284), modular add (a15, evaluator.cpp:217-231).
 */
/* SEAL refuses "transparent" results (c1 identically zero, ciphertext.h:451-456; evaluator.cpp:233-239) and
 * EncodingElem::operator+= answers by replacing that ring limb with an empty zero ciphertext (seal_ring.tcc:493-504):
 * on flat words, the limb's c0 is dropped as well.  Returns 1 if the limb was transparent. */
static int drop_if_transparent(uint64_t *ct, size_t N_E, size_t L_E) {
  for (size_t w = L_E * N_E; w < 2 * L_E * N_E; w++)
    if (ct[w]) return 0;
  memset(ct, 0, 2 * L_E * N_E * 8);
  return 1;
}

size_t ro_inner_product(const uint64_t *crs, const uint64_t *coeff, const uint8_t *tag, size_t T, size_t N_R, size_t L_R,
                        const uint64_t *q, size_t N_E, size_t L_E, const uint64_t *Q, uint64_t *out) {
  size_t per_ct = 2 * L_E * N_E, per_enc = L_R * per_ct, used = 0;
  uint64_t *plain = malloc(N_E * 8), *pntt = malloc(L_E * N_E * 8);
  memset(out, 0, per_enc * 8);
  for (size_t i = 0; i < T; i++) {
    if (tag[i] == 0) continue;
    used++;
    for (size_t j = 0; j < L_R; j++) {
      const uint64_t *ct = crs + i * per_enc + j * per_ct;
      uint64_t *acc = out + j * per_ct;
      if (tag[i] == 2) {
        ro_batch_encode(coeff + (i * L_R + j) * N_R, N_R, N_E, q[j], plain);
        ro_plain_lift_ntt(plain, N_E, q[j], Q, L_E, pntt);
      }
      for (size_t k = 0; k < 2; k++)
        for (size_t l = 0; l < L_E; l++) {
          const uint64_t *c = ct + (k * L_E + l) * N_E;
          uint64_t *a = acc + (k * L_E + l) * N_E;
          for (size_t x = 0; x < N_E; x++) {
            uint64_t term = tag[i] == 2 ? mulmod(c[x], pntt[l * N_E + x], Q[l]) : c[x];
            a[x] = addmod(a[x], term, Q[l]);
          }
        }
      /* the first summed term is a copy (seal_ring.tcc:485-488); every later one goes through add_inplace */
      if (used > 1) drop_if_transparent(acc, N_E, L_E);
    }
  }
  free(plain); free(pntt);
  return used;
}

/* EncodingElem::operator+= on two non-empty elements: evaluator.cpp:217-231 (add_poly_coeffmod per limb). */
void ro_enc_add(uint64_t *acc, const uint64_t *other, size_t L_R, size_t N_E, size_t L_E, const uint64_t *Q) {
  for (size_t j = 0; j < L_R; j++)
    for (size_t k = 0; k < 2; k++)
      for (size_t l = 0; l < L_E; l++) {
        size_t off = ((j * 2 + k) * L_E + l) * N_E;
        for (size_t x = 0; x < N_E; x++) acc[off + x] = addmod(acc[off + x], other[off + x], Q[l]);
      }
  for (size_t j = 0; j < L_R; j++) drop_if_transparent(acc + j * 2 * L_E * N_E, N_E, L_E);
}

/* ---------------------------------------------------------------------------------------------------------
 * Witness map, one slot at a time (every slot of every ring limb is an independent copy of the same
 * computation over Z_p).
 * ------------------------------------------------------------------------------------------------------- */

/* ringsnark/util/polynomials.tcc:9-43 (interpolate) on the domain x_j = j (util/evaluation_domain.tcc:7-13):
 * s[] = coefficients of the master polynomial, phi_j = Z'(x_j) by Horner, f = y_j / phi_j, synthetic division. */
static void interpolate_slot(size_t n, const uint64_t *y, uint64_t p, uint64_t *coeffs, uint64_t *s) {
  for (size_t k = 0; k < n; k++) { coeffs[k] = 0; s[k] = 0; }
  s[n - 1] = submod(0, 0 % p, p); /* -x[0] = 0 */
  for (size_t i = 1; i < n; i++) {
    uint64_t xi = i % p;
    for (size_t j = n - i - 1; j < n - 1; j++) s[j] = submod(s[j], mulmod(xi, s[j + 1], p), p);
    s[n - 1] = submod(s[n - 1], xi, p);
  }
  for (size_t j = 0; j < n; j++) {
    uint64_t xj = j % p, phi = n % p;
    for (size_t k = n - 1; k > 0; k--) { phi = mulmod(phi, xj, p); phi = addmod(phi, mulmod(s[k], k % p, p), p); }
    uint64_t inv = 0;
    ro_try_invert(phi, p, &inv);
    uint64_t ff = mulmod(y[j], inv, p), b = 1 % p;
    for (size_t k = n; k-- > 0;) {
      coeffs[k] = addmod(coeffs[k], mulmod(b, ff, p), p);
      b = mulmod(b, xj, p);
      b = addmod(b, s[k], p);
    }
  }
}

/* ringsnark/util/evaluation_domain.tcc:53-60 (vanishing_polynomial): Z(x) = prod_{i<n} (x - i); n+1 coefficients. */
void ro_vanishing(size_t n, uint64_t p, uint64_t *Z) {
  memset(Z, 0, (n + 1) * 8);
  Z[0] = 1 % p; /* running product, degree i after i factors */
  for (size_t i = 0; i < n; i++) {
    uint64_t neg = submod(0, i % p, p);
    for (size_t k = i + 2; k-- > 0;) { /* new[k] = old[k-1] - i * old[k] */
      uint64_t lower = k ? Z[k - 1] : 0;
      uint64_t same = k <= i ? mulmod(Z[k], neg, p) : 0;
      Z[k] = addmod(same, lower, p);
    }
  }
}

/* One interpolation for a whole vector of ring elements. y, coeffs: [n][L_R][N_R]. */
void ro_interpolate(size_t n, const uint64_t *y, size_t N_R, size_t L_R, const uint64_t *q, uint64_t *coeffs) {
  size_t W = N_R * L_R;
#pragma omp parallel
  {
    uint64_t *ys = malloc(n * 8), *cs = malloc(n * 8), *s = malloc(n * 8);
#pragma omp for schedule(static)
    for (size_t w = 0; w < W; w++) {
      uint64_t p = q[w / N_R];
      for (size_t k = 0; k < n; k++) ys[k] = y[k * W + w];
      interpolate_slot(n, ys, p, cs, s);
      for (size_t k = 0; k < n; k++) coeffs[k * W + w] = cs[k];
    }
    free(ys); free(cs); free(s);
  }
}

/*
 * r1cs_to_qrp_witness_map, ringsnark/reductions/r1cs_to_qrp/r1cs_to_qrp.tcc:148-259, non-ZK call
 * (d1 = d2 = d3 = 0 as in groth16.tcc:82-84), values only:
 *   aA, aB, aC = interpolate(full evaluations)            :216-223
 *   prod = aA * aB (schoolbook, polynomials.tcc:61-66), diff = prod - aC (:68-73, :237-243)
 *   H = diff / Z  (long division by the monic Z, polynomials.tcc:75-81, evaluation_domain.tcc:80-84)
 * lenA, lenB, lenC are the operand lengths AFTER Boost's normalize() as the reference sees them (trailing elements
 * for which RingElem::operator== says "== 0", i.e. the is_equal prefix quirk) -- computed by the caller with
 * ro_is_equal_quirk on whole ring elements; trailing coefficients beyond them are treated as zero.
 * yfull: [3][n][W]; H out: [n-1][W] (coefficients the reference adds into coefficients_for_H[0..n-2]).
 */
void ro_witness_H(size_t n, const uint64_t *aA, const uint64_t *aB, const uint64_t *aC, size_t lenA, size_t lenB,
                  size_t lenC, size_t N_R, size_t L_R, const uint64_t *q, uint64_t *H, size_t *H_len_out) {
  size_t W = N_R * L_R;
  size_t Hn = n >= 1 ? n - 1 : 0;
  memset(H, 0, Hn * W * 8);
  size_t lenP = (lenA && lenB) ? lenA + lenB - 1 : 0;
  size_t lenD = lenP > lenC ? lenP : lenC;
  /* quotient length: |diff| - |Z| + 1 with |Z| = n + 1, or empty when |diff| < |Z| */
  size_t lenQ = lenD >= n + 1 ? lenD - (n + 1) + 1 : 0;
  if (H_len_out) *H_len_out = lenQ;
  if (!lenQ) return;
#pragma omp parallel
  {
    uint64_t *a = malloc(n * 8), *b = malloc(n * 8), *u = malloc((2 * n + 1) * 8), *Z = malloc((n + 1) * 8);
    uint64_t lastp = 0;
#pragma omp for schedule(static)
    for (size_t w = 0; w < W; w++) {
      uint64_t p = q[w / N_R];
      if (p != lastp) { ro_vanishing(n, p, Z); lastp = p; }
      for (size_t k = 0; k < n; k++) { a[k] = k < lenA ? aA[k * W + w] : 0; b[k] = k < lenB ? aB[k * W + w] : 0; }
      memset(u, 0, (2 * n + 1) * 8);
      for (size_t i = 0; i < lenB; i++)
        for (size_t j = 0; j < lenA; j++) u[i + j] = addmod(u[i + j], mulmod(a[j], b[i], p), p);
      for (size_t k = 0; k < lenC; k++) u[k] = submod(u[k], aC[k * W + w], p);
      /* Knuth long division by monic Z (degree n): q[k] = u[n+k]; u[j] -= q[k] Z[j-k] */
      for (size_t k = lenQ; k-- > 0;) {
        uint64_t qk = u[n + k];
        if (k < Hn) H[k * W + w] = qk;
        for (size_t j = n + k; j-- > k;) u[j] = submod(u[j], mulmod(qk, Z[j - k], p), p);
      }
    }
    free(a); free(b); free(u); free(Z);
  }
}

/*
 * r1cs_to_qrp_instance_map_with_evaluation, ringsnark/reductions/r1cs_to_qrp/r1cs_to_qrp.tcc:75-116, values only, as the
 * reference computes it:
 *   u[j] = prod_{i != j} (t - x_i) / prod_{i != j} (x_j - x_i)      evaluate_all_lagrange_polynomials,
 *                                                                   util/evaluation_domain.tcc:20-41 (O(m^2), one division)
 *   Zt   = prod_i (t - x_i)                                         compute_vanishing_polynomial, :43-51
 *   At[idx] += u[i] * coeff  for every term of constraint i's a (likewise b, c)     r1cs_to_qrp.tcc:92-107
 *   Ht[i] = t^i, i = 0..m                                           :109-113
 * on the domain x_i = i (get_evaluation_domain).  t: [L_R][N_R]; CSR system as in ro-side tests: rows m*n + i, col 0 = constant
 * wire, uint64 coefficients reduced mod q_j (multiply_poly_scalar_coeffmod semantics).  ABCt: [3][nvars1][W], Ht: [n+1][W],
 * Zt: [W].  Returns 0, or -1 if some slot of t hits a domain point (the reference needs the divisions to exist).
 */
int ro_instance_map(size_t n, size_t nvars1, const uint32_t *row_ptr, const uint32_t *col, const uint64_t *coeff,
                    const uint64_t *t, size_t N_R, size_t L_R, const uint64_t *q, uint64_t *ABCt, uint64_t *Ht, uint64_t *Zt) {
  size_t W = N_R * L_R;
  int bad = 0;
  memset(ABCt, 0, 3 * nvars1 * W * 8);
#pragma omp parallel
  {
    uint64_t *u = malloc((n ? n : 1) * 8);
#pragma omp for schedule(static)
    for (size_t w = 0; w < W; w++) {
      uint64_t p = q[w / N_R], tv = t[w];
      for (size_t j = 0; j < n; j++) {
        uint64_t num = 1, den = 1;
        for (size_t i = 0; i < n; i++) {
          if (i == j) continue;
          num = mulmod(num, submod(tv, i % p, p), p);
          den = mulmod(den, submod(j % p, i % p, p), p);
        }
        uint64_t inv;
        if (!ro_try_invert(den, p, &inv)) { bad = 1; inv = 0; }
        u[j] = mulmod(num, inv, p);
      }
      uint64_t z = 1, pw = 1;
      for (size_t i = 0; i < n; i++) z = mulmod(z, submod(tv, i % p, p), p);
      Zt[w] = z;
      for (size_t i = 0; i <= n; i++) { Ht[i * W + w] = pw; pw = mulmod(pw, tv, p); }
      for (size_t m = 0; m < 3; m++)
        for (size_t i = 0; i < n; i++)
          for (size_t e = row_ptr[m * n + i]; e < row_ptr[m * n + i + 1]; e++) {
            uint64_t *dst = ABCt + (m * nvars1 + col[e]) * W + w;
            *dst = addmod(*dst, mulmod(u[i], coeff[e] % p, p), p);
          }
    }
    free(u);
  }
  return bad ? -1 : 0;
}

/*
 * EncodingElem::decode, ringsnark/seal/seal_ring.tcc:435-477, for ONE ring limb (one BGV ciphertext of size 2, NTT form,
 * first level, correction factor 1), values only:
 *   budget  Decryptor::invariant_noise_budget   depends/SEAL/native/src/seal/decryptor.cpp:383-461
 *           (c0 + c1 s, inverse NTT, CRT-compose, centred infinity norm, bitcount(Q) - bitcount(norm) - 1, floored at 0)
 *   plain   Decryptor::bgv_decrypt               decryptor.cpp:189-231 with RNSTool::decrypt_modt = BaseConverter::
 *           exact_convert_array, util/rns.cpp:466-539: temp_l = x_l (Q/Q_l)^-1 mod Q_l, v = (uint64)(sum_l double(temp_l) /
 *           double(Q_l) + 0.5) summed in limb order, out = sum_l temp_l (Q/Q_l mod t) - v (Q mod t)  mod t
 *   slots   BatchEncoder::decode                 batchencoder.cpp:278-315: forward NTT mod t, out[i] = temp[index_map[i]]
 * ct: [2][L_E][N_E]; sk: [L_E][N_E] (NTT form); out: N_R slot values (seal_ring.tcc:466 keeps the first N_R).
 * Returns the noise budget in bits.  Big integers: little-endian 64-bit words, at most 16 limbs.
 */
int ro_decode_limb(const uint64_t *ct, const uint64_t *sk, size_t N_E, size_t L_E, const uint64_t *Q, uint64_t t, size_t N_R,
                   uint64_t *out) {
  enum { MAXW = 16 };
  uint64_t *phase = malloc(L_E * N_E * 8), *plain = malloc(N_E * 8), *map = malloc(N_E * 8);
  for (size_t l = 0; l < L_E; l++) {
    for (size_t i = 0; i < N_E; i++)
      phase[l * N_E + i] = addmod(mulmod(ct[(L_E + l) * N_E + i], sk[l * N_E + i], Q[l]), ct[l * N_E + i], Q[l]);
    ro_ntt_inverse(phase + l * N_E, N_E, Q[l]);
  }
  uint64_t inv_punct[MAXW], punct_mod_t[MAXW], Q_mod_t = 1, Qw[MAXW + 1] = {1}, halfw[MAXW];
  for (size_t l = 0; l < L_E; l++) {
    uint64_t pr = 1, pm = 1;
    for (size_t k = 0; k < L_E; k++)
      if (k != l) { pr = mulmod(pr, Q[k] % Q[l], Q[l]); pm = mulmod(pm, Q[k] % t, t); }
    ro_try_invert(pr, Q[l], &inv_punct[l]);
    punct_mod_t[l] = pm;
    Q_mod_t = mulmod(Q_mod_t, Q[l] % t, t);
    uint64_t carry = 0;
    for (size_t w = 0; w < MAXW; w++) { u128 x = (u128)Qw[w] * Q[l] + carry; Qw[w] = (uint64_t)x; carry = (uint64_t)(x >> 64); }
  }
  int Qbits = 0;
  for (int w = MAXW - 1; w >= 0; w--) if (Qw[w]) { Qbits = w * 64 + 64 - __builtin_clzll(Qw[w]); break; }
  { uint64_t tmp[MAXW + 1], carry = 1;
    for (size_t w = 0; w < MAXW; w++) { tmp[w] = Qw[w] + carry; carry = tmp[w] < carry; }
    tmp[MAXW] = carry;
    for (size_t w = 0; w < MAXW; w++) halfw[w] = (tmp[w] >> 1) | (tmp[w + 1] << 63); }
  int norm_bits = 0;
  for (size_t i = 0; i < N_E; i++) {
    uint64_t temp[MAXW];
    double agg = 0.0;
    u128 acc = 0;
    for (size_t l = 0; l < L_E; l++) {
      temp[l] = mulmod(phase[l * N_E + i], inv_punct[l], Q[l]);
      agg += (double)temp[l] / (double)Q[l];
      acc = (acc + (u128)mulmod(temp[l] % t, punct_mod_t[l], t)) % t;
    }
    agg += 0.5;
    uint64_t v = (uint64_t)agg;
    plain[i] = submod((uint64_t)acc, mulmod(v % t, Q_mod_t, t), t);
    /* CRT-compose: X = sum_l temp_l * (Q / Q_l)  mod Q  -- here through mixed radix digits (same integer in [0, Q)) */
    uint64_t a[MAXW], X[MAXW] = {0};
    for (size_t l = 0; l < L_E; l++) {
      uint64_t x = phase[l * N_E + i];
      for (size_t k = 0; k < l; k++) {
        uint64_t inv;
        ro_try_invert(Q[k] % Q[l], Q[l], &inv);
        x = mulmod(submod(x, a[k] % Q[l], Q[l]), inv, Q[l]);
      }
      a[l] = x;
    }
    X[0] = a[L_E - 1];
    for (int l = (int)L_E - 2; l >= 0; l--) {
      uint64_t carry = a[l];
      for (size_t w = 0; w < L_E; w++) { u128 x = (u128)X[w] * Q[l] + carry; X[w] = (uint64_t)x; carry = (uint64_t)(x >> 64); }
    }
    int ge = 1;
    for (int w = (int)L_E - 1; w >= 0; w--) if (X[w] != halfw[w]) { ge = X[w] > halfw[w]; break; }
    if (ge) {
      uint64_t borrow = 0;
      for (size_t w = 0; w < L_E; w++) {
        uint64_t d = Qw[w] - X[w] - borrow;
        borrow = (Qw[w] < X[w]) || (Qw[w] == X[w] && borrow);
        X[w] = d;
      }
    }
    for (int w = (int)L_E - 1; w >= 0; w--)
      if (X[w]) { int nb = w * 64 + 64 - __builtin_clzll(X[w]); if (nb > norm_bits) norm_bits = nb; break; }
  }
  ro_ntt_forward(plain, N_E, t);
  ro_batch_index_map(N_E, map);
  for (size_t i = 0; i < N_R; i++) out[i] = plain[map[i]];
  free(phase); free(plain); free(map);
  int budget = Qbits - norm_bits - 1;
  return budget < 0 ? 0 : budget;
}

/* ======================================================================================================================
 * EncodingElem::encode (ringsnark/seal/seal_ring.tcc:324-359): BatchEncoder::encode + Encryptor::encrypt_symmetric
 * (depends/SEAL/native/src/seal/encryptor.cpp:242-312 -> util/rlwe.cpp:276-390) for one (ring element, ring limb).
 * Randomness is SEAL's Blake2xbPRNG (randomgen.cpp:201-211): the byte stream of a seed is the concatenation over
 * counter = 0, 1, 2, ... of blake2xb(outlen = 4096, in = counter as 8 LE bytes, key = the 64 seed bytes)
 * (util/blake2xb.c: root hash with the XOF parameter block, then one BLAKE2b call per 64 output bytes with node_offset = block
 * index; util/blake2b.c is RFC 7693).
 * ==================================================================================================================== */
static const uint64_t B2_IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                  0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
static const uint8_t B2_SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
static inline uint64_t rotr64(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }
/* one BLAKE2b compression: h (8 words) absorbs the 128-byte block m (16 LE words); t = bytes hashed so far incl. this block */
static void b2_compress(uint64_t h[8], const uint64_t m[16], uint64_t t, int last) {
  uint64_t v[16];
  for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = B2_IV[i]; }
  v[12] ^= t;
  if (last) v[14] = ~v[14];
#define B2_G(a, b, c, d, x, y)          \
  v[a] = v[a] + v[b] + (x); v[d] = rotr64(v[d] ^ v[a], 32); v[c] = v[c] + v[d]; v[b] = rotr64(v[b] ^ v[c], 24); \
  v[a] = v[a] + v[b] + (y); v[d] = rotr64(v[d] ^ v[a], 16); v[c] = v[c] + v[d]; v[b] = rotr64(v[b] ^ v[c], 63);
  for (int r = 0; r < 12; r++) {
    const uint8_t *s = B2_SIGMA[r];
    B2_G(0, 4, 8, 12, m[s[0]], m[s[1]]) B2_G(1, 5, 9, 13, m[s[2]], m[s[3]]) B2_G(2, 6, 10, 14, m[s[4]], m[s[5]])
    B2_G(3, 7, 11, 15, m[s[6]], m[s[7]]) B2_G(0, 5, 10, 15, m[s[8]], m[s[9]]) B2_G(1, 6, 11, 12, m[s[10]], m[s[11]])
    B2_G(2, 7, 8, 13, m[s[12]], m[s[13]]) B2_G(3, 4, 9, 14, m[s[14]], m[s[15]])
  }
#undef B2_G
  for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}
/* one 4096-byte PRNG buffer: out[512] words = blake2xb(4096, counter, key = seed) */
void ro_blake2xb_buffer(const uint64_t seed[8], uint64_t counter, uint64_t *out) {
  uint64_t h[8], m[16], root[8];
  /* root: digest 64, key 64, fanout 1, depth 1, leaf 0 | node_offset 0, xof_length 4096 | node_depth 0, inner 0 */
  for (int i = 0; i < 8; i++) h[i] = B2_IV[i];
  h[0] ^= 0x40ull | (0x40ull << 8) | (1ull << 16) | (1ull << 24);
  h[1] ^= 4096ull << 32;
  for (int i = 0; i < 16; i++) m[i] = i < 8 ? seed[i] : 0;   /* the key, padded to one block */
  b2_compress(h, m, 128, 0);
  for (int i = 0; i < 16; i++) m[i] = 0;
  m[0] = counter;
  b2_compress(h, m, 136, 1);
  for (int i = 0; i < 8; i++) root[i] = h[i];
  for (uint64_t b = 0; b < 64; b++) {
    /* digest 64, key 0, fanout 0, depth 0, leaf_length 64 | node_offset b, xof_length 4096 | node_depth 0, inner_length 64 */
    for (int i = 0; i < 8; i++) h[i] = B2_IV[i];
    h[0] ^= 0x40ull | (64ull << 32);
    h[1] ^= b | (4096ull << 32);
    h[2] ^= 64ull << 8;
    for (int i = 0; i < 16; i++) m[i] = i < 8 ? root[i] : 0;
    b2_compress(h, m, 64, 1);
    for (int i = 0; i < 8; i++) out[b * 8 + i] = h[i];
  }
}
/* bytes [off, off + n) of the PRNG stream of `seed` */
void ro_prng_bytes(const uint64_t seed[8], uint64_t off, size_t n, uint8_t *out) {
  uint64_t buf[512];
  uint64_t have = (uint64_t)-1;
  for (size_t k = 0; k < n; k++) {
    const uint64_t pos = off + k, ctr = pos / 4096;
    if (ctr != have) { ro_blake2xb_buffer(seed, ctr, buf); have = ctr; }
    out[k] = ((const uint8_t *)buf)[pos % 4096];
  }
}
/* slots: N_R values mod t of ring limb j; sk: [L_E][N_E] secret key (NTT form, first L_E limbs); seed: what the context's
 * random generator factory hands to every PRNG it creates (randomgen.h:440-448); ct: [2][L_E][N_E]. */
void ro_encrypt_limb(const uint64_t *slots, size_t N_R, uint64_t t, size_t N_E, size_t L_E, const uint64_t *Q, const uint64_t *sk,
                     const uint64_t seed[8], uint64_t *ct) {
  uint64_t *plain = malloc(N_E * 8), *pntt = malloc(L_E * N_E * 8), *e = malloc(N_E * 8);
  ro_batch_encode(slots, N_R, N_E, t, plain);                    /* batchencoder.cpp:110-149 */
  ro_plain_lift_ntt(plain, N_E, t, Q, L_E, pntt);                /* encryptor.cpp:260-308: same lift as transform_to_ntt */
  /* rlwe.cpp:321-328: bootstrap PRNG -> 64-byte public seed -> ciphertext PRNG */
  uint64_t pub[8];
  ro_prng_bytes(seed, 0, 64, (uint8_t *)pub);
  /* rlwe.cpp:339 -> sample_poly_uniform (rlwe.cpp:106-131): bulk fill, then rejected words redrawn from the same stream */
  uint64_t *c0 = ct, *c1 = ct + L_E * N_E;
  ro_prng_bytes(pub, 0, L_E * N_E * 8, (uint8_t *)c1);
  uint64_t extra_off = (uint64_t)L_E * N_E * 8;
  for (size_t l = 0; l < L_E; l++) {
    const uint64_t max_multiple = 0xFFFFFFFFFFFFFFFFull - (0xFFFFFFFFFFFFFFFFull % Q[l]) - 1;
    for (size_t i = 0; i < N_E; i++) {
      uint64_t r = c1[l * N_E + i];
      while (r >= max_multiple) { ro_prng_bytes(pub, extra_off, 8, (uint8_t *)&r); extra_off += 8; }
      c1[l * N_E + i] = r % Q[l];
    }
  }
  /* rlwe.cpp:354 -> sample_poly_cbd (rlwe.cpp:68-104): 6 bytes of the BOOTSTRAP stream per coefficient */
  uint8_t *nb = malloc(6 * N_E);
  ro_prng_bytes(seed, 64, 6 * N_E, nb);
  for (size_t l = 0; l < L_E; l++) {
    for (size_t i = 0; i < N_E; i++) {
      const uint8_t *x = nb + 6 * i;
      const int noise = __builtin_popcount(x[0]) + __builtin_popcount(x[1]) + __builtin_popcount(x[2] & 0x1F) - __builtin_popcount(x[3]) -
                        __builtin_popcount(x[4]) - __builtin_popcount(x[5] & 0x1F);
      e[i] = noise < 0 ? Q[l] - (uint64_t)(-noise) : (uint64_t)noise;
    }
    ro_ntt_forward(e, N_E, Q[l]);                                /* rlwe.cpp:366 */
    const uint64_t tq = t % Q[l];
    for (size_t i = 0; i < N_E; i++) {
      uint64_t v = addmod(mulmod(sk[l * N_E + i], c1[l * N_E + i], Q[l]), mulmod(e[i], tq, Q[l]), Q[l]);   /* a s + t e */
      v = v ? Q[l] - v : 0;                                      /* rlwe.cpp:386 */
      c0[l * N_E + i] = addmod(v, pntt[l * N_E + i], Q[l]);      /* encryptor.cpp:310-311 */
    }
  }
  free(plain); free(pntt); free(e); free(nb);
}
