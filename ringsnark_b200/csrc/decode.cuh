// EncodingElem::decode on the device (the verifier's front half).  Restates, for first-level BGV
// ciphertexts of size 2 in NTT form with correction factor 1 (what the prover's inner products produce),
//   EncodingElem::decode                 ringsnark/seal/seal_ring.tcc:435-477
//   Decryptor::bgv_decrypt               depends/SEAL/native/src/seal/decryptor.cpp:189-231  (c0 + c1 s, inverse NTT, decrypt_modt)
//   BaseConverter::exact_convert_array   depends/SEAL/native/src/seal/util/rns.cpp:466-539   (base q -> t with the double-
//                                         precision rounding term: same operations in the same order, IEEE round-to-nearest)
//   Decryptor::invariant_noise_budget    decryptor.cpp:383-461  (CRT-compose, centred infinity norm, bit count)
//   BatchEncoder::decode                 batchencoder.cpp:278-315 (forward NTT mod t, read through the index map)
// Scratch layouts keep the polynomials of one modulus contiguous so that the raw NTT launches of k_ntt apply:
//   phase [L_E][count][L_R][N_E]   (mod Q_l),   plain [L_R][count][N_E]   (mod t = q_j)
#pragma once
#include "kernels.cuh"

namespace rsg {

constexpr int DEC_MAXW = MAX_LE;   // words of a CRT-composed coefficient (one per limb is enough: every prime < 2^64)

struct DecodeConsts {
  uint64_t inv_punct[MAX_LE];           // (Q / Q_l)^-1 mod Q_l
  uint64_t punct_mod_t[MAX_LR][MAX_LE]; // (Q / Q_l) mod q_j
  uint64_t Q_mod_t[MAX_LR];             // Q mod q_j
  uint64_t garner_inv[MAX_LE][MAX_LE];  // [i][k] = Q_k^-1 mod Q_i, k < i
  uint64_t Qw[DEC_MAXW];                // Q as little-endian words
  uint64_t halfw[DEC_MAXW];             // (Q + 1) >> 1
};

// phase[l][e][j][i] = c0 + c1 * s mod Q_l  (dot_product_ct_sk_array, NTT form).  enc: count encodings [L_R][2][L_E][N_E];
// sk: [L_R][L_E][N_E] (the secret key of ring limb j's encoding context, NTT form).  grid (N_E/256, L_E * L_R, count)
__global__ void __launch_bounds__(256) k_dec_phase(const DevParams *__restrict__ P, const uint64_t *__restrict__ enc,
                                                   const uint64_t *__restrict__ sk, uint32_t count, uint64_t *__restrict__ phase) {
  const uint32_t N_E = P->N_E, L_E = P->L_E, L_R = P->L_R;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y / L_E, l = blockIdx.y - j * L_E, e = blockIdx.z;
  if (i >= N_E) return;
  const ModConst m = P->Q[l];
  const size_t ct = ((size_t)e * L_R + j) * 2 * L_E * N_E;
  const uint64_t c0 = enc[ct + (size_t)l * N_E + i], c1 = enc[ct + ((size_t)L_E + l) * N_E + i];
  const uint64_t s = sk[((size_t)j * L_E + l) * N_E + i];
  phase[(((size_t)l * count + e) * L_R + j) * N_E + i] = add_mod(mul_mod(c1, s, m), c0, m.p);
}

// One thread per coefficient of one (encoding, ring limb): exact base conversion q -> t and the bit length of the centred
// CRT-composed coefficient (max-reduced into bits[e][j]).  grid (N_E/128, L_R, count)
__global__ void __launch_bounds__(128) k_dec_modt(const DevParams *__restrict__ P, const DecodeConsts *__restrict__ K,
                                                  const uint64_t *__restrict__ phase, uint32_t count, uint64_t *__restrict__ plain,
                                                  int *__restrict__ bits) {
  const uint32_t N_E = P->N_E, L_E = P->L_E, L_R = P->L_R;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, e = blockIdx.z;
  if (i >= N_E) return;
  const ModConst mt = P->q[j];
  uint64_t x[MAX_LE];
  for (uint32_t l = 0; l < L_E; l++) x[l] = phase[(((size_t)l * count + e) * L_R + j) * N_E + i];
  // exact_convert_array: temp_l = x_l * inv_punct_l mod Q_l; v = trunc(sum_l temp_l / Q_l + 0.5); out = sum temp_l M_l - v [Q]_t
  double agg = 0.0;
  Acc192 acc;
  acc.clear();
  for (uint32_t l = 0; l < L_E; l++) {
    const ModConst m = P->Q[l];
    const uint64_t temp = mul_mod(x[l], K->inv_punct[l], m);
    agg = __dadd_rn(agg, __ddiv_rn((double)temp, (double)m.p));
    acc.mac(temp, K->punct_mod_t[j][l]);
  }
  agg = __dadd_rn(agg, 0.5);
  const uint64_t v = (uint64_t)agg;   // truncation, as static_cast<uint64_t>
  const uint64_t sum = acc.reduce(mt);
  const uint64_t vq = mul_mod(reduce64(v, mt), K->Q_mod_t[j], mt);
  plain[((size_t)j * count + e) * N_E + i] = sub_mod(sum, vq, mt.p);

  // invariant_noise_budget: compose (Garner mixed radix, then Horner), centre, bit length
  uint64_t a[MAX_LE];
  for (uint32_t l = 0; l < L_E; l++) {
    const ModConst m = P->Q[l];
    uint64_t t = x[l];
    for (uint32_t k = 0; k < l; k++) t = mul_mod(sub_mod(t, reduce64(a[k], m), m.p), K->garner_inv[l][k], m);
    a[l] = t;
  }
  uint64_t X[DEC_MAXW];
  for (int w = 0; w < DEC_MAXW; w++) X[w] = 0;
  X[0] = a[L_E - 1];
  for (int l = (int)L_E - 2; l >= 0; l--) {   // X = X * Q_l + a_l
    const uint64_t ql = P->Q[l].p;
    uint64_t carry = a[l];
    for (uint32_t w = 0; w < L_E; w++) {
      const uint64_t lo = X[w] * ql, hi = __umul64hi(X[w], ql);
      const uint64_t s = lo + carry;
      carry = hi + (s < lo);
      X[w] = s;
    }
  }
  bool ge = true;   // X >= half ?
  for (int w = (int)L_E - 1; w >= 0; w--)
    if (X[w] != K->halfw[w]) { ge = X[w] > K->halfw[w]; break; }
  if (ge) {         // X = Q - X
    uint64_t borrow = 0;
    for (uint32_t w = 0; w < L_E; w++) {
      const uint64_t qw = K->Qw[w], d = qw - X[w] - borrow;
      borrow = (qw < X[w]) || (qw == X[w] && borrow);
      X[w] = d;
    }
  }
  int nb = 0;
  for (int w = (int)L_E - 1; w >= 0; w--)
    if (X[w]) { nb = w * 64 + (64 - __clzll((long long)X[w])); break; }
  if (nb) atomicMax(bits + (size_t)e * L_R + j, nb);
}

// ring[e][j][s] = transformed plain[j][e][index_map[s]], s < N_R  (BatchEncoder::decode reads the top row first; the ring
// element keeps the first N_R slots, seal_ring.tcc:466).  grid (ceil(N_R/256), L_R, count)
__global__ void __launch_bounds__(256) k_dec_gather(const DevParams *__restrict__ P, const uint64_t *__restrict__ plain, uint32_t count,
                                                    uint64_t *__restrict__ ring) {
  const uint32_t N_E = P->N_E, N_R = P->N_R, L_R = P->L_R;
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, e = blockIdx.z;
  if (s >= N_R) return;
  ring[((size_t)e * L_R + j) * N_R + s] = plain[((size_t)j * count + e) * N_E + __ldg(P->index_map + s)];
}

}  // namespace rsg
