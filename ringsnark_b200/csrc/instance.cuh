// Instance map with evaluation: the O(m^2) part of ringGroth16 / Rinocchio setup and of EVERY
// verification, slot-parallel on the device.  Restates
//   evaluate_all_lagrange_polynomials   ringsnark/util/evaluation_domain.tcc:20-41
//   compute_vanishing_polynomial        ringsnark/util/evaluation_domain.tcc:43-51
//   r1cs_to_qrp_instance_map_with_evaluation  ringsnark/reductions/r1cs_to_qrp/r1cs_to_qrp.tcc:75-116
// The reference forms, for every j, prod_{i != j} (t - x_i) and prod_{i != j} (x_j - x_i) from scratch (2 m^2 ring
// multiplications and m ring divisions).  On the domain {0..m-1} the denominators are per-prime constants
// (j! (m-1-j)! (-1)^(m-1-j)) and the numerators are prefix x suffix products of (t - i), so a slot needs 3m modular
// multiplications -- same canonical residues (exact arithmetic in Z_q, the division is by a unit).
#pragma once
#include "modarith.cuh"

namespace rsg {

// One thread per (slot, limb).  u[j] = prefix_j * suffix_j * w_j, Ht[i] = t^i (i <= m), Zt = prod_i (t - i).
//   t: one ring element [L_R][N_R];  w: [L_R][m] constants as Shoup pairs;  u: m elements, Ht: m+1 elements, Zt: 1.
// The prefix products are parked in u itself on the way up and completed on the way down.
__global__ void __launch_bounds__(128) k_lagrange_at(const ModConst *__restrict__ mods, const uint64_t *__restrict__ t,
                                                     const Twiddle *__restrict__ w, uint32_t m, uint64_t *__restrict__ u,
                                                     uint64_t *__restrict__ Ht, uint64_t *__restrict__ Zt, uint32_t N_R,
                                                     uint32_t L_R) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x, limb = blockIdx.y;
  if (slot >= N_R) return;
  const size_t W = (size_t)N_R * L_R, o = (size_t)limb * N_R + slot;
  const ModConst mc = mods[limb];
  const uint64_t tv = t[o];
  uint64_t pre = 1, pw = 1;
  for (uint32_t j = 0; j < m; j++) {
    u[(size_t)j * W + o] = pre;                                   // prod_{i<j} (t - i)
    Ht[(size_t)j * W + o] = pw;                                   // t^j
    pre = mul_mod(pre, sub_mod(tv, reduce64(j, mc), mc.p), mc);
    pw = mul_mod(pw, tv, mc);
  }
  Ht[(size_t)m * W + o] = pw;
  Zt[o] = pre;
  uint64_t suf = 1;
  const Twiddle *wl = w + (size_t)limb * m;
  for (uint32_t j = m; j-- > 0;) {
    const uint64_t v = mul_mod(u[(size_t)j * W + o], suf, mc);
    const ulonglong2 ws = __ldg(reinterpret_cast<const ulonglong2 *>(wl) + j);
    Twiddle tw;
    tw.w = ws.x;
    tw.wq = ws.y;
    u[(size_t)j * W + o] = mul_shoup(v, tw, mc.p);
    suf = mul_mod(suf, sub_mod(tv, reduce64(j, mc), mc.p), mc);
  }
}

// At[k] = sum over the constraints i whose linear combination A_i mentions variable k of u[i] * coeff  (likewise Bt, Ct):
// the transposed sparse product, one block per (variable, matrix), threads over slots.
//   col_ptr: [3 * (vars + 1) + 1], rows / coeff: the constraint index and coefficient of each entry (CSC of A | B | C).
//   out element (mat * (vars + 1) + k).   grid (vars + 1, 3, L_R * ceil(N_R / 128))
__global__ void __launch_bounds__(128) k_instance_accum(const ModConst *__restrict__ mods, const uint32_t *__restrict__ col_ptr,
                                                        const uint32_t *__restrict__ rows, const uint64_t *__restrict__ coeff,
                                                        uint32_t nvars1, const uint64_t *__restrict__ u, uint64_t *__restrict__ out,
                                                        uint32_t N_R, uint32_t L_R) {
  const uint32_t k = blockIdx.x, mat = blockIdx.y;
  const uint32_t sblocks = (N_R + 127) / 128;
  const uint32_t limb = blockIdx.z / sblocks, slot = (blockIdx.z - limb * sblocks) * 128 + threadIdx.x;
  if (slot >= N_R) return;
  const size_t W = (size_t)N_R * L_R, o = (size_t)limb * N_R + slot;
  const ModConst mc = mods[limb];
  const uint32_t c = mat * nvars1 + k;
  Acc192 acc;
  acc.clear();
  for (uint32_t e = col_ptr[c]; e < col_ptr[c + 1]; e++) acc.mac(reduce64(coeff[e], mc), u[(size_t)rows[e] * W + o]);
  out[(size_t)c * W + o] = acc.reduce(mc);
}

}  // namespace rsg
