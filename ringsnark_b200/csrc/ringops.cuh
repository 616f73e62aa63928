// Element-wise RingElem operators on device-resident ring vectors -- the generic fall-backs of the backend concept
// (SURVEY.md section 8(b)): ringsnark/seal/seal_ring.tcc:105-263 over depends/SEAL-Polytools/src/poly_arith.cpp:164-350.
// Every slot of every limb is independent: one thread per word, 128-bit loads where the shape allows.
//   binop  : add_inplace / subtract_inplace / multiply_inplace (poly_arith.cpp:191-266 -> add/sub/dyadic_product_coeffmod)
//   scalar : add_scalar / subtract_scalar / multiply_scalar (poly_arith.cpp:164-189,241-252); the scalar is the same in
//            every slot and is NOT reduced by SEAL's add/sub (one conditional correction, uintarithsmallmod.h:114-150)
//   negate : negate_inplace (poly_arith.cpp:342-350)
//   invert : invert_inplace (poly_arith.cpp:304-340): per-slot inverse, fails if any slot is zero
#pragma once
#include "modarith.cuh"

namespace rsg {

enum RingOp { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2 };

// grid (ceil(N_R/256), L_R, count)
__global__ void __launch_bounds__(256) k_ring_binop(const ModConst *__restrict__ mods, int op, const uint64_t *__restrict__ a,
                                                    const uint64_t *__restrict__ b, uint64_t *__restrict__ out, uint32_t N_R,
                                                    uint32_t L_R) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x, limb = blockIdx.y;
  if (slot >= N_R) return;
  const size_t w = ((size_t)blockIdx.z * L_R + limb) * N_R + slot;
  const ModConst m = mods[limb];
  const uint64_t x = a[w], y = b[w];
  out[w] = op == OP_ADD ? add_mod(x, y, m.p) : (op == OP_SUB ? sub_mod(x, y, m.p) : mul_mod(x, y, m));
}

// b_scalar mode: the second operand is the uint64 scalar (RingElem's scalar alternative, seal_ring.hpp:22-26)
__global__ void __launch_bounds__(256) k_ring_scalar(const ModConst *__restrict__ mods, int op, const uint64_t *__restrict__ a,
                                                     uint64_t scalar, uint64_t *__restrict__ out, uint32_t N_R, uint32_t L_R) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x, limb = blockIdx.y;
  if (slot >= N_R) return;
  const size_t w = ((size_t)blockIdx.z * L_R + limb) * N_R + slot;
  const ModConst m = mods[limb];
  const uint64_t x = a[w];
  uint64_t r;
  if (op == OP_ADD) {            // add_uint_mod: one conditional subtraction, operand taken as is
    r = x + scalar;
    r = r >= m.p ? r - m.p : r;
  } else if (op == OP_SUB) {     // sub_uint_mod: one conditional addition
    r = x - scalar;
    r = x < scalar ? r + m.p : r;
  } else {                       // multiply_poly_scalar_coeffmod reduces the scalar first
    r = mul_mod(x, reduce64(scalar, m), m);
  }
  out[w] = r;
}

__global__ void __launch_bounds__(256) k_ring_negate(const ModConst *__restrict__ mods, const uint64_t *__restrict__ a,
                                                     uint64_t *__restrict__ out, uint32_t N_R, uint32_t L_R) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x, limb = blockIdx.y;
  if (slot >= N_R) return;
  const size_t w = ((size_t)blockIdx.z * L_R + limb) * N_R + slot;
  out[w] = neg_mod(a[w], mods[limb].p);
}

// a^(p-2) mod p per slot; bad[e] |= 1 if element e has a zero slot (not invertible in the ring)
__global__ void __launch_bounds__(256) k_ring_invert(const ModConst *__restrict__ mods, const uint64_t *__restrict__ a,
                                                     uint64_t *__restrict__ out, uint32_t N_R, uint32_t L_R,
                                                     uint32_t *__restrict__ bad) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x, limb = blockIdx.y;
  if (slot >= N_R) return;
  const size_t w = ((size_t)blockIdx.z * L_R + limb) * N_R + slot;
  const ModConst m = mods[limb];
  const uint64_t x = a[w];
  if (x == 0) {
    atomicOr(bad + blockIdx.z, 1u);
    out[w] = 0;
    return;
  }
  uint64_t r = 1, base = x, e = m.p - 2;
  while (e) {
    if (e & 1) r = mul_mod(r, base, m);
    base = mul_mod(base, base, m);
    e >>= 1;
  }
  out[w] = r;
}

}  // namespace rsg
