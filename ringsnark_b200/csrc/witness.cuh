// Witness-map kernels (hot path (a)): slot-parallel restatement of
//   interpolate            ringsnark/util/polynomials.tcc:9-43        (coeffs = V^-1 y on the domain {0..n-1})
//   multiply / add / divide ringsnark/util/polynomials.tcc:61-81       (H = (A*B - C) / Z, Z monic)
//   r1cs_to_qrp_witness_map ringsnark/reductions/r1cs_to_qrp/r1cs_to_qrp.tcc:148-259
// The reference runs an O(n^2) scalar routine per ring element with every "scalar" a length-N_R*L_R vector.
// Every slot of one ring prime sees the SAME domain, so V^-1 (Lagrange basis coefficients) and the power-series
// inverse of rev(Z) are per-prime constants built once on the host; the device work is then
//   (1) k_modmat   : C = M * Y over Z_p, M constant (n x n), Y = n ring elements x slots    [interpolation, division]
//   (2) k_conv_top : top half of the per-slot product A*B (only coefficients n..2n-2 reach the quotient)
// Exactly the residues the reference computes: all results are canonical and the quotient of the long division
// by a monic Z depends only on the dividend's coefficients of degree >= n (Knuth 4.6.1 D).
#pragma once
#include "modarith.cuh"

namespace rsg {

constexpr int MM_ROWS = 8;      // output rows per thread
constexpr int MM_THREADS = 128; // slots per block
constexpr int MM_KTILE = 64;    // matrix columns staged per shared-memory tile

// C[v][r][slot] = sum_c M[limb][r][c] * Y[v][c][slot]  mod q_limb
//   Y element (v*K + c), C element (v*Mrows + r); element layout [L_R][N_R].
//   c_lo[r] (nullable): first non-zero column of row r is >= c_lo_of_tile -- used for triangular matrices.
// grid (ceil(Mrows/MM_ROWS), N_R/MM_THREADS, batch*L_R): row tiles vary fastest so co-resident blocks share Y in L2.
__global__ void __launch_bounds__(MM_THREADS) k_modmat(const ModConst *__restrict__ mods, const uint64_t *__restrict__ M,
                                                       uint32_t Mrows, uint32_t K, const uint64_t *__restrict__ Y,
                                                       uint64_t *__restrict__ C, uint32_t N_R, uint32_t L_R,
                                                       uint32_t upper_triangular, uint32_t K_valid) {
  __shared__ uint64_t tile[MM_ROWS][MM_KTILE];
  const uint32_t r0 = blockIdx.x * MM_ROWS;
  const uint32_t slot = blockIdx.y * MM_THREADS + threadIdx.x;
  const uint32_t v = blockIdx.z / L_R, limb = blockIdx.z - v * L_R;
  const size_t W = (size_t)N_R * L_R;
  const uint64_t *Mp = M + (size_t)limb * Mrows * K;
  const uint64_t *Yp = Y + (size_t)v * K * W + (size_t)limb * N_R + slot;
  Acc192 acc[MM_ROWS];
#pragma unroll
  for (int r = 0; r < MM_ROWS; r++) acc[r].clear();
  // triangular: row r only has columns >= r, so this row tile starts at column r0 (rounded down to a tile)
  const uint32_t c_begin = upper_triangular ? (r0 / MM_KTILE) * MM_KTILE : 0;
  const uint32_t c_end = min(K, K_valid);
  for (uint32_t c0 = c_begin; c0 < c_end; c0 += MM_KTILE) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < MM_ROWS * MM_KTILE; i += MM_THREADS) {
      const uint32_t r = i / MM_KTILE, c = i % MM_KTILE;
      tile[r][c] = (r0 + r < Mrows && c0 + c < c_end) ? Mp[(size_t)(r0 + r) * K + c0 + c] : 0;
    }
    __syncthreads();
    const uint32_t cn = min((uint32_t)MM_KTILE, c_end - c0);
    if (slot < N_R) {
#pragma unroll 4
      for (uint32_t c = 0; c < cn; c++) {
        const uint64_t y = Yp[(size_t)(c0 + c) * W];
#pragma unroll
        for (int r = 0; r < MM_ROWS; r++) acc[r].mac(tile[r][c], y);
      }
    }
  }
  if (slot < N_R) {
    const ModConst m = mods[limb];
    uint64_t *Cp = C + (size_t)v * Mrows * W + (size_t)limb * N_R + slot;
#pragma unroll
    for (int r = 0; r < MM_ROWS; r++)
      if (r0 + r < Mrows) Cp[(size_t)(r0 + r) * W] = acc[r].reduce(m);
  }
}

// Ptop[i][slot] = sum_j A[j][slot] * B[n + i - j][slot],  i in [0, n-1), j in [i+1, n)  -- coefficient n+i of A*B.
// Elements j >= lenA of A and >= lenB of B count as zero (Boost normalize() through RingElem::operator==).
// grid (ceil((n-1)/MM_ROWS), N_R/MM_THREADS, L_R)
__global__ void __launch_bounds__(MM_THREADS) k_conv_top(const ModConst *__restrict__ mods, const uint64_t *__restrict__ A,
                                                         const uint64_t *__restrict__ B, uint32_t n, uint32_t lenA,
                                                         uint32_t lenB, uint64_t *__restrict__ Ptop, uint32_t N_R,
                                                         uint32_t L_R) {
  const uint32_t i0 = blockIdx.x * MM_ROWS;
  const uint32_t slot = blockIdx.y * MM_THREADS + threadIdx.x;
  const uint32_t limb = blockIdx.z;
  if (slot >= N_R) return;
  const size_t W = (size_t)N_R * L_R;
  const uint64_t *Ap = A + (size_t)limb * N_R + slot;
  const uint64_t *Bp = B + (size_t)limb * N_R + slot;
  Acc192 acc[MM_ROWS];
  uint64_t win[MM_ROWS];   // win[r] = B[n + i0 + r - j]
#pragma unroll
  for (int r = 0; r < MM_ROWS; r++) { acc[r].clear(); win[r] = 0; }
  // j runs from i0+1; at that point only r = 0 has an in-range B index (n-1)
  for (uint32_t j = i0 + 1; j < n; j++) {
    // shift the window: index for r at this j equals index for r-1 at j-1
#pragma unroll
    for (int r = MM_ROWS - 1; r > 0; r--) win[r] = win[r - 1];
    const uint32_t bi = n + i0 - j;                     // index for r = 0
    win[0] = bi < lenB ? Bp[(size_t)bi * W] : 0;
    const uint64_t a = j < lenA ? Ap[(size_t)j * W] : 0;
#pragma unroll
    for (int r = 0; r < MM_ROWS; r++) acc[r].mac(a, win[r]);
  }
  const ModConst m = mods[limb];
  uint64_t *Pp = Ptop + (size_t)limb * N_R + slot;
#pragma unroll
  for (int r = 0; r < MM_ROWS; r++)
    if (i0 + r < n - 1) Pp[(size_t)(i0 + r) * W] = acc[r].reduce(m);
}

// Zero-knowledge patch of H (r1cs_to_qrp.tcc:225-235, evaluation_domain.tcc:62-76):
//   H[i] += d2*A[i] + d1*B[i]  (i < n);   H[0] -= d3;   H[i] += (d1*d2) * Z[i]  (i <= n)
// d = 3 ring elements d1,d2,d3; Z = per-prime constants [L_R][n+1].  grid (n+1, ceil(W/256)).
__global__ void __launch_bounds__(256) k_h_patch(const ModConst *__restrict__ mods, uint64_t *__restrict__ H,
                                                 const uint64_t *__restrict__ A, const uint64_t *__restrict__ B,
                                                 const uint64_t *__restrict__ d, const uint64_t *__restrict__ Z, uint32_t n,
                                                 uint32_t N_R, uint32_t L_R) {
  const uint32_t i = blockIdx.x;
  const uint32_t w = blockIdx.y * blockDim.x + threadIdx.x;
  const uint32_t W = N_R * L_R;
  if (w >= W) return;
  const uint32_t limb = w / N_R;
  const ModConst m = mods[limb];
  const uint64_t d1 = d[w], d2 = d[(size_t)W + w], d3 = d[2 * (size_t)W + w];
  Acc192 acc;
  acc.clear();
  acc.add(H[(size_t)i * W + w]);
  if (i < n) {
    acc.mac(d2, A[(size_t)i * W + w]);
    acc.mac(d1, B[(size_t)i * W + w]);
  }
  acc.mac(mul_mod(d1, d2, m), Z[(size_t)limb * (n + 1) + i]);
  uint64_t r = acc.reduce(m);
  if (i == 0) r = sub_mod(r, d3, m.p);
  H[(size_t)i * W + w] = r;
}

// linear_combination::evaluate (ringsnark/relations/variable.tcc:246-254) for all 3n linear combinations and the
// three assignments the witness map uses (r1cs_to_qrp.tcc:167-219): "mid" (primary inputs zeroed), "io" (auxiliary
// inputs zeroed), "full".  The constant wire (index 0) contributes its coefficient under ALL three assignments, as
// the reference does.  Coefficients are uint64 scalars reduced mod q_j
// (multiply_poly_scalar_coeffmod semantics).  CSR rows r = m*n + i for matrix m in {A,B,C}.
//   evals element ((variant*3 + m)*n + i), variant 0 = mid, 1 = io, 2 = full.
// grid (n, 3, L_R * ceil(N_R/MM_THREADS))
__global__ void __launch_bounds__(MM_THREADS) k_r1cs_eval(const ModConst *__restrict__ mods, const uint32_t *__restrict__ row_ptr,
                                                          const uint32_t *__restrict__ col, const uint64_t *__restrict__ coeff,
                                                          uint32_t n, uint32_t n_io, const uint64_t *__restrict__ assign,
                                                          uint64_t *__restrict__ evals, uint32_t N_R, uint32_t L_R) {
  const uint32_t i = blockIdx.x, m = blockIdx.y;
  const uint32_t sblocks = (N_R + MM_THREADS - 1) / MM_THREADS;
  const uint32_t limb = blockIdx.z / sblocks, slot = (blockIdx.z - limb * sblocks) * MM_THREADS + threadIdx.x;
  if (slot >= N_R) return;
  const size_t W = (size_t)N_R * L_R;
  const ModConst mc = mods[limb];
  const uint32_t r = m * n + i;
  Acc192 a_mid, a_io, a_full;
  a_mid.clear(); a_io.clear(); a_full.clear();
  for (uint32_t t = row_ptr[r]; t < row_ptr[r + 1]; t++) {
    const uint32_t idx = col[t];
    const uint64_t c = reduce64(coeff[t], mc);
    if (idx == 0) {
      a_mid.add(c); a_io.add(c); a_full.add(c);
    } else {
      const uint64_t x = assign[(size_t)(idx - 1) * W + (size_t)limb * N_R + slot];
      a_full.mac(c, x);
      if (idx - 1 < n_io) a_io.mac(c, x);
      else a_mid.mac(c, x);
    }
  }
  const size_t o = (size_t)limb * N_R + slot;
  evals[((size_t)(0 * 3 + m) * n + i) * W + o] = a_mid.reduce(mc);
  evals[((size_t)(1 * 3 + m) * n + i) * W + o] = a_io.reduce(mc);
  evals[((size_t)(2 * 3 + m) * n + i) * W + o] = a_full.reduce(mc);
}

}  // namespace rsg
