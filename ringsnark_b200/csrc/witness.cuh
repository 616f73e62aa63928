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

// The same product for primes below 2^54 (every ring prime of the reference's configs: 36..54 bit) on the FP64 pipe.
// Measured on B200 (scratch/ubench2.cu): DFMA issues at 64 lanes/clk/SM, the same as a 32-bit IMAD, whereas a 32x32->64
// IMAD.WIDE costs ~3 IMAD slots and a 64x64 high product ~9 -- so the cheapest exact wide multiply-accumulate on this
// part is the double-precision FMA.  The constant matrix entry is cut into two 27-bit halves and the variable operand
// into three 18-bit thirds: every partial product is < 2^45 and a double accumulates 128 of them exactly (< 2^52).
// Six DFMA per 54x54-bit MAC, no carries, no reductions in the inner loop; the six sums (weights 2^0, 2^18, 2^27, 2^36,
// 2^45, 2^63) are folded into a running canonical residue once per tile of 128 columns.  Integer in, integer out: the
// doubles hold exact integers throughout, so the result is the same canonical residue as the integer kernel's.
// Each thread owns one slot x MM_ROWS rows; the matrix tile is staged in shared memory as doubles, [column][row].
constexpr int MMF_KTILE = 128;
__device__ __forceinline__ double u32_to_double_exact(uint32_t v) {   // v as a double via the 2^52 mantissa trick
  return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
}
__device__ __forceinline__ uint64_t double_to_u64_exact(double d) {   // exact integer 0 <= d < 2^52
  return (uint64_t)__double_as_longlong(d + 4503599627370496.0) & 0x000FFFFFFFFFFFFFull;
}
// Register tile: MM_ROWS rows x MMF_SLOTS slots per thread (96 double accumulators).  Shared-memory return bandwidth is
// the second limit after the FP64 pipe -- every lane of a warp needs the same matrix entry, and a broadcast LDS.128 still
// returns 512 B -- so each fetched entry must feed two slots (LDS at ~2/3 of the DFMA time instead of 4/3).
template <int ROWS, int MMF_SLOTS, int MMF_PF, int MINB>
__global__ void __launch_bounds__(MM_THREADS, MINB) k_modmat_f64(const ModConst *__restrict__ mods, const uint64_t *__restrict__ M,
                                                              uint32_t Mrows, uint32_t K, const uint64_t *__restrict__ Y,
                                                              uint64_t *__restrict__ C, uint32_t N_R, uint32_t L_R,
                                                              uint32_t upper_triangular) {
  __shared__ __align__(16) double2 tile[MMF_KTILE][ROWS];   // (low 27 bits, high 27 bits) of M[r0 + r][c0 + c]
  const uint32_t r0 = blockIdx.x * ROWS;
  const uint32_t slot0 = blockIdx.y * (MM_THREADS * MMF_SLOTS) + threadIdx.x;
  const uint32_t v = blockIdx.z / L_R, limb = blockIdx.z - v * L_R;
  const size_t W = (size_t)N_R * L_R;
  const uint64_t *Mp = M + (size_t)limb * Mrows * K;
  const uint64_t *Yp[MMF_SLOTS];
  bool live[MMF_SLOTS];
#pragma unroll
  for (int s = 0; s < MMF_SLOTS; s++) {
    live[s] = slot0 + s * MM_THREADS < N_R;
    Yp[s] = Y + (size_t)v * K * W + (size_t)limb * N_R + (live[s] ? slot0 + s * MM_THREADS : 0);   // dead lanes re-read slot 0
  }
  const ModConst m = mods[limb];
  uint64_t *Cp = C + (size_t)v * Mrows * W + (size_t)limb * N_R + slot0;
  const uint32_t c_begin = upper_triangular ? (r0 / MMF_KTILE) * MMF_KTILE : 0;
  for (uint32_t c0 = c_begin; c0 < K; c0 += MMF_KTILE) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < ROWS * MMF_KTILE; i += MM_THREADS) {
      const uint32_t r = i / MMF_KTILE, c = i % MMF_KTILE;   // coalesced along a matrix row
      const uint64_t mv = (r0 + r < Mrows && c0 + c < K) ? Mp[(size_t)(r0 + r) * K + c0 + c] : 0;
      tile[c][r] = make_double2(u32_to_double_exact((uint32_t)mv & 0x7FFFFFFu), u32_to_double_exact((uint32_t)(mv >> 27)));
    }
    __syncthreads();
    const uint32_t cn = min((uint32_t)MMF_KTILE, K - c0);
    double acc[MMF_SLOTS][ROWS][6];
#pragma unroll
    for (int s = 0; s < MMF_SLOTS; s++)
#pragma unroll
      for (int r = 0; r < ROWS; r++)
#pragma unroll
        for (int k = 0; k < 6; k++) acc[s][r][k] = 0.0;
    // columns in groups of MMF_PF, the next group's Y words already in flight; columns past K meet zero tile entries
    uint64_t ynext[MMF_SLOTS][MMF_PF];
#pragma unroll
    for (int s = 0; s < MMF_SLOTS; s++)
#pragma unroll
      for (int k = 0; k < MMF_PF; k++) ynext[s][k] = Yp[s][(size_t)min(c0 + k, K - 1) * W];
    for (uint32_t c = 0; c < cn; c += MMF_PF) {
      uint64_t ycur[MMF_SLOTS][MMF_PF];
#pragma unroll
      for (int s = 0; s < MMF_SLOTS; s++)
#pragma unroll
        for (int k = 0; k < MMF_PF; k++) ycur[s][k] = ynext[s][k];
      if (c + MMF_PF < cn) {
#pragma unroll
        for (int s = 0; s < MMF_SLOTS; s++)
#pragma unroll
          for (int k = 0; k < MMF_PF; k++) ynext[s][k] = Yp[s][(size_t)min(c0 + c + MMF_PF + k, K - 1) * W];
      }
#pragma unroll
      for (int k = 0; k < MMF_PF; k++) {
        double y0[MMF_SLOTS], y1[MMF_SLOTS], y2[MMF_SLOTS];
#pragma unroll
        for (int s = 0; s < MMF_SLOTS; s++) {
          y0[s] = u32_to_double_exact((uint32_t)ycur[s][k] & 0x3FFFFu);
          y1[s] = u32_to_double_exact((uint32_t)(ycur[s][k] >> 18) & 0x3FFFFu);
          y2[s] = u32_to_double_exact((uint32_t)(ycur[s][k] >> 36));
        }
#pragma unroll
        for (int r = 0; r < ROWS; r++) {
          const double2 mm = tile[c + k][r];
#pragma unroll
          for (int s = 0; s < MMF_SLOTS; s++) {
            acc[s][r][0] = fma(mm.x, y0[s], acc[s][r][0]);
            acc[s][r][1] = fma(mm.x, y1[s], acc[s][r][1]);
            acc[s][r][2] = fma(mm.x, y2[s], acc[s][r][2]);
            acc[s][r][3] = fma(mm.y, y0[s], acc[s][r][3]);
            acc[s][r][4] = fma(mm.y, y1[s], acc[s][r][4]);
            acc[s][r][5] = fma(mm.y, y2[s], acc[s][r][5]);
          }
        }
      }
    }
    // fold: sum_k acc_k * 2^w_k, w = {0, 18, 36, 27, 45, 63} (< 2^116): one Barrett reduction per output per tile, added to
    // the running canonical residue kept in C itself (this thread is the only writer of its outputs)
#pragma unroll
    for (int s = 0; s < MMF_SLOTS; s++) {
      if (!live[s]) continue;
#pragma unroll
      for (int r = 0; r < ROWS; r++) {
        if (r0 + r >= Mrows) continue;
        uint64_t lo = double_to_u64_exact(acc[s][r][0]), hi = 0;
        const int wsh[5] = {18, 36, 27, 45, 63};
#pragma unroll
        for (int k = 0; k < 5; k++) {
          const uint64_t x = double_to_u64_exact(acc[s][r][k + 1]);
          const uint64_t t = x << wsh[k];
          lo += t;
          hi += (lo < t) + (x >> (64 - wsh[k]));
        }
        uint64_t *o = Cp + (size_t)(r0 + r) * W + s * MM_THREADS;
        const uint64_t part = reduce128(lo, hi, m);
        *o = c0 == c_begin ? part : add_mod(*o, part, m.p);
      }
    }
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

// out[limb][r] = sum_c M[limb][r][c] * x[limb][c] mod q_limb: the interpolant of a per-constraint CONSTANT (the same in
// every slot), i.e. of the constant wire's contribution.  grid (ceil(n/128), L_R), tiny.
__global__ void __launch_bounds__(128) k_matvec(const ModConst *__restrict__ mods, const uint64_t *__restrict__ M, uint32_t n,
                                                const uint64_t *__restrict__ x, uint64_t *__restrict__ out) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x, limb = blockIdx.y;
  if (r >= n) return;
  const uint64_t *row = M + ((size_t)limb * n + r) * n, *xv = x + (size_t)limb * n;
  Acc192 acc;
  acc.clear();
  for (uint32_t c = 0; c < n; c++) acc.mac(row[c], xv[c]);
  out[(size_t)limb * n + r] = acc.reduce(mods[limb]);
}
// Interpolation is linear and full = mid + io - constant (the constant wire is counted under both partial assignments,
// r1cs_to_qrp.tcc:167-201 with variable.tcc:246-254), so the interpolants of the FULL assignment need no product of their own:
//   aA[k] = A_io[k] + A_mid[k] - cc[0][limb][k],  aB likewise with cc[1].   coeffs order: A_io,B_io,C_io,A_mid,B_mid,C_mid.
// grid (n, ceil(W/256), 2)
__global__ void __launch_bounds__(256) k_full_from_parts(const ModConst *__restrict__ mods, const uint64_t *__restrict__ coeffs,
                                                         const uint64_t *__restrict__ cc, uint64_t *__restrict__ aAB, uint32_t n,
                                                         uint32_t N_R, uint32_t L_R) {
  const uint32_t k = blockIdx.x, w = blockIdx.y * blockDim.x + threadIdx.x, m = blockIdx.z;
  const uint32_t W = N_R * L_R;
  if (w >= W) return;
  const uint32_t limb = w / N_R;
  const uint64_t p = mods[limb].p;
  const uint64_t io = coeffs[((size_t)m * n + k) * W + w], mid = coeffs[((size_t)(3 + m) * n + k) * W + w];
  const uint64_t c = cc[((size_t)m * L_R + limb) * n + k];
  aAB[((size_t)m * n + k) * W + w] = sub_mod(add_mod(io, mid, p), c, p);
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
