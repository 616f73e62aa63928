// Quasi-linear witness map: the same canonical residues as witness.cuh's dense products (and as the reference's
//   interpolate            ringsnark/util/polynomials.tcc:9-43
//   multiply / divide      ringsnark/util/polynomials.tcc:61-81
//   r1cs_to_qrp_witness_map ringsnark/reductions/r1cs_to_qrp/r1cs_to_qrp.tcc:148-259 ),
// in O(n log^2 n) instead of O(n^2) work per slot.
//
// The domain is the arithmetic progression {0..n-1} (util/evaluation_domain.tcc:53-84) and every ring prime q_j is
// = 1 mod 2*N_E (it is a batching plaintext modulus), so Z_{q_j} has negacyclic NTTs of every size S <= N_E -- with the
// twiddle tables the batch encoder already holds: SEAL's root_powers_ for size N_E/2^k are the first N_E/2^k entries of
// the size-N_E table (psi^bitrev_logN(i) = (psi^2)^bitrev_{logN-1}(i) for i < N_E/2).
//
//   interpolation, per slot:
//     (1) Newton coefficients on the nodes 0,1,2,..:  c_k = sum_{i<=k} (y_i / i!) * ((-1)^(k-i) / (k-i)!) -- one product
//         with a constant series (forward differences written as a convolution);
//     (2) Newton -> monomial basis by divide and conquer on the fixed subproduct tree of the nodes:
//         F_block = F_lo + P_block * F_hi,  P_block = prod_{j<m} (x - (first node of the block + j))  (a per-prime
//         constant, held NTT-transformed); blocks of 16 coefficients are converted by Horner's rule in registers.
//   quotient H = (A*B - C) / Z, per slot:  only the coefficients n..2n-2 of A*B reach the quotient of the division by the
//         monic Z (deg C < n), and rev(H) = rev(top) * rev(Z)^-1 mod x^(n-1): two more products.
//
// Products are taken modulo x^S + 1 with S >= n a power of two; the few coefficients that wrap around (2n-1-S of them,
// at most WF_WC_MAX or S is doubled) are short sums and are computed directly, so S = 2048 serves n = 1031.
// One CTA owns `SL` slots of one ring limb; all polynomials of those slots stay in shared memory from the first load
// to the last store.  Integer pipe (Shoup butterflies): ring primes go up to 61 bit.
#pragma once
#include "kernels.cuh"
#include "ntt.cuh"

namespace rsg {

constexpr int WF_B = 16;         // coefficients converted by Horner's rule at the leaves
constexpr int WF_WC_MAX = 32;    // wrapped coefficients fixed up directly
constexpr int WF_HMAX = 16;      // a trailing block with at most this many high coefficients is multiplied directly
constexpr int WF_MAX_LEVELS = 12;

struct FastTables {              // device pointers; [L_R] major
  const Twiddle *invfact;        // [L_R][n]      1 / i!
  const Twiddle *pts;            // [L_R][npad]   the node j as a Shoup operand, npad = n rounded up to WF_B
  const Twiddle *Ghat;           // [L_R][S]      NTT_S((-1)^j / j!) / S, forward-output order
  const uint64_t *g_nat;         // [L_R][n]      (-1)^j / j!
  const Twiddle *Phat;           // [L_R][levels][S]   level l, block b at [b*2m, (b+1)*2m): NTT_2m(P_block) / 2m
  const uint64_t *Pnat;          // [L_R][levels][S/2+1]  coefficients of the LAST active block's P at each level
  const Twiddle *Vhat;           // [L_R][S]      NTT_S(rev(Z)^-1 mod x^(n-1)) / S
  const uint64_t *v_nat;         // [L_R][n]      rev(Z)^-1 mod x^(n-1)
  Twiddle invS[MAX_LR];          // 1 / S
  uint32_t n, S, logS, wc, levels;
};

// One radix-2^RL pass (levels [s, s+RL) of the forward transform, or the same levels of the inverse) over a batch of
// negacyclic transforms of size 2^lg: per slot, transform b < nb occupies words [b << lg, (b+1) << lg) of the slot's
// padded buffer.  Same butterflies, table order and laziness as ntt.cuh; sizes are run-time values here because one
// kernel walks all levels of the divide and conquer.
// LAZY (every ring prime below 2^57): butterflies without range corrections where the bound allows it -- forward values grow
// by 4p per level from a canonical input (at most 13 levels here: < 53p < 2^64) and any 64-bit value is a valid operand of
// the next Shoup multiplication; the inverse keeps its values in [0, 4p) with the three-product Shoup quotient.
// FUSE: element-wise work folded into a pass so that it costs no shared-memory round trip and no barrier of its own.
enum : int {
  WF_PLAIN = 0,
  WF_LOAD_DUP = 1,       // first forward pass of a tree level: read (F_hi | F_hi) of block b from the coefficient buffer
  WF_STORE_MUL = 2,      // last forward pass: multiply by the transformed constant before the store
  WF_STORE_COMBINE = 4,  // last inverse pass of a tree level: F_lo + product straight into the coefficient buffer
  WF_STORE_NEWTON = 8    // last inverse pass of the Newton-coefficient product: canonicalise, add the wrapped terms, cut at n
};
struct WfFuse {
  uint64_t *coef = nullptr;        // LOAD_DUP (read) / STORE_COMBINE (read-modify-write): the slots' coefficient buffers
  uint32_t m = 0;                  // half block size of the level
  const Twiddle *mul = nullptr;    // STORE_MUL: constant for position idx of the slot buffer
  const uint64_t *wr = nullptr;    // STORE_NEWTON: wrapped terms [slot][WF_WC_MAX]
  uint32_t n = 0, wc = 0;
};
template <int RL, bool INVERSE, bool LAZY, int FUSE = WF_PLAIN>
__device__ __forceinline__ void wf_pass(uint64_t *buf, uint32_t slot_stride, uint32_t nslots, uint32_t nb, uint32_t lg,
                                        uint32_t s, const Twiddle *__restrict__ tab, uint64_t p, const WfFuse &f = WfFuse()) {
  constexpr int R = 1 << RL;
  const uint32_t lgi = lg - RL, lgg = lg - s - RL, g = 1u << lgg;
  const uint32_t per_slot = nb << lgi, total = per_slot * nslots;
  const uint64_t two_p = p << 1, four_p = p << 2;
  for (uint32_t t = threadIdx.x; t < total; t += blockDim.x) {
    const uint32_t slot = t / per_slot, r = t - slot * per_slot;
    const uint32_t b = r >> lgi, li = r & ((1u << lgi) - 1);
    const uint32_t o = li & (g - 1), blk = li >> lgg;
    const uint32_t base = (b << lg) + (blk << (lg - s)) + o;
    uint64_t *sp = buf + slot * slot_stride;
    uint64_t v[R];
    if (FUSE & WF_LOAD_DUP) {
      const uint64_t *cp = f.coef + slot * slot_stride;
      const uint32_t two_m = 2 * f.m;
#pragma unroll
      for (int k = 0; k < R; k++) {
        const uint32_t idx = base + k * g, e = idx & (two_m - 1);
        v[k] = cp[pad_idx(idx - e + f.m + (e & (f.m - 1)))];
      }
    } else {
#pragma unroll
      for (int k = 0; k < R; k++) v[k] = sp[pad_idx(base + k * g)];
    }
    if (!INVERSE) {
#pragma unroll
      for (int u = 0; u < RL; u++) {
        const int half = R >> (u + 1);
        const uint32_t tbase = (1u << (s + u)) + (blk << u);
#pragma unroll
        for (int grp = 0; grp < (1 << u); grp++) {
          const Twiddle tw = load_tw(tab, tbase + grp);
#pragma unroll
          for (int k = 0; k < half; k++) {
            if (LAZY) bfly_fwd_lazy(v[grp * 2 * half + k], v[grp * 2 * half + k + half], tw, p, four_p);
            else bfly_fwd(v[grp * 2 * half + k], v[grp * 2 * half + k + half], tw, p, two_p);
          }
        }
      }
    } else {
#pragma unroll
      for (int u = RL - 1; u >= 0; u--) {
        const int half = R >> (u + 1);
        const uint32_t tbase = (1u << (s + u)) + (blk << u);
#pragma unroll
        for (int grp = 0; grp < (1 << u); grp++) {
          const Twiddle tw = load_tw(tab, tbase + grp);
#pragma unroll
          for (int k = 0; k < half; k++) {
            if (LAZY) bfly_inv_lazy(v[grp * 2 * half + k], v[grp * 2 * half + k + half], tw, p, four_p);
            else bfly_inv(v[grp * 2 * half + k], v[grp * 2 * half + k + half], tw, p, two_p);
          }
        }
      }
    }
    if (FUSE & WF_STORE_MUL) {
#pragma unroll
      for (int k = 0; k < R; k++) sp[pad_idx(base + k * g)] = mul_shoup_lazy(v[k], load_tw(f.mul, base + k * g), p);
    } else if (FUSE & WF_STORE_COMBINE) {
      uint64_t *cp = f.coef + slot * slot_stride;
#pragma unroll
      for (int k = 0; k < R; k++) {
        const uint32_t idx = base + k * g;
        uint64_t x = v[k] >= (p << 1) ? v[k] - (p << 1) : v[k];
        x = x >= p ? x - p : x;
        uint64_t *a = cp + pad_idx(idx);
        if ((idx & (2 * f.m - 1)) < f.m) x = add_mod(x, *a, p);
        *a = x;
      }
    } else if (FUSE & WF_STORE_NEWTON) {
#pragma unroll
      for (int k = 0; k < R; k++) {
        const uint32_t idx = base + k * g;
        uint64_t x = 0;
        if (idx < f.n) {
          x = v[k] >= (p << 1) ? v[k] - (p << 1) : v[k];
          x = x >= p ? x - p : x;
          if (idx < f.wc) x = add_mod(x, f.wr[slot * WF_WC_MAX + idx], p);
        }
        sp[pad_idx(idx)] = x;
      }
    } else {
#pragma unroll
      for (int k = 0; k < R; k++) sp[pad_idx(base + k * g)] = v[k];
    }
  }
}

// levels [s0, lg) forward: input < 4p (canonical when LAZY) natural order, output bit-reversed order, < 4p or (LAZY) any
// 64-bit representative.  Ends with a barrier.
template <bool LAZY>
__device__ __noinline__ void wf_ntt_fwd(uint64_t *buf, uint32_t slot_stride, uint32_t nslots, uint32_t nb, uint32_t lg, uint32_t s0,
                                        const Twiddle *tab, uint64_t p) {
  uint32_t s = s0;
  while (lg - s >= 4) {
    wf_pass<4, false, LAZY>(buf, slot_stride, nslots, nb, lg, s, tab, p);
    __syncthreads();
    s += 4;
  }
  if (lg - s == 3) wf_pass<3, false, LAZY>(buf, slot_stride, nslots, nb, lg, s, tab, p);
  else if (lg - s == 2) wf_pass<2, false, LAZY>(buf, slot_stride, nslots, nb, lg, s, tab, p);
  else if (lg - s == 1) wf_pass<1, false, LAZY>(buf, slot_stride, nslots, nb, lg, s, tab, p);
  __syncthreads();
}
// all lg levels inverse: input < 2p bit-reversed order, output natural order, < 2p or (LAZY) < 4p, NOT scaled.  Ends with a barrier.
template <bool LAZY>
__device__ __noinline__ void wf_ntt_inv(uint64_t *buf, uint32_t slot_stride, uint32_t nslots, uint32_t nb, uint32_t lg,
                                        const Twiddle *tab, uint64_t p) {
  uint32_t rem = lg;
  const uint32_t first = rem & 3;
  if (first == 3) wf_pass<3, true, LAZY>(buf, slot_stride, nslots, nb, lg, rem - 3, tab, p);
  else if (first == 2) wf_pass<2, true, LAZY>(buf, slot_stride, nslots, nb, lg, rem - 2, tab, p);
  else if (first == 1) wf_pass<1, true, LAZY>(buf, slot_stride, nslots, nb, lg, rem - 1, tab, p);
  if (first) __syncthreads();
  rem -= first;
  while (rem) {
    wf_pass<4, true, LAZY>(buf, slot_stride, nslots, nb, lg, rem - 4, tab, p);
    __syncthreads();
    rem -= 4;
  }
}


// ---- drivers with the element-wise steps fused into the first / last pass (see WfFuse) -------------------------------
// forward, levels [0, lg), last pass multiplies by `mul` (lg >= 5)
template <bool LAZY>
__device__ __noinline__ void wf_fwd_mul(uint64_t *buf, uint32_t stride, uint32_t nslots, uint32_t lg, const Twiddle *tab, uint64_t p,
                                        const Twiddle *mul) {
  WfFuse f;
  f.mul = mul;
  uint32_t s = 0;
  while (lg - s > 4) {
    wf_pass<4, false, LAZY>(buf, stride, nslots, 1, lg, s, tab, p);
    __syncthreads();
    s += 4;
  }
  if (lg - s == 4) wf_pass<4, false, LAZY, WF_STORE_MUL>(buf, stride, nslots, 1, lg, s, tab, p, f);
  else if (lg - s == 3) wf_pass<3, false, LAZY, WF_STORE_MUL>(buf, stride, nslots, 1, lg, s, tab, p, f);
  else if (lg - s == 2) wf_pass<2, false, LAZY, WF_STORE_MUL>(buf, stride, nslots, 1, lg, s, tab, p, f);
  else wf_pass<1, false, LAZY, WF_STORE_MUL>(buf, stride, nslots, 1, lg, s, tab, p, f);
  __syncthreads();
}
// inverse, all lg levels (lg >= 5), the last pass (levels 3..0) finishes with FUSE
template <bool LAZY, int FUSE>
__device__ __noinline__ void wf_inv_fused(uint64_t *buf, uint32_t stride, uint32_t nslots, uint32_t nb, uint32_t lg, const Twiddle *tab,
                                          uint64_t p, const WfFuse &f) {
  uint32_t rem = lg;
  const uint32_t first = rem & 3;
  if (first == 3) wf_pass<3, true, LAZY>(buf, stride, nslots, nb, lg, rem - 3, tab, p);
  else if (first == 2) wf_pass<2, true, LAZY>(buf, stride, nslots, nb, lg, rem - 2, tab, p);
  else if (first == 1) wf_pass<1, true, LAZY>(buf, stride, nslots, nb, lg, rem - 1, tab, p);
  if (first) __syncthreads();
  rem -= first;
  while (rem > 4) {
    wf_pass<4, true, LAZY>(buf, stride, nslots, nb, lg, rem - 4, tab, p);
    __syncthreads();
    rem -= 4;
  }
  wf_pass<4, true, LAZY, FUSE>(buf, stride, nslots, nb, lg, 0, tab, p, f);
  __syncthreads();
}
// forward of one tree level: levels [1, lg) of (F_hi | F_hi) read from the coefficient buffer, last pass times P-hat
template <bool LAZY>
__device__ __noinline__ void wf_level_fwd(uint64_t *buf, uint32_t stride, uint32_t nslots, uint32_t nb, uint32_t lg, const Twiddle *tab,
                                          uint64_t p, const WfFuse &f) {
  if (lg == 5) {
    wf_pass<4, false, LAZY, WF_LOAD_DUP | WF_STORE_MUL>(buf, stride, nslots, nb, lg, 1, tab, p, f);
    __syncthreads();
    return;
  }
  wf_pass<4, false, LAZY, WF_LOAD_DUP>(buf, stride, nslots, nb, lg, 1, tab, p, f);
  __syncthreads();
  uint32_t s = 5;
  while (lg - s > 4) {
    wf_pass<4, false, LAZY>(buf, stride, nslots, nb, lg, s, tab, p);
    __syncthreads();
    s += 4;
  }
  if (lg - s == 4) wf_pass<4, false, LAZY, WF_STORE_MUL>(buf, stride, nslots, nb, lg, s, tab, p, f);
  else if (lg - s == 3) wf_pass<3, false, LAZY, WF_STORE_MUL>(buf, stride, nslots, nb, lg, s, tab, p, f);
  else if (lg - s == 2) wf_pass<2, false, LAZY, WF_STORE_MUL>(buf, stride, nslots, nb, lg, s, tab, p, f);
  else wf_pass<1, false, LAZY, WF_STORE_MUL>(buf, stride, nslots, nb, lg, s, tab, p, f);
  __syncthreads();
}

__device__ __forceinline__ uint64_t canon2(uint64_t x, uint64_t p) { return x >= p ? x - p : x; }
__host__ __device__ constexpr uint32_t wf_slot_stride(uint32_t S) { return padded_words(S) + 1; }   // odd: slots land on different banks
// two polynomial buffers per slot (coefficients A, scratch / second operand B) + the small staging areas; the
// buffers that do not fit an SM live in a per-CTA global scratch range instead (S = 16384: one buffer is 136 KiB)
// n_global = 1: the scratch buffers, = 2: both buffers (S = 32768: even one buffer is 272 KiB) live in global memory.
__host__ __device__ constexpr size_t wf_smem_bytes(uint32_t S, uint32_t SL, int n_global = 0) {
  return ((size_t)(2 - n_global) * SL * wf_slot_stride(S) + (size_t)SL * (2 * WF_WC_MAX + WF_HMAX)) * 8;
}

// wr[slot][k] = sum_{i+j = k+S} u_i * v_j for k < wc: the coefficients a product modulo x^S + 1 folds back (with a minus
// sign) onto its low end.  u: padded shared-memory polynomial per slot, lu entries; v: lv entries, shared (per slot,
// padded) when v_sm != nullptr, else the global constant v_gl.
__device__ __forceinline__ void wf_wrapped(uint64_t *wr, const uint64_t *u_sm, uint32_t lu, const uint64_t *v_sm, const uint64_t *v_gl,
                                           uint32_t lv, uint32_t S, uint32_t wc, uint32_t slot_stride, uint32_t nslots,
                                           const ModConst &mc) {
  for (uint32_t t = threadIdx.x; t < wc * nslots; t += blockDim.x) {
    const uint32_t k = t / nslots, s = t - k * nslots;
    const uint32_t deg = k + S;
    Acc192 acc;
    acc.clear();
    const uint32_t i_lo = deg >= lv ? deg - lv + 1 : 0;
    for (uint32_t i = i_lo; i < lu && i <= deg; i++) {
      const uint64_t a = u_sm[s * slot_stride + pad_idx(i)];
      const uint64_t b = v_sm ? v_sm[s * slot_stride + pad_idx(deg - i)] : __ldg(v_gl + (deg - i));
      acc.mac(a, b);
    }
    wr[s * WF_WC_MAX + k] = acc.reduce(mc);
  }
}

// Newton -> monomial basis of the 16 coefficients of one leaf (nodes pt0, pt0+1, ..): Horner's rule from the top,
//   f <- f * (x - (pt0 + k)) + c_k,  k = 14 .. 0,  in registers.  Missing coefficients (beyond n) are zeros.
__device__ __forceinline__ void wf_leaf(uint64_t *f /* WF_B words in shared memory, padded-contiguous */, const Twiddle *__restrict__ pts,
                                        uint64_t p) {
  uint64_t c[WF_B];
#pragma unroll
  for (int k = 0; k < WF_B; k++) c[k] = f[k];
  // before step k, c[k+1 .. 15] hold f (constant term at c[k+1]); the step leaves the new f in c[k .. 15]:
  //   new_0 = c_k - a f_0,  new_j = f_(j-1) - a f_j,  leading coefficient unchanged
#pragma unroll
  for (int k = WF_B - 2; k >= 0; k--) {
    const Twiddle pt = load_tw(pts, k);
#pragma unroll
    for (int j = k; j < WF_B - 1; j++) c[j] = sub_mod(c[j], mul_shoup(c[j + 1], pt, p), p);
  }
#pragma unroll
  for (int k = 0; k < WF_B; k++) f[k] = c[k];
}

// Shared body: buffer A of every slot holds canonical Newton coefficients c_k (k < n, zeros beyond); on return it holds
// the monomial coefficients.  B is scratch of the same shape, hs a [nslots][WF_HMAX] staging area.
template <bool LAZY>
__device__ __forceinline__ void wf_newton_to_monomial(uint64_t *A, uint64_t *B, uint64_t *hs, const FastTables &T, uint32_t limb,
                                                      uint32_t nsl, uint32_t stride, const Twiddle *fw, const Twiddle *iv,
                                                      uint64_t p, const ModConst &mc) {
  const uint32_t n = T.n, S = T.S;
  const uint32_t npad = (n + WF_B - 1) / WF_B * WF_B;
  {   // leaves: one thread per (slot, 16-coefficient block); a leaf never straddles a pad word (16-aligned)
    const uint32_t nleaf = npad / WF_B;
    const Twiddle *pts = T.pts + (size_t)limb * npad;
    for (uint32_t t = threadIdx.x; t < nleaf * nsl; t += blockDim.x) {
      const uint32_t s = t % nsl, blk = t / nsl;
      wf_leaf(A + s * stride + pad_idx(blk * WF_B), pts + blk * WF_B, p);
    }
    __syncthreads();
  }
  uint32_t lvl = 0;
  for (uint32_t m = WF_B; m < n; m <<= 1, lvl++) {
    const uint32_t lg = 32 - __clz(m), two_m = 2 * m;          // log2(2m)
    const uint32_t nb_active = (n - m + two_m - 1) / two_m;
    const uint32_t last = nb_active - 1;
    const uint32_t h_last = min(m, n - (last * two_m + m));
    const bool shortp = h_last <= WF_HMAX;
    const uint32_t nbN = nb_active - (shortp ? 1 : 0);
    const Twiddle *Ph = T.Phat + ((size_t)limb * T.levels + lvl) * S;
    const uint64_t *Pn = T.Pnat + ((size_t)limb * T.levels + lvl) * (S / 2 + 1);
    if (shortp) {   // the trailing block's own scratch range is free: stage its high coefficients and P there
      for (uint32_t t = threadIdx.x; t < h_last * nsl; t += blockDim.x) {
        const uint32_t s = t % nsl, i = t / nsl;
        hs[s * WF_HMAX + i] = A[s * stride + pad_idx(last * two_m + m + i)];
      }
      for (uint32_t i = threadIdx.x; i <= m; i += blockDim.x) B[pad_idx(last * two_m + i)] = __ldg(Pn + i);
      if (!nbN) __syncthreads();
    }
    if (nbN) {
      // (F_hi | 0) of every transformed block: its first butterfly level is a copy, so the first pass reads (F_hi | F_hi)
      // straight from A; the last forward pass multiplies by P-hat; the last inverse pass adds F_lo and writes A
      WfFuse f;
      f.coef = A;
      f.m = m;
      f.mul = Ph;
      wf_level_fwd<LAZY>(B, stride, nsl, nbN, lg, fw, p, f);
      wf_inv_fused<LAZY, WF_STORE_COMBINE>(B, stride, nsl, nbN, lg, iv, p, f);
    }
    if (shortp) {   // trailing block: out[j] = sum_{i < h_last, 0 <= j-i <= m} hi[i] * P[j-i]
      const uint64_t *Ps = B + 0 * stride;   // slot 0's scratch holds P (staged above)
      for (uint32_t t = threadIdx.x; t < two_m * nsl; t += blockDim.x) {
        const uint32_t s = t % nsl, j = t / nsl;
        Acc192 acc;
        acc.clear();
#pragma unroll 4
        for (uint32_t i = (j > m ? j - m : 0); i < h_last && i <= j; i++)
          acc.mac(hs[s * WF_HMAX + i], Ps[pad_idx(last * two_m + j - i)]);
        uint64_t x = acc.reduce(mc);
        uint64_t *a = A + s * stride + pad_idx(last * two_m + j);
        if (j < m) x = add_mod(x, *a, p);
        *a = x;
      }
    }
    __syncthreads();
  }
}

// Interpolation of `batch` vectors of n ring elements on {0..n-1}.  Strided addressing so that the same kernel serves ring
// vectors ([element][L_R][N_R]: coef_stride = W, limb_stride = N_R, vec_stride = n*W, nslots = N_R) and per-constraint
// constants ([vector][L_R][n]: coef_stride = 1, limb_stride = n, vec_stride = L_R*n, nslots = 1).
// grid (nslots / SL, batch * L_R); SL divides nslots
template <int SL, bool LAZY>
__global__ void __launch_bounds__(512) k_interp_fast(const DevParams *__restrict__ P, FastTables T, const uint64_t *__restrict__ Y,
                                                     uint64_t *__restrict__ C, size_t coef_stride, size_t limb_stride,
                                                     size_t vec_stride, uint64_t *gB, uint64_t *gA) {
  extern __shared__ uint64_t sm[];
  const uint32_t L_R = P->L_R;
  const uint32_t v = blockIdx.y / L_R, limb = blockIdx.y - v * L_R;
  constexpr uint32_t nsl = SL;   // the host picks SL dividing the slot count
  const uint32_t slot0 = blockIdx.x * SL;
  const uint32_t n = T.n, S = T.S, wc = T.wc;
  const uint32_t stride = wf_slot_stride(S);
  // gB != nullptr: the scratch buffers of this CTA are a private range of global memory (barriers order global accesses
  // within the block just as they order shared ones; the range stays in L2)
  const size_t cta_off = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * SL * stride;
  uint64_t *A = gA ? gA + cta_off : sm, *B = gB ? gB + cta_off : sm + (size_t)SL * stride;
  uint64_t *wr = sm + (size_t)((gA ? 0 : 1) + (gB ? 0 : 1)) * SL * stride, *hs = wr + (size_t)SL * 2 * WF_WC_MAX;
  const ModConst mc = P->q[limb];
  const uint64_t p = mc.p;
  const Twiddle *fw = P->fwdq[limb], *iv = P->invq[limb];
  const size_t goff = (size_t)v * vec_stride + (size_t)limb * limb_stride + slot0;
  // y_i / i!
  const Twiddle *invfact = T.invfact + (size_t)limb * n;
  for (uint32_t t = threadIdx.x; t < S * nsl; t += blockDim.x) {
    const uint32_t s = t % nsl, i = t / nsl;
    A[s * stride + pad_idx(i)] = i < n ? mul_shoup(Y[goff + (size_t)i * coef_stride + s], load_tw(invfact, i), p) : 0;
  }
  __syncthreads();
  if (wc) wf_wrapped(wr, A, n, nullptr, T.g_nat + (size_t)limb * n, n, S, wc, stride, nsl, mc);
  __syncthreads();
  {
    WfFuse f;
    f.wr = wr;
    f.n = n;
    f.wc = wc;
    wf_fwd_mul<LAZY>(A, stride, nsl, T.logS, fw, p, T.Ghat + (size_t)limb * S);
    wf_inv_fused<LAZY, WF_STORE_NEWTON>(A, stride, nsl, 1, T.logS, iv, p, f);
  }
  wf_newton_to_monomial<LAZY>(A, B, hs, T, limb, nsl, stride, fw, iv, p, mc);
  for (uint32_t t = threadIdx.x; t < n * nsl; t += blockDim.x) {
    const uint32_t s = t % nsl, i = t / nsl;
    C[goff + (size_t)i * coef_stride + s] = A[s * stride + pad_idx(i)];
  }
}

// H[i] (i < n-1) = coefficient i of the quotient of A*B by Z, A and B given by n monomial coefficients each
// ([element][L_R][N_R]).  grid (N_R / SL, L_R); SL divides N_R
template <int SL, bool LAZY>
__global__ void __launch_bounds__(512) k_quotient_fast(const DevParams *__restrict__ P, FastTables T, const uint64_t *__restrict__ Ac,
                                                       const uint64_t *__restrict__ Bc, uint64_t *__restrict__ H, uint64_t *gB, uint64_t *gA) {
  extern __shared__ uint64_t sm[];
  const uint32_t N_R = P->N_R, L_R = P->L_R, limb = blockIdx.y;
  const size_t W = (size_t)N_R * L_R;
  constexpr uint32_t nsl = SL;
  const uint32_t slot0 = blockIdx.x * SL;
  const uint32_t n = T.n, S = T.S, wc = T.wc;
  const uint32_t stride = wf_slot_stride(S);
  const size_t cta_off = ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * SL * stride;
  uint64_t *A = gA ? gA + cta_off : sm, *B = gB ? gB + cta_off : sm + (size_t)SL * stride;
  uint64_t *wr = sm + (size_t)((gA ? 0 : 1) + (gB ? 0 : 1)) * SL * stride, *wr2 = wr + (size_t)SL * WF_WC_MAX;
  const ModConst mc = P->q[limb];
  const uint64_t p = mc.p;
  const Twiddle *fw = P->fwdq[limb], *iv = P->invq[limb];
  const size_t goff = (size_t)limb * N_R + slot0;
  for (uint32_t t = threadIdx.x; t < S * nsl; t += blockDim.x) {
    const uint32_t s = t % nsl, i = t / nsl;
    A[s * stride + pad_idx(i)] = i < n ? Ac[goff + (size_t)i * W + s] : 0;
    B[s * stride + pad_idx(i)] = i < n ? Bc[goff + (size_t)i * W + s] : 0;
  }
  __syncthreads();
  // coefficients S .. 2n-2 of A*B, directly
  if (wc) wf_wrapped(wr, A, n, B, nullptr, n, S, wc, stride, nsl, mc);
  __syncthreads();
  // A and B are adjacent: one batch of 2*SL transforms (slot index SL + s addresses B's slot s)
  if (gB) {
    wf_ntt_fwd<LAZY>(A, stride, nsl, 1, T.logS, 0, fw, p);
    wf_ntt_fwd<LAZY>(B, stride, nsl, 1, T.logS, 0, fw, p);
  } else {
    wf_ntt_fwd<LAZY>(A, stride, 2 * SL, 1, T.logS, 0, fw, p);
  }
  for (uint32_t t = threadIdx.x; t < S * nsl; t += blockDim.x) {
    const uint32_t s = t % nsl, i = t / nsl;
    uint64_t *w = A + s * stride + pad_idx(i);
    const uint64_t b = B[s * stride + pad_idx(i)];
    *w = LAZY ? mul_mod(reduce64(*w, mc), reduce64(b, mc), mc) : mul_mod(canon4(*w, p), canon4(b, p), mc);
  }
  __syncthreads();
  wf_ntt_inv<LAZY>(A, stride, nsl, 1, T.logS, iv, p);
  // u_i = coefficient 2n-2-i of A*B, i < n-1 (the dividend's top, reversed), zero-padded, into B
  const Twiddle invS = T.invS[limb];
  const uint32_t lu = n - 1;
  for (uint32_t t = threadIdx.x; t < S * nsl; t += blockDim.x) {
    const uint32_t s = t % nsl, i = t / nsl;
    uint64_t x = 0;
    if (i < lu) {
      const uint32_t k = 2 * n - 2 - i;
      x = k >= S ? wr[s * WF_WC_MAX + (k - S)] : mul_shoup(A[s * stride + pad_idx(k)], invS, p);
    }
    B[s * stride + pad_idx(i)] = x;
  }
  __syncthreads();
  // rq = u * rev(Z)^-1 mod x^(n-1): the coefficients k + S <= 2(n-2) of the full product fold back onto k
  const uint32_t wc2 = 2 * lu > S + 1 ? 2 * lu - 1 - S : 0;
  if (wc2) wf_wrapped(wr2, B, lu, nullptr, T.v_nat + (size_t)limb * n, lu, S, wc2, stride, nsl, mc);
  __syncthreads();
  wf_ntt_fwd<LAZY>(B, stride, nsl, 1, T.logS, 0, fw, p);
  {
    const Twiddle *Vh = T.Vhat + (size_t)limb * S;
    for (uint32_t t = threadIdx.x; t < S * nsl; t += blockDim.x) {
      const uint32_t s = t % nsl, i = t / nsl;
      uint64_t *w = B + s * stride + pad_idx(i);
      *w = mul_shoup_lazy(*w, load_tw(Vh, i), p);
    }
  }
  __syncthreads();
  wf_ntt_inv<LAZY>(B, stride, nsl, 1, T.logS, iv, p);
  for (uint32_t t = threadIdx.x; t < lu * nsl; t += blockDim.x) {
    const uint32_t s = t % nsl, i = t / nsl;
    const uint32_t k = lu - 1 - i;                      // H_i = rq_(n-2-i)
    uint64_t x = canon4(B[s * stride + pad_idx(k)], p);
    if (k < wc2) x = add_mod(x, wr2[s * WF_WC_MAX + k], p);
    H[goff + (size_t)i * W + s] = x;
  }
}

}  // namespace rsg
