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
  // blocked mode (`big`): n beyond what one transform of size TS <= N_E serves (n <= 2*TS).  Polynomials are cut into blocks
  // of h = TS/2 coefficients; S is then the coefficient-buffer size (2h or 4h) and only the tree levels with 2m <= TS have a
  // Phat / Pnat entry.  Constants as plain residues: the partial products of an output block are summed before one reduction.
  const uint64_t *Gblk;          // [L_R][nx][TS]  NTT_TS(block j of (-1)^k / k!) / TS
  const uint64_t *Ptop;          // [L_R][2][TS]   NTT_TS(block j of prod_{x<TS}(X - x) - X^TS) / TS   (n > TS)
  const uint64_t *Vblk;          // [L_R][nx][TS]  NTT_TS(block j of rev(Z)^-1 mod x^(n-1)) / TS
  Twiddle invTS[MAX_LR];         // 1 / TS
  uint32_t big, TS, logTS, nx;
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
                                                      uint64_t p, const ModConst &mc, uint32_t max_tr = 0xffffffffu) {
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
  for (uint32_t m = WF_B; m < n && 2 * m <= max_tr; m <<= 1, lvl++) {   // max_tr: the largest transform (blocked mode stops early)
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


// ---- blocked mode: n beyond one transform (FastTables::big) ---------------------------------------------------------
// A product whose length exceeds TS (the largest transform the ring primes support, or the largest the buffers are laid out
// for) is assembled from block products: operands cut into blocks of h = TS/2 coefficients, each block transformed once at
// size TS (a block product has < 2h coefficients: nothing wraps), the partial products of output block k summed in the
// transform domain (192-bit accumulator, one reduction) and inverse-transformed once; output block k covers coefficients
// [k*h, k*h + 2h).  C5's headline n = 2^16 at N_E = 2^15 is four blocks of 16384.  The tree levels with 2m <= TS are the ones
// above (wf_newton_to_monomial); the one level beyond, m = TS, is F_lo + x^m F_hi + (P - x^m) F_hi with P - x^m as two blocks.
// One CTA walks slots (persistent grid); its buffers are a private range of global memory: A (coefficients), X / Y
// (transformed blocks; X doubles as the tree's scratch), U, Tb.  tests/test_witness_fast_model.py holds the same steps in Python.
__host__ __device__ constexpr size_t wf_big_words(uint32_t Sb, uint32_t TS) {
  return 2 * ((size_t)padded_words(Sb) + 16) + 2 * ((size_t)padded_words(2 * Sb) + 16) + (size_t)padded_words(TS) + 16;
}
struct WfBig {
  uint64_t *A, *X, *Y, *U, *Tb;
  __device__ __forceinline__ WfBig(uint64_t *base, uint32_t Sb, uint32_t TS) {
    A = base;
    U = A + padded_words(Sb) + 16;
    X = U + padded_words(Sb) + 16;
    Y = X + padded_words(2 * Sb) + 16;
    Tb = Y + padded_words(2 * Sb) + 16;
  }
};
// X[j*TS + i] = src[off + j*h + i] for i < h while j*h + i < len, zero elsewhere (j < nblk).  Ends with a barrier.
__device__ __forceinline__ void wf_blk_spread(uint64_t *X, const uint64_t *src, uint32_t off, uint32_t len, uint32_t nblk, uint32_t lgT) {
  const uint32_t TS = 1u << lgT, h = TS >> 1;
  for (uint32_t t = threadIdx.x; t < (nblk << lgT); t += blockDim.x) {
    const uint32_t j = t >> lgT, i = t & (TS - 1), pos = j * h + i;
    X[pad_idx(t)] = (i < h && pos < len) ? src[pad_idx(off + pos)] : 0;
  }
  __syncthreads();
}
// Tb = sum_{i+j=k, i<nx, j<nc} X_i * C_j (canonical; times *scale when given).  C: plain residues in global memory
// ([nc][TS], CONSTC) or a second buffer of transformed blocks.  Ends with a barrier.
template <bool LAZY, bool CONSTC>
__device__ __forceinline__ void wf_blk_mac(uint64_t *Tb, const uint64_t *X, uint32_t nx, const uint64_t *Cc, uint32_t nc, uint32_t k,
                                           uint32_t lgT, const ModConst &mc, const Twiddle *scale) {
  const uint32_t TS = 1u << lgT;
  const uint32_t i_lo = k >= nc ? k - nc + 1 : 0, i_hi = min(k, nx - 1);
  for (uint32_t e = threadIdx.x; e < TS; e += blockDim.x) {
    Acc192 acc;
    acc.clear();
    for (uint32_t i = i_lo; i <= i_hi; i++) {
      const uint64_t xv = X[pad_idx((i << lgT) + e)];
      const uint64_t a = LAZY ? reduce64(xv, mc) : canon4(xv, mc.p);
      uint64_t b;
      if (CONSTC) {
        b = __ldg(Cc + ((size_t)(k - i) << lgT) + e);
      } else {
        const uint64_t yv = Cc[pad_idx(((k - i) << lgT) + e)];
        b = LAZY ? reduce64(yv, mc) : canon4(yv, mc.p);
      }
      acc.mac(a, b);
    }
    uint64_t r = acc.reduce(mc);
    if (scale) r = mul_shoup(r, *scale, mc.p);
    Tb[pad_idx(e)] = r;
  }
  __syncthreads();
}

// Interpolation, blocked mode: same arguments as k_interp_fast; `items` = nslots * batch * L_R, grid-stride over them.
template <bool LAZY>
__global__ void __launch_bounds__(256) k_interp_big(const DevParams *__restrict__ P, FastTables T, const uint64_t *__restrict__ Y,
                                                    uint64_t *__restrict__ C, size_t coef_stride, size_t limb_stride, size_t vec_stride,
                                                    uint32_t nslots, uint32_t items, uint64_t *scratch) {
  __shared__ uint64_t hs[WF_HMAX];
  const uint32_t L_R = P->L_R, n = T.n, Sb = T.S, TS = T.TS, lgT = T.logTS, h = TS >> 1, nx = T.nx;
  const WfBig W((uint64_t *)scratch + (size_t)blockIdx.x * wf_big_words(Sb, TS), Sb, TS);
  for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
    const uint32_t slot = item % nslots, vl = item / nslots, v = vl / L_R, limb = vl - v * L_R;
    const ModConst mc = P->q[limb];
    const uint64_t p = mc.p;
    const Twiddle *fw = P->fwdq[limb], *iv = P->invq[limb];
    const size_t goff = (size_t)v * vec_stride + (size_t)limb * limb_stride + slot;
    const Twiddle *invfact = T.invfact + (size_t)limb * n;
    for (uint32_t t = threadIdx.x; t < Sb; t += blockDim.x)
      W.A[pad_idx(t)] = t < n ? mul_shoup(Y[goff + (size_t)t * coef_stride], load_tw(invfact, t), p) : 0;
    __syncthreads();
    // Newton coefficients: the low n coefficients of (y_i / i!) * ((-1)^k / k!)
    wf_blk_spread(W.X, W.A, 0, n, nx, lgT);
    wf_ntt_fwd<LAZY>(W.X, 0, 1, nx, lgT, 0, fw, p);
    const uint64_t *G = T.Gblk + ((size_t)limb * nx << lgT);
    for (uint32_t k = 0; k < nx; k++) {
      wf_blk_mac<LAZY, true>(W.Tb, W.X, nx, G, nx, k, lgT, mc, nullptr);
      wf_ntt_inv<LAZY>(W.Tb, 0, 1, 1, lgT, iv, p);
      for (uint32_t i = threadIdx.x; i < TS; i += blockDim.x) {
        const uint32_t pos = k * h + i;
        if (pos >= Sb) continue;
        uint64_t x = 0;
        if (pos < n) {
          x = canon4(W.Tb[pad_idx(i)], p);
          if (i < h && k) x = add_mod(x, W.A[pad_idx(pos)], p);     // the lower half overlaps the previous block's upper half
        }
        W.A[pad_idx(pos)] = x;
      }
      __syncthreads();
    }
    wf_newton_to_monomial<LAZY>(W.A, W.X, hs, T, limb, 1, 0, fw, iv, p, mc, TS);
    if (n > TS) {   // the level m = TS: F_lo + x^m F_hi + (P - x^m) * F_hi
      const uint32_t nf = (n - TS + h - 1) / h;
      wf_blk_spread(W.X, W.A, TS, n - TS, nf, lgT);
      wf_ntt_fwd<LAZY>(W.X, 0, 1, nf, lgT, 0, fw, p);
      const uint64_t *Pt = T.Ptop + ((size_t)limb * 2 << lgT);
      for (uint32_t k = 0; k <= nf; k++) {
        wf_blk_mac<LAZY, true>(W.Tb, W.X, nf, Pt, 2, k, lgT, mc, nullptr);
        wf_ntt_inv<LAZY>(W.Tb, 0, 1, 1, lgT, iv, p);
        for (uint32_t i = threadIdx.x; i < TS; i += blockDim.x) {
          uint64_t *a = W.A + pad_idx(k * h + i);
          *a = add_mod(canon4(W.Tb[pad_idx(i)], p), *a, p);
        }
        __syncthreads();
      }
    }
    for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) C[goff + (size_t)t * coef_stride] = W.A[pad_idx(t)];
    __syncthreads();
  }
}

// Quotient, blocked mode: H[i] (i < n-1) as k_quotient_fast.  items = N_R * L_R.
template <bool LAZY>
__global__ void __launch_bounds__(256) k_quotient_big(const DevParams *__restrict__ P, FastTables T, const uint64_t *__restrict__ Ac,
                                                      const uint64_t *__restrict__ Bc, uint64_t *__restrict__ H, uint32_t items,
                                                      uint64_t *scratch) {
  const uint32_t N_R = P->N_R, L_R = P->L_R, n = T.n, Sb = T.S, TS = T.TS, lgT = T.logTS, h = TS >> 1, nx = T.nx;
  const size_t Wd = (size_t)N_R * L_R;
  const WfBig W((uint64_t *)scratch + (size_t)blockIdx.x * wf_big_words(Sb, TS), Sb, TS);
  const uint32_t lu = n - 1, nu = (lu + h - 1) / h;
  for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
    const uint32_t slot = item % N_R, limb = item / N_R;
    const ModConst mc = P->q[limb];
    const uint64_t p = mc.p;
    const Twiddle *fw = P->fwdq[limb], *iv = P->invq[limb];
    const size_t goff = (size_t)limb * N_R + slot;
    for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) {
      W.A[pad_idx(t)] = Ac[goff + (size_t)t * Wd];
      W.U[pad_idx(t)] = Bc[goff + (size_t)t * Wd];
    }
    __syncthreads();
    wf_blk_spread(W.X, W.A, 0, n, nx, lgT);
    wf_blk_spread(W.Y, W.U, 0, n, nx, lgT);
    for (uint32_t t = threadIdx.x; t < Sb; t += blockDim.x) { W.U[pad_idx(t)] = 0; W.A[pad_idx(t)] = 0; }
    wf_ntt_fwd<LAZY>(W.X, 0, 1, nx, lgT, 0, fw, p);
    wf_ntt_fwd<LAZY>(W.Y, 0, 1, nx, lgT, 0, fw, p);
    // u_i = coefficient 2n-2-i of A*B (i < n-1): the output blocks that reach position n and beyond
    const Twiddle invTS = T.invTS[limb];
    for (uint32_t k = 0; k + 1 < 2 * nx; k++) {
      if (k * h + TS <= n) continue;
      wf_blk_mac<LAZY, false>(W.Tb, W.X, nx, W.Y, nx, k, lgT, mc, &invTS);
      wf_ntt_inv<LAZY>(W.Tb, 0, 1, 1, lgT, iv, p);
      for (uint32_t i = threadIdx.x; i < TS; i += blockDim.x) {
        const uint32_t pos = k * h + i;
        if (pos >= n && pos <= 2 * n - 2) {
          uint64_t *u = W.U + pad_idx(2 * n - 2 - pos);
          *u = add_mod(canon4(W.Tb[pad_idx(i)], p), *u, p);
        }
      }
      __syncthreads();
    }
    // rq = u * rev(Z)^-1 mod x^(n-1), accumulated into A
    wf_blk_spread(W.X, W.U, 0, lu, nu, lgT);
    wf_ntt_fwd<LAZY>(W.X, 0, 1, nu, lgT, 0, fw, p);
    const uint64_t *V = T.Vblk + ((size_t)limb * nx << lgT);
    for (uint32_t k = 0; k < nu; k++) {
      wf_blk_mac<LAZY, true>(W.Tb, W.X, nu, V, nu, k, lgT, mc, nullptr);
      wf_ntt_inv<LAZY>(W.Tb, 0, 1, 1, lgT, iv, p);
      for (uint32_t i = threadIdx.x; i < TS; i += blockDim.x) {
        const uint32_t pos = k * h + i;
        if (pos < lu) {
          uint64_t *a = W.A + pad_idx(pos);
          *a = add_mod(canon4(W.Tb[pad_idx(i)], p), *a, p);
        }
      }
      __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < lu; i += blockDim.x) H[goff + (size_t)i * Wd] = W.A[pad_idx(lu - 1 - i)];   // H_i = rq_(n-2-i)
    __syncthreads();
  }
}

}  // namespace rsg
