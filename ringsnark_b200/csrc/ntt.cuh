// Negacyclic NTT of one polynomial held in shared memory, SEAL-compatible conventions:
//   forward : natural order in -> bit-reversed order out   (util/ntt.cpp:407-436, util/dwthandler.h:94-190)
//   inverse : bit-reversed in  -> natural order out, x N^-1 (util/ntt.cpp:452-474, util/dwthandler.h:202-330)
//   psi     : SEAL's minimal primitive 2N-th root           (util/numth.cpp:386-412), tables built on the host.
// Not a translation of SEAL's loops: the transform is cut into register-resident radix-16 passes (four
// butterfly levels per shared-memory round trip) over a padded, bank-conflict-free layout; values stay lazy in
// [0, 4p) (forward) / [0, 2p) (inverse) and are canonicalised once at the end, which yields the same residues.
//
// Twiddle tables (device, one per prime): fwd[(1<<s) + g] = psi^bitrev(...) for level s (gap N>>(s+1)), group g
// -- SEAL's root_powers_ order; inv[(1<<s) + g] = fwd[(1<<s) + g]^-1.  Each entry is a Shoup pair.
#pragma once
#include "modarith.cuh"

namespace rsg {

// one pad word per 16: keeps every access pattern of every pass on 16 distinct 8-byte bank pairs per half-warp
__device__ __forceinline__ uint32_t pad_idx(uint32_t i) { return i + (i >> 4); }
__host__ __device__ constexpr uint32_t padded_words(uint32_t n) { return n + (n >> 4); }

__device__ __forceinline__ Twiddle load_tw(const Twiddle *tab, uint32_t i) {
  const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(tab) + i);
  Twiddle t;
  t.w = v.x;
  t.wq = v.y;
  return t;
}

// Cooley-Tukey butterfly, lazy: x, y in [0, 4p) -> [0, 4p)
__device__ __forceinline__ void bfly_fwd(uint64_t &x, uint64_t &y, const Twiddle &t, uint64_t p, uint64_t two_p) {
  uint64_t u = x >= two_p ? x - two_p : x;
  uint64_t v = mul_shoup_lazy(y, t, p);
  x = u + v;
  y = u - v + two_p;
}
// Gentleman-Sande butterfly, lazy: x, y in [0, 2p) -> [0, 2p)
__device__ __forceinline__ void bfly_inv(uint64_t &x, uint64_t &y, const Twiddle &t, uint64_t p, uint64_t two_p) {
  uint64_t s = x + y;
  uint64_t d = x - y + two_p;
  x = s >= two_p ? s - two_p : s;
  y = mul_shoup_lazy(d, t, p);
}

// Lazy forward butterfly for primes below 2^58: no range correction at all.  The Shoup quotient is taken from three of
// the four 32x32 partial products (one IMAD.WIDE + two IMAD.HI instead of a full 64x64 high product), which
// under-estimates it by at most 2, so y*w - q*p lies in [0, 4p) for ANY 64-bit y; with the constant 4p added to the
// difference every level raises the bound of the values by 4p: p -> (4*levels + 1) p <= 61 p < 2^64 for 15 levels.
// The caller canonicalises once at the end (reduce64).  Same residues as SEAL's schedule.
__device__ __forceinline__ uint64_t mul_shoup_approx(uint64_t y, const Twiddle &t, uint64_t p) {
  const uint32_t y0 = (uint32_t)y, y1 = (uint32_t)(y >> 32), q0 = (uint32_t)t.wq, q1 = (uint32_t)(t.wq >> 32);
  const uint64_t q = (uint64_t)y1 * q1 + __umulhi(y1, q0) + __umulhi(y0, q1);
  return y * t.w - q * p;
}
__device__ __forceinline__ void bfly_fwd_lazy(uint64_t &x, uint64_t &y, const Twiddle &t, uint64_t p, uint64_t four_p) {
  const uint64_t v = mul_shoup_approx(y, t, p);
  y = x - v + four_p;
  x = x + v;
}

// Gentleman-Sande butterfly on [0, 4p) values with the three-product Shoup quotient (one IMAD.WIDE + two IMAD.HI instead of a
// full 64x64 high product): x, y in [0, 4p) -> [0, 4p), p < 2^61.
__device__ __forceinline__ void bfly_inv_lazy(uint64_t &x, uint64_t &y, const Twiddle &t, uint64_t p, uint64_t four_p) {
  const uint64_t s = x + y;
  const uint64_t d = x - y + four_p;
  x = s >= four_p ? s - four_p : s;
  y = mul_shoup_approx(d, t, p);
}

// One pass = RL consecutive levels [S, S+RL) done in registers on 2^RL elements spaced g = n >> (S+RL) apart.
// n = local transform size (1 << LOGN); `lvl0` = levels already applied outside (0 unless the polynomial was
// pre-split in global memory), `blk` = index of this local block among the 1 << lvl0 blocks.
// Padded addresses: pad_idx(base + k*g) == pad_idx(base) + k*g + ((k*g) >> 4) for every pass shape used here (the low
// four bits of base and of k*g never carry), so the 2^RL offsets are compile-time constants.
template <int LOGN, int RL, int S, bool INVERSE, bool LAZY>
__device__ __forceinline__ void ntt_pass(uint64_t *sm, const Twiddle *tab, uint64_t p, uint32_t lvl0, uint32_t blk) {
  constexpr uint32_t n = 1u << LOGN;
  constexpr int R = 1 << RL;
  constexpr uint32_t g = n >> (S + RL);             // element stride inside an item
  constexpr uint32_t items = n >> RL;
  const uint64_t two_p = p << 1, four_p = p << 2;
  for (uint32_t item = threadIdx.x; item < items; item += blockDim.x) {
    const uint32_t o = item & (g - 1);
    const uint32_t b = item / g;                    // block index at level S (g is a power of two)
    const uint32_t base = b * (n >> S) + o;
    uint64_t *ptr = sm + pad_idx(base);
    uint64_t v[R];
#pragma unroll
    for (int k = 0; k < R; k++) v[k] = ptr[k * g + ((k * g) >> 4)];
    if (!INVERSE) {
#pragma unroll
      for (int u = 0; u < RL; u++) {
        const int half = R >> (u + 1);
        const uint32_t tbase = (1u << (lvl0 + S + u)) + (blk << (S + u)) + (b << u);
#pragma unroll
        for (int grp = 0; grp < (1 << u); grp++) {
          const Twiddle t = load_tw(tab, tbase + grp);
#pragma unroll
          for (int k = 0; k < half; k++) {
            if (LAZY) bfly_fwd_lazy(v[grp * 2 * half + k], v[grp * 2 * half + k + half], t, p, four_p);
            else bfly_fwd(v[grp * 2 * half + k], v[grp * 2 * half + k + half], t, p, two_p);
          }
        }
      }
    } else {
#pragma unroll
      for (int u = RL - 1; u >= 0; u--) {
        const int half = R >> (u + 1);
        const uint32_t tbase = (1u << (lvl0 + S + u)) + (blk << (S + u)) + (b << u);
#pragma unroll
        for (int grp = 0; grp < (1 << u); grp++) {
          const Twiddle t = load_tw(tab, tbase + grp);
#pragma unroll
          for (int k = 0; k < half; k++) {
            if (LAZY) bfly_inv_lazy(v[grp * 2 * half + k], v[grp * 2 * half + k + half], t, p, four_p);
            else bfly_inv(v[grp * 2 * half + k], v[grp * 2 * half + k + half], t, p, two_p);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < R; k++) ptr[k * g + ((k * g) >> 4)] = v[k];
  }
}

template <int LOGN, int S, bool INVERSE, bool LAZY>
struct PassChain {   // levels [S, LOGN): radix-16 passes while four levels remain, then one pass for the rest
  static __device__ __forceinline__ void fwd(uint64_t *sm, const Twiddle *tab, uint64_t p, uint32_t lvl0, uint32_t blk) {
    constexpr int REMAIN = LOGN - S;
    if constexpr (REMAIN > 0) {
      constexpr int RL = REMAIN >= 4 ? 4 : REMAIN;
      ntt_pass<LOGN, RL, S, false, LAZY>(sm, tab, p, lvl0, blk);
      __syncthreads();
      PassChain<LOGN, S + RL, INVERSE, LAZY>::fwd(sm, tab, p, lvl0, blk);
    }
  }
  // inverse: levels run LOGN-1 .. 0; the odd-sized pass comes first (at the top levels)
  static __device__ __forceinline__ void inv(uint64_t *sm, const Twiddle *tab, uint64_t p, uint32_t lvl0, uint32_t blk) {
    constexpr int REMAIN = LOGN - S;   // here S counts the levels already undone from the top
    if constexpr (REMAIN > 0) {
      constexpr int RL = (REMAIN % 4) ? (REMAIN % 4) : 4;
      ntt_pass<LOGN, RL, REMAIN - RL, true, LAZY>(sm, tab, p, lvl0, blk);
      __syncthreads();
      PassChain<LOGN, S + RL, INVERSE, LAZY>::inv(sm, tab, p, lvl0, blk);
    }
  }
};

// All LOGN levels of the local block, forward.  Input canonical (or < 4p when !LAZY); output lazy: < 4p (!LAZY) or
// < (4*LOGN + 1) p (LAZY, primes below 2^58 only).
template <int LOGN, bool LAZY = false>
__device__ __forceinline__ void ntt_forward_smem(uint64_t *sm, const Twiddle *tab, uint64_t p, uint32_t lvl0,
                                                 uint32_t blk) {
  PassChain<LOGN, 0, false, LAZY>::fwd(sm, tab, p, lvl0, blk);
}

// All LOGN levels, inverse (levels run LOGN-1 .. 0).  Input in [0, 2p), output lazy in [0, 2p) -- or [0, 4p) with LAZY
// (bfly_inv_lazy, p < 2^61) --, NOT yet scaled.
template <int LOGN, bool LAZY = false>
__device__ __forceinline__ void ntt_inverse_smem(uint64_t *sm, const Twiddle *tab, uint64_t p, uint32_t lvl0,
                                                 uint32_t blk) {
  PassChain<LOGN, 0, true, LAZY>::inv(sm, tab, p, lvl0, blk);
}

// ---- inverse transform of a batch-encoded slot vector that uses the first matrix row only (N_R <= N_E / 2) ------------
// BatchEncoder's index map (batchencoder.cpp:64-88) sends slot i of the first row to position bitrev((3^i mod 2N - 1) / 2).
// 3 generates the residues = 1, 3 mod 8, so (3^i - 1) / 2 = 0, 1 mod 4: bit 1 of the natural index -- the second-highest bit of
// the bit-reversed position -- is always 0.  Quarters 1 and 3 of the transform's input are therefore structurally zero:
//   * every pass below the top two levels works inside a quarter: only the items of quarters 0 and 2 are run (half the
//     butterflies of 12 of the 14 levels at N = 2^14);
//   * in the last pass the butterflies of levels 3, 2 on the zero quarters vanish and level 1 (quarter 0 <-> 1, 2 <-> 3)
//     degenerates to a copy and one multiplication.
// Same residues as the dense transform: a butterfly on (0, 0) is (0, 0) and on (x, 0) is (x, x w).
template <int LOGN, int RL, int S, bool LAZY>
__device__ __forceinline__ void ntt_pass_inv_q02(uint64_t *sm, const Twiddle *tab, uint64_t p) {
  static_assert(S >= 2, "the pass must stay inside a quarter");
  constexpr uint32_t n = 1u << LOGN;
  constexpr int R = 1 << RL;
  constexpr uint32_t g = n >> (S + RL);
  constexpr uint32_t items = n >> RL, quarter = items / 4;
  const uint64_t two_p = p << 1, four_p = p << 2;
  for (uint32_t it = threadIdx.x; it < 2 * quarter; it += blockDim.x) {
    const uint32_t item = it < quarter ? it : it + quarter;     // items are block-major: quarter q = [q, q+1) * items/4
    const uint32_t o = item & (g - 1);
    const uint32_t b = item / g;
    const uint32_t base = b * (n >> S) + o;
    uint64_t *ptr = sm + pad_idx(base);
    uint64_t v[R];
#pragma unroll
    for (int k = 0; k < R; k++) v[k] = ptr[k * g + ((k * g) >> 4)];
#pragma unroll
    for (int u = RL - 1; u >= 0; u--) {
      const int half = R >> (u + 1);
      const uint32_t tbase = (1u << (S + u)) + (b << u);
#pragma unroll
      for (int grp = 0; grp < (1 << u); grp++) {
        const Twiddle t = load_tw(tab, tbase + grp);
#pragma unroll
        for (int k = 0; k < half; k++) {
          if (LAZY) bfly_inv_lazy(v[grp * 2 * half + k], v[grp * 2 * half + k + half], t, p, four_p);
          else bfly_inv(v[grp * 2 * half + k], v[grp * 2 * half + k + half], t, p, two_p);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < R; k++) ptr[k * g + ((k * g) >> 4)] = v[k];
  }
}
// last pass: levels 3..0 on 16 elements spaced n/16 apart; elements 4..7 and 12..15 (quarters 1, 3) are zero on entry
template <int LOGN, bool LAZY>
__device__ __forceinline__ void ntt_pass_inv_top_q02(uint64_t *sm, const Twiddle *tab, uint64_t p) {
  constexpr uint32_t n = 1u << LOGN;
  constexpr uint32_t g = n >> 4;
  const uint64_t two_p = p << 1, four_p = p << 2;
  for (uint32_t item = threadIdx.x; item < g; item += blockDim.x) {
    uint64_t *ptr = sm + pad_idx(item);
    uint64_t v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = ((k >> 2) & 1) ? 0 : ptr[k * g + ((k * g) >> 4)];
    // level 3 (pairs k, k+1) and level 2 (pairs k, k+2) inside the non-zero quarters only
#pragma unroll
    for (int u = 3; u >= 2; u--) {
      const int half = 16 >> (u + 1);
#pragma unroll
      for (int grp = 0; grp < (1 << u); grp++) {
        if (((grp * 2 * half) >> 2) & 1) continue;   // this group lies in quarter 1 or 3
        const Twiddle t = load_tw(tab, (1u << u) + grp);
#pragma unroll
        for (int k = 0; k < half; k++) {
          if (LAZY) bfly_inv_lazy(v[grp * 2 * half + k], v[grp * 2 * half + k + half], t, p, four_p);
          else bfly_inv(v[grp * 2 * half + k], v[grp * 2 * half + k + half], t, p, two_p);
        }
      }
    }
    // level 1: (x, 0) -> (x, x w)
#pragma unroll
    for (int grp = 0; grp < 2; grp++) {
      const Twiddle t = load_tw(tab, 2 + grp);
#pragma unroll
      for (int k = 0; k < 4; k++) v[grp * 8 + k + 4] = LAZY ? mul_shoup_approx(v[grp * 8 + k], t, p) : mul_shoup_lazy(v[grp * 8 + k], t, p);
    }
    {   // level 0
      const Twiddle t = load_tw(tab, 1);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        if (LAZY) bfly_inv_lazy(v[k], v[k + 8], t, p, four_p);
        else bfly_inv(v[k], v[k + 8], t, p, two_p);
      }
    }
#pragma unroll
    for (int k = 0; k < 16; k++) ptr[k * g + ((k * g) >> 4)] = v[k];
  }
}
template <int LOGN, int S, bool LAZY>
struct PassChainInvQ02 {   // S counts the levels already undone from the bottom, as in PassChain::inv
  static __device__ __forceinline__ void run(uint64_t *sm, const Twiddle *tab, uint64_t p) {
    constexpr int REMAIN = LOGN - S;
    if constexpr (REMAIN > 4) {
      constexpr int RL = (REMAIN % 4) ? (REMAIN % 4) : 4;
      ntt_pass_inv_q02<LOGN, RL, REMAIN - RL, LAZY>(sm, tab, p);
      __syncthreads();
      PassChainInvQ02<LOGN, S + RL, LAZY>::run(sm, tab, p);
    } else {
      static_assert(REMAIN == 4, "needs at least eight levels");
      ntt_pass_inv_top_q02<LOGN, LAZY>(sm, tab, p);
      __syncthreads();
    }
  }
};
// Input: canonical values in quarters 0 and 2 of the bit-reversed array (quarters 1, 3 are not read); output as
// ntt_inverse_smem: natural order, lazy, not scaled.
template <int LOGN, bool LAZY>
__device__ __forceinline__ void ntt_inverse_smem_q02(uint64_t *sm, const Twiddle *tab, uint64_t p) {
  PassChainInvQ02<LOGN, 0, LAZY>::run(sm, tab, p);
}

}  // namespace rsg
