// Forward negacyclic NTT on the FP64 pipe, for primes below 2^49 (every data-level prime of BFVDefault(8192) and
// BFVDefault(16384): 43..49 bit).  Same transform, same ordering and therefore the same canonical residues as ntt.cuh /
// SEAL's ntt_negacyclic_harvey (util/ntt.cpp:407-436, util/dwthandler.h:94-190); only the arithmetic differs.
//
// Why: the 64-bit integer butterfly costs ~11 IMAD-class instructions and the kernel saturates the one pipe that executes
// them (ncu, round 1: fmaheavy 63 % busy, math-pipe throttle the top stall) while the FP64 pipe -- 64 lanes/clk/SM on B200,
// the same issue rate as IMAD -- sits idle.  With every value held as an EXACT integer in a double the butterfly is 8 FP64
// instructions and no integer multiply at all:
//     h = RN(y*w)             l = fma(y, w, -h)          (h + l == y*w exactly)
//     q = RN(y*(w/p) + M) - M                            (M = 1.5 * 2^52: q = nearest integer, |q - y*w/p| <= 1/2 + |y| 2^-53)
//     t = fma(-q, p, h)       v = t + l                  (v == y*w - q*p exactly, |v| <= (1/2 + |y| 2^-53) p)
//     x' = x + v              y' = x - v
// Exactness: h - q*p is an integer of magnitude < 2^51, so the fma returns it unrounded; l is the exact low part of the
// product; all sums stay below 2^53.  Requirement for the rounding trick: |y * w/p| < 2^51, i.e. |y| < 2^51.
// Growth (p < 2^49, values in units of p): a reduced value is <= 0.51; each level adds |v| <= 0.5 + b/16, giving
// 0.51 -> 1.04 -> 1.61 -> 2.21 -> 2.85 -> 3.53 after 1..5 levels; multiplier inputs therefore stay below 2.85 p < 2^51 for
// passes of up to five levels, and every value below 3.6 p < 2^53.  Values are re-centred (x - rint(x/p) p, 3 instructions)
// when a pass stores them, so each pass starts from <= 0.51 p again.
#pragma once
#include "ntt.cuh"

namespace rsg {

struct TwiddleF {   // w as a double and RN(w / p)
  double w, wp;
};

constexpr double F64_MAGIC = 6755399441055744.0;   // 1.5 * 2^52

__device__ __forceinline__ TwiddleF load_twf(const TwiddleF *tab, uint32_t i) {
  const double2 v = __ldg(reinterpret_cast<const double2 *>(tab) + i);
  TwiddleF t;
  t.w = v.x;
  t.wp = v.y;
  return t;
}

// x - rint(x / p) * p: |result| <= p/2 (+ |x| 2^-53), exact for |x| < 2^51
__device__ __forceinline__ double recentre_f64(double x, double p, double pinv) {
  const double q = __dadd_rn(__fma_rn(x, pinv, F64_MAGIC), -F64_MAGIC);
  return __fma_rn(-q, p, x);
}

__device__ __forceinline__ void bfly_fwd_f64(double &x, double &y, const TwiddleF &t, double p) {
  const double h = __dmul_rn(y, t.w);
  const double l = __fma_rn(y, t.w, -h);
  const double q = __dadd_rn(__fma_rn(y, t.wp, F64_MAGIC), -F64_MAGIC);
  const double v = __dadd_rn(__fma_rn(-q, p, h), l);
  y = __dadd_rn(x, -v);
  x = __dadd_rn(x, v);
}

// canonical residue in [0, p) of an exact-integer double with |x| < 2^51
__device__ __forceinline__ uint64_t canon_f64(double x, double p, double pinv, uint64_t pi) {
  long long r = __double2ll_rn(recentre_f64(x, p, pinv));
  return (uint64_t)(r < 0 ? r + (long long)pi : r);
}
// canonical residue -> centred exact double in (-p/2, p/2]
__device__ __forceinline__ double centre_to_f64(uint64_t r, uint64_t pi) {
  const long long s = r > (pi >> 1) ? (long long)r - (long long)pi : (long long)r;
  return __ll2double_rn(s);
}

// Levels [S, S+RL) of the local block, in registers, on 2^RL elements spaced g = n >> (S+RL) apart (see ntt_pass in
// ntt.cuh for the index conventions; the padded-offset identity holds for every (RL, S) produced by f64_chain_rl).
template <int LOGN, int RL, int S, bool RECENTRE>
__device__ __forceinline__ void ntt_pass_f64(double *sm, const TwiddleF *tab, double p, double pinv, uint32_t lvl0, uint32_t blk) {
  constexpr uint32_t n = 1u << LOGN;
  constexpr int R = 1 << RL;
  constexpr uint32_t g = n >> (S + RL);
  constexpr uint32_t items = n >> RL;
  for (uint32_t item = threadIdx.x; item < items; item += blockDim.x) {
    const uint32_t o = item & (g - 1);
    const uint32_t b = item / g;
    const uint32_t base = b * (n >> S) + o;
    double *ptr = sm + pad_idx(base);
    double v[R];
#pragma unroll
    for (int k = 0; k < R; k++) v[k] = ptr[k * g + ((k * g) >> 4)];
#pragma unroll
    for (int u = 0; u < RL; u++) {
      const int half = R >> (u + 1);
      const uint32_t tbase = (1u << (lvl0 + S + u)) + (blk << (S + u)) + (b << u);
#pragma unroll
      for (int grp = 0; grp < (1 << u); grp++) {
        const TwiddleF t = load_twf(tab, tbase + grp);
#pragma unroll
        for (int k = 0; k < half; k++) bfly_fwd_f64(v[grp * 2 * half + k], v[grp * 2 * half + k + half], t, p);
      }
    }
#pragma unroll
    for (int k = 0; k < R; k++) ptr[k * g + ((k * g) >> 4)] = RECENTRE ? recentre_f64(v[k], p, pinv) : v[k];
  }
}

// pass sizes: as few passes as MAXRL allows, levels spread evenly (14 -> 5,5,4 or 4,4,3,3; 13 -> 5,4,4)
__host__ __device__ constexpr int f64_chain_rl(int remain, int maxrl) {
  const int passes = (remain + maxrl - 1) / maxrl;
  return (remain + passes - 1) / passes;
}

template <int LOGN, int S, int MAXRL>
struct PassChainF {
  static __device__ __forceinline__ void fwd(double *sm, const TwiddleF *tab, double p, double pinv, uint32_t lvl0, uint32_t blk) {
    constexpr int REMAIN = LOGN - S;
    if constexpr (REMAIN > 0) {
      constexpr int RL = f64_chain_rl(REMAIN, MAXRL);
      // the last pass leaves its values un-centred: the caller canonicalises them anyway
      ntt_pass_f64<LOGN, RL, S, (REMAIN > RL)>(sm, tab, p, pinv, lvl0, blk);
      __syncthreads();
      PassChainF<LOGN, S + RL, MAXRL>::fwd(sm, tab, p, pinv, lvl0, blk);
    }
  }
};

// All LOGN levels of the local block.  Input: centred exact doubles (|x| <= 0.51 p); output: exact doubles, |x| < 3.6 p.
template <int LOGN, int MAXRL>
__device__ __forceinline__ void ntt_forward_smem_f64(double *sm, const TwiddleF *tab, double p, double pinv, uint32_t lvl0,
                                                     uint32_t blk) {
  PassChainF<LOGN, 0, MAXRL>::fwd(sm, tab, p, pinv, lvl0, blk);
}

}  // namespace rsg
