// Forward negacyclic NTT on the FP64 pipe, for primes below 2^49 (every data-level prime of BFVDefault(8192) and
// BFVDefault(16384): 43..49 bit).  Same transform, same ordering and therefore the same canonical residues as ntt.cuh /
// SEAL's ntt_negacyclic_harvey (util/ntt.cpp:407-436, util/dwthandler.h:94-190); only the arithmetic differs.
//
// Why: the 64-bit integer butterfly costs ~11 IMAD-class instructions and the kernel saturates the one pipe that executes
// them (ncu, round 1: fmaheavy 63 % busy, math-pipe throttle the top stall) while the FP64 pipe -- 64 lanes/clk/SM on B200,
// the same issue rate as IMAD -- sits idle.  With every value held as an EXACT integer in a double the butterfly is 8 FP64
// instructions and no integer multiply at all:
//     h = RN(y*w)             l = fma(y, w, -h)          (h + l == y*w exactly)
//     q = RN(y*wp + M) - M,  wp = RN(w * RN(1/p))        (M = 1.5 * 2^52: q = nearest integer, |q - y*w/p| <= 1/2 + |y| 2^-52)
//     t = fma(-q, p, h)       v = t + l                  (v == y*w - q*p exactly, |v| <= (1/2 + |y| 2^-52) p)
//     x' = x + v              y' = x - v
// Exactness: h - q*p is an integer of magnitude < 2^51, so the fma returns it unrounded; l is the exact low part of the
// product; all sums stay below 2^53.  Requirement for the rounding trick: |y * w/p| < 2^51, i.e. |y| < 2^51.
// Growth (p < 2^49, values in units of p): a re-centred value is <= 0.501; each level adds |v| <= 0.5 + b/8, giving
// 0.501 -> 1.07 -> 1.70 -> 2.41 -> 3.22 after 1..4 levels; multiplier inputs therefore stay below 2.41 p < 2^51 and every
// value below 3.22 p < 2^51 for passes of up to four levels.  Values are re-centred (x - rint(x/p) p, 3 instructions) when a
// pass stores them, so each pass starts from <= 0.501 p again.
// Twiddles: one double per entry (w, exact); wp is formed per pass with one multiply (a twiddle serves 2^k butterflies).
// The twiddles of the NEXT pass are requested before the barrier that ends the current one, so their L2 latency is hidden
// behind the barrier wait (ncu, first FP64 version: long-scoreboard on twiddle loads was the top stall, FP64 pipe 37 %).
#pragma once
#include "ntt.cuh"

namespace rsg {

constexpr double F64_MAGIC = 6755399441055744.0;   // 1.5 * 2^52

// x - rint(x / p) * p: |result| <= p/2 (+ |x| 2^-53), exact for |x| < 2^51
__device__ __forceinline__ double recentre_f64(double x, double p, double pinv) {
  const double q = __dadd_rn(__fma_rn(x, pinv, F64_MAGIC), -F64_MAGIC);
  return __fma_rn(-q, p, x);
}

__device__ __forceinline__ void bfly_fwd_f64(double &x, double &y, double w, double wp, double p) {
  const double h = __dmul_rn(y, w);
  const double l = __fma_rn(y, w, -h);
  const double q = __dadd_rn(__fma_rn(y, wp, F64_MAGIC), -F64_MAGIC);
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

// volatile: keeps the request where it is written (ahead of the barrier) instead of being sunk next to its first use
__device__ __forceinline__ double ldg_f64_here(const double *p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// The 2^RL - 1 twiddles one item of a pass over levels [S, S+RL) needs, in level order (see ntt_pass in ntt.cuh).
template <int LOGN, int RL, int S>
struct PassTwF {
  double w[(1 << RL) - 1];
  __device__ __forceinline__ void load(const double *tab, uint32_t lvl0, uint32_t blk, uint32_t item) {
    constexpr uint32_t g = (1u << LOGN) >> (S + RL);
    const uint32_t b = item / g;
#pragma unroll
    for (int u = 0; u < RL; u++) {
      const uint32_t tbase = (1u << (lvl0 + S + u)) + (blk << (S + u)) + (b << u);
#pragma unroll
      for (int grp = 0; grp < (1 << u); grp++) w[(1 << u) - 1 + grp] = ldg_f64_here(tab + tbase + grp);
    }
  }
};

// L1 prefetch of every twiddle this thread will need in the pass over levels [S, S+RL) (all its items): issued one pass
// ahead, it costs no registers and turns the L2 round trips of the late levels (a 128 KiB table slice per limb, which does
// not survive in L1 between CTAs) into L1 hits.
template <int LOGN, int RL, int S>
__device__ __forceinline__ void prefetch_pass_tw(const double *tab, uint32_t lvl0, uint32_t blk) {
  constexpr uint32_t g = (1u << LOGN) >> (S + RL);
  constexpr uint32_t items = (1u << LOGN) >> RL;
  for (uint32_t item = threadIdx.x; item < items; item += blockDim.x) {
    const uint32_t b = item / g;
#pragma unroll
    for (int u = 0; u < RL; u++) {
      const double *a = tab + (1u << (lvl0 + S + u)) + (blk << (S + u)) + (b << u);
#pragma unroll
      for (int k = 0; k < (1 << u); k += 4) asm volatile("prefetch.global.L1 [%0];" ::"l"(a + k));
    }
  }
}

// Levels [S, S+RL) of the local block, in registers, on 2^RL elements spaced g = n >> (S+RL) apart (index conventions and
// the padded-offset identity as in ntt_pass, ntt.cuh).  `first` holds the twiddles of this thread's first item.
// IN_GLOBAL (first pass only): the elements come straight from global memory through io.load(index) instead of shared
// memory -- element o + k*g of thread o is a coalesced read for every k.  OUT_GLOBAL (last pass only, g = 1): the 2^RL
// consecutive results of an item go straight to global memory through io.store(index, values).  Both remove a shared-memory
// round trip and a barrier, and let the global latency of one warp hide behind the butterflies of the others.
template <int LOGN, int RL, int S, bool IN_GLOBAL, bool OUT_GLOBAL, class Io>
__device__ __forceinline__ void ntt_pass_f64(double *sm, const double *tab, double p, double pinv, uint32_t lvl0, uint32_t blk,
                                             const PassTwF<LOGN, RL, S> &first, const Io &io) {
  constexpr uint32_t n = 1u << LOGN;
  constexpr int R = 1 << RL;
  constexpr uint32_t g = n >> (S + RL);
  constexpr uint32_t items = n >> RL;
  static_assert(!IN_GLOBAL || S == 0, "global input is the first pass");
  static_assert(!OUT_GLOBAL || g == 1, "global output is the last pass");
  for (uint32_t item = threadIdx.x; item < items; item += blockDim.x) {
    PassTwF<LOGN, RL, S> tw;
    if (item == threadIdx.x) tw = first;
    else tw.load(tab, lvl0, blk, item);
    const uint32_t o = item & (g - 1);
    const uint32_t b = item / g;
    const uint32_t base = b * (n >> S) + o;
    double *ptr = sm + pad_idx(base);
    double v[R];
    if (IN_GLOBAL) {
      typename Io::Raw raw[R];
#pragma unroll
      for (int k = 0; k < R; k++) raw[k] = io.load_raw(base + k * g);
#pragma unroll
      for (int k = 0; k < R; k++) v[k] = io.lift(raw[k]);
    } else {
#pragma unroll
      for (int k = 0; k < R; k++) v[k] = ptr[k * g + ((k * g) >> 4)];
    }
#pragma unroll
    for (int u = 0; u < RL; u++) {
      const int half = R >> (u + 1);
#pragma unroll
      for (int grp = 0; grp < (1 << u); grp++) {
        const double w = tw.w[(1 << u) - 1 + grp], wp = __dmul_rn(w, pinv);
#pragma unroll
        for (int k = 0; k < half; k++) bfly_fwd_f64(v[grp * 2 * half + k], v[grp * 2 * half + k + half], w, wp, p);
      }
    }
    if (OUT_GLOBAL) {
      io.template store<R>(base, v);
    } else {
#pragma unroll
      for (int k = 0; k < R; k++) ptr[k * g + ((k * g) >> 4)] = recentre_f64(v[k], p, pinv);
    }
  }
}

// pass sizes: the small pass LAST (14 -> 4,4,4,2), the shape whose padded layout is bank-conflict free for every pass
__host__ __device__ constexpr int f64_chain_rl(int remain) { return remain >= 4 ? 4 : remain; }

// CRL = levels per pass.  4 (radix-16 passes, 512 threads x 126 registers) is what ships; CRL = 3 with a 1024-thread CTA at
// 64 registers (twice the warps, one more shared-memory round trip) measured 5.67 ms against 5.02 ms per C4 proof on B200.
template <int LOGN, int S, bool IN_GLOBAL, int CRL = 4>
struct PassChainF {
  static constexpr int RL = (LOGN - S) >= CRL ? CRL : (LOGN - S);
  using Tw = PassTwF<LOGN, RL, S>;
  // precondition: `tw` loaded for item threadIdx.x and (unless the pass reads global memory) a barrier passed since shared
  // memory was last written.  The last pass writes global memory through io.
  template <class Io>
  static __device__ __forceinline__ void fwd(double *sm, const double *tab, double p, double pinv, uint32_t lvl0, uint32_t blk,
                                             const Tw &tw, const Io &io) {
    constexpr bool LAST = S + RL >= LOGN;
    if constexpr (!LAST && S + RL >= 7) prefetch_pass_tw<LOGN, PassChainF<LOGN, S + RL, false, CRL>::RL, S + RL>(tab, lvl0, blk);
    ntt_pass_f64<LOGN, RL, S, IN_GLOBAL && S == 0, LAST>(sm, tab, p, pinv, lvl0, blk, tw, io);
    if constexpr (!LAST) {
      typename PassChainF<LOGN, S + RL, false, CRL>::Tw next;
      next.load(tab, lvl0, blk, threadIdx.x);
      __syncthreads();
      PassChainF<LOGN, S + RL, false, CRL>::fwd(sm, tab, p, pinv, lvl0, blk, next, io);
    }
  }
};

}  // namespace rsg
