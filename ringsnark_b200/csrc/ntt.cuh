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

// One pass = RL consecutive levels [s, s+RL) done in registers on 2^RL elements spaced g = n >> (s+RL) apart.
// n = local transform size (1 << LOGN); `lvl0` = levels already applied outside (0 unless the polynomial was
// pre-split in global memory), `blk` = index of this local block among the 1 << lvl0 blocks.
template <int LOGN, int RL, bool INVERSE>
__device__ __forceinline__ void ntt_pass(uint64_t *sm, int s, const Twiddle *tab, uint64_t p, uint32_t lvl0,
                                         uint32_t blk) {
  constexpr uint32_t n = 1u << LOGN;
  constexpr int R = 1 << RL;
  const uint64_t two_p = p << 1;
  const uint32_t g = n >> (s + RL);                 // element stride inside an item
  const uint32_t items = n >> RL;
  for (uint32_t item = threadIdx.x; item < items; item += blockDim.x) {
    const uint32_t o = item & (g - 1);
    const uint32_t b = item / g;                    // block index at level s (g is a power of two)
    const uint32_t base = b * (n >> s) + o;
    uint64_t v[R];
#pragma unroll
    for (int k = 0; k < R; k++) v[k] = sm[pad_idx(base + k * g)];
    if (!INVERSE) {
#pragma unroll
      for (int u = 0; u < RL; u++) {
        const int half = R >> (u + 1);
        const uint32_t tbase = (1u << (lvl0 + s + u)) + (blk << (s + u)) + (b << u);
#pragma unroll
        for (int grp = 0; grp < (1 << u); grp++) {
          const Twiddle t = load_tw(tab, tbase + grp);
#pragma unroll
          for (int k = 0; k < half; k++) bfly_fwd(v[grp * 2 * half + k], v[grp * 2 * half + k + half], t, p, two_p);
        }
      }
    } else {
#pragma unroll
      for (int u = RL - 1; u >= 0; u--) {
        const int half = R >> (u + 1);
        const uint32_t tbase = (1u << (lvl0 + s + u)) + (blk << (s + u)) + (b << u);
#pragma unroll
        for (int grp = 0; grp < (1 << u); grp++) {
          const Twiddle t = load_tw(tab, tbase + grp);
#pragma unroll
          for (int k = 0; k < half; k++) bfly_inv(v[grp * 2 * half + k], v[grp * 2 * half + k + half], t, p, two_p);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < R; k++) sm[pad_idx(base + k * g)] = v[k];
  }
}

// All LOGN levels of the local block, forward.  Input in [0, 4p) (canonical is fine), output lazy in [0, 4p).
template <int LOGN>
__device__ __forceinline__ void ntt_forward_smem(uint64_t *sm, const Twiddle *tab, uint64_t p, uint32_t lvl0,
                                                 uint32_t blk) {
  constexpr int FULL = LOGN / 4, REM = LOGN % 4;
  int s = 0;
#pragma unroll
  for (int i = 0; i < FULL; i++) {
    ntt_pass<LOGN, 4, false>(sm, s, tab, p, lvl0, blk);
    s += 4;
    __syncthreads();
  }
  if (REM == 3) ntt_pass<LOGN, 3, false>(sm, s, tab, p, lvl0, blk);
  if (REM == 2) ntt_pass<LOGN, 2, false>(sm, s, tab, p, lvl0, blk);
  if (REM == 1) ntt_pass<LOGN, 1, false>(sm, s, tab, p, lvl0, blk);
  if (REM) __syncthreads();
}

// All LOGN levels, inverse (levels run LOGN-1 .. 0).  Input in [0, 2p), output lazy in [0, 2p), NOT yet scaled.
template <int LOGN>
__device__ __forceinline__ void ntt_inverse_smem(uint64_t *sm, const Twiddle *tab, uint64_t p, uint32_t lvl0,
                                                 uint32_t blk) {
  constexpr int FULL = LOGN / 4, REM = LOGN % 4;
  int s = LOGN;
  if (REM == 3) ntt_pass<LOGN, 3, true>(sm, s - 3, tab, p, lvl0, blk);
  if (REM == 2) ntt_pass<LOGN, 2, true>(sm, s - 2, tab, p, lvl0, blk);
  if (REM == 1) ntt_pass<LOGN, 1, true>(sm, s - 1, tab, p, lvl0, blk);
  if (REM) __syncthreads();
  s -= REM;
#pragma unroll
  for (int i = 0; i < FULL; i++) {
    s -= 4;
    ntt_pass<LOGN, 4, true>(sm, s, tab, p, lvl0, blk);
    __syncthreads();
  }
}

}  // namespace rsg
