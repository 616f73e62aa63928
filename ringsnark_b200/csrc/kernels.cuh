// CUDA kernels of the prover hot path (sm_100a).  See DESIGN.md for the roofline of each.
//   k_encode_intt     BatchEncoder::encode            batchencoder.cpp:64-88,110-149 + util/ntt.cpp:452-474
//   k_lift_fwd_ntt    transform_to_ntt_inplace        evaluator.cpp:2174-2265 + util/ntt.cpp:407-436
//   k_crs_lincomb     multiply_plain_ntt + add_inplace evaluator.cpp:2135-2172,155-240 fused over all terms
//   k_enc_sum         add_inplace over partial sums   evaluator.cpp:217-231
//   k_is_zero_prefix  SealPoly::is_zero (with its bug) poly_arith.cpp:147-153
//   k_ntt             raw forward / inverse NTT        util/ntt.cpp:407-474
#pragma once
#include <type_traits>
#include "ntt.cuh"
#include "ntt_f64.cuh"

namespace rsg {

constexpr int MAX_LR = 8;
constexpr int MAX_LE = 16;

// Read-only parameter block living in device global memory (one per context).
struct DevParams {
  uint32_t N_R, L_R, N_E, L_E, logN_E;
  ModConst q[MAX_LR];
  ModConst Q[MAX_LE];
  const Twiddle *fwdQ[MAX_LE];   // forward tables mod Q_l
  const Twiddle *invQ[MAX_LE];   // inverse tables mod Q_l (raw NTT entry point only)
  const Twiddle *fwdq[MAX_LR];   // forward tables mod q_j (raw NTT entry point only)
  const Twiddle *invq[MAX_LR];   // inverse tables mod q_j (batch encoder)
  Twiddle invN_q[MAX_LR];        // N_E^-1 mod q_j
  Twiddle invN_Q[MAX_LE];        // N_E^-1 mod Q_l
  Twiddle invNw_q[MAX_LR];       // N_E^-1 * psi^-(N/2) mod q_j: last inverse stage of the split transform (N_E = 2^15)
  Twiddle invNw_Q[MAX_LE];
  uint64_t thr[MAX_LR];          // ceil(q_j / 2): plain_upper_half_threshold (context.cpp:329)
  uint64_t tmodQ[MAX_LR][MAX_LE];  // q_j mod Q_l
  const uint32_t *index_map;     // matrix_reps_index_map_ (batchencoder.cpp:64-88), N_E entries
  const double *fwdQ_f64[MAX_LE];    // forward tables mod Q_l as doubles (null unless every Q_l < 2^49; ntt_f64.cuh)
  double Qinv_f64[MAX_LE];       // RN(1 / Q_l)
};

__device__ __forceinline__ uint64_t canon4(uint64_t x, uint64_t p) {  // [0,4p) -> [0,p)
  const uint64_t two_p = p << 1;
  x = x >= two_p ? x - two_p : x;
  return x >= p ? x - p : x;
}

// ---------------------------------------------------------------------------------------------------------
// Batch encode: ring limb (N_R slot values) -> plaintext polynomial coefficients mod t = q_j (N_E words).
// grid (count << LVL0, L_R); elem_idx (nullable) selects which ring element each block encodes.
// N_E = 2^(LOGN + LVL0).  A 2^15-point polynomial does not fit one SM's shared memory (256 KiB + padding > 227 KiB), so for
// LVL0 = 1 each CTA owns one HALF: the inverse transform's levels LOGN..1 stay inside a half (they are the independent
// sub-transforms of the bit-reversed input), the CTA leaves its lazy values in `plain`, and k_intt_finish applies the
// last level (pairs i, i + N/2) together with the N^-1 scaling SEAL merges into it (util/dwthandler.h:60-73).
// encode_body: the work of one CTA.  CENTRE: store the centred representative v - t for v >= ceil(t/2) (two's complement
// int64) instead of the residue -- what the centred lift of evaluator.cpp:2220-2259 consumes (prover_fast.cuh).
template <int LOGN, int LVL0, bool CENTRE>
__device__ __forceinline__ void encode_body(const DevParams *__restrict__ P, const uint64_t *__restrict__ src, uint64_t *__restrict__ dst,
                                            uint32_t j, uint32_t h) {
  extern __shared__ uint64_t sm[];
  constexpr uint32_t n = 1u << LOGN;
  const uint32_t N_R = P->N_R;
  const uint64_t p = P->q[j].p;
  for (uint32_t i = threadIdx.x; i < padded_words(n); i += blockDim.x) sm[i] = 0;
  __syncthreads();
  const uint32_t *map = P->index_map;
  for (uint32_t k = threadIdx.x; k < N_R; k += blockDim.x) {
    const uint32_t pos = __ldg(map + k);
    if ((pos >> LOGN) == h) sm[pad_idx(pos & (n - 1))] = src[k];
  }
  __syncthreads();
  // unsplit transform: values may stay in [0, 4p) (the exact Shoup multiplication by N^-1 below accepts any 64-bit word);
  // the split one hands [0, 2p) values to k_intt_finish
  if (LVL0 == 0 && 2 * N_R <= n) ntt_inverse_smem_q02<LOGN, true>(sm, P->invq[j], p);   // first matrix row only: half the input is zero
  else ntt_inverse_smem<LOGN, LVL0 == 0>(sm, P->invq[j], p, LVL0, h);
  if (LVL0 == 0) {
    const Twiddle invn = P->invN_q[j];
    const uint64_t thr = P->thr[j];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      uint64_t v = mul_shoup(sm[pad_idx(i)], invn, p);
      if (CENTRE) v = v >= thr ? v - p : v;
      dst[i] = v;
    }
  } else {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = sm[pad_idx(i)];   // lazy, < 2p
  }
}

template <int LOGN, int LVL0>
__global__ void __launch_bounds__(512) k_encode_intt(const DevParams *__restrict__ P, const uint64_t *__restrict__ ring,
                                                     const uint32_t *__restrict__ elem_idx,
                                                     uint64_t *__restrict__ plain) {
  const uint32_t j = blockIdx.y, e = blockIdx.x >> LVL0, h = blockIdx.x & ((1u << LVL0) - 1);
  const uint32_t N_R = P->N_R, L_R = P->L_R;
  const uint32_t src_e = elem_idx ? elem_idx[e] : e;
  const uint64_t *src = ring + ((size_t)src_e * L_R + j) * N_R;
  uint64_t *dst = plain + (((size_t)e * L_R + j) << (LOGN + LVL0)) + ((size_t)h << LOGN);
  encode_body<LOGN, LVL0, false>(P, src, dst, j, h);
}

// ---- Batch encode at N_E = 2^14 when the slot vector uses the first matrix row only (N_R <= N_E / 2): compact rows ----------
// The index map leaves quarters 1 and 3 of the transform's input empty (ntt.cuh, ntt_inverse_smem_q02), i.e. rows 4-7 and 12-15 of
// the 16 x 1024 matrix: only the EIGHT non-zero rows are kept in shared memory (68 KiB instead of 136: two 256-thread CTAs per SM,
// so one polynomial's 128 KiB of stores drain under the butterflies of another -- the lesson of k_lift_fwd_ntt_f64_cl).  The inverse
// transform runs small strides first: levels 13..4 stay inside a row, one WARP per row with __syncwarp only; one CTA barrier; the
// last pass (levels 3..0, column-wise) reads the eight rows, produces all sixteen output rows in registers -- levels 3, 2 on the
// non-zero groups only, level 1 as products (x, 0) -> (x, x w), level 0 with N^-1 folded into its two multiplications, the way
// SEAL merges the scaling into its last stage (util/dwthandler.h:60-73) -- centres them and writes them straight to global memory
// (no scaling pass over shared memory).  Same residues as encode_body.
constexpr uint32_t ENC_ROWW = 1024 + 64;
constexpr size_t ENC_ROWS_SMEM = 8 * (size_t)ENC_ROWW * 8;

template <int RL, int S>
__device__ __forceinline__ void enc_row_pass(uint64_t *rp, const Twiddle *__restrict__ tab, uint64_t p, uint32_t row16, uint32_t lane) {
  constexpr int LS = S - 4, R = 1 << RL;            // the row is the block of 1024 at level 4
  constexpr uint32_t g = 1024u >> (LS + RL), items = 1024u >> RL;
  const uint64_t four_p = p << 2;
  for (uint32_t item = lane; item < items; item += 32) {
    const uint32_t o = item & (g - 1), lb = item / g;
    const uint32_t b = (row16 << LS) + lb;           // block index at level S
    uint64_t *ptr = rp + pad_idx(lb * (1024u >> LS) + o);
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
        for (int k = 0; k < half; k++) bfly_inv_lazy(v[grp * 2 * half + k], v[grp * 2 * half + k + half], t, p, four_p);
      }
    }
#pragma unroll
    for (int k = 0; k < R; k++) ptr[k * g + ((k * g) >> 4)] = v[k];
  }
}

template <bool CENTRE>
__device__ __forceinline__ void encode_rows_body(const DevParams *__restrict__ P, const uint64_t *__restrict__ src, uint64_t *__restrict__ dst,
                                                 uint32_t j) {
  extern __shared__ uint64_t sm[];
  const uint32_t N_R = P->N_R, tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const uint64_t p = P->q[j].p, four_p = p << 2;
  const Twiddle *tab = P->invq[j];
  for (uint32_t i = tid; i < 8 * ENC_ROWW / 2; i += blockDim.x) reinterpret_cast<ulonglong2 *>(sm)[i] = make_ulonglong2(0, 0);
  __syncthreads();
  const uint32_t *map = P->index_map;
  for (uint32_t k = tid; k < N_R; k += blockDim.x) {
    const uint32_t pos = __ldg(map + k), row16 = pos >> 10;              // rows 0-3 and 8-11 only
    sm[((row16 & 3) | ((row16 >> 3) << 2)) * ENC_ROWW + pad_idx(pos & 1023)] = src[k];
  }
  __syncthreads();
  {   // levels 13..4 inside the rows: warp `wrp` owns compact row `wrp`
    const uint32_t row16 = (wrp & 3) | ((wrp >> 2) << 3);
    uint64_t *rp = sm + wrp * ENC_ROWW;
    enc_row_pass<2, 12>(rp, tab, p, row16, lane);
    __syncwarp();
    enc_row_pass<4, 8>(rp, tab, p, row16, lane);
    __syncwarp();
    enc_row_pass<4, 4>(rp, tab, p, row16, lane);
  }
  __syncthreads();
  // levels 3..0 across the rows: v[k] = row k at this column; rows 4-7 and 12-15 are zero on entry
  Twiddle t3[4], t2[2], t1[2];
  t3[0] = load_tw(tab, 8 + 0); t3[1] = load_tw(tab, 8 + 1); t3[2] = load_tw(tab, 8 + 4); t3[3] = load_tw(tab, 8 + 5);
  t2[0] = load_tw(tab, 4 + 0); t2[1] = load_tw(tab, 4 + 2);
  t1[0] = load_tw(tab, 2 + 0); t1[1] = load_tw(tab, 2 + 1);
  const Twiddle invn = P->invN_q[j], invnw = P->invNw_q[j];
  const uint64_t thr = P->thr[j];
  for (uint32_t col = tid; col < 1024; col += blockDim.x) {
    const uint64_t *cp = sm + pad_idx(col);
    uint64_t v[16];
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll
      for (int k = 0; k < 4; k++) v[8 * h + k] = cp[(4 * h + k) * ENC_ROWW];
      bfly_inv_lazy(v[8 * h + 0], v[8 * h + 1], t3[2 * h + 0], p, four_p);     // level 3: rows (0,1), (2,3) | (8,9), (10,11)
      bfly_inv_lazy(v[8 * h + 2], v[8 * h + 3], t3[2 * h + 1], p, four_p);
      bfly_inv_lazy(v[8 * h + 0], v[8 * h + 2], t2[h], p, four_p);             // level 2: rows (0,2), (1,3) | (8,10), (9,11)
      bfly_inv_lazy(v[8 * h + 1], v[8 * h + 3], t2[h], p, four_p);
#pragma unroll
      for (int k = 0; k < 4; k++) v[8 * h + 4 + k] = mul_shoup_approx(v[8 * h + k], t1[h], p);   // level 1: (x, 0) -> (x, x w)
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {   // level 0 with N^-1 folded in: exact Shoup products of any 64-bit word, canonical
      const uint64_t x = v[k], y = v[k + 8];
      uint64_t a = mul_shoup(x + y, invn, p), b = mul_shoup(x - y + four_p, invnw, p);
      if (CENTRE) {
        a = a >= thr ? a - p : a;
        b = b >= thr ? b - p : b;
      }
      dst[col + 1024 * k] = a;
      dst[col + 1024 * (k + 8)] = b;
    }
  }
}

__global__ void __launch_bounds__(256, 2) k_encode_intt_rows(const DevParams *__restrict__ P, const uint64_t *__restrict__ ring,
                                                             const uint32_t *__restrict__ elem_idx, uint64_t *__restrict__ plain) {
  const uint32_t j = blockIdx.y, e = blockIdx.x;
  const uint32_t N_R = P->N_R, L_R = P->L_R;
  const uint32_t src_e = elem_idx ? elem_idx[e] : e;
  encode_rows_body<false>(P, ring + ((size_t)src_e * L_R + j) * N_R, plain + (((size_t)e * L_R + j) << 14), j);
}

// Last Gentleman-Sande level of a split inverse transform + scaling: (x, y) -> ((x + y) N^-1, (x - y) w N^-1), canonical.
// data: `polys` polynomials of 2*half words, values < 2p.  which_q: per-polynomial modulus index = poly % n_mod.
__global__ void __launch_bounds__(256) k_intt_finish(uint64_t *__restrict__ data, uint32_t half, size_t polys,
                                                     const ModConst *__restrict__ mods, const Twiddle *__restrict__ invn,
                                                     const Twiddle *__restrict__ invnw, uint32_t n_mod, uint32_t fixed_mod,
                                                     uint32_t centre = 0) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= polys * half) return;
  const size_t poly = t / half;
  const uint32_t i = (uint32_t)(t - poly * half);
  const uint32_t mi = fixed_mod != 0xFFFFFFFFu ? fixed_mod : (uint32_t)(poly % n_mod);
  const uint64_t p = mods[mi].p, two_p = p << 1;
  uint64_t *d = data + poly * 2 * half;
  const uint64_t x = d[i], y = d[i + half];
  uint64_t s2 = x + y;
  s2 = s2 >= two_p ? s2 - two_p : s2;
  uint64_t a = mul_shoup(s2, invn[mi], p), b = mul_shoup(x - y + two_p, invnw[mi], p);
  if (centre) {   // centred representatives (two's complement), threshold ceil(p / 2) as context.cpp:329
    const uint64_t thr = (p + 1) >> 1;
    a = a >= thr ? a - p : a;
    b = b >= thr ? b - p : b;
  }
  d[i] = a;
  d[i + half] = b;
}

// ---------------------------------------------------------------------------------------------------------
// Centred lift into Q_l + forward NTT.  grid ((count * L_E) << LVL0, L_R): the L_E limbs of one plaintext are neighbours in
// launch order, so the CTAs that read the same coefficient vector run together and all but one of the reads hit L2.
// LAZY (every Q_l < 2^58): correction-free butterflies, one Barrett reduction per word at the store.
// LVL0 = 1 (N_E = 2^15): the first Cooley-Tukey level pairs i with i + N/2 and is applied while loading (each of the two
// CTAs lifts both words and keeps its own half); the remaining levels are two independent 2^14-point transforms.
// is_signed: `plain` holds the SUM of two centred plaintexts as int64 words (k_centre_add) instead of residues mod t --
// NTT_Q is linear, so two inner products over the same CRS range share one transform and one CRS pass.
template <int LOGN, int LVL0, bool LAZY>
__global__ void __launch_bounds__(512) k_lift_fwd_ntt(const DevParams *__restrict__ P, const uint64_t *__restrict__ plain,
                                                      uint32_t is_signed, uint64_t *__restrict__ out,
                                                      const uint8_t *__restrict__ slot_skip = nullptr) {
  extern __shared__ uint64_t sm[];
  constexpr uint32_t n = 1u << LOGN;
  const uint32_t L_R = P->L_R, L_E = P->L_E;
  const uint32_t h = blockIdx.x & ((1u << LVL0) - 1), el = blockIdx.x >> LVL0;
  const uint32_t e = el / L_E, l = el - e * L_E, j = blockIdx.y;
  if (slot_skip && slot_skip[e]) return;   // term skipped on the device (prover_fast.cuh): nothing reads this slot
  const ModConst m = P->Q[l];
  const uint64_t thr = P->thr[j], tm = P->tmodQ[j][l];
  const uint64_t *src = plain + (((size_t)e * L_R + j) << (LOGN + LVL0));
  auto lift = [&](uint32_t i) {
    const uint64_t v = src[i];
    if (is_signed) {
      const long long sv = (long long)v;
      const uint64_t r = reduce64((uint64_t)(sv < 0 ? -sv : sv), m);
      return sv < 0 ? neg_mod(r, m.p) : r;
    }
    uint64_t r = reduce64(v, m);
    if (v >= thr) r = sub_mod(r, tm, m.p);   // v + (Q - t)  ==  v - t  (mod Q_l)
    return r;
  };
  if (LVL0 == 0) {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) sm[pad_idx(i)] = lift(i);
  } else {
    const Twiddle t0 = load_tw(P->fwdQ[l], 1);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      uint64_t x = lift(i), y = lift(i + n);
      if (LAZY) bfly_fwd_lazy(x, y, t0, m.p, m.p << 2);
      else bfly_fwd(x, y, t0, m.p, m.p << 1);
      sm[pad_idx(i)] = h ? y : x;
    }
  }
  __syncthreads();
  ntt_forward_smem<LOGN, LAZY>(sm, P->fwdQ[l], m.p, LVL0, h);
  uint64_t *dst = out + (((((size_t)e * L_R + j) * L_E + l)) << (LOGN + LVL0)) + (size_t)h * n;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
    dst[i] = LAZY ? reduce64(sm[pad_idx(i)], m) : canon4(sm[pad_idx(i)], m.p);
}

// The same kernel on the FP64 pipe (every Q_l < 2^49; ntt_f64.cuh): the lift stays integer (one Barrett reduction per word),
// the residue is centred and converted once, all LOGN levels run on exact doubles, and the store canonicalises.  For
// LVL0 = 0 the first pass reads the coefficients straight from global memory and the last pass writes the result straight
// to global memory (no separate load / store phases).
template <bool SIGNED>
struct LiftIoF64 {
  using Raw = uint64_t;
  const uint64_t *src;
  uint64_t *dst;
  ModConst m;
  uint64_t thr, tm;
  double pd, pinv;
  __device__ __forceinline__ void init(const DevParams *__restrict__ P, uint32_t j, uint32_t l) {
    m = P->Q[l];
    pd = (double)m.p;
    pinv = P->Qinv_f64[l];
    thr = P->thr[j];
    tm = P->tmodQ[j][l];
  }
  __device__ __forceinline__ Raw load_raw(uint32_t i) const { return __ldg(src + i); }
  __device__ __forceinline__ double lift(Raw v) const {
    if constexpr (SIGNED) {   // sum of two centred plaintexts (k_centre_add), |v| < 2^61
      const long long sv = (long long)v;
      const uint64_t r = reduce64((uint64_t)(sv < 0 ? -sv : sv), m);
      return centre_to_f64(sv < 0 ? neg_mod(r, m.p) : r, m.p);
    } else {
      uint64_t r = reduce64(v, m);
      if (v >= thr) r = sub_mod(r, tm, m.p);   // v + (Q - t)  ==  v - t  (mod Q_l)
      return centre_to_f64(r, m.p);
    }
  }
  __device__ __forceinline__ uint64_t canon(double x) const { return canon_f64(x, pd, pinv, m.p); }
  template <int R>
  __device__ __forceinline__ void store(uint32_t base, const double (&v)[R]) const {
    static_assert(R % 2 == 0, "pairs");
#pragma unroll
    for (int k = 0; k < R; k += 2)
      *reinterpret_cast<ulonglong2 *>(dst + base + k) = make_ulonglong2(canon_f64(v[k], pd, pinv, m.p), canon_f64(v[k + 1], pd, pinv, m.p));
  }
};

// The same lift with a SMALL QUOTIENT (context flag lift_smallq: t < 2^54 and t / min Q_l < 2^11, true for every reference
// configuration): |v| < 2^55 fits 31 bits after a shift by 24, one float multiply estimates v / Q_l to within 1/2 + 2^-12, and
// x = v - q Q_l is one 32 x 64 product: |x| <= (1/2 + 2^-12) Q_l, the same residue class as the Barrett lift above and, after the
// transform, the same canonical words.  About twenty instructions per coefficient instead of about forty; measured on B200
// (tools/ntt_lab.cu, C4): 5.05 -> 4.43 ms per proof -- the kernel is bound by instruction issue (FP64 instructions hold the
// issue port two cycles, every other instruction about one: tools/fp64_lab.cu), not by the FP64 pipe alone.
// Integer <-> double conversions add the bit pattern of 1.5 * 2^52 (exact for |x| < 2^51) instead of using the XU pipe.
constexpr long long F64_MAGIC_BITS = 0x4338000000000000LL;   // bit pattern of F64_MAGIC
__device__ __forceinline__ double ll2double_magic(long long x) { return __dadd_rn(__longlong_as_double(x + F64_MAGIC_BITS), -F64_MAGIC); }
__device__ __forceinline__ long long double2ll_magic(double x) { return __double_as_longlong(__dadd_rn(x, F64_MAGIC)) - F64_MAGIC_BITS; }

template <bool SIGNED>
struct LiftIoSmallQ {
  using Raw = uint64_t;
  const uint64_t *src;
  uint64_t *dst;
  uint64_t p, thr, t;
  float qinv24;        // 2^24 / Q_l
  double pd, pinv;
  __device__ __forceinline__ void init(const DevParams *__restrict__ P, uint32_t j, uint32_t l) {
    p = P->Q[l].p;
    pd = (double)p;
    pinv = P->Qinv_f64[l];
    qinv24 = (float)(16777216.0 * pinv);
    thr = P->thr[j];
    t = P->q[j].p;
  }
  __device__ __forceinline__ Raw load_raw(uint32_t i) const { return __ldg(src + i); }
  __device__ __forceinline__ double lift(Raw raw) const {
    long long v = (long long)raw;
    if constexpr (!SIGNED) v = raw >= thr ? (long long)(raw - t) : (long long)raw;   // context.cpp:329, evaluator.cpp:2243-2259
    const int vh = (int)(v >> 24);
    const float qf = __fadd_rn(__fmul_rn(__int2float_rn(vh), qinv24), 12582912.0f);   // 1.5 * 2^23: mantissa = round(quotient)
    const int q = __float_as_int(qf) - 0x4B400000;
    return ll2double_magic(v - (long long)q * (long long)p);
  }
  __device__ __forceinline__ uint64_t canon(double x) const {
    const long long r = double2ll_magic(recentre_f64(x, pd, pinv));
    return (uint64_t)(r + ((r >> 63) & (long long)p));
  }
  template <int R>
  __device__ __forceinline__ void store(uint32_t base, const double (&v)[R]) const {
    static_assert(R % 2 == 0, "pairs");
#pragma unroll
    for (int k = 0; k < R; k += 2) *reinterpret_cast<ulonglong2 *>(dst + base + k) = make_ulonglong2(canon(v[k]), canon(v[k + 1]));
  }
};

template <int LOGN, int LVL0, bool SIGNED, bool SMALLQ>
__device__ __forceinline__ void lift_fwd_ntt_f64_body(const DevParams *__restrict__ P, const uint64_t *__restrict__ plain,
                                                      uint64_t *__restrict__ out, const uint8_t *__restrict__ slot_skip) {
  extern __shared__ double smf[];
  constexpr uint32_t n = 1u << LOGN;
  const uint32_t L_R = P->L_R, L_E = P->L_E;
  const uint32_t h = blockIdx.x & ((1u << LVL0) - 1), el = blockIdx.x >> LVL0;
  const uint32_t e = el / L_E, l = el - e * L_E, j = blockIdx.y;
  if (slot_skip && slot_skip[e]) return;   // term skipped on the device (prover_fast.cuh): nothing reads this slot
  std::conditional_t<SMALLQ, LiftIoSmallQ<SIGNED>, LiftIoF64<SIGNED>> io;
  io.init(P, j, l);
  io.src = plain + (((size_t)e * L_R + j) << (LOGN + LVL0));
  io.dst = out + (((((size_t)e * L_R + j) * L_E + l)) << (LOGN + LVL0)) + (size_t)h * n;
  const double *tab = P->fwdQ_f64[l];
  typename PassChainF<LOGN, 0, LVL0 == 0>::Tw tw0;
  tw0.load(tab, LVL0, h, threadIdx.x);
  if (LVL0 != 0) {
    const double w0 = __ldg(tab + 1), w0p = __dmul_rn(w0, io.pinv);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      double x = io.lift(io.load_raw(i)), y = io.lift(io.load_raw(i + n));
      bfly_fwd_f64(x, y, w0, w0p, io.pd);
      smf[pad_idx(i)] = recentre_f64(h ? y : x, io.pd, io.pinv);
    }
    __syncthreads();
  }
  PassChainF<LOGN, 0, LVL0 == 0>::fwd(smf, tab, io.pd, io.pinv, LVL0, h, tw0, io);
}
template <int LOGN, int LVL0, bool SIGNED, bool SMALLQ = false>
__global__ void __launch_bounds__(512) k_lift_fwd_ntt_f64(const DevParams *__restrict__ P, const uint64_t *__restrict__ plain,
                                                          uint64_t *__restrict__ out,
                                                          const uint8_t *__restrict__ slot_skip = nullptr) {
  lift_fwd_ntt_f64_body<LOGN, LVL0, SIGNED, SMALLQ>(P, plain, out, slot_skip);
}
// ---- N_E = 2^14 on thread-block clusters (what ships for the C4 / C3' shape) ------------------------------------------------
// Measured on B200 (tools/ntt_lab.cu): the single-CTA kernel above spends 39 % of its time with the FP64
// pipe idle behind its own global traffic -- 128 KiB of coefficients in before the first butterfly, 128 KiB of results out
// after the last, one CTA per SM (a 2^14-point polynomial fills the shared memory) and therefore nothing to overlap them
// with -- while its radix-16 passes themselves run at 91 % of the pipe's peak.  This kernel removes that:
//   * one polynomial per CLUSTER of four 128-thread CTAs; CTA r keeps rows 4r..4r+3 of the 16 x 1024 matrix (34 KiB), so five
//     CTAs of different clusters share an SM and one polynomial's loads / stores run under another's butterflies;
//   * pass 1 (levels 0-3, column-wise): CTA r transforms columns [256 r, 256 r + 256) and sends each of the 16 rows to its
//     owner through distributed shared memory (st.shared::cluster, 128 bit); one cluster barrier;
//   * passes 2-4 stay inside a row of 1024: one WARP per row, __syncwarp only (no CTA barrier after the first pass);
//   * every radix-16 pass handles two adjacent columns per thread: one LDS.128 / STS.128 moves both, the fifteen twiddles
//     serve 64 butterflies; padding of four words per 64 (pad2) keeps pairs 16-byte aligned and the passes conflict-free;
//   * results leave as one 256-bit store per four words (a whole 32-byte sector per lane: 128-bit stores at 32-byte stride
//     wrote half sectors and cost 0.4 ms per proof).
// Same butterflies in the same order as ntt_pass_f64: identical words (tools/ntt_lab.cu compares them).  5.05 -> 3.46 ms per
// C4 proof for the 37 056 transforms.
__device__ __forceinline__ uint32_t pad2(uint32_t i) { return i + ((i >> 6) << 2); }

__device__ __forceinline__ void radix16_pair(double (&a)[16], double (&b)[16], const double (&w)[15], double p, double pinv) {
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const int half = 16 >> (u + 1);
#pragma unroll
    for (int grp = 0; grp < (1 << u); grp++) {
      const double ww = w[(1 << u) - 1 + grp], wp = __dmul_rn(ww, pinv);
#pragma unroll
      for (int k = 0; k < half; k++) {
        bfly_fwd_f64(a[grp * 2 * half + k], a[grp * 2 * half + k + half], ww, wp, p);
        bfly_fwd_f64(b[grp * 2 * half + k], b[grp * 2 * half + k + half], ww, wp, p);
      }
    }
  }
}
// the fifteen twiddles of levels lvl..lvl+3 for block b of level lvl: entries (1 << (lvl + u)) + (b << u) + grp, grp < 2^u --
// contiguous and 16-byte aligned for u >= 1
__device__ __forceinline__ void ld_tw15(double (&w)[15], const double *tab, uint32_t lvl, uint32_t b) {
  w[0] = ldg_f64_here(tab + (1u << lvl) + b);
#pragma unroll
  for (int u = 1; u < 4; u++) {
    const double *q = tab + (1u << (lvl + u)) + (b << u);
#pragma unroll
    for (int g2 = 0; g2 < (1 << u); g2 += 2)
      asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(w[(1 << u) - 1 + g2]), "=d"(w[(1 << u) + g2]) : "l"(q + g2));
  }
}

constexpr int NTT_CL = 4;                        // CTAs per cluster
constexpr int NTT_CL_THREADS = 512 / NTT_CL;     // 128
constexpr uint32_t NTT_CL_ROWW = 1024 + 64;      // one padded row
constexpr size_t NTT_CL_SMEM = (size_t)(16 / NTT_CL) * NTT_CL_ROWW * 8;

template <bool SIGNED, bool SMALLQ>
__global__ void __cluster_dims__(NTT_CL, 1, 1) __launch_bounds__(NTT_CL_THREADS, 5)
    k_lift_fwd_ntt_f64_cl(const DevParams *__restrict__ P, const uint64_t *__restrict__ plain, uint64_t *__restrict__ out,
                          const uint8_t *__restrict__ slot_skip) {
  constexpr int LOGN = 14;
  constexpr uint32_t RPC = 16 / NTT_CL;   // rows per CTA
  extern __shared__ double smf[];
  uint32_t r;
  asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  const uint32_t L_R = P->L_R, L_E = P->L_E;
  const uint32_t el = blockIdx.x / NTT_CL, e = el / L_E, l = el - e * L_E, j = blockIdx.y;
  if (slot_skip && slot_skip[e]) return;   // all CTAs of the cluster together
  // the peers' shared memory may be written once they run: arrive now, wait just before the first remote store
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  std::conditional_t<SMALLQ, LiftIoSmallQ<SIGNED>, LiftIoF64<SIGNED>> io;
  io.init(P, j, l);
  io.src = plain + (((size_t)e * L_R + j) << LOGN);
  io.dst = out + (((((size_t)e * L_R + j) * L_E + l)) << LOGN);
  const double *tab = P->fwdQ_f64[l];
  const double pd = io.pd, pinv = io.pinv;
  // Q_l < 2^48: the re-centring after passes 1 and 3 is skipped.  With |v| <= (1/2 + |y| 2^-52) p per butterfly and p < 2^48 a
  // bound of b p grows to (b + 1/2 + b/16) p per level: 0.501 -> 2.84 p after four levels, 5.81 p after eight -- below 2^51 = 8 p,
  // the limit on a multiplied operand, and far below 2^53; pass 2 re-centres, passes 3 + 4 end at 4.3 p and the store canonicalises.
  // (49-bit primes: 3.22 p after four levels, the fifth would pass 4 p = 2^51: every pass re-centres.)
  const bool lazy48 = P->Q[l].p < (1ull << 48);
  const uint32_t tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  double a[16], b[16], w[15];
  {   // pass 1: column pair o of all 16 rows
    const uint32_t o = 2 * (NTT_CL_THREADS * r + tid);
    ulonglong2 raw[16];
#pragma unroll
    for (int k = 0; k < 16; k++) raw[k] = __ldg(reinterpret_cast<const ulonglong2 *>(io.src + o + 1024 * k));
    ld_tw15(w, tab, 0, 0);
#pragma unroll
    for (int k = 0; k < 16; k++) { a[k] = io.lift(raw[k].x); b[k] = io.lift(raw[k].y); }
    radix16_pair(a, b, w, pd, pinv);
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(smf) + pad2(o) * 8;
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (!lazy48) {
#pragma unroll
      for (int k = 0; k < 16; k++) { a[k] = recentre_f64(a[k], pd, pinv); b[k] = recentre_f64(b[k], pd, pinv); }
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
      uint32_t dstaddr;
      asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dstaddr) : "r"(local + (uint32_t)(k % RPC) * NTT_CL_ROWW * 8), "r"((uint32_t)(k / RPC)));
      asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(dstaddr), "d"(a[k]), "d"(b[k]) : "memory");
    }
  }
  const uint32_t row = RPC * r + wrp;   // the row (block of 1024) this warp owns from here on
  double *rp = smf + wrp * NTT_CL_ROWW;
  ld_tw15(w, tab, 4, row);
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  {   // pass 2: levels 4..7, elements 64 apart
    double *ptr = rp + 2 * lane;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const double2 t = *reinterpret_cast<const double2 *>(ptr + 68 * k);
      a[k] = t.x; b[k] = t.y;
    }
    radix16_pair(a, b, w, pd, pinv);
#pragma unroll
    for (int k = 0; k < 16; k++) *reinterpret_cast<double2 *>(ptr + 68 * k) = make_double2(recentre_f64(a[k], pd, pinv), recentre_f64(b[k], pd, pinv));
  }
  {   // pass 3: levels 8..11, elements 4 apart inside block bb of 64
    const uint32_t bb = lane >> 1, o = 2 * (lane & 1);
    ld_tw15(w, tab, 8, row * 16 + bb);
    __syncwarp();
    double *ptr = rp + bb * 68 + o;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const double2 t = *reinterpret_cast<const double2 *>(ptr + 4 * k);
      a[k] = t.x; b[k] = t.y;
    }
    radix16_pair(a, b, w, pd, pinv);
    if (!lazy48) {
#pragma unroll
      for (int k = 0; k < 16; k++) { a[k] = recentre_f64(a[k], pd, pinv); b[k] = recentre_f64(b[k], pd, pinv); }
    }
#pragma unroll
    for (int k = 0; k < 16; k++) *reinterpret_cast<double2 *>(ptr + 4 * k) = make_double2(a[k], b[k]);
  }
  {   // pass 4: levels 12, 13 on four consecutive elements; items lane + 32 m of the row's 256
    double w12[8], w13a[8], w13b[8];
#pragma unroll
    for (int m = 0; m < 8; m++) {
      const uint32_t i = row * 256 + lane + 32 * m;
      w12[m] = ldg_f64_here(tab + 4096 + i);
      asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(w13a[m]), "=d"(w13b[m]) : "l"(tab + 8192 + 2 * i));
    }
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 8; m++) {
      const uint32_t il = lane + 32 * m;
      const double *ptr = rp + pad2(4 * il);
      const double2 t0 = *reinterpret_cast<const double2 *>(ptr), t1 = *reinterpret_cast<const double2 *>(ptr + 2);
      double v0 = t0.x, v1 = t0.y, v2 = t1.x, v3 = t1.y;
      const double wa = w12[m], wap = __dmul_rn(wa, pinv);
      bfly_fwd_f64(v0, v2, wa, wap, pd);
      bfly_fwd_f64(v1, v3, wa, wap, pd);
      bfly_fwd_f64(v0, v1, w13a[m], __dmul_rn(w13a[m], pinv), pd);
      bfly_fwd_f64(v2, v3, w13b[m], __dmul_rn(w13b[m], pinv), pd);
      asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(io.dst + 1024 * row + 4 * il), "l"(io.canon(v0)), "l"(io.canon(v1)),
                   "l"(io.canon(v2)), "l"(io.canon(v3))
                   : "memory");
    }
  }
}

// ---- the integer transform on clusters (every Q_l < 2^58; what C5's 55-bit limbs run) ---------------------------------------
// k_lift_fwd_ntt_f64_cl's structure with Shoup butterflies: one polynomial of 2^(10 + LOGR) coefficients = 2^LOGR rows of 1024 per
// cluster of 2^LOGR / 4 CTAs (128 threads, four rows = 34 KiB each, so four CTAs of different clusters share an SM).  Pass 1:
// every thread owns ONE column, loads its 2^LOGR coefficients (coalesced across the warp), lifts them, runs levels 0..LOGR-1 in
// registers (radix-32 at N_E = 2^15) and sends each row's value to the row's owner through distributed shared memory; one cluster
// barrier; levels LOGR.. stay inside a row -- one warp per row, __syncwarp only (radix 16, 16, 4) -- and the last pass writes
// canonical words with 256-bit stores.  Correction-free butterflies (bfly_fwd_lazy: values grow by 4p per level, 61p < 2^64 after 15
// levels), one Barrett reduction per word at the store: the same residues as k_lift_fwd_ntt.
template <int RL, int LS>   // levels [LS, LS + RL) of the row-local transform of 1024; gl0 = global level of row-local level 0
__device__ __forceinline__ void int_row_pass_fwd(uint64_t *rp, const Twiddle *__restrict__ tab, uint64_t p, uint32_t gl0, uint32_t row, uint32_t lane) {
  constexpr int R = 1 << RL;
  constexpr uint32_t g = 1024u >> (LS + RL), items = 1024u >> RL;
  const uint64_t four_p = p << 2;
  for (uint32_t item = lane; item < items; item += 32) {
    const uint32_t o = item & (g - 1), lb = item / g;
    const uint32_t b = (row << LS) + lb;   // block index at global level gl0 + LS
    uint64_t *ptr = rp + pad_idx(lb * (1024u >> LS) + o);
    uint64_t v[R];
#pragma unroll
    for (int k = 0; k < R; k++) v[k] = ptr[k * g + ((k * g) >> 4)];
#pragma unroll
    for (int u = 0; u < RL; u++) {
      const int half = R >> (u + 1);
      const uint32_t tbase = (1u << (gl0 + LS + u)) + (b << u);
#pragma unroll
      for (int grp = 0; grp < (1 << u); grp++) {
        const Twiddle t = load_tw(tab, tbase + grp);
#pragma unroll
        for (int k = 0; k < half; k++) bfly_fwd_lazy(v[grp * 2 * half + k], v[grp * 2 * half + k + half], t, p, four_p);
      }
    }
#pragma unroll
    for (int k = 0; k < R; k++) ptr[k * g + ((k * g) >> 4)] = v[k];
  }
}

template <int LOGR, bool SIGNED>
__global__ void __cluster_dims__((1 << LOGR) / 4, 1, 1) __launch_bounds__(128, 4)
    k_lift_fwd_ntt_int_cl(const DevParams *__restrict__ P, const uint64_t *__restrict__ plain, uint64_t *__restrict__ out,
                          const uint8_t *__restrict__ slot_skip) {
  constexpr int ROWS = 1 << LOGR, CS = ROWS / 4, LOGN = 10 + LOGR;
  constexpr uint32_t ROWW = 1024 + 64;
  extern __shared__ uint64_t smi[];
  uint32_t r;
  asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  const uint32_t L_R = P->L_R, L_E = P->L_E;
  const uint32_t el = blockIdx.x / CS, e = el / L_E, l = el - e * L_E, j = blockIdx.y;
  if (slot_skip && slot_skip[e]) return;   // all CTAs of the cluster together
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  const ModConst m = P->Q[l];
  const uint64_t p = m.p, four_p = p << 2;
  const uint64_t thr = P->thr[j], tm = P->tmodQ[j][l];
  const Twiddle *tab = P->fwdQ[l];
  const uint64_t *src = plain + (((size_t)e * L_R + j) << LOGN);
  uint64_t *dst = out + (((((size_t)e * L_R + j) * L_E + l)) << LOGN);
  const uint32_t tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
#pragma unroll 1
  for (uint32_t it = 0; it < 1024 / CS / 128; it++) {   // pass 1: column `col` of all rows, levels 0..LOGR-1 (1024 / CS columns per CTA)
    const uint32_t col = (1024 / CS) * r + 128 * it + tid;
    uint64_t v[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; k++) v[k] = __ldg(src + col + 1024 * k);
#pragma unroll
    for (int k = 0; k < ROWS; k++) {
      if (SIGNED) {   // sum of two centred plaintexts (k_centre_add)
        const long long sv = (long long)v[k];
        const uint64_t x = reduce64((uint64_t)(sv < 0 ? -sv : sv), m);
        v[k] = sv < 0 ? neg_mod(x, p) : x;
      } else {
        uint64_t x = reduce64(v[k], m);
        if (v[k] >= thr) x = sub_mod(x, tm, p);   // v + (Q - t)  ==  v - t  (mod Q_l)
        v[k] = x;
      }
    }
#pragma unroll
    for (int u = 0; u < LOGR; u++) {
      const int half = ROWS >> (u + 1);
#pragma unroll
      for (int grp = 0; grp < (1 << u); grp++) {
        const Twiddle t = load_tw(tab, (1u << u) + grp);
#pragma unroll
        for (int k = 0; k < half; k++) bfly_fwd_lazy(v[grp * 2 * half + k], v[grp * 2 * half + k + half], t, p, four_p);
      }
    }
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(smi) + pad_idx(col) * 8;
    if (it == 0) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // the peers run: their shared memory exists
#pragma unroll
    for (int k = 0; k < ROWS; k++) {
      uint32_t dstaddr;
      asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dstaddr) : "r"(local + (uint32_t)(k & 3) * ROWW * 8), "r"((uint32_t)(k >> 2)));
      asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(dstaddr), "l"(v[k]) : "memory");
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  const uint32_t row = 4 * r + wrp;   // block of 1024 at level LOGR
  uint64_t *rp = smi + wrp * ROWW;
  int_row_pass_fwd<4, 0>(rp, tab, p, LOGR, row, lane);
  __syncwarp();
  int_row_pass_fwd<4, 4>(rp, tab, p, LOGR, row, lane);
  __syncwarp();
  {   // row-local levels 8, 9 on four consecutive elements, straight to global memory
    const uint32_t lvl8 = 1u << (LOGR + 8), lvl9 = 1u << (LOGR + 9);
#pragma unroll 2
    for (uint32_t it = lane; it < 256; it += 32) {
      const uint32_t b = row * 256 + it;
      const uint64_t *ptr = rp + pad_idx(4 * it);
      uint64_t v0 = ptr[0], v1 = ptr[1], v2 = ptr[2], v3 = ptr[3];
      const Twiddle ta = load_tw(tab, lvl8 + b), tb = load_tw(tab, lvl9 + 2 * b), tc = load_tw(tab, lvl9 + 2 * b + 1);
      bfly_fwd_lazy(v0, v2, ta, p, four_p);
      bfly_fwd_lazy(v1, v3, ta, p, four_p);
      bfly_fwd_lazy(v0, v1, tb, p, four_p);
      bfly_fwd_lazy(v2, v3, tc, p, four_p);
      asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(dst + 1024 * row + 4 * it), "l"(reduce64(v0, m)), "l"(reduce64(v1, m)),
                   "l"(reduce64(v2, m)), "l"(reduce64(v3, m))
                   : "memory");
    }
  }
}

// The same transform capped at 96 registers (a few twiddles spill to L1): 512 x 96 = 48 Ki registers leave room for one
// 256-thread CTA of k_crs_lincomb_r64 on the same SM, so the HBM-bound stream of one term group runs UNDER the FP64-bound
// transforms of the next (prover_fast.cuh, RSG_OVERLAP).
template <int LOGN, int LVL0, bool SIGNED, bool SMALLQ = false>
__global__ void __maxnreg__(96) k_lift_fwd_ntt_f64_r96(const DevParams *__restrict__ P, const uint64_t *__restrict__ plain,
                                                       uint64_t *__restrict__ out, const uint8_t *__restrict__ slot_skip) {
  lift_fwd_ntt_f64_body<LOGN, LVL0, SIGNED, SMALLQ>(P, plain, out, slot_skip);
}

// Raw NTT of `batch` polynomials; grid (batch << LVL0).  In place for LVL0 = 0.  For LVL0 = 1 the forward transform reads
// `src` and writes `dst` (they must differ: both CTAs of a polynomial read both halves), the inverse works in place on its
// own half and is completed by k_intt_finish.
template <int LOGN, int LVL0, bool INVERSE>
__global__ void __launch_bounds__(512) k_ntt(const uint64_t *__restrict__ src, uint64_t *__restrict__ dst,
                                             const Twiddle *__restrict__ tab, uint64_t p, Twiddle invn) {
  extern __shared__ uint64_t sm[];
  constexpr uint32_t n = 1u << LOGN;
  const uint32_t h = blockIdx.x & ((1u << LVL0) - 1);
  const size_t poly = (size_t)(blockIdx.x >> LVL0) << (LOGN + LVL0);
  const uint64_t *s = src + poly;
  uint64_t *d = dst + poly + (size_t)h * n;
  if (INVERSE) {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) sm[pad_idx(i)] = s[(size_t)h * n + i];
    __syncthreads();
    ntt_inverse_smem<LOGN>(sm, tab, p, LVL0, h);
    if (LVL0 == 0) {
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) d[i] = mul_shoup(sm[pad_idx(i)], invn, p);
    } else {
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) d[i] = sm[pad_idx(i)];
    }
  } else {
    if (LVL0 == 0) {
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) sm[pad_idx(i)] = s[i];
    } else {
      const Twiddle t0 = load_tw(tab, 1);
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        uint64_t x = s[i], y = s[i + n];
        bfly_fwd(x, y, t0, p, p << 1);
        sm[pad_idx(i)] = h ? y : x;
      }
    }
    __syncthreads();
    ntt_forward_smem<LOGN>(sm, tab, p, LVL0, h);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) d[i] = canon4(sm[pad_idx(i)], p);
  }
}

// ---------------------------------------------------------------------------------------------------------
// The CRS linear combination: one coalesced, 128-bit-vectorised streaming pass over the CRS in HBM.
//   partial[z][j][k][l][x] = sum_{t in split z} crs[term[t]][j][k][l][x] * pntt[pidx[t]][j][l][x]   mod Q_l
// Each thread owns two adjacent x for one (j, l) and both ciphertext polynomials k = 0, 1: per term it loads
// 3 x 16 B and issues 4 multiply-accumulates into 192-bit accumulators (one Barrett reduction per output word per
// launch instead of one per term -- same canonical residue).
// grid (N_E / (2*blockDim), L_R*L_E, splits).
__device__ __forceinline__ ulonglong2 ld_stream(const uint64_t *p) {
  ulonglong2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  return v;
}

template <int UNROLL>
__device__ __forceinline__ void crs_lincomb_body(const DevParams *__restrict__ P, const uint64_t *__restrict__ crs,
                                                 const uint32_t *__restrict__ term, const uint32_t *__restrict__ pidx,
                                                 uint32_t n_terms, uint32_t terms_per_split,
                                                 const uint64_t *__restrict__ pntt, uint64_t *__restrict__ partial,
                                                 const uint32_t *__restrict__ zoff, const uint8_t *__restrict__ slot_skip,
                                                 const uint64_t *const *__restrict__ term_ptr = nullptr) {
  const uint32_t N_E = P->N_E, L_E = P->L_E, L_R = P->L_R;
  const uint32_t x = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const uint32_t j = blockIdx.y / L_E, l = blockIdx.y - j * L_E;
  // zoff (prover_fast.cuh): split z sums the terms [zoff[z], zoff[z+1]) -- several outputs in one launch
  const uint32_t t0 = zoff ? zoff[blockIdx.z] : blockIdx.z * terms_per_split;
  const uint32_t t1 = zoff ? zoff[blockIdx.z + 1] : min(n_terms, t0 + terms_per_split);
  const size_t poly = (size_t)N_E;                       // words per (k, l) row
  const size_t ct_words = 2 * (size_t)L_E * poly;        // one ciphertext
  const size_t enc_words = (size_t)L_R * ct_words;       // one encoding
  const size_t c_off = (size_t)j * ct_words + (size_t)l * poly + x;          // k = 0 row inside an encoding
  const size_t k_stride = (size_t)L_E * poly;
  const size_t p_off = ((size_t)j * L_E + l) * poly + x;
  const size_t p_stride = (size_t)L_R * L_E * poly;

  Acc192 a00, a01, a10, a11;
  a00.clear(); a01.clear(); a10.clear(); a11.clear();

  uint32_t t = t0;
  for (; t + UNROLL <= t1; t += UNROLL) {
    ulonglong2 c0[UNROLL], c1[UNROLL], pp[UNROLL];
    uint32_t pi[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      pi[u] = __ldg(pidx + t + u);
      // term_ptr (prover_fast.cuh): the encodings of a term list may live in different arenas -- one device pointer per term
      const uint64_t *c = (term_ptr ? term_ptr[t + u] : crs + (size_t)__ldg(term + t + u) * enc_words) + c_off;
      if (slot_skip && pi[u] != 0xFFFFFFFFu && slot_skip[pi[u]]) {   // skipped on the device: contributes nothing
        c0[u] = c1[u] = pp[u] = make_ulonglong2(0, 0);
        continue;
      }
      c0[u] = ld_stream(c);
      c1[u] = ld_stream(c + k_stride);
      if (pi[u] != 0xFFFFFFFFu) pp[u] = ld_stream(pntt + (size_t)pi[u] * p_stride + p_off);
      else pp[u] = make_ulonglong2(1, 1);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      a00.mac(c0[u].x, pp[u].x);
      a01.mac(c0[u].y, pp[u].y);
      a10.mac(c1[u].x, pp[u].x);
      a11.mac(c1[u].y, pp[u].y);
    }
  }
  for (; t < t1; t++) {
    const uint32_t pi = __ldg(pidx + t);
    if (slot_skip && pi != 0xFFFFFFFFu && slot_skip[pi]) continue;
    const uint64_t *c = (term_ptr ? term_ptr[t] : crs + (size_t)__ldg(term + t) * enc_words) + c_off;
    const ulonglong2 c0 = ld_stream(c), c1 = ld_stream(c + k_stride);
    const ulonglong2 pp = pi != 0xFFFFFFFFu ? ld_stream(pntt + (size_t)pi * p_stride + p_off) : make_ulonglong2(1, 1);
    a00.mac(c0.x, pp.x);
    a01.mac(c0.y, pp.y);
    a10.mac(c1.x, pp.x);
    a11.mac(c1.y, pp.y);
  }
  const ModConst m = P->Q[l];
  uint64_t *o = partial + (size_t)blockIdx.z * enc_words + c_off;
  *reinterpret_cast<ulonglong2 *>(o) = make_ulonglong2(a00.reduce(m), a01.reduce(m));
  *reinterpret_cast<ulonglong2 *>(o + k_stride) = make_ulonglong2(a10.reduce(m), a11.reduce(m));
}

template <int UNROLL>
__global__ void __launch_bounds__(256) k_crs_lincomb(const DevParams *__restrict__ P, const uint64_t *__restrict__ crs,
                                                     const uint32_t *__restrict__ term, const uint32_t *__restrict__ pidx,
                                                     uint32_t n_terms, uint32_t terms_per_split,
                                                     const uint64_t *__restrict__ pntt, uint64_t *__restrict__ partial,
                                                     const uint32_t *__restrict__ zoff = nullptr,
                                                     const uint8_t *__restrict__ slot_skip = nullptr,
                                                     const uint64_t *const *__restrict__ term_ptr = nullptr) {
  crs_lincomb_body<UNROLL>(P, crs, term, pidx, n_terms, terms_per_split, pntt, partial, zoff, slot_skip, term_ptr);
}
// The same stream with FOUR adjacent x per thread: 256-bit loads (one whole 32-byte sector per lane and instruction), eight 192-bit
// accumulators, 256-bit stores.  grid (N_E / (4*blockDim), L_R*L_E, splits).  What the static plan launches (RSG_LIN=narrow: the
// two-x kernel above): 5.5 -> 6.45 TB/s inside the C4 proof = 98.6 % of the measured copy peak.
struct u64x4 {
  uint64_t a, b, c, d;
};
__device__ __forceinline__ u64x4 ld_stream4(const uint64_t *p) {
  u64x4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
  return v;
}
template <int UNROLL>
__global__ void __launch_bounds__(256) k_crs_lincomb_wide(const DevParams *__restrict__ P, const uint32_t *__restrict__ pidx,
                                                          const uint64_t *__restrict__ pntt, uint64_t *__restrict__ partial,
                                                          const uint32_t *__restrict__ zoff, const uint8_t *__restrict__ slot_skip,
                                                          const uint64_t *const *__restrict__ term_ptr,
                                                          const uint32_t *__restrict__ zorder = nullptr, uint32_t n_paired = 0) {
  // zorder: launch position -> split.  Two splits that stream the SAME CRS encodings with different plaintexts (the ringGroth16 A
  // and B inner products both run over s_pows, groth16.tcc:89-104) come first, pair by pair, and their CTAs alternate in launch
  // order -- CTA 2i works on the first split of the pair and CTA 2i+1 on the same (x, limb) tile of the second -- so the two run
  // side by side and the second read of every encoding is an L2 hit instead of a second trip to HBM.
  uint32_t z = blockIdx.z, bx = blockIdx.x, by = blockIdx.y;
  if (zorder) {
    const uint32_t per = gridDim.x * gridDim.y, b = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    if (b < n_paired * per) {
      const uint32_t q = b / (2 * per), r = b - q * 2 * per, xy = r >> 1;
      by = xy / gridDim.x;
      bx = xy - by * gridDim.x;
      z = zorder[2 * q + (r & 1)];
    } else {
      z = zorder[blockIdx.z];
    }
  }
  const uint32_t N_E = P->N_E, L_E = P->L_E, L_R = P->L_R;
  const uint32_t x = 4 * (bx * blockDim.x + threadIdx.x);
  const uint32_t j = by / L_E, l = by - j * L_E;
  const uint32_t t0 = zoff[z], t1 = zoff[z + 1];
  const size_t poly = (size_t)N_E, ct_words = 2 * (size_t)L_E * poly, enc_words = (size_t)L_R * ct_words;
  const size_t c_off = (size_t)j * ct_words + (size_t)l * poly + x, k_stride = (size_t)L_E * poly;
  const size_t p_off = ((size_t)j * L_E + l) * poly + x, p_stride = (size_t)L_R * L_E * poly;
  Acc192 a0[4], a1[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { a0[i].clear(); a1[i].clear(); }
  auto mac4 = [&](const u64x4 &c0, const u64x4 &c1, const u64x4 &pp) {
    a0[0].mac(c0.a, pp.a); a0[1].mac(c0.b, pp.b); a0[2].mac(c0.c, pp.c); a0[3].mac(c0.d, pp.d);
    a1[0].mac(c1.a, pp.a); a1[1].mac(c1.b, pp.b); a1[2].mac(c1.c, pp.c); a1[3].mac(c1.d, pp.d);
  };
  const u64x4 ones = {1, 1, 1, 1}, zeros = {0, 0, 0, 0};
  uint32_t t = t0;
  for (; t + UNROLL <= t1; t += UNROLL) {
    u64x4 c0[UNROLL], c1[UNROLL], pp[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      const uint32_t pi = __ldg(pidx + t + u);
      const uint64_t *c = term_ptr[t + u] + c_off;
      if (slot_skip && pi != 0xFFFFFFFFu && slot_skip[pi]) {
        c0[u] = c1[u] = pp[u] = zeros;
        continue;
      }
      c0[u] = ld_stream4(c);
      c1[u] = ld_stream4(c + k_stride);
      pp[u] = pi != 0xFFFFFFFFu ? ld_stream4(pntt + (size_t)pi * p_stride + p_off) : ones;
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) mac4(c0[u], c1[u], pp[u]);
  }
  for (; t < t1; t++) {
    const uint32_t pi = __ldg(pidx + t);
    if (slot_skip && pi != 0xFFFFFFFFu && slot_skip[pi]) continue;
    const uint64_t *c = term_ptr[t] + c_off;
    const u64x4 c0 = ld_stream4(c), c1 = ld_stream4(c + k_stride);
    mac4(c0, c1, pi != 0xFFFFFFFFu ? ld_stream4(pntt + (size_t)pi * p_stride + p_off) : ones);
  }
  const ModConst m = P->Q[l];
  uint64_t *o = partial + (size_t)z * enc_words + c_off;
  asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(o), "l"(a0[0].reduce(m)), "l"(a0[1].reduce(m)), "l"(a0[2].reduce(m)), "l"(a0[3].reduce(m)) : "memory");
  asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(o + k_stride), "l"(a1[0].reduce(m)), "l"(a1[1].reduce(m)), "l"(a1[2].reduce(m)), "l"(a1[3].reduce(m)) : "memory");
}

// 64 registers x 256 threads = 16 Ki registers: the CTA that fits next to k_lift_fwd_ntt_f64_r96 on one SM.
__global__ void __maxnreg__(64) k_crs_lincomb_r64(const DevParams *__restrict__ P, const uint64_t *__restrict__ crs,
                                                  const uint32_t *__restrict__ term, const uint32_t *__restrict__ pidx,
                                                  uint32_t n_terms, uint32_t terms_per_split,
                                                  const uint64_t *__restrict__ pntt, uint64_t *__restrict__ partial,
                                                  const uint32_t *__restrict__ zoff, const uint8_t *__restrict__ slot_skip,
                                                  const uint64_t *const *__restrict__ term_ptr) {
  crs_lincomb_body<2>(P, crs, term, pidx, n_terms, terms_per_split, pntt, partial, zoff, slot_skip, term_ptr);
}

// out[w] = sum_s parts[s][w] mod Q_l(w): the modular-add kernel (after split-K or after the NCCL all-gather).
// `n_enc` encodings per part (a whole proof = 3), parts stored back to back.
// part_stride (words, 0 = back to back): distance between consecutive parts.
__global__ void __launch_bounds__(256) k_enc_sum(const DevParams *__restrict__ P, const uint64_t *__restrict__ parts,
                                                 uint32_t n_parts, uint32_t n_enc, uint64_t *__restrict__ out, size_t part_stride = 0) {
  const uint32_t N_E = P->N_E, L_E = P->L_E;
  const size_t enc_words = (size_t)n_enc * P->L_R * 2 * L_E * N_E;
  if (!part_stride) part_stride = enc_words;
  const size_t w = 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (w >= enc_words) return;
  const uint32_t l = (uint32_t)((w / N_E) % L_E);
  const uint64_t p = P->Q[l].p;
  ulonglong2 acc = *reinterpret_cast<const ulonglong2 *>(parts + w);
  for (uint32_t s = 1; s < n_parts; s++) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(parts + (size_t)s * part_stride + w);
    acc.x = add_mod(acc.x, v.x, p);
    acc.y = add_mod(acc.y, v.y, p);
  }
  *reinterpret_cast<ulonglong2 *>(out + w) = acc;
}

// out = a + b mod Q_l over one encoding (EncodingElem::operator+=, evaluator.cpp:217-231); out may alias a.
__global__ void __launch_bounds__(256) k_enc_add(const DevParams *__restrict__ P, const uint64_t *a, const uint64_t *__restrict__ b,
                                                 uint64_t *out) {
  const uint32_t N_E = P->N_E, L_E = P->L_E;
  const size_t enc_words = (size_t)P->L_R * 2 * L_E * N_E;
  const size_t w = 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (w >= enc_words) return;
  const uint64_t p = P->Q[(uint32_t)((w / N_E) % L_E)].p;
  const ulonglong2 x = *reinterpret_cast<const ulonglong2 *>(a + w), y = *reinterpret_cast<const ulonglong2 *>(b + w);
  *reinterpret_cast<ulonglong2 *>(out + w) = make_ulonglong2(add_mod(x.x, y.x, p), add_mod(x.y, y.y, p));
}

// comb[m][j][i] = centred(plain[pair[2m]][j][i]) + centred(plain[pair[2m+1]][j][i]) as int64 (second index 0xFFFFFFFF =
// none), centred(v) = v - t for v >= ceil(t/2) (evaluator.cpp:2220-2259 lifts exactly these representatives into Q).
// The sum of the lifts is the lift of the sum for EVERY Q_l, so the merged term needs one forward transform per limb.
// grid (M, L_R), 256 threads, two words per access.
__global__ void __launch_bounds__(256) k_centre_add(const DevParams *__restrict__ P, const uint64_t *__restrict__ plain,
                                                    const uint32_t *__restrict__ pair, uint64_t *__restrict__ comb) {
  const uint32_t mi = blockIdx.x, j = blockIdx.y, N_E = P->N_E, L_R = P->L_R;
  const uint64_t t = P->q[j].p, thr = P->thr[j];
  const uint32_t ia = pair[2 * mi], ib = pair[2 * mi + 1];
  const ulonglong2 *a = reinterpret_cast<const ulonglong2 *>(plain + ((size_t)ia * L_R + j) * N_E);
  const ulonglong2 *b = ib != 0xFFFFFFFFu ? reinterpret_cast<const ulonglong2 *>(plain + ((size_t)ib * L_R + j) * N_E) : nullptr;
  ulonglong2 *o = reinterpret_cast<ulonglong2 *>(comb + ((size_t)mi * L_R + j) * N_E);
  auto centre = [&](uint64_t v) { return v >= thr ? v - t : v; };   // two's complement
  for (uint32_t i = threadIdx.x; i < N_E / 2; i += blockDim.x) {
    ulonglong2 x = a[i];
    x.x = centre(x.x); x.y = centre(x.y);
    if (b) {
      const ulonglong2 y = b[i];
      x.x += centre(y.x); x.y += centre(y.y);
    }
    o[i] = x;
  }
}

// pval[g][j] = NTT_{Q_0}(lift(plain[g][j]))[0] = sum_i lift(plain[g][j][i]) * psi^i mod Q_0: the one NTT-domain word k_probe
// needs per term, evaluated directly (an N_E-term dot product) for the parts of a merged lincomb, whose separate
// transforms are never formed.  psi_pow[i] = psi^i mod Q_0 (psi = SEAL's minimal primitive 2N-th root).  grid (G, L_R).
__global__ void __launch_bounds__(256) k_probe_eval(const DevParams *__restrict__ P, const uint64_t *__restrict__ plain,
                                                    const uint64_t *__restrict__ psi_pow, uint64_t *__restrict__ pval) {
  __shared__ uint64_t part[8];
  const uint32_t g = blockIdx.x, j = blockIdx.y, N_E = P->N_E, L_R = P->L_R;
  const ModConst m = P->Q[0];
  const uint64_t thr = P->thr[j], tm = P->tmodQ[j][0];
  const uint64_t *src = plain + ((size_t)g * L_R + j) * N_E;
  Acc192 acc;
  acc.clear();
  for (uint32_t i = threadIdx.x; i < N_E; i += blockDim.x) {
    const uint64_t v = src[i];
    uint64_t r = reduce64(v, m);
    if (v >= thr) r = sub_mod(r, tm, m.p);
    acc.mac(r, __ldg(psi_pow + i));
  }
  uint64_t s = acc.reduce(m);
#pragma unroll
  for (int off = 16; off; off >>= 1) s = add_mod(s, __shfl_xor_sync(0xFFFFFFFFu, s, off), m.p);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t t = part[0];
    for (int w = 1; w < 8; w++) t = add_mod(t, part[w], m.p);
    pval[(size_t)g * L_R + j] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------
// "Transparent ciphertext" semantics.  SEAL throws when an addition yields a ciphertext whose c1 is identically zero
// (evaluator.cpp:233-239) and EncodingElem::operator+= answers by replacing that limb with an empty zero ciphertext
// (seal_ring.tcc:493-504): the running sum of an inner product is DROPPED whenever a prefix of it is transparent.
// k_probe makes that observable at negligible cost: per ring limb j it forms the running sum of the c1 contributions
// at one fixed slot (k = 1, l = 0, x = 0) over the term list and flags the terms where it is zero -- a necessary
// condition for the prefix to be transparent.  The host resolves flagged prefixes exactly (rsgpu.cu).
// grid (L_R), 256 threads; carry[j] holds the running sum across chunks of the term list.
// pv_stride != 0: the NTT-domain plaintext value of term t at the probe slot is pv[pidx[t] * pv_stride + j] (k_probe_eval
// output of a merged lincomb) instead of being read from pntt.
__global__ void __launch_bounds__(256) k_probe(const DevParams *__restrict__ P, const uint64_t *__restrict__ crs,
                                               const uint32_t *__restrict__ term, const uint32_t *__restrict__ pidx,
                                               uint32_t n_terms, const uint64_t *__restrict__ pntt, uint64_t *__restrict__ carry,
                                               uint8_t *__restrict__ flags, uint32_t flag_stride, uint32_t pv_stride) {
  __shared__ uint64_t warp_tot[8];
  __shared__ uint64_t run;
  const uint32_t j = blockIdx.x, N_E = P->N_E, L_E = P->L_E, L_R = P->L_R;
  const ModConst m = P->Q[0];
  const size_t ct_words = 2 * (size_t)L_E * N_E, enc_words = (size_t)L_R * ct_words;
  const size_t c_off = (size_t)j * ct_words + (size_t)L_E * N_E;   // k = 1, l = 0, x = 0
  const size_t p_off = pv_stride ? j : (size_t)j * L_E * N_E, p_stride = pv_stride ? pv_stride : (size_t)L_R * L_E * N_E;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) run = carry[j];
  __syncthreads();
  for (uint32_t base = 0; base < n_terms; base += 256) {
    const uint32_t t = base + threadIdx.x;
    uint64_t v = 0;
    if (t < n_terms) {
      const uint64_t cw = crs[(size_t)term[t] * enc_words + c_off];
      const uint32_t pi = pidx[t];
      v = pi != 0xFFFFFFFFu ? mul_mod(cw, pntt[(size_t)pi * p_stride + p_off], m) : cw;
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint64_t o = __shfl_up_sync(0xFFFFFFFFu, v, off);
      if (lane >= off) v = add_mod(v, o, m.p);
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    uint64_t pre = run;
    for (uint32_t w = 0; w < warp; w++) pre = add_mod(pre, warp_tot[w], m.p);
    v = add_mod(v, pre, m.p);
    if (t < n_terms) flags[(size_t)j * flag_stride + t] = v == 0;
    __syncthreads();
    if (threadIdx.x == 255) run = v;
    __syncthreads();
  }
  if (threadIdx.x == 0) carry[j] = run;
}

// nz[j] |= 1 iff the c1 polynomial (all L_E limbs) of ring limb j of one encoding has a non-zero word.
// grid (blocks, L_R); nz must be zeroed by the caller.
__global__ void __launch_bounds__(256) k_c1_nonzero(const DevParams *__restrict__ P, const uint64_t *__restrict__ enc,
                                                    uint32_t *__restrict__ nz) {
  const uint32_t j = blockIdx.y;
  const size_t half = (size_t)P->L_E * P->N_E;
  const uint64_t *c1 = enc + (size_t)j * 2 * half + half;
  uint32_t any = 0;
  for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < half; w += (size_t)gridDim.x * blockDim.x) any |= c1[w] != 0;
  if (__syncthreads_or((int)any) && threadIdx.x == 0) atomicOr(nz + j, 1u);
}
// Ring limbs whose c1 is identically zero become all-zero words (the reference's empty zero ciphertext). grid (blocks, L_R)
__global__ void __launch_bounds__(256) k_zero_transparent(const DevParams *__restrict__ P, uint64_t *__restrict__ enc,
                                                          const uint32_t *__restrict__ nz) {
  const uint32_t j = blockIdx.y;
  if (nz[j]) return;
  const size_t ct_words = 2 * (size_t)P->L_E * P->N_E;
  uint64_t *ct = enc + (size_t)j * ct_words;
  for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < ct_words; w += (size_t)gridDim.x * blockDim.x) ct[w] = 0;
}

// flags[e] = 1 iff bytes [0, W + 7) of element e are zero (W = words per element); one block per element.
__global__ void __launch_bounds__(256) k_is_zero_prefix(const uint64_t *__restrict__ ring, uint32_t W,
                                                        uint8_t *__restrict__ flags) {
  const uint64_t *src = ring + (size_t)blockIdx.x * W;
  const uint32_t bytes = W + 7, full = min(bytes / 8, W), rem = bytes % 8;
  uint32_t nz = 0;
  for (uint32_t i = threadIdx.x; i < full; i += blockDim.x) nz |= (src[i] != 0);
  if (threadIdx.x == 0 && rem && full < W) nz |= ((src[full] & ((1ull << (8 * rem)) - 1)) != 0);
  const int any = __syncthreads_or((int)nz);
  if (threadIdx.x == 0) flags[blockIdx.x] = any ? 0 : 1;
}

// Counter-based uniform residues (synthetic CRS / ring elements): word w of row r uses modulus mods[r % n_mods].
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void __launch_bounds__(256) k_fill_uniform(uint64_t *__restrict__ dst, size_t words, uint32_t row_words,
                                                      const ModConst *__restrict__ mods, uint32_t n_mods,
                                                      uint32_t rows_per_mod_cycle, uint64_t seed, uint64_t w_base = 0) {
  // w_base: index of dst[0] in the virtual (unsharded) array -- a shard filled with its global offset holds the same
  // words the whole array would (w_base must be a multiple of n_mods * rows_per_mod_cycle * row_words)
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += stride) {
    const size_t row = w / row_words;
    const uint64_t p = mods[(row / rows_per_mod_cycle) % n_mods].p;
    dst[w] = __umul64hi(splitmix64(seed ^ ((w + w_base) * 0xD1342543DE82EF95ull)), p);
  }
}

}  // namespace rsg
