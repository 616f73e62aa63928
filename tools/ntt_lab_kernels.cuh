// tools/ntt_lab_kernels.cuh -- EXPERIMENTAL variants of the lift + forward-NTT kernel, timed by tools/ntt_lab.cu against the
// shipped kernel (ringsnark_b200/csrc/kernels.cuh).  None of them ships: see DESIGN.md section 8 for the measurements.
//
//  * LiftIoSmallQ: the lift evaluator.cpp:2220-2259 performs (v mod Q_l for the centred plaintext coefficient v) with a SMALL
//    QUOTIENT: |v| < 2^55 and Q_l > 2^41 give |v / Q_l| < 2^14, so one float multiply estimates the quotient to within
//    1/2 + 2^-9 and x = v - q Q_l is a 32x64 product -- about a dozen integer instructions where the 64-bit Barrett
//    reduction + negate + centre + I2F.F64.S64 of LiftIoF64 took about forty.  Integer <-> double conversions use the
//    1.5 * 2^52 constant (exact for |x| < 2^51) instead of the XU conversion instructions.
//  * k_lift_fwd_ntt_f64_c2: one polynomial per 2-CTA thread-block cluster.  Each CTA holds HALF of the polynomial (68 KiB of
//    shared memory, 256 threads), so two CTAs of different clusters are resident per SM and the load / lift / barrier phases
//    of one polynomial overlap the butterflies of another (the 512-thread single-CTA kernel owns a whole SM: every barrier
//    and the whole lift drain the FP64 pipe).  Levels 0-1 are one radix-4 pass from global memory whose results go to the
//    owning CTA's shared memory -- the partner's half through distributed shared memory (st.shared::cluster) --, then one
//    cluster barrier, then twelve levels local to each half (three radix-16 passes, the last one straight to global memory).
//    Same butterflies in the same order as the single-CTA kernel: identical words.
#pragma once
#include <cooperative_groups.h>
#include "../ringsnark_b200/csrc/kernels.cuh"

namespace rsg {

// One 2^14-point polynomial per 2-CTA cluster.  grid (2 * count * L_E, L_R), 256 threads, padded_words(2^13) * 8 bytes.
template <bool SIGNED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 2)
    k_lift_fwd_ntt_f64_c2(const DevParams *__restrict__ P, const uint64_t *__restrict__ plain, uint64_t *__restrict__ out,
                          const uint8_t *__restrict__ slot_skip) {
  namespace cg = cooperative_groups;
  constexpr int LOGN = 14, LOGH = 13;
  constexpr uint32_t n = 1u << LOGN, quarter = n >> 2, half = n >> 1;
  extern __shared__ double smf[];
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t r = cluster.block_rank();
  const uint32_t L_R = P->L_R, L_E = P->L_E;
  const uint32_t el = blockIdx.x >> 1, e = el / L_E, l = el - e * L_E, j = blockIdx.y;
  if (slot_skip && slot_skip[e]) return;   // both CTAs of the cluster take this branch together
  LiftIoSmallQ<SIGNED> io;
  io.init(P, j, l);
  io.src = plain + (((size_t)e * L_R + j) << LOGN);
  io.dst = out + (((((size_t)e * L_R + j) * L_E + l)) << LOGN) + (size_t)r * half;
  const double *tab = P->fwdQ_f64[l];
  const double pd = io.pd, pinv = io.pinv;
  double *sm_peer = cluster.map_shared_rank(smf, r ^ 1);
  double *sm_lo = r == 0 ? smf : sm_peer;   // quarters 0, 1 live in rank 0
  double *sm_hi = r == 0 ? sm_peer : smf;   // quarters 2, 3 live in rank 1
  // levels 0, 1: radix-4 items o in [r * 2048, (r + 1) * 2048), elements o + k * 4096
  const double w0 = __ldg(tab + 1), w1a = __ldg(tab + 2), w1b = __ldg(tab + 3);
  const double w0p = __dmul_rn(w0, pinv), w1ap = __dmul_rn(w1a, pinv), w1bp = __dmul_rn(w1b, pinv);
  typename PassChainF<LOGH, 1, false>::Tw tw1;
  tw1.load(tab, 1, r, threadIdx.x);
  // the peer's shared memory may be written only once the peer CTA is running: arrive now, wait before the first store
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  constexpr int IPT = (int)(quarter / 2 / 256);   // items per thread (8)
  constexpr int BATCH = 4;                        // items whose loads are in flight together
#pragma unroll 1
  for (int it0 = 0; it0 < IPT; it0 += BATCH) {
    uint64_t raw[BATCH][4];
#pragma unroll
    for (int b = 0; b < BATCH; b++) {
      const uint32_t o = r * (quarter / 2) + (uint32_t)(it0 + b) * 256 + threadIdx.x;
#pragma unroll
      for (int k = 0; k < 4; k++) raw[b][k] = io.load_raw(o + k * quarter);
    }
    if (it0 == 0) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
#pragma unroll
    for (int b = 0; b < BATCH; b++) {
      const uint32_t o = r * (quarter / 2) + (uint32_t)(it0 + b) * 256 + threadIdx.x;
      double v0 = io.lift(raw[b][0]), v1 = io.lift(raw[b][1]), v2 = io.lift(raw[b][2]), v3 = io.lift(raw[b][3]);
      bfly_fwd_f64(v0, v2, w0, w0p, pd);
      bfly_fwd_f64(v1, v3, w0, w0p, pd);
      bfly_fwd_f64(v0, v1, w1a, w1ap, pd);
      bfly_fwd_f64(v2, v3, w1b, w1bp, pd);
      const uint32_t a0 = pad_idx(o), a1 = pad_idx(o + quarter);
      sm_lo[a0] = recentre_f64(v0, pd, pinv);
      sm_lo[a1] = recentre_f64(v1, pd, pinv);
      sm_hi[a0] = recentre_f64(v2, pd, pinv);
      sm_hi[a1] = recentre_f64(v3, pd, pinv);
    }
  }
  cluster.sync();
  // levels 2..13 inside this CTA's half: local transform of size 2^13 whose first level is done (S = 1), global block r
  PassChainF<LOGH, 1, false>::fwd(smf, tab, pd, pinv, 1, r, tw1, io);
}

// ---- persistent variants (lab): one CTA per SM loops over the (term, limb) polynomials ----------------------------------
// v3: the single-CTA kernel body in a grid-stride loop (no CTA turnover).
template <int LOGN, bool SIGNED>
__global__ void __launch_bounds__(512) k_lift_fwd_ntt_f64_v3(const DevParams *__restrict__ P, const uint64_t *__restrict__ plain,
                                                             uint64_t *__restrict__ out, const uint8_t *__restrict__ slot_skip,
                                                             uint32_t total) {
  extern __shared__ double smf[];
  const uint32_t L_R = P->L_R, L_E = P->L_E;
  for (uint32_t wk = blockIdx.x; wk < total; wk += gridDim.x) {
    const uint32_t el = wk / L_R, j = wk - el * L_R, e = el / L_E, l = el - e * L_E;
    if (slot_skip && slot_skip[e]) continue;
    LiftIoSmallQ<SIGNED> io;
    io.init(P, j, l);
    io.src = plain + (((size_t)e * L_R + j) << LOGN);
    io.dst = out + (((((size_t)e * L_R + j) * L_E + l)) << LOGN);
    const double *tab = P->fwdQ_f64[l];
    typename PassChainF<LOGN, 0, true>::Tw tw0;
    tw0.load(tab, 0, 0, threadIdx.x);
    PassChainF<LOGN, 0, true>::fwd(smf, tab, io.pd, io.pinv, 0, 0, tw0, io);
    __syncthreads();
  }
}

// v4: v3 + the raw coefficients of the NEXT polynomial's first item are requested before the last pass of the current one,
// and those of a thread's second first-pass item before the butterflies of its first: no exposed global-load latency.
template <bool SIGNED>
__global__ void __launch_bounds__(512) k_lift_fwd_ntt_f64_v4(const DevParams *__restrict__ P, const uint64_t *__restrict__ plain,
                                                             uint64_t *__restrict__ out, const uint8_t *__restrict__ slot_skip,
                                                             uint32_t total) {
  constexpr int LOGN = 14;
  constexpr uint32_t n = 1u << LOGN, g1 = n >> 4;
  extern __shared__ double smf[];
  const uint32_t L_R = P->L_R, L_E = P->L_E;
  auto decode = [&](uint32_t wk, uint32_t &e, uint32_t &l, uint32_t &j) {
    const uint32_t el = wk / L_R;
    j = wk - el * L_R;
    e = el / L_E;
    l = el - e * L_E;
  };
  auto next_live = [&](uint32_t wk) {   // first work item >= wk of this CTA that is not skipped
    while (wk < total) {
      uint32_t e, l, j;
      decode(wk, e, l, j);
      if (!(slot_skip && slot_skip[e])) break;
      wk += gridDim.x;
    }
    return wk;
  };
  uint32_t wk = next_live(blockIdx.x);
  uint64_t raw[16];
  if (wk < total) {
    uint32_t e, l, j;
    decode(wk, e, l, j);
    const uint64_t *src = plain + (((size_t)e * L_R + j) << LOGN);
#pragma unroll
    for (int k = 0; k < 16; k++) raw[k] = __ldg(src + threadIdx.x + k * g1);
  }
  while (wk < total) {
    uint32_t e, l, j;
    decode(wk, e, l, j);
    LiftIoSmallQ<SIGNED> io;
    io.init(P, j, l);
    io.src = plain + (((size_t)e * L_R + j) << LOGN);
    io.dst = out + (((((size_t)e * L_R + j) * L_E + l)) << LOGN);
    const double *tab = P->fwdQ_f64[l];
    const double pd = io.pd, pinv = io.pinv;
    {   // pass 1: levels 0..3, items threadIdx.x and threadIdx.x + 512 (same 15 twiddles)
      double w[15];
#pragma unroll
      for (int k = 0; k < 15; k++) w[k] = __ldg(tab + 1 + k);
      uint64_t rawB[16];
#pragma unroll
      for (int k = 0; k < 16; k++) rawB[k] = __ldg(io.src + threadIdx.x + 512 + k * g1);
#pragma unroll
      for (int it = 0; it < 2; it++) {
        double v[16];
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = io.lift(it == 0 ? raw[k] : rawB[k]);
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int half = 16 >> (u + 1);
#pragma unroll
          for (int grp = 0; grp < (1 << u); grp++) {
            const double ww = w[(1 << u) - 1 + grp], wp = __dmul_rn(ww, pinv);
#pragma unroll
            for (int k = 0; k < half; k++) bfly_fwd_f64(v[grp * 2 * half + k], v[grp * 2 * half + k + half], ww, wp, pd);
          }
        }
        double *ptr = smf + pad_idx(threadIdx.x + it * 512);
#pragma unroll
        for (int k = 0; k < 16; k++) ptr[k * g1 + ((k * g1) >> 4)] = recentre_f64(v[k], pd, pinv);
      }
    }
    PassTwF<LOGN, 4, 4> tw2;
    tw2.load(tab, 0, 0, threadIdx.x);
    __syncthreads();
    prefetch_pass_tw<LOGN, 4, 8>(tab, 0, 0);
    ntt_pass_f64<LOGN, 4, 4, false, false>(smf, tab, pd, pinv, 0, 0, tw2, io);
    PassTwF<LOGN, 4, 8> tw3;
    tw3.load(tab, 0, 0, threadIdx.x);
    __syncthreads();
    prefetch_pass_tw<LOGN, 2, 12>(tab, 0, 0);
    ntt_pass_f64<LOGN, 4, 8, false, false>(smf, tab, pd, pinv, 0, 0, tw3, io);
    PassTwF<LOGN, 2, 12> tw4;
    tw4.load(tab, 0, 0, threadIdx.x);
    // the next polynomial's first-item coefficients: in flight during the last pass
    const uint32_t nwk = next_live(wk + gridDim.x);
    if (nwk < total) {
      uint32_t e2, l2, j2;
      decode(nwk, e2, l2, j2);
      const uint64_t *src = plain + (((size_t)e2 * L_R + j2) << LOGN);
#pragma unroll
      for (int k = 0; k < 16; k++) raw[k] = __ldg(src + threadIdx.x + k * g1);
    }
    __syncthreads();
    ntt_pass_f64<LOGN, 2, 12, false, true>(smf, tab, pd, pinv, 0, 0, tw4, io);
    __syncthreads();   // shared memory is rewritten by the next polynomial's first pass
    wk = nwk;
  }
}

// ---- v5: pair items + 128-bit shared-memory accesses ---------------------------------------------------------------------
// Every radix-16 pass handles TWO adjacent items per thread (elements i, i+1 of each of the sixteen rows): one LDS.128 /
// STS.128 moves both, the fifteen twiddles (and their w/p products) serve 64 butterflies instead of 32, and a 512-thread CTA
// covers a 2^14-point pass with exactly one pair per thread (no item loop).  Padding: four words per 64 (pad2), which keeps
// pairs adjacent and 16-byte aligned and every pass free of bank conflicts (the last, radix-4 pass: two-way).
// MODE 5 = source reads alias one 8 KiB row (L1 hits); 6 = results stored only under a never-true condition; 7 = both
// MODE (diagnosis only, wrong results unless 0): 1 = pass 2 three times; 2 = no lift, no butterflies in pass 1;
// 3 = no butterflies / canonicalisation in pass 4; 4 = pass 2 three times without its shared-memory traffic
template <bool SIGNED, int MODE = 0>
__global__ void __launch_bounds__(512) k_lift_fwd_ntt_f64_v5(const DevParams *__restrict__ P, const uint64_t *__restrict__ plain,
                                                             uint64_t *__restrict__ out, const uint8_t *__restrict__ slot_skip) {
  constexpr int LOGN = 14;
  extern __shared__ double smf[];
  const uint32_t L_R = P->L_R, L_E = P->L_E;
  const uint32_t el = blockIdx.x, e = el / L_E, l = el - e * L_E, j = blockIdx.y;
  if (slot_skip && slot_skip[e]) return;
  LiftIoSmallQ<SIGNED> io;
  io.init(P, j, l);
  io.src = plain + (((size_t)e * L_R + j) << LOGN);
  io.dst = out + (((((size_t)e * L_R + j) * L_E + l)) << LOGN);
  const double *tab = P->fwdQ_f64[l];
  const double pd = io.pd, pinv = io.pinv;
  const uint32_t tid = threadIdx.x;
  double a[16], b[16], w[15];
  {   // pass 1: levels 0..3, rows 1024 apart, pair o = 2 tid
    const uint32_t o = 2 * tid;
    ulonglong2 raw[16];
#pragma unroll
    for (int k = 0; k < 16; k++) raw[k] = __ldg(reinterpret_cast<const ulonglong2 *>(io.src + o + ((MODE == 5 || MODE == 7 || MODE == 9) ? 0 : 1024 * k)));
    ld_tw15(w, tab, 0, 0);
    if (MODE == 2) {
#pragma unroll
      for (int k = 0; k < 16; k++) { a[k] = __hiloint2double(0x43300000, (int)raw[k].x); b[k] = __hiloint2double(0x43300000, (int)raw[k].y); }
    } else {
#pragma unroll
      for (int k = 0; k < 16; k++) { a[k] = io.lift(raw[k].x); b[k] = io.lift(raw[k].y); }
      radix16_pair(a, b, w, pd, pinv);
    }
    double *ptr = smf + pad2(o);
#pragma unroll
    for (int k = 0; k < 16; k++)
      *reinterpret_cast<double2 *>(ptr + 1024 * k + ((1024 * k) >> 4)) = make_double2(recentre_f64(a[k], pd, pinv), recentre_f64(b[k], pd, pinv));
  }
  {   // pass 2: levels 4..7, rows 64 apart inside block bb of 1024
    const uint32_t bb = tid >> 5, o = 2 * (tid & 31);
    ld_tw15(w, tab, 4, bb);
    __syncthreads();
    double *ptr = smf + pad2(bb * 1024 + o);
#pragma unroll 1
    for (int rep = 0; rep < ((MODE == 1 || MODE == 4) ? 3 : 1); rep++) {
      if (MODE != 4 || rep == 0) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
          const double2 t = *reinterpret_cast<const double2 *>(ptr + 64 * k + 4 * k);
          a[k] = t.x; b[k] = t.y;
        }
      }
      radix16_pair(a, b, w, pd, pinv);
      if (MODE == 4 && rep < 2) {
#pragma unroll
        for (int k = 0; k < 16; k++) { a[k] = recentre_f64(a[k], pd, pinv); b[k] = recentre_f64(b[k], pd, pinv); }
      } else {
#pragma unroll
        for (int k = 0; k < 16; k++)
          *reinterpret_cast<double2 *>(ptr + 64 * k + 4 * k) = make_double2(recentre_f64(a[k], pd, pinv), recentre_f64(b[k], pd, pinv));
      }
      if (MODE == 1 && rep < 2) __syncthreads();
    }
  }
  {   // pass 3: levels 8..11, rows 4 apart inside block bb of 64
    const uint32_t bb = tid >> 1, o = 2 * (tid & 1);
    ld_tw15(w, tab, 8, (MODE == 9 || MODE == 10) ? (bb & 15) : bb);
    __syncthreads();
    double *ptr = smf + pad2(bb * 64 + o);
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const double2 t = *reinterpret_cast<const double2 *>(ptr + 4 * k);
      a[k] = t.x; b[k] = t.y;
    }
    radix16_pair(a, b, w, pd, pinv);
#pragma unroll
    for (int k = 0; k < 16; k++)
      *reinterpret_cast<double2 *>(ptr + 4 * k) = make_double2(recentre_f64(a[k], pd, pinv), recentre_f64(b[k], pd, pinv));
  }
  {   // pass 4: levels 12, 13 on four consecutive elements; items tid + 512 m
    double w12[8];
    double2 w13[8];
#pragma unroll
    for (int m = 0; m < 8; m++) {
      const uint32_t i = (MODE == 9 || MODE == 10) ? tid : tid + 512 * m;
      w12[m] = ldg_f64_here(tab + 4096 + i);
      asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(w13[m].x), "=d"(w13[m].y) : "l"(tab + 8192 + 2 * i));
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 8; m++) {
      const uint32_t i = tid + 512 * m;
      const double *ptr = smf + pad2(4 * i);
      const double2 t0 = *reinterpret_cast<const double2 *>(ptr), t1 = *reinterpret_cast<const double2 *>(ptr + 2);
      double v0 = t0.x, v1 = t0.y, v2 = t1.x, v3 = t1.y;
      uint64_t *d = io.dst + 4 * i;
      if (MODE == 3) {
        *reinterpret_cast<ulonglong2 *>(d) = make_ulonglong2(__double_as_longlong(v0) + __double_as_longlong(w12[m]), __double_as_longlong(v1) + __double_as_longlong(w13[m].x));
        *reinterpret_cast<ulonglong2 *>(d + 2) = make_ulonglong2(__double_as_longlong(v2), __double_as_longlong(v3) + __double_as_longlong(w13[m].y));
        continue;
      }
      const double wa = w12[m], wap = __dmul_rn(wa, pinv);
      bfly_fwd_f64(v0, v2, wa, wap, pd);
      bfly_fwd_f64(v1, v3, wa, wap, pd);
      bfly_fwd_f64(v0, v1, w13[m].x, __dmul_rn(w13[m].x, pinv), pd);
      bfly_fwd_f64(v2, v3, w13[m].y, __dmul_rn(w13[m].y, pinv), pd);
      if ((MODE == 6 || MODE == 7 || MODE == 9) && v0 != 1.2345e-300) continue;
      if (MODE == 8 || MODE == 10) {   // one 256-bit store per item: a whole 32-byte sector per lane
        asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(d), "l"(io.canon(v0)), "l"(io.canon(v1)), "l"(io.canon(v2)), "l"(io.canon(v3)) : "memory");
        continue;
      }
      *reinterpret_cast<ulonglong2 *>(d) = make_ulonglong2(io.canon(v0), io.canon(v1));
      *reinterpret_cast<ulonglong2 *>(d + 2) = make_ulonglong2(io.canon(v2), io.canon(v3));
    }
  }
}

// ---- v6: v5's passes on a 2-CTA cluster ------------------------------------------------------------------------------------
// One polynomial per cluster of two 256-thread CTAs; CTA r keeps rows 8r..8r+7 (8 x 1024 doubles + padding = 68 KiB), so two
// CTAs of different clusters share an SM and the global loads / stores of one polynomial run under the butterflies of another.
// Pass 1 (levels 0-3, column pairs): CTA r transforms columns [512 r, 512 r + 512) and sends every row to its owner -- half of
// them through distributed shared memory.  One cluster barrier.  Passes 2-4 stay inside a row: one WARP per row, __syncwarp only.
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, double x, double y) {
  asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
template <bool SIGNED, int CS, int MINB, int MODE = 0>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(512 / CS, MINB)
    k_lift_fwd_ntt_f64_v6(const DevParams *__restrict__ P, const uint64_t *__restrict__ plain, uint64_t *__restrict__ out,
                          const uint8_t *__restrict__ slot_skip) {
  constexpr int LOGN = 14;
  constexpr uint32_t ROWW = 1024 + 64;   // padded row
  extern __shared__ double smf[];
  uint32_t r;
  asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  const uint32_t L_R = P->L_R, L_E = P->L_E;
  const uint32_t el = blockIdx.x / CS, e = el / L_E, l = el - e * L_E, j = blockIdx.y;
  if (slot_skip && slot_skip[e]) return;   // both CTAs of the cluster together
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  LiftIoSmallQ<SIGNED> io;
  io.init(P, j, l);
  io.src = plain + (((size_t)e * L_R + j) << LOGN);
  io.dst = out + ((((((size_t)e * L_R + j) * L_E + l)) % ((MODE & 4) ? 64 : 0x7FFFFFFF)) << LOGN);   // MODE 4: results into an 8 MiB window (L2)
  const double *tab = P->fwdQ_f64[l];
  const double pd = io.pd, pinv = io.pinv;
  const uint32_t tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  double a[16], b[16], w[15];
  {   // pass 1: column pair o, all 16 rows
    constexpr uint32_t TH = 512 / CS, RPC = 16 / CS;   // threads and rows per CTA
    const uint32_t o = 2 * (TH * r + tid);
    ulonglong2 raw[16];
#pragma unroll
    for (int k = 0; k < 16; k++) raw[k] = __ldg(reinterpret_cast<const ulonglong2 *>(io.src + o + ((MODE & 1) ? 0 : 1024 * k)));
    ld_tw15(w, tab, 0, 0);
#pragma unroll
    for (int k = 0; k < 16; k++) { a[k] = io.lift(raw[k].x); b[k] = io.lift(raw[k].y); }
    radix16_pair(a, b, w, pd, pinv);
    const uint32_t col = pad2(o);   // o < 1024: offset inside a padded row
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(smf) + col * 8;
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // the peer CTAs are running: their shared memory exists
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const uint32_t owner = k / RPC, rowoff = (uint32_t)(k % RPC) * ROWW * 8;
      uint32_t dstaddr;
      asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dstaddr) : "r"(local + rowoff), "r"(owner));
      st_cluster_v2(dstaddr, recentre_f64(a[k], pd, pinv), recentre_f64(b[k], pd, pinv));
    }
  }
  const uint32_t row = (16 / CS) * r + wrp;            // global row (block of 1024) this warp owns from here on
  double *rp = smf + wrp * ROWW;
  ld_tw15(w, tab, 4, row);
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  {   // pass 2: levels 4..7, elements 64 apart inside the row
    double *ptr = rp + pad2(2 * lane);
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
#pragma unroll
    for (int k = 0; k < 16; k++) *reinterpret_cast<double2 *>(ptr + 4 * k) = make_double2(recentre_f64(a[k], pd, pinv), recentre_f64(b[k], pd, pinv));
  }
  {   // pass 4: levels 12, 13; items lane + 32 m of the row's 256
    double w12[8];
    double2 w13[8];
#pragma unroll
    for (int m = 0; m < 8; m++) {
      const uint32_t i = row * 256 + lane + 32 * m;
      w12[m] = ldg_f64_here(tab + 4096 + i);
      asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(w13[m].x), "=d"(w13[m].y) : "l"(tab + 8192 + 2 * i));
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
      bfly_fwd_f64(v0, v1, w13[m].x, __dmul_rn(w13[m].x, pinv), pd);
      bfly_fwd_f64(v2, v3, w13[m].y, __dmul_rn(w13[m].y, pinv), pd);
      uint64_t *d = io.dst + 1024 * row + 4 * il;
      if ((MODE & 2) && v0 != 1.2345e-300) continue;
      asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(d), "l"(io.canon(v0)), "l"(io.canon(v1)), "l"(io.canon(v2)), "l"(io.canon(v3)) : "memory");
    }
  }
}

}  // namespace rsg
