// The N-GPU prover's two exchange steps as kernels over NVLink peer memory: every rank maps the others'
// buffers (CUDA IPC / symmetric memory, set up by the host driver) and the data moves by plain loads and stores through
// NVSwitch -- no NCCL call, no pack / unpack pass.  At C4 on 8 GPUs each NCCL collective costs 70-170 us (latency, not
// bandwidth: 0.75-6 MiB per peer) against a proof of 1.8 ms; these kernels move the same bytes in a few microseconds and are
// ordered by the driver's device-side barriers (one after each).
//   k_exchange_p2p  slots <-> terms: a rank holds slots [r S, (r+1) S) of EVERY witness coefficient and writes them straight
//                   into the term owner's coefficient buffer, already in the layout the lincomb phase reads ([5][per][L_R][N_R]).
//   k_enc_sum_p2p   the modular sum of the G partial proofs, reduce-scatter + all-gather in one launch: rank r loads slice r of
//                   every rank's partial, adds mod Q_l and stores the sum into every rank's proof buffer; the probe blocks that
//                   ride behind the partials (rsg_groth16_lincombs_shard) are copied to a local buffer for the host check.
#pragma once
#include "kernels.cuh"

namespace rsg {

constexpr int P2P_MAX = 16;
struct PeerPtrs {
  uint64_t *p[P2P_MAX];
};

// wit: [7n+2 rows][L_R * S] = A_io | B_io | C_io | A_mid | B_mid | C_mid (n rows each) | H (n+1) | one zero row, this rank's slots.
// full[d]: [5 (A_io, A_mid, B_io, B_mid, H)][per][L_R][G * S] on rank d = all slots of d's term range.  grid (5 per, G, L_R).
__global__ void __launch_bounds__(128) k_exchange_p2p(const uint64_t *__restrict__ wit, PeerPtrs full, uint32_t n, uint32_t per, uint32_t S,
                                                      uint32_t L_R, uint32_t G, uint32_t rank) {
  const uint32_t v = blockIdx.x / per, i = blockIdx.x - v * per, d = blockIdx.y, limb = blockIdx.z;
  const uint32_t k = d * per + i;
  const uint32_t base = v == 0 ? 0 : (v == 1 ? 3 * n : (v == 2 ? n : (v == 3 ? 4 * n : 6 * n)));
  const uint32_t length = v == 4 ? n + 1 : n;
  const ulonglong2 *src = k < length ? reinterpret_cast<const ulonglong2 *>(wit + ((size_t)(base + k) * L_R + limb) * S) : nullptr;
  ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(full.p[d] + (((size_t)v * per + i) * L_R + limb) * ((size_t)G * S) + (size_t)rank * S);
  for (uint32_t s = threadIdx.x; s < S / 2; s += blockDim.x) dst[s] = src ? src[s] : make_ulonglong2(0, 0);
}

// parts[s]: rank s's record [words3 = n_enc encodings | bw probe-block words]; finals[d]: rank d's proof buffer (words3).
// grid covers words3 / G / 2 pairs (the slice) -- plus the probe blocks, copied by the same threads in a grid-stride loop.
__global__ void __launch_bounds__(256) k_enc_sum_p2p(const DevParams *__restrict__ P, PeerPtrs parts, PeerPtrs finals, uint32_t G, uint32_t rank,
                                                     size_t words3, uint32_t bw, uint64_t *__restrict__ blocks) {
  const uint32_t N_E = P->N_E, L_E = P->L_E;
  const size_t slice = words3 / G;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t idx = t; idx < (size_t)G * bw; idx += (size_t)gridDim.x * blockDim.x) blocks[idx] = parts.p[idx / bw][words3 + idx % bw];
  const size_t w = 2 * t;
  if (w >= slice) return;
  const size_t gw = (size_t)rank * slice + w;
  const uint64_t p = P->Q[(uint32_t)((gw / N_E) % L_E)].p;
  ulonglong2 acc = *reinterpret_cast<const ulonglong2 *>(parts.p[0] + gw);
  for (uint32_t s = 1; s < G; s++) {
    const ulonglong2 x = *reinterpret_cast<const ulonglong2 *>(parts.p[s] + gw);
    acc.x = add_mod(acc.x, x.x, p);
    acc.y = add_mod(acc.y, x.y, p);
  }
  for (uint32_t d = 0; d < G; d++) *reinterpret_cast<ulonglong2 *>(finals.p[d] + gw) = acc;
}

}  // namespace rsg
