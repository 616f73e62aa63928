// tools/ntt_lab.cu -- stand-alone timing bench for variants of the lift + forward-NTT kernel (C4 parameters: N_E = 2^14,
// eight 48/49-bit limbs, one 54-bit plaintext modulus).  Every variant is checked word for word against the shipped
// kernel.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/ntt_lab tools/ntt_lab.cu
// Run (GPU box): tools/ntt_lab [terms]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../ringsnark_b200/csrc/kernels.cuh"
#include "ntt_lab_kernels.cuh"

using namespace rsg;
typedef unsigned __int128 u128;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

static uint64_t mulmod(uint64_t a, uint64_t b, uint64_t p) { return (uint64_t)(((u128)a * b) % p); }
static uint64_t powmod(uint64_t a, uint64_t e, uint64_t p) {
  uint64_t r = 1;
  while (e) { if (e & 1) r = mulmod(r, a, p); a = mulmod(a, a, p); e >>= 1; }
  return r;
}
static uint64_t min_root(uint64_t degree, uint64_t p) {
  uint64_t root = 0;
  for (uint64_t g = 2; g < p; g++) {
    uint64_t r = powmod(g, (p - 1) / degree, p);
    if (powmod(r, degree / 2, p) == p - 1) { root = r; break; }
  }
  uint64_t sq = mulmod(root, root, p), cur = root, best = root;
  for (uint64_t i = 0; i < degree / 2; i++) { if (cur < best) best = cur; cur = mulmod(cur, sq, p); }
  return best;
}
static uint32_t bitrev(uint32_t x, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
}

__global__ void k_diff(const uint64_t *a, const uint64_t *b, size_t n, unsigned long long *cnt) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && a[i] != b[i]) atomicAdd(cnt, 1ull);
}
__global__ void k_fill_src(uint64_t *d, size_t n, uint64_t seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t x = splitmix64(seed ^ (i * 0xD1342543DE82EF95ull));
  long long v = (long long)(x >> 9) - (1ll << 54);     // uniform in [-2^54, 2^54)
  if ((i & 1023) == 0) v = (1ll << 54) - 1;            // extremes
  if ((i & 1023) == 1) v = -(1ll << 54) + 1;
  d[i] = (uint64_t)v;
}

int main(int argc, char **argv) {
  const int logn = 14;
  const size_t N = 1u << logn, L_E = 8;
  const size_t terms = argc > 1 ? atoi(argv[1]) : 1158;
  const int reps = 5;
  const uint64_t Q[8] = {281474976546817ull, 281474976317441ull, 281474975662081ull, 562949952798721ull,
                         562949952700417ull, 562949952274433ull, 562949951979521ull, 562949951881217ull};
  const uint64_t t = 18014398508400641ull;
  DevParams hp;
  memset(&hp, 0, sizeof(hp));
  hp.N_R = 2048; hp.L_R = 1; hp.N_E = (uint32_t)N; hp.L_E = (uint32_t)L_E; hp.logN_E = logn;
  hp.thr[0] = (t + 1) >> 1;
  for (size_t l = 0; l < L_E; l++) {
    const uint64_t p = Q[l];
    hp.Q[l].p = p;
    u128 all = ~(u128)0, ratio = all / p;
    hp.Q[l].ratio0 = (uint64_t)ratio; hp.Q[l].ratio1 = (uint64_t)(ratio >> 64); hp.Q[l].r128 = (uint64_t)((all % p + 1) % p);
    hp.tmodQ[0][l] = t % p;
    const uint64_t psi = min_root(2 * N, p);
    std::vector<double> tf(N, 1.0);
    uint64_t pw = 1;
    for (size_t i = 1; i < N; i++) { pw = mulmod(pw, psi, p); tf[bitrev((uint32_t)i, logn)] = (double)pw; }
    double *d;
    CK(cudaMalloc(&d, N * 8));
    CK(cudaMemcpy(d, tf.data(), N * 8, cudaMemcpyHostToDevice));
    hp.fwdQ_f64[l] = d;
    hp.Qinv_f64[l] = (double)(1.0L / (long double)p);
  }
  DevParams *dP;
  CK(cudaMalloc(&dP, sizeof(hp)));
  CK(cudaMemcpy(dP, &hp, sizeof(hp), cudaMemcpyHostToDevice));
  uint64_t *src, *out0, *out1;
  unsigned long long *cnt;
  CK(cudaMalloc(&src, terms * N * 8));
  CK(cudaMalloc(&out0, terms * L_E * N * 8));
  CK(cudaMalloc(&out1, terms * L_E * N * 8));
  CK(cudaMalloc(&cnt, 8));
  k_fill_src<<<(unsigned)((terms * N + 255) / 256), 256>>>(src, terms * N, 12345);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const size_t sm_full = (size_t)padded_words(1u << 14) * 8, sm_half = (size_t)padded_words(1u << 13) * 8;
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64<14, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_full));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64<14, 0, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_full));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_c2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_half));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v3<14, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_full));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v5<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_full));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v5<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_full));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v5<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_full));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v5<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_full));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v5<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_full));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v5<true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_full));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v5<true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_full));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v5<true, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_full));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v6<true, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (1024 + 64) * 8));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v6<true, 2, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (1024 + 64) * 8));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v6<true, 2, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (1024 + 64) * 8));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v6<true, 2, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (1024 + 64) * 8));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v6<true, 2, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (1024 + 64) * 8));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v6<true, 2, 2, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (1024 + 64) * 8));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v6<true, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (1024 + 64) * 8));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v6<true, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (1024 + 64) * 8));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v6<true, 4, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (1024 + 64) * 8));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v6<true, 8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (1024 + 64) * 8));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v5<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_full));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v5<true, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_full));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v5<true, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_full));
  CK(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_v4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_full));
  auto run = [&](int variant, uint64_t *out) {
    dim3 grid((unsigned)(terms * L_E), 1);
    if (variant == 0) k_lift_fwd_ntt_f64<14, 0, true><<<grid, 512, sm_full>>>(dP, src, out, nullptr);
    else if (variant == 1) k_lift_fwd_ntt_f64<14, 0, true, true><<<grid, 512, sm_full>>>(dP, src, out, nullptr);
    else if (variant == 2) k_lift_fwd_ntt_f64_c2<true><<<dim3((unsigned)(terms * L_E * 2), 1), 256, sm_half>>>(dP, src, out, nullptr);
    else if (variant == 3) k_lift_fwd_ntt_f64_v3<14, true><<<148, 512, sm_full>>>(dP, src, out, nullptr, (uint32_t)(terms * L_E));
    else if (variant == 5) k_lift_fwd_ntt_f64_v5<true><<<grid, 512, sm_full>>>(dP, src, out, nullptr);
    else if (variant == 6) k_lift_fwd_ntt_f64_v5<true, 1><<<grid, 512, sm_full>>>(dP, src, out, nullptr);
    else if (variant == 7) k_lift_fwd_ntt_f64_v5<true, 2><<<grid, 512, sm_full>>>(dP, src, out, nullptr);
    else if (variant == 8) k_lift_fwd_ntt_f64_v5<true, 3><<<grid, 512, sm_full>>>(dP, src, out, nullptr);
    else if (variant == 9) k_lift_fwd_ntt_f64_v5<true, 4><<<grid, 512, sm_full>>>(dP, src, out, nullptr);
    else if (variant == 10) k_lift_fwd_ntt_f64_v5<true, 5><<<grid, 512, sm_full>>>(dP, src, out, nullptr);
    else if (variant == 11) k_lift_fwd_ntt_f64_v5<true, 6><<<grid, 512, sm_full>>>(dP, src, out, nullptr);
    else if (variant == 12) k_lift_fwd_ntt_f64_v5<true, 7><<<grid, 512, sm_full>>>(dP, src, out, nullptr);
    else if (variant == 13) k_lift_fwd_ntt_f64_v5<true, 8><<<grid, 512, sm_full>>>(dP, src, out, nullptr);
    else if (variant == 14) k_lift_fwd_ntt_f64_v5<true, 9><<<grid, 512, sm_full>>>(dP, src, out, nullptr);
    else if (variant == 15) k_lift_fwd_ntt_f64_v5<true, 10><<<grid, 512, sm_full>>>(dP, src, out, nullptr);
    else if (variant == 16) k_lift_fwd_ntt_f64_v6<true, 2, 2><<<dim3((unsigned)(terms * L_E * 2), 1), 256, 8 * (1024 + 64) * 8>>>(dP, src, out, nullptr);
    else if (variant == 21) k_lift_fwd_ntt_f64_v6<true, 2, 2, 1><<<dim3((unsigned)(terms * L_E * 2), 1), 256, 8 * (1024 + 64) * 8>>>(dP, src, out, nullptr);
    else if (variant == 22) k_lift_fwd_ntt_f64_v6<true, 2, 2, 2><<<dim3((unsigned)(terms * L_E * 2), 1), 256, 8 * (1024 + 64) * 8>>>(dP, src, out, nullptr);
    else if (variant == 23) k_lift_fwd_ntt_f64_v6<true, 2, 2, 3><<<dim3((unsigned)(terms * L_E * 2), 1), 256, 8 * (1024 + 64) * 8>>>(dP, src, out, nullptr);
    else if (variant == 25) k_lift_fwd_ntt_f64_v6<true, 2, 2, 4><<<dim3((unsigned)(terms * L_E * 2), 1), 256, 8 * (1024 + 64) * 8>>>(dP, src, out, nullptr);
    else if (variant == 26) k_lift_fwd_ntt_f64_v6<true, 2, 2, 5><<<dim3((unsigned)(terms * L_E * 2), 1), 256, 8 * (1024 + 64) * 8>>>(dP, src, out, nullptr);
    else if (variant == 17) k_lift_fwd_ntt_f64_v6<true, 2, 3><<<dim3((unsigned)(terms * L_E * 2), 1), 256, 8 * (1024 + 64) * 8>>>(dP, src, out, nullptr);
    else if (variant == 18) k_lift_fwd_ntt_f64_v6<true, 4, 4><<<dim3((unsigned)(terms * L_E * 4), 1), 128, 4 * (1024 + 64) * 8>>>(dP, src, out, nullptr);
    else if (variant == 19) k_lift_fwd_ntt_f64_v6<true, 4, 5><<<dim3((unsigned)(terms * L_E * 4), 1), 128, 4 * (1024 + 64) * 8>>>(dP, src, out, nullptr);
    else if (variant == 20) k_lift_fwd_ntt_f64_v6<true, 8, 8><<<dim3((unsigned)(terms * L_E * 8), 1), 64, 2 * (1024 + 64) * 8>>>(dP, src, out, nullptr);
    else if (variant == 24) k_lift_fwd_ntt_f64_cl<true, true><<<dim3((unsigned)(terms * L_E * NTT_CL), 1), NTT_CL_THREADS, NTT_CL_SMEM>>>(dP, src, out, nullptr);
    else if (variant == 4) k_lift_fwd_ntt_f64_v4<true><<<148, 512, sm_full>>>(dP, src, out, nullptr, (uint32_t)(terms * L_E));
  };
  const char *names[] = {"v0 shipped (512 thr, 1 CTA/SM)", "v1 cheap lift + magic conversions", "v2 2-CTA cluster, DSMEM first pass, 2 CTAs/SM",
                         "v3 persistent v1 (148 CTAs, grid-stride)", "v4 persistent + next polynomial's loads in flight",
                         "v5 pair items, 128-bit smem, shared twiddles",
                         "v5 diag: pass 2 three times", "v5 diag: pass 1 without lift and butterflies", "v5 diag: pass 4 without butterflies / canon",
                         "v5 diag: pass 2 three times, registers only", "v5 diag: source reads hit L1", "v5 diag: no result stores",
                         "v5 diag: L1 source + no stores", "v5 + 256-bit result stores", "v5 diag: L1 source + no stores + hot twiddles",
                         "v5 diag: 256-bit stores + hot twiddles", "v6 cluster of 2 x 256 thr, warp per row, 2 CTAs/SM",
                         "v6 cluster 2 x 256, 3 CTAs/SM (85 regs)", "v6 cluster 4 x 128, 4 CTAs/SM", "v6 cluster 4 x 128, 5 CTAs/SM (102 regs)",
                         "v6 cluster 8 x 64, 8 CTAs/SM", "v6 diag: L1 source", "v6 diag: no stores", "v6 diag: L1 source + no stores", "SHIPPED k_lift_fwd_ntt_f64_cl (kernels.cuh)",
                         "v6 diag: results into an 8 MiB window (L2-resident)", "v6 diag: L1 source + results into 8 MiB window"};
  run(0, out0);
  CK(cudaDeviceSynchronize());
  for (int v = 0; v < 27; v++) {
    run(v, out1);   // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; r++) run(v, out1);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    CK(cudaMemset(cnt, 0, 8));
    const size_t words = terms * L_E * N;
    k_diff<<<(unsigned)((words + 255) / 256), 256>>>(out0, out1, words, cnt);
    unsigned long long bad = 0;
    CK(cudaMemcpy(&bad, cnt, 8, cudaMemcpyDeviceToHost));
    const double polys = (double)terms * L_E;
    printf("%-48s %8.3f ms  %7.1f G butterflies/s  scaled to 37056 polys: %6.3f ms  mismatches %llu\n", names[v], ms,
           polys * 114688.0 / ms * 1e-6, ms * 37056.0 / polys, bad);
  }
  return 0;
}
