// tools/fp64_lab.cu -- what the FP64 pipe of a B200 SM sustains on the NTT's instruction mix (registers only, no memory):
//   m1: independent DFMA chains                    -> the pipe's peak
//   m2: radix-16 passes (4 levels, 8 FP64 ops per butterfly + recentre) on register values
//   m3: m2 + shared-memory round trip and __syncthreads per pass
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/fp64_lab tools/fp64_lab.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../ringsnark_b200/csrc/ntt_f64.cuh"
using namespace rsg;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int ILP>
__global__ void m1(double *out, int iters, double a, double b) {
  double v[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) v[k] = threadIdx.x + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) v[k] = __fma_rn(v[k], a, b);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += v[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <bool SMEM>
__global__ void m2(double *out, const double *tw, int iters, double p, double pinv) {
  extern __shared__ double sm[];
  double v[16], w[15];
#pragma unroll
  for (int k = 0; k < 16; k++) v[k] = (double)((threadIdx.x * 16 + k) * 7919 % 100003);
#pragma unroll
  for (int k = 0; k < 15; k++) w[k] = tw[k];
  for (int i = 0; i < iters; i++) {
    if (SMEM) {
#pragma unroll
      for (int k = 0; k < 16; k++) v[k] = sm[pad_idx(threadIdx.x + k * blockDim.x)];
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int half = 16 >> (u + 1);
#pragma unroll
      for (int grp = 0; grp < (1 << u); grp++) {
        const double ww = w[(1 << u) - 1 + grp], wp = __dmul_rn(ww, pinv);
#pragma unroll
        for (int k = 0; k < half; k++) bfly_fwd_f64(v[grp * 2 * half + k], v[grp * 2 * half + k + half], ww, wp, p);
      }
    }
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = recentre_f64(v[k], p, pinv);
    if (SMEM) {
#pragma unroll
      for (int k = 0; k < 16; k++) sm[pad_idx(threadIdx.x * 16 + k)] = v[k];
      __syncthreads();
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) s += v[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m4: m2 + EXTRA independent integer (KIND 0: 32-bit add/xor; KIND 1: 64-bit IMAD-style multiply-add) or FP32 (KIND 2) instructions
// per item: does the FP64 stream leave issue slots for them?
template <int EXTRA, int KIND>
__global__ void m4(double *out, const double *tw, int iters, double p, double pinv) {
  double v[16], w[15];
#pragma unroll
  for (int k = 0; k < 16; k++) v[k] = (double)((threadIdx.x * 16 + k) * 7919 % 100003);
#pragma unroll
  for (int k = 0; k < 15; k++) w[k] = tw[k];
  uint32_t a[8];
  float f[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { a[k] = threadIdx.x * 3 + k; f[k] = threadIdx.x + k; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int half = 16 >> (u + 1);
#pragma unroll
      for (int grp = 0; grp < (1 << u); grp++) {
        const double ww = w[(1 << u) - 1 + grp], wp = __dmul_rn(ww, pinv);
#pragma unroll
        for (int k = 0; k < half; k++) bfly_fwd_f64(v[grp * 2 * half + k], v[grp * 2 * half + k + half], ww, wp, p);
      }
#pragma unroll
      for (int x = 0; x < EXTRA / 4; x++) {
        if (KIND == 0) a[x & 7] = (a[x & 7] + a[(x + 3) & 7]) ^ (uint32_t)i;
        else if (KIND == 1) a[x & 7] = a[x & 7] * a[(x + 3) & 7] + (uint32_t)i;
        else f[x & 7] = __fmaf_rn(f[x & 7], f[(x + 3) & 7], 1.0f);
      }
    }
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = recentre_f64(v[k], p, pinv);
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) s += v[k];
  uint32_t t = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) t += a[k] + (uint32_t)f[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}

int main() {
  double *out, *tw;
  CK(cudaMalloc(&out, 148 * 8 * 1024 * 8));
  CK(cudaMalloc(&tw, 16 * 8));
  double htw[16];
  for (int i = 0; i < 16; i++) htw[i] = 123456789012345.0 + 1000003.0 * i;
  CK(cudaMemcpy(tw, htw, sizeof(htw), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double p = 562949952798721.0, pinv = 1.0 / p;
  const double peak = 148.0 * 64 * 1.965e9;
  auto time = [&](auto launch, double ops, const char *name) {
    launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-58s %8.3f ms  %6.2f T lane-ops/s = %5.1f %% of 148 x 64 x 1.965 GHz\n", name, ms, ops / ms * 1e-9, ops / ms * 1e-9 / (peak * 1e-12) * 100);
  };
  const int it1 = 20000;
  time([&] { m1<8><<<148, 512>>>(out, it1, 1.0000001, 0.5); }, 148.0 * 512 * 8 * it1, "m1 DFMA, ILP 8, 512 thr x 1 CTA/SM");
  time([&] { m1<8><<<148, 1024>>>(out, it1, 1.0000001, 0.5); }, 148.0 * 1024 * 8 * it1, "m1 DFMA, ILP 8, 1024 thr x 1 CTA/SM");
  time([&] { m1<2><<<148, 512>>>(out, it1, 1.0000001, 0.5); }, 148.0 * 512 * 2 * it1, "m1 DFMA, ILP 2, 512 thr (latency probe)");
  time([&] { m1<1><<<148, 128>>>(out, it1, 1.0000001, 0.5); }, 148.0 * 128 * 1 * it1, "m1 DFMA, ILP 1, 128 thr (1 warp/SMSP: latency)");
  const int it2 = 2000;
  const double ops2 = 32 * 8 + 15 + 16 * 3;   // per item
  time([&] { m2<false><<<148, 512>>>(out, tw, it2, p, pinv); }, 148.0 * 512 * ops2 * it2, "m2 radix-16 items in registers, 512 thr x 1 CTA/SM");
  time([&] { m2<false><<<148 * 2, 256>>>(out, tw, it2, p, pinv); }, 148.0 * 512 * ops2 * it2, "m2 radix-16 items in registers, 256 thr x 2 CTA/SM");
  time([&] { m2<false><<<148, 256>>>(out, tw, it2, p, pinv); }, 148.0 * 256 * ops2 * it2, "m2 radix-16 items in registers, 256 thr x 1 CTA/SM");
  time([&] { m2<false><<<148, 128>>>(out, tw, it2, p, pinv); }, 148.0 * 128 * ops2 * it2, "m2 radix-16 items in registers, 128 thr x 1 CTA/SM");
  const size_t smb = (size_t)padded_words(512 * 16) * 8;
  CK(cudaFuncSetAttribute(m2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb));
  time([&] { m2<true><<<148, 512, smb>>>(out, tw, it2, p, pinv); }, 148.0 * 512 * ops2 * it2, "m3 + smem round trip + barrier, 512 thr x 1 CTA/SM");
  const size_t smb2 = (size_t)padded_words(256 * 16) * 8;
  time([&] { m2<true><<<148 * 2, 256, smb2>>>(out, tw, it2, p, pinv); }, 148.0 * 512 * ops2 * it2, "m3 + smem round trip + barrier, 256 thr x 2 CTA/SM");
  time([&] { m4<0, 0><<<148, 512>>>(out, tw, it2, p, pinv); }, 148.0 * 512 * ops2 * it2, "m4 + 0 extra");
  time([&] { m4<64, 0><<<148, 512>>>(out, tw, it2, p, pinv); }, 148.0 * 512 * ops2 * it2, "m4 + 64 x (IADD+LOP) per item (319 FP64)");
  time([&] { m4<128, 0><<<148, 512>>>(out, tw, it2, p, pinv); }, 148.0 * 512 * ops2 * it2, "m4 + 128 x (IADD+LOP) per item");
  time([&] { m4<256, 0><<<148, 512>>>(out, tw, it2, p, pinv); }, 148.0 * 512 * ops2 * it2, "m4 + 256 x (IADD+LOP) per item");
  time([&] { m4<128, 1><<<148, 512>>>(out, tw, it2, p, pinv); }, 148.0 * 512 * ops2 * it2, "m4 + 128 x IMAD per item");
  time([&] { m4<256, 1><<<148, 512>>>(out, tw, it2, p, pinv); }, 148.0 * 512 * ops2 * it2, "m4 + 256 x IMAD per item");
  time([&] { m4<128, 2><<<148, 512>>>(out, tw, it2, p, pinv); }, 148.0 * 512 * ops2 * it2, "m4 + 128 x FFMA per item");
  time([&] { m4<256, 2><<<148, 512>>>(out, tw, it2, p, pinv); }, 148.0 * 512 * ops2 * it2, "m4 + 256 x FFMA per item");
  return 0;
}
