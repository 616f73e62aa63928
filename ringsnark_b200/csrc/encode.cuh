// EncodingElem::encode on the device: BatchEncoder::encode + Encryptor::encrypt_symmetric for BGV,
//   ringsnark/seal/seal_ring.tcc:324-359 -> depends/SEAL/native/src/seal/encryptor.cpp:242-312 -> util/rlwe.cpp:276-390,
// straight into a CRS arena, with SEAL's own randomness so that a seeded context reproduces SEAL's ciphertext words:
//   * Blake2xbPRNG (randomgen.cpp:201-211): the stream of a 64-byte seed is, for counter = 0, 1, ..., the 4096 bytes
//     blake2xb(outlen 4096, in = counter, key = seed) (util/blake2xb.c; BLAKE2b is RFC 7693).  Buffers and their 64-byte
//     blocks are independent, so the 1 MiB uniform polynomial of one ciphertext is 16 384 parallel compressions;
//   * rlwe.cpp:321-328: bootstrap stream -> 64-byte public seed -> ciphertext stream;
//   * sample_poly_uniform (rlwe.cpp:106-131): bulk fill, then every word >= max_multiple is redrawn from the words that follow
//     the bulk, in limb-major order (SEAL's default primes sit just below a power of two, which makes this a ~2^-31 event;
//     it is resolved by one thread per ciphertext, in order);
//   * sample_poly_cbd (rlwe.cpp:68-104): 6 bytes of the bootstrap stream per coefficient, after the 64 seed bytes;
//   * c1 = a, c0 = -(a s + t NTT(e)) + NTT(lift(plain)) mod Q_l (rlwe.cpp:358-386, encryptor.cpp:260-311).
#pragma once
#include "kernels.cuh"

namespace rsg {

__device__ __constant__ uint64_t B2_IV_D[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                               0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
__device__ __forceinline__ uint64_t rotr64(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }

// One BLAKE2b compression whose message block has only its first 8 words non-zero (all this PRNG ever hashes).
// t = bytes hashed so far including this block.
__device__ __forceinline__ void b2_compress8(uint64_t (&h)[8], const uint64_t (&m8)[8], uint64_t t, bool last) {
  uint64_t v[16], m[16];
#pragma unroll
  for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = B2_IV_D[i]; m[i] = m8[i]; m[i + 8] = 0; }
  v[12] ^= t;
  if (last) v[14] = ~v[14];
#define RSG_B2_G(a, b, c, d, x, y)                                                                                  \
  v[a] = v[a] + v[b] + (x); v[d] = rotr64(v[d] ^ v[a], 32); v[c] = v[c] + v[d]; v[b] = rotr64(v[b] ^ v[c], 24); \
  v[a] = v[a] + v[b] + (y); v[d] = rotr64(v[d] ^ v[a], 16); v[c] = v[c] + v[d]; v[b] = rotr64(v[b] ^ v[c], 63);
#pragma unroll
  for (int r = 0; r < 12; r++) {
    // sigma is a compile-time table after unrolling: the message words are picked by constant index
    constexpr uint8_t S[12][16] = {
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
    RSG_B2_G(0, 4, 8, 12, m[S[r][0]], m[S[r][1]])
    RSG_B2_G(1, 5, 9, 13, m[S[r][2]], m[S[r][3]])
    RSG_B2_G(2, 6, 10, 14, m[S[r][4]], m[S[r][5]])
    RSG_B2_G(3, 7, 11, 15, m[S[r][6]], m[S[r][7]])
    RSG_B2_G(0, 5, 10, 15, m[S[r][8]], m[S[r][9]])
    RSG_B2_G(1, 6, 11, 12, m[S[r][10]], m[S[r][11]])
    RSG_B2_G(2, 7, 8, 13, m[S[r][12]], m[S[r][13]])
    RSG_B2_G(3, 4, 9, 14, m[S[r][14]], m[S[r][15]])
  }
#undef RSG_B2_G
#pragma unroll
  for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}
// root hash of buffer `counter` of the stream keyed by `seed` (blake2xb.c: blake2xb_init_key + update + blake2b_final)
__device__ __forceinline__ void b2x_root(const uint64_t (&seed)[8], uint64_t counter, uint64_t (&root)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i++) root[i] = B2_IV_D[i];
  root[0] ^= 0x40ull | (0x40ull << 8) | (1ull << 16) | (1ull << 24);   // digest 64, key 64, fanout 1, depth 1, leaf_length 0
  root[1] ^= 4096ull << 32;                                           // node_offset 0, xof_length 4096
  b2_compress8(root, seed, 128, false);                               // the key, padded to one block
  const uint64_t m[8] = {counter, 0, 0, 0, 0, 0, 0, 0};
  b2_compress8(root, m, 136, true);
}
// 64-byte block `b` (< 64) of a buffer (blake2xb_final's counter construction)
__device__ __forceinline__ void b2x_block(const uint64_t (&root)[8], uint32_t b, uint64_t (&out)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = B2_IV_D[i];
  out[0] ^= 0x40ull | (64ull << 32);          // digest 64, key 0, fanout 0, depth 0, leaf_length 64
  out[1] ^= (uint64_t)b | (4096ull << 32);    // node_offset b, xof_length 4096
  out[2] ^= 64ull << 8;                       // node_depth 0, inner_length 64
  b2_compress8(out, root, 64, true);
}
// word `idx` of the stream (rare path: redraws of sample_poly_uniform)
__device__ uint64_t b2x_stream_word(const uint64_t (&seed)[8], uint64_t idx) {
  uint64_t root[8], blk[8];
  b2x_root(seed, idx / 512, root);
  b2x_block(root, (uint32_t)((idx % 512) / 8), blk);
  return blk[idx % 8];
}

// roots[s][c][8] for streams s < n_streams, buffers c < n_buf.  The seed of stream s is seeds[s * seed_stride .. + 8).
__global__ void __launch_bounds__(128) k_b2x_roots(const uint64_t *__restrict__ seeds, size_t seed_stride, uint32_t n_streams, uint32_t n_buf,
                                                   uint64_t *__restrict__ roots) {
  const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (size_t)n_streams * n_buf) return;
  const uint32_t s = (uint32_t)(id / n_buf), cbuf = (uint32_t)(id % n_buf);
  uint64_t seed[8], root[8];
#pragma unroll
  for (int i = 0; i < 8; i++) seed[i] = seeds[(size_t)s * seed_stride + i];
  b2x_root(seed, cbuf, root);
#pragma unroll
  for (int i = 0; i < 8; i++) roots[id * 8 + i] = root[i];
}
// out[s * out_stride + c * 512 + b * 8 + w]: one thread per 64-byte block
__global__ void __launch_bounds__(128) k_b2x_blocks(const uint64_t *__restrict__ roots, uint32_t n_streams, uint32_t n_buf,
                                                    uint64_t *__restrict__ out, size_t out_stride) {
  const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (size_t)n_streams * n_buf * 64) return;
  const size_t rb = id / 64;
  const uint32_t b = (uint32_t)(id % 64), s = (uint32_t)(rb / n_buf), cbuf = (uint32_t)(rb % n_buf);
  uint64_t root[8], blk[8];
#pragma unroll
  for (int i = 0; i < 8; i++) root[i] = roots[rb * 8 + i];
  b2x_block(root, b, blk);
  uint64_t *dst = out + (size_t)s * out_stride + (size_t)cbuf * 512 + (size_t)b * 8;
  *reinterpret_cast<ulonglong2 *>(dst) = make_ulonglong2(blk[0], blk[1]);
  *reinterpret_cast<ulonglong2 *>(dst + 2) = make_ulonglong2(blk[2], blk[3]);
  *reinterpret_cast<ulonglong2 *>(dst + 4) = make_ulonglong2(blk[4], blk[5]);
  *reinterpret_cast<ulonglong2 *>(dst + 6) = make_ulonglong2(blk[6], blk[7]);
}

// sample_poly_uniform's second half, in place on the bulk words of one ciphertext's c1 ([L_E][N_E], stream s = blockIdx.x):
// redraw the rejected words in order, then reduce every word mod Q_l.  pub_seeds[s * seed_stride ..]: the ciphertext
// stream's seed.  256 threads.
constexpr int ENC_MAX_REJ = 256;
__global__ void __launch_bounds__(256) k_enc_uniform_fix(const DevParams *__restrict__ P, uint64_t *__restrict__ arena_first, size_t enc_first,
                                                         const uint64_t *__restrict__ pub_seeds, size_t seed_stride, uint32_t *__restrict__ err,
                                                         const uint64_t *__restrict__ bulk, size_t bulk_stride) {
  __shared__ uint32_t n_rej;
  __shared__ uint32_t pos[ENC_MAX_REJ];
  const uint32_t s = blockIdx.x, L_E = P->L_E, N_E = P->N_E, L_R = P->L_R;
  const size_t ct_words = 2 * (size_t)L_E * N_E;
  uint64_t *c1 = arena_first + (enc_first * L_R + s) * ct_words + (size_t)L_E * N_E;   // stream s = (element, ring limb)
  const uint32_t words = L_E * N_E;
  // bulk (nullable): the stream's first `words` words when they were generated aside (L_E * N_E not a whole number of
  // 512-word PRNG buffers); otherwise they already sit in c1
  if (bulk) {
    const uint64_t *src = bulk + (size_t)s * bulk_stride;
    for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) c1[w] = src[w];
  }
  if (threadIdx.x == 0) n_rej = 0;
  __syncthreads();
  for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) {
    const uint64_t p = P->Q[w / N_E].p;
    const uint64_t max_multiple = 0xFFFFFFFFFFFFFFFFull - (0xFFFFFFFFFFFFFFFFull % p) - 1;
    if (c1[w] >= max_multiple) {
      const uint32_t k = atomicAdd(&n_rej, 1u);
      if (k < ENC_MAX_REJ) pos[k] = w;
    }
  }
  __syncthreads();
  if (n_rej && threadIdx.x == 0) {
    if (n_rej > ENC_MAX_REJ) {
      atomicExch(err, 1u);   // cannot happen for primes the context accepts (< 2^61: rejection probability < 1/8 needs > 2^11 of 2^17)
    } else {
      const uint32_t k = n_rej;
      for (uint32_t a = 1; a < k; a++) {   // ascending positions = SEAL's processing order
        const uint32_t x = pos[a];
        uint32_t b = a;
        while (b > 0 && pos[b - 1] > x) { pos[b] = pos[b - 1]; b--; }
        pos[b] = x;
      }
      uint64_t seed[8];
      for (int i = 0; i < 8; i++) seed[i] = pub_seeds[(size_t)s * seed_stride + i];
      uint64_t next = words;   // the redraws follow the bulk in the stream
      for (uint32_t a = 0; a < k; a++) {
        const uint64_t p = P->Q[pos[a] / N_E].p;
        const uint64_t max_multiple = 0xFFFFFFFFFFFFFFFFull - (0xFFFFFFFFFFFFFFFFull % p) - 1;
        uint64_t r;
        do r = b2x_stream_word(seed, next++); while (r >= max_multiple);
        c1[pos[a]] = r;
      }
    }
  }
  __syncthreads();
  for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) c1[w] = reduce64(c1[w], P->Q[w / N_E]);
}

// e_l[i] of every stream from the bootstrap bytes: noise[l][s][i] = cbd(bytes 64 + 6 i ..) mod Q_l.  grid (N_E / 256, streams).
__global__ void __launch_bounds__(256) k_enc_noise(const DevParams *__restrict__ P, const uint64_t *__restrict__ boot, size_t boot_stride,
                                                   uint32_t n_streams, uint64_t *__restrict__ noise) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y, N_E = P->N_E, L_E = P->L_E;
  if (i >= N_E) return;
  const uint8_t *x = reinterpret_cast<const uint8_t *>(boot + (size_t)s * boot_stride) + 64 + 6 * (size_t)i;
  const int v = __popc(x[0]) + __popc(x[1]) + __popc(x[2] & 0x1F) - __popc(x[3]) - __popc(x[4]) - __popc(x[5] & 0x1F);
  for (uint32_t l = 0; l < L_E; l++) noise[((size_t)l * n_streams + s) * N_E + i] = v < 0 ? P->Q[l].p - (uint64_t)(-v) : (uint64_t)v;
}

// c0 = -(a s + (t mod Q_l) e^) + m^  with a = c1 (already reduced), e^ = NTT(e) [l][s][i], m^ = pntt [s][l][i].
// grid (N_E / 256, L_E, streams).
__global__ void __launch_bounds__(256) k_enc_finish(const DevParams *__restrict__ P, uint64_t *__restrict__ arena_first, size_t enc_first,
                                                    const uint64_t *__restrict__ sk, const uint64_t *__restrict__ noise_ntt,
                                                    const uint64_t *__restrict__ pntt, uint32_t n_streams) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y, s = blockIdx.z;
  const uint32_t N_E = P->N_E, L_E = P->L_E, L_R = P->L_R, j = s % L_R;
  if (i >= N_E) return;
  const ModConst m = P->Q[l];
  const size_t ct_words = 2 * (size_t)L_E * N_E;
  uint64_t *ct = arena_first + (enc_first * L_R + s) * ct_words;
  const uint64_t a = ct[(size_t)(L_E + l) * N_E + i];
  const uint64_t as = mul_mod(a, sk[((size_t)j * L_E + l) * N_E + i], m);
  const uint64_t te = mul_mod(noise_ntt[((size_t)l * n_streams + s) * N_E + i], P->tmodQ[j][l], m);
  const uint64_t v = neg_mod(add_mod(as, te, m.p), m.p);
  ct[(size_t)l * N_E + i] = add_mod(v, pntt[((size_t)s * L_E + l) * N_E + i], m.p);
}

}  // namespace rsg
