// 64-bit modular arithmetic for primes p < 2^61 on the sm_100a integer pipes (32-bit IMAD/IADD3 underneath).
// Every routine returns what the reference's routine returns -- a canonical residue in [0, p) -- or says
// explicitly that it is lazy.  Restates (does not copy) the arithmetic of
//   depends/SEAL/native/src/seal/util/uintarithsmallmod.h:255-326 (Shoup operand, lazy multiply),
//   depends/SEAL/native/src/seal/util/uintarithsmallmod.h:114-190 (add/sub mod, Barrett 64/128),
//   depends/SEAL/native/src/seal/modulus.cpp:87-98 (const_ratio = floor(2^128 / p)).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rsg {

// Per-prime constants, computed on the host at context creation.
struct ModConst {
  uint64_t p;        // the prime
  uint64_t ratio0;   // floor(2^128 / p), low word
  uint64_t ratio1;   // floor(2^128 / p), high word  (= floor(2^64 / p))
  uint64_t r128;     // 2^128 mod p   (folds the third accumulator word)
};

struct Twiddle {     // Shoup pair: w and floor(w * 2^64 / p)
  uint64_t w, wq;
};

__device__ __forceinline__ uint64_t add_mod(uint64_t a, uint64_t b, uint64_t p) {
  uint64_t s = a + b;
  return s >= p ? s - p : s;
}
__device__ __forceinline__ uint64_t sub_mod(uint64_t a, uint64_t b, uint64_t p) {
  uint64_t d = a - b;
  return a >= b ? d : d + p;
}
__device__ __forceinline__ uint64_t neg_mod(uint64_t a, uint64_t p) { return a ? p - a : 0; }

// x * w mod p, lazy: result in [0, 2p) for ANY 64-bit x (Harvey / Shoup).
__device__ __forceinline__ uint64_t mul_shoup_lazy(uint64_t x, const Twiddle &t, uint64_t p) {
  uint64_t q = __umul64hi(x, t.wq);
  return x * t.w - q * p;
}
__device__ __forceinline__ uint64_t mul_shoup(uint64_t x, const Twiddle &t, uint64_t p) {
  uint64_t r = mul_shoup_lazy(x, t, p);
  return r >= p ? r - p : r;
}

// x mod p for a 64-bit x (Barrett with floor(2^64/p)); canonical.
__device__ __forceinline__ uint64_t reduce64(uint64_t x, const ModConst &m) {
  uint64_t q = __umul64hi(x, m.ratio1);
  uint64_t r = x - q * m.p;
  return r >= m.p ? r - m.p : r;
}

// (hi:lo) mod p for a 128-bit value, p < 2^61; canonical.
__device__ __forceinline__ uint64_t reduce128(uint64_t lo, uint64_t hi, const ModConst &m) {
  // q = floor((hi:lo) * ratio / 2^128), only the low word of q is needed
  uint64_t carry = __umul64hi(lo, m.ratio0);
  uint64_t t_lo = lo * m.ratio1, t_hi = __umul64hi(lo, m.ratio1);
  uint64_t s1 = t_lo + carry;
  uint64_t c1 = s1 < t_lo;
  uint64_t mid = t_hi + c1;
  uint64_t u_lo = hi * m.ratio0, u_hi = __umul64hi(hi, m.ratio0);
  uint64_t s2 = s1 + u_lo;
  uint64_t c2 = s2 < s1;
  uint64_t q = hi * m.ratio1 + mid + u_hi + c2;
  uint64_t r = lo - q * m.p;
  return r >= m.p ? r - m.p : r;
}

__device__ __forceinline__ uint64_t mul_mod(uint64_t a, uint64_t b, const ModConst &m) {
  return reduce128(a * b, __umul64hi(a, b), m);
}

// 192-bit accumulator for sums of 64x64 products: never overflows for < 2^32 terms.
struct Acc192 {
  uint64_t lo, hi;
  uint32_t top;
  __device__ __forceinline__ void clear() { lo = 0; hi = 0; top = 0; }
  __device__ __forceinline__ void mac(uint64_t a, uint64_t b) {
    asm("mad.lo.cc.u64 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u64 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+l"(lo), "+l"(hi), "+r"(top)
        : "l"(a), "l"(b));
  }
  __device__ __forceinline__ void add(uint64_t a) {
    asm("add.cc.u64 %0, %0, %3;\n\t"
        "addc.cc.u64 %1, %1, 0;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+l"(lo), "+l"(hi), "+r"(top)
        : "l"(a));
  }
  __device__ __forceinline__ void add128(uint64_t a_lo, uint64_t a_hi) {
    asm("add.cc.u64 %0, %0, %3;\n\t"
        "addc.cc.u64 %1, %1, %4;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+l"(lo), "+l"(hi), "+r"(top)
        : "l"(a_lo), "l"(a_hi));
  }
  __device__ __forceinline__ uint64_t reduce(const ModConst &m) const {
    uint64_t r = reduce128(lo, hi, m);
    if (top) {
      // top * (2^128 mod p) < 2^32 * 2^61; add the reduced low part and reduce once more
      uint64_t t_lo = (uint64_t)top * m.r128, t_hi = __umul64hi((uint64_t)top, m.r128);
      uint64_t s = t_lo + r;
      t_hi += s < t_lo;
      r = reduce128(s, t_hi, m);
    }
    return r;
  }
};

}  // namespace rsg
