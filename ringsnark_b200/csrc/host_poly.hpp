// Host-side polynomial arithmetic over Z_p (p < 2^62) for the per-(n, q_j) constants of the witness map: the vanishing
// polynomial Z = prod_{x<n} (X - x) (util/evaluation_domain.tcc:53-84 gives the domain, polynomials.tcc:45-59 the product),
// rev(Z)^-1 mod x^(n-1) (what the long division by Z of polynomials.tcc:70-81 multiplies by), and the subproduct tree of the
// nodes.  Small operands use the schoolbook product; large ones are cut into blocks of TS/2 coefficients and multiplied
// through negacyclic transforms of size TS with the device's own twiddle order, so n = 2^16 costs seconds, not minutes.
// Every result is the canonical residue of the exact polynomial, so which product ran cannot be seen in the tables.
// Plain C++ (no CUDA): tests/test_host_poly.py compiles it with g++ and checks each routine against its naive form.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace rsg_host {

typedef unsigned __int128 u128;
typedef std::vector<uint64_t> Poly;

inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t p) { return (uint64_t)(((u128)a * b) % p); }
inline uint64_t powmod(uint64_t a, uint64_t e, uint64_t p) {
  uint64_t r = 1;
  a %= p;
  while (e) {
    if (e & 1) r = mulmod(r, a, p);
    a = mulmod(a, a, p);
    e >>= 1;
  }
  return r;
}
inline uint64_t invmod(uint64_t a, uint64_t p) { return powmod(a, p - 2, p); }

// Transform tables of one prime: tw[k] = psi^bitrev(k) as the device holds them (size-N table; the table of a smaller size is
// its prefix), itw[k] = tw[k]^-1.
struct NttTables {
  uint64_t p = 0;
  int logN = 0;
  Poly tw, itw;
  void set(uint64_t prime, const Poly &forward) {
    p = prime;
    tw = forward;
    logN = 0;
    while (((size_t)1 << logN) < tw.size()) logN++;
    // batch inversion: one modular inverse for the whole table
    const size_t n = tw.size();
    itw.assign(n, 1);
    Poly pre(n, 1);
    uint64_t acc = 1;
    for (size_t i = 1; i < n; i++) { pre[i] = acc; acc = mulmod(acc, tw[i], p); }
    uint64_t inv = invmod(acc, p);
    for (size_t i = n; i-- > 1;) { itw[i] = mulmod(inv, pre[i], p); inv = mulmod(inv, tw[i], p); }
  }
};

// forward negacyclic transform of size 2^lg: natural order in, bit-reversed order out (the device's wf_ntt_fwd)
inline void ntt_fwd(Poly &a, int lg, const NttTables &t) {
  const size_t n = (size_t)1 << lg;
  const uint64_t p = t.p;
  for (int s = 0; s < lg; s++) {
    const size_t gap = n >> (s + 1);
    for (size_t blk = 0; blk < ((size_t)1 << s); blk++) {
      const uint64_t w = t.tw[((size_t)1 << s) + blk];
      for (size_t o = 0; o < gap; o++) {
        const size_t i = blk * 2 * gap + o;
        const uint64_t x = a[i], y = mulmod(a[i + gap], w, p);
        a[i] = x + y >= p ? x + y - p : x + y;
        a[i + gap] = x >= y ? x - y : x + p - y;
      }
    }
  }
}
// its inverse, scaled: bit-reversed in, natural out
inline void ntt_inv(Poly &a, int lg, const NttTables &t) {
  const size_t n = (size_t)1 << lg;
  const uint64_t p = t.p;
  for (int s = lg - 1; s >= 0; s--) {
    const size_t gap = n >> (s + 1);
    for (size_t blk = 0; blk < ((size_t)1 << s); blk++) {
      const uint64_t w = t.itw[((size_t)1 << s) + blk];
      for (size_t o = 0; o < gap; o++) {
        const size_t i = blk * 2 * gap + o;
        const uint64_t x = a[i], y = a[i + gap];
        a[i] = x + y >= p ? x + y - p : x + y;
        a[i + gap] = mulmod(x >= y ? x - y : x + p - y, w, p);
      }
    }
  }
  const uint64_t ninv = invmod(n % p, p);
  for (size_t i = 0; i < n; i++) a[i] = mulmod(a[i], ninv, p);
}

inline Poly polymul_school(const Poly &a, const Poly &b, uint64_t p) {
  if (a.empty() || b.empty()) return Poly();
  Poly r(a.size() + b.size() - 1, 0);
  for (size_t i = 0; i < a.size(); i++) {
    if (!a[i]) continue;
    for (size_t k = 0; k < b.size(); k++) r[i + k] = (uint64_t)(((u128)a[i] * b[k] + r[i + k]) % p);
  }
  return r;
}

// transformed copies of the h-coefficient blocks of a (zero-padded to TS = 2h each)
inline std::vector<Poly> blocks_fwd(const Poly &a, size_t h, int lgT, const NttTables &t) {
  const size_t nb = (a.size() + h - 1) / h;
  std::vector<Poly> out(nb, Poly(2 * h, 0));
  for (size_t j = 0; j < nb; j++) {
    for (size_t i = 0; i < h && j * h + i < a.size(); i++) out[j][i] = a[j * h + i];
    ntt_fwd(out[j], lgT, t);
  }
  return out;
}

// a * b (all a.size() + b.size() - 1 coefficients); `school_below`: operand-size product under which the schoolbook loop runs
inline Poly polymul(const Poly &a, const Poly &b, const NttTables &t, size_t school_below = (size_t)1 << 16) {
  if (a.empty() || b.empty()) return Poly();
  const uint64_t p = t.p;
  if (a.size() * b.size() <= school_below || t.logN < 2) return polymul_school(a, b, p);
  // transform size: the smallest that holds the whole product, capped at the table size (then blocks)
  const size_t rs = a.size() + b.size() - 1;
  int lgT = 2;
  while (lgT < t.logN && ((size_t)1 << lgT) < rs) lgT++;
  const size_t TS = (size_t)1 << lgT;
  if (TS >= rs) {   // one transform serves
    Poly x(TS, 0), y(TS, 0);
    std::copy(a.begin(), a.end(), x.begin());
    std::copy(b.begin(), b.end(), y.begin());
    ntt_fwd(x, lgT, t);
    ntt_fwd(y, lgT, t);
    for (size_t i = 0; i < TS; i++) x[i] = mulmod(x[i], y[i], p);
    ntt_inv(x, lgT, t);
    x.resize(rs);
    return x;
  }
  const size_t h = TS / 2;
  const std::vector<Poly> A = blocks_fwd(a, h, lgT, t), B = blocks_fwd(b, h, lgT, t);
  Poly r(rs, 0), T(TS);
  for (size_t k = 0; k + 1 < A.size() + B.size(); k++) {
    std::fill(T.begin(), T.end(), 0);
    for (size_t i = 0; i < A.size(); i++) {
      if (k < i || k - i >= B.size()) continue;
      const Poly &x = A[i], &y = B[k - i];
      for (size_t e = 0; e < TS; e++) {
        const uint64_t v = T[e] + mulmod(x[e], y[e], p);
        T[e] = v >= p ? v - p : v;
      }
    }
    ntt_inv(T, lgT, t);
    for (size_t e = 0; e < TS && k * h + e < rs; e++) {
      const uint64_t v = r[k * h + e] + T[e];
      r[k * h + e] = v >= p ? v - p : v;
    }
  }
  return r;
}

// prod_{x in [lo, hi)} (X - x), monic, hi - lo + 1 coefficients
inline Poly node_product(uint64_t lo, uint64_t hi, const NttTables &t) {
  const uint64_t p = t.p;
  if (hi - lo <= 32) {
    Poly f{1};
    for (uint64_t x = lo; x < hi; x++) f = polymul_school(f, Poly{(p - x % p) % p, 1}, p);
    return f;
  }
  const uint64_t mid = lo + (hi - lo) / 2;
  return polymul(node_product(lo, mid, t), node_product(mid, hi, t), t);
}

// f^-1 mod x^m for f_0 = 1 (Newton iteration u <- u (2 - f u))
inline Poly series_inverse(const Poly &f, size_t m, const NttTables &t) {
  const uint64_t p = t.p;
  Poly u{1};
  if (m <= 1) return u;
  for (size_t k = 1; k < m;) {
    const size_t k2 = std::min(2 * k, m);
    Poly fk(f.begin(), f.begin() + std::min(k2, f.size()));
    Poly e = polymul(fk, u, t);
    e.resize(k2, 0);
    for (size_t i = 0; i < k2; i++) e[i] = e[i] ? p - e[i] : 0;      // -f u
    e[0] = (e[0] + 2) % p;                                            // 2 - f u
    u = polymul(u, e, t);
    u.resize(k2, 0);
    k = k2;
  }
  return u;
}
// the recurrence the small-n tables use (and what the fast one must reproduce): u_i = -sum_{1<=t<=i} f_t u_(i-t)
inline Poly series_inverse_naive(const Poly &f, size_t m, uint64_t p) {
  Poly u(std::max<size_t>(m, 1), 0);
  u[0] = 1;
  for (size_t i = 1; i < m; i++) {
    u128 acc = 0;
    for (size_t s = 1; s <= i && s < f.size(); s++) acc += (u128)mulmod(f[s], u[i - s], p);
    u[i] = (p - (uint64_t)(acc % p)) % p;
  }
  return u;
}

}  // namespace rsg_host
