// tools/host_poly_check.cpp -- CPU check of ringsnark_b200/csrc/host_poly.hpp: every fast routine against its naive form, on a
// 20-bit and a 60-bit prime.  Built and run by tests/test_host_poly.py:  g++ -O2 -std=c++17 -o <bin> tools/host_poly_check.cpp
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../ringsnark_b200/csrc/host_poly.hpp"
using namespace rsg_host;

static bool is_prime(uint64_t n) {
  if (n < 2) return false;
  for (uint64_t q : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
    if (n % q == 0) return n == q;
  }
  uint64_t d = n - 1;
  int r = 0;
  while (!(d & 1)) { d >>= 1; r++; }
  for (uint64_t a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
    uint64_t x = powmod(a, d, n);
    if (x == 1 || x == n - 1) continue;
    bool comp = true;
    for (int i = 1; i < r && comp; i++) {
      x = mulmod(x, x, n);
      if (x == n - 1) comp = false;
    }
    if (comp) return false;
  }
  return true;
}
static unsigned bitrev(unsigned x, int bits) {
  unsigned r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
}
static NttTables make_tables(uint64_t p, int logN) {
  const size_t N = (size_t)1 << logN;
  uint64_t psi = 0;
  for (uint64_t g = 2; g < 1000 && !psi; g++) {
    const uint64_t r = powmod(g, (p - 1) / (2 * N), p);
    if (powmod(r, N, p) == p - 1) psi = r;
  }
  Poly tw(N, 1);
  uint64_t cur = 1;
  for (size_t i = 0; i < N; i++) { tw[bitrev((unsigned)i, logN)] = cur; cur = mulmod(cur, psi, p); }
  NttTables t;
  t.set(p, tw);
  return t;
}
#define CHECK(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static int run(uint64_t p, int logN) {
  const NttTables t = make_tables(p, logN);
  const size_t N = (size_t)1 << logN;
  std::mt19937_64 rng(p);
  auto rnd = [&](size_t n) { Poly a(n); for (auto &x : a) x = rng() % p; return a; };
  for (size_t k = 1; k < N; k++) CHECK(mulmod(t.tw[k], t.itw[k], p) == 1);
  {   // transform round trip and the negacyclic product
    Poly a = rnd(N), b = a;
    ntt_fwd(b, logN, t);
    ntt_inv(b, logN, t);
    CHECK(a == b);
  }
  const size_t sizes[][2] = {{1, 1}, {3, 700}, {300, 300}, {N / 2, N / 2}, {N / 2 + 1, N / 2}, {N, N}, {N + 1, N + 1}, {3 * N / 2 + 5, 2 * N - 3},
                             {2 * N + 1, 37}, {2 * N + 1, 2 * N + 1}};
  for (auto &s : sizes) {
    Poly a = rnd(s[0]), b = rnd(s[1]);
    CHECK(polymul(a, b, t, 64) == polymul_school(a, b, p));
  }
  for (uint64_t n : {1ull, 2ull, 33ull, 100ull, (unsigned long long)N, (unsigned long long)(2 * N), (unsigned long long)(2 * N + 77)}) {
    Poly z{1};
    for (uint64_t x = 0; x < n; x++) z = polymul_school(z, Poly{(p - x % p) % p, 1}, p);
    const Poly zf = node_product(0, n, t);
    CHECK(z == zf);
    CHECK(node_product(5, 5 + n, t).size() == n + 1);
    if (n >= 2) {
      Poly rz(n + 1);
      for (uint64_t i = 0; i <= n; i++) rz[i] = z[n - i];
      const size_t m = n - 1;
      Poly u1 = series_inverse(rz, m, t), u2 = series_inverse_naive(rz, m, p);
      u2.resize(std::max<size_t>(m, 1));
      u1.resize(std::max<size_t>(m, 1), 0);
      CHECK(u1 == u2);
    }
  }
  return 0;
}

int main() {
  if (run(786433, 10)) return 1;                       // 3 * 2^18 + 1
  uint64_t p = ((uint64_t)1 << 60) + 1;
  while (!is_prime(p)) p += (uint64_t)1 << 12;         // = 1 mod 2^12: negacyclic transforms up to size 2^11
  if (run(p, 9)) return 1;
  printf("ok\n");
  return 0;
}
