// oracle/ringelem_check.cpp -- TEST INFRASTRUCTURE (built into oracle/_ref/, needs /root/reference at build time).
// ringsnark::seal_gpu::RingElem's own host arithmetic against the reference's ringsnark::seal::RingElem
// (ringsnark/seal/seal_ring.tcc:5-302 over depends/SEAL-Polytools/src/poly_arith.cpp:147-350): every operator on every pairing
// of scalar / polynomial operands incl. the edge values the variant rules branch on (0, 1, values at and above q_1's bit length,
// unreduced scalars, zero-prefix and non-invertible polynomials) -- same variant, same words, same is_zero / == / hash.
// No device is touched: runs in the CPU test tier.
#include <cstdio>
#include <functional>
#include <ringsnark/seal_gpu/seal_ring.hpp>

typedef ringsnark::seal::RingElem R;
typedef ringsnark::seal_gpu::RingElem G;

struct RAccess : R { static void seed(uint64_t s) { prng = seal::Blake2xbPRNGFactory(seal::prng_seed_type{s, 1, 2, 3, 0, 0, 0, 0}).create(); } };

static std::vector<uint64_t> words_of(const R &r) {
  R t(r);
  t.to_poly_inplace();
  auto &p = t.get_poly();
  std::vector<uint64_t> w;
  for (size_t j = 0; j < p.get_coeff_modulus_count(); j++) {
    auto limb = p.get_limb(j);
    w.insert(w.end(), limb.begin(), limb.end());
  }
  return w;
}
static long checks = 0, bad = 0;
static void same(const char *what, const R &r, const G &g, int ia, int ib) {
  checks++;
  bool ok = r.is_scalar() == g.is_scalar();
  if (ok && r.is_scalar()) ok = r.get_scalar() == g.get_scalar();
  if (ok && !r.is_scalar()) ok = words_of(r) == g.words();
  ok = ok && r.is_zero() == g.is_zero() && r.hash() == g.hash() && r.size_in_bits() == g.size_in_bits() && r.fast_is_zero() == g.fast_is_zero();
  if (!ok) {
    bad++;
    if (bad < 20) fprintf(stderr, "MISMATCH %s operands %d,%d: ref %s gpu %s\n", what, ia, ib, r.is_scalar() ? "scalar" : "poly", g.is_scalar() ? "scalar" : "poly");
  }
}

int main() {
  seal::EncryptionParameters parms(seal::scheme_type::bgv);
  const size_t N = 1024;
  parms.set_poly_modulus_degree(N);
  parms.set_coeff_modulus(seal::CoeffModulus::Create(N, {30, 31}));   // two ring limbs; first level keeps one... use both via BFV-style
  parms.set_plain_modulus(seal::PlainModulus::Batching(N, 20));
  seal::SEALContext ctx(parms);
  G::set_context(ctx);   // also sets the reference's
  const auto fp = ctx.first_context_data()->parms();
  const size_t L = fp.coeff_modulus().size();
  const uint64_t q1 = fp.coeff_modulus()[0].value();
  RAccess::seed(7);
  std::vector<R> ops;
  const std::vector<uint64_t> scalars = {0ull, 1ull, 2ull, 5ull, 1000003ull, (1ull << 14) + 1, (1ull << 15), (1ull << 29) - 1, q1 - 1, q1, q1 + 5,
                                         2 * q1 + 3, (1ull << 40) + 7, ~0ull};
  // inversion of an unreduced scalar (>= q_j) overflows SEAL's xgcd inside the reference's noexcept is_invertible(): not compared
  auto invertible_domain = [&](const R &r) { return !r.is_scalar() || r.get_scalar() < fp.coeff_modulus()[L - 1].value(); };
  for (uint64_t s : scalars) ops.push_back(R(s));
  for (int k = 0; k < 3; k++) ops.push_back(R::random_element());
  {   // zero-prefix polynomial (SealPoly::is_zero says zero), an all-zero polynomial, a polynomial with a non-invertible slot
    R r = R::random_element();
    auto w = words_of(r);
    for (size_t i = 0; i < w.size() / 8 + 1; i++) w[i] = 0;
    ops.push_back(R(polytools::SealPoly(ctx, w, &ctx.first_parms_id())));
    std::fill(w.begin(), w.end(), 0);
    ops.push_back(R(polytools::SealPoly(ctx, w, &ctx.first_parms_id())));
    w = words_of(R::random_element());
    w[N * L - 3] = 0;
    ops.push_back(R(polytools::SealPoly(ctx, w, &ctx.first_parms_id())));
    w = words_of(ops[14]);
    for (size_t i = w.size() / 8; i < w.size(); i++) w[i] ^= (i * 2654435761u) & 0xFFFF;   // equal to ops[14] on the compared prefix only
    ops.push_back(R(polytools::SealPoly(ctx, w, &ctx.first_parms_id())));
  }
  std::vector<G> gops;
  for (const auto &r : ops) gops.push_back(G(r));
  const int n = (int)ops.size();
  for (int a = 0; a < n; a++) {
    same("copy", ops[a], gops[a], a, -1);
    {
      R r(ops[a]); G g(gops[a]);
      r.negate_inplace(); g.negate_inplace();
      same("negate", r, g, a, -1);
      R r2 = ops[a].to_poly(); G g2 = gops[a].to_poly();
      same("to_poly", r2, g2, a, -1);
      if (invertible_domain(ops[a])) {
        checks++;
        if (ops[a].is_invertible() != gops[a].is_invertible()) { bad++; fprintf(stderr, "MISMATCH is_invertible %d\n", a); }
        bool tr = false, tg = false;
        R ri; G gi;
        try { ri = ops[a].inverse(); } catch (const std::invalid_argument &) { tr = true; }
        try { gi = gops[a].inverse(); } catch (const std::invalid_argument &) { tg = true; }
        checks++;
        if (tr != tg) { bad++; fprintf(stderr, "MISMATCH inverse throws %d\n", a); }
        if (!tr && !tg) same("inverse", ri, gi, a, -1);
      }
    }
    for (int b = 0; b < n; b++) {
      { R r(ops[a]); G g(gops[a]); r += ops[b]; g += gops[b]; same("+=", r, g, a, b); }
      { R r(ops[a]); G g(gops[a]); r -= ops[b]; g -= gops[b]; same("-=", r, g, a, b); }
      { R r(ops[a]); G g(gops[a]); r *= ops[b]; g *= gops[b]; same("*=", r, g, a, b); }
      checks++;
      if ((ops[a] == ops[b]) != (gops[a] == gops[b])) { bad++; fprintf(stderr, "MISMATCH == %d,%d\n", a, b); }
      if (!invertible_domain(ops[b])) continue;
      bool tr = false, tg = false;
      R rd; G gd;
      try { rd = ops[a] / ops[b]; } catch (const std::invalid_argument &) { tr = true; }
      try { gd = gops[a] / gops[b]; } catch (const std::invalid_argument &) { tg = true; }
      checks++;
      if (tr != tg) { bad++; fprintf(stderr, "MISMATCH / throws %d,%d\n", a, b); }
      if (!tr && !tg) same("/", rd, gd, a, b);
    }
  }
  // chains: values that grow through the scalar -> polynomial promotion
  {
    R r(3); G g(3);
    for (int i = 0; i < 40; i++) { r *= R(7); g *= G(7); r += R(i); g += G(i); same("chain", r, g, i, -1); }
    for (int i = 0; i < 10; i++) { r -= R(5); g -= G(5); r = -r; g = -g; same("chain2", r, g, i, -1); }
  }
  printf("{\"checks\": %ld, \"mismatches\": %ld}\n", checks, bad);
  return bad ? 1 : 0;
}
