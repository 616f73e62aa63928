// Test-case catalogue shared by oracle/ref_harness.cpp (reference on CPU) and
// oracle/dropin_harness.cpp (reference templates over the B200 backend).  TEST INFRASTRUCTURE.
//
// A case = (ring parameters, encoding parameters, circuit shape).  The named cases restate
// C1 = examples/example_SEAL.cpp:15-54, C3' = benchmarks/bench_mul_SEAL.cpp
// restated with N_E = 2 N_R, C4 = benchmarks/bench_logistic_regression_inference.cpp:20-27 (shape only),
// plus tiny parameter sets (sec_level none) whose dumps are small enough to commit under tests/golden/.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "seal/seal.h"

namespace cases {

struct CaseSpec {
  std::string name;
  size_t N_R = 0;               // ring degree
  std::vector<int> ring_bits;   // bit sizes handed to CoeffModulus::Create(N_E, .) -- incl. the special prime SEAL drops
  size_t N_E = 0;               // encoding degree
  std::vector<int> enc_bits;    // empty => CoeffModulus::BFVDefault(N_E), as EncodingElem::set_context does
  bool sec_none = false;        // tiny parameter sets need sec_level_type::none
  // circuit
  size_t n = 0, io = 0, aux = 0;
  bool use_const = false;       // constraints that touch the constant wire (reference verifier then rejects)
  bool quirks = false;          // scalar 0 / 1 / k inputs and a zero-prefix element among the assignment
  bool ntt_demo = false;        // ONE constraint (x_1 + sum_i row^i x_(i+1)) * 1 = x_io with RING-ELEMENT coefficients, no
                                // auxiliary input: the shape of benchmarks/bench_ntt_SEAL.cpp:29-55 (SURVEY 8(d) C2')
};

inline CaseSpec get_case(const std::string &name) {
  CaseSpec c;
  c.name = name;
  if (name == "tiny_fast") {            // t (25 bit) < every Q_l (30 bit): SEAL's fast plain lift
    c.N_R = 128; c.ring_bits = {25, 25, 25}; c.N_E = 256; c.enc_bits = {30, 30, 30, 30}; c.sec_none = true;
    c.n = 5; c.io = 4; c.aux = 5;
  } else if (name == "tiny_slow") {     // t (40 bit) > Q_l (30 bit): slow multi-word lift
    c.N_R = 128; c.ring_bits = {40}; c.N_E = 256; c.enc_bits = {30, 30, 30, 30, 30, 30}; c.sec_none = true;
    c.n = 7; c.io = 3; c.aux = 9;
  } else if (name == "tiny_quirks") {   // scalar / zero / one / zero-prefix coefficients + constant wire
    c.N_R = 128; c.ring_bits = {25, 25, 25}; c.N_E = 256; c.enc_bits = {30, 30, 30, 30}; c.sec_none = true;
    c.n = 6; c.io = 2; c.aux = 9; c.use_const = true; c.quirks = true;
  } else if (name == "tiny_full") {     // N_R == N_E / 1 (every slot used), one limb
    c.N_R = 256; c.ring_bits = {26}; c.N_E = 256; c.enc_bits = {31, 31, 31, 31}; c.sec_none = true;
    c.n = 3; c.io = 2; c.aux = 4;
  } else if (name == "c1") {            // examples/example_SEAL.cpp
    c.N_R = 4096; c.ring_bits = {36, 36, 37}; c.N_E = 8192; c.n = 2; c.io = 5; c.aux = 1;
  } else if (name == "c2p") {           // bench_ntt_SEAL.cpp shape, shortened (the reference uses io = N_R + 1 = 4097).
    // NOT part of any test: with n = 1 and no auxiliary input the proof's C is an EMPTY encoding, and groth16::prover copies
    // it into the proof -- ringsnark/seal/seal_ring.hpp:246 asserts !other.ciphertexts.empty(), so the reference built with
    // assertions (as oracle/Makefile.ref builds it) aborts on this shape
    c.N_R = 4096; c.ring_bits = {36, 36, 37}; c.N_E = 8192; c.n = 1; c.io = 258; c.aux = 0; c.ntt_demo = true; c.use_const = true;
  } else if (name == "c3p") {           // bench_mul circuit shape, N_R = 8192, N_E = 16384
    c.N_R = 8192; c.ring_bits = {43, 43, 44, 44, 44}; c.N_E = 16384; c.n = 4; c.io = 7; c.aux = 1;
  } else if (name == "c4s") {           // C4 parameters, small circuit (bench_plaintext_check size)
    c.N_R = 2048; c.ring_bits = {54}; c.N_E = 16384; c.n = 33; c.io = 17; c.aux = 48;
  } else if (name == "c4m") {           // C4 parameters, medium circuit
    c.N_R = 2048; c.ring_bits = {54}; c.N_E = 16384; c.n = 129; c.io = 65; c.aux = 192;
  } else if (name == "c4") {            // logistic-regression shape
    c.N_R = 2048; c.ring_bits = {54}; c.N_E = 16384; c.n = 1031; c.io = 517; c.aux = 1538; c.use_const = true;
  } else {
    throw std::invalid_argument("unknown case " + name);
  }
  return c;
}

// Ring context: mirrors how every reference driver builds it (examples/example_SEAL.cpp:15-24).
inline seal::SEALContext make_ring_context(const CaseSpec &c) {
  seal::EncryptionParameters p(seal::scheme_type::bgv);
  p.set_poly_modulus_degree(c.N_R);
  p.set_coeff_modulus(seal::CoeffModulus::Create(c.N_E, c.ring_bits));
  p.set_plain_modulus(seal::PlainModulus::Batching(c.N_R, 20));
  return seal::SEALContext(p, true, c.sec_none ? seal::sec_level_type::none : seal::sec_level_type::tc128);
}

// Encoding contexts: mirrors EncodingElem::set_context (ringsnark/seal/seal_ring.hpp:266-306) but with a
// seeded PRNG factory per context so that keygen/encrypt (the CRS) are reproducible from one seed.
inline std::vector<seal::SEALContext> make_enc_contexts(const CaseSpec &c, const seal::SEALContext &ring, uint64_t seed) {
  auto ring_parms = ring.first_context_data()->parms();
  auto coeff_modulus = c.enc_bits.empty() ? seal::CoeffModulus::BFVDefault(c.N_E)
                                          : seal::CoeffModulus::Create(c.N_E, c.enc_bits);
  std::vector<seal::SEALContext> out;
  for (size_t j = 0; j < ring_parms.coeff_modulus().size(); j++) {
    seal::EncryptionParameters p(seal::scheme_type::bgv);
    p.set_poly_modulus_degree(c.N_E);
    p.set_plain_modulus(ring_parms.coeff_modulus()[j].value());
    p.set_coeff_modulus(coeff_modulus);
    p.set_random_generator(std::make_shared<seal::Blake2xbPRNGFactory>(seal::prng_seed_type{seed, j + 1, 0xB200, 0, 0, 0, 0, 0}));
    seal::SEALContext ctx(p, true, c.sec_none ? seal::sec_level_type::none : seal::sec_level_type::tc128);
    if (ctx.first_context_data()->qualifiers().parameter_error != seal::EncryptionParameterQualifiers::error_type::success)
      throw std::invalid_argument(std::string("encoding context: ") + ctx.first_context_data()->qualifiers().parameter_error_message());
    out.push_back(ctx);
  }
  return out;
}

// xorshift-style deterministic stream for circuit wiring (NOT for ring elements).
struct Wiring {
  uint64_t s;
  explicit Wiring(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
  uint64_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
  size_t below(size_t m) { return (size_t)(next() % m); }
};

// Synthetic satisfiable R1CS of the requested shape over any RingT with the reference's concept.
//   variables 1..nfree are free (random), every constraint i defines variable nfree+1+i:
//     (x_a [+ x_b] [+ k])  *  (x_c [+ 2 x_d])  =  x_out
//   the first `io` variables are the primary input.  The constant wire appears only if use_const.
template <typename R, typename CS, typename Constraint, typename LC, typename Var>
void build_circuit(const CaseSpec &c, uint64_t seed, CS &cs, std::vector<R> &assignment,
                   R (*make_elem)(int kind)) {  // kind 0: uniform random element; 1: zero-prefix element (is_zero quirk)
  const size_t nv = c.io + c.aux;
  if (c.ntt_demo) {
    // sum = x_1 + row x_2 + row^2 x_3 + ... ; constraint sum * 1 = x_io (bench_ntt_SEAL.cpp:47-55): polynomial coefficients,
    // the constant wire on the B side, n = 1 (domain {0}, Z = x, H = 0), no auxiliary input
    if (c.n != 1 || c.aux != 0 || c.io < 3) throw std::invalid_argument("ntt_demo needs n = 1, aux = 0, io >= 3");
    assignment.assign(nv, R(0));
    for (size_t v = 0; v + 1 < c.io; v++) assignment[v] = make_elem(0);
    cs.primary_input_size = c.io;
    cs.auxiliary_input_size = 0;
    const R rs = make_elem(0);
    R row(rs);
    LC sum = LC(Var(1));
    R val = assignment[0];
    for (size_t i = 1; i + 1 < c.io; i++) {
      sum = sum + Var(i + 1) * row;
      val += row * assignment[i];
      row *= rs;
    }
    cs.add_constraint(Constraint(sum, LC((long)1), LC(Var(c.io))));
    assignment[c.io - 1] = val;
    return;
  }
  if (nv < c.n + 1) throw std::invalid_argument("case has too few variables");
  const size_t nfree = nv - c.n;
  Wiring w(seed);
  assignment.assign(nv, R(0));
  for (size_t v = 0; v < nfree; v++) assignment[v] = make_elem(0);
  if (c.quirks) {
    // scalar inputs of each kind the reference special-cases (seal_ring.tcc:509-548) and an element whose
    // first N_R*L_R/8 words are zero but which is not zero (poly_arith.cpp:147-153 treats it as zero)
    if (nfree < 5) throw std::invalid_argument("quirks case needs >= 5 free variables");
    assignment[0] = R(1);
    assignment[1] = R(0);
    assignment[2] = R(7);
    assignment[3] = make_elem(1);
    assignment[4] = R(1);
  }
  cs.primary_input_size = c.io;
  cs.auxiliary_input_size = c.aux;
  for (size_t i = 0; i < c.n; i++) {
    const size_t avail = nfree + i;  // variables defined so far (0-based count)
    size_t a = w.below(avail), b = w.below(avail), cc = w.below(avail), d = w.below(avail);
    const bool two_a = (w.next() & 1), two_b = (w.next() & 3) == 0, konst = c.use_const && (w.next() % 3 == 0);
    LC la = LC(Var(a + 1));
    R va = assignment[a];
    if (two_a) { la = la + LC(Var(b + 1)); va += assignment[b]; }
    if (konst) { la = la + LC((long)3); va += R(3); }
    LC lb = LC(Var(cc + 1));
    R vb = assignment[cc];
    if (two_b) { lb = lb + Var(d + 1) * (long)2; vb += assignment[d] * R(2); }
    const size_t out = nfree + i;
    cs.add_constraint(Constraint(la, lb, LC(Var(out + 1))));
    assignment[out] = va * vb;
  }
}

}  // namespace cases
