// Reference harness: runs the UNMODIFIED zkFHE/ringSNARK code (headers under /root/reference, SEAL 4.1.1,
// SEAL-Polytools) on the CPU and either dumps golden vectors or times the prover.  Built by
// oracle/Makefile.ref into oracle/_ref/ref_harness.  TEST INFRASTRUCTURE: it is the parity checker and the
// timed CPU baseline ("kind": "reference"); nothing in the product path links or calls it.
//
//   ref_harness dump <case> <out.rsgv> [seed]      golden vectors for one case (see oracle/cases.hpp)
//   ref_harness time <case> <what> [args]          JSON timing line; <what> in {prover, lincomb, witness}
//   ref_harness list
#include <chrono>
#include <cstring>
#include <iostream>
#include <sstream>

#include <ringsnark/seal/seal_ring.hpp>
#include <ringsnark/seal/seal_util.hpp>
#include <ringsnark/zk_proof_systems/groth16/groth16.hpp>
#include <ringsnark/zk_proof_systems/rinocchio/rinocchio.hpp>

#include "cases.hpp"
#include "poly_arith.h"
#include "rsgv_io.hpp"
#include "seal/seal.h"
#ifdef _OPENMP
#include <omp.h>
#endif

typedef ringsnark::seal::RingElem R;
typedef ringsnark::seal::EncodingElem E;
using std::vector;

// --- access to protected state through derived classes (no reference file is modified) ---------------
struct RingAccess : R {
  static void seed(uint64_t s) {
    prng = seal::Blake2xbPRNGFactory(seal::prng_seed_type{s, 0x52494e47, 0, 0, 0, 0, 0, 0}).create();
  }
};
struct EncAccess : E {
  static const vector<seal::Ciphertext> &cts(const E &e) { return e.*(&EncAccess::ciphertexts); }
};

static seal::SEALContext *g_ring = nullptr;
static cases::CaseSpec g_case;

static R make_elem(int kind) {
  R r = R::random_element();
  if (kind == 1) {
    // zero-prefix element: first size/8 + 1 words zero, rest random (non-zero tail)
    auto &p = r.get_poly();
    size_t L = p.get_coeff_modulus_count(), N = p.get_coeff_count();
    vector<uint64_t> w(L * N);
    for (size_t j = 0; j < L; j++) {
      auto limb = p.get_limb(j);
      for (size_t i = 0; i < N; i++) w[j * N + i] = limb[i];
    }
    for (size_t i = 0; i < L * N / 8 + 1; i++) w[i] = 0;
    w[L * N - 1] |= 1;
    r = R(polytools::SealPoly(*g_ring, w, &g_ring->first_parms_id()));
  }
  return r;
}

// --- flattening helpers ---------------------------------------------------------------------------------
static size_t ring_words() {
  auto parms = g_ring->first_context_data()->parms();
  return parms.poly_modulus_degree() * parms.coeff_modulus().size();
}

// tag: 0 = scalar (value in `scalar`), 1 = poly.  Dense words are what to_poly() yields (seal_ring.tcc:265-277).
static void flatten_ring(const vector<R> &v, vector<uint64_t> &words, vector<uint64_t> &tags, vector<uint64_t> &scalars) {
  const size_t W = ring_words();
  auto parms = g_ring->first_context_data()->parms();
  const size_t L = parms.coeff_modulus().size(), N = parms.poly_modulus_degree();
  words.assign(v.size() * W, 0);
  tags.assign(v.size(), 0);
  scalars.assign(v.size(), 0);
  for (size_t e = 0; e < v.size(); e++) {
    polytools::SealPoly p(*g_ring);
    if (v[e].is_scalar()) {
      tags[e] = 0;
      scalars[e] = v[e].get_scalar();
      R tmp(v[e]);
      tmp.to_poly_inplace();
      p = tmp.get_poly();
    } else {
      tags[e] = 1;
      p = v[e].get_poly();
    }
    for (size_t j = 0; j < L; j++) {
      auto limb = p.get_limb(j);
      for (size_t i = 0; i < N; i++) words[e * W + j * N + i] = limb[i];
    }
  }
}

static void put_ring(rsgv::Writer &w, const std::string &name, const vector<R> &v) {
  vector<uint64_t> words, tags, scalars;
  flatten_ring(v, words, tags, scalars);
  w.put(name, words);
  w.put(name + ".tag", tags);
  w.put(name + ".scalar", scalars);
}

// One EncodingElem -> [L_R][2][L_E][N_E] words + per-limb ciphertext size (0 for SEAL's empty zero ciphertext).
static void flatten_enc(const E &e, size_t L_R, size_t L_E, size_t N_E, uint64_t *words, uint64_t *sizes) {
  const size_t per = 2 * L_E * N_E;
  memset(words, 0, L_R * per * 8);
  if (e.is_empty()) {
    for (size_t j = 0; j < L_R; j++) sizes[j] = ~0ull;  // whole element empty
    return;
  }
  const auto &cts = EncAccess::cts(e);
  for (size_t j = 0; j < L_R; j++) {
    sizes[j] = cts[j].size();
    if (cts[j].size() == 0) continue;
    if (cts[j].size() != 2 || cts[j].coeff_modulus_size() != L_E || cts[j].poly_modulus_degree() != N_E)
      throw std::logic_error("unexpected ciphertext shape");
    memcpy(words + j * per, cts[j].data(), per * 8);
  }
}

static void put_enc_vec(rsgv::Writer &w, const std::string &name, const vector<E> &v, size_t L_R, size_t L_E, size_t N_E) {
  const size_t per = L_R * 2 * L_E * N_E;
  vector<uint64_t> words(v.size() * per), sizes(v.size() * L_R);
  for (size_t i = 0; i < v.size(); i++) flatten_enc(v[i], L_R, L_E, N_E, words.data() + i * per, sizes.data() + i * L_R);
  w.put(name, words);
  w.put(name + ".size", sizes);
}

// --- case set-up ----------------------------------------------------------------------------------------
struct Setup {
  ringsnark::r1cs_constraint_system<R> cs;
  vector<R> assignment, primary, auxiliary;
  size_t L_R, L_E, N_E, N_R;
  vector<uint64_t> q, Q;
};

static Setup setup_case(const std::string &name, uint64_t seed) {
  g_case = cases::get_case(name);
  g_ring = new seal::SEALContext(cases::make_ring_context(g_case));
  if (!g_ring->parameters_set()) throw std::invalid_argument(std::string("ring context: ") + g_ring->parameter_error_message());
  R::set_context(*g_ring);
  E::set_contexts(cases::make_enc_contexts(g_case, *g_ring, seed));
  RingAccess::seed(seed);
  Setup s;
  cases::build_circuit<R, ringsnark::r1cs_constraint_system<R>, ringsnark::r1cs_constraint<R>,
                       ringsnark::linear_combination<R>, ringsnark::variable<R>>(g_case, seed, s.cs, s.assignment, &make_elem);
  s.primary.assign(s.assignment.begin(), s.assignment.begin() + g_case.io);
  s.auxiliary.assign(s.assignment.begin() + g_case.io, s.assignment.end());
  auto rp = g_ring->first_context_data()->parms();
  auto ep = E::get_contexts()[0].first_context_data()->parms();
  s.N_R = rp.poly_modulus_degree();
  s.L_R = rp.coeff_modulus().size();
  s.N_E = ep.poly_modulus_degree();
  s.L_E = ep.coeff_modulus().size();
  for (auto &m : rp.coeff_modulus()) s.q.push_back(m.value());
  for (auto &m : ep.coeff_modulus()) s.Q.push_back(m.value());
  return s;
}

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// --- dump -----------------------------------------------------------------------------------------------
static int cmd_dump(const std::string &name, const std::string &out, uint64_t seed) {
  Setup s = setup_case(name, seed);
  const size_t n = g_case.n;
  rsgv::Writer w;
  w.put("params", vector<uint64_t>{s.N_R, s.L_R, s.N_E, s.L_E, n, g_case.io, g_case.aux, seed,
                                   (uint64_t)g_case.use_const, (uint64_t)g_case.quirks});
  w.put("ring_q", s.q);
  w.put("enc_Q", s.Q);
  {
    vector<uint64_t> fast;
    for (auto &c : E::get_contexts()) fast.push_back(c.first_context_data()->qualifiers().using_fast_plain_lift);
    w.put("fast_plain_lift", fast);
  }
  const bool sat = s.cs.is_satisfied(s.primary, s.auxiliary);
  w.put1("r1cs_satisfied", sat);

  // (1) the nine evaluation vectors the witness map interpolates (r1cs_to_qrp.tcc:167-223)
  {
    vector<R> mid_assign(s.primary.size(), R::zero()), io_assign(s.primary);
    mid_assign.insert(mid_assign.end(), s.auxiliary.begin(), s.auxiliary.end());
    vector<R> zeros(s.auxiliary.size(), R::zero());
    io_assign.insert(io_assign.end(), zeros.begin(), zeros.end());
    const vector<R> *assigns[3] = {&mid_assign, &io_assign, &s.assignment};
    const char *names[3] = {"mid", "io", "full"};
    for (int k = 0; k < 3; k++) {
      vector<R> ya, yb, yc;
      for (size_t i = 0; i < n; i++) {
        ya.push_back(s.cs.constraints[i].a.evaluate(*assigns[k]));
        yb.push_back(s.cs.constraints[i].b.evaluate(*assigns[k]));
        yc.push_back(s.cs.constraints[i].c.evaluate(*assigns[k]));
      }
      put_ring(w, std::string("eval_A_") + names[k], ya);
      put_ring(w, std::string("eval_B_") + names[k], yb);
      put_ring(w, std::string("eval_C_") + names[k], yc);
    }
  }
  // (1b) the constraint system itself in CSR form (rows m*n + i, m in {A,B,C}; col 0 = constant wire)
  {
    vector<uint64_t> row_ptr{0}, col, coeff;
    bool scalar_coeffs = true;
    for (int m = 0; m < 3; m++)
      for (size_t i = 0; i < n; i++) {
        const auto &lc = m == 0 ? s.cs.constraints[i].a : (m == 1 ? s.cs.constraints[i].b : s.cs.constraints[i].c);
        for (const auto &lt : lc.terms) {
          if (!lt.coeff.is_scalar()) { scalar_coeffs = false; continue; }
          col.push_back(lt.index);
          coeff.push_back(lt.coeff.get_scalar());
        }
        row_ptr.push_back(col.size());
      }
    w.put1("r1cs_scalar_coeffs", scalar_coeffs);   // 0: ring-element coefficients (bench_ntt shape) -- no CSR form exists
    if (scalar_coeffs) {
      w.put("r1cs_row_ptr", row_ptr);
      w.put("r1cs_col", col);
      w.put("r1cs_coeff", coeff);
    }
  }
  put_ring(w, "primary_input", s.primary);
  put_ring(w, "auxiliary_input", s.auxiliary);

  // (2) witness map output (r1cs_to_qrp.tcc:148-259), non-ZK as groth16::prover calls it (groth16.tcc:82-84)
  double t0 = now_s();
  const auto wit = ringsnark::r1cs_to_qrp_witness_map(s.cs, s.primary, s.auxiliary, R::zero(), R::zero(), R::zero());
  double t_wit = now_s() - t0;
  put_ring(w, "wit_A_io", wit.coefficients_for_A_io);
  put_ring(w, "wit_B_io", wit.coefficients_for_B_io);
  put_ring(w, "wit_C_io", wit.coefficients_for_C_io);
  put_ring(w, "wit_A_mid", wit.coefficients_for_A_mid);
  put_ring(w, "wit_B_mid", wit.coefficients_for_B_mid);
  put_ring(w, "wit_C_mid", wit.coefficients_for_C_mid);
  put_ring(w, "wit_Z", wit.coefficients_for_Z);
  put_ring(w, "wit_H", wit.coefficients_for_H);

  // (3) CRS from the reference generator (groth16.tcc:4-67), seeded through the encoding contexts
  t0 = now_s();
  const auto kp = ringsnark::groth16::generator<R, E>(s.cs);
  double t_gen = now_s() - t0;
  put_enc_vec(w, "crs_s_pows", kp.pk.s_pows, s.L_R, s.L_E, s.N_E);
  put_enc_vec(w, "crs_delta_ts", kp.pk.delta_ts, s.L_R, s.L_E, s.N_E);
  put_enc_vec(w, "crs_delta_mid", kp.pk.delta_mid, s.L_R, s.L_E, s.N_E);
  put_enc_vec(w, "crs_alpha", vector<E>{kp.pk.alpha}, s.L_R, s.L_E, s.N_E);
  put_enc_vec(w, "crs_beta", vector<E>{kp.pk.beta}, s.L_R, s.L_E, s.N_E);

  // (4) each inner product of the prover on its own (seal_ring.tcc:361-433) ...
  const auto &sp = kp.pk.s_pows;
  vector<E> ips(6);  // default-constructed = empty; += avoids copying an empty result (seal_ring.hpp:245-247 asserts)
  ips[0] += E::inner_product(sp.begin(), sp.end() - 1, wit.coefficients_for_A_io.begin(), wit.coefficients_for_A_io.end());
  ips[1] += E::inner_product(sp.begin(), sp.end() - 1, wit.coefficients_for_A_mid.begin(), wit.coefficients_for_A_mid.end());
  ips[2] += E::inner_product(sp.begin(), sp.end() - 1, wit.coefficients_for_B_io.begin(), wit.coefficients_for_B_io.end());
  ips[3] += E::inner_product(sp.begin(), sp.end() - 1, wit.coefficients_for_B_mid.begin(), wit.coefficients_for_B_mid.end());
  ips[4] += E::inner_product(kp.pk.delta_ts.begin(), kp.pk.delta_ts.end(), wit.coefficients_for_H.begin(), wit.coefficients_for_H.end());
  if (!s.auxiliary.empty())
    ips[5] += E::inner_product(kp.pk.delta_mid.begin(), kp.pk.delta_mid.end(), s.auxiliary.begin(), s.auxiliary.end());
  put_enc_vec(w, "ip", ips, s.L_R, s.L_E, s.N_E);  // order: A_io, A_mid, B_io, B_mid, H, aux

  // ... (5) and the proof itself (groth16.tcc:69-115) with the verifier's verdict (groth16.tcc:117-170)
  t0 = now_s();
  const auto proof = ringsnark::groth16::prover(kp.pk, s.primary, s.auxiliary);
  double t_prove = now_s() - t0;
  put_enc_vec(w, "proof", vector<E>{proof.A, proof.B, proof.C}, s.L_R, s.L_E, s.N_E);
  bool ok = false;
  // an EMPTY proof element (C2': n = 1, no auxiliary input -> C is the empty sum) makes the reference verifier index an empty
  // ciphertext vector (seal_ring.tcc:435-445 only asserts the size): not runnable, recorded as verified = 2
  const bool has_empty = proof.A.is_empty() || proof.B.is_empty() || proof.C.is_empty();
  if (!has_empty) {
    try {
      ok = ringsnark::groth16::verifier(kp.vk, s.primary, proof);
    } catch (const std::exception &ex) {
      std::cerr << "verifier threw: " << ex.what() << std::endl;
    }
  }
  w.put1("verified", has_empty ? 2 : (uint64_t)ok);

  // (6) primitive-level known answers from SEAL itself for term 0 / ring limb 0 of (s_pows, A_mid-like poly):
  //     BatchEncoder::encode output (batchencoder.cpp:110-149) and transform_to_ntt_inplace output
  //     (evaluator.cpp:2174-2265), plus one multiply_plain and one forward/inverse NTT pair per Q_l.
  {
    R elem = make_elem(0);
    put_ring(w, "kat_elem", vector<R>{elem});
    vector<uint64_t> pc, pn, prod;
    for (size_t j = 0; j < s.L_R; j++) {
      seal::BatchEncoder be(E::get_contexts()[j]);
      seal::Evaluator ev(E::get_contexts()[j]);
      seal::Plaintext pt;
      be.encode(elem.get_poly().get_limb(j), pt);
      pc.insert(pc.end(), pt.data(), pt.data() + pt.coeff_count());
      seal::Plaintext pt2 = pt;
      ev.transform_to_ntt_inplace(pt2, E::get_contexts()[j].first_parms_id());
      pn.insert(pn.end(), pt2.data(), pt2.data() + s.L_E * s.N_E);
      seal::Ciphertext ct = EncAccess::cts(sp[0])[j];
      ev.multiply_plain_inplace(ct, pt);
      prod.insert(prod.end(), ct.data(), ct.data() + 2 * s.L_E * s.N_E);
    }
    w.put("kat_plain_coeff", pc);   // [L_R][N_E]
    w.put("kat_plain_ntt", pn);     // [L_R][L_E][N_E]
    w.put("kat_mul_plain", prod);   // [L_R][2][L_E][N_E] = s_pows[0] * elem
    // raw NTT pair mod Q_0 and mod q_0 on a ramp
    auto tabQ = E::get_contexts()[0].first_context_data()->small_ntt_tables();
    vector<uint64_t> x(s.N_E);
    for (size_t i = 0; i < s.N_E; i++) x[i] = (i * 0x9E3779B97F4A7C15ull + 12345) % s.Q[0];
    w.put("kat_ntt_in", x);
    seal::util::ntt_negacyclic_harvey(x.data(), tabQ[0]);
    w.put("kat_ntt_fwd_Q0", x);
    auto tabt = E::get_contexts()[0].first_context_data()->plain_ntt_tables();
    vector<uint64_t> y(s.N_E);
    for (size_t i = 0; i < s.N_E; i++) y[i] = (i * 0xD1B54A32D192ED03ull + 777) % s.q[0];
    w.put("kat_intt_in", y);
    seal::util::inverse_ntt_negacyclic_harvey(y.data(), *tabt);
    w.put("kat_intt_inv_q0", y);
    w.put1("kat_root_Q0", tabQ[0].get_root());
    w.put1("kat_root_q0", tabt->get_root());
  }
  // (7) instance map with evaluation (r1cs_to_qrp.tcc:75-116) at a fresh exceptional point, as generator
  //     (groth16.tcc:11-12) and verifier (groth16.tcc:127-128) call it.  Drawn AFTER everything above so that the ring
  //     PRNG stream of (1)-(6) is what it was before this section existed.
  double t_inst = 0;
  {
    const auto domain = ringsnark::get_evaluation_domain<R>(s.cs.num_constraints());
    const R t = R::random_exceptional_element(domain);
    t0 = now_s();
    const auto inst = ringsnark::r1cs_to_qrp_instance_map_with_evaluation(s.cs, t);
    t_inst = now_s() - t0;
    put_ring(w, "inst_t", vector<R>{t});
    put_ring(w, "inst_At", inst.At);
    put_ring(w, "inst_Bt", inst.Bt);
    put_ring(w, "inst_Ct", inst.Ct);
    put_ring(w, "inst_Ht", inst.Ht);
    put_ring(w, "inst_Zt", vector<R>{inst.Zt});
  }
  // (8) EncodingElem::decode (seal_ring.tcc:435-477) of the three proof elements, as the verifier does first
  //     (groth16.tcc:121-123): the secret keys in NTT form, the decoded ring elements and SEAL's invariant noise budgets
  {
    const auto &sk = kp.vk.sk_enc;
    vector<uint64_t> skw, budgets, decoded_ok;
    for (size_t j = 0; j < s.L_R; j++) skw.insert(skw.end(), sk[j].data().data(), sk[j].data().data() + s.L_E * s.N_E);
    w.put("dec_sk", skw);   // [L_R][L_E][N_E]: first-level limbs of the key-level secret key
    vector<R> dec;
    for (const E *e : {&proof.A, &proof.B, &proof.C}) {
      bool ok = false;
      R r = R::zero();
      if (!e->is_empty()) {
        const auto &cts = EncAccess::cts(*e);
        for (size_t j = 0; j < s.L_R; j++) {
          seal::Decryptor d(E::get_contexts()[j], sk[j]);
          budgets.push_back(cts[j].size() ? (uint64_t)d.invariant_noise_budget(cts[j]) : ~0ull);
        }
        try {
          r = E::decode(sk, *e);
          ok = true;
        } catch (const std::exception &ex) {
          std::cerr << "decode threw: " << ex.what() << std::endl;
        }
      } else {
        for (size_t j = 0; j < s.L_R; j++) budgets.push_back(~0ull);
      }
      decoded_ok.push_back(ok);
      dec.push_back(ok ? r.to_poly() : R::zero().to_poly());
    }
    put_ring(w, "dec_proof", dec);
    w.put("dec_budget", budgets);   // [3][L_R]; ~0 = empty ciphertext (no budget defined)
    w.put("dec_ok", decoded_ok);
  }
  w.put("timing_us", vector<uint64_t>{(uint64_t)(t_wit * 1e6), (uint64_t)(t_gen * 1e6), (uint64_t)(t_prove * 1e6), (uint64_t)(t_inst * 1e6)});
  w.save(out);
  std::cerr << "case " << name << ": satisfied=" << sat << " verified=" << ok << " witness_map=" << t_wit
            << "s generator=" << t_gen << "s prover=" << t_prove << "s -> " << out << std::endl;
  return 0;
}

// --- timing ---------------------------------------------------------------------------------------------
// Synthetic CRS element: uniform words in [0, Q_l), NTT form, first level -- what a real CRS looks like to the
// prover (the reference generator is O(n^2) ring inversions and takes ~90 s at C4).
static E synth_enc(uint64_t &st) {
  vector<seal::Ciphertext> cts;
  for (auto &ctx : E::get_contexts()) {
    seal::Ciphertext ct(ctx);
    ct.resize(ctx, ctx.first_parms_id(), 2);
    ct.is_ntt_form() = true;
    auto &mods = ctx.first_context_data()->parms().coeff_modulus();
    const size_t N = ctx.first_context_data()->parms().poly_modulus_degree();
    for (size_t k = 0; k < 2; k++)
      for (size_t l = 0; l < mods.size(); l++) {
        uint64_t *p = ct.data(k) + l * N;
        for (size_t i = 0; i < N; i++) {
          st ^= st << 13; st ^= st >> 7; st ^= st << 17;
          p[i] = st % mods[l].value();
        }
      }
    cts.push_back(ct);
  }
  return E(cts);
}

static int cmd_time(const std::string &name, const std::string &what, int argc, char **argv) {
  // args: terms=<T> n=<n> reps=<r> threads=<t>
  size_t terms = 32, n_override = 0, reps = 1, threads = 1;
  for (int i = 0; i < argc; i++) {
    std::string a(argv[i]);
    if (a.rfind("terms=", 0) == 0) terms = std::stoul(a.substr(6));
    if (a.rfind("n=", 0) == 0) n_override = std::stoul(a.substr(2));
    if (a.rfind("reps=", 0) == 0) reps = std::stoul(a.substr(5));
    if (a.rfind("threads=", 0) == 0) threads = std::stoul(a.substr(8));
  }
  g_case = cases::get_case(name);
  if (n_override) {  // shrink the circuit, keep the shape ratios
    double f = (double)n_override / g_case.n;
    g_case.n = n_override;
    g_case.io = std::max<size_t>(1, (size_t)(g_case.io * f));
    g_case.aux = std::max<size_t>(n_override + 1 - g_case.io, (size_t)(g_case.aux * f));
  }
  const cases::CaseSpec cspec = g_case;
  g_ring = new seal::SEALContext(cases::make_ring_context(cspec));
  R::set_context(*g_ring);
  E::set_contexts(cases::make_enc_contexts(cspec, *g_ring, 1));
  RingAccess::seed(1);
#ifdef _OPENMP
  omp_set_num_threads((int)threads);
#endif
  std::ostringstream js;
  if (what == "lincomb") {
    // `threads` independent inner products of `terms` terms each, run concurrently (the way
    // rinocchio.tcc:106-163 uses OpenMP sections); throughput = threads*terms / wall.
    uint64_t st = 0x243F6A8885A308D3ull;
    vector<E> crs;
    vector<R> coeffs;
    for (size_t i = 0; i < terms; i++) { crs.push_back(synth_enc(st)); coeffs.push_back(R::random_element()); }
    vector<double> ts;
    for (size_t r = 0; r < reps; r++) {
      double t0 = now_s();
#pragma omp parallel for num_threads(threads)
      for (size_t th = 0; th < threads; th++) {
        E res = E::inner_product(crs.begin(), crs.end(), coeffs.begin(), coeffs.end());
        if (res.is_empty()) abort();
      }
      ts.push_back(now_s() - t0);
    }
    std::sort(ts.begin(), ts.end());
    double med = ts[ts.size() / 2];
    js << "{\"what\":\"lincomb\",\"case\":\"" << name << "\",\"terms\":" << terms << ",\"threads\":" << threads
       << ",\"reps\":" << reps << ",\"seconds\":" << med << ",\"ms_per_term_per_thread\":" << 1e3 * med / terms
       << ",\"terms_per_s\":" << threads * terms / med << "}";
  } else if (what == "witness") {
    ringsnark::r1cs_constraint_system<R> cs;
    vector<R> assignment;
    cases::build_circuit<R, ringsnark::r1cs_constraint_system<R>, ringsnark::r1cs_constraint<R>,
                         ringsnark::linear_combination<R>, ringsnark::variable<R>>(cspec, 1, cs, assignment, &make_elem);
    vector<R> primary(assignment.begin(), assignment.begin() + cspec.io), auxiliary(assignment.begin() + cspec.io, assignment.end());
    vector<double> ts;
    for (size_t r = 0; r < reps; r++) {
      double t0 = now_s();
      const auto wit = ringsnark::r1cs_to_qrp_witness_map(cs, primary, auxiliary, R::zero(), R::zero(), R::zero());
      ts.push_back(now_s() - t0);
      if (wit.coefficients_for_H.empty()) abort();
    }
    std::sort(ts.begin(), ts.end());
    js << "{\"what\":\"witness\",\"case\":\"" << name << "\",\"n\":" << cspec.n << ",\"threads\":1,\"reps\":" << reps
       << ",\"seconds\":" << ts[ts.size() / 2] << "}";
  } else if (what == "instance") {
    // r1cs_to_qrp_instance_map_with_evaluation (r1cs_to_qrp.tcc:75-116): the O(m^2) step of generator and verifier
    ringsnark::r1cs_constraint_system<R> cs;
    vector<R> assignment;
    cases::build_circuit<R, ringsnark::r1cs_constraint_system<R>, ringsnark::r1cs_constraint<R>,
                         ringsnark::linear_combination<R>, ringsnark::variable<R>>(cspec, 1, cs, assignment, &make_elem);
    const auto domain = ringsnark::get_evaluation_domain<R>(cs.num_constraints());
    const R t = R::random_exceptional_element(domain);
    vector<double> ts;
    for (size_t r = 0; r < reps; r++) {
      double t0 = now_s();
      const auto inst = ringsnark::r1cs_to_qrp_instance_map_with_evaluation(cs, t);
      ts.push_back(now_s() - t0);
      if (inst.Ht.empty()) abort();
    }
    std::sort(ts.begin(), ts.end());
    js << "{\"what\":\"instance\",\"case\":\"" << name << "\",\"n\":" << cspec.n << ",\"threads\":1,\"reps\":" << reps
       << ",\"seconds\":" << ts[ts.size() / 2] << "}";
  } else if (what == "decode") {
    // EncodingElem::decode (seal_ring.tcc:435-477) of `terms` fresh encodings: noise budget + decrypt + batch decode each
    auto keys = E::keygen();
    vector<R> rs;
    for (size_t i = 0; i < terms; i++) rs.push_back(R::random_element());
    const auto encs = E::encode(std::get<1>(keys), rs);
    vector<double> ts;
    for (size_t r = 0; r < reps; r++) {
      double t0 = now_s();
      for (size_t i = 0; i < terms; i++) {
        const R d = E::decode(std::get<1>(keys), encs[i]);
        if (d != rs[i]) abort();
      }
      ts.push_back(now_s() - t0);
    }
    std::sort(ts.begin(), ts.end());
    js << "{\"what\":\"decode\",\"case\":\"" << name << "\",\"encodings\":" << terms << ",\"threads\":1,\"reps\":" << reps
       << ",\"seconds\":" << ts[ts.size() / 2] << ",\"ms_per_encoding\":" << 1e3 * ts[ts.size() / 2] / terms << "}";
  } else if (what == "prover") {
    // full groth16::prover on a synthetic CRS of the right shape
    ringsnark::r1cs_constraint_system<R> cs;
    vector<R> assignment;
    cases::build_circuit<R, ringsnark::r1cs_constraint_system<R>, ringsnark::r1cs_constraint<R>,
                         ringsnark::linear_combination<R>, ringsnark::variable<R>>(cspec, 1, cs, assignment, &make_elem);
    vector<R> primary(assignment.begin(), assignment.begin() + cspec.io), auxiliary(assignment.begin() + cspec.io, assignment.end());
    uint64_t st = 0x243F6A8885A308D3ull;
    vector<E> s_pows, gamma_io, delta_mid, delta_ts;
    for (size_t i = 0; i < cspec.n + 1; i++) { s_pows.push_back(synth_enc(st)); delta_ts.push_back(synth_enc(st)); }
    for (size_t i = 0; i < cspec.io + 1; i++) gamma_io.push_back(s_pows[0]);
    for (size_t i = 0; i < cspec.aux; i++) delta_mid.push_back(synth_enc(st));
    ringsnark::groth16::proving_key<R, E> pk(cs, synth_enc(st), synth_enc(st), s_pows, gamma_io, delta_mid, delta_ts, nullptr);
    vector<double> ts;
    std::streambuf *old = std::cout.rdbuf(std::cerr.rdbuf());  // prover chats on stdout
    for (size_t r = 0; r < reps; r++) {
      double t0 = now_s();
      const auto proof = ringsnark::groth16::prover(pk, primary, auxiliary);
      ts.push_back(now_s() - t0);
      if (proof.A.is_empty()) abort();
    }
    std::cout.rdbuf(old);
    std::sort(ts.begin(), ts.end());
    js << "{\"what\":\"prover\",\"case\":\"" << name << "\",\"n\":" << cspec.n << ",\"io\":" << cspec.io << ",\"aux\":" << cspec.aux
       << ",\"threads\":1,\"reps\":" << reps << ",\"seconds\":" << ts[ts.size() / 2] << "}";
  } else {
    std::cerr << "unknown timing target " << what << std::endl;
    return 2;
  }
  std::cout << js.str() << std::endl;
  return 0;
}

int main(int argc, char **argv) {
  try {
    if (argc >= 2 && std::string(argv[1]) == "list") {
      for (const char *n : {"tiny_fast", "tiny_slow", "tiny_quirks", "tiny_full", "c1", "c2p", "c3p", "c4s", "c4m", "c4"}) std::cout << n << "\n";
      return 0;
    }
    if (argc >= 4 && std::string(argv[1]) == "dump") {
      uint64_t seed = argc >= 5 ? std::stoull(argv[4]) : 0xB200;
      std::streambuf *old = std::cout.rdbuf(std::cerr.rdbuf());
      int rc = cmd_dump(argv[2], argv[3], seed);
      std::cout.rdbuf(old);
      return rc;
    }
    if (argc >= 4 && std::string(argv[1]) == "time") return cmd_time(argv[2], argv[3], argc - 4, argv + 4);
    std::cerr << "usage: ref_harness dump <case> <out.rsgv> [seed] | time <case> <prover|lincomb|witness|instance|decode> [terms=T n=N reps=R threads=T] | list\n";
    return 2;
  } catch (const std::exception &ex) {
    std::cerr << "ref_harness: " << ex.what() << std::endl;
    return 1;
  }
}
