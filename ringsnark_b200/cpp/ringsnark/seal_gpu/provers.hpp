// ringsnark/seal_gpu/provers.hpp -- groth16::prover and rinocchio::prover for the B200 backend as ONE library call each.
// Included at the end of seal_gpu/seal_ring.hpp.  Explicit specialisations of the reference's function templates
// (zk_proof_systems/groth16/groth16.hpp, zk_proof_systems/rinocchio/rinocchio.hpp) for <seal_gpu::RingElem,
// seal_gpu::EncodingElem>: drivers keep calling ringsnark::groth16::prover(pk, primary, auxiliary) unchanged.
//
// The un-specialised templates work over this backend too (every operator they use is defined), but they pay per-element host
// work: a RingElem per coefficient, one C call per inner product, and -- in rinocchio::prover -- the plaintexts of a_mid, b_mid,
// c_mid, h, z are encoded and transformed twice (s_pows and alpha_s_pows).  The fused calls keep the witness on the device from
// the assignment to the proof (rsg_groth16_prove_refs / rsg_rinocchio_prove).  They decline (and the generic sequence below,
// a restatement of groth16.tcc:69-115 / rinocchio.tcc:74-190 over the same operators, runs instead) when
//   * a linear-combination coefficient is a ring element rather than an integer (no CSR form: the NTT demo circuit),
//   * a key vector is not one contiguous arena range (it was not produced by one encode() call),
//   * the library reports a transparent-ciphertext candidate (seal_ring.tcc:493-504) or a term set above its scratch budget.
#ifndef RINGSNARK_SEAL_GPU_PROVERS_HPP
#define RINGSNARK_SEAL_GPU_PROVERS_HPP

#include <ringsnark/zk_proof_systems/groth16/groth16.hpp>
#include <ringsnark/zk_proof_systems/rinocchio/rinocchio.hpp>

namespace ringsnark::seal_gpu::detail {

struct FusedInputs {
  std::vector<uint32_t> row_ptr{0}, col;
  std::vector<uint64_t> coeff;
  std::vector<uint8_t> aux_kind;
  std::shared_ptr<DevRing> assignment;
  rsg_r1cs *r1cs = nullptr;
  ~FusedInputs() { rsg_r1cs_destroy(r1cs); }
};
// CSR form of the constraint system + the assignment in HBM; false when a coefficient is not an integer
inline bool fused_inputs(const r1cs_constraint_system<RingElem> &cs, const r1cs_primary_input<RingElem> &primary,
                         const r1cs_auxiliary_input<RingElem> &auxiliary, FusedInputs &in) {
  auto &b = backend();
  const size_t n = cs.num_constraints(), n_io = primary.size(), n_aux = auxiliary.size();
  if (n_io != cs.primary_input_size || n_aux != cs.auxiliary_input_size)
    throw std::invalid_argument("assignment does not match the constraint system");
  for (int m = 0; m < 3; m++)
    for (size_t i = 0; i < n; i++) {
      const auto &lc = m == 0 ? cs.constraints[i].a : (m == 1 ? cs.constraints[i].b : cs.constraints[i].c);
      for (const auto &lt : lc.terms) {
        if (!lt.coeff.is_scalar()) return false;
        in.col.push_back((uint32_t)lt.index);
        in.coeff.push_back(lt.coeff.get_scalar());
      }
      in.row_ptr.push_back((uint32_t)in.col.size());
    }
  std::vector<uint64_t> w;
  w.reserve((n_io + n_aux) * b.ring_words);
  for (const auto &r : primary) r.append_words(w);
  for (const auto &r : auxiliary) r.append_words(w);
  // the reference's per-term dispatch for auxiliary inputs held as scalars (seal_ring.tcc:514-529); polynomials get
  // SealPoly::is_zero's prefix test on the device
  in.aux_kind.assign(n_aux, (uint8_t)RSG_AUX_POLY);
  for (size_t i = 0; i < n_aux; i++)
    if (auxiliary[i].is_scalar()) {
      const uint64_t s = auxiliary[i].get_scalar();
      in.aux_kind[i] = s == 0 ? RSG_TERM_SKIP : (s == 1 ? RSG_TERM_ONE : RSG_TERM_GENERAL);
    }
  in.assignment = std::make_shared<DevRing>();
  check(rsg_ringvec_create(b.ctx, n_io + n_aux ? n_io + n_aux : 1, &in.assignment->v));
  if (n_io + n_aux) check(rsg_ringvec_upload(in.assignment->v, 0, n_io + n_aux, w.data()));
  check(rsg_r1cs_create(b.ctx, n, n_io, n_aux, in.row_ptr.data(), in.col.data(), in.coeff.data(), &in.r1cs));
  return true;
}
inline bool single(const EncodingElem &e, rsg_crs_ref *ref) {
  ref->crs = e.arena_handle();
  ref->first = e.arena_index();
  return ref->crs != nullptr;
}
// encodings [0, count) of a fresh arena as EncodingElems; used[k] == 0 -> the reference's empty encoding
inline std::vector<EncodingElem> wrap_proof(const std::shared_ptr<DevEnc> &arena, const size_t *used, size_t count) {
  std::vector<EncodingElem> out(count);
  for (size_t k = 0; k < count; k++)
    if (used[k]) out[k] = EncodingElem(arena, k);
  return out;
}

// groth16.tcc:69-115 over this backend's operators (what the un-specialised template does)
inline groth16::proof<RingElem, EncodingElem> groth16_prover_generic(const groth16::proving_key<RingElem, EncodingElem> &pk,
                                                                     const r1cs_primary_input<RingElem> &primary,
                                                                     const r1cs_auxiliary_input<RingElem> &auxiliary) {
  using E = EncodingElem;
  const qrp_witness<RingElem> w =
      r1cs_to_qrp_witness_map(pk.constraint_system, primary, auxiliary, RingElem::zero(), RingElem::zero(), RingElem::zero());
  const auto sb = pk.s_pows.begin(), se = pk.s_pows.end() - 1;
  E a = E::inner_product(sb, se, w.coefficients_for_A_io.begin(), w.coefficients_for_A_io.end());
  a += E::inner_product(sb, se, w.coefficients_for_A_mid.begin(), w.coefficients_for_A_mid.end());
  a += pk.alpha;
  E bb = E::inner_product(sb, se, w.coefficients_for_B_io.begin(), w.coefficients_for_B_io.end());
  bb += E::inner_product(sb, se, w.coefficients_for_B_mid.begin(), w.coefficients_for_B_mid.end());
  bb += pk.beta;
  E c = E::inner_product(pk.delta_ts.begin(), pk.delta_ts.end(), w.coefficients_for_H.begin(), w.coefficients_for_H.end());
  if (!auxiliary.empty()) c += E::inner_product(pk.delta_mid.begin(), pk.delta_mid.end(), auxiliary.begin(), auxiliary.end());
  return groth16::proof<RingElem, EncodingElem>(a, bb, c);
}
// rinocchio.tcc:74-190 likewise (d1, d2, d3 are drawn by the caller, in the reference's order)
inline rinocchio::proof<RingElem, EncodingElem> rinocchio_prover_generic(const rinocchio::proving_key<RingElem, EncodingElem> &pk,
                                                                         const r1cs_primary_input<RingElem> &primary,
                                                                         const r1cs_auxiliary_input<RingElem> &auxiliary, bool use_zk,
                                                                         const RingElem &d1, const RingElem &d2, const RingElem &d3) {
  using E = EncodingElem;
  const qrp_witness<RingElem> w = r1cs_to_qrp_witness_map(pk.constraint_system, primary, auxiliary, d1, d2, d3);
  const auto &am = w.coefficients_for_A_mid, &bm = w.coefficients_for_B_mid, &cm = w.coefficients_for_C_mid;
  const auto &z = w.coefficients_for_Z, &h = w.coefficients_for_H;
  const auto sb = pk.s_pows.begin(), se = pk.s_pows.end(), ab = pk.alpha_s_pows.begin(), ae = pk.alpha_s_pows.end();
  E a = E::inner_product(sb, se - 1, am.begin(), am.end()), aa = E::inner_product(ab, ae - 1, am.begin(), am.end());
  E b = E::inner_product(sb, se - 1, bm.begin(), bm.end()), ba = E::inner_product(ab, ae - 1, bm.begin(), bm.end());
  E c = E::inner_product(sb, se - 1, cm.begin(), cm.end()), ca = E::inner_product(ab, ae - 1, cm.begin(), cm.end());
  E d = E::inner_product(sb, se, h.begin(), h.end()), da = E::inner_product(ab, ae, h.begin(), h.end());
  E ze = E::inner_product(sb, se, z.begin(), z.end()), za = E::inner_product(ab, ae, z.begin(), z.end());
  if (use_zk) {
    a += d1 * ze; aa += d1 * za;
    b += d2 * ze; ba += d2 * za;
    c += d3 * ze; ca += d3 * za;
  }
  E f;
  if (!auxiliary.empty()) {
    f = E::inner_product(pk.beta_prods.begin(), pk.beta_prods.end(), auxiliary.begin(), auxiliary.end());
    if (use_zk) {
      f += d1 * pk.beta_rv_ts;
      f += d2 * pk.beta_rw_ts;
      f += d3 * pk.beta_ry_ts;
    }
  }
  return rinocchio::proof<RingElem, EncodingElem>(a, aa, b, ba, c, ca, d, da, f);
}
inline bool fused_enabled() {
  const char *m = std::getenv("RSG_FUSED");
  return !(m && std::string(m) == "0");
}
}  // namespace ringsnark::seal_gpu::detail

namespace ringsnark::groth16 {
template <>
inline proof<seal_gpu::RingElem, seal_gpu::EncodingElem> prover<seal_gpu::RingElem, seal_gpu::EncodingElem>(
    const proving_key<seal_gpu::RingElem, seal_gpu::EncodingElem> &pk, const r1cs_primary_input<seal_gpu::RingElem> &primary_input,
    const r1cs_auxiliary_input<seal_gpu::RingElem> &auxiliary_input) {
  using E = seal_gpu::EncodingElem;
  namespace D = seal_gpu::detail;
  cout << "[Prover] " << "using non-zero-knowledge SNARK" << endl;   // groth16.tcc:76-80
  rsg_crs_ref refs[5];
  D::FusedInputs in;
  const bool fused = D::fused_enabled() && E::contiguous(pk.s_pows.begin(), pk.s_pows.end(), &refs[0]) &&
                     E::contiguous(pk.delta_ts.begin(), pk.delta_ts.end(), &refs[1]) &&
                     E::contiguous(pk.delta_mid.begin(), pk.delta_mid.end(), &refs[2]) && D::single(pk.alpha, &refs[3]) &&
                     D::single(pk.beta, &refs[4]) && D::fused_inputs(pk.constraint_system, primary_input, auxiliary_input, in);
  if (fused) {
    auto &b = D::backend();
    auto arena = D::new_arena(3);
    size_t used[3] = {0, 0, 0};
    const int rc = rsg_groth16_prove_refs(b.ctx, in.r1cs, refs, in.assignment->v, nullptr, in.aux_kind.data(), nullptr,
                                          rsg_crs_device_ptr(arena->c), used);
    if (rc == RSG_OK) {
      const auto e = D::wrap_proof(arena, used, 3);
      return proof<seal_gpu::RingElem, seal_gpu::EncodingElem>(e[0], e[1], e[2]);
    }
    if (rc != RSG_ERR_TRANSPARENT && rc != RSG_ERR_UNSUPPORTED) D::check(rc);
  }
  return D::groth16_prover_generic(pk, primary_input, auxiliary_input);
}
}  // namespace ringsnark::groth16

namespace ringsnark::rinocchio {
template <>
inline proof<seal_gpu::RingElem, seal_gpu::EncodingElem> prover<seal_gpu::RingElem, seal_gpu::EncodingElem>(
    const proving_key<seal_gpu::RingElem, seal_gpu::EncodingElem> &pk, const r1cs_primary_input<seal_gpu::RingElem> &primary_input,
    const r1cs_auxiliary_input<seal_gpu::RingElem> &auxiliary_input) {
  using R = seal_gpu::RingElem;
  using E = seal_gpu::EncodingElem;
  namespace D = seal_gpu::detail;
  const bool use_zk = !auxiliary_input.empty();   // rinocchio.tcc:81-90
  if (!use_zk) cout << "[Prover] " << "using non-zero-knowledge SNARK, since no auxiliary inputs are " "defined" << endl;
  const R d1 = use_zk ? R::random_invertible_element() : R::zero();
  const R d2 = use_zk ? R::random_invertible_element() : R::zero();
  const R d3 = use_zk ? R::random_invertible_element() : R::zero();
  rsg_crs_ref refs[6] = {};
  D::FusedInputs in;
  bool fused = D::fused_enabled() && E::contiguous(pk.s_pows.begin(), pk.s_pows.end(), &refs[0]) &&
               E::contiguous(pk.alpha_s_pows.begin(), pk.alpha_s_pows.end(), &refs[1]) &&
               E::contiguous(pk.beta_prods.begin(), pk.beta_prods.end(), &refs[2]);
  if (fused && use_zk) fused = D::single(pk.beta_rv_ts, &refs[3]) && D::single(pk.beta_rw_ts, &refs[4]) && D::single(pk.beta_ry_ts, &refs[5]);
  fused = fused && D::fused_inputs(pk.constraint_system, primary_input, auxiliary_input, in);
  if (fused) {
    auto &b = D::backend();
    std::vector<uint64_t> dw;
    if (use_zk) {
      d1.append_words(dw);
      d2.append_words(dw);
      d3.append_words(dw);
    }
    auto arena = D::new_arena(9);
    size_t used[9] = {0};
    const int rc = rsg_rinocchio_prove(b.ctx, in.r1cs, refs, in.assignment->v, nullptr, in.aux_kind.data(), use_zk ? dw.data() : nullptr,
                                       nullptr, rsg_crs_device_ptr(arena->c), used);
    if (rc == RSG_OK) {
      const auto e = D::wrap_proof(arena, used, 9);
      return proof<R, E>(e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7], e[8]);
    }
    if (rc != RSG_ERR_TRANSPARENT && rc != RSG_ERR_UNSUPPORTED) D::check(rc);
  }
  return D::rinocchio_prover_generic(pk, primary_input, auxiliary_input, use_zk, d1, d2, d3);
}
}  // namespace ringsnark::rinocchio

#endif  // RINGSNARK_SEAL_GPU_PROVERS_HPP
