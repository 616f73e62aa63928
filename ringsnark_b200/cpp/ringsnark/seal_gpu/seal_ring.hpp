// ringsnark/seal_gpu/seal_ring.hpp -- B200 backend for zkFHE/ringSNARK: a RingElem / EncodingElem pair with the
// surface of ringsnark/seal/seal_ring.hpp:18-420 whose prover hot path runs on the GPU through the C ABI of
// include/rsgpu.h (librsgpu.so).  Drop this directory next to ringsnark/seal/ in a ringSNARK checkout, include it
// instead of <ringsnark/seal/seal_ring.hpp>, and
//     ringsnark::groth16::prover<seal_gpu::RingElem, seal_gpu::EncodingElem>(pk, primary, auxiliary)
//     ringsnark::rinocchio::prover<...>(...)
// instantiate from the UNCHANGED templates in ringsnark/zk_proof_systems.
//
// What runs where
//   GPU : EncodingElem::inner_product (seal_ring.tcc:361-433), EncodingElem::operator*= (:509-548) and operator+=
//         (:479-507), and -- through the explicit specialisation at the end of this file --
//         r1cs_to_qrp_witness_map (reductions/r1cs_to_qrp/r1cs_to_qrp.tcc:148-259) including the sparse
//         linear_combination::evaluate pass in front of it (relations/variable.tcc:246-254), and
//         r1cs_to_qrp_instance_map_with_evaluation (r1cs_to_qrp.tcc:75-116: the O(m^2) step of setup and of every
//         verification).  There is no CPU fallback for these: without librsgpu.so and a CUDA device they throw.
//   host: ring elements that drivers build (circuits, assignments) are host values exactly as in the reference --
//         this class keeps a ringsnark::seal::RingElem inside and forwards the element-wise operators to it, so the
//         scalar/polynomial variant rules (seal_ring.tcc:105-263) and SealPoly::is_zero's prefix quirk
//         (depends/SEAL-Polytools/src/poly_arith.cpp:147-153) are the reference's own code, not a re-statement.
//         Setup (keygen, encode) delegates to ringsnark::seal::EncodingElem: SURVEY.md section 8 keeps it on the SEAL
//         path.  EncodingElem::decode (the verifier's front half) runs on the GPU (rsg_decode).
//   Ring elements PRODUCED by the GPU witness map stay in HBM (a RingElem then holds a ref-counted slice of a device
//   vector plus its is_zero flag) and are fed to inner_product without a round trip; they are downloaded lazily only
//   if host code looks at them.  Encodings live in HBM arenas: one arena per encode() call, so a proving-key vector
//   is contiguous and a lincomb over an iterator range of it is one streaming pass.
#ifndef RINGSNARK_SEAL_GPU_RING_HPP
#define RINGSNARK_SEAL_GPU_RING_HPP

#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <tuple>
#include <vector>

#include <ringsnark/seal/seal_ring.hpp>
#include <ringsnark/reductions/r1cs_to_qrp/r1cs_to_qrp.hpp>
#include <ringsnark/util/polynomials.hpp>

#include "rsgpu.h"

namespace ringsnark::seal_gpu {

namespace detail {
// C-ABI status -> the exception types the reference throws (SURVEY.md section 8(b) "Errors").
inline void check(int rc) {
  if (rc == RSG_OK) return;
  const std::string msg = rsg_last_error();
  switch (rc) {
    case RSG_ERR_STATE: throw std::invalid_argument("context not set");
    case RSG_ERR_NOTINV: throw std::invalid_argument("element is not invertible in ring");
    case RSG_ERR_ARG: throw std::invalid_argument("rsgpu: " + msg);
    default: throw std::runtime_error("rsgpu: " + msg);
  }
}

struct Backend {
  rsg_context *ctx = nullptr;
  size_t N_R = 0, L_R = 0, N_E = 0, L_E = 0, ring_words = 0, enc_words = 0, enc_bits = 0;
  std::vector<uint64_t> q, Q;
};
inline Backend &backend_storage() {
  static Backend b;
  return b;
}
inline Backend &backend() {
  Backend &b = backend_storage();
  if (!b.ctx) throw std::invalid_argument("context not set");
  return b;
}
inline std::mutex &host_mutex() {
  static std::mutex m;
  return m;
}

struct DevRing {   // a vector<RingElem> resident in HBM
  rsg_ringvec *v = nullptr;
  std::vector<uint8_t> zero;   // SealPoly::is_zero (prefix semantics) of every element, computed on the device
  ~DevRing() { rsg_ringvec_destroy(v); }
};
struct DevEnc {    // a vector<EncodingElem> resident in HBM
  rsg_crs *c = nullptr;
  size_t n = 0;
  ~DevEnc() { rsg_crs_destroy(c); }
};
inline std::shared_ptr<DevEnc> new_arena(size_t n) {
  auto a = std::make_shared<DevEnc>();
  check(rsg_crs_create(backend().ctx, n, &a->c));
  a->n = n;
  return a;
}

// access to the protected ciphertext vector of the reference's EncodingElem (no reference file is modified)
struct SealEncAccess : ::ringsnark::seal::EncodingElem {
  static const std::vector<::seal::Ciphertext> &cts(const ::ringsnark::seal::EncodingElem &e) {
    return e.*(&SealEncAccess::ciphertexts);
  }
};
}  // namespace detail

// =====================================================================================================================
class RingElem {
 public:
  using Host = ::ringsnark::seal::RingElem;

 private:
  mutable Host host_;                      // valid iff host_valid_
  mutable bool host_valid_ = true;
  std::shared_ptr<detail::DevRing> dev_;   // non-null: the value lives in HBM at element dev_idx_ of *dev_
  size_t dev_idx_ = 0;
  inline static bool bound_ = false;

  Host &mut() {                            // about to be modified on the host: materialise, then detach from HBM
    host();
    dev_.reset();
    return host_;
  }

 public:
  RingElem() = default;
  RingElem(const RingElem &) = default;
  RingElem(RingElem &&) = default;
  RingElem &operator=(const RingElem &) = default;
  RingElem &operator=(RingElem &&) = default;
  virtual ~RingElem() = default;
  RingElem(uint64_t value) : host_(value) {}
  explicit RingElem(const polytools::SealPoly &poly) : host_(poly) {}
  RingElem(const Host &h) : host_(h) {}
  static RingElem from_device(std::shared_ptr<detail::DevRing> dev, size_t idx) {
    RingElem r;
    r.host_valid_ = false;
    r.dev_ = std::move(dev);
    r.dev_idx_ = idx;
    return r;
  }

  // The host value (downloads once if the element was produced on the GPU).
  const Host &host() const {
    if (!host_valid_) {
      std::lock_guard<std::mutex> g(detail::host_mutex());
      if (!host_valid_) {
        auto rp = Host::get_context().first_context_data()->parms();
        std::vector<uint64_t> w(rp.poly_modulus_degree() * rp.coeff_modulus().size());
        detail::check(rsg_ringvec_download(dev_->v, dev_idx_, 1, w.data()));
        host_ = Host(polytools::SealPoly(Host::get_context(), w, &Host::get_context().first_parms_id()));
        host_valid_ = true;
      }
    }
    return host_;
  }
  const std::shared_ptr<detail::DevRing> &device_vector() const { return dev_; }
  size_t device_index() const { return dev_idx_; }

  // [L_R][N_R] words of the element as to_poly() would give them (seal_ring.tcc:265-277), appended to `out`.
  void append_words(std::vector<uint64_t> &out) const {
    const Host &h = host();
    if (h.is_scalar()) {
      Host tmp(h);
      tmp.to_poly_inplace();
      auto &p = tmp.get_poly();
      for (size_t j = 0; j < p.get_coeff_modulus_count(); j++) {
        auto limb = p.get_limb(j);
        out.insert(out.end(), limb.begin(), limb.end());
      }
    } else {
      auto p = h.get_poly();
      for (size_t j = 0; j < p.get_coeff_modulus_count(); j++) {
        auto limb = p.get_limb(j);
        out.insert(out.end(), limb.begin(), limb.end());
      }
    }
  }

  /* Static (seal_ring.hpp:52-118) */
  static void set_context(::seal::SEALContext &context_) {
    if (bound_) throw std::invalid_argument("cannot re-set context once set");
    try {
      Host::set_context(context_);
    } catch (const std::invalid_argument &) {
      // the SEAL backend of this process already holds a ring context: share it
    }
    bound_ = true;
  }
  static ::seal::SEALContext &get_context() {
    if (!bound_) throw std::invalid_argument("context not set");
    return Host::get_context();
  }
  static RingElem one() { return RingElem(1); }
  static RingElem zero() { return RingElem(0); }
  static RingElem random_exceptional_element(const std::shared_ptr<evaluation_domain<RingElem>> domain = nullptr) {
    std::shared_ptr<evaluation_domain<Host>> hd;
    if (domain) hd = std::make_shared<evaluation_domain<Host>>(domain->m);
    return RingElem(Host::random_exceptional_element(hd));
  }
  static RingElem random_element() { return RingElem(Host::random_element()); }
  static RingElem random_invertible_element() { return RingElem(Host::random_invertible_element()); }
  static RingElem random_nonzero_element() { return RingElem(Host::random_nonzero_element()); }

  /* Members (seal_ring.hpp:123-182) */
  [[nodiscard]] size_t size_in_bits() const { return host().size_in_bits(); }
  [[nodiscard]] bool is_zero() const {
    if (!host_valid_ && dev_) return dev_->zero[dev_idx_] != 0;   // same prefix test, done on the device
    return host().is_zero();
  }
  [[nodiscard]] bool fast_is_zero() const { return host_valid_ ? host_.fast_is_zero() : false; }
  [[nodiscard]] bool is_poly() const { return host_valid_ ? host_.is_poly() : true; }
  [[nodiscard]] bool is_scalar() const { return host_valid_ ? host_.is_scalar() : false; }
  void negate_inplace() { mut().negate_inplace(); }
  RingElem operator-() const {
    RingElem res(*this);
    res.negate_inplace();
    return res;
  }
  [[nodiscard]] bool is_invertible() const noexcept {
    try {
      return host().is_invertible();
    } catch (...) {
      return false;
    }
  }
  void invert_inplace() { mut().invert_inplace(); }
  [[nodiscard]] RingElem inverse() const {
    RingElem res(*this);
    res.invert_inplace();
    return res;
  }
  RingElem &operator+=(const RingElem &o) { mut() += o.host(); return *this; }
  RingElem &operator-=(const RingElem &o) { mut() -= o.host(); return *this; }
  RingElem &operator*=(const RingElem &o) { mut() *= o.host(); return *this; }
  RingElem &operator/=(const RingElem &o) { mut() /= o.host(); return *this; }
  RingElem &to_poly_inplace() { mut().to_poly_inplace(); return *this; }
  [[nodiscard]] RingElem to_poly() const {
    RingElem res(*this);
    res.to_poly_inplace();
    return res;
  }
  [[nodiscard]] size_t hash() const { return host().hash(); }
  using invalid_ring_elem_types = Host::invalid_ring_elem_types;
  [[nodiscard]] uint64_t get_scalar() const { return host().get_scalar(); }
  [[nodiscard]] polytools::SealPoly get_poly() const { return host().get_poly(); }
  [[nodiscard]] polytools::SealPoly &get_poly() { return mut().get_poly(); }
};

inline RingElem operator+(const RingElem &l, const RingElem &r) { RingElem x(l); x += r; return x; }
inline RingElem operator-(const RingElem &l, const RingElem &r) { RingElem x(l); x -= r; return x; }
inline RingElem operator*(const RingElem &l, const RingElem &r) { RingElem x(l); x *= r; return x; }
inline RingElem operator/(const RingElem &l, const RingElem &r) { RingElem x(l); x /= r; return x; }
inline bool operator==(const RingElem &l, const RingElem &r) { return l.host() == r.host(); }
inline bool operator!=(const RingElem &l, const RingElem &r) { return !(l == r); }
inline std::ostream &operator<<(std::ostream &out, const RingElem &e) { return out << e.host(); }

namespace detail {
// The backend ring-only callers use (interpolate<RingElem> before any EncodingElem::set_context, as in the reference's
// util/interpolation_test.cpp:87-91): the full backend when it exists, else a device context built from the ring parameters
// alone (N_E = N_R, one stand-in encoding limb) -- the witness kernels only touch the ring primes.
inline Backend &ring_backend() {
  Backend &full = backend_storage();
  if (full.ctx) return full;
  static Backend rb;
  std::lock_guard<std::mutex> g(host_mutex());
  if (!rb.ctx) {
    auto rp = RingElem::get_context().first_context_data()->parms();
    rb.N_R = rb.N_E = rp.poly_modulus_degree();
    rb.L_R = rp.coeff_modulus().size();
    rb.L_E = 1;
    for (auto &m : rp.coeff_modulus()) rb.q.push_back(m.value());
    rb.Q.push_back(rb.q[0]);
    rb.ring_words = rb.N_R * rb.L_R;
    rb.enc_words = rb.L_R * 2 * rb.N_E;
    const char *dev = std::getenv("RSG_DEVICE");
    check(rsg_context_create(&rb.ctx, rb.N_R, rb.L_R, rb.q.data(), rb.N_E, rb.L_E, rb.Q.data(), dev ? std::atoi(dev) : 0));
  }
  return rb;
}
}  // namespace detail

// =====================================================================================================================
class EncodingElem {
 public:
  using SealEnc = ::ringsnark::seal::EncodingElem;
  using PublicKey = SealEnc::PublicKey;
  using SecretKey = SealEnc::SecretKey;
  using decoding_error = SealEnc::decoding_error;

 private:
  std::shared_ptr<detail::DevEnc> arena_;   // null: empty (additive identity, seal_ring.tcc:482-488) or zero_
  size_t idx_ = 0;
  bool zero_ = false;                       // the size-0 "zero ciphertexts" operator*= assigns for r == 0 (:514-523)

  uint64_t *dptr() const { return rsg_crs_device_ptr(arena_->c) + idx_ * detail::backend().enc_words; }
  void make_unique() {                      // value semantics: never write into an arena somebody else can see
    if (arena_ && (arena_.use_count() > 1 || arena_->n > 1)) {
      auto fresh = detail::new_arena(1);
      detail::check(rsg_crs_copy(fresh->c, 0, arena_->c, idx_, 1));
      arena_ = fresh;
      idx_ = 0;
    }
  }
  static void init_backend() {
    detail::Backend &b = detail::backend_storage();
    if (b.ctx) throw std::invalid_argument("cannot re-set contexts once set");
    auto &ring = RingElem::get_context();
    auto rp = ring.first_context_data()->parms();
    auto &ctxs = SealEnc::get_contexts();
    auto ep = ctxs[0].first_context_data()->parms();
    b.N_R = rp.poly_modulus_degree();
    b.L_R = rp.coeff_modulus().size();
    b.N_E = ep.poly_modulus_degree();
    b.L_E = ep.coeff_modulus().size();
    for (auto &m : rp.coeff_modulus()) b.q.push_back(m.value());
    for (auto &m : ep.coeff_modulus()) {
      b.Q.push_back(m.value());
      b.enc_bits += (size_t)m.bit_count();
    }
    if (ctxs.size() != b.L_R) throw std::invalid_argument("one encoding context per ring limb expected");
    for (size_t j = 0; j < b.L_R; j++)
      if (ctxs[j].first_context_data()->parms().plain_modulus().value() != b.q[j])
        throw std::invalid_argument("encoding context j must use plain modulus q_j");
    b.ring_words = b.N_R * b.L_R;
    b.enc_words = b.L_R * 2 * b.L_E * b.N_E;
    const char *dev = std::getenv("RSG_DEVICE");
    detail::check(rsg_context_create(&b.ctx, b.N_R, b.L_R, b.q.data(), b.N_E, b.L_E, b.Q.data(), dev ? std::atoi(dev) : 0));
  }

 public:
  EncodingElem() = default;
  EncodingElem(const EncodingElem &) = default;
  EncodingElem &operator=(const EncodingElem &) = default;
  EncodingElem(std::shared_ptr<detail::DevEnc> arena, size_t idx) : arena_(std::move(arena)), idx_(idx) {}
  [[nodiscard]] bool is_empty() const { return !arena_ && !zero_; }
  [[nodiscard]] bool is_zero_ciphertext() const { return zero_; }
  // (arena handle, index) of this encoding -- null handle for empty / zero encodings; used by the fused provers
  [[nodiscard]] const rsg_crs *arena_handle() const { return arena_ ? arena_->c : nullptr; }
  [[nodiscard]] size_t arena_index() const { return idx_; }
  // true when [begin, end) are consecutive encodings of ONE arena (what one encode() call returns)
  static bool contiguous(std::vector<EncodingElem>::const_iterator begin, std::vector<EncodingElem>::const_iterator end, rsg_crs_ref *ref) {
    ref->crs = nullptr;
    ref->first = 0;
    if (begin == end) return true;
    if (!begin->arena_) return false;
    for (auto it = begin; it != end; ++it)
      if (it->arena_ != begin->arena_ || it->idx_ != begin->idx_ + (size_t)(it - begin)) return false;
    ref->crs = begin->arena_->c;
    ref->first = begin->idx_;
    return true;
  }

  /* Static (seal_ring.hpp:254-341) */
  static std::tuple<PublicKey, SecretKey> keygen() { return SealEnc::keygen(); }
  static void set_context(size_t N = 0) {
    bool have = true;
    try {
      SealEnc::get_contexts();
    } catch (const std::invalid_argument &) {
      have = false;
    }
    if (!have) SealEnc::set_context(N);
    init_backend();
  }
  static void set_contexts(const std::vector<::seal::SEALContext> &contexts_) {
    bool have = true;
    try {
      SealEnc::get_contexts();
    } catch (const std::invalid_argument &) {
      have = false;
    }
    if (!have) SealEnc::set_contexts(contexts_);
    init_backend();
  }
  static std::vector<::seal::SEALContext> &get_contexts() {
    detail::backend();
    return SealEnc::get_contexts();
  }

  // One encoding as host words [L_R][2][L_E][N_E]; and back to / from the reference's type (setup / verify side).
  std::vector<uint64_t> words() const {
    auto &b = detail::backend();
    std::vector<uint64_t> w(b.enc_words, 0);
    if (arena_) detail::check(rsg_crs_download(arena_->c, idx_, 1, w.data()));
    return w;
  }
  SealEnc to_seal() const {
    if (is_empty()) return SealEnc();
    auto &b = detail::backend();
    auto &ctxs = SealEnc::get_contexts();
    std::vector<::seal::Ciphertext> cts;
    std::vector<uint64_t> w;
    if (!zero_) w = words();
    for (size_t j = 0; j < b.L_R; j++) {
      ::seal::Ciphertext ct(ctxs[j], ctxs[j].first_parms_id());
      ct.is_ntt_form() = true;
      if (!zero_) {
        ct.resize(ctxs[j], ctxs[j].first_parms_id(), 2);
        std::memcpy(ct.data(), w.data() + j * 2 * b.L_E * b.N_E, 2 * b.L_E * b.N_E * sizeof(uint64_t));
      }
      cts.push_back(std::move(ct));
    }
    return SealEnc(cts);
  }
  static std::vector<EncodingElem> from_seal(const std::vector<SealEnc> &encs) {
    auto &b = detail::backend();
    std::vector<EncodingElem> out(encs.size());
    if (encs.empty()) return out;
    auto arena = detail::new_arena(encs.size());
    std::vector<uint64_t> w(b.enc_words);
    const size_t per = 2 * b.L_E * b.N_E;
    for (size_t i = 0; i < encs.size(); i++) {
      if (encs[i].is_empty()) continue;
      const auto &cts = detail::SealEncAccess::cts(encs[i]);
      bool all_zero_size = true;
      for (size_t j = 0; j < b.L_R; j++) {
        if (cts[j].size() == 0) {
          std::memset(w.data() + j * per, 0, per * sizeof(uint64_t));
          continue;
        }
        all_zero_size = false;
        if (cts[j].size() != 2 || cts[j].coeff_modulus_size() != b.L_E || cts[j].poly_modulus_degree() != b.N_E ||
            !cts[j].is_ntt_form())
          throw std::invalid_argument("seal_gpu: only fresh first-level NTT-form ciphertexts of size 2 are supported");
        std::memcpy(w.data() + j * per, cts[j].data(), per * sizeof(uint64_t));
      }
      if (all_zero_size) {
        out[i].zero_ = true;
        continue;
      }
      detail::check(rsg_crs_upload(arena->c, i, 1, w.data()));
      out[i] = EncodingElem(arena, i);
    }
    return out;
  }

  // seal_ring.tcc:324-359 on the GPU: batch encode + symmetric BGV encryption by rsg_encode, straight
  // into ONE HBM arena -- the CRS never exists on the host.  Randomness is SEAL's: per (element, ring limb) the 64-byte seed the
  // context's random generator factory would hand out (randomgen.h:440-448 -- fresh system randomness, or the fixed seed of a
  // seeded factory, in which case the ciphertext words equal SEAL's bit for bit).  RSG_ENCODE=seal keeps SEAL's own path.
  static std::vector<EncodingElem> encode(const SecretKey &sk, const std::vector<RingElem> &rs) {
    auto &b = detail::backend();
    const char *mode = std::getenv("RSG_ENCODE");
    if (mode && std::string(mode) == "seal") {
      std::vector<RingElem::Host> hosts;
      hosts.reserve(rs.size());
      for (const auto &r : rs) hosts.push_back(r.host());
      return from_seal(SealEnc::encode(sk, hosts));
    }
    std::vector<EncodingElem> out(rs.size());
    if (rs.empty()) return out;
    if (sk.size() != b.L_R) throw std::invalid_argument("one secret key per ring limb expected");
    std::vector<uint64_t> skw;
    skw.reserve(b.L_R * b.L_E * b.N_E);
    for (size_t j = 0; j < b.L_R; j++) skw.insert(skw.end(), sk[j].data().data(), sk[j].data().data() + b.L_E * b.N_E);
    std::vector<uint64_t> w;
    w.reserve(rs.size() * b.ring_words);
    for (const auto &r : rs) r.append_words(w);   // scalars as polynomials with every slot set (seal_ring.tcc:343-344)
    auto ring = std::make_shared<detail::DevRing>();
    detail::check(rsg_ringvec_create(b.ctx, rs.size(), &ring->v));
    detail::check(rsg_ringvec_upload(ring->v, 0, rs.size(), w.data()));
    auto &ctxs = SealEnc::get_contexts();
    std::vector<uint64_t> seeds(rs.size() * b.L_R * 8);
    for (size_t j = 0; j < b.L_R; j++) {
      auto factory = ctxs[j].first_context_data()->parms().random_generator();
      if (!factory) factory = ::seal::UniformRandomGeneratorFactory::DefaultFactory();
      for (size_t i = 0; i < rs.size(); i++) {
        ::seal::prng_seed_type seed;
        if (factory->use_random_seed()) ::seal::random_bytes(reinterpret_cast<::seal::seal_byte *>(seed.data()), ::seal::prng_seed_byte_count);
        else seed = factory->default_seed();
        std::copy(seed.begin(), seed.end(), seeds.begin() + (i * b.L_R + j) * 8);
      }
    }
    auto arena = detail::new_arena(rs.size());
    detail::check(rsg_encode(b.ctx, skw.data(), ring->v, 0, rs.size(), seeds.data(), arena->c, 0));
    for (size_t i = 0; i < rs.size(); i++) out[i] = EncodingElem(arena, i);
    return out;
  }
  // seal_ring.tcc:435-477 on the GPU: noise budget, c0 + c1 s, exact base conversion q -> t, batch
  // decode -- rsg_decode, bit-identical to SEAL's Decryptor + BatchEncoder.  Empty / zero encodings (no arena) keep the
  // reference's own handling.
  static RingElem decode(const SecretKey &sk, const EncodingElem &e) {
    if (!e.arena_) return RingElem(SealEnc::decode(sk, e.to_seal()));
    auto &b = detail::backend();
    if (sk.size() != b.L_R) throw std::invalid_argument("one secret key per ring limb expected");
    std::vector<uint64_t> skw;
    skw.reserve(b.L_R * b.L_E * b.N_E);
    for (size_t j = 0; j < b.L_R; j++) skw.insert(skw.end(), sk[j].data().data(), sk[j].data().data() + b.L_E * b.N_E);
    std::vector<uint64_t> w(b.ring_words);
    std::vector<int32_t> budget(b.L_R, 0);
    const int rc = rsg_decode(b.ctx, skw.data(), e.dptr(), nullptr, 1, w.data(), budget.data());
    if (rc == RSG_ERR_NOISE)
      for (size_t j = 0; j < b.L_R; j++)
        if (budget[j] <= 0)
          throw decoding_error("ciphertext #" + std::to_string(j) + " has remaining noise budget " + std::to_string(budget[j]) + " <= 0");
    detail::check(rc);
    return RingElem(polytools::SealPoly(RingElem::get_context(), w, &RingElem::get_context().first_parms_id()));
  }

  // seal_ring.tcc:361-433 on the GPU.
  static EncodingElem inner_product(std::vector<EncodingElem>::const_iterator a_start,
                                    std::vector<EncodingElem>::const_iterator a_end,
                                    std::vector<RingElem>::const_iterator b_start,
                                    std::vector<RingElem>::const_iterator b_end) {
    auto &b = detail::backend();
    const size_t count = (size_t)(a_end - a_start);
    if ((size_t)(b_end - b_start) < count) throw std::invalid_argument("inner_product: mismatched sizes");
    // the reference's per-term dispatch (is_zero -> skipped, seal_ring.tcc:390-396,416; scalar 1 -> ciphertext taken
    // unchanged, :525-528; everything else batch-encoded and multiplied, :533-543)
    std::vector<uint8_t> tags(count);
    std::shared_ptr<detail::DevRing> dev = count ? (b_start)->device_vector() : nullptr;
    bool all_dev = (bool)dev;
    for (size_t i = 0; i < count; i++) {
      const RingElem &r = *(b_start + i);
      const EncodingElem &a = *(a_start + i);
      if (r.is_zero() || !a.arena_) {
        tags[i] = RSG_TERM_SKIP;
      } else if (r.is_scalar() && r.get_scalar() == 1) {
        tags[i] = RSG_TERM_ONE;
      } else {
        tags[i] = RSG_TERM_GENERAL;
      }
      if (r.device_vector() != dev) all_dev = false;
    }
    // coefficients: device-resident (witness-map output) as they are; host values are staged into one upload
    std::vector<uint32_t> cidx(count, 0);
    std::shared_ptr<detail::DevRing> staged;
    if (all_dev) {
      for (size_t i = 0; i < count; i++) cidx[i] = (uint32_t)(b_start + i)->device_index();
    } else {
      std::vector<uint64_t> w;
      size_t g = 0;
      for (size_t i = 0; i < count; i++)
        if (tags[i] == RSG_TERM_GENERAL) {
          (b_start + i)->append_words(w);
          cidx[i] = (uint32_t)g++;
        }
      staged = std::make_shared<detail::DevRing>();
      detail::check(rsg_ringvec_create(b.ctx, g ? g : 1, &staged->v));
      if (g) detail::check(rsg_ringvec_upload(staged->v, 0, g, w.data()));
      dev = staged;
    }
    // CRS side: group the terms by arena (one group when the range comes from one encode() call)
    EncodingElem res;
    std::vector<uint8_t> done(count, 0);
    for (size_t first = 0; first < count; first++) {
      if (done[first] || tags[first] == RSG_TERM_SKIP) continue;
      const auto &arena = (a_start + first)->arena_;
      std::vector<uint32_t> ci, ri;
      std::vector<uint8_t> tg;
      for (size_t i = first; i < count; i++)
        if (!done[i] && tags[i] != RSG_TERM_SKIP && (a_start + i)->arena_ == arena) {
          ci.push_back((uint32_t)(a_start + i)->idx_);
          ri.push_back(cidx[i]);
          tg.push_back(tags[i]);
          done[i] = 1;
        }
      auto out = detail::new_arena(1);
      size_t used = 0;
      detail::check(rsg_inner_product_idx(b.ctx, arena->c, ci.data(), dev->v, ri.data(), ci.size(), tg.data(), nullptr,
                                          rsg_crs_device_ptr(out->c), &used));
      res += EncodingElem(out, 0);
    }
    return res;   // empty when every term was skipped, like the reference's default-constructed `res`
  }

  /* Members (seal_ring.hpp:346-373) */
  [[nodiscard]] size_t size_in_bits() const {
    if (!arena_) return 0;
    auto &b = detail::backend();
    return b.L_R * b.N_E * 2 * b.enc_bits;
  }
  [[nodiscard]] static size_t size_in_bits_pk(const PublicKey &pk) { return SealEnc::size_in_bits_pk(pk); }
  [[nodiscard]] static size_t size_in_bits_sk(const SecretKey &sk) { return SealEnc::size_in_bits_sk(sk); }

  // seal_ring.tcc:479-507 (SEAL's add_inplace treats the size-0 zero ciphertext as the identity as well)
  EncodingElem &operator+=(const EncodingElem &other) {
    if (other.is_empty() || other.zero_) return *this;
    if (!arena_) {
      *this = other;
      return *this;
    }
    make_unique();
    detail::check(rsg_enc_add(detail::backend().ctx, dptr(), other.dptr()));
    return *this;
  }
  // seal_ring.tcc:509-548
  EncodingElem &operator*=(const RingElem &r) {
    if (r.is_zero()) {
      arena_.reset();
      zero_ = true;
      return *this;
    }
    if (!arena_) return *this;                                  // empty stays empty, zero stays zero
    if (r.is_scalar() && r.get_scalar() == 1) return *this;
    std::vector<EncodingElem> a{*this};
    std::vector<RingElem> bvec{r};
    *this = inner_product(a.begin(), a.end(), bvec.begin(), bvec.end());
    return *this;
  }
  friend bool operator==(const EncodingElem &lhs, const EncodingElem &rhs);
};

inline EncodingElem operator+(const EncodingElem &l, const EncodingElem &r) { EncodingElem x(l); x += r; return x; }
inline EncodingElem operator*(const EncodingElem &l, const RingElem &r) { EncodingElem x(l); x *= r; return x; }
inline EncodingElem operator*(const RingElem &l, const EncodingElem &r) { EncodingElem x(r); x *= l; return x; }
// seal_ring.hpp:391-409 compares ciphertext counts first (so `proof.F == EncT()` is "is F empty"), then contents.
inline bool operator==(const EncodingElem &lhs, const EncodingElem &rhs) {
  if (lhs.is_empty() || rhs.is_empty()) return lhs.is_empty() && rhs.is_empty();
  if (lhs.zero_ || rhs.zero_) return lhs.zero_ && rhs.zero_;
  return lhs.words() == rhs.words();
}
}  // namespace ringsnark::seal_gpu

namespace std {
template <>
struct hash<ringsnark::seal_gpu::RingElem> {
  size_t operator()(const ringsnark::seal_gpu::RingElem &r) const { return r.hash(); }
};
}  // namespace std

// =====================================================================================================================
// Hot path (a): the QRP witness map on the GPU.  Explicit specialisation of the reference's function template
// (reductions/r1cs_to_qrp/r1cs_to_qrp.hpp:46-51) for this backend's ring type; groth16::prover (groth16.tcc:82-84)
// and rinocchio::prover (rinocchio.tcc:92-93) pick it up without being edited.
namespace ringsnark {
template <>
inline qrp_witness<seal_gpu::RingElem> r1cs_to_qrp_witness_map<seal_gpu::RingElem>(
    const r1cs_constraint_system<seal_gpu::RingElem> &cs, const r1cs_primary_input<seal_gpu::RingElem> &primary_input,
    const r1cs_auxiliary_input<seal_gpu::RingElem> &auxiliary_input, const seal_gpu::RingElem &d1,
    const seal_gpu::RingElem &d2, const seal_gpu::RingElem &d3) {
  using R = seal_gpu::RingElem;
  namespace D = seal_gpu::detail;
  auto &b = D::backend();
  const size_t n = cs.num_constraints(), n_io = primary_input.size(), n_aux = auxiliary_input.size(), W = b.ring_words;
  if (n_io != cs.primary_input_size || n_aux != cs.auxiliary_input_size)
    throw std::invalid_argument("assignment does not match the constraint system");

  // the nine evaluation vectors of r1cs_to_qrp.tcc:167-223 -> HBM
  auto make_vec = [&](size_t count) {
    auto v = std::make_shared<D::DevRing>();
    D::check(rsg_ringvec_create(b.ctx, count ? count : 1, &v->v));
    return v;
  };
  auto evals = make_vec(9 * n);
  bool scalar_coeffs = true;
  std::vector<uint32_t> row_ptr{0}, col;
  std::vector<uint64_t> coeff;
  for (int m = 0; m < 3 && scalar_coeffs; m++)
    for (size_t i = 0; i < n && scalar_coeffs; i++) {
      const auto &lc = m == 0 ? cs.constraints[i].a : (m == 1 ? cs.constraints[i].b : cs.constraints[i].c);
      for (const auto &lt : lc.terms) {
        if (!lt.coeff.is_scalar()) {
          scalar_coeffs = false;
          break;
        }
        col.push_back((uint32_t)lt.index);
        coeff.push_back(lt.coeff.get_scalar());
      }
      row_ptr.push_back((uint32_t)col.size());
    }
  rsg_r1cs *r1cs = nullptr;   // non-null: the evaluations were produced on the device from this CSR system
  struct R1csGuard {
    rsg_r1cs *&r;
    ~R1csGuard() { rsg_r1cs_destroy(r); }
  } guard{r1cs};
  if (scalar_coeffs) {
    // integer coefficients (every reference driver except the NTT demo): sparse evaluate on the device
    std::vector<uint64_t> w;
    w.reserve((n_io + n_aux) * W);
    for (const auto &r : primary_input) r.append_words(w);
    for (const auto &r : auxiliary_input) r.append_words(w);
    auto assignment = make_vec(n_io + n_aux);
    if (n_io + n_aux) D::check(rsg_ringvec_upload(assignment->v, 0, n_io + n_aux, w.data()));
    D::check(rsg_r1cs_create(b.ctx, n, n_io, n_aux, row_ptr.data(), col.data(), coeff.data(), &r1cs));
    D::check(rsg_r1cs_evaluate(b.ctx, r1cs, assignment->v, evals->v));
  } else {
    // ring-element coefficients: linear_combination::evaluate as written (relations/variable.tcc:246-254), then upload
    r1cs_variable_assignment<R> mid(n_io, R::zero()), io(primary_input), full(primary_input);
    mid.insert(mid.end(), auxiliary_input.begin(), auxiliary_input.end());
    io.insert(io.end(), n_aux, R::zero());
    full.insert(full.end(), auxiliary_input.begin(), auxiliary_input.end());
    const r1cs_variable_assignment<R> *as[3] = {&mid, &io, &full};
    std::vector<uint64_t> w;
    w.reserve(9 * n * W);
    for (int v = 0; v < 3; v++)
      for (int m = 0; m < 3; m++)
        for (size_t i = 0; i < n; i++) {
          const auto &lc = m == 0 ? cs.constraints[i].a : (m == 1 ? cs.constraints[i].b : cs.constraints[i].c);
          lc.evaluate(*as[v]).append_words(w);
        }
    D::check(rsg_ringvec_upload(evals->v, 0, 9 * n, w.data()));
  }

  // interpolation, product, division by Z, zero-knowledge patch: all on the device
  auto coeffs = make_vec(6 * n), H = make_vec(n + 1);
  std::vector<uint64_t> dw;
  const bool zk = !(d1.is_zero() && d2.is_zero() && d3.is_zero());
  if (zk) {
    d1.append_words(dw);
    d2.append_words(dw);
    d3.append_words(dw);
  }
  if (r1cs) D::check(rsg_witness_map_r1cs(b.ctx, r1cs, evals->v, zk ? dw.data() : nullptr, coeffs->v, H->v));
  else D::check(rsg_witness_map_zk(b.ctx, n, evals->v, zk ? dw.data() : nullptr, coeffs->v, H->v));
  coeffs->zero.resize(6 * n);
  H->zero.resize(n + 1);
  D::check(rsg_ringvec_is_zero_prefix(coeffs->v, 0, 6 * n, coeffs->zero.data()));
  D::check(rsg_ringvec_is_zero_prefix(H->v, 0, n + 1, H->zero.data()));
  auto slice = [&](const std::shared_ptr<D::DevRing> &v, size_t first, size_t count) {
    std::vector<R> out;
    out.reserve(count);
    for (size_t i = 0; i < count; i++) out.push_back(R::from_device(v, first + i));
    return out;
  };
  // Z: per-prime constants; the leading coefficient is the scalar 1 exactly as Boost's product of (x - i) leaves it
  // (evaluation_domain.tcc:53-60), every other coefficient has been through a negation and is a polynomial.
  std::vector<uint64_t> hZ(b.L_R * (n + 1));
  D::check(rsg_vanishing(b.ctx, n, hZ.data()));
  std::vector<R> Z;
  Z.reserve(n + 1);
  for (size_t k = 0; k < n; k++) {
    std::vector<uint64_t> w(W);
    for (size_t j = 0; j < b.L_R; j++) std::fill(w.begin() + j * b.N_R, w.begin() + (j + 1) * b.N_R, hZ[j * (n + 1) + k]);
    Z.push_back(R(polytools::SealPoly(R::get_context(), w, &R::get_context().first_parms_id())));
  }
  Z.push_back(R::one());

  r1cs_variable_assignment<R> full_variable_assignment(primary_input);
  full_variable_assignment.insert(full_variable_assignment.end(), auxiliary_input.begin(), auxiliary_input.end());
  // coefficient order in HBM: A_io, B_io, C_io, A_mid, B_mid, C_mid (include/rsgpu.h, rsg_witness_map)
  return qrp_witness<R>(cs.num_variables(), n, cs.num_inputs(), d1, d2, d3, full_variable_assignment, slice(coeffs, 0, n),
                        slice(coeffs, n, n), slice(coeffs, 2 * n, n), slice(coeffs, 3 * n, n), slice(coeffs, 4 * n, n),
                        slice(coeffs, 5 * n, n), Z, slice(H, 0, n + 1));
}
// =====================================================================================================================
// the instance map with evaluation on the GPU.  Explicit specialisation of
// reductions/r1cs_to_qrp/r1cs_to_qrp.hpp:48-50; generator (groth16.tcc:11-12, rinocchio.tcc:12-13) and verifier
// (groth16.tcc:127-128, rinocchio.tcc:220-221) pick it up unedited.  Falls back to the reference's loop (as written,
// r1cs_to_qrp.tcc:75-116) when a linear-term coefficient is a ring element rather than an integer.
template <>
inline qrp_instance_evaluation<seal_gpu::RingElem> r1cs_to_qrp_instance_map_with_evaluation<seal_gpu::RingElem>(
    const r1cs_constraint_system<seal_gpu::RingElem> &cs, const seal_gpu::RingElem &t) {
  using R = seal_gpu::RingElem;
  namespace D = seal_gpu::detail;
  auto &b = D::backend();
  const auto domain = get_evaluation_domain<R>(cs.num_constraints());
  const size_t n = cs.num_constraints(), nv1 = cs.num_variables() + 1;
  // evaluate_all_lagrange_polynomials' guard (evaluation_domain.tcc:22-24)
  for (size_t i = 0; i < domain->m; i++)
    if (domain->get_domain_element(i) == t) throw std::invalid_argument("t cannot be one of the values in the domain");
  bool scalar_coeffs = true;
  std::vector<uint32_t> row_ptr{0}, col;
  std::vector<uint64_t> coeff;
  for (int m = 0; m < 3 && scalar_coeffs; m++)
    for (size_t i = 0; i < n && scalar_coeffs; i++) {
      const auto &lc = m == 0 ? cs.constraints[i].a : (m == 1 ? cs.constraints[i].b : cs.constraints[i].c);
      for (const auto &lt : lc.terms) {
        if (!lt.coeff.is_scalar()) {
          scalar_coeffs = false;
          break;
        }
        col.push_back((uint32_t)lt.index);
        coeff.push_back(lt.coeff.get_scalar());
      }
      row_ptr.push_back((uint32_t)col.size());
    }
  if (!scalar_coeffs) {
    std::vector<R> At(nv1, R::zero()), Bt(nv1, R::zero()), Ct(nv1, R::zero()), Ht;
    const R Zt = domain->compute_vanishing_polynomial(t);
    const std::vector<R> u = domain->evaluate_all_lagrange_polynomials(t);
    for (size_t i = 0; i < n; ++i) {
      for (const auto &lt : cs.constraints[i].a.terms) At[lt.index] += u[i] * lt.coeff;
      for (const auto &lt : cs.constraints[i].b.terms) Bt[lt.index] += u[i] * lt.coeff;
      for (const auto &lt : cs.constraints[i].c.terms) Ct[lt.index] += u[i] * lt.coeff;
    }
    R ti = R::one();
    for (size_t i = 0; i < domain->m + 1; ++i) {
      Ht.emplace_back(ti);
      ti *= t;
    }
    return qrp_instance_evaluation<R>(domain, cs.num_variables(), domain->m, cs.num_inputs(), t, std::move(At), std::move(Bt),
                                      std::move(Ct), std::move(Ht), Zt);
  }
  auto make_vec = [&](size_t count) {
    auto v = std::make_shared<D::DevRing>();
    D::check(rsg_ringvec_create(b.ctx, count ? count : 1, &v->v));
    return v;
  };
  rsg_r1cs *r1cs = nullptr;
  struct R1csGuard {
    rsg_r1cs *&r;
    ~R1csGuard() { rsg_r1cs_destroy(r); }
  } guard{r1cs};
  D::check(rsg_r1cs_create(b.ctx, n, cs.primary_input_size, cs.auxiliary_input_size, row_ptr.data(), col.data(), coeff.data(), &r1cs));
  std::vector<uint64_t> tw;
  t.append_words(tw);
  auto tv = make_vec(1), ABCt = make_vec(3 * nv1), Ht = make_vec(n + 1), Zt = make_vec(1);
  D::check(rsg_ringvec_upload(tv->v, 0, 1, tw.data()));
  D::check(rsg_instance_map(b.ctx, r1cs, tv->v, 0, ABCt->v, Ht->v, Zt->v));
  for (auto *v : {ABCt.get(), Ht.get(), Zt.get()}) {
    const size_t cnt = rsg_ringvec_size(v->v);
    v->zero.resize(cnt);
    D::check(rsg_ringvec_is_zero_prefix(v->v, 0, cnt, v->zero.data()));
  }
  auto slice = [&](const std::shared_ptr<D::DevRing> &v, size_t first, size_t count) {
    std::vector<R> out;
    out.reserve(count);
    for (size_t i = 0; i < count; i++) out.push_back(R::from_device(v, first + i));
    return out;
  };
  return qrp_instance_evaluation<R>(domain, cs.num_variables(), domain->m, cs.num_inputs(), t, slice(ABCt, 0, nv1),
                                    slice(ABCt, nv1, nv1), slice(ABCt, 2 * nv1, nv1), slice(Ht, 0, n + 1),
                                    R::from_device(Zt, 0));
}
}  // namespace ringsnark

// =====================================================================================================================
// The verifier's three interpolations (groth16.tcc:147-153, rinocchio.tcc: the same shape): explicit specialisation of the
// global function template interpolate (util/polynomials.hpp:17-18) for this backend's ring type.  On the evaluation
// domain {0..n-1} -- the only point set the proof systems use -- it runs on the GPU (rsg_interpolate: the quasi-linear or
// the dense kernels of the witness map); any other point set goes through the reference's own algorithm on the host
// type.  With the device instance map and the device decode this takes groth16::verifier at the logistic-regression shape
// (n = 1031) from minutes to well under a second.
template <>
inline std::vector<ringsnark::seal_gpu::RingElem> interpolate<ringsnark::seal_gpu::RingElem>(
    const std::vector<ringsnark::seal_gpu::RingElem> &x, const std::vector<ringsnark::seal_gpu::RingElem> &y) {
  using R = ringsnark::seal_gpu::RingElem;
  namespace D = ringsnark::seal_gpu::detail;
  const size_t n = x.size();
  if (y.size() != n) throw std::invalid_argument("interpolate: mismatched sizes");
  bool domain = n >= 1;
  for (size_t i = 0; i < n && domain; i++) domain = x[i].is_scalar() && x[i].get_scalar() == i;
  if (!domain) {
    std::vector<R::Host> hx, hy;
    hx.reserve(n);
    hy.reserve(n);
    for (const auto &e : x) hx.push_back(e.host());
    for (const auto &e : y) hy.push_back(e.host());
    const auto hc = interpolate<R::Host>(hx, hy);
    return std::vector<R>(hc.begin(), hc.end());
  }
  auto &b = D::ring_backend();
  auto make_vec = [&](size_t count) {
    auto v = std::make_shared<D::DevRing>();
    D::check(rsg_ringvec_create(b.ctx, count, &v->v));
    return v;
  };
  std::vector<uint64_t> w;
  w.reserve(n * b.ring_words);
  for (const auto &e : y) e.append_words(w);
  auto yv = make_vec(n), out = make_vec(n);
  D::check(rsg_ringvec_upload(yv->v, 0, n, w.data()));
  D::check(rsg_interpolate(b.ctx, n, 1, yv->v, 0, out->v, 0));
  out->zero.resize(n);
  D::check(rsg_ringvec_is_zero_prefix(out->v, 0, n, out->zero.data()));
  std::vector<R> coeffs;
  coeffs.reserve(n);
  for (size_t i = 0; i < n; i++) coeffs.push_back(R::from_device(out, i));
  return coeffs;
}

#include "provers.hpp"

#endif  // RINGSNARK_SEAL_GPU_RING_HPP
