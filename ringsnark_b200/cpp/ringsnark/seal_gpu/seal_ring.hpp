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
//   host: ring elements that drivers build (circuits, assignments) are host values; RingElem is this backend's OWN class -- the
//         scalar / polynomial variant rules (seal_ring.tcc:105-263), SEAL's modular routines and SealPoly::is_zero / is_equal's
//         byte-count quirks (depends/SEAL-Polytools/src/poly_arith.cpp:147-162) restated, checked operator by operator against
//         the reference's class (oracle/ringelem_check.cpp, tests/test_ringelem.py).  Operators on two HBM-resident operands run
//         on the GPU (rsg_ring_*).  EncodingElem::keygen and the decode of EMPTY encodings still call the reference's SEAL path;
//         EncodingElem::encode (rsg_encode) and decode (rsg_decode) run on the GPU.
//   Ring elements PRODUCED by the GPU witness map stay in HBM (a RingElem then holds a ref-counted slice of a device
//   vector plus its is_zero flag) and are fed to inner_product without a round trip; they are downloaded lazily only
//   if host code looks at them.  Encodings live in HBM arenas: one arena per encode() call, so a proving-key vector
//   is contiguous and a lincomb over an iterator range of it is one streaming pass.
#ifndef RINGSNARK_SEAL_GPU_RING_HPP
#define RINGSNARK_SEAL_GPU_RING_HPP

#include <atomic>
#include <cmath>
#include <cstring>
#include <memory>
#include <variant>
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
  rsg_context *ctx = nullptr;   // the device context the vector belongs to
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
// RingElem: the ring R_q = prod_j Z_{q_j}^{N_R} in slot (double-CRT) form, with the reference's value semantics
// (ringsnark/seal/seal_ring.hpp:18-214, seal_ring.tcc:5-302) restated over this backend's own storage -- no
// ringsnark::seal::RingElem, no polytools arithmetic:
//   * a SCALAR (uint64, "the same value in every slot", kept scalar while its bit length stays below q_1's,
//     seal_ring.tcc:126-147,220-239), or
//   * a POLYNOMIAL of L_R x N_R residues that lives on the host (copy-on-write word vector), in HBM (a slice of a device
//     vector the witness / instance map produced) or both.
// Element-wise operators on two HBM-resident operands run on the GPU (rsg_ring_binop / rsg_ring_scalar_op / rsg_ring_negate /
// rsg_ring_invert, csrc/ringops.cuh) and leave their result there; anything else runs on the host words with the same
// modular routines SEAL applies (util/uintarithsmallmod.h: add_uint_mod / sub_uint_mod take the scalar operand UNREDUCED with
// one conditional correction, multiply reduces it first; poly_arith.cpp:164-350).  SealPoly::is_zero / is_equal keep their
// byte-count quirks (poly_arith.cpp:147-162): only bytes [0, W + 7) resp. [0, W) of the W words are looked at.
namespace detail {
struct RingParams {
  ::seal::SEALContext *ctx = nullptr;
  size_t N = 0, L = 0;
  std::vector<uint64_t> q;
  std::vector<int> q_bits;
};
inline RingParams &ring_params_storage() {
  static RingParams p;
  return p;
}
inline const RingParams &ring_params() {
  const RingParams &p = ring_params_storage();
  if (!p.ctx) throw std::invalid_argument("context not set");
  return p;
}
inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t p) { return (uint64_t)(((unsigned __int128)a * b) % p); }
// 1 + floor(log2(x)) exactly as seal_ring.tcc:127-128,221-222 evaluates it (double arithmetic, then converted to size_t)
inline size_t bitsize_like_reference(uint64_t x) { return 1 + std::floor(std::log2(x)); }
}  // namespace detail

class RingElem {
 public:
  using Host = ::ringsnark::seal::RingElem;   // the reference's type: conversions only (harnesses, RSG_ENCODE=seal)
  using Words = std::vector<uint64_t>;

 protected:
  inline static std::shared_ptr<::seal::UniformRandomGenerator> prng = nullptr;   // random_element (seal_ring.hpp:27,90-102)

 private:
  bool poly_ = false;
  uint64_t scalar_ = 0;
  mutable std::shared_ptr<const Words> host_;   // polynomial words [L_R][N_R] once they exist on the host
  std::shared_ptr<detail::DevRing> dev_;        // non-null: the polynomial lives in HBM at element dev_idx_ of *dev_
  size_t dev_idx_ = 0;

  static std::shared_ptr<const Words> make_words(Words &&w) { return std::make_shared<const Words>(std::move(w)); }
  void set_poly(Words &&w) {
    poly_ = true;
    scalar_ = 0;
    dev_.reset();
    std::atomic_store(&host_, make_words(std::move(w)));
  }
  void set_scalar(uint64_t v) {
    poly_ = false;
    scalar_ = v;
    dev_.reset();
    std::atomic_store(&host_, std::shared_ptr<const Words>());
  }
  bool device_only() const { return poly_ && dev_ && !std::atomic_load(&host_); }
  // words of the polynomial "scalar in every slot" as SealPoly(context).add_scalar_inplace(scalar) leaves them (seal_ring.tcc:265-277)
  static Words scalar_words(uint64_t v) {
    const auto &rp = detail::ring_params();
    Words w(rp.N * rp.L, 0);
    if (v)
      for (size_t j = 0; j < rp.L; j++) std::fill(w.begin() + j * rp.N, w.begin() + (j + 1) * rp.N, v >= rp.q[j] ? v - rp.q[j] : v);
    return w;
  }
  template <class F>
  void map_words(F f) {   // w[j][i] <- f(w[j][i], q_j), copy-on-write
    const auto &rp = detail::ring_params();
    Words w(words());
    for (size_t j = 0; j < rp.L; j++)
      for (size_t i = 0; i < rp.N; i++) w[j * rp.N + i] = f(w[j * rp.N + i], rp.q[j]);
    set_poly(std::move(w));
  }
  template <class F>
  void zip_words(const Words &o, F f) {
    const auto &rp = detail::ring_params();
    Words w(words());
    for (size_t j = 0; j < rp.L; j++)
      for (size_t i = 0; i < rp.N; i++) w[j * rp.N + i] = f(w[j * rp.N + i], o[j * rp.N + i], rp.q[j]);
    set_poly(std::move(w));
  }
  // device-side operators (defined after detail::ring_backend): true when the operation ran in HBM
  bool dev_binop(int op, const RingElem &o);
  bool dev_scalar_op(int op, uint64_t v);
  bool dev_negate();
  bool dev_invert();

 public:
  RingElem() = default;
  RingElem(const RingElem &o) : poly_(o.poly_), scalar_(o.scalar_), host_(std::atomic_load(&o.host_)), dev_(o.dev_), dev_idx_(o.dev_idx_) {}
  RingElem(RingElem &&o) noexcept : poly_(o.poly_), scalar_(o.scalar_), host_(std::atomic_load(&o.host_)), dev_(std::move(o.dev_)), dev_idx_(o.dev_idx_) {}
  RingElem &operator=(const RingElem &o) {
    if (this != &o) {
      poly_ = o.poly_;
      scalar_ = o.scalar_;
      std::atomic_store(&host_, std::atomic_load(&o.host_));
      dev_ = o.dev_;
      dev_idx_ = o.dev_idx_;
    }
    return *this;
  }
  RingElem &operator=(RingElem &&o) noexcept { return *this = static_cast<const RingElem &>(o); }
  virtual ~RingElem() = default;
  RingElem(uint64_t value) : scalar_(value) {}
  explicit RingElem(const polytools::SealPoly &poly) : poly_(true) {
    polytools::SealPoly p(poly);   // get_limb is not const upstream
    Words w;
    for (size_t j = 0; j < p.get_coeff_modulus_count(); j++) {
      auto limb = p.get_limb(j);
      w.insert(w.end(), limb.begin(), limb.end());
    }
    host_ = make_words(std::move(w));
  }
  RingElem(const Host &h) {   // from the reference's type (harnesses that run both backends side by side)
    if (h.is_scalar()) scalar_ = h.get_scalar();
    else *this = RingElem(h.get_poly());
  }
  static RingElem from_words(Words &&w) {
    RingElem r;
    r.set_poly(std::move(w));
    return r;
  }
  static RingElem from_device(std::shared_ptr<detail::DevRing> dev, size_t idx) {
    RingElem r;
    r.poly_ = true;
    r.dev_ = std::move(dev);
    r.dev_idx_ = idx;
    return r;
  }

  // The polynomial's words (downloads once if the element was produced on the GPU; safe under concurrent readers).
  const Words &words() const {
    auto h = std::atomic_load(&host_);
    if (!h) {
      std::lock_guard<std::mutex> g(detail::host_mutex());
      h = std::atomic_load(&host_);
      if (!h) {
        const auto &rp = detail::ring_params();
        Words w(rp.N * rp.L);
        detail::check(rsg_ringvec_download(dev_->v, dev_idx_, 1, w.data()));
        h = make_words(std::move(w));
        std::atomic_store(&host_, h);
      }
    }
    return *h;
  }
  // The reference's representation of this value.
  Host host() const {
    if (!poly_) return Host(scalar_);
    auto &ctx = get_context();
    return Host(polytools::SealPoly(ctx, words(), &ctx.first_parms_id()));
  }
  const std::shared_ptr<detail::DevRing> &device_vector() const { return dev_; }
  size_t device_index() const { return dev_idx_; }

  // [L_R][N_R] words of the element as to_poly() would give them (seal_ring.tcc:265-277), appended to `out`.
  void append_words(std::vector<uint64_t> &out) const {
    if (poly_) {
      const Words &w = words();
      out.insert(out.end(), w.begin(), w.end());
    } else {
      const Words w = scalar_words(scalar_);
      out.insert(out.end(), w.begin(), w.end());
    }
  }

  /* Static (seal_ring.hpp:52-118) */
  static void set_context(::seal::SEALContext &context_) {
    auto &rp = detail::ring_params_storage();
    if (rp.ctx) throw std::invalid_argument("cannot re-set context once set");
    rp.ctx = new ::seal::SEALContext(context_);
    const auto parms = rp.ctx->first_context_data()->parms();
    rp.N = parms.poly_modulus_degree();
    rp.L = parms.coeff_modulus().size();
    for (const auto &m : parms.coeff_modulus()) {
      rp.q.push_back(m.value());
      rp.q_bits.push_back(m.bit_count());
    }
    try {   // harnesses that also instantiate the reference backend in this process share the ring ...
      Host::set_context(context_);
    } catch (const std::invalid_argument &) {   // ... which must then be the SAME ring
      if (Host::get_context().first_parms_id() != context_.first_parms_id()) {
        delete rp.ctx;
        rp = detail::RingParams();
        throw std::invalid_argument("the reference backend of this process is bound to another ring context");
      }
    }
  }
  static ::seal::SEALContext &get_context() { return *detail::ring_params().ctx; }
  static RingElem one() { return RingElem(1); }
  static RingElem zero() { return RingElem(0); }
  // seal_ring.hpp:70-88: rejection sampling of a scalar below q_1 (and above the domain) under the reference's mask
  static RingElem random_exceptional_element(const std::shared_ptr<evaluation_domain<RingElem>> domain = nullptr) {
    const uint64_t q1 = detail::ring_params().q[0];
    const uint64_t bit_width = 1ULL + (uint64_t)std::floor(std::log2l(q1));
    const uint64_t mask = (1 << (bit_width + 1)) - 1;   // an int shift, as upstream: the effective width is (bit_width + 1) mod 32
    uint64_t rand = ::seal::random_uint64() & mask;
    while (rand >= q1 || (domain && rand <= domain->m)) rand = ::seal::random_uint64() & mask;
    return RingElem(rand);
  }
  static RingElem random_element() {   // seal_ring.hpp:90-102: uniform residues from SEAL's sampler
    if (prng == nullptr) prng = ::seal::UniformRandomGeneratorFactory::DefaultFactory()->create();
    const auto &rp = detail::ring_params();
    Words w(rp.N * rp.L);
    ::seal::util::sample_poly_uniform(prng, rp.ctx->first_context_data()->parms(), w.data());
    return from_words(std::move(w));
  }
  static RingElem random_invertible_element() {
    RingElem res;
    do res = random_element();
    while (!res.is_invertible());
    return res;
  }
  static RingElem random_nonzero_element() {
    RingElem res;
    do res = random_element();
    while (res.is_zero());
    return res;
  }

  /* Members (seal_ring.hpp:123-182) */
  [[nodiscard]] size_t size_in_bits() const {
    if (!poly_) return 8 * sizeof(uint64_t);
    const auto &rp = detail::ring_params();
    size_t size = 0;
    for (int b : rp.q_bits) size += (size_t)b * rp.N;
    return size;
  }
  [[nodiscard]] bool is_zero() const;   // after detail::dev_zero_flag
  [[nodiscard]] bool fast_is_zero() const { return !poly_ && scalar_ == 0; }
  [[nodiscard]] bool is_poly() const { return poly_; }
  [[nodiscard]] bool is_scalar() const { return !poly_; }
  void negate_inplace() {   // seal_ring.tcc:58-69: a scalar becomes a polynomial first
    to_poly_inplace();
    if (dev_negate()) return;
    map_words([](uint64_t x, uint64_t p) { return x ? p - x : 0; });
  }
  RingElem operator-() const {
    RingElem res(*this);
    res.negate_inplace();
    return res;
  }
  [[nodiscard]] bool is_invertible() const noexcept {
    try {
      RingElem tmp(*this);
      tmp.invert_inplace();
      return true;
    } catch (...) {
      return false;
    }
  }
  void invert_inplace() {   // seal_ring.tcc:87-103 over poly_arith.cpp:304-340: every slot must be invertible
    to_poly_inplace();
    if (dev_invert()) return;
    const auto &rp = detail::ring_params();
    Words w(words());
    for (size_t j = 0; j < rp.L; j++)
      for (size_t i = 0; i < rp.N; i++) {
        const uint64_t x = w[j * rp.N + i], p = rp.q[j];
        if (x % p == 0) throw std::invalid_argument("element is not invertible in ring");
        uint64_t r = 1, base = x % p, e = p - 2;   // q_j prime: x^(p-2), the unique inverse in [1, p) that try_invert_uint_mod returns
        while (e) {
          if (e & 1) r = detail::mulmod(r, base, p);
          base = detail::mulmod(base, base, p);
          e >>= 1;
        }
        w[j * rp.N + i] = r;
      }
    set_poly(std::move(w));
  }
  [[nodiscard]] RingElem inverse() const {
    RingElem res(*this);
    res.invert_inplace();
    return res;
  }
  RingElem &operator+=(const RingElem &o) {   // seal_ring.tcc:105-152
    if (poly_) {
      if (o.poly_) {
        if (!dev_binop(0, o)) zip_words(o.words(), [](uint64_t a, uint64_t b, uint64_t p) { const uint64_t s = a + b; return s >= p ? s - p : s; });
      } else if (o.scalar_ != 0) {
        const uint64_t v = o.scalar_;
        if (!dev_scalar_op(0, v)) map_words([v](uint64_t a, uint64_t p) { const uint64_t s = a + v; return s >= p ? s - p : s; });
      }
    } else {
      if (scalar_ == 0) return *this = o;
      if (o.poly_) {
        const uint64_t v = scalar_;
        *this = o;
        return *this += RingElem(v);
      }
      const size_t a = detail::bitsize_like_reference(scalar_), b = detail::bitsize_like_reference(o.scalar_);
      const size_t res = a == b ? a + 1 : std::max(a, b);
      if (res < detail::bitsize_like_reference(detail::ring_params().q[0])) scalar_ += o.scalar_;
      else {
        to_poly_inplace();
        return *this += o;
      }
    }
    return *this;
  }
  RingElem &operator-=(const RingElem &o) {   // seal_ring.tcc:154-185
    if (poly_) {
      if (o.poly_) {
        if (!dev_binop(1, o)) zip_words(o.words(), [](uint64_t a, uint64_t b, uint64_t p) { return a >= b ? a - b : a - b + p; });
      } else if (o.scalar_ != 0) {
        const uint64_t v = o.scalar_;
        if (!dev_scalar_op(1, v)) map_words([v](uint64_t a, uint64_t p) { return a >= v ? a - v : a - v + p; });
      }
    } else if (o.poly_) {   // scalar - poly = -poly + scalar
      const uint64_t v = scalar_;
      *this = o;
      negate_inplace();
      if (v) *this += RingElem(v);
    } else {                // scalar - scalar always becomes a polynomial (seal_ring.tcc:175-178)
      to_poly_inplace();
      if (o.scalar_ != 0) *this -= o;
    }
    return *this;
  }
  RingElem &operator*=(const RingElem &o) {   // seal_ring.tcc:187-247
    if (poly_) {
      if (o.poly_) {
        if (!dev_binop(2, o)) zip_words(o.words(), [](uint64_t a, uint64_t b, uint64_t p) { return detail::mulmod(a, b, p); });
      } else {
        if (o.scalar_ == 1) return *this;
        if (o.scalar_ == 0) {
          set_scalar(0);
          return *this;
        }
        const uint64_t v = o.scalar_;
        if (!dev_scalar_op(2, v)) map_words([v](uint64_t a, uint64_t p) { return detail::mulmod(a, v % p, p); });
      }
    } else {
      if (scalar_ == 1) return *this = o;
      if (scalar_ == 0) return *this;
      if (o.fast_is_zero()) {
        set_scalar(0);
        return *this;
      }
      if (o.poly_) {
        const uint64_t v = scalar_;
        *this = o;
        return *this *= RingElem(v);
      }
      const size_t res = detail::bitsize_like_reference(scalar_) + detail::bitsize_like_reference(o.scalar_);
      if (res < detail::bitsize_like_reference(detail::ring_params().q[0])) scalar_ *= o.scalar_;
      else {
        to_poly_inplace();
        return *this *= o;
      }
    }
    return *this;
  }
  RingElem &operator/=(const RingElem &o) { return *this *= o.inverse(); }
  RingElem &to_poly_inplace() {
    if (!poly_) set_poly(scalar_words(scalar_));
    return *this;
  }
  [[nodiscard]] RingElem to_poly() const {
    RingElem res(*this);
    res.to_poly_inplace();
    return res;
  }
  [[nodiscard]] size_t hash() const {
    if (!poly_) return scalar_;
    size_t h = 0;
    for (uint64_t v : words()) h ^= v;
    return h;
  }
  using invalid_ring_elem_types = Host::invalid_ring_elem_types;
  [[nodiscard]] uint64_t get_scalar() const {
    if (poly_) throw std::bad_variant_access();
    return scalar_;
  }
  [[nodiscard]] polytools::SealPoly get_poly() const {
    if (!poly_) throw std::bad_variant_access();
    auto &ctx = get_context();
    return polytools::SealPoly(ctx, words(), &ctx.first_parms_id());
  }
  // SealPoly::is_equal (poly_arith.cpp:155-162): memcmp over data.size() BYTES, i.e. the first W / 8 words
  static bool prefix_equal(const Words &a, const Words &b) { return a.size() == b.size() && !std::memcmp(a.data(), b.data(), a.size()); }
};

inline RingElem operator+(const RingElem &l, const RingElem &r) { RingElem x(l); x += r; return x; }
inline RingElem operator-(const RingElem &l, const RingElem &r) { RingElem x(l); x -= r; return x; }
inline RingElem operator*(const RingElem &l, const RingElem &r) { RingElem x(l); x *= r; return x; }
inline RingElem operator/(const RingElem &l, const RingElem &r) { RingElem x(l); x /= r; return x; }
inline bool operator==(const RingElem &l, const RingElem &r) {   // seal_ring.tcc:249-263
  if (l.is_scalar() && r.is_scalar()) return l.get_scalar() == r.get_scalar();
  return RingElem::prefix_equal(l.to_poly().words(), r.to_poly().words());
}
inline bool operator!=(const RingElem &l, const RingElem &r) { return !(l == r); }
inline std::ostream &operator<<(std::ostream &out, const RingElem &e) {
  if (e.is_scalar()) return out << e.get_scalar();
  return out << e.get_poly().to_json();
}

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
// SealPoly::is_zero (poly_arith.cpp:147-153) of element idx of a device vector: the flags are computed on the device, once
inline bool dev_zero_flag(const std::shared_ptr<DevRing> &d, size_t idx) {
  std::lock_guard<std::mutex> g(host_mutex());
  if (d->zero.empty()) {
    const size_t cnt = rsg_ringvec_size(d->v);
    d->zero.resize(cnt);
    check(rsg_ringvec_is_zero_prefix(d->v, 0, cnt, d->zero.data()));
  }
  return d->zero[idx] != 0;
}
inline std::shared_ptr<DevRing> new_dev_elem(rsg_context *ctx) {
  auto v = std::make_shared<DevRing>();
  check(rsg_ringvec_create(ctx, 1, &v->v));
  v->ctx = ctx;
  return v;
}
}  // namespace detail

inline bool RingElem::is_zero() const {
  if (!poly_) return scalar_ == 0;
  if (device_only()) return detail::dev_zero_flag(dev_, dev_idx_);
  const Words &w = words();   // bytes [0, W + 7) of the W words
  const size_t W = w.size(), bytes = W + 7, full = std::min(bytes / 8, W), rem = bytes % 8;
  for (size_t i = 0; i < full; i++)
    if (w[i]) return false;
  return !(rem && full < W && (w[full] & ((1ull << (8 * rem)) - 1)));
}
// Both operands in HBM (same device context), no host copy yet: the operator runs there and the result stays there.
inline bool RingElem::dev_binop(int op, const RingElem &o) {
  if (!device_only() || !o.device_only() || !dev_->ctx || dev_->ctx != o.dev_->ctx) return false;
  auto out = detail::new_dev_elem(dev_->ctx);
  detail::check(rsg_ring_binop(dev_->ctx, op, dev_->v, dev_idx_, o.dev_->v, o.dev_idx_, out->v, 0, 1));
  dev_ = out;
  dev_idx_ = 0;
  return true;
}
inline bool RingElem::dev_scalar_op(int op, uint64_t v) {
  if (!device_only() || !dev_->ctx) return false;
  auto out = detail::new_dev_elem(dev_->ctx);
  detail::check(rsg_ring_scalar_op(dev_->ctx, op, dev_->v, dev_idx_, v, out->v, 0, 1));
  dev_ = out;
  dev_idx_ = 0;
  return true;
}
inline bool RingElem::dev_negate() {
  if (!device_only() || !dev_->ctx) return false;
  auto out = detail::new_dev_elem(dev_->ctx);
  detail::check(rsg_ring_negate(dev_->ctx, dev_->v, dev_idx_, out->v, 0, 1));
  dev_ = out;
  dev_idx_ = 0;
  return true;
}
inline bool RingElem::dev_invert() {
  if (!device_only() || !dev_->ctx) return false;
  auto out = detail::new_dev_elem(dev_->ctx);
  uint8_t ok = 0;
  detail::check(rsg_ring_invert(dev_->ctx, dev_->v, dev_idx_, out->v, 0, 1, &ok));   // RSG_ERR_NOTINV -> std::invalid_argument
  dev_ = out;
  dev_idx_ = 0;
  return true;
}

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
    else if (N && SealEnc::get_contexts()[0].first_context_data()->parms().poly_modulus_degree() != N)
      throw std::invalid_argument("the encoding contexts of this process were set for another N");
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
    else {
      auto &held = SealEnc::get_contexts();
      if (held.size() != contexts_.size()) throw std::invalid_argument("the encoding contexts of this process differ from the ones supplied");
      for (size_t i = 0; i < held.size(); i++)
        if (held[i].first_parms_id() != contexts_[i].first_parms_id())
          throw std::invalid_argument("the encoding contexts of this process differ from the ones supplied");
    }
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
    if (other.is_empty()) return *this;
    if (!arena_) {   // empty += anything non-empty (also the size-0 zero ciphertexts: the result is then NOT empty); zero += encoding
      if (is_empty() || other.arena_) *this = other;
      return *this;
    }
    if (other.zero_) return *this;   // SEAL's add_inplace with a size-0 ciphertext adds nothing
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
    v->ctx = b.ctx;
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
    v->ctx = b.ctx;
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
    v->ctx = b.ctx;
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
