// Stand-in for boost::math::tools::polynomial<T> (Boost.Math >= 1.80 is required by the
// reference, /root/reference/CMakeLists.txt:37, but is neither vendored there nor installed
// in this image). TEST INFRASTRUCTURE ONLY: used to compile the reference under oracle/_ref.
//
// Only the members the reference calls are provided (ringsnark/util/polynomials.tcc:61-81,
// ringsnark/util/evaluation_domain.tcc:53-60): construction from std::vector<T> (normalises),
// data(), size(), +=, *=, /=.  Semantics restated from Boost's published behaviour:
//   * normalize(): drop trailing coefficients c for which (c != T(0)) is false;
//   * += : zero-extend to the longer operand, add element-wise, normalise;
//   * *= : schoolbook product into a T(0)-filled vector of size |a|+|b|-1 (zero operand ->
//          zero polynomial), NOT re-normalised;
//   * /= : quotient of classical long division (Knuth 4.6.1 algorithm D) with
//          q of size |u|-|v|+1, NOT re-normalised; zero if |u| < |v|.
// All arithmetic is exact modular arithmetic in the reference's ring, so the only thing
// that could differ from real Boost is the *length* of a returned vector (trailing zeros);
// the caller (r1cs_to_qrp.tcc:250-253) tolerates shorter vectors.  Parity with real Boost
// is therefore unpinned only at that level.
#pragma once
#include <algorithm>
#include <cstddef>
#include <utility>
#include <vector>

namespace boost {
namespace math {
namespace tools {

template <class T>
class polynomial {
 public:
  typedef typename std::vector<T>::size_type size_type;

  polynomial() {}
  polynomial(const std::vector<T> &p) : m_data(p) { normalize(); }
  polynomial(std::vector<T> &&p) : m_data(std::move(p)) { normalize(); }

  std::vector<T> &data() { return m_data; }
  const std::vector<T> &data() const { return m_data; }
  size_type size() const { return m_data.size(); }
  bool is_zero() const { return m_data.empty(); }

  void normalize() {
    auto rit = std::find_if(m_data.rbegin(), m_data.rend(),
                            [](const T &x) { return x != T(0); });
    m_data.erase(rit.base(), m_data.end());
  }

  polynomial &operator+=(const polynomial &v) {
    if (m_data.size() < v.size()) m_data.resize(v.size(), T(0));
    for (size_type i = 0; i < v.size(); ++i) m_data[i] += v.m_data[i];
    normalize();
    return *this;
  }

  polynomial &operator*=(const polynomial &v) {
    if (v.is_zero() || is_zero()) {
      m_data.clear();
      return *this;
    }
    std::vector<T> prod(size() + v.size() - 1, T(0));
    for (size_type i = 0; i < v.size(); ++i)
      for (size_type j = 0; j < size(); ++j) prod[i + j] += m_data[j] * v.m_data[i];
    m_data.swap(prod);
    return *this;
  }

  polynomial &operator/=(const polynomial &v) {
    if (size() < v.size()) {
      m_data.clear();
      return *this;
    }
    std::vector<T> u(m_data);
    const size_type m = u.size() - 1, n = v.size() - 1;
    size_type k = m - n;
    std::vector<T> q(m - n + 1, T(0));
    do {
      q[k] = u[n + k] / v.m_data[n];
      for (size_type j = n + k; j > k;) {
        j--;
        u[j] -= q[k] * v.m_data[j - k];
      }
    } while (k-- != 0);
    m_data.swap(q);
    return *this;
  }

 private:
  std::vector<T> m_data;
};

}  // namespace tools
}  // namespace math
}  // namespace boost
