// TEST INFRASTRUCTURE (oracle) -- never linked into the product library.
//
// Minimal forward-mode dual number, written from scratch, standing in for
// ceres::Jet<double, N> of Ceres Solver 1.9.0 (a third-party dependency that is
// pinned by /root/reference/.travis.yml:33 but is NOT vendored in the reference
// tree).  Only the operations that the reference's hot-path headers need
// (src/rsba/mat/cam.h, src/rsba/mat/core.h, src/rsba/video_bundler_free.h,
// src/rsba/video_bundler_rs_inter.h) are provided.  Semantics follow the
// published definition of a Jet: value `a` plus an N-vector of partials `v`,
// every arithmetic rule the exact chain rule.
#ifndef RSBA_ORACLE_SHIM_JET_H_
#define RSBA_ORACLE_SHIM_JET_H_

#include <cmath>
#include <ostream>

namespace ceres {

template <typename T, int N>
struct Jet {
  T a;
  T v[N];

  Jet() : a() { for (int i = 0; i < N; ++i) v[i] = T(); }
  Jet(const T& value) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(); }  // NOLINT
  Jet(int value) : a(T(value)) { for (int i = 0; i < N; ++i) v[i] = T(); }    // NOLINT
  Jet(const T& value, int k) : a(value) {
    for (int i = 0; i < N; ++i) v[i] = T();
    v[k] = T(1);
  }

  Jet& operator+=(const Jet& y) { a += y.a; for (int i = 0; i < N; ++i) v[i] += y.v[i]; return *this; }
  Jet& operator-=(const Jet& y) { a -= y.a; for (int i = 0; i < N; ++i) v[i] -= y.v[i]; return *this; }
  Jet& operator*=(const Jet& y) { *this = *this * y; return *this; }
  Jet& operator/=(const Jet& y) { *this = *this / y; return *this; }
};

// ---- unary
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f) { return f; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f) {
  Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h;
}

// ---- Jet (op) Jet
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h;
}
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h;
}
template <typename T, int N> inline Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h;
}
template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  // (f/g)' = (f' - (f/g) g') / g
  Jet<T, N> h;
  const T g_inv = T(1.0) / g.a;
  h.a = f.a * g_inv;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - h.a * g.v[i]) * g_inv;
  return h;
}

// ---- Jet (op) scalar and scalar (op) Jet
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a += s; return h; }
template <typename T, int N> inline Jet<T, N> operator+(T s, const Jet<T, N>& f) { Jet<T, N> h = f; h.a += s; return h; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a -= s; return h; }
template <typename T, int N> inline Jet<T, N> operator-(T s, const Jet<T, N>& f) {
  Jet<T, N> h; h.a = s - f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h;
}
template <typename T, int N> inline Jet<T, N> operator*(const Jet<T, N>& f, T s) {
  Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h;
}
template <typename T, int N> inline Jet<T, N> operator*(T s, const Jet<T, N>& f) { return f * s; }
template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, T s) {
  const T s_inv = T(1.0) / s;
  Jet<T, N> h; h.a = f.a * s_inv; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s_inv; return h;
}
template <typename T, int N> inline Jet<T, N> operator/(T s, const Jet<T, N>& g) {
  // (s/g)' = -s g' / g^2
  Jet<T, N> h;
  h.a = s / g.a;
  const T minus_s_g_a_inverse2 = -s / (g.a * g.a);
  for (int i = 0; i < N; ++i) h.v[i] = g.v[i] * minus_s_g_a_inverse2;
  return h;
}

// ---- comparisons act on the value part only
#define RSBA_SHIM_JET_CMP(op)                                                                       \
  template <typename T, int N> inline bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) {   \
    return f.a op g.a;                                                                               \
  }                                                                                                  \
  template <typename T, int N> inline bool operator op(const Jet<T, N>& f, const T& s) {            \
    return f.a op s;                                                                                 \
  }                                                                                                  \
  template <typename T, int N> inline bool operator op(const T& s, const Jet<T, N>& g) {            \
    return s op g.a;                                                                                 \
  }
RSBA_SHIM_JET_CMP(<)
RSBA_SHIM_JET_CMP(<=)
RSBA_SHIM_JET_CMP(>)
RSBA_SHIM_JET_CMP(>=)
RSBA_SHIM_JET_CMP(==)
RSBA_SHIM_JET_CMP(!=)
#undef RSBA_SHIM_JET_CMP

// ---- elementary functions
template <typename T, int N> inline Jet<T, N> abs(const Jet<T, N>& f) { return f.a < T(0.0) ? -f : f; }
template <typename T, int N> inline Jet<T, N> sqrt(const Jet<T, N>& f) {
  Jet<T, N> h;
  h.a = std::sqrt(f.a);
  const T two_a_inverse = T(1.0) / (T(2.0) * h.a);
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * two_a_inverse;
  return h;
}
template <typename T, int N> inline Jet<T, N> cos(const Jet<T, N>& f) {
  Jet<T, N> h;
  h.a = std::cos(f.a);
  const T minus_sin = -std::sin(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = minus_sin * f.v[i];
  return h;
}
template <typename T, int N> inline Jet<T, N> sin(const Jet<T, N>& f) {
  Jet<T, N> h;
  h.a = std::sin(f.a);
  const T c = std::cos(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i];
  return h;
}

template <typename T, int N> inline std::ostream& operator<<(std::ostream& s, const Jet<T, N>& z) {
  s << "[" << z.a << " ; ...]";
  return s;
}

}  // namespace ceres

#endif  // RSBA_ORACLE_SHIM_JET_H_
