// TEST INFRASTRUCTURE (oracle) -- never linked into the product library.
//
// Stand-in for the sliver of <ceres/ceres.h> (Ceres Solver 1.9.0, third-party, not
// vendored in the reference) that the reference's hot-path headers touch:
// CostFunction, LossFunction and AutoDiffCostFunction
// (src/rsba/video_bundler_free.h:70-91, src/rsba/video_bundler_rs_inter.h:36-47,88-98).
// AutoDiffCostFunction here is functional: Evaluate() seeds one Jet per parameter
// scalar and returns per-block row-major Jacobians, which is the documented
// contract of ceres::AutoDiffCostFunction.
#ifndef RSBA_ORACLE_SHIM_CERES_H_
#define RSBA_ORACLE_SHIM_CERES_H_

#include <cmath>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>
#include "ceres/jet.h"

namespace ceres {

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int>& parameter_block_sizes() const { return sizes_; }
  int num_residuals() const { return num_residuals_; }
 protected:
  std::vector<int> sizes_;
  int num_residuals_ = 0;
};

class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};

namespace shim_detail {
template <int... Ns> struct Sum;
template <> struct Sum<> { static const int value = 0; };
template <int N0, int... Ns> struct Sum<N0, Ns...> { static const int value = N0 + Sum<Ns...>::value; };

template <typename F, typename T>
inline bool Call(const F& f, const T* const* p, T* r, std::integral_constant<int, 1>) { return f(p[0], r); }
template <typename F, typename T>
inline bool Call(const F& f, const T* const* p, T* r, std::integral_constant<int, 2>) { return f(p[0], p[1], r); }
template <typename F, typename T>
inline bool Call(const F& f, const T* const* p, T* r, std::integral_constant<int, 3>) { return f(p[0], p[1], p[2], r); }
template <typename F, typename T>
inline bool Call(const F& f, const T* const* p, T* r, std::integral_constant<int, 4>) { return f(p[0], p[1], p[2], p[3], r); }
template <typename F, typename T>
inline bool Call(const F& f, const T* const* p, T* r, std::integral_constant<int, 5>) {
  return f(p[0], p[1], p[2], p[3], p[4], r);
}
}  // namespace shim_detail

template <typename Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public CostFunction {
 public:
  static const int kNumBlocks = sizeof...(Ns);
  static const int kNumParams = shim_detail::Sum<Ns...>::value;
  typedef Jet<double, kNumParams> JetT;

  explicit AutoDiffCostFunction(Functor* functor) : functor_(functor) {
    const int sizes[] = {Ns...};
    sizes_.assign(sizes, sizes + kNumBlocks);
    num_residuals_ = kNumResiduals;
  }

  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
    const int sizes[] = {Ns...};
    typedef std::integral_constant<int, kNumBlocks> Arity;
    if (jacobians == nullptr) {
      const double* p[kNumBlocks];
      for (int b = 0; b < kNumBlocks; ++b) p[b] = parameters[b];
      return shim_detail::Call(*functor_, p, residuals, Arity());
    }
    JetT x[kNumParams];
    JetT y[kNumResiduals];
    const JetT* p[kNumBlocks];
    int off = 0;
    for (int b = 0; b < kNumBlocks; ++b) {
      p[b] = x + off;
      for (int i = 0; i < sizes[b]; ++i) x[off + i] = JetT(parameters[b][i], off + i);
      off += sizes[b];
    }
    if (!shim_detail::Call(*functor_, p, y, Arity())) return false;
    for (int r = 0; r < kNumResiduals; ++r) residuals[r] = y[r].a;
    off = 0;
    for (int b = 0; b < kNumBlocks; ++b) {
      if (jacobians[b] != nullptr) {
        for (int r = 0; r < kNumResiduals; ++r)
          for (int i = 0; i < sizes[b]; ++i) jacobians[b][r * sizes[b] + i] = y[r].v[off + i];
      }
      off += sizes[b];
    }
    return true;
  }

 private:
  std::unique_ptr<Functor> functor_;
};

}  // namespace ceres
#endif  // RSBA_ORACLE_SHIM_CERES_H_
