// Robust cost functions used for measurement re-weighting (GNC / M-estimation).  Scalar host
// math outside the CUDA hot path; API-compatible with the reference's
// include/DPGO/DPGO_robust.h (RobustCostParameters, RobustCost::weight/reset/update).
#ifndef DPGO_B200_ROBUST_H
#define DPGO_B200_ROBUST_H

#include <cmath>
#include <cstddef>
#include <iostream>
#include <stdexcept>
#include <string>

namespace DPGO {

double chi2inv(double quantile, size_t dof);  // DPGO_utils

class RobustCostParameters {
 public:
  enum class Type { L2, L1, TLS, Huber, GM, GNC_TLS };

  Type costType;
  unsigned GNCMaxNumIters;
  double GNCBarc, GNCMuStep, GNCInitMu;
  double HuberThreshold;
  double TLSThreshold;

  explicit RobustCostParameters(Type type = Type::L2, unsigned gncMaxIters = 20, double gncBarc = 5.0,
                                double gncMuStep = 1.4, double gncInitMu = 1e-4, double huberThresh = 3,
                                double TLSThresh = 10)
      : costType(type), GNCMaxNumIters(gncMaxIters), GNCBarc(gncBarc), GNCMuStep(gncMuStep),
        GNCInitMu(gncInitMu), HuberThreshold(huberThresh), TLSThreshold(TLSThresh) {}

  static std::string robustCostName(Type type) {
    static const char *names[] = {"L2", "L1", "TLS", "Huber", "GM", "GNC_TLS"};
    return names[static_cast<int>(type)];
  }

  friend std::ostream &operator<<(std::ostream &os, const RobustCostParameters &p) {
    os << "Robust cost parameters: \nCost function: " << robustCostName(p.costType)
       << "\nGNC maximum iterations: " << p.GNCMaxNumIters << "\nGNC mu step: " << p.GNCMuStep
       << "\nGNC initial mu: " << p.GNCInitMu << "\nGNC threshold (barc): " << p.GNCBarc
       << "\nHuber threshold: " << p.HuberThreshold << "\nTLS threshold: " << p.TLSThreshold << "\n";
    return os;
  }
};

class RobustCost {
 public:
  explicit RobustCost(const RobustCostParameters &params) : mParams(params), mu(params.GNCInitMu) { reset(); }

  /// weight of a measurement with (unsquared) residual r
  double weight(double r) const {
    using T = RobustCostParameters::Type;
    switch (mParams.costType) {
      case T::L2: return 1.0;
      case T::L1: return 1.0 / r;
      case T::Huber: return r < mParams.HuberThreshold ? 1.0 : mParams.HuberThreshold / r;
      case T::TLS: return r < mParams.TLSThreshold ? 1.0 : 0.0;
      case T::GM: {
        const double a = 1.0 + r * r;
        return 1.0 / (a * a);
      }
      case T::GNC_TLS: {  // eq. (14) of Yang et al., "Graduated Non-Convexity for Robust Spatial Perception"
        const double c2 = mParams.GNCBarc * mParams.GNCBarc, r2 = r * r;
        if (r2 >= (mu + 1.0) / mu * c2) return 0.0;
        if (r2 <= mu / (mu + 1.0) * c2) return 1.0;
        return std::sqrt(c2 * mu * (mu + 1.0) / r2) - mu;
      }
    }
    throw std::runtime_error("weight function for selected cost function is not implemented !");
  }

  void reset() {
    if (mParams.costType == RobustCostParameters::Type::GNC_TLS) {
      mu = mParams.GNCInitMu;
      mGNCIteration = 0;
    }
  }

  /// GNC: advance the mu schedule
  void update() {
    if (mParams.costType != RobustCostParameters::Type::GNC_TLS) return;
    if (++mGNCIteration > mParams.GNCMaxNumIters) {
      std::printf("GNC: reached maximum iterations.");
      return;
    }
    mu *= mParams.GNCMuStep;
  }

  static double computeErrorThresholdAtQuantile(double quantile, size_t dimension) {
    if (dimension != 3) throw std::runtime_error("quantile function currently only supports 3D problem.");
    if (!(quantile > 0)) throw std::runtime_error("quantile must be positive");
    return quantile < 1 ? std::sqrt(chi2inv(quantile, 6)) : 1e5;
  }

 private:
  const RobustCostParameters mParams;
  size_t mGNCIteration = 0;
  double mu;
};

}  // namespace DPGO
#endif
