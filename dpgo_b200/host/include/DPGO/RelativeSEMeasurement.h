// One relative SE(d) measurement from pose (r1,p1) to pose (r2,p2); field-compatible with the
// reference's include/DPGO/RelativeSEMeasurement.h:23-88.
#ifndef DPGO_B200_RELATIVESEMEASUREMENT_H
#define DPGO_B200_RELATIVESEMEASUREMENT_H

#include <DPGO/DPGO_types.h>

namespace DPGO {

struct RelativeSEMeasurement {
  size_t r1 = 0, r2 = 0;  ///< robot ids of tail / head
  size_t p1 = 0, p2 = 0;  ///< frame ids of tail / head
  Matrix R;               ///< d x d rotation
  Matrix t;               ///< d x 1 translation
  double kappa = 0;       ///< rotational precision
  double tau = 0;         ///< translational precision
  bool fixedWeight = false;
  double weight = 1.0;    ///< GNC weight in [0,1]

  RelativeSEMeasurement() = default;
  RelativeSEMeasurement(size_t first_robot, size_t second_robot, size_t first_pose, size_t second_pose,
                        const Matrix &relative_rotation, const Matrix &relative_translation,
                        double rotational_precision, double translational_precision)
      : r1(first_robot), r2(second_robot), p1(first_pose), p2(second_pose), R(relative_rotation),
        t(relative_translation), kappa(rotational_precision), tau(translational_precision) {}

  friend std::ostream &operator<<(std::ostream &os, const RelativeSEMeasurement &m) {
    os << "r1: " << m.r1 << "\np1: " << m.p1 << "\nr2: " << m.r2 << "\np2: " << m.p2 << "\nR:\n" << m.R
       << "\nt:\n" << m.t << "\nKappa: " << m.kappa << "\nTau: " << m.tau << "\nFixed weight: " << m.fixedWeight
       << "\nWeight: " << m.weight << "\n";
    return os;
  }
};

}  // namespace DPGO
#endif
