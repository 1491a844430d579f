// PGOLogger of the drop-in API (see include/DPGO/PGOLogger.h).  Written against the file format of
// the reference (src/PGOLogger.cpp:18-224), not its code: rotation <-> quaternion conversions are
// implemented here because the drop-in carries no Eigen.
#include <DPGO/PGOLogger.h>

#include <cmath>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <utility>

#include "check.h"

namespace DPGO {

namespace {

struct Quat {
  double x, y, z, w;
};

// unit quaternion of a rotation matrix (branch on the largest of trace / diagonal entries, the
// numerically stable choice; Eigen makes the same choice, so signs agree with the reference)
Quat quatFromRotation(const Matrix &R) {
  Quat q;
  const double tr = R(0, 0) + R(1, 1) + R(2, 2);
  if (tr > 0) {
    double s = std::sqrt(tr + 1.0);
    q.w = 0.5 * s;
    s = 0.5 / s;
    q.x = (R(2, 1) - R(1, 2)) * s;
    q.y = (R(0, 2) - R(2, 0)) * s;
    q.z = (R(1, 0) - R(0, 1)) * s;
    return q;
  }
  int i = 0;
  if (R(1, 1) > R(0, 0)) i = 1;
  if (R(2, 2) > R(i, i)) i = 2;
  const int j = (i + 1) % 3, k = (j + 1) % 3;
  double s = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
  double v[3];
  v[i] = 0.5 * s;
  s = 0.5 / s;
  q.w = (R(k, j) - R(j, k)) * s;
  v[j] = (R(j, i) + R(i, j)) * s;
  v[k] = (R(k, i) + R(i, k)) * s;
  q.x = v[0]; q.y = v[1]; q.z = v[2];
  return q;
}

Matrix rotationFromQuat(Quat q) {  // normalises first, like the reference's loaders
  const double nrm = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  if (nrm > 0) { q.x /= nrm; q.y /= nrm; q.z /= nrm; q.w /= nrm; }
  Matrix R(3, 3);
  const double xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
  const double xy = q.x * q.y, xz = q.x * q.z, yz = q.y * q.z;
  const double wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
  R(0, 0) = 1 - 2 * (yy + zz); R(0, 1) = 2 * (xy - wz);     R(0, 2) = 2 * (xz + wy);
  R(1, 0) = 2 * (xy + wz);     R(1, 1) = 1 - 2 * (xx + zz); R(1, 2) = 2 * (yz - wx);
  R(2, 0) = 2 * (xz - wy);     R(2, 1) = 2 * (yz + wx);     R(2, 2) = 1 - 2 * (xx + yy);
  return R;
}

// next comma-separated field of a line
bool field(std::istringstream &ss, std::string &tok) { return static_cast<bool>(std::getline(ss, tok, ',')); }

}  // namespace

PGOLogger::PGOLogger(std::string logDir) : logDirectory(std::move(logDir)) {}
PGOLogger::~PGOLogger() = default;

void PGOLogger::logMeasurements(std::vector<RelativeSEMeasurement> &measurements, const std::string &filename) {
  if (measurements.empty()) return;
  std::ofstream file(logDirectory + filename);
  if (!file.is_open()) return;
  if (measurements[0].R.rows() == 2) return;
  file << "robot_src,pose_src,robot_dst,pose_dst,qx,qy,qz,qw,tx,ty,tz,kappa,tau,is_known_inlier,weight\n";
  for (const RelativeSEMeasurement &m : measurements) {
    const Quat q = quatFromRotation(m.R);
    file << m.r1 << "," << m.p1 << "," << m.r2 << "," << m.p2 << ",";
    file << q.x << "," << q.y << "," << q.z << "," << q.w << ",";
    file << m.t(0, 0) << "," << m.t(1, 0) << "," << m.t(2, 0) << ",";
    file << m.kappa << "," << m.tau << "," << m.fixedWeight << "," << m.weight << "\n";
  }
}

void PGOLogger::logTrajectory(unsigned d, unsigned n, const Matrix &T, const std::string &filename) {
  if (d == 2) return;
  DPGO_CHECK(static_cast<unsigned>(T.rows()) == d);
  DPGO_CHECK(static_cast<unsigned>(T.cols()) == (d + 1) * n);
  std::ofstream file(logDirectory + filename);
  if (!file.is_open()) return;
  file << "pose_index,qx,qy,qz,qw,tx,ty,tz\n";
  for (unsigned i = 0; i < n; ++i) {
    const Matrix R = T.block(0, i * (d + 1), d, d);
    const Quat q = quatFromRotation(R);
    file << i << "," << q.x << "," << q.y << "," << q.z << "," << q.w << ",";
    file << T(0, i * (d + 1) + d) << "," << T(1, i * (d + 1) + d) << "," << T(2, i * (d + 1) + d) << "\n";
  }
}

Matrix PGOLogger::loadTrajectory(const std::string &filename) {
  std::ifstream infile(logDirectory + filename);
  std::cout << "Loading trajectory from " << logDirectory + filename << "..." << std::endl;
  if (!infile.is_open()) {
    std::cout << "Could not open specified file!" << std::endl;
    return Matrix(0, 0);
  }
  std::map<unsigned, Matrix> poses;
  std::string line, tok;
  std::getline(infile, line);  // header
  unsigned num_poses = 0;
  while (std::getline(infile, line)) {
    if (line.empty()) continue;
    std::istringstream ss(line);
    double v[8];
    for (double &x : v) {
      DPGO_CHECK(field(ss, tok));
      x = std::stod(tok);
    }
    Matrix Ti(3, 4);
    Ti.block(0, 0, 3, 3) = rotationFromQuat(Quat{v[1], v[2], v[3], v[4]});
    Ti(0, 3) = v[5]; Ti(1, 3) = v[6]; Ti(2, 3) = v[7];
    poses[static_cast<unsigned>(v[0])] = Ti;
    num_poses++;
  }
  Matrix T(3, 4 * static_cast<std::ptrdiff_t>(num_poses));
  for (unsigned i = 0; i < num_poses; ++i) {
    const auto it = poses.find(i);
    DPGO_CHECK(it != poses.end());  // pose ids must be 0..n-1 (the reference uses map::at)
    T.block(0, 4 * i, 3, 4) = it->second;
  }
  std::cout << "Loaded " << num_poses << " poses." << std::endl;
  return T;
}

std::vector<RelativeSEMeasurement> PGOLogger::loadMeasurements(const std::string &filename, bool load_weight) {
  std::vector<RelativeSEMeasurement> measurements;
  std::cout << "Loading measurements from " << filename << "..." << std::endl;
  std::ifstream infile(filename);
  if (!infile.is_open()) {
    std::cout << "Could not open specified file!" << std::endl;
    return measurements;
  }
  std::string line, tok;
  std::getline(infile, line);  // header
  while (std::getline(infile, line)) {
    if (line.empty()) continue;
    std::istringstream ss(line);
    double v[15];
    for (double &x : v) {
      DPGO_CHECK(field(ss, tok));
      x = std::stod(tok);
    }
    Matrix t(3, 1);
    t(0, 0) = v[8]; t(1, 0) = v[9]; t(2, 0) = v[10];
    RelativeSEMeasurement m(static_cast<size_t>(v[0]), static_cast<size_t>(v[2]), static_cast<size_t>(v[1]),
                            static_cast<size_t>(v[3]), rotationFromQuat(Quat{v[4], v[5], v[6], v[7]}), t, v[11],
                            v[12]);
    m.fixedWeight = (static_cast<int>(v[13]) != 0);
    if (load_weight) m.weight = v[14];
    measurements.push_back(m);
  }
  std::printf("Loaded %zu measurements.\n", measurements.size());
  return measurements;
}

}  // namespace DPGO
