// Scalar control logic of the Riemannian trust-region / truncated-CG solve, shared verbatim by
// the host-driven path (one launch per op) and the persistent fused kernel (everything on
// device).  Restates ROPTLIB's SolversTR::Run / tCG_TR as configured by the reference at
// src/QuadraticOptimizer.cpp:61-100 (Stop_Criterion GRAD_F, Tolerance = gradnorm_tol,
// initial_Delta, maximum_Delta = 5 initial_Delta, Max_Inner_Iter, Min_Inner_Iter = 0) with
// ROPTLIB's defaults for everything the reference does not set (theta = 1, kappa = 0.1,
// Acceptence_Rho = 0.1, Shrinked_tau = 0.25, Magnified_tau = 2).  ROPTLIB itself is not
// vendored in the reference (cmake/roptlib.cmake:7-8), so this follows the published
// Steihaug-Toint tCG / RTR algorithm (Absil, Baker, Gallivan 2007).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define DPGO_HD __host__ __device__ __forceinline__
#else
#define DPGO_HD inline
#endif

namespace dpgo {

enum TcgStatus { TCG_LCON = 0, TCG_SCON = 1, TCG_NEGCURV = 2, TCG_EXCREGION = 3, TCG_MAXITER = 4 };

struct TcgState {
  double e_Pe, d_Pd, z_r, e_Pd, norm_r0, alpha;
  int status;
};

// after z0 = P(r0):  r_r = <r0,r0>, z_r = <z0,r0>
DPGO_HD void tcg_begin(TcgState &s, double r_r, double z_r) {
  s.e_Pe = 0.0;
  s.norm_r0 = sqrt(r_r);
  s.z_r = z_r;
  s.d_Pd = z_r;
  s.e_Pd = 0.0;
  s.alpha = 0.0;
  s.status = TCG_MAXITER;
}

// after Hd and d_Hd = <delta, H delta>.  Returns true when the step leaves the trust region or
// meets non-positive curvature; then *step = tau (eta += tau delta) and the loop ends.
// Otherwise *step = alpha (eta += alpha delta, r += alpha Hd).
DPGO_HD bool tcg_curvature(TcgState &s, double d_Hd, double Delta, double *step) {
  const double alpha = s.z_r / d_Hd;
  const double e_Pe_new = s.e_Pe + 2.0 * alpha * s.e_Pd + alpha * alpha * s.d_Pd;
  if (d_Hd <= 0.0 || e_Pe_new >= Delta * Delta) {
    *step = (-s.e_Pd + sqrt(s.e_Pd * s.e_Pd + s.d_Pd * (Delta * Delta - s.e_Pe))) / s.d_Pd;
    s.status = (d_Hd <= 0.0) ? TCG_NEGCURV : TCG_EXCREGION;
    return true;
  }
  s.e_Pe = e_Pe_new;
  s.alpha = alpha;
  *step = alpha;
  return false;
}

// after r += alpha Hd, r_r = <r,r>.  Returns true when the residual test fires.
DPGO_HD bool tcg_converged(TcgState &s, double r_r, double theta, double kappa) {
  const double norm_r = sqrt(r_r);
  const double p = pow(s.norm_r0, theta);
  if (norm_r <= s.norm_r0 * (p < kappa ? p : kappa)) {
    s.status = (kappa < p) ? TCG_LCON : TCG_SCON;
    return true;
  }
  return false;
}

// after z = P(r), z_r_new = <z,r>.  Returns beta (delta = -z + beta delta).
DPGO_HD double tcg_direction(TcgState &s, double z_r_new) {
  const double beta = z_r_new / s.z_r;
  s.z_r = z_r_new;
  s.e_Pd = beta * (s.e_Pd + s.alpha * s.d_Pd);
  s.d_Pd = z_r_new + beta * beta * s.d_Pd;
  return beta;
}

// Trust-region ratio test.  eta_g = <eta, grad>, eta_Heta = <eta, Hess eta>.
// Updates Delta; returns true when the candidate is accepted.
DPGO_HD bool rtr_accept(double f1, double f2, double eta_g, double eta_Heta, int tcg_status,
                        double accept_rho, double shrink, double magnify, double max_Delta,
                        double *Delta, double *rho_out) {
  const double rho = (f1 - f2) / (-(eta_g + 0.5 * eta_Heta));
  if (rho > 0.75) {
    if (tcg_status == TCG_EXCREGION || tcg_status == TCG_NEGCURV) {
      const double g = magnify * (*Delta);
      *Delta = g < max_Delta ? g : max_Delta;
    }
  } else if (rho < 0.25) {
    *Delta = shrink * (*Delta);
  }
  *rho_out = rho;
  const double sqeps = 1.4901161193847656e-08;  // sqrt(DBL_EPSILON)
  return (rho > accept_rho) || (fabs(f1 - f2) / (fabs(f1) + 1.0) < sqeps && f2 < f1);
}

}  // namespace dpgo
