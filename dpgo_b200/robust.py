"""Robust cost functions for measurement re-weighting (GNC / M-estimators): scalar host math that
mirrors the reference's RobustCost (include/DPGO/DPGO_robust.h, src/DPGO_robust.cpp:45-134); the
C++ drop-in has the same class in dpgo_b200/host/include/DPGO/DPGO_robust.h.  The per-edge
residuals it consumes come from the device (dpgo_measurement_errors)."""
import math

import numpy as np

L2, L1, TLS, HUBER, GM, GNC_TLS = "L2", "L1", "TLS", "Huber", "GM", "GNC_TLS"


class RobustCost:
    def __init__(self, cost_type=GNC_TLS, gnc_max_iters=20, gnc_barc=5.0, gnc_mu_step=1.4, gnc_init_mu=1e-4,
                 huber_threshold=3.0, tls_threshold=10.0):
        if cost_type not in (L2, L1, TLS, HUBER, GM, GNC_TLS):
            raise ValueError("unknown robust cost %r" % (cost_type,))
        self.type = cost_type
        self.gnc_max_iters, self.barc = int(gnc_max_iters), float(gnc_barc)
        self.mu_step, self.init_mu = float(gnc_mu_step), float(gnc_init_mu)
        self.huber, self.tls = float(huber_threshold), float(tls_threshold)
        self.reset()

    def reset(self):                                      # src/DPGO_robust.cpp:98-112
        self.mu = self.init_mu
        self.gnc_iteration = 0

    def weights(self, residuals):
        """Vectorised RobustCost::weight (src/DPGO_robust.cpp:54-96) for an array of (unsquared)
        residuals."""
        r = np.asarray(residuals, dtype=np.float64)
        if self.type == L2:
            return np.ones_like(r)
        if self.type == L1:
            return 1.0 / r
        if self.type == HUBER:
            return np.where(r < self.huber, 1.0, self.huber / np.maximum(r, 1e-300))
        if self.type == TLS:
            return np.where(r < self.tls, 1.0, 0.0)
        if self.type == GM:
            a = 1.0 + r * r
            return 1.0 / (a * a)
        r2, c2, mu = r * r, self.barc * self.barc, self.mu     # GNC_TLS, eq. (14) of the GNC paper
        mid = np.sqrt(c2 * mu * (mu + 1.0) / np.maximum(r2, 1e-300)) - mu
        return np.where(r2 >= (mu + 1.0) / mu * c2, 0.0, np.where(r2 <= mu / (mu + 1.0) * c2, 1.0, mid))

    def weight(self, r):
        return float(self.weights(np.array([r]))[0])

    def update(self):                                     # src/DPGO_robust.cpp:114-132
        if self.type != GNC_TLS:
            return
        self.gnc_iteration += 1
        if self.gnc_iteration > self.gnc_max_iters:
            return
        self.mu *= self.mu_step


def chi2_threshold_3d(quantile):
    """RobustCost::computeErrorThresholdAtQuantile for 3-D problems (6 degrees of freedom): the
    chi-square quantile by bisection on the regularised lower incomplete gamma function
    P(3, x/2) = 1 - exp(-y)(1 + y + y^2/2), y = x/2."""
    if not quantile > 0:
        raise ValueError("quantile must be positive")
    if quantile >= 1:
        return 1e5
    lo, hi = 0.0, 200.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        y = 0.5 * mid
        cdf = 1.0 - math.exp(-y) * (1.0 + y + 0.5 * y * y)
        lo, hi = (mid, hi) if cdf < quantile else (lo, mid)
    return math.sqrt(0.5 * (lo + hi))
