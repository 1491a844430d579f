// placeholder until the persistent kernel lands: route to the host-driven solver
#include "device_state.h"
namespace dpgo {
int solve_host(dpgo_dev *h, const dpgo_ropt_params *P, const double *x_in, double *x_out,
               dpgo_ropt_result *res);
int solve_fused(dpgo_dev *h, const dpgo_ropt_params *P, const double *x_in, double *x_out,
                dpgo_ropt_result *res) {
  return solve_host(h, P, x_in, x_out, res);
}
}  // namespace dpgo
