// Internal state behind a dpgo_handle (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/dpgo_b200.h"
#include "../../include/dpgo_b200_dev.h"

struct dpgo_mailbox_s;

namespace dpgo {

// one edge's contribution to one block of Q (host-side assembly, device_lib.cu)
struct Contribution {
  int32_t row, col;
  int32_t src;   // index into the edge set
  int8_t kind;   // 0: T W T^T, 1: -T W, 2: -W T^T, 3: W   (private) ; 4/5 shared out/in ; 6 prior ; 7 zero
};


struct EdgeSet {
  int m = 0;
  std::vector<int32_t> a, b;  // private: p1,p2 ; shared: my_idx, nbr_slot
  std::vector<uint8_t> outgoing;
  std::vector<double> R, t, kappa, tau, weight;
};

void set_error(const char *fmt, ...);

}  // namespace dpgo

struct dpgo_dev;
namespace dpgo {
// device_lib.cu
int read_partials(dpgo_dev *h, int nblocks, int K, double *out);
int sync_host_blocks(dpgo_dev *h);   // host copy of the Q blocks <- device, when a device-side re-weighting made it stale
// precon_dd.cu: two-level (domain decomposition) exact preconditioner, precon_mode == 2
int dd_build(dpgo_dev *h);
void dd_free(dpgo_dev *h);
int op_precon_dd(dpgo_dev *h, const double *Y, const double *rvec, double *z, double *neg_out, double *z_r);
int dd_time_apply(dpgo_dev *h, const double *vec);   // the streaming phases only (no finish)
double dd_bytes(const dpgo_dev *h);
int two_level_max_domain_poses(int dh);   // interior domains hold at most this many poses
}  // namespace dpgo

struct dpgo_dev {
  int device = 0, n = 0, d = 0, r = 0;
  int N = 0;       // (d+1) n
  int ld = 0;      // N rounded up to a multiple of 64 (Pinv rows)
  int ldk = 0;     // nsplit * KT >= ld (Pinv columns / vector padding)
  size_t vlen = 0; // r * N   (doubles in a lifted pose array)
  size_t vpad = 0; // r * ldk (allocated doubles per array)
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int num_sms = 148;

  // host-side graph (reference: PoseGraph members, include/DPGO/PoseGraph.h)
  dpgo::EdgeSet priv, shared;
  int num_nbr_slots = 0;
  std::vector<int32_t> prior_idx;
  std::vector<double> prior_poses;
  double prior_kappa = 10000.0, prior_tau = 100.0;

  // block-CSR Q (host copy + device)
  std::vector<int32_t> rowptr, colidx;
  std::vector<double> blocks;
  int nnzb = 0;
  int *d_rowptr = nullptr, *d_colidx = nullptr, *d_browidx = nullptr;
  double *d_blocks = nullptr;
  // cross block-CSR for G = Gconst + Xnbr * C
  int cnnzb = 0;
  int *d_crowptr = nullptr, *d_ccolidx = nullptr;
  double *d_cblocks = nullptr;
  double *d_Gconst = nullptr, *d_G = nullptr, *d_nbr = nullptr;
  double *d_nbr_xy[2] = {nullptr, nullptr};   // dpgo_neighbor_buffer: neighbours' X / auxiliary Y (exchange.cu)
  std::vector<dpgo_mailbox_s *> mailboxes;    // receiver-side mailboxes of the asynchronous publication (exchange.cu)
  void *d_mailbox_views = nullptr;
  bool mailbox_views_dirty = true;

  // dense preconditioner
  double *d_Pinv = nullptr, *d_zpart = nullptr;
  int KT = 0, nsplit = 0;
  // symmetric half-storage variant (precon_mode == 1): T blocks of 128, NG groups of kSymS
  // blocks, work items (ig >= kg), partial buffers zD (in d_zpart) and zT
  // precon_mode is the RESOLVED storage variant (0 full dense inverse, 1 symmetric half, 2 two-level);
  // precon_request is what the caller asked for (-1 = choose by size at build time)
  int precon_mode = 0;
  int precon_request = -1;
  int gemv_occ = 0;
  int partial_blocks = 0;  // CTAs the partials buffer can serve (8 doubles each)
  bool finalized = false, has_precon = false;
  bool weights_only_update = false;   // inside dpgo_update_weights: the pattern of Q is unchanged
  std::vector<dpgo::Contribution> q_contribs;   // sorted contributions of the edges to the blocks of Q (the pattern)
  int uploaded_nnzb = -1;
  // device copy of what a weight-only refresh (GNC) needs: the contribution lists per block of Q, the sorted shared
  // edges of the cross blocks, the edge data and the weights; Q and the cross blocks are then re-weighted by kernels
  // and the host copy of the blocks goes stale until somebody asks for it (dpgo_get_Q_bsr)
  struct DeviceEdges { double *R = nullptr, *t = nullptr, *kappa = nullptr, *tau = nullptr, *w = nullptr; };
  DeviceEdges de_priv, de_shared;
  int *d_cptr = nullptr, *d_csrc = nullptr, *d_corder = nullptr;
  signed char *d_ckind = nullptr;
  unsigned char *d_sout = nullptr;
  bool refresh_plan_valid = false, host_blocks_stale = false;

  // lifted pose arrays
  double *d_slot[4] = {nullptr, nullptr, nullptr, nullptr};
  double *d_xa = nullptr, *d_xb = nullptr;
  double *d_EG = nullptr, *d_EG2 = nullptr, *d_grad = nullptr, *d_grad2 = nullptr;
  double *d_S = nullptr, *d_S2 = nullptr;
  double *d_eta = nullptr, *d_r = nullptr, *d_z = nullptr, *d_delta = nullptr, *d_Hd = nullptr;
  double *d_t0 = nullptr, *d_t1 = nullptr, *d_t2 = nullptr;
  double *d_partials = nullptr, *d_scalars = nullptr;
  double *h_scalars = nullptr;  // pinned
  void *d_fused = nullptr;      // fused-kernel parameter / result block
  void *h_fused = nullptr;      // pinned mirror
  void *d_trace = nullptr;      // -DDPGO_TRACE builds: per-CTA phase times of the last fused solve
  int last_grid = 0;            // CTAs of the last fused launch
  void *dd = nullptr;           // dpgo::DdState (precon_mode >= 2)
  int dd_split1 = 0, dd_split3 = 0;   // inner splits of the interior / Schur strips (0 = by size)
  int dd_prefetch = 1;                // issue the next strip phase's first stages before the barrier
  // stand-alone Q*X (dpgo_set_qx_variant; measurement variants): -1 / 0 = lane-group kernel (default),
  // 1 = + L2 prefetch hints, 2 = X tiles staged in shared memory, 3 = two blocks per step
  int qx_variant = -1, qx_prefetch_dist = 0;
  int dd_max_domain = 0;              // poses per interior domain (0 = one wave of strip stages)
  int *d_public_idx = nullptr;
  int num_public = 0;
  double *d_flush = nullptr;
  size_t flush_bytes = 0;

  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t launches = 0;
  // dpgo_optimize_slot_async / dpgo_optimize_result
  bool pending = false;
  int pending_verbose = 0;
  int64_t pending_l0 = 0, pending_launches = 0;
};
