/*
 * dpgo_b200 -- C-ABI of the B200-native RBCD local-solve hot path.
 *
 * This is the drop-in boundary: everything the reference's QuadraticProblem /
 * QuadraticOptimizer / PoseGraph data matrices / PGOAgent Nesterov step compute on the CPU
 * (Eigen + ROPTLIB + CHOLMOD) is computed behind these entry points by hand-written sm_100a
 * CUDA kernels.  Plain pointers and sizes only; no C++/torch types cross the boundary.
 *
 * Conventions
 *   - every dense argument is FP64, column-major, r x (d+1)n ("lifted pose array",
 *     reference layout pinned by tests/testEigenMap.cpp:12-36 and
 *     include/DPGO/manifold/Poses.h:21-118): pose i occupies columns (d+1)i..(d+1)i+d,
 *     first d columns = Stiefel block, last column = translation.
 *   - "host" pointers are ordinary host memory, "dev" pointers are device memory on the
 *     handle's device; sizes are implied by (n, d, r) of the handle.
 *   - every function returns 0 on success, a negative DPGO_E* code otherwise; nothing throws
 *     across the boundary; dpgo_last_error() returns a thread-local message.
 *   - a handle is bound to one device and one stream; calls on one handle must be serialized by
 *     the caller (PGOAgent's mutexes already do, src/PGOAgent.cpp:940-942).
 *   - there is NO CPU fallback: every compute entry point runs CUDA kernels or fails.
 *
 * Citations "ref:" are file:line under the reference repository (mit-acl/dpgo @ a238090c).
 */
#ifndef DPGO_B200_H
#define DPGO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPGO_OK 0
#define DPGO_EINVAL (-1)   /* bad argument / contract violation (reference: glog CHECK abort) */
#define DPGO_ECUDA (-2)    /* CUDA runtime / cuSOLVER failure                                 */
#define DPGO_ESTATE (-3)   /* call order violated (e.g. solve before finalize)                */
#define DPGO_ENUMERIC (-4) /* Cholesky of Q + 0.1 I failed (ref: src/PoseGraph.cpp:606-607)   */

typedef struct dpgo_dev *dpgo_handle;

/* ROptParameters, ref: include/DPGO/DPGO_types.h:44-86 (+ ROPTLIB SolversTR defaults the
 * reference leaves untouched: theta, kappa, accept_rho, shrink, magnify). */
typedef struct dpgo_ropt_params {
  int32_t method;                 /* 0 = RTR, 1 = RGD                          */
  int32_t verbose;
  double gradnorm_tol;            /* 1e-2                                      */
  double RGD_stepsize;            /* 1e-3                                      */
  int32_t RGD_use_preconditioner; /* 1                                         */
  int32_t RTR_iterations;         /* 3                                         */
  int32_t RTR_tCG_iterations;     /* 50                                        */
  int32_t fused;                  /* 1 = single persistent kernel (default), 0 = one launch per op */
  double RTR_initial_radius;      /* 100                                       */
  double tcg_theta;               /* 1                                         */
  double tcg_kappa;               /* 0.1                                       */
  double accept_rho;              /* 0.1                                       */
  double shrink;                  /* 0.25                                      */
  double magnify;                 /* 2                                         */
} dpgo_ropt_params;

/* ROPTResult, ref: include/DPGO/DPGO_types.h:91-107, plus work counters. */
typedef struct dpgo_ropt_result {
  int32_t success;
  int32_t tcg_status;  /* 0 LCON, 1 SCON, 2 NEGCURVTURE, 3 EXCREGION, 4 MAXITER (ROPTLIB tCGstatusSet) */
  double f_init, gradnorm_init, f_opt, gradnorm_opt;
  double elapsed_ms;   /* device time of the solve (CUDA events on the handle's stream) */
  int32_t outer_iters, inner_iters, accepted, rejected;
  /* work actually executed on the device (for the roofline accounting) */
  int64_t n_qx;        /* block-CSR SpMM passes (Q applied to an r x N array)  */
  int64_t n_precon;    /* dense (Q+0.1I)^-1 applications                       */
  int64_t n_pose_sweeps; /* per-pose sweeps (projection / retraction / ...)    */
  int64_t n_launches;  /* kernels launched by this call                        */
  /* fused solver only: device time (ms, CTA 0's globaltimer, barrier waits included) spent in
   * 0 cost+gradient, 1 preconditioner GEMV, 2 preconditioner finish (sum+projection),
   * 3 Hessian-vector, 4 tCG vector update, 5 tCG direction update, 6 retraction / copies, 7 unused;
   * two-level preconditioner (mode 2) only, the parts of slot 1: 8 interior strips (y), 9 separator
   * right-hand side, 10 Schur strips, 11 back-substitution right-hand side, 12 interior strips (w) */
  double phase_ms[16];
  int64_t n_barriers;  /* grid-wide barriers executed by the fused kernel          */
} dpgo_ropt_result;

void dpgo_default_params(dpgo_ropt_params *p);
const char *dpgo_last_error(void);
const char *dpgo_version(void);

/* ---- lifecycle ------------------------------------------------------------------------- */
/* One handle = one agent's PoseGraph + QuadraticProblem + optimizer state on one GPU.
 * ref: PoseGraph(id, r, d) src/PoseGraph.cpp:14-21; QuadraticProblem ctor
 * src/QuadraticProblem.cpp:17-23.  `stream` is a cudaStream_t (or NULL: the library creates
 * its own non-blocking stream). */
int dpgo_create(int device, int n, int d, int r, void *stream, dpgo_handle *out);
int dpgo_destroy(dpgo_handle h);
int dpgo_dims(dpgo_handle h, int *n, int *d, int *r);
/* number of CUDA kernels this handle has launched so far (bench.py's gpu_launches) */
int dpgo_launch_count(dpgo_handle h, int64_t *count);
int dpgo_sync(dpgo_handle h);

/* ---- data matrices (ref: PoseGraph::constructQ/G, constructConnectionLaplacianSE) -------- */
/* Private (intra-robot) edges, struct-of-arrays; R is m x d x d row-major, t is m x d.
 * ref: src/DPGO_utils.cpp:272-344, src/PoseGraph.cpp:381-391. */
int dpgo_set_private_edges(dpgo_handle h, int m, const int32_t *p1, const int32_t *p2,
                           const double *R, const double *t, const double *kappa,
                           const double *tau, const double *weight);
/* Shared (inter-robot) edges: my_idx = my pose, nbr_slot = index of the neighbour's pose in
 * the neighbour-pose buffer (order chosen by the caller, e.g. PoseGraph::neighborPublicPoseIDs
 * order), outgoing[k] = 1 if my pose is the tail (ref: m.r1 == id_, src/PoseGraph.cpp:409).
 * ref: src/PoseGraph.cpp:403-458 (Q diagonal terms), :505-563 (G terms). */
int dpgo_set_shared_edges(dpgo_handle h, int m, int num_nbr_slots, const int32_t *my_idx,
                          const int32_t *nbr_slot, const uint8_t *outgoing, const double *R,
                          const double *t, const double *kappa, const double *tau,
                          const double *weight);
/* Priors on my poses (ref: PoseGraph::setPrior src/PoseGraph.cpp:176-181, :462-469, :566-575);
 * poses is num x r x (d+1) column-major tiles. */
int dpgo_set_priors(dpgo_handle h, int num, const int32_t *idx, const double *poses,
                    double prior_kappa, double prior_tau);
/* Build the block-CSR Q on the host, upload it, build the cross blocks for G, and build the
 * exact preconditioner (Q + 0.1 I)^-1 with the library's own Cholesky / inverse kernels (the reference
 * uses a CHOLMOD factorization, src/PoseGraph.cpp:598-613 -- both are exact solves; how the inverse is
 * stored is the library's choice by size, see dpgo_b200_dev.h).
 * build_precon = 0 skips the preconditioner (then only unpreconditioned ops are available). */
int dpgo_finalize(dpgo_handle h, int build_precon);
/* Update only the measurement weights (GNC): Q, the cross blocks of G and -- when asked for -- the preconditioner
 * are re-weighted on the device from the new weights (same pattern: no host assembly, the symbolic part of the
 * preconditioner set-up is kept; the result has the bits of a from-scratch dpgo_finalize with these weights).
 * NULL keeps the current weights of that edge set.
 * ref: PoseGraph::clearDataMatrices after weight updates, src/PGOAgent.cpp:1062-1142. */
int dpgo_update_weights(dpgo_handle h, const double *w_private, const double *w_shared,
                        int build_precon);

/* Q as the library built it (for parity tests): nnzb, and optionally the arrays. */
int dpgo_get_Q_bsr(dpgo_handle h, int *nnzb, int32_t *rowptr, int32_t *colidx, double *blocks);

/* Linear term.  Either set G directly (host, r x N) or provide neighbour poses
 * (num_nbr_slots tiles of r x (d+1)) and let the device build it.
 * ref: PoseGraph::setNeighborPoses + constructG, src/PoseGraph.cpp:183-186, :493-580. */
int dpgo_set_G(dpgo_handle h, const double *G_host);
int dpgo_set_neighbor_poses(dpgo_handle h, const double *tiles_host);
int dpgo_set_neighbor_poses_dev(dpgo_handle h, const double *tiles_dev);
int dpgo_get_G(dpgo_handle h, double *G_host);

/* ---- QuadraticProblem operators (host in / host out; used by parity tests and by the C++
 *      QuadraticProblem shell).  ref: src/QuadraticProblem.cpp:29-83 ------------------------ */
int dpgo_qx(dpgo_handle h, const double *X, double *out);                  /* X*Q            */
int dpgo_f(dpgo_handle h, const double *X, double *f);                     /* :29-41         */
int dpgo_egrad(dpgo_handle h, const double *X, double *out);               /* :43-47         */
int dpgo_rgrad(dpgo_handle h, const double *X, double *out, double *norm); /* :71-83         */
int dpgo_hessvec(dpgo_handle h, const double *X, const double *V, double *out); /* Riemannian Hess[V] at X (:49-54 + Stiefel::EucHvToHv) */
int dpgo_precon(dpgo_handle h, const double *X, const double *V, double *out);  /* :56-69    */
int dpgo_tangent_project(dpgo_handle h, const double *X, const double *V, double *out);
int dpgo_retract(dpgo_handle h, const double *X, const double *V, double *out); /* QF retraction */
int dpgo_project_manifold(dpgo_handle h, const double *M, double *out);  /* LiftedSEManifold::project, src/manifold/LiftedSEManifold.cpp:34-45 */

/* ---- QuadraticOptimizer (ref: src/QuadraticOptimizer.cpp:26-137) ------------------------- */
/* optimize(): X0/Xout are host r x N arrays; either may be NULL to use / keep the
 * device-resident iterate of the handle (slot DPGO_SLOT_X). */
int dpgo_optimize(dpgo_handle h, const dpgo_ropt_params *params, const double *X0,
                  double *Xout, dpgo_ropt_result *result);

/* ---- device-resident agent state (ref: PGOAgent X / Y / V / XPrev, src/PGOAgent.cpp) ----- */
#define DPGO_SLOT_X 0
#define DPGO_SLOT_Y 1
#define DPGO_SLOT_V 2
#define DPGO_SLOT_XPREV 3
int dpgo_slot_set(dpgo_handle h, int slot, const double *host);
int dpgo_slot_get(dpgo_handle h, int slot, double *host);
int dpgo_slot_copy(dpgo_handle h, int dst, int src);
/* Y = project((1-alpha) X + alpha V)  ref: PGOAgent::updateY src/PGOAgent.cpp:922-928 */
int dpgo_nesterov_update_Y(dpgo_handle h, double alpha);
/* V = project(V + gamma (X - Y))      ref: PGOAgent::updateV src/PGOAgent.cpp:930-936 */
int dpgo_nesterov_update_V(dpgo_handle h, double gamma);
/* X <- optimize(starting from slot `from`), result left in slot X.  ref: updateX :938-995 */
int dpgo_optimize_slot(dpgo_handle h, const dpgo_ropt_params *params, int from,
                       dpgo_ropt_result *result);
/* Stream-ordered form of the same call for drivers that queue whole RBCD rounds ahead of the device
 * (the reference's updateX blocks its caller; a one-process-per-GPU driver does not have to):
 * dpgo_optimize_slot_async queues the single-launch RTR solve (params->method == 0 and
 * params->fused != 0, else DPGO_EINVAL) and the copy of its result block on the handle's stream and
 * returns without waiting; dpgo_optimize_result waits for the stream and returns the result of the
 * most recent asynchronous solve (DPGO_ESTATE if there is none). */
int dpgo_optimize_slot_async(dpgo_handle h, const dpgo_ropt_params *params, int from);
int dpgo_optimize_result(dpgo_handle h, dpgo_ropt_result *result);
/* Pack the public poses (indices given once) of slot `slot` into a device buffer of
 * num_public tiles -- the payload of getSharedPoseDict / getAuxSharedPoseDict
 * (ref: src/PGOAgent.cpp:97-110, :132-146). */
int dpgo_set_public_indices(dpgo_handle h, int num_public, const int32_t *idx);
int dpgo_pack_public_dev(dpgo_handle h, int slot, double *tiles_dev);
/* Same with a caller-owned DEVICE index list: out[k] = tile idx_dev[k] of slot `slot` -- used to
 * pack, per neighbour, exactly the poses that neighbour needs (getSharedPoseDictWithNeighbor,
 * ref: src/PGOAgent.cpp:112-130) straight into the NCCL send buffer. */
int dpgo_gather_tiles_dev(dpgo_handle h, int slot, int num, const int32_t *idx_dev,
                          double *tiles_dev);
/* ---- public-pose exchange between agents, inside the library ------------------------------------
 * What PGOAgent::getSharedPoseDict / getAuxSharedPoseDict hand out (ref: src/PGOAgent.cpp:97-146) and
 * updateNeighborPoses / updateAuxNeighborPoses take in (:650-702), as the reference's driver moves them
 * every iteration (examples/MultiRobotExample.cpp:183-204): packed device tiles, NCCL send/recv between
 * ranks (one process per GPU), gathered straight into the receiver's buffer when both agents share a
 * device.  A communicator is bound to one device + stream; all handles it serves use that stream, so a
 * round (solve -> pack -> send/recv -> G -> solve) is ordered by the stream alone. */
typedef struct dpgo_comm_s *dpgo_comm;
#define DPGO_COMM_ID_BYTES 128
int dpgo_comm_unique_id(unsigned char *id);                       /* rank 0 creates it, every rank gets a copy */
int dpgo_comm_create(int device, int rank, int world, const unsigned char *id, void *stream, dpgo_comm *out);
int dpgo_comm_destroy(dpgo_comm c);
int dpgo_comm_launch_count(dpgo_comm c, int64_t *n);              /* kernels + NCCL groups queued so far */
/* Neighbour pose buffers owned by the handle (num_nbr_slots tiles): aux = 0 the neighbours' X
 * (neighborPoseDict), aux = 1 their auxiliary Y (neighborAuxPoseDict). */
int dpgo_neighbor_buffer(dpgo_handle h, int aux, double **dev_ptr);
/* G from one of them: setNeighborPoses + constructG (ref: src/PoseGraph.cpp:183-186, :493-580). */
int dpgo_use_neighbor_poses(dpgo_handle h, int aux);
/* One message: the `count` tiles of slot `slot` of agent `src` listed in the device array d_frames go to
 * neighbour slots [dst_offset, dst_offset + count) of agent `dst` (buffer `aux`).  src == NULL: the sender
 * lives on rank `peer` (receive); dst == NULL: the receiver lives on rank `peer` (send).  Both sides list
 * the messages between a pair of ranks in the same order. */
typedef struct dpgo_message {
  dpgo_handle src;
  dpgo_handle dst;
  int peer;
  int slot;
  int aux;
  int count;
  const int32_t *d_frames;
  int dst_offset;
} dpgo_message;
int dpgo_exchange(dpgo_comm c, const dpgo_message *msgs, int n);

/* ---- asynchronous publication of public poses through peer memory ---------------------------------
 * The reference's asynchronous mode (PGOAgent::startOptimizationLoop / runOptimizationLoop,
 * src/PGOAgent.cpp:475-499) lets every agent iterate at its own rate with whatever neighbour poses have
 * arrived (updateNeighborPoses from another thread, :650-678).  Here the sender stores its public poses
 * straight into a mailbox in the RECEIVER's GPU memory (CUDA IPC mapping, NVLink peer stores) and the
 * receiver takes consistent snapshots before a solve: no rendezvous, no collective, no host in between.
 *   receiver:  dpgo_mailbox_create(h_a, first_slot_of_b, count, &mb, ipc)   one per neighbour b; send `ipc` to b's process
 *   sender:    dpgo_mailbox_open(device, ipc, NULL, count, tile, &to)       (or local = mb when a and b share a process)
 *   sender, after every solve:   dpgo_publish(h_b, DPGO_SLOT_X, to, frames_a_needs_dev)
 *   receiver, before every solve: dpgo_collect(h_a); dpgo_use_neighbor_poses(h_a, 0)                              */
typedef struct dpgo_mailbox_s *dpgo_mailbox;
#define DPGO_IPC_HANDLE_BYTES 64
int dpgo_mailbox_create(dpgo_handle h, int dst_offset, int count, dpgo_mailbox *out, unsigned char *ipc_handle);
int dpgo_mailbox_open(int device, const unsigned char *ipc_handle, dpgo_mailbox local, int count, int tile,
                      dpgo_mailbox *out);
int dpgo_mailbox_close(dpgo_mailbox m);
int dpgo_publish(dpgo_handle h, int slot, dpgo_mailbox to, const int32_t *d_frames);
int dpgo_collect(dpgo_handle h);

/* Squared residual of every measurement of this agent at the poses in `slot`:
 *   err = kappa |Y1 R~ - Y2|_F^2 + tau |p2 - p1 - Y1 t~|^2
 * (ref: computeMeasurementError, src/DPGO_utils.cpp:501-507; PGOAgent::computeMeasurementResidual,
 * src/PGOAgent.cpp:1062-1102 takes the square root), one device thread per edge -- the input of
 * the GNC / M-estimator weight update (PGOAgent::updateMeasurementWeights :1104-1142; follow with
 * dpgo_update_weights).  err_private / err_shared are host arrays in the order of
 * dpgo_set_private_edges / dpgo_set_shared_edges.  Shared edges read the neighbours' poses from
 * `nbr_poses_dev` (num_nbr_slots tiles, device memory) or, when it is NULL, from the buffer last
 * given to dpgo_set_neighbor_poses(_dev). */
int dpgo_measurement_errors(dpgo_handle h, int slot, const double *nbr_poses_dev, double *err_private,
                            double *err_shared);
/* Rounding of the lifted iterate in `slot` to SE(d) poses in the frame of an anchor pose:
 *   R_i = projectToRotationGroup(Ya^T Y_i), t_i = Ya^T (p_i - pa), one device thread per pose
 * (ref: PGOAgent::getTrajectoryInLocalFrame / getTrajectoryInGlobalFrame src/PGOAgent.cpp:718-767,
 * projectToRotationGroup src/DPGO_utils.cpp:464-478).  anchor_tile = the r x (d+1) lifted anchor pose (host;
 * the globalAnchor), or NULL for the local frame (pose 0 of the slot).  T_host: d x (d+1)n, column-major
 * (PoseArray layout). */
int dpgo_round_trajectory(dpgo_handle h, int slot, const double *anchor_tile, double *T_host);
/* max_i || p_i(a) - p_i(b) ||  (LiftedPoseArray::maxTranslationDistance, used for
 * PGOAgentStatus.relativeChange, src/PGOAgent.cpp:404) */
int dpgo_max_translation_distance(dpgo_handle h, int slot_a, int slot_b, double *out);

/* ---- chordal initialization (ref: chordalInitialization src/DPGO_solver.cpp:220-269) ------------ */
/* Relax-and-round initial guess from the measurements alone, on the device: rotations from the linear least-squares
 * problem min sum kappa ||R_j - R_i R_ij||^2 with R_0 = I (every block then projected to SO(d),
 * projectToRotationGroup), translations from min sum tau ||t_j - t_i - R_i t_ij||^2 with t_0 = 0
 * (recoverTranslations).  The reference factorizes with SPQR; here both normal equations are solved by conjugate
 * gradients on the device with the library's Q*X kernel and its exact (Q + 0.1 I)^-1 as the preconditioner.
 * Measurement weights are not used (as in the reference).  Edges as in dpgo_set_private_edges.
 * T_host: d x (d+1)n, column-major (PoseArray layout).  info (optional): iterations and final relative residuals of
 * the two solves (a residual above ~1e-8 means the graph is badly conditioned or not connected to pose 0). */
typedef struct {
  int rotation_iterations, translation_iterations;
  double rotation_residual, translation_residual;
  int64_t launches;
} dpgo_chordal_info;
int dpgo_chordal_initialization(int device, int n, int d, int m, const int32_t *p1, const int32_t *p2,
                                const double *R, const double *t, const double *kappa, const double *tau,
                                double *T_host, dpgo_chordal_info *info);

#ifdef __cplusplus
}
#endif
#endif /* DPGO_B200_H */
