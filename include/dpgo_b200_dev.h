/* dpgo_b200_dev.h -- development and measurement entry points of libdpgo_b200.so.
 *
 * NOT part of the drop-in boundary (include/dpgo_b200.h holds every call the reference-facing C++ shells
 * make): storage / tuning choices of the exact preconditioner, kernel variants kept for A/B measurements,
 * timing helpers used by bench.py and tools/, and host-only inspection of the nested dissection.  Nothing here
 * changes results beyond summation order. */
#ifndef DPGO_B200_DEV_H
#define DPGO_B200_DEV_H
#include "dpgo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* How the exact preconditioner (Q + 0.1 I)^{-1} is stored and applied.  Both forms give the same operator (up to
 * summation order); they differ in bytes streamed per application:
 *  -1 (default) choose by size when the preconditioner is built: 2 when N = (d+1)n >= 3000, else 0;
 *   0 full dense inverse, N^2*8 bytes per application (one streaming pass, 2 grid phases);
 *   2 two-level: nested-dissection domains with dense interior inverses A_II^{-1} and a dense
 *     inverse of the separator Schur complement S = A_SS - A_SI A_II^{-1} A_IS; one application is
 *     z_S = S^{-1}(r_S - A_SI A_II^{-1} r_I), z_I = A_II^{-1}(r_I - A_IS z_S): 3 strip GEMVs and 2
 *     sparse couplings, ~N^2*8/14 bytes on sphere2500 (L2 resident), 5 grid phases.
 * (Round 2 measured and removed three more forms: symmetric half storage, and two three-phase forms of the
 * two-level elimination -- all slower on every data set once the grid barrier was fixed, DESIGN.md.)
 * Takes effect at the next dpgo_finalize(h, 1) / dpgo_update_weights(..., 1). */
int dpgo_set_precon_mode(dpgo_handle h, int mode);
/* The form in use (0 or 2) once the preconditioner is built. */
int dpgo_get_precon_mode(dpgo_handle h, int *mode);
/* Host-only inspection of the partition the two-level variant is built on (no device needed): the
 * nested dissection of a pose graph given as a block-CSR pattern (n block rows, rowptr[n+1],
 * colidx) into interior domains of at most max_domain_poses poses (<= 0: the library's value for
 * (d+1) = dh scalars per pose) and a vertex separator.  group[i] = domain id of pose i, or -1 for
 * a separator pose; *num_domains = number of domains.  No block of the pattern joins two
 * different domains. */
int dpgo_two_level_partition(int n, const int32_t *rowptr, const int32_t *colidx, int dh,
                             int max_domain_poses, int32_t *group, int *num_domains);
/* Tuning of the two-level variant (measurement knobs; 0 / negative = library default): how many
 * partial slots the inner dimension of the interior strips and of the Schur strips is split into
 * (more splits = more CTAs busy per phase, more partial sums to add), and whether the first
 * pipeline stages of a strip phase are issued before the grid barrier that precedes it
 * (prefetch: 1 on, 0 off, negative = default on).  Takes effect at the next preconditioner build. */
int dpgo_set_precon_tuning(dpgo_handle h, int split_interior, int split_schur, int prefetch);
/* Measurement knob of the two-level variants: poses per interior domain of the nested dissection
 * (0 = library default: a domain is one wave of strip stages, 80 poses for d = 3).  Larger domains mean
 * fewer separator poses and larger interior inverses; max_domain_poses >= n gives a single domain and no
 * separator, i.e. the full dense inverse applied by the one-CTA-per-SM strip kernel (strips then take
 * several waves).  Takes effect at the next preconditioner build. */
int dpgo_set_two_level_domain_size(dpgo_handle h, int max_domain_poses);

/* ---- measurement helpers ----------------------------------------------------------------- */
/* Time `reps` back-to-back launches of the Q*X kernel / preconditioner kernel on the handle's
 * stream with CUDA events; flush_l2 != 0 streams a >L2-sized buffer between launches
 * (outside the timed intervals).  Returns mean microseconds per launch. */
int dpgo_time_qx(dpgo_handle h, int reps, int flush_l2, double *usec);
int dpgo_time_precon(dpgo_handle h, int reps, int flush_l2, double *usec);
/* Same for the per-pose kernels: op 0 = QF retraction (slot 0 + slot 1), 1 = polar projection of
 * 0.5 slot0 + 0.3 slot1 + 0.2 slot2 (the Nesterov updateY / updateV form), 2 = rounding of slot 0 in the frame of
 * its first pose.  Algorithmic bytes: 3, 4 and (r + d)/r tile arrays of r(d+1)n doubles. */
int dpgo_time_pose_op(dpgo_handle h, int op, int reps, int flush_l2, double *usec);
/* Measurement knob for the stand-alone Q*X (dpgo_qx, dpgo_time_qx; the solver's fused passes are not
 * affected): variant -1 / 0 = the default kernel; 1 = the same product with a software prefetch -- every pose
 * group asks the L2 (cp.async.bulk.prefetch.L2) for the Q blocks, column indices and X tile of the pose
 * `prefetch_distance` rows further on (0 = the poses covered by the CTAs that are resident together);
 * 2 = the X tiles of a warp step fetched once (16-byte pieces dealt over the lanes) and staged in shared memory;
 * 3 = the block row walked two blocks per step with the column indices one step ahead.  All give the same bits. */
int dpgo_set_qx_variant(dpgo_handle h, int variant, int prefetch_distance);
/* Measurement builds only (library compiled with -DDPGO_TRACE, `python dpgo_b200/build.py --trace`;
 * otherwise DPGO_ESTATE): how long every CTA of the last fused solve worked in each phase before
 * reaching the phase's grid barrier, busy_ms[cta * 16 + phase] with the phase ids of
 * dpgo_ropt_result.phase_ms.  *num_ctas = CTAs of that launch; nothing is written when cap_ctas is
 * smaller. */
int dpgo_phase_trace(dpgo_handle h, double *busy_ms, int cap_ctas, int *num_ctas);
/* algorithmic bytes of one Q*X / one preconditioner application (SURVEY 8(d) formula) */
int dpgo_bytes_qx(dpgo_handle h, double *bytes);
int dpgo_bytes_precon(dpgo_handle h, double *bytes);

#ifdef __cplusplus
}
#endif
#endif /* DPGO_B200_DEV_H */
