"""ctypes binding of the C-ABI in include/dpgo_b200.h.  There is no CPU fallback: if the native
library is missing this module raises at import time, and every compute call fails loudly
when no CUDA device is present."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DPGO_B200_LIB: measurement builds of the same library (tools/phase_trace.py loads the -DDPGO_TRACE one)
LIB_PATH = os.environ.get("DPGO_B200_LIB") or os.path.join(_HERE, "libdpgo_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(dpgo_b200 has no CPU fallback)")

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


class RoptParams(C.Structure):
    _fields_ = [("method", C.c_int32), ("verbose", C.c_int32), ("gradnorm_tol", C.c_double),
                ("RGD_stepsize", C.c_double), ("RGD_use_preconditioner", C.c_int32),
                ("RTR_iterations", C.c_int32), ("RTR_tCG_iterations", C.c_int32),
                ("fused", C.c_int32), ("RTR_initial_radius", C.c_double),
                ("tcg_theta", C.c_double), ("tcg_kappa", C.c_double), ("accept_rho", C.c_double),
                ("shrink", C.c_double), ("magnify", C.c_double)]


class RoptResult(C.Structure):
    _fields_ = [("success", C.c_int32), ("tcg_status", C.c_int32), ("f_init", C.c_double),
                ("gradnorm_init", C.c_double), ("f_opt", C.c_double), ("gradnorm_opt", C.c_double),
                ("elapsed_ms", C.c_double), ("outer_iters", C.c_int32), ("inner_iters", C.c_int32),
                ("accepted", C.c_int32), ("rejected", C.c_int32), ("n_qx", C.c_int64),
                ("n_precon", C.c_int64), ("n_pose_sweeps", C.c_int64), ("n_launches", C.c_int64),
                ("phase_ms", C.c_double * 16), ("n_barriers", C.c_int64)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["phase_ms"] = list(self.phase_ms)
        return d


class ChordalInfo(C.Structure):
    _fields_ = [("rotation_iterations", C.c_int32), ("translation_iterations", C.c_int32),
                ("rotation_residual", C.c_double), ("translation_residual", C.c_double), ("launches", C.c_int64)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)
H = C.c_void_p

class Message(C.Structure):
    """dpgo_message: one public-pose message of an exchange (include/dpgo_b200.h)."""
    _fields_ = [("src", H), ("dst", H), ("peer", C.c_int), ("slot", C.c_int), ("aux", C.c_int),
                ("count", C.c_int), ("d_frames", C.c_void_p), ("dst_offset", C.c_int)]


COMM = C.c_void_p
COMM_ID_BYTES = 128
MAILBOX = C.c_void_p
IPC_HANDLE_BYTES = 64

# name -> (restype, argtypes); every symbol declared in include/dpgo_b200.h and include/dpgo_b200_dev.h
SIGNATURES = {
    "dpgo_default_params": (None, [C.POINTER(RoptParams)]),
    "dpgo_last_error": (C.c_char_p, []),
    "dpgo_version": (C.c_char_p, []),
    "dpgo_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(H)]),
    "dpgo_destroy": (C.c_int, [H]),
    "dpgo_dims": (C.c_int, [H, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dpgo_sync": (C.c_int, [H]),
    "dpgo_launch_count": (C.c_int, [H, C.POINTER(C.c_int64)]),
    "dpgo_set_private_edges": (C.c_int, [H, C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, _dp]),
    "dpgo_set_shared_edges": (C.c_int, [H, C.c_int, C.c_int, _ip, _ip, _bp, _dp, _dp, _dp, _dp, _dp]),
    "dpgo_set_priors": (C.c_int, [H, C.c_int, _ip, _dp, C.c_double, C.c_double]),
    "dpgo_finalize": (C.c_int, [H, C.c_int]),
    "dpgo_set_precon_mode": (C.c_int, [H, C.c_int]),
    "dpgo_two_level_partition": (C.c_int, [C.c_int, _ip, _ip, C.c_int, C.c_int, _ip, C.POINTER(C.c_int)]),
    "dpgo_get_precon_mode": (C.c_int, [H, C.POINTER(C.c_int)]),
    "dpgo_set_precon_tuning": (C.c_int, [H, C.c_int, C.c_int, C.c_int]),
    "dpgo_set_two_level_domain_size": (C.c_int, [H, C.c_int]),
    "dpgo_update_weights": (C.c_int, [H, _dp, _dp, C.c_int]),
    "dpgo_get_Q_bsr": (C.c_int, [H, C.POINTER(C.c_int), _ip, _ip, _dp]),
    "dpgo_set_G": (C.c_int, [H, _dp]),
    "dpgo_set_neighbor_poses": (C.c_int, [H, _dp]),
    "dpgo_set_neighbor_poses_dev": (C.c_int, [H, C.c_void_p]),
    "dpgo_get_G": (C.c_int, [H, _dp]),
    "dpgo_qx": (C.c_int, [H, _dp, _dp]),
    "dpgo_f": (C.c_int, [H, _dp, _dp]),
    "dpgo_egrad": (C.c_int, [H, _dp, _dp]),
    "dpgo_rgrad": (C.c_int, [H, _dp, _dp, _dp]),
    "dpgo_hessvec": (C.c_int, [H, _dp, _dp, _dp]),
    "dpgo_precon": (C.c_int, [H, _dp, _dp, _dp]),
    "dpgo_tangent_project": (C.c_int, [H, _dp, _dp, _dp]),
    "dpgo_retract": (C.c_int, [H, _dp, _dp, _dp]),
    "dpgo_project_manifold": (C.c_int, [H, _dp, _dp]),
    "dpgo_optimize": (C.c_int, [H, C.POINTER(RoptParams), _dp, _dp, C.POINTER(RoptResult)]),
    "dpgo_slot_set": (C.c_int, [H, C.c_int, _dp]),
    "dpgo_slot_get": (C.c_int, [H, C.c_int, _dp]),
    "dpgo_slot_copy": (C.c_int, [H, C.c_int, C.c_int]),
    "dpgo_nesterov_update_Y": (C.c_int, [H, C.c_double]),
    "dpgo_nesterov_update_V": (C.c_int, [H, C.c_double]),
    "dpgo_optimize_slot": (C.c_int, [H, C.POINTER(RoptParams), C.c_int, C.POINTER(RoptResult)]),
    "dpgo_optimize_slot_async": (C.c_int, [H, C.POINTER(RoptParams), C.c_int]),
    "dpgo_optimize_result": (C.c_int, [H, C.POINTER(RoptResult)]),
    "dpgo_set_public_indices": (C.c_int, [H, C.c_int, _ip]),
    "dpgo_pack_public_dev": (C.c_int, [H, C.c_int, C.c_void_p]),
    "dpgo_gather_tiles_dev": (C.c_int, [H, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dpgo_comm_unique_id": (C.c_int, [C.c_char_p]),
    "dpgo_comm_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p, C.POINTER(COMM)]),
    "dpgo_comm_destroy": (C.c_int, [COMM]),
    "dpgo_comm_launch_count": (C.c_int, [COMM, C.POINTER(C.c_int64)]),
    "dpgo_neighbor_buffer": (C.c_int, [H, C.c_int, C.POINTER(C.c_void_p)]),
    "dpgo_use_neighbor_poses": (C.c_int, [H, C.c_int]),
    "dpgo_exchange": (C.c_int, [COMM, C.POINTER(Message), C.c_int]),
    "dpgo_mailbox_create": (C.c_int, [H, C.c_int, C.c_int, C.POINTER(MAILBOX), C.c_char_p]),
    "dpgo_mailbox_open": (C.c_int, [C.c_int, C.c_char_p, MAILBOX, C.c_int, C.c_int, C.POINTER(MAILBOX)]),
    "dpgo_mailbox_close": (C.c_int, [MAILBOX]),
    "dpgo_publish": (C.c_int, [H, C.c_int, MAILBOX, C.c_void_p]),
    "dpgo_collect": (C.c_int, [H]),
    "dpgo_measurement_errors": (C.c_int, [H, C.c_int, C.c_void_p, _dp, _dp]),
    "dpgo_round_trajectory": (C.c_int, [H, C.c_int, _dp, _dp]),
    "dpgo_max_translation_distance": (C.c_int, [H, C.c_int, C.c_int, _dp]),
    "dpgo_chordal_initialization": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp, _dp, _dp,
                                              C.POINTER(ChordalInfo)]),
    "dpgo_time_qx": (C.c_int, [H, C.c_int, C.c_int, _dp]),
    "dpgo_time_precon": (C.c_int, [H, C.c_int, C.c_int, _dp]),
    "dpgo_time_pose_op": (C.c_int, [H, C.c_int, C.c_int, C.c_int, _dp]),
    "dpgo_set_qx_variant": (C.c_int, [H, C.c_int, C.c_int]),
    "dpgo_phase_trace": (C.c_int, [H, _dp, C.c_int, C.POINTER(C.c_int)]),
    "dpgo_bytes_qx": (C.c_int, [H, _dp]),
    "dpgo_bytes_precon": (C.c_int, [H, _dp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header / library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


class DpgoError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise DpgoError(f"dpgo_b200 error {rc}: {lib.dpgo_last_error().decode()}")
