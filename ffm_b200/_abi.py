"""ctypes mirror of include/ffm_b200.h -- the only way Python reaches the kernels.

There is no CPU fallback: if the library is missing or a call fails the error is raised.
"""
import ctypes as C
import os

from . import build as _build

ABI_VERSION = 8
FFM_HOST, FFM_DEVICE = 0, 1
FFM_NEUMANN, FFM_MOORE = 4, 8
FFM_F32, FFM_F64 = 0, 1
E_INVALID, E_CUDA, E_UNSUPPORTED, E_STATE = -1, -2, -3, -4
MODEL_CORE, MODEL_UNIFIED_CRITIC, MODEL_UNIFIED_ACTOR, MODEL_UNIFIED_BOTH, MODEL_TRAINED, MODEL_MCQ = range(6)
LEARN_NONE, LEARN_EXACT, LEARN_BATCHED = range(3)
LEGACY_AC, LEGACY_ACTOR_ONLY = 0, 1
LEGACY_TABLE_V, LEGACY_TABLE_H = 0, 1
SFF_L1, SFF_L2, SFF_LINF, SFF_BFS4, SFF_BFS8, SFF_DIJKSTRA8 = range(6)


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32),
        ("height", C.c_int32), ("width", C.c_int32),
        ("neighborhood", C.c_int32), ("sff_dtype", C.c_int32),
        ("n_episodes", C.c_int32), ("n_max", C.c_int32),
        ("track_dff", C.c_int32), ("reserved0", C.c_int32),
        ("k_S", C.c_double), ("k_D", C.c_double),
        ("dff_c0", C.c_float), ("dff_c1", C.c_float), ("dff_threshold", C.c_float), ("reserved1", C.c_float),
        ("seed", C.c_uint64), ("episode_base", C.c_uint32), ("reserved2", C.c_uint32),
        ("model", C.c_int32), ("learn", C.c_int32), ("block_size", C.c_int32), ("reserved3", C.c_int32),
        ("k_A", C.c_double), ("gamma", C.c_double), ("alpha_v", C.c_double), ("alpha_h", C.c_double),
        ("exit_reward", C.c_double), ("step_penalty", C.c_double), ("collision_penalty", C.c_double),
        ("epsilon", C.c_double), ("sff_min", C.c_double), ("sff_max", C.c_double),
        ("stop_penalty", C.c_double), ("timeout_penalty", C.c_double), ("step_cap", C.c_int32), ("q_log2_capacity", C.c_int32),
    ]


class LegacyConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32),
        ("height", C.c_int32), ("width", C.c_int32),
        ("neighborhood", C.c_int32), ("sff_dtype", C.c_int32),
        ("n_episodes", C.c_int32), ("n_max", C.c_int32),
        ("model", C.c_int32), ("learn", C.c_int32),
        ("block_size", C.c_int32), ("table_log2_capacity", C.c_int32),
        ("k_S", C.c_double), ("k_D", C.c_double), ("k_A", C.c_double),
        ("dff_c0", C.c_float), ("dff_c1", C.c_float), ("dff_threshold", C.c_float), ("reserved0", C.c_float),
        ("gamma", C.c_double), ("alpha_v", C.c_double), ("alpha_h", C.c_double), ("exit_reward", C.c_double),
        ("step_penalty", C.c_double), ("collision_penalty", C.c_double), ("epsilon", C.c_double),
        ("sff_min", C.c_double), ("sff_max", C.c_double),
        ("seed", C.c_uint64), ("episode_base", C.c_uint32), ("reserved1", C.c_uint32),
    ]


class Draws(C.Structure):
    _fields_ = [("move", C.c_void_p), ("conflict", C.c_void_p), ("steps", C.c_int32),
                ("first_step", C.c_int32), ("space", C.c_int32), ("reserved", C.c_int32)]


class RolloutOut(C.Structure):
    _fields_ = [("traj_cells", C.c_void_p), ("traj_n", C.c_void_p), ("traj_steps", C.c_int32),
                ("reserved", C.c_int32), ("rec_state", C.c_void_p), ("rec_action", C.c_void_p),
                ("rec_reward", C.c_void_p), ("rec_len", C.c_void_p),
                ("ctraj", C.c_void_p), ("ctraj_off", C.c_void_p), ("ctraj_cap", C.c_int64)]


# name -> (restype, argtypes); tests/test_abi.py checks this table against the header
SIGNATURES = {
    "ffm_abi_version": (C.c_int, []),
    "ffm_last_error": (C.c_char_p, []),
    "ffm_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "ffm_destroy": (C.c_int, [C.c_void_p]),
    "ffm_set_fields": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ffm_set_positions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ffm_place": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "ffm_get_positions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ffm_set_dff": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ffm_get_dff": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ffm_update_dff": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ffm_rollout": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(Draws), C.POINTER(RolloutOut), C.c_void_p]),
    "ffm_move_probs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ffm_get_counters": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ffm_tables_shape": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ffm_tables_set": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ffm_tables_get": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ffm_tables_bind_deltas": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ffm_tables_apply_deltas": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ffm_set_epsilon": (C.c_int, [C.c_void_p, C.c_double]),
    "ffm_bind_dynamic": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ffm_set_episode_base": (C.c_int, [C.c_void_p, C.c_uint32]),
    "ffm_q_shape": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "ffm_q_get": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ffm_q_set": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "ffm_mcq_set_forced": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ffm_mcq_backup_ordered": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ffm_mcq_accumulate": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ffm_mcq_export_deltas": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ffm_mcq_import_deltas": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "ffm_mcq_fold": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ffm_set_beta": (C.c_int, [C.c_void_p, C.c_double]),
    "ffm_mcq_finalize_timeouts": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ffm_sff_generate": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int,
                                   C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]),
    "ffm_rollout_returns": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_void_p,
                                      C.c_int32, C.c_void_p]),
    "ffm_measure_smem_bandwidth": (C.c_int, [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ffm_launch_count": (C.c_int64, [C.c_void_p]),
    "ffm_kernel_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ffm_cluster_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "ffm_legacy_create": (C.c_int, [C.POINTER(LegacyConfig), C.POINTER(C.c_void_p)]),
    "ffm_legacy_destroy": (C.c_int, [C.c_void_p]),
    "ffm_legacy_set_fields": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ffm_legacy_set_positions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ffm_legacy_get_positions": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ffm_legacy_set_dff": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ffm_legacy_get_dff": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ffm_legacy_zero_dff": (C.c_int, [C.c_void_p]),
    "ffm_legacy_update_dff": (C.c_int, [C.c_void_p]),
    "ffm_legacy_rollout": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
    "ffm_legacy_get_counters": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ffm_legacy_table_size": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]),
    "ffm_legacy_table_get": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "ffm_legacy_table_set": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_double]),
    "ffm_legacy_set_epsilon": (C.c_int, [C.c_void_p, C.c_double]),
    "ffm_legacy_set_episode_base": (C.c_int, [C.c_void_p, C.c_uint32]),
    "ffm_legacy_launch_count": (C.c_int64, [C.c_void_p]),
}

_lib = None


class FfmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libffm_b200 error {code}: {msg}")
        self.code = code


def lib():
    """Load libffm_b200.so (must have been built: ``python -m ffm_b200.build``)."""
    global _lib
    if _lib is None:
        path = _build.LIB
        if not os.path.exists(path):
            raise ImportError(f"{path} not built; run `python -m ffm_b200.build` (no CPU fallback exists)")
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.ffm_abi_version() != ABI_VERSION:
            raise ImportError("libffm_b200.so ABI version mismatch; rebuild")
        _lib = L
    return _lib


def check(rc):
    """Map an FFM_E_* return code to the exception the reference would raise."""
    if rc == 0:
        return
    msg = lib().ffm_last_error().decode("utf-8", "replace")
    if rc == E_INVALID:
        raise ValueError(msg)
    raise FfmError(rc, msg)
