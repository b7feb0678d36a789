"""LegacySim: the legacy 13-cell models (model/ffm_ac_core.py, model/ffm_actor_only.py) on the device.

Thin ctypes wrapper over the ``ffm_legacy_*`` entry points of include/ffm_b200.h (kernels:
csrc/ffm_legacy.cu).  The dict tables of the reference are keyed here by
``((bx * nby + by) << 26) | sum_j cell_j << 2j``; ``key_to_state`` / ``state_to_key`` convert to the
``(tuple(state_13), (bx, by))`` tuples the reference pickles (ffm_ac_core.py:106-109).
"""
import ctypes as C

import numpy as np
import torch

from . import _abi
from .sim import MOORE, NEUMANN

AC_DEFAULTS = {                           # ffm_ac_core.py:10-23
    "k_S": 10, "k_D": 1, "diffuse": 0.2, "decay": 0.2, "neighborhood": "neumann",
    "alpha_v": 0.1, "gamma": 0.95, "exit_reward": 100.0, "step_penalty": 0.0, "collision_penalty": -1.0,
    "block_size": 3,
}
ACTOR_ONLY_DEFAULTS = {                   # ffm_actor_only.py:24-39
    "k_D": 1, "k_A": 10, "diffuse": 0.2, "decay": 0.2, "neighborhood": "neumann",
    "alpha_v": 0.1, "gamma": 0.95, "exit_reward": 100.0, "step_penalty": 0.0, "collision_penalty": -1.0,
    "alpha_h": 0.1, "epsilon": 0.0,
}
ACTOR_ONLY_BLOCK = 5                      # ffm_actor_only.py:144


def state_to_key(state, nby):
    cells, (bx, by) = state
    code = 0
    for j, v in enumerate(cells):
        code |= int(v) << (2 * j)
    return ((int(bx) * nby + int(by)) << 26) | code


def key_to_state(key, nby):
    key = int(key)
    code, blk = key & ((1 << 26) - 1), key >> 26
    return (tuple((code >> (2 * j)) & 3 for j in range(13)), (blk // nby, blk % nby))


class LegacySim:
    """B episodes of a legacy model sharing map, SFF, parameters and tables.

    model   "ac" (ffm_ac_core.FloorFieldModel) | "actor_only" (ffm_actor_only.FloorFieldModelActorOnly)
    learn   "exact": the reference's sequential per-agent table updates (B must be 1); "none": frozen tables
    """

    def __init__(self, map_array, sff, n_episodes, n_max, model="ac", learn="exact", params=None, seed=0, episode_base=0,
                 table_log2_capacity=0, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("ffm_b200 needs a CUDA device (no CPU fallback)")
        if model not in ("ac", "actor_only"):
            raise ValueError(model)
        base = AC_DEFAULTS if model == "ac" else ACTOR_ONLY_DEFAULTS
        self.params = dict(base) if params is None else {**base, **params}
        self.model = model
        self.map_array = np.ascontiguousarray(np.asarray(map_array).astype(np.uint8))
        sff = np.asarray(sff)
        if sff.shape != self.map_array.shape:
            raise ValueError("sff shape differs from map shape")
        if model == "ac":
            self.sff = np.ascontiguousarray(sff if sff.dtype == np.float32 else sff.astype(np.float64))   # file dtype (:28)
        else:
            self.sff = np.ascontiguousarray(np.where(np.isinf(sff), 0.0, sff).astype(np.float32))         # ffm_actor_only.py:45-48
        self.H, self.W = self.map_array.shape
        self.B, self.n_max = int(n_episodes), int(n_max)
        self.neighbors = list(NEUMANN) if self.params["neighborhood"] == "neumann" else list(MOORE)
        self.A = len(self.neighbors) + 1
        self.block_size = int(self.params["block_size"]) if model == "ac" else ACTOR_ONLY_BLOCK
        self.nby = (self.W + self.block_size - 1) // self.block_size
        self.device = torch.cuda.current_device() if device is None else int(device)
        p = self.params
        decay, diffuse = p["decay"], p["diffuse"]
        cfg = _abi.LegacyConfig()
        cfg.abi_version, cfg.device = _abi.ABI_VERSION, self.device
        cfg.height, cfg.width = self.H, self.W
        cfg.neighborhood = len(self.neighbors)
        cfg.sff_dtype = _abi.FFM_F32 if self.sff.dtype == np.float32 else _abi.FFM_F64
        cfg.n_episodes, cfg.n_max = self.B, self.n_max
        cfg.model = _abi.LEGACY_AC if model == "ac" else _abi.LEGACY_ACTOR_ONLY
        cfg.learn = {"none": _abi.LEARN_NONE, "exact": _abi.LEARN_EXACT}[learn]
        cfg.block_size, cfg.table_log2_capacity = self.block_size, int(table_log2_capacity)
        cfg.k_S, cfg.k_D, cfg.k_A = float(p.get("k_S", 0.0)), float(p["k_D"]), float(p.get("k_A", 0.0))
        cfg.dff_c0 = float(np.float32((1 - decay) * (1 - diffuse)))                      # ffm_ac_core.py:301
        cfg.dff_c1 = float(np.float32(decay * (1 - diffuse) / len(self.neighbors)))      # :305-309
        cfg.dff_threshold = float(np.float32(1e-4))                                      # :317
        cfg.gamma, cfg.alpha_v, cfg.alpha_h = float(p["gamma"]), float(p["alpha_v"]), float(p.get("alpha_h", 0.0))
        cfg.exit_reward, cfg.step_penalty = float(p["exit_reward"]), float(p["step_penalty"])
        cfg.collision_penalty, cfg.epsilon = float(p["collision_penalty"]), float(p.get("epsilon", 0.0))
        cfg.sff_min, cfg.sff_max = float(np.min(self.sff)), float(np.max(self.sff))
        cfg.seed, cfg.episode_base = int(seed) & 0xFFFFFFFFFFFFFFFF, int(episode_base) & 0xFFFFFFFF
        self._lib = _abi.lib()
        self._h = C.c_void_p()
        _abi.check(self._lib.ffm_legacy_create(C.byref(cfg), C.byref(self._h)))
        _abi.check(self._lib.ffm_legacy_set_fields(self._h, self.map_array.ctypes.data, self.sff.ctypes.data))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.ffm_legacy_destroy(h)
            self._h = None

    # -- state --------------------------------------------------------------------------------------
    def set_positions(self, pos_rc, n):
        pos_rc = np.ascontiguousarray(pos_rc, dtype=np.int32).reshape(self.B, self.n_max, 2)
        n = np.ascontiguousarray(n, dtype=np.int32).reshape(self.B)
        _abi.check(self._lib.ffm_legacy_set_positions(self._h, pos_rc.ctypes.data, n.ctypes.data))

    def get_positions(self):
        pos = np.empty((self.B, self.n_max, 2), np.int32)
        n = np.empty((self.B,), np.int32)
        _abi.check(self._lib.ffm_legacy_get_positions(self._h, pos.ctypes.data, n.ctypes.data))
        return pos, n

    def set_dff(self, dff):
        dff = np.ascontiguousarray(dff, dtype=np.float32).reshape(self.B, self.H, self.W)
        _abi.check(self._lib.ffm_legacy_set_dff(self._h, dff.ctypes.data))

    def get_dff(self):
        dff = np.empty((self.B, self.H, self.W), np.float32)
        _abi.check(self._lib.ffm_legacy_get_dff(self._h, dff.ctypes.data))
        return dff

    def zero_dff(self):
        _abi.check(self._lib.ffm_legacy_zero_dff(self._h))

    def update_dff(self):
        _abi.check(self._lib.ffm_legacy_update_dff(self._h))

    def set_epsilon(self, epsilon):
        _abi.check(self._lib.ffm_legacy_set_epsilon(self._h, float(epsilon)))

    def set_episode_base(self, episode_base):
        _abi.check(self._lib.ffm_legacy_set_episode_base(self._h, int(episode_base) & 0xFFFFFFFF))

    # -- stepping -----------------------------------------------------------------------------------
    def rollout(self, max_steps, record=0):
        """Up to ``max_steps`` CA steps per episode.  ``record`` > 0 returns (cells [B][record][n_max] uint32,
        counts [B][record] int32): the positions after each of the first ``record`` steps of this call."""
        if record:
            traj = np.empty((self.B, int(record), self.n_max), np.uint32)
            cnt = np.empty((self.B, int(record)), np.int32)
            _abi.check(self._lib.ffm_legacy_rollout(self._h, int(max_steps), traj.ctypes.data, cnt.ctypes.data, int(record)))
            return traj, cnt
        _abi.check(self._lib.ffm_legacy_rollout(self._h, int(max_steps), None, None, 0))
        return None

    def counters(self):
        t = np.empty((self.B,), np.int32)
        ps = np.empty((self.B,), np.uint64)
        _abi.check(self._lib.ffm_legacy_get_counters(self._h, t.ctypes.data, ps.ctypes.data))
        return t, ps

    def launch_count(self):
        return int(self._lib.ffm_legacy_launch_count(self._h))

    # -- tables -------------------------------------------------------------------------------------
    def _which(self, which):
        return {"V": _abi.LEGACY_TABLE_V, "H": _abi.LEGACY_TABLE_H}[which]

    def table_size(self, which="V"):
        n = C.c_int64()
        _abi.check(self._lib.ffm_legacy_table_size(self._h, self._which(which), C.byref(n)))
        return int(n.value)

    def get_table(self, which="V"):
        """(keys uint64 [n], rows float64 [n][width]) in insertion order."""
        n = self.table_size(which)
        width = 1 if which == "V" else self.A
        keys = np.empty((n,), np.uint64)
        rows = np.empty((n, width), np.float64)
        got = C.c_int64()
        if n:
            _abi.check(self._lib.ffm_legacy_table_get(self._h, self._which(which), keys.ctypes.data, rows.ctypes.data, n, C.byref(got)))
        return keys, rows

    def set_table(self, keys, rows, which="V", default=0.0):
        width = 1 if which == "V" else self.A
        keys = np.ascontiguousarray(keys, dtype=np.uint64).reshape(-1)
        rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(len(keys), width)
        _abi.check(self._lib.ffm_legacy_table_set(self._h, self._which(which), keys.ctypes.data, rows.ctypes.data, len(keys), float(default)))
