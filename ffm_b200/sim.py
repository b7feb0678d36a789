"""BatchSim: B independent evacuation episodes on one map, stepped by the persistent rollout kernel.

Thin host wrapper over the C ABI (include/ffm_b200.h); torch supplies caller-visible device
buffers and the CUDA stream, nothing else.
"""
import ctypes as C

import numpy as np
import torch

from . import _abi

NEUMANN = [(-1, 0), (1, 0), (0, -1), (0, 1)]                                     # ffm_core.py:30
MOORE = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]   # ffm_core.py:32-34
CORE_DEFAULTS = {"k_S": 3, "k_D": 1, "diffuse": 0.2, "decay": 0.2, "neighborhood": "moore"}  # ffm_core.py:8-14


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


class BatchSim:
    """B episodes of ``FloorFieldModel`` (model/ffm_core.py) sharing map, SFF and parameters.

    params        the reference's ``params`` dict (merged over its defaults, ffm_core.py:8-15)
    seed          Philox key; draws are keyed (episode_base + e, step, stream, entity), so results do
                  not depend on how episodes are spread over handles / GPUs
    track_dff     None: track iff k_D != 0.  False is only legal when k_D == 0 (the "SFF only"
                  configuration): the DFF cannot influence any move and is not computed at all.
    """

    def __init__(self, map_array, sff, n_episodes, n_max, params=None, seed=0, episode_base=0,
                 track_dff=None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("ffm_b200 needs a CUDA device (no CPU fallback)")
        self.params = dict(CORE_DEFAULTS) if params is None else {**CORE_DEFAULTS, **params}
        self.map_array = np.ascontiguousarray(np.asarray(map_array).astype(np.uint8))   # ffm_core.py:16
        if self.map_array.ndim != 2:
            raise ValueError("map_array must be 2-D")
        sff = np.asarray(sff)
        if sff.shape != self.map_array.shape:
            raise ValueError("sff shape differs from map shape")
        # NumPy computes the scores in promote(sff.dtype, float32): float32 stays, all else -> float64
        self.sff = np.ascontiguousarray(sff if sff.dtype == np.float32 else sff.astype(np.float64))
        self.H, self.W = self.map_array.shape
        self.B, self.n_max = int(n_episodes), int(n_max)
        self.neighbors = list(NEUMANN) if self.params["neighborhood"] == "neumann" else list(MOORE)
        self.device = torch.cuda.current_device() if device is None else int(device)
        k_D = float(self.params["k_D"])
        if track_dff is None:
            track_dff = k_D != 0.0
        self.track_dff = bool(track_dff)
        decay, diffuse = self.params["decay"], self.params["diffuse"]
        cfg = _abi.Config()
        cfg.abi_version = _abi.ABI_VERSION
        cfg.device = self.device
        cfg.height, cfg.width = self.H, self.W
        cfg.neighborhood = len(self.neighbors)
        cfg.sff_dtype = _abi.FFM_F32 if self.sff.dtype == np.float32 else _abi.FFM_F64
        cfg.n_episodes, cfg.n_max = self.B, self.n_max
        cfg.track_dff = int(self.track_dff)
        cfg.k_S, cfg.k_D = float(self.params["k_S"]), k_D
        # scalars are formed in Python floats and cast to float32 when they meet the float32 field
        cfg.dff_c0 = float(np.float32((1 - decay) * (1 - diffuse)))                      # ffm_core.py:109
        cfg.dff_c1 = float(np.float32(decay * (1 - diffuse) / len(self.neighbors)))      # ffm_core.py:113
        cfg.dff_threshold = float(np.float32(1e-4))                                      # ffm_core.py:116
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.episode_base = int(episode_base) & 0xFFFFFFFF
        self._lib = _abi.lib()
        self._h = C.c_void_p()
        _abi.check(self._lib.ffm_create(C.byref(cfg), C.byref(self._h)))
        self._keep = []
        _abi.check(self._lib.ffm_set_fields(self._h, _ptr(self.map_array), _ptr(self.sff), _abi.FFM_HOST, _stream()))

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.ffm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state ---------------------------------------------------------------------------------
    def set_positions(self, pos_rc, n):
        """pos_rc int32 [B, n_max, 2] (row, col), n int32 [B]; NumPy (host) or CUDA tensors (device).
        Resets DFF and the step counters (reset(): ffm_unified.py:800-812)."""
        if isinstance(pos_rc, torch.Tensor):
            assert pos_rc.is_cuda and n.is_cuda and pos_rc.dtype == torch.int32 and n.dtype == torch.int32
            assert pos_rc.is_contiguous() and tuple(pos_rc.shape) == (self.B, self.n_max, 2)
            space = _abi.FFM_DEVICE
        else:
            pos_rc = np.ascontiguousarray(pos_rc, dtype=np.int32)
            n = np.ascontiguousarray(n, dtype=np.int32)
            assert pos_rc.shape == (self.B, self.n_max, 2) and n.shape == (self.B,)
            space = _abi.FFM_HOST
        _abi.check(self._lib.ffm_set_positions(self._h, _ptr(pos_rc), _ptr(n), space, _stream()))
        if space == _abi.FFM_HOST:
            torch.cuda.current_stream().synchronize()   # host buffers may be freed by the caller

    def get_positions(self):
        """-> (pos_rc int32 [B, n_max, 2] with -1 padding, n int32 [B]) as NumPy arrays."""
        pos = np.empty((self.B, self.n_max, 2), dtype=np.int32)
        n = np.empty((self.B,), dtype=np.int32)
        _abi.check(self._lib.ffm_get_positions(self._h, _ptr(pos), _ptr(n), _abi.FFM_HOST, _stream()))
        return pos, n

    def get_dff(self):
        dff = np.empty((self.B, self.H, self.W), dtype=np.float32)
        _abi.check(self._lib.ffm_get_dff(self._h, _ptr(dff), _abi.FFM_HOST, _stream()))
        return dff

    def set_dff(self, dff):
        dff = np.ascontiguousarray(dff, dtype=np.float32)
        assert dff.shape == (self.B, self.H, self.W)
        _abi.check(self._lib.ffm_set_dff(self._h, _ptr(dff), _abi.FFM_HOST, _stream()))
        torch.cuda.current_stream().synchronize()

    def counters(self):
        """-> (steps int32 [B], ped_steps int64 [B]) since the last set_positions."""
        steps = np.empty((self.B,), dtype=np.int32)
        ps = np.empty((self.B,), dtype=np.int64)
        _abi.check(self._lib.ffm_get_counters(self._h, _ptr(steps), _ptr(ps), _abi.FFM_HOST, _stream()))
        return steps, ps

    def counters_into(self, steps_t, ped_steps_t):
        """Device-side copy of the counters into caller tensors (no host sync)."""
        _abi.check(self._lib.ffm_get_counters(self._h, _ptr(steps_t), _ptr(ped_steps_t), _abi.FFM_DEVICE, _stream()))

    # -- stepping ------------------------------------------------------------------------------
    def rollout(self, max_steps, draws=None, record=0):
        """Run up to ``max_steps`` CA steps per episode (asynchronous on the current stream).

        draws   optional dict(move=float64 [B, T, n_max], conflict=float64 [B, T, H*W, 2],
                first_step=int) of recorded uniforms (CUDA tensors) overriding the Philox streams
        record  > 0: also return (traj_cells uint32-as-int32 [B, record, n_max], traj_n int32
                [B, record]) CUDA tensors with the positions after each step
        """
        dptr = None
        if draws is not None:
            d = _abi.Draws()
            mv, cf = draws.get("move"), draws.get("conflict")
            steps = None
            if mv is not None:
                assert mv.is_cuda and mv.dtype == torch.float64 and mv.is_contiguous()
                assert mv.shape[0] == self.B and mv.shape[2] == self.n_max
                steps = mv.shape[1]
                d.move = mv.data_ptr()
            if cf is not None:
                assert cf.is_cuda and cf.dtype == torch.float64 and cf.is_contiguous()
                assert cf.shape[0] == self.B and cf.shape[2] == self.H * self.W and cf.shape[3] == 2
                assert steps is None or steps == cf.shape[1]
                steps = cf.shape[1]
                d.conflict = cf.data_ptr()
            d.steps = int(steps or 0)
            d.first_step = int(draws.get("first_step", 0))
            d.space = _abi.FFM_DEVICE
            dptr = C.byref(d)
            self._keep = [mv, cf]
        optr, ret = None, None
        if record:
            cells = torch.zeros((self.B, record, self.n_max), dtype=torch.int32, device=f"cuda:{self.device}")
            cnt = torch.zeros((self.B, record), dtype=torch.int32, device=f"cuda:{self.device}")
            o = _abi.RolloutOut()
            o.traj_cells, o.traj_n, o.traj_steps = cells.data_ptr(), cnt.data_ptr(), int(record)
            optr = C.byref(o)
            ret = (cells, cnt)
        _abi.check(self._lib.ffm_rollout(self._h, int(max_steps), dptr, optr, _stream()))
        return ret

    # -- introspection -------------------------------------------------------------------------
    @property
    def launch_count(self):
        return int(self._lib.ffm_launch_count(self._h))

    def kernel_info(self):
        a, b, c, d = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _abi.check(self._lib.ffm_kernel_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(smem_bytes=a.value, threads=b.value, ctas_per_sm=c.value, fields_in_smem=bool(d.value))
