"""Static-floor-field generation on the GPU (csrc/ffm_sff_kernels.cuh through ffm_sff_generate).

``generate_sff(map, metric)`` covers the reference's generators (Create_SFF.py, create_12x12_map_and_sff.py:
obstacle-blind ``"L1" | "L2" | "Linf"``) and the obstacle-aware fields of the north star
(``"bfs4" | "bfs8" | "dijkstra8"``).  ``maps`` may be one map [H, W] or a batch [n, H, W].
"""
import ctypes as C

import numpy as np
import torch

from . import _abi

MODES = {"L1": _abi.SFF_L1, "L2": _abi.SFF_L2, "Linf": _abi.SFF_LINF, "bfs4": _abi.SFF_BFS4, "bfs8": _abi.SFF_BFS8,
         "dijkstra8": _abi.SFF_DIJKSTRA8}


def generate_sff(maps, metric="L1", dtype=np.float64, device=None, return_rounds=False):
    """-> SFF array(s) of ``dtype`` (float32 | float64), inf on non-walkable / unreachable cells.
    NumPy in -> NumPy out (host buffers); CUDA uint8 tensor in -> CUDA tensor out."""
    if metric not in MODES:
        raise ValueError(f"metric must be one of {sorted(MODES)}")
    if not torch.cuda.is_available():
        raise RuntimeError("ffm_b200 needs a CUDA device (no CPU fallback)")
    dt = np.dtype(dtype)
    if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise ValueError("dtype must be float32 or float64")
    code = _abi.FFM_F32 if dt == np.float32 else _abi.FFM_F64
    device = torch.cuda.current_device() if device is None else int(device)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rounds = C.c_int32(0)
    if isinstance(maps, torch.Tensor):
        assert maps.is_cuda and maps.dtype == torch.uint8 and maps.is_contiguous()
        shape = tuple(maps.shape)
        single = maps.dim() == 2
        n = 1 if single else shape[0]
        H, W = shape[-2:]
        out = torch.empty(shape, dtype=torch.float32 if code == _abi.FFM_F32 else torch.float64, device=maps.device)
        _abi.check(_abi.lib().ffm_sff_generate(C.c_void_p(maps.data_ptr()), n, H, W, MODES[metric], code,
                                               C.c_void_p(out.data_ptr()), _abi.FFM_DEVICE, device, stream, C.byref(rounds) if return_rounds else None))
    else:
        m = np.ascontiguousarray(np.asarray(maps).astype(np.uint8))
        single = m.ndim == 2
        n = 1 if single else m.shape[0]
        H, W = m.shape[-2:]
        out = np.empty(m.shape, dtype=dt)
        _abi.check(_abi.lib().ffm_sff_generate(C.c_void_p(m.ctypes.data), n, H, W, MODES[metric], code,
                                               C.c_void_p(out.ctypes.data), _abi.FFM_HOST, device, stream, C.byref(rounds) if return_rounds else None))
    return (out, rounds.value) if return_rounds else out
