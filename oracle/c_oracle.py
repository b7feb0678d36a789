"""ctypes wrapper of the C restatement (oracle/c/ffm_oracle.c) -- test infrastructure / CPU baseline."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "c", "libffm_oracle.so")
_lib = None


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "c")], check=True)


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        _lib = C.CDLL(LIB)
        _lib.ffm_oracle_core_run.restype = C.c_int
    return _lib


def run_core_batch(map_array, sff, pos_rc, n, params, seed=0, episode_base=0, max_steps=1 << 30, threads=1,
                   track_dff=True, traj_steps=0, want_state=False, guard=0.0, record_moves=0):
    """B episodes of model/ffm_core.py semantics.  pos_rc int32 [B, n_max, 2], n int32 [B].
    -> (steps int32 [B], ped_steps int64 [B])  or, with want_state / traj_steps, a dict.
    guard > 0 re-draws move uniforms closer than `guard` to a CDF boundary; record_moves = T returns
    the uniforms actually used as move_draws float64 [B, T, n_max] (0 where nothing was drawn)."""
    m = np.ascontiguousarray(map_array, dtype=np.uint8)
    H, W = m.shape
    sff = np.asarray(sff)
    f64 = sff.dtype != np.float32
    sff = np.ascontiguousarray(sff, dtype=np.float64 if f64 else np.float32)
    pos_rc = np.ascontiguousarray(pos_rc, dtype=np.int32)
    n = np.ascontiguousarray(n, dtype=np.int32)
    B, n_max = pos_rc.shape[0], pos_rc.shape[1]
    p = {"k_S": 3, "k_D": 1, "diffuse": 0.2, "decay": 0.2, "neighborhood": "moore", **(params or {})}
    nbr = 4 if p["neighborhood"] == "neumann" else 8
    c0 = np.float32((1 - p["decay"]) * (1 - p["diffuse"]))
    c1 = np.float32(p["decay"] * (1 - p["diffuse"]) / nbr)
    thr = np.float32(1e-4)
    steps = np.zeros(B, np.int32)
    ped = np.zeros(B, np.int64)
    margin = np.zeros(B, np.float64)
    fpos = np.full((B, n_max), -1, np.int32) if want_state else None
    fn = np.zeros(B, np.int32) if want_state else None
    fdff = np.zeros((B, H, W), np.float32) if want_state else None
    traj = np.full((B, traj_steps, n_max), -1, np.int32) if traj_steps else None
    traj_n = np.zeros((B, traj_steps), np.int32) if traj_steps else None
    moves = np.zeros((B, record_moves, n_max), np.float64) if record_moves else None

    def ptr(a):
        return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)

    lib().ffm_oracle_core_run(
        ptr(m), ptr(sff), C.c_int(int(f64)), C.c_int(H), C.c_int(W), C.c_int(nbr), C.c_double(float(p["k_S"])),
        C.c_double(float(p["k_D"])), C.c_float(float(c0)), C.c_float(float(c1)), C.c_float(float(thr)), ptr(pos_rc), ptr(n),
        C.c_int(B), C.c_int(n_max), C.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), C.c_uint32(int(episode_base)),
        C.c_int(int(min(max_steps, 2**31 - 1))), C.c_int(int(bool(track_dff))), C.c_int(int(threads)), ptr(steps), ptr(ped),
        ptr(margin), ptr(fpos), ptr(fn), ptr(fdff), ptr(traj), ptr(traj_n), C.c_int(int(traj_steps)),
        C.c_double(float(guard)), ptr(moves), C.c_int(int(record_moves)))
    if want_state or traj_steps or record_moves:
        return dict(steps=steps, ped_steps=ped, min_margin=margin, final_pos=fpos, final_n=fn, final_dff=fdff,
                    traj=traj, traj_n=traj_n, move_draws=moves)
    return steps, ped


def geodesic(map_array, mode):
    """Geodesic SFF (float32; inf on non-walkable / unreachable cells).  mode: "bfs4" | "bfs8" | "dijkstra8"."""
    m = np.ascontiguousarray(map_array, dtype=np.uint8)
    H, W = m.shape
    out = np.empty((H, W), np.float32)
    w_diag = {"bfs4": -1.0, "bfs8": 1.0, "dijkstra8": float(np.float32(np.sqrt(2.0)))}[mode]
    L = lib()
    L.ffm_oracle_geodesic.restype = C.c_int
    rc = L.ffm_oracle_geodesic(C.c_void_p(m.ctypes.data), C.c_int(H), C.c_int(W), C.c_float(1.0), C.c_float(w_diag),
                               C.c_void_p(out.ctypes.data))
    assert rc == 0
    return out
