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


def seed_from_numpy_state():
    """Philox seed of a drop-in object whose params carry no "seed": a hash of the process-global NumPy generator's
    state, taken WITHOUT advancing it -- after ``np.random.seed(s)`` the run is reproducible and the placement draw
    that follows (initialize_agents, ffm_core.py:25) sees exactly the stream the reference would see."""
    import hashlib
    st = np.random.get_state()
    h = hashlib.blake2b(st[1].tobytes() + int(st[2]).to_bytes(4, "little"), digest_size=8).digest()
    return int.from_bytes(h, "little") >> 2


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

    DEFAULTS = CORE_DEFAULTS

    def _configure(self, cfg):
        """Hook for subclasses: fill the model-specific part of the ffm_config_t."""
        cfg.model = _abi.MODEL_CORE

    def _score_field(self, sff):
        """The SFF array the kernels score with (and its dtype decides float32 vs float64 arithmetic)."""
        return sff if sff.dtype == np.float32 else sff.astype(np.float64)

    def __init__(self, map_array, sff, n_episodes, n_max, params=None, seed=0, episode_base=0,
                 track_dff=None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("ffm_b200 needs a CUDA device (no CPU fallback)")
        self.params = dict(self.DEFAULTS) if params is None else {**self.DEFAULTS, **params}
        self.map_array = np.ascontiguousarray(np.asarray(map_array).astype(np.uint8))   # ffm_core.py:16
        if self.map_array.ndim != 2:
            raise ValueError("map_array must be 2-D")
        sff = np.asarray(sff)
        if sff.shape != self.map_array.shape:
            raise ValueError("sff shape differs from map shape")
        # NumPy computes the scores in promote(sff.dtype, float32): float32 stays, all else -> float64
        self.sff = np.ascontiguousarray(self._score_field(sff))
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
        cfg.k_S, cfg.k_D = float(self.params.get("k_S", 0.0)), k_D
        # scalars are formed in Python floats and cast to float32 when they meet the float32 field
        cfg.dff_c0 = float(np.float32((1 - decay) * (1 - diffuse)))                      # ffm_core.py:109
        cfg.dff_c1 = float(np.float32(decay * (1 - diffuse) / len(self.neighbors)))      # ffm_core.py:113
        cfg.dff_threshold = float(np.float32(1e-4))                                      # ffm_core.py:116
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.episode_base = int(episode_base) & 0xFFFFFFFF
        self._seed_base = (int(seed), int(episode_base))
        self._configure(cfg)
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

    def place(self, n, exit_pos=None, radius=None):
        """Device-side initialize_agents(): n[e] pedestrians per episode on distinct free cells, uniformly without
        replacement (optionally within L1 `radius` of `exit_pos`, count clamped like ffm_unified.py:160-162),
        keyed by the global episode id.  Resets DFF and counters."""
        n = np.ascontiguousarray(np.broadcast_to(np.asarray(n, dtype=np.int32), (self.B,)))
        er, ec, rad = (-1, -1, -1) if exit_pos is None or radius is None else (int(exit_pos[0]), int(exit_pos[1]), int(radius))
        _abi.check(self._lib.ffm_place(self._h, _ptr(n), er, ec, rad, _stream()))   # synchronous; reports device-side errors

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

    def update_dff(self):
        """One stand-alone decay + diffusion pass (update_dff(), ffm_core.py:106-117) over every episode's DFF, on the device."""
        _abi.check(self._lib.ffm_update_dff(self._h, _stream()))

    def move_probs(self):
        """Probe: (probs float64 [B, n_max, neighbours+1] in slot order, kind int32 [B, n_max]) for the current state
        -- the distribution each pedestrian samples from in the next step (0 no request, 1 forced exit, 2 sampled)."""
        A = len(self.neighbors) + 1
        probs = np.empty((self.B, self.n_max, A), np.float64)
        kind = np.empty((self.B, self.n_max), np.int32)
        _abi.check(self._lib.ffm_move_probs(self._h, _ptr(probs), _ptr(kind), _abi.FFM_HOST, _stream()))
        return probs, kind

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
    def rollout(self, max_steps, draws=None, record=0, record_buffer=False, compact_cap=None):
        """Run up to ``max_steps`` CA steps per episode (asynchronous on the current stream).

        draws   optional dict(move=float64 [B, T, n_max], conflict=float64 [B, T, H*W, 2],
                first_step=int) of recorded uniforms (CUDA tensors) overriding the Philox streams
        record  > 0: also return (traj_cells uint32-as-int32 [B, record, n_max], traj_n int32
                [B, record]) CUDA tensors with the positions after each step
        compact_cap  with record > 0: return the COMPACT record instead -- dict(ctraj=int16 [B, compact_cap, 2] (row, col)
                pairs, the rows of consecutive steps back to back (each padded with -1 to a multiple of 4 entries),
                off=int32 [B, record + 1] CSR offsets, n=int32 [B, record] row lengths): 4 bytes per pedestrian-step
                (see unpack_trajectory)
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
            dev = f"cuda:{self.device}"
            o = _abi.RolloutOut()
            o.traj_steps = int(record)
            if record_buffer:
                # rollout buffer of the unified models: state / action / reward per agent-step, SoA [B, T, n_max]
                rs = torch.zeros((self.B, record, self.n_max), dtype=torch.int32, device=dev)
                ra = torch.zeros((self.B, record, self.n_max), dtype=torch.uint8, device=dev)
                rr = torch.zeros((self.B, record, self.n_max), dtype=torch.float32, device=dev)
                rl = torch.zeros((self.B, self.n_max), dtype=torch.int32, device=dev)
                o.rec_state, o.rec_action, o.rec_reward, o.rec_len = rs.data_ptr(), ra.data_ptr(), rr.data_ptr(), rl.data_ptr()
                ret = dict(state=rs, action=ra, reward=rr, length=rl)
            elif compact_cap:
                cap = (int(compact_cap) + 3) & ~3
                ct = torch.empty((self.B, cap, 2), dtype=torch.int16, device=dev)
                off = torch.full((self.B, record + 1), -1, dtype=torch.int32, device=dev)
                cnt = torch.zeros((self.B, record), dtype=torch.int32, device=dev)
                o.ctraj, o.ctraj_off, o.ctraj_cap, o.traj_n = ct.data_ptr(), off.data_ptr(), cap, cnt.data_ptr()
                ret = dict(ctraj=ct, off=off, n=cnt)
            else:
                cells = torch.zeros((self.B, record, self.n_max), dtype=torch.int32, device=dev)
                cnt = torch.zeros((self.B, record), dtype=torch.int32, device=dev)
                o.traj_cells, o.traj_n = cells.data_ptr(), cnt.data_ptr()
                ret = (cells, cnt)
            optr = C.byref(o)
        _abi.check(self._lib.ffm_rollout(self._h, int(max_steps), dptr, optr, _stream()))
        return ret

    def set_episode_base(self, episode_base):
        _abi.check(self._lib.ffm_set_episode_base(self._h, int(episode_base) & 0xFFFFFFFF))
        self._seed_base = (self._seed_base[0], int(episode_base))

    # -- introspection -------------------------------------------------------------------------
    @property
    def launch_count(self):
        return int(self._lib.ffm_launch_count(self._h))

    def kernel_info(self):
        a, b, c, d = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _abi.check(self._lib.ffm_kernel_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        info = dict(smem_bytes=a.value, threads=b.value, ctas_per_sm=c.value, fields_in_smem=bool(d.value))
        _abi.check(self._lib.ffm_cluster_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        info.update(cluster=a.value, max_clusters=b.value, score_in_smem=bool(c.value),
                    name={1: "ffm_cell_rollout_kernel", 0: "ffm_core_rollout_kernel"}.get(
                        d.value, "ffm_mcq_rollout_kernel" if isinstance(self, McqSim) else "ffm_unified_rollout_kernel"))
        return info


def unpack_trajectory(ctraj, off, n, steps=None):
    """One episode of a compact record (NumPy: ctraj int16 [cap, 2], off int32 [T + 1], n int32 [T]) -> the list run()
    collects: one int64 [n_t, 2] array of (row, col) per step (ffm_core.py:125; `np.array(buffer, dtype=object)` of it is
    main.py:52's positions.npy)."""
    T = len(n) if steps is None else int(steps)
    out = []
    for t in range(T):
        if off[t] < 0 or off[t + 1] < 0:
            raise ValueError(f"trajectory record overflowed at step {t}: raise compact_cap")
        out.append(ctraj[off[t]:off[t] + n[t]].astype(np.int64))
    return out


UNIFIED_DEFAULTS = {                      # ffm_unified.py:36-53
    "k_S": 10, "k_D": 1, "k_A": 10, "diffuse": 0.2, "decay": 0.2, "neighborhood": "neumann",
    "alpha_v": 0.1, "gamma": 0.95, "exit_reward": 100.0, "step_penalty": 0.0, "collision_penalty": -1.0,
    "block_size": 5, "alpha_h": 0.1, "epsilon": 0.0,
}
TRAINED_DEFAULTS = {"k_D": 1, "k_A": 10, "diffuse": 0.2, "decay": 0.2, "neighborhood": "neumann", "block_size": 5}  # ffm_trained_core.py:29-36
_MODES = {"critic_only": _abi.MODEL_UNIFIED_CRITIC, "actor_only": _abi.MODEL_UNIFIED_ACTOR, "both": _abi.MODEL_UNIFIED_BOTH,
          "trained": _abi.MODEL_TRAINED}
_LEARN = {"none": _abi.LEARN_NONE, "exact": _abi.LEARN_EXACT, "batched": _abi.LEARN_BATCHED}


class UnifiedSim(BatchSim):
    """B episodes of ``FloorFieldModelUnified`` (model/ffm_unified.py) or of the trained-actor model
    (model/ffm_trained_core.py) sharing map, SFF, parameters and the V / H tables.

    mode   "critic_only" | "actor_only" | "both" | "trained"
    learn  "exact"   the reference's sequential per-agent table updates (n_episodes must be 1)
           "batched" synchronous batched TD: rollouts accumulate TD errors / visit counts / alpha_h*delta into
                     ``dV`` / ``dN`` / ``dH`` against frozen tables; ``apply_deltas()`` (after the caller's
                     all-reduce) folds them in
           "none"    frozen tables
    Tables are dense: state id = (bx*nby + by)*256 + rU*64 + rD*16 + rL*4 + rR (``key_to_id``).
    """

    DEFAULTS = UNIFIED_DEFAULTS

    def __init__(self, map_array, sff, n_episodes, n_max, mode="critic_only", learn="exact", params=None, seed=0,
                 episode_base=0, device=None):
        if mode not in _MODES:
            raise ValueError(f"learning_mode must be one of {sorted(_MODES)}, got {mode}")
        self.mode, self.learn = mode, learn
        if mode == "trained":
            self.DEFAULTS = TRAINED_DEFAULTS
        super().__init__(map_array, sff, n_episodes, n_max, params, seed, episode_base, True, device)
        S, A = C.c_int32(), C.c_int32()
        _abi.check(self._lib.ffm_tables_shape(self._h, C.byref(S), C.byref(A)))
        self.S, self.A = S.value, A.value
        self.bs = int(self.params["block_size"])
        self.nby = -(-self.W // self.bs)
        self.delta = self.dV = self.dN = self.dF = self.dH = None
        if learn == "batched":
            self.bind_deltas(self.new_delta_buffer())

    def _score_field(self, sff):
        if self.mode == "critic_only":
            return sff if sff.dtype == np.float32 else sff.astype(np.float64)           # file dtype kept (ffm_unified.py:70)
        return np.where(np.isinf(sff), 0.0, sff).astype(np.float32)                       # ffm_unified.py:72-76

    def _configure(self, cfg):
        p = self.params
        cfg.model = _MODES[self.mode]
        cfg.learn = _LEARN[self.learn]
        cfg.block_size = int(p["block_size"])
        cfg.k_A = float(p.get("k_A", 0.0))
        cfg.gamma, cfg.alpha_v, cfg.alpha_h = float(p.get("gamma", 0.0)), float(p.get("alpha_v", 0.0)), float(p.get("alpha_h", 0.0))
        cfg.exit_reward = float(p.get("exit_reward", 0.0))
        cfg.step_penalty = float(p.get("step_penalty", 0.0))
        cfg.collision_penalty = float(p.get("collision_penalty", 0.0))
        cfg.epsilon = float(p.get("epsilon", 0.0) or 0.0)
        cfg.sff_min, cfg.sff_max = float(np.min(self.sff)), float(np.max(self.sff))       # ffm_unified.py:425-426

    # -- tables ----------------------------------------------------------------------------------
    def key_to_id(self, key):
        r, (bx, by) = key
        return (int(bx) * self.nby + int(by)) * 256 + int(r[0]) * 64 + int(r[1]) * 16 + int(r[2]) * 4 + int(r[3])

    def id_to_key(self, sid):
        blk, code = divmod(int(sid), 256)
        return ((code >> 6, (code >> 4) & 3, (code >> 2) & 3, code & 3), (blk // self.nby, blk % self.nby))

    def get_tables(self):
        """-> (V float64 [S], v_seen bool [S], H float64 [S, A], h_seen bool [S]) NumPy copies."""
        V = np.empty(self.S, np.float64); vs = np.empty(self.S, np.uint8)
        H = np.empty((self.S, self.A), np.float64); hs = np.empty(self.S, np.uint8)
        _abi.check(self._lib.ffm_tables_get(self._h, _ptr(V), _ptr(vs), _ptr(H), _ptr(hs), _abi.FFM_HOST, _stream()))
        return V, vs.astype(bool), H, hs.astype(bool)

    def set_tables(self, V=None, v_seen=None, H=None, h_seen=None):
        conv = lambda a, dt, shape: None if a is None else np.ascontiguousarray(np.asarray(a, dtype=dt).reshape(shape))
        V, H = conv(V, np.float64, (self.S,)), conv(H, np.float64, (self.S, self.A))
        vs, hs = conv(v_seen, np.uint8, (self.S,)), conv(h_seen, np.uint8, (self.S,))
        _abi.check(self._lib.ffm_tables_set(self._h, _ptr(V), _ptr(vs), _ptr(H), _ptr(hs), _abi.FFM_HOST, _stream()))

    def v_dict(self):
        """The reference's ``dict(self.V)`` (ffm_unified.py:821)."""
        V, vs, _, _ = self.get_tables()
        return {self.id_to_key(s): float(V[s]) for s in np.flatnonzero(vs)}

    def h_dict(self):
        """The reference's ``dict(self.H)`` (ffm_unified.py:855): rows as lists of Python floats."""
        _, _, H, hs = self.get_tables()
        return {self.id_to_key(s): [float(v) for v in H[s]] for s in np.flatnonzero(hs)}

    def load_v_dict(self, d):
        V = np.zeros(self.S, np.float64); vs = np.zeros(self.S, np.uint8)
        for k, v in d.items():
            V[self.key_to_id(k)] = v
            vs[self.key_to_id(k)] = 1
        self.set_tables(V=V, v_seen=vs)

    def load_h_dict(self, d):
        H = np.zeros((self.S, self.A), np.float64); hs = np.zeros(self.S, np.uint8)
        for k, v in d.items():
            if len(v) == self.A:
                H[self.key_to_id(k)] = v
                hs[self.key_to_id(k)] = 1
        self.set_tables(H=H, h_seen=hs)

    def set_epsilon(self, epsilon):
        _abi.check(self._lib.ffm_set_epsilon(self._h, float(epsilon)))

    def bind_dynamic(self, dyn):
        """Read epsilon / episode_base from a device buffer from now on (CUDA-graph replays; see RoundParams); None unbinds."""
        self._dyn = dyn
        _abi.check(self._lib.ffm_bind_dynamic(self._h, None if dyn is None else _ptr(dyn)))

    # -- batched learning: ONE flat float64 buffer [dV | dN | dF | dH] per sync, so that the cross-GPU exchange is a
    #    single all-reduce(sum) -----------------------------------------------------------------------
    def new_delta_buffer(self):
        return torch.zeros((3 + self.A) * self.S, dtype=torch.float64, device=f"cuda:{self.device}")

    def bind_deltas(self, flat):
        """Make ``flat`` (from new_delta_buffer) the buffer the following rollouts accumulate into."""
        S = self.S
        assert flat.is_cuda and flat.dtype == torch.float64 and flat.is_contiguous() and flat.numel() == (3 + self.A) * S
        self.delta = flat
        self.dV, self.dN, self.dF, self.dH = flat[:S], flat[S:2 * S], flat[2 * S:3 * S], flat[3 * S:].view(S, self.A)
        _abi.check(self._lib.ffm_tables_bind_deltas(self._h, _ptr(self.dV), _ptr(self.dN), _ptr(self.dF), _ptr(self.dH)))

    def apply_deltas(self):
        """Fold the bound delta buffer in (call after all-reducing it): V += (1-(1-alpha_v)^n) * mean TD error,
        H += dH, touched keys marked present, deltas zeroed, H extremes refreshed.  Stream-ordered."""
        _abi.check(self._lib.ffm_tables_apply_deltas(self._h, _stream()))


class RoundParams:
    """The two per-round parameters of a unified-model rollout as a device-resident struct {double epsilon; uint32
    episode_base; uint32 pad}, so that a CUDA graph holding the round's launches can be replayed with new values:
    ``set()`` copies them stream-ordered."""

    def __init__(self, device):
        self.dev = torch.zeros(16, dtype=torch.uint8, device=device)

    def set(self, epsilon, episode_base):
        h = np.zeros(16, np.uint8)
        h[:8] = np.frombuffer(np.float64(epsilon).tobytes(), np.uint8)
        h[8:12] = np.frombuffer(np.uint32(int(episode_base) & 0xFFFFFFFF).tobytes(), np.uint8)
        self.dev.copy_(torch.from_numpy(h))          # pageable source: staged before the call returns, stream-ordered on the device


def rollout_returns(reward, length, gamma):
    """Discounted returns G[b, t, n] = r + gamma * G[b, t+1, n] (float64) of a rollout buffer
    (reward float32 [B, T, N], length int32 [B, N], CUDA tensors) -- ffm_learning_core.py:262-278."""
    assert reward.is_cuda and reward.dtype == torch.float32 and reward.is_contiguous() and reward.dim() == 3
    assert length.is_cuda and length.dtype == torch.int32 and length.is_contiguous()
    B, T, N = reward.shape
    out = torch.empty((B, T, N), dtype=torch.float64, device=reward.device)
    _abi.check(_abi.lib().ffm_rollout_returns(_ptr(reward), _ptr(length), B, T, N, float(gamma), _ptr(out),
                                              reward.device.index or 0, _stream()))
    return out


MCQ_DEFAULTS = {"k_S": 3.0, "k_D": 1.0, "k_Q": 1.0, "diffuse": 0.2, "decay": 0.2, "neighborhood": "neumann",
                "step_penalty": 0.0, "stop_penalty": 0.0, "collision_penalty": 0.0, "exit_reward": 100.0,
                "timeout_penalty": 50.0, "max_steps": 500}        # ffm_learning_core.py:45-59


class McqSim(BatchSim):
    """B episodes of the target-centric Monte-Carlo Q-learning model (model/ffm_learning_core.py).

    learn  "exact"    the reference's reverse Monte-Carlo backups at arrivals / timeouts (n_episodes must be 1)
           "none"     frozen Q table, any number of episodes
           "batched"  frozen table during the rollout, every path's finish order recorded; afterwards either
                      ``backup_ordered()`` (the reference's backups in episode order: exact when the policy did not read
                      Q, i.e. beta = 1 -- coverage pretrain, warm-up episodes) or ``accumulate()`` + ``fold()`` (returns
                      reduced per (state, action) and folded in with the visit count; ``McqBatchedLearner`` adds the
                      multi-GPU exchange)
    The Q dict is a hash table on the device: 64-bit key = ((tx//3)*nby + ty//3) * 4**9 + sum(v_i * 4**i) over the 3x3
    window around the target cell (``key_to_id``); rows float32 [5] in the reference's action order
    FROM_UP/DOWN/LEFT/RIGHT/SELF.  ``q_log2_capacity``: log2 of its slots (default 21; rows may fill half of them).
    """

    DEFAULTS = MCQ_DEFAULTS

    def __init__(self, map_array, sff, n_episodes, n_max, learn="exact", params=None, seed=0, episode_base=0,
                 alpha=0.1, gamma=0.99, device=None, q_log2_capacity=0):
        self.learn, self.alpha, self.gamma = learn, float(alpha), float(gamma)
        self._q_log2 = int(q_log2_capacity)
        if params is not None and params.get("neighborhood", "neumann") != "neumann":
            params = {**params, "neighborhood": "neumann"}          # the model forces von Neumann (ffm_learning_core.py:72-73)
        super().__init__(map_array, sff, n_episodes, n_max, params, seed, episode_base, True, device)
        cap = C.c_int64()
        _abi.check(self._lib.ffm_q_shape(self._h, C.byref(cap)))
        self.q_capacity = cap.value
        self.nby = -(-self.W // 3)

    def _configure(self, cfg):
        p = self.params
        cfg.model = _abi.MODEL_MCQ
        cfg.learn = _LEARN[self.learn]
        cfg.k_A = float(p["k_Q"])
        cfg.alpha_v, cfg.gamma = self.alpha, self.gamma
        cfg.step_penalty, cfg.stop_penalty = float(p["step_penalty"]), float(p["stop_penalty"])
        cfg.collision_penalty, cfg.exit_reward = float(p["collision_penalty"]), float(p["exit_reward"])
        cfg.timeout_penalty, cfg.step_cap = float(p["timeout_penalty"]), int(p["max_steps"])
        cfg.q_log2_capacity = self._q_log2
        # _update_dff (ffm_learning_core.py:307-321) is always Moore: the neighbour weight divides by 8
        decay, diffuse = float(p["decay"]), float(p["diffuse"])
        cfg.dff_c0 = float(np.float32((1.0 - decay) * (1.0 - diffuse)))
        cfg.dff_c1 = float(np.float32(decay * (1.0 - diffuse) / 8))

    def set_beta(self, beta):
        _abi.check(self._lib.ffm_set_beta(self._h, float(beta)))

    def finalize_timeouts(self):
        """finalize_timeouts() (ffm_learning_core.py:326-360) before the step cap."""
        _abi.check(self._lib.ffm_mcq_finalize_timeouts(self._h, _stream()))

    def key_to_id(self, key):
        cells, (bx, by) = key
        return (int(bx) * self.nby + int(by)) * 4 ** 9 + sum(int(v) << (2 * k) for k, v in enumerate(bytes(cells)))

    def id_to_key(self, sid):
        blk, code = divmod(int(sid), 4 ** 9)
        return (bytes((code >> (2 * k)) & 3 for k in range(9)), (blk // self.nby, blk % self.nby))

    def get_q(self):
        """-> (ids int64 [K] ascending, rows float32 [K, 5]) of the rows that exist."""
        keys = np.empty(self.q_capacity, np.uint64)
        rows = np.empty((self.q_capacity, 5), np.float32)
        _abi.check(self._lib.ffm_q_get(self._h, _ptr(keys), _ptr(rows), _abi.FFM_HOST, _stream()))
        used = np.flatnonzero(keys != np.uint64(0xFFFFFFFFFFFFFFFF))
        order = used[np.argsort(keys[used], kind="stable")]
        return keys[order].astype(np.int64), rows[order].copy()

    def q_dict(self):
        """The reference's ``self.Q``: {(combined3x3 bytes, (bx, by)): float32[5]}."""
        ids, rows = self.get_q()
        return {self.id_to_key(i): rows[k] for k, i in enumerate(ids)}

    def load_q_dict(self, d):
        keys = np.fromiter((self.key_to_id(k) for k in d), dtype=np.uint64, count=len(d))
        rows = np.ascontiguousarray(np.stack([np.asarray(v, np.float32) for v in d.values()]) if len(d) else np.zeros((0, 5), np.float32))
        _abi.check(self._lib.ffm_q_set(self._h, _ptr(keys), _ptr(rows), len(d), _abi.FFM_HOST, _stream()))

    # -- coverage pretrain / batched learning ------------------------------------------------------------------------
    def set_forced(self, target_rc, from_dir, step_cap):
        """Teacher-forced first transition of every episode (force_first_step_and_roll,
        run_coverage_pretrain_and_training.py:91-166); call after set_positions placed one agent per episode on its
        source cell.  target_rc int [B, 2] (row < 0: none), from_dir int [B] (FROM_* 0..4), step_cap int [B]."""
        t = np.asarray(target_rc, np.int64).reshape(self.B, 2)
        cell = np.where(t[:, 0] < 0, -1, t[:, 0] * self.W + t[:, 1]).astype(np.int32)
        fd = np.ascontiguousarray(np.asarray(from_dir, np.int32).reshape(self.B))
        cap = np.ascontiguousarray(np.asarray(step_cap, np.int32).reshape(self.B))
        _abi.check(self._lib.ffm_mcq_set_forced(self._h, _ptr(cell), _ptr(fd), _ptr(cap), _stream()))

    def backup_ordered(self):
        _abi.check(self._lib.ffm_mcq_backup_ordered(self._h, _stream()))

    def accumulate(self):
        _abi.check(self._lib.ffm_mcq_accumulate(self._h, _stream()))

    def fold(self):
        _abi.check(self._lib.ffm_mcq_fold(self._h, _stream()))

    def delta_device(self):
        return f"cuda:{self.device}"

    def export_deltas(self, capacity, out=None):
        """-> (keys int64 [capacity], rows float64 [capacity, 10], count int32 [1]) CUDA tensors: the rows touched since the
        last fold, by key; the local delta tables are cleared.  ``out``: a flat float64 CUDA tensor of 1 + 11 * capacity
        elements to write into ([count | keys | rows], the views are returned) so that the whole list travels as one message."""
        dev = f"cuda:{self.device}"
        if out is None:
            out = torch.zeros(1 + 11 * capacity, dtype=torch.float64, device=dev)
        assert out.is_cuda and out.dtype == torch.float64 and out.is_contiguous() and out.numel() == 1 + 11 * capacity
        count = out[:1].view(torch.int32)[:1]
        keys = out[1:1 + capacity].view(torch.int64)
        rows = out[1 + capacity:].view(capacity, 10)
        _abi.check(self._lib.ffm_mcq_export_deltas(self._h, _ptr(keys), _ptr(rows), int(capacity), _ptr(count), _stream()))
        return keys, rows, count

    def import_deltas(self, keys, rows, count, count_dev=None):
        """Add one rank's exported list into the local delta tables: ``count`` rows, or min(count, *count_dev) when the
        exporter's count arrives on the device with its list."""
        assert keys.is_cuda and rows.is_cuda and keys.is_contiguous() and rows.is_contiguous()
        self._keep = [keys, rows, count_dev]
        _abi.check(self._lib.ffm_mcq_import_deltas(self._h, _ptr(keys), _ptr(rows), int(count), _ptr(count_dev), _stream()))
