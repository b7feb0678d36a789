"""Drop-in ``FloorFieldModelActorOnly`` with the interface of the reference's legacy ``model/ffm_actor_only.py``.

    FloorFieldModelActorOnly(map_array, sff_path, N, pretrained_v_path=None, params=None)   ffm_actor_only.py:21-80
    .step() .run(save_prefix, save_interval, max_steps, return_trajectory) .reset() .update_dff()
    .initialize_agents() .get_neighbors() ._encode_state(x, y, occupancy) .set_epsilon(e)
    .get_v_table() .get_v_table_size() .get_h_table() .get_h_table_size()
    attributes: params, map_array, sff, dff, N, positions, neighbors, V, H, alpha_v, gamma, alpha_h, epsilon, initial_v_size

The step reproduces the reference as written, including the repetition of its decision block once per neighbour slot
(:214-355; csrc/ffm_legacy.cu).  V and H live on the device as hash tables; ``V`` / ``H`` / ``get_*_table()`` materialise the
reference's dicts keyed by ``pickle.dumps((tuple(state_13), (bx, by)))`` (:145-148).  A pretrained V table is loaded under tuple
keys exactly like the reference does (:61-66) -- which means, there as here, that the step (which looks states up by their
pickled form) never reads it: it only counts towards ``get_v_table_size()``.
"""
import pickle
from collections import defaultdict

import numpy as np

from ..legacy import ACTOR_ONLY_DEFAULTS, LegacySim, key_to_state
from ..sim import MOORE, NEUMANN, seed_from_numpy_state

MAX_CAPACITY = 16380


class FloorFieldModelActorOnly:
    def __init__(self, map_array, sff_path, N, pretrained_v_path=None, params=None):
        self.params = dict(ACTOR_ONLY_DEFAULTS) if params is None else {**ACTOR_ONLY_DEFAULTS, **params}   # :24-42
        self.map_array = map_array.astype(np.uint8)
        sff_loaded = np.load(sff_path, mmap_mode="r")
        self.sff = np.where(np.isinf(sff_loaded), 0.0, sff_loaded).astype(np.float32)                      # :45-48
        self.N = N
        self.neighbors = self.get_neighbors()
        self._inert_v = {}
        if pretrained_v_path:                                                                              # :56-69
            with open(pretrained_v_path, "rb") as f:
                pretrained = pickle.load(f)
            for k, v in pretrained.items():
                real_key = pickle.loads(k)
                self._inert_v[tuple(tuple(int(x) for x in sub) for sub in real_key)] = v
            self.initial_v_size = len(self._inert_v)
            print(f"✓ 事前学習済みCriticを読み込みました: {self.initial_v_size}状態")
        else:
            self.initial_v_size = 0
            print("⚠ 事前学習済みCriticなしで開始します")
        self.alpha_v = self.params["alpha_v"]
        self.gamma = self.params["gamma"]
        self.alpha_h = self.params["alpha_h"]
        self.epsilon = self.params.get("epsilon", 0.0)
        seed = self.params.get("seed")
        self._seed = seed_from_numpy_state() if seed is None else int(seed)
        self._episode = 0
        self._cap = min(max(int((self.map_array == 0).sum()), int(N), 1), MAX_CAPACITY)
        self._sim = LegacySim(self.map_array, self.sff, 1, self._cap, model="actor_only", learn="exact", params=self.params,
                              seed=self._seed)
        self._host_pos = self._host_dff = self._v_cache = self._h_cache = None
        self._upload(self.initialize_agents(), keep_dff=False)

    # -- reference helpers ----------------------------------------------------------------------
    def initialize_agents(self):
        free_cells = np.argwhere(self.map_array == 0)                                                     # :80-85
        return free_cells[np.random.choice(len(free_cells), self.N, replace=False)]

    def get_neighbors(self):
        return list(NEUMANN) if self.params["neighborhood"] == "neumann" else list(MOORE)

    def _encode_state(self, x, y, occupancy):
        """ffm_actor_only.py:102-148 (host helper; cells beyond the map read 0 here, :120,137)."""
        h, w = occupancy.shape
        cells = [int(occupancy[x + a, y + b]) if (0 <= x + a < h and 0 <= y + b < w) else 0
                 for a in (-1, 0, 1) for b in (-1, 0, 1)]
        cells += [int(occupancy[x + a, y + b]) if (0 <= x + a < h and 0 <= y + b < w) else 0
                  for a, b in [(-2, 0), (2, 0), (0, -2), (0, 2)]]
        return pickle.dumps((tuple(cells), (x // 5, y // 5)))

    # -- state ----------------------------------------------------------------------------------
    def _upload(self, positions, keep_dff):
        positions = np.asarray(positions).reshape(-1, 2)
        if len(positions) > self._cap:
            raise ValueError(f"{len(positions)} pedestrians exceed the capacity {self._cap}")
        buf = np.full((1, self._cap, 2), -1, dtype=np.int32)
        buf[0, :len(positions)] = positions
        self._sim.set_episode_base(self._episode)
        self._sim.set_positions(buf, np.array([len(positions)], dtype=np.int32))
        if not keep_dff:
            self._sim.zero_dff()
        self._episode += 1
        self._host_pos = positions.astype(np.int64) if len(positions) else positions
        self._host_dff = None

    @property
    def positions(self):
        if self._host_pos is None:
            pos, n = self._sim.get_positions()
            self._host_pos = pos[0, :n[0]].astype(np.int64)
        return self._host_pos

    @positions.setter
    def positions(self, value):
        self._upload(value, keep_dff=True)

    @property
    def dff(self):
        if self._host_dff is None:
            self._host_dff = self._sim.get_dff()[0]
        return self._host_dff

    @dff.setter
    def dff(self, value):
        self._sim.set_dff(np.asarray(value, dtype=np.float32).reshape((1,) + self.map_array.shape))
        self._host_dff = None

    def _table(self, which):
        keys, rows = self._sim.get_table(which)
        nby = self._sim.nby
        if which == "V":
            return {pickle.dumps(key_to_state(k, nby)): float(r[0]) for k, r in zip(keys, rows)}
        return {pickle.dumps(key_to_state(k, nby)): [float(x) for x in r] for k, r in zip(keys, rows)}

    @property
    def V(self):
        if self._v_cache is None:
            self._v_cache = defaultdict(lambda: 0.0, {**self._inert_v, **self._table("V")})
        return self._v_cache

    @property
    def H(self):
        if self._h_cache is None:
            self._h_cache = defaultdict(lambda: [], self._table("H"))
        return self._h_cache

    # -- stepping -------------------------------------------------------------------------------
    def _invalidate(self):
        self._host_pos = self._host_dff = self._v_cache = self._h_cache = None

    def step(self):
        self._sim.rollout(1)
        self._invalidate()

    def update_dff(self):
        self._sim.update_dff()
        self._host_dff = None

    def reset(self):
        """ffm_actor_only.py:562-568: new placement, zero DFF, V / H kept."""
        self._upload(self.initialize_agents(), keep_dff=False)

    def run(self, save_prefix=None, save_interval=100, max_steps=None, return_trajectory=False):
        """ffm_actor_only.py:612-664."""
        W = self.map_array.shape[1]
        step, buffer, trajectory = 0, [], ([] if return_trajectory else None)
        record = bool(save_prefix) or return_trajectory
        t0 = int(self._sim.counters()[0][0])
        while self.positions.shape[0] > 0 and (max_steps is None or step < max_steps):
            chunk = int(save_interval) if save_prefix else 256
            if max_steps is not None:
                chunk = min(chunk, max_steps - step)
            out = self._sim.rollout(chunk, record=chunk if record else 0)
            done = int(self._sim.counters()[0][0]) - t0 - step
            self._invalidate()
            if record and done > 0:
                cells, cnt = out[0][0], out[1][0]
                rows = [np.stack(np.divmod(cells[t, :cnt[t]].astype(np.int64), W), axis=1) for t in range(done)]
                buffer += rows
                if return_trajectory:
                    trajectory += rows
            step += done
            if save_prefix and step % save_interval == 0 and buffer:
                np.savez_compressed(f"{save_prefix}_{step}.npz", positions=np.array(buffer, dtype=np.int32))
                buffer = []
            if done == 0:
                break
        if save_prefix and buffer:
            np.savez_compressed(f"{save_prefix}_final.npz", positions=np.array(buffer, dtype=np.int32))
        if return_trajectory:
            return step, np.array(trajectory, dtype=object)
        return step

    # -- tables ---------------------------------------------------------------------------------
    def get_v_table(self):
        return dict(self.V)                                                                                # :570-577

    def get_v_table_size(self):
        current = self.initial_v_size + self._sim.table_size("V")                                          # :579-588
        return (self.initial_v_size, current, current - self.initial_v_size)

    def get_h_table(self):
        return dict(self.H)                                                                                # :590-597

    def set_epsilon(self, epsilon):
        self.epsilon = float(np.clip(epsilon, 0.0, 1.0))                                                   # :599-606
        self._sim.set_epsilon(self.epsilon)

    def get_h_table_size(self):
        rows = self._sim.table_size("H")                                                                   # :608-616
        return (rows, rows * self._sim.A)
