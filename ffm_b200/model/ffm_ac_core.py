"""Drop-in ``FloorFieldModel`` with the interface of the reference's legacy ``model/ffm_ac_core.py``.

    FloorFieldModel(map_array, sff_path, N, params=None)                                ffm_ac_core.py:9-38
    .step() .run(save_prefix, save_interval, max_steps) .reset() .update_dff()
    .initialize_agents() .get_neighbors() ._encode_state(x, y, state_map)
    .get_v_table() .set_v_table(d) .get_v_table_size()
    attributes: params, map_array, sff, dff, N, positions, neighbors, V, alpha_v, gamma, block_size

The V table lives on the device as a hash table (csrc/ffm_legacy.cu); ``V`` / ``get_v_table()`` materialise
the reference's dict with its own keys, ``pickle.dumps((tuple(state_13), (bx, by)))`` (:106-109), so
run_critic_training.py's table pickling works unchanged.  Learning is the reference's sequential per-agent
TD(0) update.  Random numbers: Philox streams keyed (seed, episode, step, stream, agent | cell); the episode
counter advances on every reset().
"""
import pickle
from collections import defaultdict

import numpy as np

from ..legacy import AC_DEFAULTS, LegacySim, key_to_state, state_to_key
from ..sim import MOORE, NEUMANN, seed_from_numpy_state

MAX_CAPACITY = 16380


class FloorFieldModel:
    def __init__(self, map_array, sff_path, N, params=None):
        self.params = dict(AC_DEFAULTS) if params is None else {**AC_DEFAULTS, **params}     # ffm_ac_core.py:10-26
        self.map_array = map_array.astype(np.uint8)                                         # :27
        self.sff = np.load(sff_path, mmap_mode="r")                                         # :28
        self.N = N
        self.neighbors = self.get_neighbors()
        self.alpha_v = self.params["alpha_v"]
        self.gamma = self.params["gamma"]
        self.block_size = self.params["block_size"]
        seed = self.params.get("seed")
        self._seed = seed_from_numpy_state() if seed is None else int(seed)
        self._episode = 0
        self._cap = min(max(int((self.map_array == 0).sum()), int(N), 1), MAX_CAPACITY)
        self._sim = LegacySim(self.map_array, np.asarray(self.sff), 1, self._cap, model="ac", learn="exact",
                              params=self.params, seed=self._seed)
        self._v_default = 0.0                                                               # :34
        self._host_pos = self._host_dff = self._v_cache = None
        self._upload(self.initialize_agents(), keep_dff=False)                              # :31

    # -- reference helpers ----------------------------------------------------------------------
    def initialize_agents(self):
        free_cells = np.argwhere(self.map_array == 0)                                       # :40-45
        return free_cells[np.random.choice(len(free_cells), self.N, replace=False)]

    def get_neighbors(self):
        return list(NEUMANN) if self.params["neighborhood"] == "neumann" else list(MOORE)   # :47-60

    def _encode_state(self, x, y, state_map):
        """ffm_ac_core.py:62-109 (host helper for callers that poke it; the kernel has its own encoder)."""
        h, w = state_map.shape
        cells = [int(state_map[x + a, y + b]) if (0 <= x + a < h and 0 <= y + b < w) else 2
                 for a in (-1, 0, 1) for b in (-1, 0, 1)]
        cells += [int(state_map[x + a, y + b]) if (0 <= x + a < h and 0 <= y + b < w) else 2
                  for a, b in [(-2, 0), (2, 0), (0, -2), (0, 2)]]
        return pickle.dumps((tuple(cells), (x // self.block_size, y // self.block_size)))

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

    def _v_dict(self):
        keys, rows = self._sim.get_table("V")
        nby = self._sim.nby
        return {pickle.dumps(key_to_state(k, nby)): float(v) for k, v in zip(keys, rows[:, 0])}

    @property
    def V(self):
        """The reference's ``self.V`` materialised from the device table; cached until the next step / table upload."""
        if self._v_cache is None:
            d = self._v_default
            self._v_cache = defaultdict(lambda: d, self._v_dict())
        return self._v_cache

    # -- stepping -------------------------------------------------------------------------------
    def step(self):
        self._sim.rollout(1)
        self._host_pos = self._host_dff = self._v_cache = None

    def update_dff(self):
        self._sim.update_dff()                                                              # :298-318
        self._host_dff = None

    def reset(self):
        """ffm_ac_core.py:320-326: new placement, zero DFF, V kept."""
        self._upload(self.initialize_agents(), keep_dff=False)

    def run(self, save_prefix=None, save_interval=100, max_steps=None):
        """ffm_ac_core.py:362-390."""
        W = self.map_array.shape[1]
        step, buffer = 0, []
        t0 = int(self._sim.counters()[0][0])
        while self.positions.shape[0] > 0 and (max_steps is None or step < max_steps):
            chunk = int(save_interval) if save_prefix else 256
            if max_steps is not None:
                chunk = min(chunk, max_steps - step)
            out = self._sim.rollout(chunk, record=chunk if save_prefix else 0)
            done = int(self._sim.counters()[0][0]) - t0 - step
            self._host_pos = self._host_dff = self._v_cache = None
            if save_prefix and done > 0:
                cells, cnt = out[0][0], out[1][0]
                buffer += [np.stack(np.divmod(cells[t, :cnt[t]].astype(np.int64), W), axis=1) for t in range(done)]
            step += done
            if save_prefix and step % save_interval == 0 and buffer:
                np.savez_compressed(f"{save_prefix}_{step}.npz", positions=np.array(buffer, dtype=np.int32))
                buffer = []
            if done == 0:
                break
        if save_prefix and buffer:
            np.savez_compressed(f"{save_prefix}_final.npz", positions=np.array(buffer, dtype=np.int32))
        return step

    # -- tables ---------------------------------------------------------------------------------
    def get_v_table(self):
        return self._v_dict()                                                               # :328-335

    def set_v_table(self, v_table):
        """ffm_ac_core.py:337-345: afterwards an unseen state reads as -1.0 (the lambda of :345)."""
        nby = self._sim.nby
        keys, vals = [], []
        for k, v in v_table.items():
            st = pickle.loads(k) if isinstance(k, (bytes, bytearray)) else k
            keys.append(state_to_key(st, nby))
            vals.append(float(v))
        self._v_default = -1.0
        self._sim.set_table(np.array(keys, np.uint64), np.array(vals, np.float64), "V", default=-1.0)
        self._v_cache = None

    def get_v_table_size(self):
        return self._sim.table_size("V")                                                    # :347-354
