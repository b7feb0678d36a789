"""Drop-in ``FloorFieldModel`` with the interface of the reference's ``model/ffm_trained_core.py``
(inference with a trained actor table; run_trained_ffm.py:199-243 drives it).

    FloorFieldModel(map_array, sff_path, N, h_table_path, params=None)       ffm_trained_core.py:20
    .step() .run(save_prefix=None, save_interval=100, max_steps=None) .update_dff()
    .initialize_agents() .get_neighbors();  attributes positions / dff are assignable
"""
import pickle

import numpy as np

from ..sim import seed_from_numpy_state, MOORE, NEUMANN, TRAINED_DEFAULTS, UnifiedSim
from .ffm_unified import MAX_CAPACITY


class FloorFieldModel:
    def __init__(self, map_array, sff_path, N, h_table_path, params=None):
        self.params = dict(TRAINED_DEFAULTS) if params is None else {**TRAINED_DEFAULTS, **params}   # :29-37
        self.map_array = map_array.astype(np.uint8)
        sff_loaded = np.load(sff_path, mmap_mode="r")
        self.sff = np.where(np.isinf(sff_loaded), 0.0, sff_loaded).astype(np.float32)                # :41-43
        self.N = N
        self.neighbors = self.get_neighbors()
        self.block_size = self.params["block_size"]
        with open(h_table_path, "rb") as f:                                                          # :52-68
            pickled = pickle.load(f)
        self.H = {}
        for k, v in pickled.items():
            if isinstance(k, bytes):
                k = pickle.loads(k)
            self.H[(tuple(int(r) for r in k[0]), (int(k[1][0]), int(k[1][1])))] = v
        print(f"✓ 学習済みHテーブルを読み込みました: {len(self.H)}状態")
        seed = self.params.get("seed")
        self._seed = seed_from_numpy_state() if seed is None else int(seed)
        self._episode = 0
        self._cap = min(max(int((self.map_array == 0).sum()), int(N), 1), MAX_CAPACITY)
        self._sim = UnifiedSim(self.map_array, np.asarray(sff_loaded), 1, self._cap, mode="trained", learn="none",
                               params=self.params, seed=self._seed)
        self._sim.load_h_dict(self.H)
        self._host_pos = self._host_dff = None
        self.positions = self.initialize_agents()

    def initialize_agents(self):
        free_cells = np.argwhere(self.map_array == 0)                                                # :72-76
        return free_cells[np.random.choice(len(free_cells), self.N, replace=False)]

    def get_neighbors(self):
        return list(NEUMANN) if self.params["neighborhood"] == "neumann" else list(MOORE)

    @property
    def positions(self):
        if self._host_pos is None:
            pos, n = self._sim.get_positions()
            self._host_pos = pos[0, :n[0]].astype(np.int64)
        return self._host_pos

    @positions.setter
    def positions(self, value):
        """run_trained_ffm.py:235 -- starts a new episode (the DFF is kept until assigned)."""
        value = np.asarray(value).reshape(-1, 2)
        if len(value) > self._cap:
            raise ValueError(f"{len(value)} pedestrians exceed the capacity {self._cap}")
        dff = self._sim.get_dff() if self._episode > 0 else None
        buf = np.full((1, self._cap, 2), -1, dtype=np.int32)
        buf[0, :len(value)] = value
        self._sim.set_episode_base(self._episode)
        self._sim.set_positions(buf, np.array([len(value)], dtype=np.int32))
        self._sim.get_positions()
        if dff is not None:
            self._sim.set_dff(dff)
        self._episode += 1
        self._host_pos, self._host_dff = value.astype(np.int64), None

    @property
    def dff(self):
        if self._host_dff is None:
            self._host_dff = self._sim.get_dff()[0]
        return self._host_dff

    @dff.setter
    def dff(self, value):
        self._sim.set_dff(np.asarray(value, dtype=np.float32).reshape((1,) + self.map_array.shape))
        self._host_dff = None

    def step(self):
        self._sim.rollout(1)
        self._host_pos = self._host_dff = None

    def update_dff(self):
        """ffm_trained_core.py:333-353 as a stand-alone call: the kernels' stencil run once on the device (inside step() the step kernel
        does it)."""
        self._sim.update_dff()
        self._host_dff = None

    def run(self, save_prefix=None, save_interval=100, max_steps=None):
        """ffm_trained_core.py:355-391."""
        import torch
        W = self.map_array.shape[1]
        step, buffer = 0, []
        while self.positions.shape[0] > 0 and (max_steps is None or step < max_steps):
            chunk = int(save_interval) if save_prefix else 256
            if max_steps is not None:
                chunk = min(chunk, max_steps - step)
            out = self._sim.rollout(chunk, record=chunk if save_prefix else 0)
            done = int(self._sim.counters()[0][0]) - step
            self._host_pos = self._host_dff = None
            if save_prefix and done > 0:
                torch.cuda.synchronize()
                cells, cnt = out[0].cpu().numpy()[0], out[1].cpu().numpy()[0]
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
