"""Drop-in ``FloorFieldModelUnified`` with the interface of the reference's ``model/ffm_unified.py``.

    FloorFieldModelUnified(map_array, sff_path, N, learning_mode="critic_only",
                           pretrained_v_path=None, params=None)                     ffm_unified.py:27-35
    .reset(exit_pos=None, radius=None) .run(save_prefix, save_interval, max_steps, return_trajectory)
    .step() .update_dff() .initialize_agents(exit_pos, radius) .get_neighbors() ._encode_state(x, y, state_map)
    .get_v_table() .set_v_table(d) .get_v_table_size() .get_h_table() .get_h_table_size() .set_epsilon(e)
    attributes: params, learning_mode, map_array, sff, dff, N, positions, neighbors, V, H, epsilon,
                alpha_v, gamma, block_size, alpha_h, initial_v_size

The tables live on the device as dense arrays (csrc/ffm_unified_kernel.cuh); ``V`` / ``H`` /
``get_*_table()`` materialise the reference's dicts on demand (keys ``((rU,rD,rL,rR),(bx,by))``,
Python-float values / 5-element lists), so the drivers' pickling works unchanged
(run_unified_critic_training.py:290-299).  Learning uses the reference's sequential per-agent
updates (FFM_LEARN_EXACT).  Random numbers: Philox streams keyed (seed, episode, step, stream,
agent | cell); the episode counter advances on every reset().
"""
import pickle
from collections import defaultdict

import numpy as np

from ..sim import seed_from_numpy_state, MOORE, NEUMANN, UNIFIED_DEFAULTS, UnifiedSim

MAX_CAPACITY = 16380


class FloorFieldModelUnified:
    def __init__(self, map_array, sff_path, N, learning_mode="critic_only", pretrained_v_path=None, params=None):
        self.params = dict(UNIFIED_DEFAULTS) if params is None else {**UNIFIED_DEFAULTS, **params}   # ffm_unified.py:36-56
        valid_modes = ["critic_only", "actor_only", "both"]
        if learning_mode not in valid_modes:                                                         # :59-63
            raise ValueError(f"learning_mode must be one of {valid_modes}, got {learning_mode}")
        self.learning_mode = learning_mode
        self.map_array = map_array.astype(np.uint8)                                                  # :66
        sff_loaded = np.load(sff_path, mmap_mode="r")
        if learning_mode == "critic_only":
            self.sff = sff_loaded                                                                    # :70
        else:
            self.sff = np.where(np.isinf(sff_loaded), 0.0, sff_loaded).astype(np.float32)            # :72-76
        self.N = N
        self.neighbors = self.get_neighbors()
        self.alpha_v = self.params["alpha_v"]
        self.gamma = self.params["gamma"]
        self.block_size = self.params["block_size"]
        actor = learning_mode in ["actor_only", "both"]
        self.alpha_h = self.params["alpha_h"] if actor else None
        self.epsilon = self.params.get("epsilon", 0.0) if actor else None
        seed = self.params.get("seed")
        self._seed = seed_from_numpy_state() if seed is None else int(seed)
        self._episode = 0
        cap = min(max(int((self.map_array == 0).sum()), int(N), 1), MAX_CAPACITY)
        self._cap = cap
        self._sim = UnifiedSim(self.map_array, np.asarray(sff_loaded), 1, cap, mode=learning_mode, learn="exact",
                               params=self.params, seed=self._seed, episode_base=0)
        self.initial_v_size = 0
        if pretrained_v_path and actor:                                                              # :84-110
            with open(pretrained_v_path, "rb") as f:
                pickled = pickle.load(f)
            clean = {}
            for k, v in pickled.items():
                try:
                    k = pickle.loads(k)
                except TypeError:
                    pass
                clean[(tuple(int(r) for r in k[0]), (int(k[1][0]), int(k[1][1])))] = v
            self._sim.load_v_dict(clean)
            self.initial_v_size = len(clean)
            print(f"✓ 事前学習済みCriticを読み込みました: {self.initial_v_size}状態")
        elif learning_mode == "actor_only":
            print("⚠ 警告: actor_onlyモードですが事前学習済みCriticが指定されていません")
        self._host_pos = self._host_dff = self._v_cache = self._h_cache = None
        self._upload(self.initialize_agents(), keep_dff=False)                                       # :80

    # -- reference helpers ----------------------------------------------------------------------
    def initialize_agents(self, exit_pos=None, radius=None):
        """ffm_unified.py:131-171 (same global-generator draws as the reference)."""
        free_cells = np.argwhere(self.map_array == 0)
        if exit_pos is None or radius is None:
            return free_cells[np.random.choice(len(free_cells), self.N, replace=False)]
        exit_x, exit_y = exit_pos
        radius_cells = free_cells[np.abs(free_cells[:, 0] - exit_x) + np.abs(free_cells[:, 1] - exit_y) <= radius]
        actual_N = min(self.N, len(radius_cells))
        if actual_N == 0:
            return np.empty((0, 2), dtype=np.int32)
        return radius_cells[np.random.choice(len(radius_cells), actual_N, replace=False)]

    def get_neighbors(self):
        return list(NEUMANN) if self.params["neighborhood"] == "neumann" else list(MOORE)

    def _encode_state(self, x, y, state_map):
        """ffm_unified.py:188-269 (host restatement for callers that poke it; the kernel has its own)."""
        height, width = state_map.shape
        ranks = []
        for dx, dy in [(-1, 0), (1, 0), (0, -1), (0, 1)]:
            a, b = x + dx, y + dy
            if not (0 <= a < height and 0 <= b < width) or state_map[a, b] in (1, 2):
                ranks.append(0)
                continue
            diag = [(a, b - 1), (a, b + 1)] if dx != 0 else [(a - 1, b), (a + 1, b)]
            if any(0 <= p < height and 0 <= q < width and state_map[p, q] == 1 for p, q in diag):
                ranks.append(1)
                continue
            a2, b2 = x + 2 * dx, y + 2 * dy
            ranks.append(2 if (not (0 <= a2 < height and 0 <= b2 < width) or state_map[a2, b2] in (1, 2)) else 3)
        return (tuple(int(r) for r in ranks), (int(x // self.block_size), int(y // self.block_size)))

    # -- state ----------------------------------------------------------------------------------
    def _upload(self, positions, keep_dff):
        positions = np.asarray(positions).reshape(-1, 2)
        if len(positions) > self._cap:
            raise ValueError(f"{len(positions)} pedestrians exceed the capacity {self._cap}")
        dff = self._sim.get_dff() if keep_dff else None
        buf = np.full((1, self._cap, 2), -1, dtype=np.int32)
        buf[0, :len(positions)] = positions
        self._sim.set_episode_base(self._episode)
        self._sim.set_positions(buf, np.array([len(positions)], dtype=np.int32))
        self._sim.get_positions()
        if dff is not None:
            self._sim.set_dff(dff)
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

    @property
    def V(self):
        """The reference's ``self.V`` materialised from the device table; cached until the tables can have changed (a step,
        a run, a table upload), so that drivers reading ``len(model.V)`` every episode do not rebuild it per access."""
        if self._v_cache is None:
            self._v_cache = defaultdict(lambda: 0.0, self._sim.v_dict())
        return self._v_cache

    @property
    def H(self):
        if self.learning_mode not in ["actor_only", "both"]:
            return None
        if self._h_cache is None:
            self._h_cache = defaultdict(lambda: [], self._sim.h_dict())
        return self._h_cache

    # -- stepping -------------------------------------------------------------------------------
    def step(self):
        self._sim.rollout(1)
        self._host_pos = self._host_dff = self._v_cache = self._h_cache = None

    def update_dff(self):
        """ffm_unified.py:779-798 as a stand-alone call: the kernels' stencil run once on the device (inside step() the step kernel
        does it)."""
        self._sim.update_dff()
        self._host_dff = None

    def reset(self, exit_pos=None, radius=None):
        """ffm_unified.py:800-812: new placement, zero DFF, tables kept."""
        self._upload(self.initialize_agents(exit_pos=exit_pos, radius=radius), keep_dff=False)

    def run(self, save_prefix=None, save_interval=100, max_steps=None, return_trajectory=False):
        """ffm_unified.py:882-931."""
        import torch
        W = self.map_array.shape[1]
        step, trajectory, buffer = 0, ([] if return_trajectory else None), []
        record = bool(save_prefix) or return_trajectory
        while self.positions.shape[0] > 0 and (max_steps is None or step < max_steps):
            chunk = int(save_interval) if save_prefix else 256
            if max_steps is not None:
                chunk = min(chunk, max_steps - step)
            out = self._sim.rollout(chunk, record=chunk if record else 0)
            done = int(self._sim.counters()[0][0]) - step
            self._host_pos = self._host_dff = self._v_cache = self._h_cache = None
            if record and done > 0:
                torch.cuda.synchronize()
                cells, cnt = out[0].cpu().numpy()[0], out[1].cpu().numpy()[0]
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
        return self._sim.v_dict()                                   # :814-821

    def set_v_table(self, v_table):
        self._sim.load_v_dict(dict(v_table))                        # :823-830
        self._v_cache = None

    def get_v_table_size(self):
        size = int(self._sim.get_tables()[1].sum())
        if self.learning_mode == "actor_only":                      # :840-845
            return (self.initial_v_size, size, size - self.initial_v_size)
        return size

    def get_h_table(self):
        return self._sim.h_dict() if self.learning_mode in ["actor_only", "both"] else None   # :847-857

    def get_h_table_size(self):
        if self.learning_mode in ["actor_only", "both"]:            # :869-880
            rows = int(self._sim.get_tables()[3].sum())
            return (rows, rows * self._sim.A)
        return None

    def set_epsilon(self, epsilon):
        if self.learning_mode in ["actor_only", "both"]:            # :859-867
            self.epsilon = float(np.clip(epsilon, 0.0, 1.0))
            self._sim.set_epsilon(self.epsilon)
