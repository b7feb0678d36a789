"""Drop-in ``FloorFieldModel`` with the interface of the reference's ``model/ffm_learning_core.py``
(target-centric Monte-Carlo Q-learning; ``main_learning.py:60-106`` drives it).

    FloorFieldModel(map_array, sff_path, N, params=None)      ffm_learning_core.py:39-91
    .reset() .step(beta) .finalize_timeouts() .save_Q(path)
    attributes: params, map_array, sff, dff, N, positions, neighbors, Q (assignable dict), alpha, gamma,
                action_size, max_steps

``Q`` is materialised from / loaded into the dense device table on access (keys
``(combined3x3.tobytes(), (bx, by))``, float32[5] rows), so ``model.Q = shared_Q`` ... ``return model.Q``
(main_learning.py:81,106) works.  ``alpha`` / ``gamma`` may be assigned after construction (:79-80): the device
handle is created on first use.  Not provided: the per-agent ``paths`` / ``prev_direction`` internals that
``run_coverage_pretrain_and_training.py`` pokes.
"""
import pickle

import numpy as np

from ..sim import seed_from_numpy_state, MCQ_DEFAULTS, McqSim
from .ffm_unified import MAX_CAPACITY


class FloorFieldModel:
    FROM_UP, FROM_DOWN, FROM_LEFT, FROM_RIGHT, FROM_SELF = range(5)

    def __init__(self, map_array, sff_path, N, params=None):
        self.params = dict(MCQ_DEFAULTS) if params is None else {**MCQ_DEFAULTS, **params}
        self.map_array = map_array.astype(np.uint8)
        self.sff = np.load(sff_path, mmap_mode="r")
        self.N = int(N)
        self.neighbors = [(-1, 0), (1, 0), (0, -1), (0, 1)]
        self.alpha, self.gamma = 0.1, 0.99
        self.action_size = 5
        self.max_steps = int(self.params["max_steps"])
        seed = self.params.get("seed")
        self._seed = seed_from_numpy_state() if seed is None else int(seed)
        self._episode = int(self.params.get("episode", 0))
        self._sim = None
        self._made_with = None
        self._pending_q = None
        self._host_pos = self._host_dff = None
        self._cap = min(max(int((self.map_array == 0).sum()), self.N, 1), MAX_CAPACITY)
        self._start(self._initialize_agents())

    def _initialize_agents(self):
        free_cells = np.argwhere(self.map_array == 0)                     # :95-99
        return free_cells[np.random.choice(len(free_cells), self.N, replace=False)].astype(np.int16)

    # -- device handle -----------------------------------------------------------------------------
    def _ensure(self):
        if self._sim is None or self._made_with != (self.alpha, self.gamma):
            q = self._sim.q_dict() if self._sim is not None else self._pending_q
            state = (self.positions.copy(), self.dff.copy()) if self._sim is not None else None
            if self._sim is not None:
                self._sim.close()
            self._sim = McqSim(self.map_array, np.asarray(self.sff), 1, self._cap, learn="exact", params=self.params,
                               seed=self._seed, alpha=self.alpha, gamma=self.gamma)
            self._made_with = (self.alpha, self.gamma)
            if q:
                self._sim.load_q_dict(q)
            self._pending_q = None
            self._upload(self._start_pos if state is None else state[0])
            if state is not None:
                self._sim.set_dff(state[1][None])
        return self._sim

    def _upload(self, positions):
        buf = np.full((1, self._cap, 2), -1, dtype=np.int32)
        buf[0, :len(positions)] = positions
        self._sim.set_episode_base(self._episode)
        self._sim.set_positions(buf, np.array([len(positions)], dtype=np.int32))
        self._sim.get_positions()

    def _start(self, positions):
        self._start_pos = np.asarray(positions).reshape(-1, 2)
        self._host_pos = self._start_pos.astype(np.int16)
        self._host_dff = np.zeros_like(self.map_array, dtype=np.float32)
        if self._sim is not None:
            self._upload(self._start_pos)

    # -- reference API -----------------------------------------------------------------------------
    def reset(self):
        self._episode += 1                                                # new draws for the new episode
        self._start(self._initialize_agents())                            # :101-106

    @property
    def positions(self):
        if self._host_pos is None:
            pos, n = self._sim.get_positions()
            self._host_pos = pos[0, :n[0]].astype(np.int16)
        return self._host_pos

    @property
    def dff(self):
        if self._host_dff is None:
            self._host_dff = self._sim.get_dff()[0]
        return self._host_dff

    @property
    def Q(self):
        return self._pending_q if self._sim is None else self._sim.q_dict()

    @Q.setter
    def Q(self, value):
        if self._sim is None:
            self._pending_q = dict(value)
        else:
            self._sim.load_q_dict(dict(value))

    def step(self, beta):
        sim = self._ensure()
        sim.set_beta(beta)
        sim.rollout(1)
        self._host_pos = self._host_dff = None

    def finalize_timeouts(self):
        if self.positions.shape[0] == 0:
            return
        self._ensure().finalize_timeouts()
        self._host_pos = self._host_dff = None

    def save_Q(self, filepath):
        with open(filepath, "wb") as f:
            pickle.dump(self.Q or {}, f)                                   # :365-367
