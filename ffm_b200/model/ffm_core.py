"""Drop-in ``FloorFieldModel`` with the interface of the reference's ``model/ffm_core.py``.

    FloorFieldModel(map_array, sff_path, N, params=None)      ffm_core.py:7
    .step() .update_dff() .run(save_prefix=None, save_interval=100)
    .initialize_agents() .get_neighbors()
    attributes: params, map_array, sff, dff, N, positions, neighbors

Semantics are the reference's, step for step (kernel: csrc/ffm_core_kernel.cuh); what differs is
where the random numbers come from: the reference consumes the process-global NumPy / ``random``
generators in agent order, this class draws from counter-based Philox streams keyed
(seed, episode, step, stream, agent | cell).  The seed is taken from ``params["seed"]`` or, like the
reference's placement draw, from the global NumPy generator at construction -- so
``np.random.seed(s)`` before constructing makes a run reproducible.
"""
import numpy as np

from ..sim import seed_from_numpy_state, BatchSim, CORE_DEFAULTS, MOORE, NEUMANN


class FloorFieldModel:
    def __init__(self, map_array, sff_path, N, params=None):
        self.params = dict(CORE_DEFAULTS) if params is None else {**CORE_DEFAULTS, **params}   # ffm_core.py:8-15
        self.map_array = map_array.astype(np.uint8)                                            # ffm_core.py:16
        self.sff = np.load(sff_path, mmap_mode="r")                                            # ffm_core.py:17
        self.N = N
        self.neighbors = self.get_neighbors()
        self._episode = 0
        seed = self.params.get("seed")
        self._seed = seed_from_numpy_state() if seed is None else int(seed)
        positions = self.initialize_agents()                                                   # ffm_core.py:20
        self._cap = max(int(N), 1)
        self._sim = BatchSim(self.map_array, np.asarray(self.sff), 1, self._cap, self.params,
                             seed=self._seed, episode_base=0, track_dff=True)
        self._host_pos = None
        self._host_dff = None
        self.positions = positions

    # -- reference helpers ----------------------------------------------------------------------
    def initialize_agents(self):
        """ffm_core.py:23-26 (same global-generator draw, so the same seed gives the same placement)."""
        free_cells = np.argwhere(self.map_array == 0)
        selected = free_cells[np.random.choice(len(free_cells), self.N, replace=False)]
        return selected

    def get_neighbors(self):
        """ffm_core.py:28-34."""
        return list(NEUMANN) if self.params["neighborhood"] == "neumann" else list(MOORE)

    # -- state properties (device is the source of truth between steps) ---------------------------
    @property
    def positions(self):
        if self._host_pos is None:
            pos, n = self._sim.get_positions()
            self._host_pos = pos[0, :n[0]].astype(np.int64)
        return self._host_pos

    @positions.setter
    def positions(self, value):
        """Assignment (run_trained_ffm.py:235): uploads, keeps the DFF and the step counter."""
        value = np.asarray(value).reshape(-1, 2)
        if len(value) > self._cap:
            raise ValueError(f"{len(value)} pedestrians exceed the capacity {self._cap} fixed at construction")
        dff = self._sim.get_dff() if getattr(self, "_started", False) else None
        buf = np.full((1, self._cap, 2), -1, dtype=np.int32)
        buf[0, :len(value)] = value
        self._sim.set_positions(buf, np.array([len(value)], dtype=np.int32))
        self._sim.get_positions()          # surfaces validation errors (ValueError) now, not later
        if dff is not None:
            self._sim.set_dff(dff)
        self._started = True
        self._host_pos = value.astype(np.int64)
        self._host_dff = None

    @property
    def dff(self):
        if self._host_dff is None:
            self._host_dff = self._sim.get_dff()[0]
        return self._host_dff

    @dff.setter
    def dff(self, value):
        """Assignment (run_trained_ffm.py:236)."""
        value = np.asarray(value, dtype=np.float32).reshape(self.map_array.shape)
        self._sim.set_dff(value[None])
        self._host_dff = None

    # -- stepping -------------------------------------------------------------------------------
    def step(self):
        """ffm_core.py:36-104: one CA step including update_dff()."""
        self._sim.rollout(1)
        self._host_pos = None
        self._host_dff = None

    def update_dff(self):
        """ffm_core.py:106-117 as a stand-alone call: the kernels' stencil run once on the device (inside step() the step kernel
        does it)."""
        self._sim.update_dff()
        self._host_dff = None

    def run(self, save_prefix=None, save_interval=100):
        """ffm_core.py:119-133: step until everybody has left; optional .npz dumps of the buffered
        per-step positions every ``save_interval`` steps (same file names and dtype)."""
        import torch
        step = 0
        while self.positions.shape[0] > 0:
            chunk = int(save_interval) if save_prefix else 256
            if save_prefix:
                cells, cnt = self._sim.rollout(chunk, record=chunk)
                torch.cuda.synchronize()
                cells, cnt = cells.cpu().numpy()[0], cnt.cpu().numpy()[0]
            else:
                self._sim.rollout(chunk)
            done = int(self._sim.counters()[0][0]) - step
            self._host_pos = None
            self._host_dff = None
            if save_prefix and done > 0:
                W = self.map_array.shape[1]
                buffer = [np.stack(np.divmod(cells[t, :cnt[t]].astype(np.int64), W), axis=1) for t in range(done)]
                tag = f"{step + done}" if done == chunk and (step + done) % save_interval == 0 else "final"
                # like the reference (ffm_core.py:129,133) this needs equally long rows; ragged -> ValueError
                np.savez_compressed(f"{save_prefix}_{tag}.npz", positions=np.array(buffer, dtype=np.int32))
            step += done
            if done == 0:
                break
