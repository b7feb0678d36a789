"""NumPy restatement of the reference CA step (test infrastructure -- see oracle/__init__.py).

``CoreOracle`` follows ``/root/reference/model/ffm_core.py`` function by function but replaces
the per-agent ``np.delete`` + structured ``np.isin`` occupancy test (ffm_core.py:48-60, the
O(n^2 log n) part) by an occupancy grid, and takes its uniforms from a draw source
(oracle/inject.py protocol) instead of the process-global generators.  Every floating-point
expression is the reference's own NumPy expression (same dtypes, same ufuncs), so on one machine
the restatement is bit-identical to the reference; ``tests/test_oracle_vs_golden.py`` pins that
against fixtures produced from the unmodified reference.
"""
import numpy as np

from .inject import choice_cdf

NEUMANN = [(-1, 0), (1, 0), (0, -1), (0, 1)]                       # ffm_core.py:30
MOORE = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]   # ffm_core.py:32-34

CORE_DEFAULTS = {"k_S": 3, "k_D": 1, "diffuse": 0.2, "decay": 0.2, "neighborhood": "moore"}  # ffm_core.py:8-14


def neighbors_of(params):
    """ffm_core.py:28-34."""
    return list(NEUMANN) if params["neighborhood"] == "neumann" else list(MOORE)


def update_dff(dff, params, neighbors):
    """ffm_core.py:106-117 (== ffm_unified.py:779-798, ffm_trained_core.py:333-353)."""
    diffuse = params["diffuse"]
    decay = params["decay"]
    new_dff = (1 - decay) * (1 - diffuse) * dff
    padded = np.pad(new_dff, 1, mode="constant")
    for dx, dy in neighbors:
        new_dff += decay * (1 - diffuse) / len(neighbors) * padded[1 + dx:new_dff.shape[0] + 1 + dx,
                                                                  1 + dy:new_dff.shape[1] + 1 + dy]
    threshold = 1e-4
    new_dff[new_dff < threshold] = 0
    return new_dff


class CoreOracle:
    """State + step of ``FloorFieldModel`` (ffm_core.py:6-133) with keyed draws."""

    def __init__(self, map_array, sff, positions, params=None, source=None):
        self.params = dict(CORE_DEFAULTS) if params is None else {**CORE_DEFAULTS, **params}
        self.map_array = np.asarray(map_array).astype(np.uint8)          # ffm_core.py:16
        self.sff = np.asarray(sff)                                       # file dtype kept (ffm_core.py:17)
        self.dff = np.zeros_like(self.map_array, dtype=np.float32)       # ffm_core.py:18
        self.positions = np.array(positions, dtype=np.int64).reshape(-1, 2)
        self.neighbors = neighbors_of(self.params)
        self.source = source
        self.t = 0
        self.min_margin = np.inf
        self.probs_log = None

    def step(self):
        src, t = self.source, self.t
        H, W = self.map_array.shape
        pos = self.positions
        n = pos.shape[0]
        occ = np.zeros((H, W), dtype=bool)
        occ[pos[:, 0], pos[:, 1]] = True
        move_requests = {}
        next_positions = np.copy(pos)
        offsets = np.array(self.neighbors)
        k_S, k_D = self.params["k_S"], self.params["k_D"]

        for idx in range(n):                                             # ffm_core.py:40
            cur = pos[idx]
            nb = cur + offsets
            cell = self.map_array[nb[:, 0], nb[:, 1]]
            nb = nb[(cell == 0) | (cell == 3)]                           # ffm_core.py:52-54
            if nb.shape[0] > 0:
                nb = nb[~occ[nb[:, 0], nb[:, 1]]]                        # ffm_core.py:57-60
            if nb.shape[0] == 0:
                continue                                                 # no request, no draw (ffm_core.py:63)
            nb = np.vstack([nb, cur])                                    # "stay" appended last (ffm_core.py:64)
            exit_mask = self.map_array[nb[:, 0], nb[:, 1]] == 3
            if np.any(exit_mask):                                        # forced exit (ffm_core.py:66-72)
                chosen = tuple(int(v) for v in nb[exit_mask][0])
                move_requests.setdefault(chosen, []).append(idx)
                continue
            sff_vals = self.sff[nb[:, 0], nb[:, 1]]
            dff_vals = self.dff[nb[:, 0], nb[:, 1]]
            score = -k_S * sff_vals + k_D * dff_vals                     # ffm_core.py:77
            probs = np.exp(score - np.max(score))                        # ffm_core.py:78-80
            probs_sum = probs.sum()
            if np.isfinite(probs_sum) and probs_sum != 0:                # ffm_core.py:82
                probs /= probs_sum
                cdf = choice_cdf(probs)                                  # ffm_core.py:84
                u = src.move(t, idx, cdf)
                self.min_margin = min(self.min_margin, float(np.min(np.abs(cdf - u))))
                if self.probs_log is not None:
                    self.probs_log.append((t, idx, probs.copy()))
                chosen = tuple(int(v) for v in nb[int(cdf.searchsorted(u, side="right"))])
                move_requests.setdefault(chosen, []).append(idx)

        for target, agents in move_requests.items():                     # ffm_core.py:90-98
            cellid = target[0] * W + target[1]
            if len(agents) == 1:
                a = agents[0]
                next_positions[a] = target
                self.dff[pos[a][0], pos[a][1]] += 1
            elif src.coin(t, cellid) < 0.5:
                k = len(agents)
                a = agents[int(src.winner(t, cellid, k) * k)]
                next_positions[a] = target
                self.dff[pos[a][0], pos[a][1]] += 1

        keep = self.map_array[next_positions[:, 0], next_positions[:, 1]] != 3   # ffm_core.py:101-102
        self.positions = next_positions[keep]
        self.dff = update_dff(self.dff, self.params, self.neighbors)     # ffm_core.py:104
        self.t += 1

    def run(self, max_steps=None, keep_dff=True):
        """ffm_core.py:119-133 without the file dumps; returns the same dict as inject.run_reference."""
        traj, dffs = [], []
        while self.positions.shape[0] > 0 and (max_steps is None or self.t < max_steps):
            self.step()
            traj.append(self.positions.copy())
            if keep_dff:
                dffs.append(self.dff.copy())
        return dict(steps=self.t, traj=traj, dff=dffs, min_margin=self.min_margin)
