"""NumPy restatement of the unified critic/actor model and the trained-actor inference model
(test infrastructure -- see oracle/__init__.py).

Follows ``/root/reference/model/ffm_unified.py`` (``FloorFieldModelUnified``: ``_encode_state`` :188-269,
``step`` :271-606, ``_update_critic`` :608-670, ``_get_td_errors`` :672-723, ``_update_actor`` :725-777,
``update_dff`` :779-798) and ``/root/reference/model/ffm_trained_core.py`` (``step`` :159-331), with

  * dense tables instead of dicts: state id = (bx * nby + by) * 256 + rU*64 + rD*16 + rL*4 + rR,
    ``V[S]`` float64 + ``v_seen[S]`` (a defaultdict read inserts the key, :658,661),
    ``H[S, A]`` float64 + ``h_seen[S]`` (rows are inserted as zeros, :405-411, :769-773);
  * an occupancy grid instead of per-agent set building (:301-302);
  * keyed draws from a draw source (oracle/inject.py protocol) instead of the global generators.

Every floating-point expression keeps the reference's dtype and operation order; the table
arithmetic is Python-float (float64) scalar arithmetic exactly as in the reference.
``tests/test_unified_oracle.py`` pins this file to fixtures produced by the unmodified reference.
"""
import numpy as np

from .ffm_numpy import MOORE, NEUMANN, update_dff
from .inject import choice_cdf

UNIFIED_DEFAULTS = {                      # ffm_unified.py:36-53
    "k_S": 10, "k_D": 1, "k_A": 10, "diffuse": 0.2, "decay": 0.2, "neighborhood": "neumann",
    "alpha_v": 0.1, "gamma": 0.95, "exit_reward": 100.0, "step_penalty": 0.0, "collision_penalty": -1.0,
    "block_size": 5, "alpha_h": 0.1, "epsilon": 0.0,
}
TRAINED_DEFAULTS = {"k_D": 1, "k_A": 10, "diffuse": 0.2, "decay": 0.2, "neighborhood": "neumann", "block_size": 5}  # ffm_trained_core.py:29-36

DIRS = [(-1, 0), (1, 0), (0, -1), (0, 1)]   # up, down, left, right (ffm_unified.py:209)


def encode_ranks(x, y, state_map):
    """ffm_unified.py:205-256: the four direction ranks of cell (x, y)."""
    h, w = state_map.shape
    ranks = []
    for dx, dy in DIRS:
        n1x, n1y = x + dx, y + dy
        if not (0 <= n1x < h and 0 <= n1y < w):
            ranks.append(0)
            continue
        v1 = state_map[n1x, n1y]
        if v1 == 2 or v1 == 1:
            ranks.append(0)
            continue
        diag = [(n1x, n1y - 1), (n1x, n1y + 1)] if dx != 0 else [(n1x - 1, n1y), (n1x + 1, n1y)]
        if any(0 <= a < h and 0 <= b < w and state_map[a, b] == 1 for a, b in diag):
            ranks.append(1)
            continue
        n2x, n2y = x + 2 * dx, y + 2 * dy
        if not (0 <= n2x < h and 0 <= n2y < w):
            ranks.append(2)
            continue
        v2 = state_map[n2x, n2y]
        ranks.append(2 if (v2 == 2 or v2 == 1) else 3)
    return tuple(ranks)


class UnifiedOracle:
    """mode in {"critic_only", "actor_only", "both", "trained"}."""

    def __init__(self, map_array, sff, positions, mode="critic_only", params=None, source=None,
                 v_table=None, h_table=None):
        if mode not in ("critic_only", "actor_only", "both", "trained"):
            raise ValueError(mode)
        base = TRAINED_DEFAULTS if mode == "trained" else UNIFIED_DEFAULTS
        self.params = dict(base) if params is None else {**base, **params}
        self.mode = mode
        self.map_array = np.asarray(map_array).astype(np.uint8)
        sff = np.asarray(sff)
        if mode == "critic_only":
            self.sff = sff                                                        # file dtype (:70)
        else:
            self.sff = np.where(np.isinf(sff), 0.0, sff).astype(np.float32)      # :72-76, trained :41-43
        self.dff = np.zeros_like(self.map_array, dtype=np.float32)
        self.positions = np.array(positions, dtype=np.int64).reshape(-1, 2)
        self.neighbors = list(NEUMANN) if self.params["neighborhood"] == "neumann" else list(MOORE)
        self.A = len(self.neighbors) + 1
        self.bs = self.params["block_size"]
        H, W = self.map_array.shape
        self.nbx, self.nby = -(-H // self.bs), -(-W // self.bs)
        S = self.nbx * self.nby * 256
        self.V = np.zeros(S, np.float64)
        self.v_seen = np.zeros(S, bool)
        self.H = np.zeros((S, self.A), np.float64)
        self.h_seen = np.zeros(S, bool)
        if v_table:
            for k, v in v_table.items():
                self.V[self.key_to_id(k)] = v
                self.v_seen[self.key_to_id(k)] = True
        if h_table:
            for k, v in h_table.items():
                self.H[self.key_to_id(k)] = v
                self.h_seen[self.key_to_id(k)] = True
        self.epsilon = self.params.get("epsilon", 0.0)
        self.source = source
        self.t = 0
        self.min_margin = np.inf
        self.step_log = None     # set to [] to record (states, chosen slots, rewards) per step in agent order

    # -- keys -------------------------------------------------------------------------------------
    def state_id(self, x, y, state_map):
        r = encode_ranks(int(x), int(y), state_map)
        return ((int(x) // self.bs) * self.nby + (int(y) // self.bs)) * 256 + r[0] * 64 + r[1] * 16 + r[2] * 4 + r[3]

    def key_to_id(self, key):
        r, (bx, by) = key
        return (int(bx) * self.nby + int(by)) * 256 + r[0] * 64 + r[1] * 16 + r[2] * 4 + r[3]

    def id_to_key(self, sid):
        blk, code = divmod(int(sid), 256)
        return ((code >> 6, (code >> 4) & 3, (code >> 2) & 3, code & 3), (blk // self.nby, blk % self.nby))

    def v_dict(self):
        return {self.id_to_key(s): float(self.V[s]) for s in np.flatnonzero(self.v_seen)}

    def h_dict(self):
        return {self.id_to_key(s): [float(v) for v in self.H[s]] for s in np.flatnonzero(self.h_seen)}

    # -- helpers ----------------------------------------------------------------------------------
    def _v(self, sid):
        self.v_seen[sid] = True                  # defaultdict read inserts (:658,661)
        return float(self.V[sid])

    def _reward(self, exits, coll):
        p = self.params
        r = p["step_penalty"]                    # :637
        if exits:
            r += p["exit_reward"]                # :641
        r += coll * p["collision_penalty"]       # :645-648 (every agent has a collision count)
        return r

    def _h_minmax(self):
        vals = self.H[self.h_seen]
        return (float(vals.min()), float(vals.max())) if vals.size else None

    # -- one step ----------------------------------------------------------------------------------
    def step(self):
        src, t, p = self.source, self.t, self.params
        Hh, Ww = self.map_array.shape
        pos = self.positions
        n = pos.shape[0]
        A = self.A
        learn_actor = self.mode in ("actor_only", "both")
        state_map = self.map_array.copy()
        state_map[pos[:, 0], pos[:, 1]] = 1                                       # :284-286
        occ = np.zeros((Hh, Ww), bool)
        occ[pos[:, 0], pos[:, 1]] = True
        offs = np.array(self.neighbors + [(0, 0)])
        states = np.zeros(n, np.int64)
        will_exit = np.zeros(n, bool)
        chosen_slot = np.full(n, -1, np.int64)
        slot_valid = np.zeros(n, bool)
        targets = [None] * n
        requests = {}

        for idx in range(n):
            x, y = int(pos[idx, 0]), int(pos[idx, 1])
            sid = self.state_id(x, y, state_map)                                  # :293
            states[idx] = sid
            coords = pos[idx] + offs                                              # :297-298
            inb = (coords[:, 0] >= 0) & (coords[:, 0] < Hh) & (coords[:, 1] >= 0) & (coords[:, 1] < Ww)
            valid = np.zeros(A, bool)
            is_exit = np.zeros(A, bool)
            for i in range(A):
                if not inb[i]:
                    continue
                mv = self.map_array[coords[i, 0], coords[i, 1]]
                free = True if i == A - 1 else not occ[coords[i, 0], coords[i, 1]]   # :318-321
                valid[i] = (mv == 0 or mv == 3) and free                          # :316,323
                if i < A - 1 and mv == 3:
                    is_exit[i] = True                                             # :327-332
            if is_exit.any():                                                     # :334-350
                k = int(np.flatnonzero(is_exit)[0])
                will_exit[idx] = True
                chosen_slot[idx], slot_valid[idx] = k, valid[k]
                targets[idx] = (int(coords[k, 0]), int(coords[k, 1]))
                requests.setdefault(targets[idx], []).append(idx)
                continue

            dff_vals = np.array([self.dff[c[0], c[1]] for c in coords])           # float32
            if self.mode == "critic_only":
                sff_vals = np.array([self.sff[c[0], c[1]] for c in coords])       # :355-357
                score = -p["k_S"] * sff_vals + p["k_D"] * dff_vals                # :361-364
            else:
                if self.mode == "trained":
                    h_vals = self.H[sid].astype(np.float32) if self.h_seen[sid] else np.zeros(A, np.float32)   # trained :229-236
                else:
                    self.h_seen[sid] = True                                       # zero row inserted (:405-410)
                    h_vals = self.H[sid].copy()                                   # float64 (:411)
                mm = self._h_minmax()                                             # :414-426
                if mm is not None:
                    h_min, h_max = mm
                    sff_min, sff_max = float(np.min(self.sff)), float(np.max(self.sff))
                    if np.isfinite(h_min) and np.isfinite(h_max) and h_max - h_min > 1e-6:
                        h_vals = ((h_max - h_vals) / (h_max - h_min)) * (sff_max - sff_min) + sff_min   # :435-438
                score = -p["k_A"] * h_vals + p["k_D"] * dff_vals                  # :442-445
                if np.any(np.isnan(score)) or np.any(np.isinf(score)):            # :448-455
                    score = np.zeros_like(score)
                    score[valid] = 1.0
            probs = np.exp(score - np.max(score))                                 # :367-368 / :458-459
            probs[~valid] = 0.0
            s = probs.sum()
            if np.isfinite(s) and s > 0:
                probs /= s
            else:
                probs = np.zeros_like(score)
                probs[valid] = 1.0 / valid.sum()
            k = None
            if learn_actor and self.epsilon > 0 and src.eps_coin(t, idx) < self.epsilon:   # :478-482
                vi = np.flatnonzero(valid)
                k = int(vi[int(src.eps_pick(t, idx, len(vi)) * len(vi))])         # :484-488
            if k is None:
                cdf = choice_cdf(probs)
                u = src.move(t, idx, cdf)
                self.min_margin = min(self.min_margin, float(np.min(np.abs(cdf - u))))
                k = int(cdf.searchsorted(u, side="right"))                        # :387 / :498
            chosen_slot[idx], slot_valid[idx] = k, valid[k]
            targets[idx] = (int(coords[k, 0]), int(coords[k, 1]))
            requests.setdefault(targets[idx], []).append(idx)

        # conflicts: always exactly one winner (:520-539, trained :311-323)
        nxt = pos.copy()
        coll = np.zeros(n, np.int64)
        for target, agents in requests.items():
            if len(agents) == 1:
                w = agents[0]
            else:
                k = len(agents)
                w = agents[int(src.winner(t, target[0] * Ww + target[1], k) * k)]
                coll[agents] = k - 1
            nxt[w] = target
            self.dff[pos[w, 0], pos[w, 1]] += 1

        if self.mode != "trained":
            nmap = self.map_array.copy()                                          # :543-546
            stay_in = self.map_array[nxt[:, 0], nxt[:, 1]] != 3
            nmap[nxt[stay_in, 0], nxt[stay_in, 1]] = 1
            nstate = np.full(n, -1, np.int64)
            for idx in range(n):
                if not will_exit[idx]:
                    nstate[idx] = self.state_id(nxt[idx, 0], nxt[idx, 1], nmap)
            g, av = p["gamma"], p["alpha_v"]
            td = np.zeros(n, np.float64)
            if self.step_log is not None:
                self.step_log.append(dict(states=states.copy(), slots=chosen_slot.copy(),
                                          rewards=np.array([self._reward(will_exit[i], int(coll[i])) for i in range(n)], np.float64)))
            for idx in range(n):                                                  # sequential TD(0) (:633-665)
                r = self._reward(will_exit[idx], int(coll[idx]))
                v_next = 0.0 if will_exit[idx] else self._v(nstate[idx])
                v_cur = self._v(states[idx])
                d = r + g * v_next - v_cur
                self.V[states[idx]] = v_cur + av * d
                td[idx] = d
            if self.mode == "actor_only":                                         # :568-574: recomputed with the updated V
                for idx in range(n):
                    r = self._reward(will_exit[idx], int(coll[idx]))
                    v_next = 0.0 if will_exit[idx] else self._v(nstate[idx])
                    td[idx] = r + g * v_next - self._v(states[idx])
            if learn_actor:                                                       # :745-777
                ah = p["alpha_h"]
                for idx in range(n):
                    self.h_seen[states[idx]] = True
                    if slot_valid[idx]:
                        self.H[states[idx], chosen_slot[idx]] += ah * td[idx]

        keep = self.map_array[nxt[:, 0], nxt[:, 1]] != 3                          # :601-604
        self.positions = nxt[keep]
        self.dff = update_dff(self.dff, p, self.neighbors)                        # :606
        self.t += 1

    def run(self, max_steps=None):
        traj = []
        while self.positions.shape[0] > 0 and (max_steps is None or self.t < max_steps):
            self.step()
            traj.append(self.positions.copy())
        return dict(steps=self.t, traj=traj, min_margin=self.min_margin)
