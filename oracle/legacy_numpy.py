"""NumPy restatement of the legacy 13-cell models (test infrastructure -- see oracle/__init__.py).

``AcOracle`` follows ``/root/reference/model/ffm_ac_core.py`` (``FloorFieldModel``: ``_encode_state``
:62-109, ``step`` :111-244, ``_update_critic`` :246-296, ``update_dff`` :298-318);
``ActorOnlyOracle`` follows ``/root/reference/model/ffm_actor_only.py``
(``FloorFieldModelActorOnly``: ``_encode_state`` :102-148, ``step`` :150-413 INCLUDING its
indentation slip -- the decision block sits inside the per-neighbour exit loop :214-355, so every
agent files one request per neighbour slot --, ``_update_critic`` :415-474, ``_update_actor``
:476-540).

Differences in form only:
  * the dict tables are keyed by an integer instead of the pickled tuple:
        key = ((bx * nby + by) << 26) | sum_j cell_j << (2 j)        j = 0..12 in ``state_13`` order
    (3x3 block row-major, then U2, D2, L2, R2); ``key_to_state`` / ``state_to_key`` convert;
  * an occupancy grid replaces the per-agent set building (:137-160);
  * keyed draws from a draw source (oracle/inject.py protocol) replace the global generators.
Every floating-point expression keeps the reference's dtype and operation order.
``tests/test_legacy_oracle.py`` pins this file to fixtures produced by the unmodified reference.
"""
import numpy as np

from .ffm_numpy import MOORE, NEUMANN, update_dff
from .inject import choice_cdf

AC_DEFAULTS = {                           # ffm_ac_core.py:10-23
    "k_S": 10, "k_D": 1, "diffuse": 0.2, "decay": 0.2, "neighborhood": "neumann",
    "alpha_v": 0.1, "gamma": 0.95, "exit_reward": 100.0, "step_penalty": 0.0, "collision_penalty": -1.0,
    "block_size": 3,
}
ACTOR_ONLY_DEFAULTS = {                   # ffm_actor_only.py:24-39
    "k_D": 1, "k_A": 10, "diffuse": 0.2, "decay": 0.2, "neighborhood": "neumann",
    "alpha_v": 0.1, "gamma": 0.95, "exit_reward": 100.0, "step_penalty": 0.0, "collision_penalty": -1.0,
    "alpha_h": 0.1, "epsilon": 0.0,
}
AHEAD = [(-2, 0), (2, 0), (0, -2), (0, 2)]     # U2, D2, L2, R2 (ffm_ac_core.py:89)
CELLS13 = [(a, b) for a in (-1, 0, 1) for b in (-1, 0, 1)] + AHEAD


def code13(x, y, state_map, outside):
    """The 13 cell values around (x, y) packed two bits each; ``outside`` is the value of cells
    beyond the map (2 in ffm_ac_core.py:81,98; 0 in ffm_actor_only.py:120,137)."""
    h, w = state_map.shape
    code = 0
    for j, (a, b) in enumerate(CELLS13):
        p, q = x + a, y + b
        v = int(state_map[p, q]) if (0 <= p < h and 0 <= q < w) else outside
        code |= v << (2 * j)
    return code


def state_to_key(state, nby):
    cells, (bx, by) = state
    code = 0
    for j, v in enumerate(cells):
        code |= int(v) << (2 * j)
    return ((int(bx) * nby + int(by)) << 26) | code


def key_to_state(key, nby):
    code, blk = key & ((1 << 26) - 1), key >> 26
    return (tuple((code >> (2 * j)) & 3 for j in range(13)), (blk // nby, blk % nby))


class AcOracle:
    """Legacy TD(0) critic on top of the FFM policy (ffm_ac_core.py)."""

    def __init__(self, map_array, sff, positions, params=None, source=None, v_table=None, v_default=0.0):
        self.params = dict(AC_DEFAULTS) if params is None else {**AC_DEFAULTS, **params}
        self.map_array = np.asarray(map_array).astype(np.uint8)
        self.sff = np.asarray(sff)                                              # file dtype (:28)
        self.dff = np.zeros_like(self.map_array, dtype=np.float32)
        self.positions = np.array(positions, dtype=np.int64).reshape(-1, 2)
        self.neighbors = list(NEUMANN) if self.params["neighborhood"] == "neumann" else list(MOORE)
        self.bs = self.params["block_size"]
        self.nby = (self.map_array.shape[1] + self.bs - 1) // self.bs
        self.V = dict(v_table) if v_table else {}
        self.v_default = v_default                # 0.0 (:34); -1.0 after set_v_table (:340)
        self.source = source
        self.t = 0
        self.min_margin = np.inf

    def key(self, x, y, state_map):
        return (((x // self.bs) * self.nby + (y // self.bs)) << 26) | code13(x, y, state_map, 2)

    def _v(self, k):
        if k not in self.V:                       # defaultdict read inserts (:34)
            self.V[k] = self.v_default
        return self.V[k]

    def step(self):
        P, m = self.params, self.map_array
        W = m.shape[1]
        pos = self.positions
        n = pos.shape[0]
        state_map = m.copy()                                                    # :120-122
        for x, y in pos:
            state_map[x, y] = 1
        occupied = {(int(x), int(y)) for x, y in pos}
        states, will_exit = {}, {}
        requests = {}                             # target -> [idx] in insertion order (:112)
        for idx in range(n):
            x, y = int(pos[idx, 0]), int(pos[idx, 1])
            states[idx] = self.key(x, y, state_map)                             # :129-130
            cand = [(x + dx, y + dy) for dx, dy in self.neighbors
                    if m[x + dx, y + dy] in (0, 3) and (x + dx, y + dy) not in occupied]   # :141-160
            if not cand:
                continue                                                        # :163
            cand.append((x, y))                                                 # :164
            ex = [c for c in cand if m[c] == 3]
            if ex:                                                              # :172-178
                will_exit[idx] = True
                requests.setdefault(ex[0], []).append(idx)
                continue
            rows = np.array([c[0] for c in cand]); cols = np.array([c[1] for c in cand])
            score = -P["k_S"] * self.sff[rows, cols] + P["k_D"] * self.dff[rows, cols]   # :187-190
            probs = np.exp(score - np.max(score))                               # :191-193
            s = probs.sum()
            if np.isfinite(s) and s != 0:                                       # :195
                probs /= s
                cdf = choice_cdf(probs)
                u = self.source.move(self.t, idx, cdf)
                self.min_margin = min(self.min_margin, float(np.min(np.abs(cdf - u))))
                chosen = cand[int(cdf.searchsorted(u, side="right"))]           # :197-199
                requests.setdefault(chosen, []).append(idx)
        nxt = pos.copy()
        coll = {}
        for target, agents in requests.items():                                 # :208-229
            if len(agents) == 1:
                a = agents[0]
                nxt[a] = target
                self.dff[pos[a, 0], pos[a, 1]] += 1
                coll[a] = 0
            else:
                k = len(agents)
                w = agents[int(self.source.winner(self.t, target[0] * W + target[1], k) * k)]   # random.choice (:219)
                nxt[w] = target
                self.dff[pos[w, 0], pos[w, 1]] += 1
                for a in agents:
                    coll[a] = k - 1
        state_map_next = m.copy()                                               # :233-236
        for x, y in nxt:
            if m[x, y] != 3:
                state_map_next[x, y] = 1
        for idx in range(n):                                                    # :264-296
            reward = P["step_penalty"]
            if will_exit.get(idx):
                reward += P["exit_reward"]
            if idx in coll:
                reward += coll[idx] * P["collision_penalty"]
            if will_exit.get(idx):
                v_next = 0.0
            else:
                v_next = self._v(self.key(int(nxt[idx, 0]), int(nxt[idx, 1]), state_map_next))
            v_cur = self._v(states[idx])
            td = reward + P["gamma"] * v_next - v_cur
            self.V[states[idx]] = v_cur + P["alpha_v"] * td
        self.positions = nxt[m[nxt[:, 0], nxt[:, 1]] != 3]                      # :241-244
        self.dff = update_dff(self.dff, P, self.neighbors)
        self.t += 1

    def run(self, max_steps=None):
        traj = []
        while self.positions.shape[0] > 0 and (max_steps is None or len(traj) < max_steps):
            self.step()
            traj.append(self.positions.copy())
        return traj


class ActorOnlyOracle:
    """Legacy actor (ffm_actor_only.py).  The decision block of ``step`` is nested inside the loop that collects
    the exit flags (:214-355), so an agent decides once per neighbour slot i -- with the exit flags of slots
    <= i only -- and files one request per slot; the last decision is the recorded action (:350-355).  Draw
    entities are ``idx * 8 + i``.  ``V`` lookups use the pickled state as key, so a pretrained table loaded
    under tuple keys (:61-66) is never read: ``v_inert`` only counts towards the table size."""

    BLOCK = 5                                     # :144

    def __init__(self, map_array, sff, positions, params=None, source=None, v_table=None, h_table=None, epsilon=None):
        self.params = dict(ACTOR_ONLY_DEFAULTS) if params is None else {**ACTOR_ONLY_DEFAULTS, **params}
        self.map_array = np.asarray(map_array).astype(np.uint8)
        sff = np.asarray(sff)
        self.sff = np.where(np.isinf(sff), 0.0, sff).astype(np.float32)          # :45-48
        self.dff = np.zeros_like(self.map_array, dtype=np.float32)
        self.positions = np.array(positions, dtype=np.int64).reshape(-1, 2)
        self.neighbors = list(NEUMANN) if self.params["neighborhood"] == "neumann" else list(MOORE)
        self.A = len(self.neighbors) + 1
        self.nby = (self.map_array.shape[1] + self.BLOCK - 1) // self.BLOCK
        self.V = dict(v_table) if v_table else {}
        self.H = {k: list(v) for k, v in h_table.items()} if h_table else {}
        self.epsilon = self.params.get("epsilon", 0.0) if epsilon is None else epsilon
        self.source = source
        self.t = 0
        self.min_margin = np.inf

    def key(self, x, y, occ):
        return (((x // self.BLOCK) * self.nby + (y // self.BLOCK)) << 26) | code13(x, y, occ, 0)

    def _v(self, k):
        if k not in self.V:
            self.V[k] = 0.0                       # :70
        return self.V[k]

    def step(self):
        P, m = self.params, self.map_array
        h, w = m.shape
        pos = self.positions
        n = pos.shape[0]
        nb = len(self.neighbors)
        occ = m.copy()                                                          # :163-165
        for x, y in pos:
            occ[x, y] = 1
        occupied = {(int(x), int(y)) for x, y in pos}
        states, will_exit, actions, valid_of = {}, {}, {}, {}
        requests = {}
        sff_min, sff_max = float(np.min(self.sff)), float(np.max(self.sff))
        for idx in range(n):
            x, y = int(pos[idx, 0]), int(pos[idx, 1])
            state = self.key(x, y, occ)                                         # :174-175
            states[idx] = state
            coords = [(x + dx, y + dy) for dx, dy in self.neighbors]
            inb = [0 <= a < h and 0 <= b < w for a, b in coords]
            valid = [inb[k] and m[coords[k]] in (0, 3) and coords[k] not in occupied for k in range(nb)] + [True]   # :180-206
            valid = np.array(valid, dtype=bool)
            allc = coords + [(x, y)]                                            # :210
            exit_mask = np.zeros(self.A, dtype=bool)
            for i in range(nb):                                                 # :216
                if inb[i] and m[coords[i]] == 3:
                    exit_mask[i] = True
                if np.any(exit_mask):                                           # :223-241 (nested in the loop)
                    e = int(np.where(exit_mask)[0][0])
                    will_exit[idx] = True
                    requests.setdefault(allc[e], []).append(idx)
                    actions[idx] = allc[e]
                    valid_of[idx] = valid
                    continue
                dff_vals = np.array([self.dff[c] for c in allc])                # :244-246 (float32)
                if state not in self.H or len(self.H[state]) != self.A:         # :252-257
                    self.H[state] = [0.0] * self.A
                h_vals = np.array(self.H[state])
                allh = [v for row in self.H.values() for v in row]              # :263-269
                hmin, hmax = float(np.min(allh)), float(np.max(allh))
                if not (np.isnan(hmin) or np.isnan(hmax) or np.isinf(hmin) or np.isinf(hmax)):
                    if hmax - hmin > 1e-6:                                      # :288-293
                        h_vals = ((hmax - h_vals) / (hmax - hmin)) * (sff_max - sff_min) + sff_min
                score = -P["k_A"] * h_vals + P["k_D"] * dff_vals                # :295-298
                score[~valid] = -np.inf
                if np.any(np.isnan(score)) or np.any(np.isinf(score)):          # :304-312: any invalid slot -> flat scores
                    score = np.zeros_like(score)
                    score[valid] = 1.0
                score_max = np.max(score[valid])                                # :315-319
                probs = np.exp(score - score_max)
                probs[~valid] = 0.0
                s = probs.sum()
                if np.isfinite(s) and s > 0:                                    # :324-333
                    probs /= s
                else:
                    vi = np.where(valid)[0]
                    probs = np.zeros_like(score)
                    probs[vi] = 1.0 / len(vi)
                ent = idx * 8 + i
                if self.epsilon > 0 and self.source.eps_coin(self.t, ent) < self.epsilon:   # :329-341
                    vi = np.where(valid)[0]
                    chosen = int(vi[int(self.source.eps_pick(self.t, ent, len(vi)) * len(vi))])
                else:
                    cdf = choice_cdf(probs)
                    u = self.source.move(self.t, ent, cdf)
                    self.min_margin = min(self.min_margin, float(np.min(np.abs(cdf - u))))
                    chosen = int(cdf.searchsorted(u, side="right"))             # :343
                requests.setdefault(allc[chosen], []).append(idx)               # :346-349
                actions[idx] = allc[chosen]
                valid_of[idx] = valid
        nxt = pos.copy()
        coll = {}
        for target, agents in requests.items():                                 # :360-384 (duplicates stay in the lists)
            if len(agents) == 1:
                a = agents[0]
                nxt[a] = target
                self.dff[pos[a, 0], pos[a, 1]] += 1
                coll[a] = 0
            else:
                k = len(agents)
                wn = agents[int(self.source.winner(self.t, target[0] * w + target[1], k) * k)]
                nxt[wn] = target
                self.dff[pos[wn, 0], pos[wn, 1]] += 1
                coll[wn] = k - 1
                for a in agents:
                    if a != wn:
                        coll[a] = k - 1
        occ_next = m.copy()                                                     # :388-391
        for x, y in nxt:
            if m[x, y] != 3:
                occ_next[x, y] = 1
        td = {}
        for idx in range(n):                                                    # :434-472
            reward = P["step_penalty"]
            if will_exit.get(idx):
                reward += P["exit_reward"]
            if idx in coll:
                reward += coll[idx] * P["collision_penalty"]
            if will_exit.get(idx):
                v_next = 0.0
            else:
                v_next = self._v(self.key(int(nxt[idx, 0]), int(nxt[idx, 1]), occ_next))
            v_cur = self._v(states[idx])
            td[idx] = reward + P["gamma"] * v_next - v_cur
            self.V[states[idx]] = v_cur + P["alpha_v"] * td[idx]
        for idx in range(n):                                                    # :495-536
            x, y = int(pos[idx, 0]), int(pos[idx, 1])
            allc = [(x + dx, y + dy) for dx, dy in self.neighbors] + [(x, y)]
            ci = allc.index(actions[idx])
            st = states[idx]
            if st not in self.H or len(self.H[st]) != self.A:
                self.H[st] = [0.0] * self.A
            if valid_of[idx][ci]:
                self.H[st][ci] += P["alpha_h"] * td[idx]                        # :534
        self.positions = nxt[m[nxt[:, 0], nxt[:, 1]] != 3]                      # :408-411
        self.dff = update_dff(self.dff, P, self.neighbors)
        self.t += 1

    def run(self, max_steps=None):
        traj = []
        while self.positions.shape[0] > 0 and (max_steps is None or len(traj) < max_steps):
            self.step()
            traj.append(self.positions.copy())
        return traj
