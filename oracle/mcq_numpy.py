"""NumPy restatement of the target-centric Monte-Carlo Q-learning model (test infrastructure).

Follows ``/root/reference/model/ffm_learning_core.py`` (``FloorFieldModel``: ``_combined3x3_at_target`` :115-140,
``step(beta)`` :145-285, ``_update_dff`` :307-321, ``finalize_timeouts`` :326-360) with

  * a dense table instead of the dict: state id = block * 4**9 + sum(v[i] * 4**i) over the row-major 3x3 window
    (v in {0,1,2,3}), block = (tx // 3) * nby + ty // 3;  ``Q[S, 5]`` float32 + ``q_seen[S]`` (rows are created by
    ``_ensure_qvec`` :289-291, never by the read path :190-191);
  * per-agent paths kept as (state id, action, reward) lists exactly like ``self.paths`` (:79-81);
  * keyed draws from a draw source instead of the global generators.

Arithmetic: logits are Python floats (:193); the table update is NumPy float32 scalar arithmetic
``Q[s][a] += alpha * (G - Q[s][a])`` with the Python floats weakly cast (NEP 50), kept literally.
"""
import numpy as np

from .inject import choice_cdf

MCQ_DEFAULTS = {"k_S": 3.0, "k_D": 1.0, "k_Q": 1.0, "diffuse": 0.2, "decay": 0.2, "neighborhood": "neumann",
                "step_penalty": 0.0, "stop_penalty": 0.0, "collision_penalty": 0.0, "exit_reward": 100.0,
                "timeout_penalty": 50.0, "max_steps": 500}        # ffm_learning_core.py:45-59
NEIGHBORS = [(-1, 0), (1, 0), (0, -1), (0, 1)]                     # UP, DOWN, LEFT, RIGHT (:73)
FROM_UP, FROM_DOWN, FROM_LEFT, FROM_RIGHT, FROM_SELF = range(5)
FROM_OF_MOVE = [FROM_DOWN, FROM_UP, FROM_RIGHT, FROM_LEFT]         # _dir_to_from (:294-305) per neighbour index
MOORE = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]


class McqOracle:
    def __init__(self, map_array, sff, positions, params=None, source=None, alpha=0.1, gamma=0.99):
        self.params = dict(MCQ_DEFAULTS) if params is None else {**MCQ_DEFAULTS, **params}
        self.map_array = np.asarray(map_array).astype(np.uint8)
        self.sff = np.asarray(sff)
        self.H, self.W = self.map_array.shape
        self.dff = np.zeros((self.H, self.W), np.float32)
        self.nby = -(-self.W // 3)
        self.S = (-(-self.H // 3)) * self.nby * 4 ** 9
        self.Q = np.zeros((self.S, 5), np.float32)
        self.q_seen = np.zeros(self.S, bool)
        self.alpha, self.gamma = alpha, gamma
        self.source = source
        self.reset(positions)
        self.min_margin = np.inf

    def reset(self, positions):
        self.positions = np.array(positions, dtype=np.int64).reshape(-1, 2)
        self.paths = [[] for _ in range(len(self.positions))]
        self.dff.fill(0.0)
        self.t = 0

    # -- state ------------------------------------------------------------------------------------------
    def state_id(self, tx, ty, occ):
        """combined3x3 (:115-140): map window (OOB = 2) + occupancy where the map is free; the agent itself is NOT
        excluded (the exclusion is commented out in the reference, :132-135)."""
        code = 0
        k = 0
        for a in range(tx - 1, tx + 2):
            for b in range(ty - 1, ty + 2):
                if 0 <= a < self.H and 0 <= b < self.W:
                    v = int(self.map_array[a, b])
                    if v == 0 and occ[a, b]:
                        v = 1
                else:
                    v = 2
                code += v << (2 * k)
                k += 1
        return ((tx // 3) * self.nby + ty // 3) * 4 ** 9 + code

    def key_of(self, sid):
        blk, code = divmod(int(sid), 4 ** 9)
        cells = bytes((code >> (2 * k)) & 3 for k in range(9))
        return (cells, (blk // self.nby, blk % self.nby))

    def id_of(self, key):
        cells, (bx, by) = key
        return (int(bx) * self.nby + int(by)) * 4 ** 9 + sum(int(v) << (2 * k) for k, v in enumerate(cells))

    def q_dict(self):
        return {self.key_of(s): self.Q[s].copy() for s in np.flatnonzero(self.q_seen)}

    def _backup(self, path):
        """Reverse Monte-Carlo backup (:262-267, :350-355)."""
        G = 0.0
        for sid, ac, r in reversed(path):
            G = r + self.gamma * G
            self.q_seen[sid] = True
            self.Q[sid][ac] += self.alpha * (G - self.Q[sid][ac])

    # -- one step ----------------------------------------------------------------------------------------
    def step(self, beta):
        src, t, p = self.source, self.t, self.params
        k_S, k_D, k_Q = float(p["k_S"]), float(p["k_D"]), float(p["k_Q"])
        step_pen, stop_pen, coll_pen = float(p["step_penalty"]), float(p["stop_penalty"]), float(p["collision_penalty"])
        pos = self.positions
        n = len(pos)
        occ = np.zeros((self.H, self.W), bool)
        occ[pos[:, 0], pos[:, 1]] = True
        passable = (self.map_array == 0) | (self.map_array == 3)
        requests, nxt, arrived = {}, pos.copy(), []
        for idx in range(n):
            x, y = int(pos[idx, 0]), int(pos[idx, 1])
            cand = []
            for k, (dx, dy) in enumerate(NEIGHBORS):
                tx, ty = x + dx, y + dy
                if 0 <= tx < self.H and 0 <= ty < self.W and passable[tx, ty] and not occ[tx, ty]:
                    cand.append((tx, ty, FROM_OF_MOVE[k]))
            cand.append((x, y, FROM_SELF))
            logits, sids = [], []
            for tx, ty, a in cand:
                sid = self.state_id(tx, ty, occ)
                q_val = float(self.Q[sid][a]) if self.q_seen[sid] else 0.0          # read path creates nothing (:190-191)
                logits.append(beta * (-k_S * float(self.sff[tx, ty])) + k_D * float(self.dff[tx, ty]) + (1 - beta) * k_Q * q_val)
                sids.append(sid)
            la = np.asarray(logits, dtype=np.float64)
            probs = np.exp(la - np.max(la))
            s = probs.sum()
            if not np.isfinite(s) or s <= 0:
                chosen = len(cand) - 1
            else:
                probs /= s
                cdf = choice_cdf(probs)
                u = src.move(t, idx, cdf)
                self.min_margin = min(self.min_margin, float(np.min(np.abs(cdf - u))))
                chosen = int(cdf.searchsorted(u, side="right"))
            tx, ty, a = cand[chosen]
            self.q_seen[sids[chosen]] = True                                       # _ensure_qvec (:221)
            self.paths[idx].append((sids[chosen], a, -stop_pen if a == FROM_SELF else -step_pen))
            requests.setdefault((tx, ty), []).append(idx)
        for (tx, ty), agents in requests.items():
            if len(agents) == 1:
                winner = agents[0]
            else:
                k = len(agents)
                winner = agents[int(src.winner(t, tx * self.W + ty, k) * k)]
                for i in agents:
                    if i != winner and self.paths[i]:
                        sk, ac, _ = self.paths[i][-1]
                        self.paths[i][-1] = (sk, ac, -coll_pen)                    # (:253-257)
            sx, sy = int(pos[winner, 0]), int(pos[winner, 1])
            if (tx, ty) != (sx, sy):
                self.dff[sx, sy] += 1.0                                            # only real moves leave a footprint (:235,247)
                nxt[winner] = (tx, ty)
            if self.map_array[tx, ty] == 3:
                arrived.append(winner)
        self.positions = nxt
        for idx in sorted(arrived, reverse=True):                                  # (:263-278)
            if self.paths[idx]:
                sk, ac, _ = self.paths[idx][-1]
                self.paths[idx][-1] = (sk, ac, float(p["exit_reward"]))
            self._backup(self.paths[idx])
            self.positions = np.delete(self.positions, idx, axis=0)
            del self.paths[idx]
        self._update_dff()
        self.t += 1
        if self.t >= int(p["max_steps"]) and len(self.positions) > 0:              # (:284-285)
            self.finalize_timeouts()

    def _update_dff(self):
        """ffm_learning_core.py:307-321: always Moore, sum of the shifted fields first, one multiply after."""
        diffuse, decay = float(self.params["diffuse"]), float(self.params["decay"])
        base = (1.0 - decay) * (1.0 - diffuse) * self.dff
        padded = np.pad(base, 1, mode="constant")
        acc = np.zeros_like(base)
        for dx, dy in MOORE:
            acc += padded[1 + dx:self.H + 1 + dx, 1 + dy:self.W + 1 + dy]
        acc *= decay * (1.0 - diffuse) / len(MOORE)
        self.dff = base + acc
        self.dff[self.dff < 1e-4] = 0.0

    def finalize_timeouts(self):
        """ffm_learning_core.py:326-360."""
        if len(self.positions) == 0:
            return
        occ = np.zeros((self.H, self.W), bool)
        occ[self.positions[:, 0], self.positions[:, 1]] = True
        for idx in range(len(self.positions)):
            x, y = int(self.positions[idx, 0]), int(self.positions[idx, 1])
            sid = self.state_id(x, y, occ)
            self.q_seen[sid] = True
            self.paths[idx].append((sid, FROM_SELF, -float(self.params["timeout_penalty"])))
            self._backup(self.paths[idx])
        self.positions = np.empty((0, 2), dtype=np.int64)
        self.paths = []

    def run(self, beta, max_steps=None):
        traj = []
        while len(self.positions) > 0 and (max_steps is None or self.t < max_steps):
            self.step(beta)
            traj.append(self.positions.copy())
        return dict(steps=self.t, traj=traj, min_margin=self.min_margin)


# ---- coverage pretrain (run_coverage_pretrain_and_training.py:60-216) --------------------------------------------------
DIR_TO_DXY = {FROM_UP: (-1, 0), FROM_DOWN: (1, 0), FROM_LEFT: (0, -1), FROM_RIGHT: (0, 1), FROM_SELF: (0, 0)}   # :61-67, src = T + delta


def valid_from_dirs_for_target(map_array, tx, ty):
    """run_coverage_pretrain_and_training.py:78-88: the FROM_* whose source cell is free, STOP always."""
    H, W = map_array.shape
    out = []
    for a, (dx, dy) in DIR_TO_DXY.items():
        sx, sy = tx + dx, ty + dy
        if a == FROM_SELF or (0 <= sx < H and 0 <= sy < W and map_array[sx, sy] == 0):
            out.append(a)
    return out


def coverage_order(map_array):
    """Every (tx, ty, from_dir) pattern of coverage_pretrain_empty (:173-216) in its UNSHUFFLED order: free targets in
    np.where order, directions in dict order."""
    xs, ys = np.where(np.asarray(map_array) == 0)
    return [(int(x), int(y), a) for x, y in zip(xs, ys) for a in valid_from_dirs_for_target(map_array, int(x), int(y))]


def forced_step_cap(sff, sx, sy, step_buffer=10):
    return int(min(200, max(1, float(sff[sx, sy]) + step_buffer)))       # :150-151


def force_first_step_and_roll(o, T, from_dir, step_buffer=10):
    """force_first_step_and_roll (:91-166) on a McqOracle whose table is the shared Q: one agent, teacher-forced first
    record, then step(beta=1.0) until exit or the SFF-derived cap, then finalize_timeouts.  Returns the CA steps run."""
    tx, ty = T
    dx, dy = DIR_TO_DXY[from_dir]
    sx, sy = tx + dx, ty + dy
    m = o.map_array
    if m[tx, ty] != 0:
        return 0
    if from_dir != FROM_SELF and not (0 <= sx < o.H and 0 <= sy < o.W and m[sx, sy] == 0):
        return 0
    o.reset(np.array([[sx, sy]]))
    occ = np.zeros((o.H, o.W), bool)
    occ[sx, sy] = True
    sid = o.state_id(tx, ty, occ)
    o.q_seen[sid] = True
    p = o.params
    if from_dir == FROM_SELF:
        reward = -float(p["stop_penalty"])
    else:
        reward = -float(p["step_penalty"])
        o.dff[sx, sy] += 1.0
        o.positions[0] = (tx, ty)
    o.paths[0].append((sid, from_dir, reward))
    cap = forced_step_cap(o.sff, sx, sy, step_buffer)
    steps = 0
    while len(o.positions) > 0 and steps < cap:
        o.step(1.0)
        steps += 1
    if len(o.positions) > 0:
        o.finalize_timeouts()
    return steps


def coverage_pretrain(map_array, sff, params, order, seed, alpha=0.1, gamma=0.99):
    """coverage_pretrain_empty (:173-216) over the given (tx, ty, from_dir) order; mini-episode k draws from the keyed
    streams of episode k.  Returns (oracle holding the shared Q, steps per mini-episode)."""
    from .inject import PhiloxSource
    o = McqOracle(map_array, sff, np.zeros((0, 2)), params, None, alpha, gamma)
    steps = []
    for k, (tx, ty, a) in enumerate(order):
        o.source = PhiloxSource(seed, k)
        steps.append(force_first_step_and_roll(o, (tx, ty), a))
    return o, steps
