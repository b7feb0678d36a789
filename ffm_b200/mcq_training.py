"""Batched GPU forms of the reference's Monte-Carlo Q-learning pipeline (run_coverage_pretrain_and_training.py):

  coverage_pretrain   coverage_pretrain_empty (:173-216): for every free target T and every valid FROM_* (plus STOP) one
                      mini-episode with a teacher-forced first transition (force_first_step_and_roll, :91-166), cap =
                      SFF(src) + 10 steps, shared Q.  All patterns run as ONE launch (one CTA per mini-episode); because the
                      policy of the pretrain does not read Q (beta = 1.0), applying the reverse Monte-Carlo backups afterwards
                      in pattern order (ffm_mcq_backup_ordered) gives the reference's shared dict bit for bit.
  McqBatchedLearner   the MC analogue of the unified model's batched TD learning: B episodes against a frozen table, returns
                      reduced per (state, action), exchanged BY KEY between GPUs (the hash-table slot of a key is a local
                      matter), folded in with the visit count.
  run_training        the N ramp / beta schedule of main() (:313-333: compute_agent_count :41-46, compute_beta :26-38) with
                      `batch` episodes per schedule entry.

Host logic only; the arithmetic is in csrc/ffm_mcq_kernel.cuh.
"""
import random

import numpy as np
import torch
import torch.distributed as dist

from .sim import MCQ_DEFAULTS, McqSim

FROM_UP, FROM_DOWN, FROM_LEFT, FROM_RIGHT, FROM_SELF = range(5)
DIR_TO_DXY = {FROM_UP: (-1, 0), FROM_DOWN: (1, 0), FROM_LEFT: (0, -1), FROM_RIGHT: (0, 1), FROM_SELF: (0, 0)}   # :61-67


def compute_beta(episode_step):
    """run_coverage_pretrain_and_training.py:26-38."""
    if episode_step <= 50:
        return 1.0
    if episode_step <= 650:
        return 1.0 - (episode_step - 50) / 600.0
    return 0.0


def compute_agent_count(episode, full_N):
    """run_coverage_pretrain_and_training.py:41-46."""
    if episode < 500:
        return max(1, full_N * (episode // 50 + 1) // 10)
    return full_N


def valid_from_dirs_for_target(map_array, tx, ty):
    """:78-88."""
    H, W = map_array.shape
    out = []
    for a, (dx, dy) in DIR_TO_DXY.items():
        sx, sy = tx + dx, ty + dy
        if a == FROM_SELF or (0 <= sx < H and 0 <= sy < W and map_array[sx, sy] == 0):
            out.append(a)
    return out


def coverage_patterns(map_array, shuffle=True):
    """The (tx, ty, from_dir) patterns in the order coverage_pretrain_empty (:181-199) visits them: free targets in np.where
    order, shuffled with the process-global ``random`` exactly where the reference shuffles (targets once, each target's
    directions), so ``random.seed(s)`` reproduces the reference's order."""
    xs, ys = np.where(map_array == 0)
    targets = [(int(x), int(y)) for x, y in zip(xs, ys)]
    if shuffle:
        random.shuffle(targets)
    order = []
    for tx, ty in targets:
        dirs = valid_from_dirs_for_target(map_array, tx, ty)
        random.shuffle(dirs)                                   # :199 (unconditional in the reference)
        order += [(tx, ty, a) for a in dirs]
    return order


def coverage_pretrain(map_array, sff, params, shared_Q=None, order=None, shuffle=True, seed=0, step_buffer=10, device=None,
                      return_steps=False):
    """coverage_pretrain_empty (:173-216) as one batched launch; ``shared_Q`` (a dict, updated in place and returned) plays
    the role of the reference's shared dict; alpha / gamma come from ``params`` like in force_first_step_and_roll (:113-114).
    Mini-episode k draws from the keyed Philox streams of episode k."""
    m = np.ascontiguousarray(np.asarray(map_array).astype(np.uint8))
    sff = np.asarray(sff)
    p = {**MCQ_DEFAULTS, **(params or {})}
    if order is None:
        order = coverage_patterns(m, shuffle)
    order = [(int(tx), int(ty), int(a)) for tx, ty, a in order]
    B = len(order)
    shared_Q = {} if shared_Q is None else shared_Q
    if B == 0:
        return (shared_Q, np.zeros(0, np.int32)) if return_steps else shared_Q
    src = np.array([(tx + DIR_TO_DXY[a][0], ty + DIR_TO_DXY[a][1]) for tx, ty, a in order], np.int32)
    caps = np.array([int(min(200, max(1, float(sff[sx, sy]) + step_buffer))) for sx, sy in src], np.int32)      # :150-151
    sim = McqSim(m, sff, B, 1, learn="batched", params=p, seed=seed, alpha=float(p.get("alpha", 0.1)), gamma=float(p.get("gamma", 0.99)),
                 device=device)
    if shared_Q:
        sim.load_q_dict(shared_Q)
    sim.set_beta(1.0)
    sim.set_positions(src.reshape(B, 1, 2), np.ones(B, np.int32))
    sim.set_forced(np.array([(tx, ty) for tx, ty, _ in order]), np.array([a for _, _, a in order]), caps)
    sim.rollout(int(caps.max()))
    sim.backup_ordered()
    steps = sim.counters()[0]
    shared_Q.clear()
    shared_Q.update(sim.q_dict())
    sim.close()
    return (shared_Q, steps) if return_steps else shared_Q


class McqBatchedLearner:
    """Synchronous batched Monte-Carlo learning: after a rollout of ``sim`` (learn="batched"), ``sync()`` reduces the returns
    per (state, action) on this rank, exchanges the touched rows by key with the other ranks (one all-gather of the padded
    lists: the path's only collective), imports every rank's list in rank order -- so all ranks hold bit-identical sums and
    therefore bit-identical tables -- and folds them in."""

    def __init__(self, sim, distributed=True, export_capacity=1 << 15):
        assert sim.learn == "batched"
        self.sim, self.distributed, self.capacity = sim, distributed, int(export_capacity)
        self._send = self._recv = None

    def sync(self):
        """No host synchronisation: export kernel -> ONE all-gather of the fixed-size [count | keys | rows] messages -> one
        import kernel per rank (each reads its list's count on the device) -> fold.  A list longer than ``export_capacity``
        raises at the next host read of the handle."""
        sim = self.sim
        sim.accumulate()
        world = dist.get_world_size() if self.distributed and dist.is_available() and dist.is_initialized() else 1
        if world > 1:
            cap = self.capacity
            size = 1 + 11 * cap
            if self._send is None:
                self._send = torch.zeros(size, dtype=torch.float64, device=sim.delta_device())
                self._recv = torch.zeros(world * size, dtype=torch.float64, device=sim.delta_device())
            sim.export_deltas(cap, out=self._send)
            dist.all_gather_into_tensor(self._recv, self._send)
            for r in range(world):
                msg = self._recv[r * size:(r + 1) * size]
                sim.import_deltas(msg[1:1 + cap].view(torch.int64), msg[1 + cap:].view(cap, 10), cap, count_dev=msg[:1].view(torch.int32)[:1])
        sim.fold()


def run_training(map_array, sff, params, full_N, shared_Q=None, num_episodes=1200, batch=64, seed=0, device=None, log=None,
                 entries=None):
    """The schedule of main() (:313-333) with ``batch`` episodes per entry: entry k uses N = compute_agent_count(k) agents and
    beta = 1.0 for k < 500, compute_beta(k - 500) afterwards; the table is folded after every entry (McqBatchedLearner).
    ``entries``: the schedule entries to run (default: all of range(num_episodes)).  Returns (Q dict, mean steps per entry)."""
    m = np.ascontiguousarray(np.asarray(map_array).astype(np.uint8))
    p = {**MCQ_DEFAULTS, **(params or {})}
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)
    sim = McqSim(m, np.asarray(sff), batch, int(full_N), learn="batched", params=p, seed=seed, alpha=float(p.get("alpha", 0.1)),
                 gamma=float(p.get("gamma", 0.99)), device=device)
    if shared_Q:
        sim.load_q_dict(shared_Q)
    learner = McqBatchedLearner(sim)
    mean_steps = []
    for k in (range(int(num_episodes)) if entries is None else entries):
        N = compute_agent_count(k, int(full_N))
        beta = 1.0 if k < 500 else compute_beta(k - 500)
        sim.set_episode_base((k * world + rank) * batch)
        sim.set_beta(beta)
        sim.place(np.full(batch, N, np.int32))
        sim.rollout(int(p["max_steps"]) + 1)
        learner.sync()
        steps = sim.counters()[0]
        mean_steps.append(float(steps.mean()))
        if log and (k + 1) % 50 == 0:
            log(f"entry {k + 1}/{num_episodes}: N={N} beta={beta:.3f} mean steps {mean_steps[-1]:.1f}")
    Q = sim.q_dict()
    sim.close()
    return Q, mean_steps
