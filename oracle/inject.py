"""Injected-draw harness for the UNMODIFIED reference classes (test infrastructure).

Runs only in the build container, where ``/root/reference`` exists; it is how the golden
fixtures under ``tests/golden/`` are produced (``oracle/make_golden.py``) and how the NumPy
restatement in ``oracle/ffm_numpy.py`` is pinned to the reference.

The reference draws from process-global generators, in agent order:
  * ``np.random.choice(n, p=probs)``       model/ffm_core.py:84, model/ffm_unified.py:387,498
  * ``np.random.rand() < 0.5``             model/ffm_core.py:95
  * ``random.choice(agents)``              model/ffm_core.py:96, model/ffm_unified.py:530
  * ``random.random() < epsilon``          model/ffm_unified.py:481
  * ``np.random.randint(len(valid))``      model/ffm_unified.py:487
(the legacy models draw at the same kinds of call sites: model/ffm_ac_core.py:194,215; model/ffm_actor_only.py:329,333,343,370)
Those five call sites are monkey-patched for the duration of a ``with injected(...)`` block.
Each patched function reads the caller's frame to learn WHICH agent / target cell the draw
is for and asks a *draw source* for the uniform, which it maps to a result exactly the way
the original generator maps its own uniform (``searchsorted(cdf, u, 'right')`` for
``choice`` -- numpy/random/mtrand.pyx ``RandomState.choice``).

Draw-source protocol (all uniforms are float64 in [0, 1)):
    move(step, idx, cdf) -> u           choice index = searchsorted(cdf, u, 'right')
    coin(step, cell) -> u               somebody moves iff u < 0.5          (ffm_core only)
    winner(step, cell, k) -> u          winner = agents[int(u * k)]
    eps_coin(step, idx) -> u            explore iff u < epsilon
    eps_pick(step, idx, k) -> u         valid_indices[int(u * k)]
"""
import contextlib
import random as _pyrandom
import sys

import numpy as np

from . import philox

REFERENCE_ROOT = "/root/reference"


def import_reference(module):
    """Import ``model.<module>`` from the read-only reference tree (container only)."""
    import importlib

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module("model." + module)


def choice_cdf(p):
    """CDF used by legacy ``RandomState.choice``: p -> float64, cumsum, divide by the last entry."""
    cdf = np.asarray(p, dtype=np.float64).cumsum()
    cdf /= cdf[-1]
    return cdf


class PhiloxSource:
    """Keyed draws: (seed, episode) fixed, (step, stream, entity) per call (oracle/philox.py)."""

    def __init__(self, seed, episode=0):
        self.seed = int(seed)
        self.episode = int(episode)

    def _d(self, step, stream, entity):
        return philox.draw2_scalar(self.seed, self.episode, step, stream, entity)

    def move(self, step, idx, cdf=None):
        return self._d(step, philox.STREAM_MOVE, idx)[0]

    def coin(self, step, cell):
        return self._d(step, philox.STREAM_CONFLICT, cell)[0]

    def winner(self, step, cell, k):
        return self._d(step, philox.STREAM_CONFLICT, cell)[1]

    def eps_coin(self, step, idx):
        return self._d(step, philox.STREAM_EPS, idx)[0]

    def eps_pick(self, step, idx, k):
        return self._d(step, philox.STREAM_EPS, idx)[1]


class StockSource:
    """Consumes the reference's OWN global generators (the legacy MT19937 streams of ``np.random``
    and ``random``) exactly as the unpatched code would, so that a stock seeded run
    (``main.py:23-26``) can be recorded and replayed through the CUDA path's draw buffers."""

    def __init__(self):
        self._sample = np.random.random_sample
        self._rand = np.random.rand
        self._randint = np.random.randint
        self._randbelow = _pyrandom._inst._randbelow
        self._random = _pyrandom.random

    def move(self, step, idx, cdf=None):
        return float(self._sample())

    def coin(self, step, cell):
        return float(self._rand())

    def winner(self, step, cell, k):
        return (int(self._randbelow(k)) + 0.5) / k      # random.choice == seq[_randbelow(len)]

    def eps_coin(self, step, idx):
        return float(self._random())

    def eps_pick(self, step, idx, k):
        return (int(self._randint(k)) + 0.5) / k


class RecordingSource:
    """Wraps a source and stores every draw in the dense buffers the CUDA path replays:
    ``move[t, idx]``, ``conflict[t, cell, 2]`` (coin, winner), ``eps[t, idx, 2]`` (NaN = not drawn).

    ``guard`` > 0 re-draws (from a private MT19937) any move draw that lies within ``guard`` of a
    CDF boundary, which makes the recorded stream insensitive to last-ulp differences of ``exp``
    between NumPy's SIMD kernels and CUDA's (first hard part).
    """

    def __init__(self, base, max_steps, n_agents, n_cells, guard=0.0, redraw_seed=12345):
        self.base = base
        self.guard = guard
        self.move_buf = np.full((max_steps, n_agents), np.nan)
        self.conflict_buf = np.full((max_steps, n_cells, 2), np.nan)
        self.eps_buf = np.full((max_steps, n_agents, 2), np.nan)
        self._redraw = np.random.RandomState(redraw_seed)
        self.redraws = 0

    def move(self, step, idx, cdf=None):
        u = self.base.move(step, idx, cdf)
        if self.guard > 0 and cdf is not None:
            while np.min(np.abs(cdf - u)) < self.guard:
                u = float(self._redraw.random_sample())
                self.redraws += 1
        self.move_buf[step, idx] = u
        return u

    def coin(self, step, cell):
        u = self.base.coin(step, cell)
        self.conflict_buf[step, cell, 0] = u
        return u

    def winner(self, step, cell, k):
        u = self.base.winner(step, cell, k)
        self.conflict_buf[step, cell, 1] = u
        return u

    def eps_coin(self, step, idx):
        u = self.base.eps_coin(step, idx)
        self.eps_buf[step, idx, 0] = u
        return u

    def eps_pick(self, step, idx, k):
        u = self.base.eps_pick(step, idx, k)
        self.eps_buf[step, idx, 1] = u
        return u


class _State:
    def __init__(self, source, width):
        self.source = source
        self.width = width
        self.step = 0
        self.min_margin = np.inf
        self.n_move = self.n_coin = self.n_winner = self.n_eps = 0
        self.probs_log = None
        self.sub_key = False      # model/ffm_actor_only.py draws up to `neighbours` times per agent and step: entity = idx * 8 + i


@contextlib.contextmanager
def injected(source, width, log_probs=False, sub_key=False):
    """Patch the five draw call sites; yields a state object whose ``.step`` the caller sets to
    the CA step number before each ``model.step()``."""
    st = _State(source, width)
    st.sub_key = sub_key

    def _agent(frame):
        loc = frame.f_locals
        return int(loc["idx"]) * 8 + int(loc["i"]) if st.sub_key else int(loc["idx"])
    if log_probs:
        st.probs_log = []
    o_choice, o_rand, o_randint = np.random.choice, np.random.rand, np.random.randint
    o_pychoice, o_pyrandom = _pyrandom.choice, _pyrandom.random

    def _cell(frame):
        loc = frame.f_locals
        t = loc["target"] if "target" in loc else (loc["tx"], loc["ty"])    # ffm_learning_core.py:229 unpacks (tx, ty)
        return int(t[0]) * st.width + int(t[1])

    def choice(a, size=None, replace=True, p=None):
        if p is None:                      # placement draw in the constructor (ffm_core.py:25)
            return o_choice(a, size=size, replace=replace, p=p)
        idx = _agent(sys._getframe(1))
        cdf = choice_cdf(p)
        u = st.source.move(st.step, idx, cdf)
        st.min_margin = min(st.min_margin, float(np.min(np.abs(cdf - u))))
        st.n_move += 1
        if st.probs_log is not None:
            st.probs_log.append((st.step, idx, np.array(p, copy=True)))
        return int(cdf.searchsorted(u, side="right"))

    def rand(*shape):
        if shape:
            return o_rand(*shape)
        st.n_coin += 1
        return st.source.coin(st.step, _cell(sys._getframe(1)))

    def py_choice(seq):
        st.n_winner += 1
        k = len(seq)
        return seq[int(st.source.winner(st.step, _cell(sys._getframe(1)), k) * k)]

    def py_random():
        st.n_eps += 1
        return st.source.eps_coin(st.step, _agent(sys._getframe(1)))

    def randint(low, high=None, size=None, dtype=int):
        if high is not None or size is not None:
            return o_randint(low, high, size, dtype)
        return int(st.source.eps_pick(st.step, _agent(sys._getframe(1)), low) * low)

    np.random.choice, np.random.rand, np.random.randint = choice, rand, randint
    _pyrandom.choice, _pyrandom.random = py_choice, py_random
    try:
        yield st
    finally:
        np.random.choice, np.random.rand, np.random.randint = o_choice, o_rand, o_randint
        _pyrandom.choice, _pyrandom.random = o_pychoice, o_pyrandom


def run_reference(model, source, max_steps=None, log_probs=False, keep_dff=True, sub_key=False):
    """Drive a reference model object (any ``model/ffm_*.py`` class) step by step under injected
    draws until evacuation / ``max_steps``.

    Returns dict(steps, traj=[positions after each step], dff=[dff after each step],
    min_margin, n_move, n_coin, n_winner, n_eps, probs_log)."""
    width = model.map_array.shape[1]
    traj, dffs = [], []
    with injected(source, width, log_probs=log_probs, sub_key=sub_key) as st:
        t = 0
        while model.positions.shape[0] > 0 and (max_steps is None or t < max_steps):
            st.step = t
            model.step()
            traj.append(np.array(model.positions, dtype=np.int64).reshape(-1, 2))
            if keep_dff:
                dffs.append(np.array(model.dff, dtype=np.float32, copy=True))
            t += 1
    return dict(steps=t, traj=traj, dff=dffs, min_margin=st.min_margin, n_move=st.n_move,
                n_coin=st.n_coin, n_winner=st.n_winner, n_eps=st.n_eps, probs_log=st.probs_log)


class BufferSource:
    """Replays recorded draws: ``move[t, idx]`` dense, conflicts as a dict {(t, cell): (coin, winner)}."""

    def __init__(self, move, conflict):
        self._move = move
        self._conflict = conflict

    def move(self, step, idx, cdf=None):
        return float(self._move[step, idx])

    def coin(self, step, cell):
        return float(self._conflict[(step, cell)][0])

    def winner(self, step, cell, k):
        return float(self._conflict[(step, cell)][1])
