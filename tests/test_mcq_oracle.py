"""Pins oracle/mcq_numpy.py to fixtures produced by the unmodified model/ffm_learning_core.py (multi-episode runs with
a shared Q table, beta schedules, timeouts): trajectories, the Q dict (keys and float32 rows), the DFF."""
import numpy as np
import pytest

from helpers import load_mcq, MCQ_FIXTURES
from oracle import mcq_numpy
from oracle.inject import PhiloxSource


@pytest.mark.parametrize("name", MCQ_FIXTURES)
def test_mcq_oracle_reproduces_reference(name):
    g = load_mcq(name)
    o = mcq_numpy.McqOracle(g["map"], g["sff"], g["ep"][0]["pos0"], g["params"], None, alpha=float(g["alpha"]), gamma=float(g["gamma"]))
    for ep, E in enumerate(g["ep"]):
        o.reset(E["pos0"])
        o.source = PhiloxSource(int(g["seed"]), ep)
        r = o.run(float(g["betas"][ep]))
        assert r["steps"] == int(g["steps"][ep])
        for t, want in enumerate(E["traj"]):
            assert np.array_equal(r["traj"][t], want), (ep, t)
    assert np.array_equal(np.flatnonzero(o.q_seen), g["q_ids"])
    assert np.array_equal(o.Q[g["q_ids"]].view(np.uint32), g["q_rows"].view(np.uint32))
    assert np.array_equal(o.dff.view(np.uint32), g["final_dff"].view(np.uint32))


def test_state_key_round_trip():
    m = np.zeros((9, 9), np.uint8); m[0, :] = m[-1, :] = m[:, 0] = m[:, -1] = 2; m[0, 4] = 3
    o = mcq_numpy.McqOracle(m, np.zeros((9, 9)), [[1, 4], [2, 4]])
    occ = np.zeros((9, 9), bool); occ[1, 4] = occ[2, 4] = True
    sid = o.state_id(1, 4, occ)
    cells, blk = o.key_of(sid)
    assert blk == (0, 1) and list(cells) == [2, 3, 2, 0, 1, 0, 0, 1, 0]      # wall, exit, wall / free, self, free / free, ped, free
    assert o.id_of((cells, blk)) == sid
    assert o.key_of(o.state_id(0, 4, occ))[0][:3] == bytes([2, 2, 2])        # out-of-bounds row counts as wall (:122)


from helpers import PRETRAIN_FIXTURES, load_pretrain


@pytest.mark.parametrize("name", PRETRAIN_FIXTURES)
def test_coverage_pretrain_oracle_reproduces_reference(name):
    """oracle/mcq_numpy.coverage_pretrain (restating run_coverage_pretrain_and_training.py:91-216) against the shared Q the
    unmodified driver produced: per-pattern CA step counts, Q keys, float32 rows."""
    g = load_pretrain(name)
    o, steps = mcq_numpy.coverage_pretrain(g["map"], g["sff"], g["params"], g["order"], int(g["seed"]), float(g["alpha"]), float(g["gamma"]))
    assert np.array_equal(np.array(steps), g["steps"])
    assert np.array_equal(np.flatnonzero(o.q_seen), g["q_ids"])
    assert np.array_equal(o.Q[g["q_ids"]].view(np.uint32), g["q_rows"].view(np.uint32))


def test_coverage_patterns_cover_every_target_and_direction():
    """The product-side pattern enumeration (ffm_b200/mcq_training.py) == the oracle's restatement of :70-88,:181-199, and the
    reference's shuffle is reproduced by random.seed."""
    import random
    from ffm_b200.mcq_training import coverage_patterns
    from oracle import assets
    m = assets.room_map(9, 14)
    assert coverage_patterns(m, shuffle=False) != [] and sorted(coverage_patterns(m, shuffle=False)) == sorted(mcq_numpy.coverage_order(m))
    g = load_pretrain("mcq_pretrain_9x14")
    random.seed(5)                                             # make_golden.pretrain_case's shuffle seed
    assert coverage_patterns(g["map"], shuffle=True) == [tuple(int(v) for v in row) for row in g["order"]]
