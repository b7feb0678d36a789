"""Pins oracle/unified_numpy.py to fixtures produced by the unmodified reference classes
(model/ffm_unified.py in its three learning modes, model/ffm_trained_core.py): trajectories of
multi-episode learning runs, the final V and H tables (keys and float64 values), the final DFF."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, UNIFIED_FIXTURES, load_unified
from oracle import unified_numpy
from oracle.inject import PhiloxSource


def _initial_v(g):
    if not g["v_from"]:
        return None
    z = np.load(os.path.join(GOLDEN, g["v_from"] + ".npz"))
    helper = unified_numpy.UnifiedOracle(g["map"], g["sff"], np.zeros((0, 2)), "critic_only", g["params"])
    return {helper.id_to_key(i): float(v) for i, v in zip(z["v_ids"], z["v_vals"])}


@pytest.mark.parametrize("name", UNIFIED_FIXTURES)
def test_unified_oracle_reproduces_reference(name):
    g = load_unified(name)
    o = unified_numpy.UnifiedOracle(g["map"], g["sff"], g["ep"][0]["pos0"], g["mode"], g["params"], None, v_table=_initial_v(g))
    o.epsilon = float(g["eps"])
    for ep, E in enumerate(g["ep"]):
        o.positions, o.t, o.source = E["pos0"].copy(), 0, PhiloxSource(int(g["seed"]), ep)
        o.dff[:] = 0                                            # reset() (ffm_unified.py:800-812)
        r = o.run(max_steps=int(g["max_steps"]))
        assert r["steps"] == int(g["steps"][ep])
        for t, want in enumerate(E["traj"]):
            assert np.array_equal(r["traj"][t], want), (ep, t)
    assert np.array_equal(np.flatnonzero(o.v_seen), g["v_ids"])
    assert np.array_equal(o.V[g["v_ids"]], g["v_vals"])         # bit-exact float64
    assert np.array_equal(np.flatnonzero(o.h_seen), g["h_ids"])
    assert np.array_equal(o.H[g["h_ids"]], g["h_vals"])
    assert np.array_equal(o.dff.view(np.uint32), g["final_dff"].view(np.uint32))


def test_trained_oracle_reproduces_reference():
    z = np.load(os.path.join(GOLDEN, "trained_12x12.npz"))
    params = json.loads(str(z["params"]))
    hz = np.load(os.path.join(GOLDEN, str(z["h_from"]) + ".npz"))
    helper = unified_numpy.UnifiedOracle(z["map"], z["sff"], np.zeros((0, 2)), "trained", params)
    htab = {helper.id_to_key(i): list(row) for i, row in zip(hz["h_ids"], hz["h_vals"])}
    o = unified_numpy.UnifiedOracle(z["map"], z["sff"], z["pos0"], "trained", params, PhiloxSource(int(z["seed"]), 0), h_table=htab)
    r = o.run(max_steps=int(z["max_steps"]))
    assert r["steps"] == int(z["steps"])
    offs = np.concatenate([[0], np.cumsum(z["counts"])])
    for t in range(r["steps"]):
        assert np.array_equal(r["traj"][t], z["traj"][offs[t]:offs[t + 1]].astype(np.int64)), t
    assert np.array_equal(o.dff.view(np.uint32), z["final_dff"].view(np.uint32))


def test_encode_state_known_answers():
    """Rank code semantics of _encode_state (ffm_unified.py:188-269) on a hand-built neighbourhood."""
    m = np.zeros((7, 7), np.uint8)
    m[0, :] = m[-1, :] = m[:, 0] = m[:, -1] = 2
    sm = m.copy()
    assert unified_numpy.encode_ranks(3, 3, sm) == (3, 3, 3, 3)
    assert unified_numpy.encode_ranks(1, 1, sm) == (0, 3, 0, 3)        # walls up/left
    assert unified_numpy.encode_ranks(2, 3, sm) == (2, 3, 3, 3)        # wall two steps up
    sm[2, 3] = 1
    assert unified_numpy.encode_ranks(3, 3, sm)[0] == 0                # pedestrian one step up
    sm[2, 3] = 0; sm[2, 4] = 1
    assert unified_numpy.encode_ranks(3, 3, sm) == (1, 3, 3, 1)        # diagonal-forward pedestrian: up and right
    sm[2, 4] = 0; sm[1, 3] = 1
    assert unified_numpy.encode_ranks(3, 3, sm)[0] == 2                # pedestrian two steps up
    sm[1, 3] = 0; m2 = sm.copy(); m2[2, 3] = 3
    assert unified_numpy.encode_ranks(3, 3, m2)[0] == 3                # an exit does not block
