"""The NumPy restatement of the legacy 13-cell models (oracle/legacy_numpy.py) against fixtures written by the UNMODIFIED
reference classes under keyed draws (oracle/make_golden.py::legacy_all)."""
import numpy as np
import pytest

from helpers import LEGACY_AC_FIXTURES, LEGACY_ACTOR_FIXTURES, load_legacy
from oracle import legacy_numpy
from oracle.inject import PhiloxSource


@pytest.mark.parametrize("name", LEGACY_AC_FIXTURES)
def test_ac_oracle_reproduces_reference(name):
    g = load_legacy(name)
    V, default = {}, 0.0
    for ep, e in enumerate(g["ep"]):
        o = legacy_numpy.AcOracle(g["map"], g["sff"], e["pos0"], g["params"], PhiloxSource(int(g["seed"]), ep), v_table=V, v_default=default)
        traj = o.run(max_steps=int(g["max_steps"]))
        assert len(traj) == len(e["traj"])
        for a, b in zip(traj, e["traj"]):
            assert np.array_equal(a, b)
        V = o.V
        if ep == int(g["set_v_after"]):
            default = -1.0
    keys = np.array(sorted(V), np.uint64)
    assert np.array_equal(keys, g["v_keys"])
    assert np.array_equal(np.array([V[int(k)] for k in keys]), g["v_vals"])       # bit for bit
    assert np.array_equal(o.dff, g["final_dff"])


def test_key_roundtrip():
    st = ((0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 1), (3, 2))
    assert legacy_numpy.key_to_state(legacy_numpy.state_to_key(st, 7), 7) == st


@pytest.mark.parametrize("name", LEGACY_ACTOR_FIXTURES)
def test_actor_only_oracle_reproduces_reference(name):
    g = load_legacy(name)
    V, Ht = {}, {}
    for ep, e in enumerate(g["ep"]):
        o = legacy_numpy.ActorOnlyOracle(g["map"], g["sff"], e["pos0"], g["params"], PhiloxSource(int(g["seed"]), ep),
                                         v_table=V, h_table=Ht, epsilon=float(g["eps"]))
        traj = o.run(max_steps=int(g["max_steps"]))
        assert len(traj) == len(e["traj"])
        for t, (a, b) in enumerate(zip(traj, e["traj"])):
            assert np.array_equal(a, b), (ep, t)
        V, Ht = o.V, o.H
    vk = np.array(sorted(V), np.uint64)
    hk = np.array(sorted(Ht), np.uint64)
    assert np.array_equal(vk, g["v_keys"]) and np.array_equal(hk, g["h_keys"])
    assert np.array_equal(np.array([V[int(k)] for k in vk]), g["v_vals"])
    assert np.array_equal(np.array([Ht[int(k)] for k in hk]).reshape(len(hk), -1), g["h_vals"])
    assert np.array_equal(o.dff, g["final_dff"])


def test_product_key_helpers_match_the_checker():
    """ffm_b200.legacy.state_to_key / key_to_state (what the drop-ins use to rebuild the reference's pickled keys) against the
    checker's, on the keys of a reference fixture."""
    from ffm_b200 import legacy as product
    g = load_legacy("legacy_ac_moore_f64")
    nby = (g["map"].shape[1] + g["params"]["block_size"] - 1) // g["params"]["block_size"]
    for k in g["v_keys"][:500]:
        st = legacy_numpy.key_to_state(int(k), nby)
        assert product.key_to_state(k, nby) == st and product.state_to_key(st, nby) == int(k)
        assert st[0][4] == 1 and len(st[0]) == 13            # the centre cell of a state is the pedestrian itself
