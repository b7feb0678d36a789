"""North-star bar "float move probabilities must match within 1e-6 relative": the rollout kernel's own
probability arithmetic (ffm_move_probs probe) vs the probabilities the UNMODIFIED reference passed to
np.random.choice (fixtures) and vs the oracle mid-episode, float32 and float64 SFF, with and without DFF."""
import numpy as np
import pytest

from helpers import CORE_FIXTURES, load_golden, pack_positions
from oracle import ffm_numpy
from oracle.inject import PhiloxSource

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _compact(row, mask_order):
    return row[mask_order]


@pytest.mark.parametrize("name", CORE_FIXTURES)
def test_first_step_probabilities_match_reference(cuda_device, name):
    from ffm_b200 import BatchSim
    g = load_golden(name)
    N = len(g["pos0"])
    sim = BatchSim(g["map"], g["sff"], 1, N, g["params"], seed=int(g["seed"]), episode_base=int(g["episode"]), track_dff=True)
    sim.set_positions(*pack_positions([g["pos0"].astype(np.int32)], N))
    probs, kind = sim.move_probs()
    checked = 0
    for want, (t, idx, k) in zip(g["probs"], g["probs_meta"]):
        if t != 0:
            break
        got = probs[0, idx]
        got = got[got > 0] if (got > 0).sum() == k else got[:k]
        assert kind[0, idx] == 2 and len(got) == k
        np.testing.assert_allclose(got, want[:k], rtol=RTOL, atol=0)
        checked += 1
    assert checked == int((g["probs_meta"][:, 0] == 0).sum())      # every logged draw of step 0 (none when the room starts full)


@pytest.mark.parametrize("name", ["core_12x12_moore_f32_full", "core_50x50_moore_f64", "core_20x20_moore_params"])
def test_mid_episode_probabilities_match_oracle(cuda_device, name):
    """After 30 steps the DFF is non-trivial: compare every pedestrian's distribution with the oracle's."""
    from ffm_b200 import BatchSim
    g = load_golden(name)
    N = len(g["pos0"])
    sim = BatchSim(g["map"], g["sff"], 1, N, g["params"], seed=int(g["seed"]), episode_base=int(g["episode"]), track_dff=True)
    sim.set_positions(*pack_positions([g["pos0"].astype(np.int32)], N))
    sim.rollout(30)
    probs, kind = sim.move_probs()
    o = ffm_numpy.CoreOracle(g["map"], g["sff"], g["pos0"], g["params"], PhiloxSource(int(g["seed"]), int(g["episode"])))
    for _ in range(30):
        o.step()
    o.probs_log = []
    n_before = o.positions.shape[0]
    o.step()
    assert sim.get_positions()[1][0] == n_before
    seen = set()
    for t, idx, p in o.probs_log:
        got = probs[0, idx]
        got = got[got > 0]
        assert kind[0, idx] == 2 and len(got) == len(p), (idx, got, p)
        np.testing.assert_allclose(got, p, rtol=RTOL, atol=0)
        seen.add(idx)
    assert len(seen) > 10
    assert set(np.flatnonzero(kind[0, :n_before] == 2)) == seen
