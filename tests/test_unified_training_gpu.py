"""Curriculum training of the unified model on the GPU (ffm_b200/unified_training.py: batched forms of
run_unified_critic_training.py / run_unified_actor_training.py) and the reference's acceptance criterion for the result:
with the trained H frozen (run_trained_ffm.py) the evacuation time of N pedestrians lies in [2N - 1, 2N + 14]
(analyze_steps_by_n.py:109-110; 95-100 % of the episodes in the reference's steps_range_statistics_all.csv)."""
import numpy as np
import pytest

from oracle import assets

pytestmark = pytest.mark.gpu


def test_actor_training_reaches_the_reference_band(cuda_device):
    from ffm_b200 import unified_training as ut
    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    exit_pos = (0, 6)
    V, hist_c = ut.train_critic(m, sff, exit_pos, batch=256, rounds=4, sync_every=8, seed=10)
    assert len(V) > 5000 and all(np.isfinite(v) for v in V.values())
    H, V2, hist_a = ut.train_actor(m, sff, exit_pos, V, batch=256, rounds=4, sync_every=8, seed=11)
    assert len(H) > 5000 and set(V) <= set(V2)
    rows = np.array(list(H.values()))
    assert rows.shape[1] == 5 and np.isfinite(rows).all() and np.abs(rows).max() > 1.0
    for N in (10, 30, 50, 90):
        steps, frac = ut.evaluate_trained(m, sff, exit_pos, H, N, episodes=256, seed=12)
        assert frac >= 0.95, (N, frac, steps.mean())
        assert abs(steps.mean() - (2 * N - 1)) < (6.0 if N == 10 else 3.0), (N, steps.mean())
    # control: the same evaluation with an untrained (empty) H table does not evacuate on that schedule
    steps0, frac0 = ut.evaluate_trained(m, sff, exit_pos, {}, 50, episodes=64, seed=12)
    assert frac0 < 0.5, frac0


def test_graph_replayed_rounds_equal_eager_rounds(cuda_device):
    """The round's launch chain captured in a CUDA graph (epsilon and the episode key read from a device struct,
    ffm_bind_dynamic) trains the same tables as eager launches: same keyed draws, same trajectories; the sums differ only by
    the order of the float64 atomics."""
    from ffm_b200 import unified_training as ut
    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    cfgs = [(5, 10), (9, 30), (15, 50)]
    out = []
    for use_graph in (False, True):
        V, _ = ut.train_critic(m, sff, (0, 6), configs=cfgs, batch=64, rounds=3, sync_every=8, seed=21, use_graph=use_graph)
        H, V2, _ = ut.train_actor(m, sff, (0, 6), V, configs=cfgs, batch=64, rounds=3, sync_every=8, seed=22, use_graph=use_graph)
        out.append((V, H))
    (Ve, He), (Vg, Hg) = out
    assert set(Ve) == set(Vg) and set(He) == set(Hg) and len(Ve) > 500
    assert np.allclose([Ve[k] for k in Ve], [Vg[k] for k in Ve], rtol=1e-9, atol=1e-9)
    assert np.allclose([He[k] for k in He], [Hg[k] for k in He], rtol=1e-7, atol=1e-7)
