"""Host logic of the batched training drivers (no GPU): curricula and schedules follow the reference's drivers."""
import numpy as np

from oracle import assets


def test_curriculum_and_epsilon_schedule_follow_the_drivers():
    from ffm_b200 import unified_training as ut
    m = assets.room_map(12, 12)
    cfgs = ut.curriculum(m, (0, 6))
    assert ut.RADIUS_LIST == [3, 5, 7, 9, 11, 13, 15] and ut.N_LIST == [1, 10, 20, 30, 40, 50, 60, 70, 80, 90]
    assert cfgs[0] == (3, 1) and cfgs[-1] == (15, 90)
    assert all(n <= ut.count_available_cells(m, (0, 6), r) for r, n in cfgs)
    assert (3, 10) not in cfgs and ut.count_available_cells(m, (0, 6), 3) == 9          # skipped like :221-226
    assert ut.count_available_cells(m, (0, 6), 15) == 100



def test_mcq_schedule_follows_the_driver():
    """run_coverage_pretrain_and_training.py:26-46."""
    from ffm_b200.mcq_training import compute_agent_count, compute_beta
    assert [compute_agent_count(e, 100) for e in (0, 49, 50, 499, 500, 1199)] == [10, 10, 20, 100, 100, 100]
    assert compute_agent_count(0, 5) == 1
    assert compute_beta(0) == 1.0 and compute_beta(50) == 1.0 and abs(compute_beta(350) - 0.5) < 1e-12 and compute_beta(651) == 0.0
