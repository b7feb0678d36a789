"""One frozen-table batch of the legacy TD-critic kernel (the launch profiles/r2_ncu_full_legacy_ac.json was captured from).
Usage: python profiles/quickbench_legacy.py [episodes]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffm_b200.legacy import LegacySim                     # noqa: E402
from ffm_b200.workloads import place, room_map, sff_room  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
m = room_map(50, 50)
sff = sff_room(m, "neumann").astype(np.float64)
params = {"k_S": 10, "k_D": 1, "alpha_v": 0.01, "gamma": 0.99, "exit_reward": 100.0, "step_penalty": -1.0, "collision_penalty": -1.0,
          "neighborhood": "neumann", "block_size": 5}
sim = LegacySim(m, sff, B, 100, model="ac", learn="none", params=params, seed=1)
pos = place(m, 100, B, 0, 2)
for it in range(3):
    sim.set_positions(pos, np.full((B,), 100, np.int32))
    sim.set_dff(np.zeros((B, 50, 50), np.float32))
    t0 = time.perf_counter()
    sim.rollout(500)
    dt = time.perf_counter() - t0
    ps = int(sim.counters()[1].sum())
    print(f"pass {it}: {ps} ped-steps in {dt * 1e3:.2f} ms (host clock around the synchronous call) = {ps / dt:.3e} ped-steps/s")
