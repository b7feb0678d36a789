"""B = 1 sequential-exact learning (what the drop-in classes run): ped-steps/s.  python profiles/exp_exact_mode.py"""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from ffm_b200 import UnifiedSim, workloads
m = workloads.room_map(12, 12); sff = workloads.sff_room(m, "neumann")
P = dict(k_S=10, k_D=1, k_A=10, alpha_v=0.01, alpha_h=0.1, gamma=0.99, exit_reward=100.0, step_penalty=-1.0, collision_penalty=-1.0,
         neighborhood="neumann", block_size=1, epsilon=0.1)
for mode in ("critic_only", "both"):
    N = 50
    sim = UnifiedSim(m, sff, 1, N, mode=mode, learn="exact", params=P, seed=3)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); tot = 0
        for ep in range(50):
            sim.set_episode_base(ep)
            sim.place(N, exit_pos=(0, 6), radius=15)
            sim.rollout(300)
            tot += int(sim.counters()[1][0])
        dt = time.perf_counter() - t0
    print(json.dumps({"mode": mode, "ped_steps_per_s": tot / dt, "episodes_per_s": 50 / dt}))
