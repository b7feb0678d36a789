"""C3 A/B: python profiles/quickbench_c3.py [episodes] -- prints ped-steps/s of the configuration the environment
selects (FFM_KERNEL=ped | FFM_CLUSTER=2|4|8, FFM_THREADS=512|1024, FFM_SCORE_GLOBAL=1)."""
import json, os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from ffm_b200 import BatchSim
from ffm_b200.sff import generate_sff
from ffm_b200.workloads import place, rooms_map_c3

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
N = 10000
m = rooms_map_c3(256, 256)
sff = generate_sff(m, "bfs8", np.float32)
pos = place(m, N, B, 0, 0x5EED0003)
n = np.full((B,), N, np.int32)
sim = BatchSim(m, sff, B, N, {"k_S": 3, "k_D": 1, "neighborhood": "moore"}, seed=0x5EED0003)
info = sim.kernel_info()
best = None
for it in range(3):
    sim.set_positions(pos, n)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.rollout(cap); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps, ps = sim.counters()
    best = ms if best is None else min(best, ms)
print(json.dumps(dict(tag=os.environ.get("TAG", ""), B=B, cap=cap, ms=best, ped_steps=int(ps.sum()), rate=float(ps.sum()) / best * 1e3,
                      steps_mean=float(steps.mean()), info=info)))
