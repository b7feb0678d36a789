import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from ffm_b200 import workloads
from ffm_b200 import BatchSim
B, N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 1024
m = workloads.room_map(64, 64)
sff = workloads.sff_room(m, "moore")
rng = np.random.RandomState(0)
free = np.argwhere(m == 0)
pos = np.stack([free[rng.permutation(len(free))[:N]] for _ in range(B)]).astype(np.int32)
n = np.full((B,), N, dtype=np.int32)
for track in (False, True):
    sim = BatchSim(m, sff, B, N, {"k_S": 3, "k_D": 0}, seed=1, track_dff=track)
    print(sim.kernel_info())
    for it in range(3):
        sim.set_positions(pos, n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); sim.rollout(4096); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        steps, ps = sim.counters()
        print(f"track_dff={track} B={B} ms={ms:.1f} ped_steps={ps.sum():.3e} rate={ps.sum()/ms*1e3:.3e}/s steps mean={steps.mean():.0f} max={steps.max()}")
    sim.close()
