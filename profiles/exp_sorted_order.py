import sys
sys.path.insert(0, '.')
import numpy as np, torch
from ffm_b200 import BatchSim, workloads
B, N = 3552, 1024
m = workloads.room_map(64, 64); sff = workloads.sff_room(m, "moore")
pos = workloads.place(m, N, B, 0, 1)
n = np.full((B,), N, dtype=np.int32)
def run(p, tag):
    sim = BatchSim(m, sff, B, N, {"k_S": 3, "k_D": 0}, seed=1, track_dff=False)
    for it in range(3):
        sim.set_positions(p, n); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); sim.rollout(4096); e1.record(); torch.cuda.synchronize()
        steps, ps = sim.counters()
    print(tag, "ms=%.1f rate=%.3e steps=%.0f" % (e0.elapsed_time(e1), ps.sum() / e0.elapsed_time(e1) * 1e3, steps.mean()))
run(pos, "random order ")
key = pos[:, :, 0].astype(np.int64) * 64 + pos[:, :, 1]
order = np.argsort(key, axis=1)
spos = np.take_along_axis(pos, order[:, :, None], axis=1)
run(np.ascontiguousarray(spos), "row-major sort")
# sort by distance to exit (Linf), then row-major: peds near the exit get the low indices
d = np.maximum(np.abs(pos[:, :, 0] - 0), np.abs(pos[:, :, 1] - 32)).astype(np.int64)
order = np.argsort(d * 4096 + key, axis=1)
run(np.ascontiguousarray(np.take_along_axis(pos, order[:, :, None], axis=1)), "by distance  ")
