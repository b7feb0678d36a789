import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from ffm_b200 import workloads
from ffm_b200.sff import generate_sff
m3 = workloads.rooms_map_c3()
for mode in ("bfs4", "bfs8", "dijkstra8"):
    t = time.time(); out, r = generate_sff(m3, mode, np.float32, return_rounds=True); print("c3", mode, r, "%.3fs" % (time.time() - t), float(out[np.isfinite(out)].max()), flush=True)
m5 = workloads.obstacle_map_c5(1024, 1024, index=0)
for mode in ("bfs4", "dijkstra8"):
    t = time.time(); out, r = generate_sff(torch.from_numpy(m5).cuda(), mode, np.float32, return_rounds=True); torch.cuda.synchronize(); print("c5", mode, r, "%.3fs" % (time.time() - t), flush=True)
maps = np.stack([workloads.obstacle_map_c5(200, 160, index=i, n_exits=4) for i in range(2)])
t = time.time(); out, r = generate_sff(maps, "bfs8", np.float32, return_rounds=True); print("batch", r, "%.3fs" % (time.time() - t), flush=True)
