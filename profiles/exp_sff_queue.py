"""SFF sweep timing: python profiles/exp_sff_queue.py [maps] -- ms per mode and tile visits."""
import json, os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from ffm_b200 import workloads
from ffm_b200.sff import generate_sff
M = int(sys.argv[1]) if len(sys.argv) > 1 else 16
maps = np.stack([workloads.obstacle_map_c5(1024, 1024, index=i) for i in range(M)])
dm = torch.from_numpy(maps).cuda()
for mode in ("L1", "bfs4", "bfs8", "dijkstra8"):
    _, visits = generate_sff(dm, mode, np.float32, return_rounds=True)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); generate_sff(dm, mode, np.float32); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(json.dumps(dict(tag=os.environ.get("FFM_SFF_CTAS_PER_SM", "max"), mode=mode, maps=M, ms=float(np.median(ts)), visits=visits,
                          cells_per_s=maps.size / np.median(ts) * 1e3)))
