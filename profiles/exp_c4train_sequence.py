"""Why is the c4train pipeline slower inside the default bench sequence?  Times it stand-alone, after a C2 rollout handle, and
with large torch buffers alive."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from ffm_b200 import BatchSim, workloads
from ffm_b200 import unified_training as ut
m = workloads.room_map(12, 12); sff = workloads.sff_room(m, "neumann"); exit_pos = (0, 6)
cfgs = ut.curriculum(m, exit_pos)
def pipe(tag):
    torch.cuda.synchronize(); t = time.time()
    V, _ = ut.train_critic(m, sff, exit_pos, configs=cfgs, batch=256, rounds=2, seed=1)
    torch.cuda.synchronize(); t1 = time.time()
    H, _, _ = ut.train_actor(m, sff, exit_pos, V, configs=cfgs, batch=256, rounds=2, seed=2)
    torch.cuda.synchronize(); t2 = time.time()
    print(f"{tag}: critic {1e3*(t1-t):.0f} ms, actor {1e3*(t2-t1):.0f} ms")
pipe("cold"); pipe("alone"); pipe("alone")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush.fill_(1)
pipe("with 256 MB torch buffer")
m2 = workloads.room_map(64, 64); s2 = workloads.sff_room(m2, "moore")
big = BatchSim(m2, s2, 4096, 1024, {"k_S": 3, "k_D": 0}, seed=1, track_dff=False)
big.set_positions(workloads.place(m2, 1024, 4096, 0, 1), np.full(4096, 1024, np.int32)); big.rollout(4096); torch.cuda.synchronize()
pipe("after a C2 handle (alive)")
big.close(); pipe("after closing it")
pin = torch.empty((256, 1 << 20, 2), dtype=torch.int16).pin_memory()
pipe("with 1 GB pinned host memory")
