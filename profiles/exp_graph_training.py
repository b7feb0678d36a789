"""CUDA-graph replay of the training rounds vs eager launches: python profiles/exp_graph_training.py"""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from ffm_b200 import workloads
from ffm_b200 import unified_training as ut
m = workloads.room_map(12, 12); sff = workloads.sff_room(m, "neumann"); exit_pos = (0, 6)
cfgs = ut.curriculum(m, exit_pos)
for use_graph in (False, True, False, True):
    torch.cuda.synchronize(); t = time.time()
    V, _ = ut.train_critic(m, sff, exit_pos, configs=cfgs, batch=256, rounds=2, seed=1, use_graph=use_graph)
    torch.cuda.synchronize(); t1 = time.time()
    H, _, _ = ut.train_actor(m, sff, exit_pos, V, configs=cfgs, batch=256, rounds=2, seed=2, use_graph=use_graph)
    torch.cuda.synchronize(); t2 = time.time()
    res = {N: ut.evaluate_trained(m, sff, exit_pos, H, N, episodes=256)[1] for N in (10, 50, 90)}
    print(f"graph={use_graph}: critic {1e3*(t1-t):.0f} ms, actor {1e3*(t2-t1):.0f} ms, |V|={len(V)} |H|={len(H)} band={res}")
