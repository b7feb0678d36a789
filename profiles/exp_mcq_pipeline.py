"""MC-Q pipeline timing on the reference's default configuration (50x50 room, L1 SFF, N = 100, max_steps 500):
coverage pretrain (all patterns, one launch) + a stretch of the training schedule.  python profiles/exp_mcq_pipeline.py [entries] [batch]"""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from ffm_b200 import workloads
from ffm_b200.mcq_training import coverage_patterns, coverage_pretrain, run_training
entries = int(sys.argv[1]) if len(sys.argv) > 1 else 60
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
m = workloads.room_map(50, 50)
sff = workloads.sff_room(m, "neumann").astype(np.float64)
params = {"k_S": 3, "k_D": 1, "diffuse": 0.2, "decay": 0.2, "max_steps": 500, "alpha": 0.1, "gamma": 0.99}
order = coverage_patterns(m, shuffle=False)
for it in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    Q, steps = coverage_pretrain(m, sff, params, {}, order=order, seed=1, return_steps=True)
    torch.cuda.synchronize(); t1 = time.time()
    Q2, mean_steps = run_training(m, sff, params, full_N=100, shared_Q=Q, num_episodes=entries, batch=batch, seed=2)
    torch.cuda.synchronize(); t2 = time.time()
    print(json.dumps(dict(patterns=len(order), pretrain_s=t1 - t0, pretrain_ca_steps=int(steps.sum()), q_after_pretrain=len(Q),
                          entries=entries, batch=batch, train_s=t2 - t1, q_after=len(Q2), mean_steps=mean_steps[:3] + mean_steps[-3:])))
