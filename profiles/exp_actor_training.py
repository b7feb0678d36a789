"""Curriculum training of the unified model on the GPU (batched) and the reference's acceptance band:
python profiles/exp_actor_training.py [batch] [rounds] [sync_every]"""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from ffm_b200 import workloads
from ffm_b200 import unified_training as ut
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sync_every = int(sys.argv[3]) if len(sys.argv) > 3 else 8
m = workloads.room_map(12, 12)
sff = workloads.sff_room(m, "neumann")
exit_pos = tuple(int(v) for v in np.argwhere(m == 3)[0])
t0 = time.time()
V, hc = ut.train_critic(m, sff, exit_pos, batch=batch, rounds=rounds, sync_every=sync_every)
torch.cuda.synchronize(); t1 = time.time()
H, V2, ha = ut.train_actor(m, sff, exit_pos, V, batch=batch, rounds=rounds, sync_every=sync_every)
torch.cuda.synchronize(); t2 = time.time()
res = {}
for N in (10, 30, 50, 70, 90):
    steps, frac = ut.evaluate_trained(m, sff, exit_pos, H, N, episodes=512)
    res[N] = dict(mean=float(steps.mean()), band=frac, lo=int(steps.min()), hi=int(steps.max()))
configs = ut.curriculum(m, exit_pos)
print(json.dumps(dict(batch=batch, rounds=rounds, sync_every=sync_every, configs=len(configs), episodes=len(configs) * rounds * batch,
                      critic_s=t1 - t0, actor_s=t2 - t1, V=len(V), H=len(H), eval=res,
                      actor_last=[h for h in ha[-3:]])))
