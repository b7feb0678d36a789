"""BASELINE config 4 across GPUs: batched actor-critic learning of the unified model, episodes sharded over
ranks, one NCCL all-reduce of the table deltas (dV, dN, dH + key flags) per sync.
    python -m torch.distributed.run --nproc-per-node N profiles/bench_c4_multigpu.py
Prints one JSON line (rank 0): episodes/s, ped-steps/s, all-reduce share of the round time."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffm_b200 import UnifiedSim, workloads          # noqa: E402
from ffm_b200.sharding import BatchedLearner, shard_range   # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank = dist.get_rank() if ws > 1 else 0
    m = workloads.room_map(12, 12)
    sff = workloads.sff_room(m, "neumann")
    P = dict(k_S=10, k_D=1, k_A=10, alpha_v=0.01, alpha_h=0.1, gamma=0.99, exit_reward=100.0, step_penalty=-1.0,
             collision_penalty=-1.0, neighborhood="neumann", block_size=1, epsilon=0.1)
    B, N, rounds = 4096, 50, 6                      # episodes per GPU per sync (weak scaling), MODEL_PARAMS of the drivers
    first, _ = shard_range(B * ws, rank, ws)
    out = {}
    for mode in ("critic_only", "both"):
        sim = UnifiedSim(m, sff, B, N, mode=mode, learn="batched", params=P, seed=11, episode_base=first, device=local)
        learner = BatchedLearner(sim)
        n = np.full(B, N, np.int32)
        def one_round(r):
            sim.set_episode_base(first + r * B * ws)
            sim.place(n, exit_pos=(0, 6), radius=15)
            sim.rollout(300)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); learner.sync(); e1.record()
            return e0, e1
        for r in range(2):
            one_round(r)
        torch.cuda.synchronize()
        if ws > 1:
            dist.barrier()
        t0 = time.perf_counter()
        evs = [one_round(2 + r) for r in range(rounds)]
        torch.cuda.synchronize()
        if ws > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        steps, ped = sim.counters()
        tot = torch.tensor([float(ped.sum())], device="cuda", dtype=torch.float64)
        if ws > 1:
            dist.all_reduce(tot)
        sync_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        out[mode] = dict(episodes_per_s=B * ws * rounds / dt, ped_steps_per_s=tot.item() * rounds / dt,
                         round_ms=dt / rounds * 1e3, sync_ms=sync_ms, mean_steps=float(steps.mean()))
        V, vs, H, hs = sim.get_tables()
        chk = torch.tensor([float(V.sum()), float(vs.sum())], device="cuda", dtype=torch.float64)
        if ws > 1:
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            out[mode]["tables_identical_across_ranks"] = bool(torch.equal(lo, hi))
        sim.close()
    if rank == 0:
        print(json.dumps({"bench": "c4_batched_learning", "n_gpus": ws, "episodes_per_gpu_per_sync": B, "N": N, **out}))
    if ws > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
