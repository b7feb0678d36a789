"""Run under torchrun with 2+ GPUs:  python -m torch.distributed.run --nproc-per-node 2 tests/multi_gpu_check.py
Checks (1) sharded C2-geometry episodes equal the same global episodes run on one handle, (2) the
batched learner leaves identical tables on every rank and equals a single-rank run of all episodes."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import bench
    from ffm_b200 import BatchSim, UnifiedSim
    from ffm_b200.sharding import BatchedLearner, shard_range
    from helpers import pack_positions
    from oracle import assets

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rank, ws = dist.get_rank(), dist.get_world_size()

    # (1) independent episodes, keyed by global id
    m = bench.room_map(64, 64); sff = bench.sff_room(m, "moore")
    total, N, seed = 64, 1024, 11
    first, count = shard_range(total)
    params = {"k_S": 3, "k_D": 1, "neighborhood": "moore"}
    mine = BatchSim(m, sff, count, N, params, seed=seed, episode_base=first, device=local)
    mine.set_positions(bench.place(m, N, count, first, seed), np.full(count, N, np.int32))
    mine.rollout(4096)
    steps = torch.from_numpy(mine.counters()[0]).cuda()
    gathered = [torch.zeros_like(steps) for _ in range(ws)]
    dist.all_gather(gathered, steps)
    if rank == 0:
        full = BatchSim(m, sff, total, N, params, seed=seed, episode_base=0, device=local)
        full.set_positions(bench.place(m, N, total, 0, seed), np.full(total, N, np.int32))
        full.rollout(4096)
        assert np.array_equal(torch.cat(gathered).cpu().numpy(), full.counters()[0]), "sharded != single handle"
        print("sharded episodes == single handle: OK")

    # (2) batched learner: identical tables on all ranks == one rank running all episodes
    m = assets.room_map(12, 12); sff = assets.sff_norm_min(m, "L1", np.float32)
    P = dict(k_S=10, k_D=1, k_A=10, alpha_v=0.01, alpha_h=0.1, gamma=0.99, exit_reward=100.0, step_penalty=-1.0,
             collision_penalty=-1.0, neighborhood="neumann", block_size=1, epsilon=0.1)
    free = np.argwhere(m == 0)
    rng = np.random.RandomState(5)
    E, Np = 64, 12
    pos = [free[rng.choice(len(free), Np, replace=False)] for _ in range(E)]
    first, count = shard_range(E)
    sim = UnifiedSim(m, sff, count, Np, mode="both", learn="batched", params=P, seed=7, episode_base=first, device=local)
    learner = BatchedLearner(sim)
    for r in range(3):
        sim.set_episode_base(first + r * E)
        learner.round(*pack_positions(pos[first:first + count], Np), 100)
    V, vs, H, hs = sim.get_tables()
    tv = torch.from_numpy(V).cuda()
    lo, hi = tv.clone(), tv.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks hold different V tables"
    if rank == 0:
        one = UnifiedSim(m, sff, E, Np, mode="both", learn="batched", params=P, seed=7, episode_base=0, device=local)
        l1 = BatchedLearner(one, distributed=False)
        for r in range(3):
            one.set_episode_base(r * E)
            l1.round(*pack_positions(pos, Np), 100)
        V1, vs1, H1, hs1 = one.get_tables()
        assert np.array_equal(vs, vs1) and np.array_equal(hs, hs1)
        assert np.allclose(V, V1, rtol=1e-9, atol=1e-12) and np.allclose(H, H1, rtol=1e-9, atol=1e-12)   # atomics reorder sums
        print("batched learner over", ws, "ranks == single rank: OK")
    # (2b) the same with the exchange overlapped with the next rollout (staleness 1): ranks still agree bit for bit
    sim2 = UnifiedSim(m, sff, count, Np, mode="both", learn="batched", params=P, seed=7, episode_base=first, device=local)
    l2 = BatchedLearner(sim2, overlap=True)
    for r in range(3):
        sim2.set_episode_base(first + r * E)
        l2.round(*pack_positions(pos[first:first + count], Np), 100, sync_every=25)
    tv = torch.from_numpy(sim2.get_tables()[0]).cuda()
    lo, hi = tv.clone(), tv.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "overlapped exchange: ranks hold different V tables"
    if rank == 0:
        print("overlapped exchange: identical tables on", ws, "ranks: OK")

    # (3) Monte-Carlo Q-learning: returns exchanged BY KEY (hash-table slots are a local matter) -> identical dicts on all
    #     ranks == one rank running all episodes
    from ffm_b200 import McqSim
    from ffm_b200.mcq_training import McqBatchedLearner
    sffd = assets.sff_norm_min(m, "L1", np.float64)
    QP = {"max_steps": 60, "step_penalty": 0.02, "stop_penalty": 0.1, "collision_penalty": 0.5}
    msim = McqSim(m, sffd, count, Np, learn="batched", params=QP, seed=21, episode_base=first, alpha=0.2, gamma=0.97, device=local,
                  q_log2_capacity=15)
    ml = McqBatchedLearner(msim)
    for r, beta in enumerate((1.0, 0.5, 0.0)):
        msim.set_episode_base(first + r * E)
        msim.set_beta(beta)
        msim.set_positions(*pack_positions(pos[first:first + count], Np))
        msim.rollout(61)
        ml.sync()
    ids, rows = msim.get_q()
    n_ids = torch.tensor([len(ids)], device="cuda")
    all_n = [torch.zeros_like(n_ids) for _ in range(ws)]
    dist.all_gather(all_n, n_ids)
    assert len({int(x.item()) for x in all_n}) == 1, "ranks hold Q dicts of different size"
    ti, tr = torch.from_numpy(ids).cuda(), torch.from_numpy(rows).cuda()
    for t in (ti, tr):
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks hold different Q dicts"
    if rank == 0:
        one = McqSim(m, sffd, E, Np, learn="batched", params=QP, seed=21, episode_base=0, alpha=0.2, gamma=0.97, device=local,
                     q_log2_capacity=15)
        l1 = McqBatchedLearner(one, distributed=False)
        for r, beta in enumerate((1.0, 0.5, 0.0)):
            one.set_episode_base(r * E)
            one.set_beta(beta)
            one.set_positions(*pack_positions(pos, Np))
            one.rollout(61)
            l1.sync()
        ids1, rows1 = one.get_q()
        assert np.array_equal(ids, ids1) and np.allclose(rows, rows1, rtol=1e-5, atol=1e-5)
        print("MC-Q batched learner (by-key exchange) over", ws, "ranks == single rank: OK,", len(ids), "rows")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
