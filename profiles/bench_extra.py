"""Secondary measurements (not the driver's bench contract): SFF generation sweep (C5), discounted-returns
kernel (HBM-bound), unified-model rollouts / batched learning (C4), single-step API latency.
One JSON object per line.  Run on the GPU box:  python profiles/bench_extra.py"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffm_b200 import BatchSim, UnifiedSim, workloads   # noqa: E402
from ffm_b200.sff import generate_sff                  # noqa: E402
from ffm_b200.sim import rollout_returns               # noqa: E402
from ffm_b200.sharding import BatchedLearner           # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def sff_sweep():
    maps = np.stack([workloads.obstacle_map_c5(1024, 1024, index=i) for i in range(16)])
    dm = torch.from_numpy(maps).cuda()
    for mode in ("L1", "L2", "bfs4", "bfs8", "dijkstra8"):
        out = {}
        def run():
            out["r"] = generate_sff(dm, mode, np.float32, return_rounds=True)
        ms = timed(run, warm=1, reps=3)
        cells = maps.size
        print(json.dumps({"bench": "sff_generate", "mode": mode, "maps": len(maps), "shape": "1024x1024", "ms": ms,
                          "maps_per_s": len(maps) / ms * 1e3, "cells_per_s": cells / ms * 1e3, "rounds": out["r"][1],
                          "algorithmic_GBps": cells * 5 / ms / 1e6, "hbm_frac": cells * 5 / ms / 1e6 / PEAK}))


def returns():
    B, T, N = 256, 512, 1024
    r = torch.randn(B, T, N, device="cuda")
    L = torch.randint(0, T + 1, (B, N), device="cuda", dtype=torch.int32)
    ms = timed(lambda: rollout_returns(r, L, 0.99))
    bytes_alg = float((L.clamp(max=T).sum().item()) * 4 + B * T * N * 8)      # rewards read inside the paths + returns written
    print(json.dumps({"bench": "rollout_returns", "shape": [B, T, N], "ms": ms, "algorithmic_GBps": bytes_alg / ms / 1e6,
                      "hbm_frac": bytes_alg / ms / 1e6 / PEAK, "bytes": bytes_alg}))
    L = torch.full((B, N), T, device="cuda", dtype=torch.int32)
    ms = timed(lambda: rollout_returns(r, L, 0.99))
    bytes_alg = float(B * T * N * 12)
    print(json.dumps({"bench": "rollout_returns_full_paths", "shape": [B, T, N], "ms": ms, "algorithmic_GBps": bytes_alg / ms / 1e6,
                      "hbm_frac": bytes_alg / ms / 1e6 / PEAK, "bytes": bytes_alg}))


def unified_c4():
    m = workloads.room_map(12, 12)
    sff = workloads.sff_room(m, "neumann")
    P = dict(k_S=10, k_D=1, k_A=10, alpha_v=0.01, alpha_h=0.1, gamma=0.99, exit_reward=100.0, step_penalty=-1.0,
             collision_penalty=-1.0, neighborhood="neumann", block_size=1, epsilon=0.1)
    B, N = 4096, 50
    for mode, learn in (("critic_only", "none"), ("critic_only", "batched"), ("both", "batched")):
        sim = UnifiedSim(m, sff, B, N, mode=mode, learn=learn, params=P, seed=3)
        learner = BatchedLearner(sim, distributed=False) if learn == "batched" else None
        n = np.full(B, N, np.int32)
        def run():
            sim.place(n, exit_pos=(0, 6), radius=15)
            sim.rollout(300)
            if learner:
                learner.sync()
        ms = timed(run, warm=2, reps=3)
        steps, ped = sim.counters()
        print(json.dumps({"bench": "unified_c4", "mode": mode, "learn": learn, "episodes": B, "N": N, "ms": ms,
                          "ped_steps_per_s": float(ped.sum()) / ms * 1e3, "episodes_per_s": B / ms * 1e3,
                          "mean_steps": float(steps.mean())}))
        sim.close()
    # the reference's sequential semantics, one episode at a time (B = 1, FFM_LEARN_EXACT)
    sim = UnifiedSim(m, sff, 1, N, mode="both", learn="exact", params=P, seed=3)
    t0 = time.perf_counter(); tot = 0
    for ep in range(50):
        sim.set_episode_base(ep)
        sim.place(N, exit_pos=(0, 6), radius=15)
        sim.rollout(300)
        tot += int(sim.counters()[1][0])
    dt = time.perf_counter() - t0
    print(json.dumps({"bench": "unified_exact_sequential", "mode": "both", "episodes": 50, "N": N, "ped_steps_per_s": tot / dt,
                      "episodes_per_s": 50 / dt, "note": "reference actor_only measured at ~125 ped-steps/s (SURVEY.md section 6)"}))


def step_latency():
    m = workloads.room_map(50, 50)
    sff = workloads.sff_room(m, "neumann").astype(np.float64)
    sim = BatchSim(m, sff, 1, 100, {"neighborhood": "neumann"}, seed=1)
    sim.place(100)
    ms = timed(lambda: sim.rollout(1), warm=5, reps=20)
    print(json.dumps({"bench": "single_step_api", "map": "50x50", "N": 100, "ms_per_step_launch": ms}))


if __name__ == "__main__":
    torch.cuda.set_device(0)
    which = sys.argv[1:] or ["sff", "returns", "unified", "step"]
    if "sff" in which: sff_sweep()
    if "returns" in which: returns()
    if "unified" in which: unified_c4()
    if "step" in which: step_latency()
