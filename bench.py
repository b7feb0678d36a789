#!/usr/bin/env python
"""Benchmark of the FFM hot path: pedestrian-steps/s of batched evacuation episodes (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c2dff|c3|c4|c1|c2traj|sff] [--impl ours|reference]
                    [--no-secondary]

A "step" is one pass of the hot path over one batch: B episodes placed, then rolled out by the
persistent kernel until everybody has left (or the step cap).  Workload `c2` (the configuration the
metric is quoted on): 64x64 single-exit room, SFF only (k_D = 0, DFF not tracked), Moore,
1024 pedestrians/episode, 4096 episodes per GPU, cap 4096 CA steps.  Episodes are keyed by global
episode id, sharded over ranks with no data-path collective (weak scaling).

`value`     ped-steps/s with the initial positions already resident in HBM (CUDA events, max over ranks)
`e2e`       same metric through the public API with HOST buffers: pinned-host positions -> H2D ->
            rollout -> D2H of the per-episode counters, all inside the timed region
`roofline`  dominant kernel: algorithmic bytes / CUDA-event duration.  For the shared-memory-resident rollouts (C1, C2)
            the bound is the SHARED-MEMORY bandwidth, measured live by a micro-benchmark (ffm_measure_smem_bandwidth);
            the HBM view of the same bytes is kept under roofline.hbm.  C3 (fields in L2), the trajectory record and
            the SFF sweep are reported against the measured HBM copy peak (MEASURED_PEAKS.json).
`cpu_baseline` the oracle port timed on this box's host cores on a bounded sample (rank 0, N = 1)
`secondary` (default C2 run only) the other configurations of BASELINE.json, each measured the same way with its own
            clocks / e2e / roofline: c3 (256x256 plan, DFF), c4 (batched TD learning, the NCCL all-reduce of the table
            deltas inside the timed region), c2traj (C2 with the compact trajectory record, D2H of the record in e2e),
            sff (64-map 1024x1024 geodesic sweep)

`--impl reference` times the CPU implementation (oracle port; the Python reference cannot travel to
the GPU box) on all host threads on the same workload and prints the same line with impl=reference.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pedestrian_steps_per_sec"
UNIT = "ped-steps/s"

# algorithmic bytes per pedestrian-step, SURVEY.md section 8(d): own position read 4 B + 9 cells x
# (map/occupancy 1 B + SFF 4 B [+ DFF 4 B]) + position write 4 B + occupancy update 2 B
BYTES_PER_PED_STEP = {False: 55, True: 91}
# unified model (C4): 5 slots x (1 + 4 + 4) B + position 4 + 4 + occupancy 2 + 16 B of state-encoder cell reads + V[s], V[s'] read and
# V[s] written (3 x 8 B)
BYTES_PER_PED_STEP_UNIFIED = 95
C4_PARAMS = dict(k_S=10, k_D=1, k_A=10, alpha_v=0.01, alpha_h=0.1, gamma=0.99, exit_reward=100.0, step_penalty=-1.0,
                 collision_penalty=-1.0, neighborhood="neumann", block_size=1, diffuse=0.2, decay=0.2)   # run_unified_critic_training.py:40-50

WORKLOADS = {
    # name: (H, W, peds/episode, episodes/GPU, step cap, neighbourhood, k_S, k_D, track_dff)
    "c2": dict(h=64, w=64, n=1024, episodes=4096, cap=4096, nbh="moore", k_S=3, k_D=0, track_dff=False,
               desc="C2: 4096 episodes/GPU, 64x64 single-exit room, SFF only, Moore, 1024 peds/episode, to evacuation (cap 4096)"),
    "c2dff": dict(h=64, w=64, n=1024, episodes=4096, cap=4096, nbh="moore", k_S=3, k_D=1, track_dff=True,
                  desc="C2 geometry with DFF on (k_D=1, diffuse=decay=0.2)"),
    "c3": dict(h=256, w=256, n=10000, episodes=148, cap=2048, nbh="moore", k_S=3, k_D=1, track_dff=True, plan="c3",
               desc="C3: 148 episodes/GPU, 256x256 floor plan (3x3 rooms, 4 exits), geodesic SFF, DFF on, Moore, 10000 peds/episode, cap 2048"),
    "c4": dict(h=12, w=12, n=50, episodes=4096, cap=300, nbh="neumann", k_S=10, k_D=1, track_dff=True, model="unified", radius=15,
               desc="C4: batched TD(0) critic learning (run_unified_critic_training.py MODEL_PARAMS), 12x12, N=50 within radius 15, "
                    "4096 episodes/GPU per sync, all-reduce of the table deltas, cap 300"),
    "c1": dict(h=12, w=12, n=100, episodes=4096, cap=4096, nbh="neumann", k_S=3, k_D=1, track_dff=True,
               desc="C1 geometry batched: 12x12 room, neumann, N=100, DFF on"),
    "c2traj": dict(h=64, w=64, n=1024, episodes=1024, cap=4096, nbh="moore", k_S=3, k_D=0, track_dff=False, record=True,
                   desc="C2 geometry, 1024 episodes/GPU, with the compact trajectory record (int16 row/col pairs, 4 B per "
                        "pedestrian-step; what run() collects, ffm_core.py:125 / main.py:44-52)"),
    "c4train": dict(h=12, w=12, train=True, batch=256, rounds=2,
                    desc="C4 pipeline: the unified model's curricula (run_unified_critic_training.py / run_unified_actor_training.py: radius "
                         "3..15 step 2 x N in {1, 10..90}, epsilon 0.2 -> 0.01 per configuration) as batched GPU runs -- critic, then actor on "
                         "the critic's V -- 256 episodes/GPU per round, 2 rounds per configuration, table deltas all-reduced every 8 CA steps"),
    "c5train": dict(h=50, w=50, train="mcq", batch=64, stride=10,
                    desc="C5 pipeline (run_coverage_pretrain_and_training.py on its default configuration: 50x50 room, L1 SFF float64, N = 100, "
                         "max_steps 500): coverage pretrain -- all 11 328 (target, from-direction) mini-episodes as ONE launch + the ordered backup "
                         "pass -- then the N ramp / beta schedule, every 10th of its 1200 entries, 64 episodes/GPU per entry, returns exchanged "
                         "by key between GPUs"),
    "sff": dict(h=1024, w=1024, maps=64, sff=True,
                desc="C5: static-floor-field sweep, 64 maps/GPU of 1024x1024 with 20 % random rectangular obstacles and 8 exits: "
                     "geodesic BFS-4, BFS-8, (1, sqrt2)-Dijkstra fields + the obstacle-blind L1 field of Create_SFF.py"),
}
WORKLOADS["legacy"] = dict(h=50, w=50, n=100, episodes=8, cap=500, legacy=True, batch=2048,
                           desc="legacy 13-cell models (SURVEY 8 f4) on run_critic_training.py's configuration: 50x50 room, float64 L1 SFF, "
                                "MODEL_PARAMS, N = 100, MAX_STEPS 500 -- sequential-exact TD(0) learning episodes (ffm_ac_core), the same for the "
                                "legacy actor (ffm_actor_only, N = 10, epsilon 0.1), and a frozen-table batch of 2048 episodes")
SECONDARY = ("c3", "c4", "c4train", "c5train", "c2traj", "sff", "legacy")
C4_ROUNDS = 16          # rollout + exchange rounds per timed step of the c4 workload (>= 50 syncs over the default 5 steps)


# ------------------------------------------------------------------------------------------------
# workload construction (product-side code only; no oracle)
# ------------------------------------------------------------------------------------------------
from ffm_b200.workloads import place, room_map, rooms_map_c3, sff_room  # noqa: E402


def build_fields(wl):
    """(map, sff) of a workload; the C3 field is the geodesic 8-connected BFS distance generated on the GPU."""
    if wl.get("plan") == "c3":
        from ffm_b200.sff import generate_sff
        m = rooms_map_c3(wl["h"], wl["w"])
        return m, generate_sff(m, "bfs8", np.float32)
    m = room_map(wl["h"], wl["w"])
    return m, sff_room(m, wl["nbh"])


# full episodes per host thread in the C-port CPU sample (sized for ~10-30 s of CPU work per thread-second scale)
CPU_C_EPISODES = {"c2": 64, "c2dff": 32, "c1": 2048, "c3": 1}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        """Row index now: the samples of a workload are rows[mark_before : mark_after] of the one long-running sampler
        (nvidia-smi needs ~0.5 s to start, longer than the short workloads' timed regions)."""
        return len(self.rows)

    def close(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def window(self, first, last=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[first:last]:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm (the only place this file touches oracle/)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One episode of the NumPy port (oracle/ffm_numpy.py == model/ffm_core.py semantics) for at most
    `budget` seconds; returns (ped_steps, seconds, steps)."""
    wl, pos0, budget, seed = args
    from oracle import ffm_numpy

    class MT:   # the reference's own draw source: global MT19937 (ffm_core.py:84,95,96)
        rs = np.random.RandomState(seed)
        def move(self, t, i, cdf=None): return self.rs.random_sample()
        def coin(self, t, c): return self.rs.random_sample()
        def winner(self, t, c, k): return self.rs.random_sample()

    m, sff = wl["_fields"]
    o = ffm_numpy.CoreOracle(m, sff, pos0,
                             {"k_S": wl["k_S"], "k_D": wl["k_D"], "neighborhood": wl["nbh"]}, MT())
    t0 = time.perf_counter()
    ped = 0
    while o.positions.shape[0] > 0 and o.t < wl["cap"] and time.perf_counter() - t0 < budget:
        ped += o.positions.shape[0]
        o.step()
    return ped, time.perf_counter() - t0, o.t


def _cpu_worker_unified(args):
    """Episodes of the NumPy port of model/ffm_unified.py (critic_only, sequential TD updates) for `budget` seconds."""
    wl, placements, budget, seed = args
    from oracle import unified_numpy

    class MT:
        rs = np.random.RandomState(seed)
        def move(self, t, i, cdf=None): return self.rs.random_sample()
        def winner(self, t, c, k): return self.rs.random_sample()
        def eps_coin(self, t, i): return self.rs.random_sample()
        def eps_pick(self, t, i, k): return self.rs.random_sample()

    m, sff = wl["_fields"]
    o = unified_numpy.UnifiedOracle(m, sff, placements[0], "critic_only", C4_PARAMS, MT())
    t0 = time.perf_counter()
    ped = steps = 0
    for pos0 in placements:
        o.positions, o.t = pos0.astype(np.int64), 0
        o.dff[:] = 0
        while o.positions.shape[0] > 0 and o.t < wl["cap"]:
            ped += o.positions.shape[0]
            o.step()
            steps += 1
        if time.perf_counter() - t0 > budget:
            break
    return ped, time.perf_counter() - t0, steps


def cpu_numpy_port_unified(wl, budget_s, seed=1234, cores=None):
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    m, _ = wl["_fields"]
    free = np.argwhere(m == 0)
    er, ec = np.argwhere(m == 3)[0]
    near = free[np.abs(free[:, 0] - er) + np.abs(free[:, 1] - ec) <= wl["radius"]]
    rng = np.random.RandomState(seed)
    jobs = [(wl, [near[rng.choice(len(near), min(wl["n"], len(near)), replace=False)] for _ in range(64)], budget_s, seed + i)
            for i in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker_unified, jobs)
        wall = time.perf_counter() - t0
    ped = sum(r[0] for r in res)
    return dict(value=ped / wall, unit=UNIT, cores=cores, kind="port",
                sample=f"{cores} processes of the NumPy port of model/ffm_unified.py (oracle/unified_numpy.py, critic_only, the "
                       f"reference's sequential TD updates), whole episodes for {budget_s:.0f} s each: {ped} ped-steps, "
                       f"{sum(r[2] for r in res)} CA steps")


def cpu_numpy_port_legacy(m, sff, pos, params, cap, budget_s, seed):
    """cpu_baseline of the legacy workload: the NumPy restatement of model/ffm_ac_core.py (oracle/legacy_numpy.py) running whole
    sequential-exact learning episodes on ONE host core -- the reference's own shape: one Python process, one shared V dict."""
    from oracle import legacy_numpy

    class Src:
        def __init__(self, s): self.rs = np.random.RandomState(s)
        def move(self, t, i, cdf=None): return self.rs.random_sample()
        def winner(self, t, c, k): return self.rs.random_sample()

    V, ped, eps, steps = {}, 0, 0, 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget_s and eps < len(pos):
        o = legacy_numpy.AcOracle(m, sff, pos[eps], params, Src(seed + eps), v_table=V)
        while o.positions.shape[0] > 0 and o.t < cap:
            ped += o.positions.shape[0]
            o.step()
        steps += o.t
        V = o.V
        eps += 1
    wall = time.perf_counter() - t0
    return dict(value=ped / wall, unit=UNIT, cores=1, kind="port", episodes_per_sec=eps / wall,
                sample=f"{eps} whole learning episodes of the NumPy port of model/ffm_ac_core.py (oracle/legacy_numpy.py) on one core, "
                       f"shared V dict: {ped} ped-steps, {steps} CA steps in {wall:.1f} s")


def cpu_numpy_port(wl, budget_s, seed=1234, cores=None):
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    m, _ = wl["_fields"]
    pos = place(m, wl["n"], cores, 0, seed)
    with mp.get_context("fork").Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(wl, pos[i].astype(np.int64), budget_s, seed + i) for i in range(cores)])
        wall = time.perf_counter() - t0
    ped = sum(r[0] for r in res)
    return dict(value=ped / wall, unit=UNIT, cores=cores, kind="port",
                sample=f"{cores} processes x 1 episode of the NumPy port (oracle/ffm_numpy.py, global-MT19937 draws like "
                       f"the reference), each cut at {budget_s:.0f} s: {ped} ped-steps, {sum(r[2] for r in res)} CA steps "
                       f"(dense early phase of the episode, where the port is fastest per ped-step)")


def cpu_c_port(wl, episodes, seed=1234, cores=None):
    """C restatement (oracle/c), one episode per thread; full episodes."""
    from oracle import c_oracle
    cores = cores or os.cpu_count() or 1
    m, sff = wl["_fields"]
    pos = place(m, wl["n"], episodes, 0, seed)
    t0 = time.perf_counter()
    steps, ped = c_oracle.run_core_batch(m, sff, pos, np.full((episodes,), wl["n"], np.int32),
                                        {"k_S": wl["k_S"], "k_D": wl["k_D"], "neighborhood": wl["nbh"]},
                                        seed=seed, episode_base=0, max_steps=wl["cap"], threads=cores,
                                        track_dff=wl["track_dff"])
    wall = time.perf_counter() - t0
    return dict(value=float(ped.sum()) / wall, unit=UNIT, cores=cores, kind="port",
                sample=f"{episodes} full episodes of the C restatement (oracle/c/ffm_oracle.c, keyed Philox draws), "
                       f"{cores} threads, {int(ped.sum())} ped-steps in {wall:.1f} s")


def have_c_oracle():
    try:
        from oracle import c_oracle
        return c_oracle.available()
    except Exception:
        return False


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if wl.get("plan") == "c3":     # the C3 field is a geodesic distance: CPU-side Dijkstra of the oracle
        from oracle import c_oracle
        m = rooms_map_c3(wl["h"], wl["w"])
        wl["_fields"] = (m, c_oracle.geodesic(m, "bfs8"))
    else:
        m = room_map(wl["h"], wl["w"])
        wl["_fields"] = (m, sff_room(m, wl["nbh"]))
    cores = os.cpu_count() or 1
    times, vals, last = [], [], None
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        if wl.get("model") == "unified":
            last = cpu_numpy_port_unified(wl, budget_s=args.cpu_budget, seed=1234 + it, cores=cores)
        elif have_c_oracle() and not args.numpy_port:
            last = cpu_c_port(wl, episodes=CPU_C_EPISODES.get(args.workload, 64) * cores, seed=1234 + it, cores=cores)
        else:
            last = cpu_numpy_port(wl, budget_s=args.cpu_budget, seed=1234 + it, cores=cores)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
            vals.append(last["value"])
    v = float(np.mean(vals))
    last["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "note": "each step = bounded sample of the workload on host cores"},
            "cpu_baseline": last,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide state of the GPU arm: torch / distributed handles, the L2-flush buffer, measured peaks."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the FFM kernels have no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.local}"))
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
        self.peaks = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        self.hbm_peak = float(self.peaks.get("hbm_gbs", 6650.0))
        self.hbm_peak_source = "measured (MEASURED_PEAKS.json hbm_gbs)" if self.peaks else "fallback 6650 GB/s"
        self._smem = None
        self.sampler = ClockSampler(self.local)     # one sampler for the whole run, 50 ms period; workloads take windows of it
        self.sampler.start()
        try:
            self.profile = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        except OSError:
            self.profile = {}

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, maxed, summed):
        """max over ranks of the timings, sum over ranks of the unit counts"""
        torch = self.torch
        tm = torch.tensor(maxed, dtype=torch.float64, device="cuda")
        ts = torch.tensor(summed, dtype=torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(tm, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(ts, op=self.dist.ReduceOp.SUM)
        return tm.tolist(), ts.tolist()

    def smem_peak(self):
        """(GB/s, SM MHz) of the shared-memory read micro-benchmark (conflict-free 16-byte loads on every SM)."""
        if self._smem is None:
            import ctypes as C
            from ffm_b200 import _abi
            g, mhz = C.c_double(), C.c_double()
            _abi.check(_abi.lib().ffm_measure_smem_bandwidth(self.local, C.byref(g), C.byref(mhz)))
            self._smem = (g.value, mhz.value)
        return self._smem


def event_pairs(torch, k, n):
    return [[torch.cuda.Event(enable_timing=True) for _ in range(n)] for _ in range(k)]


def run_rollout_workload(ctx, name, wl):
    """One rollout workload: W warm-up passes, K timed passes with the inputs resident in HBM (CUDA events), K timed
    passes through host buffers (e2e), clocks sampled over both.  Returns the fields of a bench line."""
    torch, args = ctx.torch, ctx.args
    from ffm_b200 import BatchSim
    world, rank, local = ctx.world, ctx.rank, ctx.local
    B, N, cap = wl["episodes"], wl["n"], wl["cap"]
    m, sff = build_fields(wl)
    wl["_fields"] = (m, sff)
    params = {"k_S": wl["k_S"], "k_D": wl["k_D"], "diffuse": 0.2, "decay": 0.2, "neighborhood": wl["nbh"]}
    ep_base = rank * B                                   # global episode ids: results independent of N
    n_np = np.full((B,), N, dtype=np.int32)
    unified = wl.get("model") == "unified"
    record = bool(wl.get("record"))
    rounds = C4_ROUNDS if unified else 1
    learner = None
    if unified:
        from ffm_b200 import UnifiedSim
        from ffm_b200.sharding import BatchedLearner
        sim = UnifiedSim(m, sff, B, N, mode="critic_only", learn="batched", params=C4_PARAMS, seed=args.seed,
                         episode_base=ep_base, device=local)
        exit_rc = tuple(int(v) for v in np.argwhere(m == 3)[0])
        sim.place(n_np, exit_pos=exit_rc, radius=wl["radius"])          # initialize_agents(exit_pos, radius), keyed per episode
        pos_np, n_np = sim.get_positions()
        learner = BatchedLearner(sim, overlap=world > 1)                # one all-reduce of the flat delta buffer per sync
    else:
        pos_np = place(m, N, B, ep_base, args.seed)
        sim = BatchSim(m, sff, B, N, params, seed=args.seed, episode_base=ep_base, track_dff=wl["track_dff"], device=local)
    info = sim.kernel_info()
    pos_dev = torch.from_numpy(pos_np).cuda()
    n_dev = torch.from_numpy(n_np).cuda()
    pos_pin = torch.from_numpy(pos_np).pin_memory()
    n_pin = torch.from_numpy(n_np).pin_memory()
    steps_dev = torch.zeros(B, dtype=torch.int32, device="cuda")
    ped_dev = torch.zeros(B, dtype=torch.int64, device="cuda")
    rec_cap, rec_host, rec = 0, None, {}
    if record:
        rec_cap = 1 << 20                                              # entries per episode (C2: ~8.9e5 pedestrian-steps + padding)
        rec_host = torch.empty((B, rec_cap, 2), dtype=torch.int16).pin_memory()

    def roll():
        if record:
            rec["r"] = sim.rollout(cap, record=cap, compact_cap=rec_cap)
        else:
            sim.rollout(cap)

    def one_pass_resident(ev=None):
        if ev: ev[0].record()
        kern = 0.0
        for r in range(rounds):
            sim.set_positions(pos_dev, n_dev)
            if ev and r == 0: ev[1].record()
            roll()
            if ev and r == 0: ev[2].record()
            if learner: learner.sync()
        if learner: learner.flush()
        sim.counters_into(steps_dev, ped_dev)
        if ev: ev[3].record()

    def one_pass_e2e():
        ped = 0
        for r in range(rounds):
            sim.set_positions(pos_pin.numpy(), n_pin.numpy())       # H2D of this round's inputs
            roll()
            if learner: learner.sync()
            if r == rounds - 1 and learner: learner.flush()
            if record:                                              # D2H of the trajectory record (+ its CSR offsets / counts)
                rec_host.copy_(rec["r"]["ctraj"], non_blocking=True)
                rec["off_h"] = rec["r"]["off"].cpu(); rec["n_h"] = rec["r"]["n"].cpu()
            st_h, ped_h = sim.counters()                             # D2H of the round's result (+ sync)
            ped += int(ped_h.sum())
        return ped

    for _ in range(args.warmup):
        ctx.flush.fill_(1)
        one_pass_resident()
    ctx.barrier()
    ped_per_round = int(ped_dev.sum().item())
    ped_per_pass = ped_per_round * rounds
    steps_host = steps_dev.cpu().numpy()

    sampler = ctx.sampler
    mark0 = sampler.mark()
    evs = event_pairs(torch, args.steps, 4)
    l0 = sim.launch_count
    ctx.barrier()
    for k in range(args.steps):
        ctx.flush.fill_(k & 0xFF)                 # L2 flush between timed iterations (not timed)
        one_pass_resident(evs[k])
    ctx.barrier()
    launches = sim.launch_count - l0
    total_ms = sum(e[0].elapsed_time(e[3]) for e in evs)
    kern_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps

    sync_ms = None
    if learner:                                   # the exchange alone: all-reduce + fold-in on the critical path (blocking form)
        from ffm_b200.sharding import BatchedLearner
        plain = BatchedLearner(sim, overlap=False)
        sim.set_positions(pos_dev, n_dev); sim.rollout(cap); plain.sync()
        ctx.barrier()
        se = event_pairs(torch, 20, 2)
        for k in range(20):
            se[k][0].record(); plain.sync(); se[k][1].record()
        ctx.barrier()
        sync_ms = float(np.median([e[0].elapsed_time(e[1]) for e in se]))
        sim.bind_deltas(learner.bufs[learner._cur])

    # ---- e2e: host buffers in, host counters (and the trajectory record) out ------------------------
    e2e_evs = event_pairs(torch, args.steps, 2)
    for _ in range(2):
        one_pass_e2e()
    ctx.barrier()
    e2e_ped = 0
    for k in range(args.steps):
        ctx.flush.fill_(k & 0xFF)
        e2e_evs[k][0].record()
        e2e_ped += one_pass_e2e()
        e2e_evs[k][1].record()
    ctx.barrier()
    clocks = sampler.window(mark0)
    e2e_ms = sum(e[0].elapsed_time(e[1]) for e in e2e_evs)

    (total_ms, e2e_ms, kern_ms), (ped_all, e2e_ped_all) = ctx.reduce([total_ms, e2e_ms, kern_ms], [float(ped_per_pass), float(e2e_ped)])
    out = None
    if rank == 0:
        value = ped_all * args.steps / (total_ms * 1e-3)
        e2e_value = e2e_ped_all / (e2e_ms * 1e-3)
        alg_bytes = (BYTES_PER_PED_STEP_UNIFIED if unified else BYTES_PER_PED_STEP[wl["track_dff"]]) * ped_per_round
        if wl["track_dff"]:
            alg_bytes += 8 * wl["h"] * wl["w"] * int(steps_host.sum())      # DFF field read+write per episode-step
        if record:
            alg_bytes += 4 * ped_per_round                                  # the record itself: 4 B per pedestrian-step to HBM
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        prof = ctx.profile.get(name, {})
        on_chip = bool(info["fields_in_smem"]) or not wl["track_dff"]      # SFF-only runs keep the bitboards / owner grid on chip either way
        hbm = {"achieved": achieved, "peak": ctx.hbm_peak, "unit": "GB/s", "frac": achieved / ctx.hbm_peak, "peak_source": ctx.hbm_peak_source}
        if on_chip and not record:
            smem_gbs, mhz = ctx.smem_peak()
            roof = {"bound": "smem", "achieved": achieved, "peak": smem_gbs, "unit": "GB/s", "frac": achieved / smem_gbs,
                    "peak_source": f"measured live: ffm_measure_smem_bandwidth (conflict-free 16-byte shared loads on all SMs, {mhz:.0f} MHz)",
                    "hbm": hbm,
                    "note": "the episode state is shared-memory resident: DRAM traffic is the prologue/epilogue only (see traffic); the "
                            "kernel is issue-bound, see issue_active_pct / smem wavefronts from the committed ncu capture"}
        else:
            roof = {"bound": "hbm", **hbm,
                    "note": ("record stores stream to HBM; " if record else "") +
                            ("score/DFF fields live in L2/HBM at this map size" if not info["fields_in_smem"] else "")}
        if prof.get("issue_active_pct") is not None:
            # what actually bounds the on-chip rollouts: instruction issue (from the committed ncu capture of this command)
            roof["issue"] = {"achieved": prof.get("ipc"), "peak": 4.0, "unit": "warp instructions / clk / SM",
                             "frac": prof["issue_active_pct"] / 100.0, "source": prof.get("source")}
        roof.update({"traffic": prof.get("dram_bytes_per_launch") if B == prof.get("episodes") else None,
                     "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": alg_bytes,
                     "ncu": {k: prof.get(k) for k in ("smem_wavefronts_pct_of_peak", "smem_bank_conflict_share", "issue_active_pct",
                                                       "lanes_active_per_instruction", "warps_active_pct", "source") if k in prof}})
        par = f"episodes sharded over {world} GPU(s)" + (
            f"; one NCCL all-reduce of the flat table-delta buffer ({(3 + sim.A) * sim.S * 8} B) per sync, "
            f"{'overlapped with the next rollout (staleness 1)' if learner.overlap else 'blocking'}" if learner else ", no collective")
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "episodes_per_gpu": B, "peds_per_episode": N, "map": f"{wl['h']}x{wl['w']}",
                       "step_cap": cap, "parallelism": par,
                       "l2": "256 MiB L2 flush between timed iterations (untimed)",
                       "kernel": info,
                       "mean_evacuation_steps": float(steps_host.mean()), "ped_steps_per_pass_per_gpu": ped_per_pass},
            "episodes_per_sec": B * rounds * world * args.steps / (total_ms * 1e-3),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(pos_np.nbytes + n_np.nbytes) * rounds,
                    "d2h_bytes_per_step": (int(B * 4 + B * 8) + (int(rec_host.numel() * 2 + B * (2 * cap + 1) * 4) if record else 0)) * rounds,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "roofline": roof,
        }
        if learner:
            out["config"]["rounds_per_step"] = rounds
            out["config"]["syncs_timed"] = rounds * args.steps
            out["sync_ms_blocking"] = sync_ms
        if record:
            off_h = rec["off_h"].numpy(); st = np.minimum(steps_host, cap)
            used = int(off_h[np.arange(B), st].sum())
            out["record"] = {"bytes_written_per_pass_per_gpu": used * 4, "ped_steps_per_pass_per_gpu": ped_per_pass,
                             "bytes_per_ped_step": used * 4 / ped_per_pass, "overflowed": bool((off_h[np.arange(B), st] < 0).any())}
    sim.close()
    del sim
    torch.cuda.empty_cache()
    return out


def run_sff_workload(ctx, name, wl):
    """C5: SFF sweep.  A step = all fields of all maps of this rank.  value = cells/s over the geodesic modes."""
    torch, args = ctx.torch, ctx.args
    from ffm_b200.sff import generate_sff
    from ffm_b200.workloads import obstacle_map_c5
    world, rank = ctx.world, ctx.rank
    M, H, W = wl["maps"], wl["h"], wl["w"]
    maps = np.stack([obstacle_map_c5(H, W, index=rank * M + i) for i in range(M)])
    dm = torch.from_numpy(maps).cuda()
    pin = torch.from_numpy(maps).pin_memory()
    modes = ("bfs4", "bfs8", "dijkstra8", "L1")
    host_out = torch.empty((M, H, W), dtype=torch.float32).pin_memory()

    def sweep(ev=None):
        for i, mode in enumerate(modes):
            if ev: ev[i].record()
            out = generate_sff(dm, mode, np.float32)
        if ev: ev[len(modes)].record()
        return out

    for _ in range(args.warmup):
        sweep()
    ctx.barrier()
    sampler = ctx.sampler
    mark0 = sampler.mark()
    evs = event_pairs(torch, args.steps, len(modes) + 1)
    ctx.barrier()
    for k in range(args.steps):
        ctx.flush.fill_(k & 0xFF)
        sweep(evs[k])
    ctx.barrier()
    total_ms = sum(e[0].elapsed_time(e[-1]) for e in evs)
    mode_ms = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) / args.steps for i in range(len(modes))]
    # e2e: maps from pinned host memory, fields back to pinned host memory
    e2 = event_pairs(torch, args.steps, 2)
    ctx.barrier()
    for k in range(args.steps):
        e2[k][0].record()
        dmk = pin.cuda(non_blocking=True)
        for mode in modes:
            host_out.copy_(generate_sff(dmk, mode, np.float32), non_blocking=True)
        torch.cuda.synchronize()
        e2[k][1].record()
    ctx.barrier()
    clocks = sampler.window(mark0)
    e2e_ms = sum(e[0].elapsed_time(e[1]) for e in e2)
    cells = float(M * H * W * len(modes))
    (total_ms, e2e_ms, *mode_ms), (cells_all,) = ctx.reduce([total_ms, e2e_ms] + mode_ms, [cells])
    if rank != 0:
        return None
    geo_ms = sum(mode_ms[:3])
    alg = 5.0 * M * H * W            # >= 5 B per cell (1 B map read + 4 B field write) per field
    return {
        "metric": "sff_cells_per_sec", "value": cells_all * args.steps / (total_ms * 1e-3), "unit": "cells/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "maps_per_gpu": M, "map": f"{H}x{W}", "fields_per_step": list(modes),
                   "parallelism": f"maps sharded over {world} GPU(s), no collective", "l2": "256 MiB L2 flush between timed iterations (untimed)"},
        "maps_per_sec": M * len(modes) * world * args.steps / (total_ms * 1e-3),
        "per_mode_ms": dict(zip(modes, mode_ms)),
        "geodesic_cells_per_sec": 3.0 * M * H * W * world / (geo_ms * 1e-3),
        "clocks": clocks,
        "e2e": {"value": cells_all * args.steps / (e2e_ms * 1e-3), "unit": "cells/s", "h2d_bytes_per_step": int(maps.nbytes),
                "d2h_bytes_per_step": int(host_out.numel() * 4 * len(modes)), "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": args.steps * (3 * 3 + 2),
        "roofline": {"bound": "hbm", "achieved": alg * 3 / (geo_ms * 1e-3) / 1e9, "peak": ctx.hbm_peak, "unit": "GB/s",
                     "frac": alg * 3 / (geo_ms * 1e-3) / 1e9 / ctx.hbm_peak, "traffic": None, "peak_source": ctx.hbm_peak_source,
                     "kernel_ms": geo_ms / 3, "algorithmic_bytes_per_launch": alg,
                     "note": "geodesic modes (sff_relax_queue_kernel): wavefront propagation is latency-bound, not bandwidth-bound; "
                             "cells/s and maps/s are the meaningful figures (SURVEY 8(d))"},
    }


def run_training_workload(ctx, name, wl):
    """C4 as a pipeline: critic curriculum, then actor curriculum, then the frozen-H evaluation the reference accepts a training
    by.  A step = both curricula from scratch; value = training episodes per second (all ranks)."""
    torch, args = ctx.torch, ctx.args
    from ffm_b200 import unified_training as ut
    m = room_map(wl["h"], wl["w"])
    sff = sff_room(m, "neumann")
    exit_pos = tuple(int(v) for v in np.argwhere(m == 3)[0])
    configs = ut.curriculum(m, exit_pos)
    batch, rounds = wl["batch"], wl["rounds"]
    episodes = 2 * len(configs) * rounds * batch                     # critic + actor, per rank
    out = {}

    def one(seed):
        V, _ = ut.train_critic(m, sff, exit_pos, configs=configs, batch=batch, rounds=rounds, seed=seed, device=ctx.local)
        H, _, _ = ut.train_actor(m, sff, exit_pos, V, configs=configs, batch=batch, rounds=rounds, seed=seed + 1, device=ctx.local)
        out["H"], out["V"] = H, V

    for k in range(min(args.warmup, 2)):
        one(10 + k)
    ctx.barrier()
    mark0 = ctx.sampler.mark()
    evs = event_pairs(torch, args.steps, 2)
    for k in range(args.steps):
        evs[k][0].record()
        one(100 + 2 * k)
        evs[k][1].record()
    ctx.barrier()
    clocks = ctx.sampler.window(mark0)
    total_ms = sum(e[0].elapsed_time(e[1]) for e in evs)
    (total_ms,), (eps_all,) = ctx.reduce([total_ms], [float(episodes)])
    band = {}
    for N in (10, 50, 90):
        steps, frac = ut.evaluate_trained(m, sff, exit_pos, out["H"], N, episodes=256, device=ctx.local)
        band[str(N)] = {"mean_steps": float(steps.mean()), "in_band_2N-1..2N+14": frac}
    if ctx.rank != 0:
        return None
    value = eps_all * args.steps / (total_ms * 1e-3)
    return {
        "metric": "training_episodes_per_sec", "value": value, "unit": "episodes/s", "n_gpus": ctx.world, "steps": args.steps,
        "warmup": min(args.warmup, 2), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "configurations": len(configs), "episodes_per_step_per_gpu": episodes,
                   "parallelism": f"episodes sharded over {ctx.world} GPU(s); one NCCL all-reduce of the flat table-delta buffer per sync",
                   "reference": "run_20260115_234109/summary.txt: 45 000 actor episodes in 17:18:51 = 0.72 episodes/s (hardware unknown)"},
        "acceptance": band, "tables": {"V_states": len(out["V"]), "H_rows": len(out["H"])},
        "clocks": clocks,
        "e2e": {"value": value, "unit": "episodes/s", "h2d_bytes_per_step": int(m.nbytes + sff.nbytes) * 2,
                "d2h_bytes_per_step": int((len(out["V"]) + 5 * len(out["H"])) * 8), "ms_per_step": total_ms / args.steps,
                "note": "the pipeline is driven from the host API as is: maps / fields in, V and H dicts out, inside the timed region"},
        "gpu_launches": int(args.steps * 2 * len(configs) * rounds * (2 + (-(-ut.MAX_STEPS // 8)) * 2)),   # place + (rollout, apply) per sync
        "roofline": {"bound": "hbm", "achieved": None, "peak": ctx.hbm_peak, "unit": "GB/s", "frac": None, "traffic": None,
                     "note": "latency-bound pipeline of short launches (ffm_unified_rollout_kernel + apply_deltas + NCCL); see c4 for the kernel"},
    }


def run_mcq_training_workload(ctx, name, wl):
    """C5 as a pipeline: coverage pretrain (replicated on every rank: its result is the shared starting table) + the training
    schedule (episodes sharded over ranks, by-key exchange of the returns).  value = training episodes per second."""
    torch, args = ctx.torch, ctx.args
    from ffm_b200.mcq_training import coverage_patterns, coverage_pretrain, run_training
    m = room_map(wl["h"], wl["w"])
    sff = sff_room(m, "neumann").astype(np.float64)
    params = {"k_S": 3, "k_D": 1, "diffuse": 0.2, "decay": 0.2, "max_steps": 500, "alpha": 0.1, "gamma": 0.99}    # config/default_config.yaml
    order = coverage_patterns(m, shuffle=False)
    entries = list(range(0, 1200, wl["stride"]))
    batch = wl["batch"]
    out = {}

    def one(seed):
        Q, steps = coverage_pretrain(m, sff, params, {}, order=order, seed=seed, device=ctx.local, return_steps=True)
        out["pre"], out["pre_steps"] = len(Q), int(steps.sum())
        Q, mean_steps = run_training(m, sff, params, full_N=100, shared_Q=Q, batch=batch, seed=seed + 1, device=ctx.local, entries=entries)
        out["Q"], out["mean_steps"] = len(Q), mean_steps

    for k in range(min(args.warmup, 2)):
        one(10 + k)
    ctx.barrier()
    mark0 = ctx.sampler.mark()
    evs = event_pairs(torch, args.steps, 2)
    for k in range(args.steps):
        evs[k][0].record()
        one(100 + 2 * k)
        evs[k][1].record()
    ctx.barrier()
    clocks = ctx.sampler.window(mark0)
    total_ms = sum(e[0].elapsed_time(e[1]) for e in evs)
    (total_ms,), (sched_all,) = ctx.reduce([total_ms], [float(len(entries) * batch)])
    if ctx.rank != 0:
        return None
    episodes = len(order) + sched_all                      # the pretrain is replicated: counted once
    value = episodes * args.steps / (total_ms * 1e-3)
    return {
        "metric": "training_episodes_per_sec", "value": value, "unit": "episodes/s", "n_gpus": ctx.world, "steps": args.steps,
        "warmup": min(args.warmup, 2), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "pretrain_patterns": len(order), "pretrain_ca_steps": out["pre_steps"],
                   "schedule_entries": len(entries), "episodes_per_entry_per_gpu": batch,
                   "parallelism": f"pretrain replicated; schedule episodes sharded over {ctx.world} GPU(s), one all-gather of the touched (key, return sums) rows per entry"},
        "tables": {"Q_rows_after_pretrain": out["pre"], "Q_rows_after_training": out["Q"]},
        "mean_steps_first_last_entry": [out["mean_steps"][0], out["mean_steps"][-1]],
        "clocks": clocks,
        "e2e": {"value": value, "unit": "episodes/s", "h2d_bytes_per_step": int(m.nbytes + sff.nbytes) * 2, "d2h_bytes_per_step": int(out["Q"] * 28),
                "ms_per_step": total_ms / args.steps,
                "note": "the pipeline is driven from the host API as is: map / field in, Q dict out, inside the timed region"},
        "gpu_launches": int(args.steps * (3 + len(entries) * 5)),     # pretrain: place-free rollout + ordered backup (+ table load); per entry: place, rollout, accumulate, fold (+ exchange)
        "roofline": {"bound": "hbm", "achieved": None, "peak": ctx.hbm_peak, "unit": "GB/s", "frac": None, "traffic": None,
                     "note": "latency-bound pipeline (ffm_mcq_rollout_kernel: hash-table lookups in L2, paths in HBM)"},
    }


LEGACY_PARAMS = {"k_S": 10, "k_D": 1, "alpha_v": 0.01, "gamma": 0.99, "exit_reward": 100.0, "step_penalty": -1.0,
                 "collision_penalty": -1.0, "neighborhood": "neumann", "block_size": 5}     # run_critic_training.py:34-44


def run_legacy_workload(ctx, name, wl):
    """Legacy 13-cell models.  A step = `episodes` sequential-exact learning episodes of the TD critic on one handle per GPU
    (the reference's own semantics: one episode after the other on a shared V dict -- "replicas only" across GPUs).
    The ffm_legacy_* calls take HOST buffers and return when done, so value and e2e are the same measurement."""
    torch, args = ctx.torch, ctx.args
    from ffm_b200.legacy import LegacySim
    m = room_map(wl["h"], wl["w"])
    sff64 = sff_room(m, "neumann").astype(np.float64)
    E, N, cap, HW = wl["episodes"], wl["n"], wl["cap"], wl["h"] * wl["w"]
    pos = place(m, N, E, ctx.rank * E, args.seed)
    ac = LegacySim(m, sff64, 1, N, model="ac", learn="exact", params=LEGACY_PARAMS, seed=args.seed, device=ctx.local)
    Na = 10
    actor = LegacySim(m, sff64, 1, Na, model="actor_only", learn="exact", params={**LEGACY_PARAMS, "k_A": 10, "alpha_h": 0.1},
                      seed=args.seed, device=ctx.local)
    actor.set_epsilon(0.1)
    B = wl["batch"]
    frozen = LegacySim(m, sff64, B, N, model="ac", learn="none", params=LEGACY_PARAMS, seed=args.seed, episode_base=ctx.rank * B,
                       device=ctx.local)
    bpos = place(m, N, B, ctx.rank * B, args.seed + 1)
    counts = {}

    def episodes(sim, n, tag):
        ps = 0
        for e in range(E):
            sim.set_episode_base(ctx.rank * E + e)
            sim.set_positions(np.ascontiguousarray(pos[e:e + 1, :n]), np.array([n], np.int32))
            sim.zero_dff()
            sim.rollout(cap)
            ps += int(sim.counters()[1][0])
        counts[tag] = ps

    def batch():
        frozen.set_positions(bpos, np.full((B,), N, np.int32))
        frozen.zero_dff()
        frozen.rollout(cap)
        counts["frozen"] = int(frozen.counters()[1].sum())

    for _ in range(min(args.warmup, 2)):
        episodes(ac, N, "ac"); episodes(actor, Na, "actor"); batch()
    ctx.barrier()
    mark0 = ctx.sampler.mark()
    evs = event_pairs(torch, args.steps, 4)
    tot = {"ac": 0, "actor": 0, "frozen": 0}
    for k in range(args.steps):
        evs[k][0].record()
        episodes(ac, N, "ac")
        evs[k][1].record()
        episodes(actor, Na, "actor")
        evs[k][2].record()
        batch()
        evs[k][3].record()
        for t in tot:
            tot[t] += counts[t]
    ctx.barrier()
    clocks = ctx.sampler.window(mark0)
    ms = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) for i in range(3)]
    ms, sums = ctx.reduce(ms, [float(tot["ac"]), float(tot["actor"]), float(tot["frozen"])])
    if ctx.rank != 0:
        return None
    value = sums[0] / (ms[0] * 1e-3)
    cpu = None
    if ctx.world == 1 and not args.no_cpu:
        try:
            cpu = cpu_numpy_port_legacy(m, sff64, pos, LEGACY_PARAMS, cap, min(args.cpu_budget, 8.0), args.seed & 0x7FFFFFFF)
        except Exception as ex:   # the baseline is reporting, never the product
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex!r}"}
    return {
        "cpu_baseline": cpu,
        "metric": "pedestrian_steps_per_sec", "value": value, "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps, "warmup": min(args.warmup, 2),
        "ms_per_step": ms[0] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "episodes_per_step_per_gpu": E, "parallelism": f"replicas only: {ctx.world} independent learner(s), no collective",
                   "reference": "critic_training/run_20251206_153157/summary.txt: 11 000 episodes in 03:25:36 = 0.89 episodes/s (hardware unknown)"},
        "episodes_per_sec": ctx.world * E * args.steps / (ms[0] * 1e-3),
        "actor_only": {"value": sums[1] / (ms[1] * 1e-3), "unit": UNIT, "episodes_per_sec": ctx.world * E * args.steps / (ms[1] * 1e-3), "N": Na,
                       "epsilon": 0.1, "ms_per_step": ms[1] / args.steps},
        "frozen_batch": {"value": sums[2] / (ms[2] * 1e-3), "unit": UNIT, "episodes_per_gpu": B, "ms_per_step": ms[2] / args.steps},
        "tables": {"V_states": ac.table_size("V"), "actor_V_states": actor.table_size("V"), "actor_H_rows": actor.table_size("H")},
        "clocks": clocks,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(E * (N * 8 + 4)), "d2h_bytes_per_step": int(E * 12),
                "ms_per_step": ms[0] / args.steps, "note": "the legacy entry points take host buffers: positions in, counters out, every episode"},
        "gpu_launches": int(args.steps * (2 * E + 1)),
        "roofline": {"bound": "hbm", "achieved": None, "peak": ctx.hbm_peak, "unit": "GB/s", "frac": None, "traffic": None,
                     "note": "sequential-exact learning: one CTA, the TD / actor updates of a step applied by ONE thread in agent order (the reference's "
                             "semantics) -- latency-bound by construction; the frozen batch shows the kernel's parallel rate"},
    }


def run_workload(ctx, name, wl):
    if wl.get("legacy"):
        return run_legacy_workload(ctx, name, wl)
    if wl.get("train") == "mcq":
        return run_mcq_training_workload(ctx, name, wl)
    if wl.get("sff"):
        return run_sff_workload(ctx, name, wl)
    if wl.get("train"):
        return run_training_workload(ctx, name, wl)
    return run_rollout_workload(ctx, name, wl)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--episodes", type=int, default=None, help="episodes per GPU (default: workload's)")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--numpy-port", action="store_true", help="reference arm: force the NumPy port")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads of the default C2 run")
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=0x5EED0002)
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.episodes:
        wl["episodes"] = args.episodes
    if args.impl == "reference":
        if wl.get("sff") or wl.get("record") or wl.get("train") or wl.get("legacy"):
            raise SystemExit("the reference arm covers the rollout workloads")
        run_reference_arm(args, wl)
        return

    # stdout carries exactly ONE JSON line: anything a library writes to fd 1 meanwhile (NCCL prints its
    # version banner there) is sent to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    ctx = Ctx(args)
    if args.warmup < 3:
        args.warmup = 3
    line = run_workload(ctx, args.workload, wl)
    secondary = {}
    if args.workload == "c2" and not args.no_secondary and not args.episodes:
        for name in SECONDARY:
            try:
                secondary[name] = run_workload(ctx, name, dict(WORKLOADS[name]))
            except Exception as ex:      # a secondary workload never takes the headline down with it
                secondary[name] = {"error": repr(ex)}
    if ctx.rank == 0:
        if secondary:
            line["secondary"] = secondary
        if not args.no_cpu and ctx.world == 1 and not wl.get("sff") and not wl.get("train") and not wl.get("legacy"):
            try:
                if wl.get("model") == "unified":
                    line["cpu_baseline"] = cpu_numpy_port_unified(wl, budget_s=min(args.cpu_budget, 10.0))
                elif have_c_oracle():
                    line["cpu_baseline"] = cpu_c_port(wl, episodes=CPU_C_EPISODES.get(args.workload, 64) * (os.cpu_count() or 1))
                    line["cpu_baseline_numpy"] = cpu_numpy_port(wl, budget_s=min(args.cpu_budget, 10.0))
                else:
                    line["cpu_baseline"] = cpu_numpy_port(wl, budget_s=args.cpu_budget)
            except Exception as ex:   # the baseline is reporting, never the product
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex!r}"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    ctx.sampler.close()
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
