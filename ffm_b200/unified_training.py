"""Batched GPU forms of the reference's curriculum drivers for the unified model (12x12 room, MODEL_PARAMS):

  train_critic      run_unified_critic_training.py:142-225 -- critic_only TD(0) over the curriculum radius x N
  train_actor       run_unified_actor_training.py:33-35,51-52,193-268 -- actor_only on a pretrained V, the same curriculum,
                    epsilon decaying linearly 0.2 -> 0.01 WITHIN every (radius, N) configuration (:251-259)
  evaluate_trained  run_trained_ffm.py:199-243 -- frozen H through the trained-actor model; the reference's acceptance band
                    for the evacuation time is [2N - 1, 2N + 14] (analyze_steps_by_n.py:109-110)

Every (radius, N) configuration runs as `rounds` synchronous batches of `batch` episodes (kernel mode FFM_LEARN_BATCHED:
TD errors against frozen tables, one all-reduce of the flat delta buffer per sync, sharding.BatchedLearner) instead of
EPISODES_PER_CONFIG sequential episodes; episodes are keyed by a global id, so a run is reproducible on any number of GPUs.
Host logic only; the arithmetic is in csrc/ffm_unified_kernel.cuh.
"""
import numpy as np

from .sharding import BatchedLearner, world
from .sim import RoundParams, UnifiedSim

MODEL_PARAMS = dict(k_S=10, k_D=1, k_A=10, alpha_v=0.01, alpha_h=0.1, gamma=0.99, exit_reward=100.0, step_penalty=-1.0,
                    collision_penalty=-1.0, neighborhood="neumann", block_size=1)       # run_unified_actor_training.py:58-70
RADIUS_LIST = list(range(3, 16, 2))                                                   # :33-35
N_LIST = [1] + list(range(10, 91, 10))                                                # :38-41 ([1, 10, 20, ..., 90])
MAX_STEPS = 300                                                                       # :47
EPSILON_START, EPSILON_END = 0.2, 0.01                                                # :51-52


def count_available_cells(map_array, exit_pos, radius):
    """run_unified_actor_training.py: free cells within L1 `radius` of the exit."""
    free = np.argwhere(np.asarray(map_array) == 0)
    return int((np.abs(free[:, 0] - exit_pos[0]) + np.abs(free[:, 1] - exit_pos[1]) <= radius).sum())


def curriculum(map_array, exit_pos, radius_list=RADIUS_LIST, n_list=N_LIST):
    """The (radius, N) configurations in driver order, those with N > available cells skipped (:221-226)."""
    return [(r, n) for r in radius_list for n in n_list if n <= count_available_cells(map_array, exit_pos, r)]


def _run_curriculum(sim, learner, configs, exit_pos, batch, rounds, max_steps, sync_every, epsilon_schedule, log, use_graph=False):
    """use_graph: the launch chain of one round -- ceil(max_steps / sync_every) x (rollout, exchange, fold-in) -- is captured
    ONCE in a CUDA graph and replayed for every round; the two parameters that change per round, epsilon and the episode
    key, then live in a device struct (RoundParams / ffm_bind_dynamic) instead of the launches' by-value parameters.
    Measured on one B200 (profiles/exp_graph_training.py): the critic curriculum 85 -> 71 ms, the actor curriculum
    unchanged -- the chain is bound by the latency of its ~115 dependent short kernels per round, not by the host's
    launch rate -- so the drivers keep eager launches by default."""
    import torch
    rank, ws = world()
    ep = 0
    history = []
    graph, rp = None, None
    if use_graph:
        assert not learner.overlap
        rp = RoundParams(f"cuda:{sim.device}")
        sim.bind_dynamic(rp.dev)

    def chain():
        done = 0
        while done < max_steps:
            k = min(sync_every, max_steps - done)
            sim.rollout(k)
            learner.sync()
            done += k

    for ci, (radius, N) in enumerate(configs):
        for r in range(rounds):
            eps = sim.params.get("epsilon", 0.0) or 0.0
            if epsilon_schedule:
                progress = (r + 1) / rounds                                             # :251-259, per configuration
                eps = EPSILON_START + (EPSILON_END - EPSILON_START) * progress
            base = (ep * ws + rank) * batch
            ep += 1
            sim.set_epsilon(eps)
            sim.set_episode_base(base)            # host copies: the placement below is keyed by the episode id as well
            if use_graph:
                rp.set(eps, base)                 # device copies: what the replayed rollouts read
            sim.place(np.full(batch, N, np.int32), exit_pos=exit_pos, radius=radius)
            if not use_graph:
                chain()
                if learner.overlap:
                    learner.flush()
            elif graph is None:
                # the first round runs eagerly on a side stream (warm-up; it is a real training round), then the same chain
                # is captured -- capturing executes nothing -- and every later round replays it
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    chain()
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    chain()
            else:
                graph.replay()
        steps = sim.counters()[0]
        history.append((radius, N, float(steps.mean())))
        if log:
            log(f"config {ci + 1}/{len(configs)} radius={radius} N={N}: mean steps {steps.mean():.1f}")
    if use_graph:
        torch.cuda.synchronize()
        sim.bind_dynamic(None)
    return history


def train_critic(map_array, sff, exit_pos, params=None, configs=None, batch=256, rounds=4, max_steps=MAX_STEPS, sync_every=8,
                 seed=0, device=None, log=None, use_graph=False):
    """-> (V dict like get_v_table(), history).  Batched TD(0) of the critic over the curriculum."""
    p = {**MODEL_PARAMS, **(params or {})}
    configs = curriculum(map_array, exit_pos) if configs is None else configs
    n_max = max(n for _, n in configs)
    sim = UnifiedSim(map_array, sff, batch, n_max, mode="critic_only", learn="batched", params=p, seed=seed, device=device)
    hist = _run_curriculum(sim, BatchedLearner(sim), configs, exit_pos, batch, rounds, max_steps, sync_every, False, log, use_graph)
    V = sim.v_dict()
    sim.close()
    return V, hist


def train_actor(map_array, sff, exit_pos, v_table, params=None, configs=None, batch=256, rounds=4, max_steps=MAX_STEPS,
                sync_every=8, seed=1, device=None, log=None, use_graph=False):
    """-> (H dict like get_h_table(), V dict, history).  actor_only on the pretrained V (new states still get their V
    learned, ffm_unified.py:561-574), epsilon-greedy exploration decaying within every configuration."""
    p = {**MODEL_PARAMS, **(params or {})}
    configs = curriculum(map_array, exit_pos) if configs is None else configs
    n_max = max(n for _, n in configs)
    sim = UnifiedSim(map_array, sff, batch, n_max, mode="actor_only", learn="batched", params=p, seed=seed, device=device)
    sim.load_v_dict(v_table)
    hist = _run_curriculum(sim, BatchedLearner(sim), configs, exit_pos, batch, rounds, max_steps, sync_every, True, log, use_graph)
    H, V = sim.h_dict(), sim.v_dict()
    sim.close()
    return H, V, hist


def evaluate_trained(map_array, sff, exit_pos, h_table, N, radius=15, episodes=256, params=None, max_steps=MAX_STEPS, seed=2,
                     device=None):
    """Frozen H through the trained-actor model (run_trained_ffm.py:199-243) -> (steps int32 [episodes], fraction of episodes
    inside the reference's band [2N - 1, 2N + 14])."""
    p = {k: v for k, v in {**MODEL_PARAMS, **(params or {})}.items() if k in ("k_D", "k_A", "diffuse", "decay", "neighborhood", "block_size")}
    sim = UnifiedSim(map_array, sff, episodes, N, mode="trained", learn="none", params=p, seed=seed, device=device)
    sim.load_h_dict(h_table)
    sim.place(np.full(episodes, N, np.int32), exit_pos=exit_pos, radius=radius)
    sim.rollout(max_steps)
    steps = sim.counters()[0]
    left = sim.get_positions()[1]
    sim.close()
    inside = (left == 0) & (steps >= 2 * N - 1) & (steps <= 2 * N + 14)
    return steps, float(inside.mean())
