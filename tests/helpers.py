"""Shared helpers of the parity tests (the checker side: may import oracle/)."""
import numpy as np

from oracle import assets, ffm_numpy
from oracle.inject import PhiloxSource

# A move draw closer than this to a CDF boundary may legitimately resolve differently between
# NumPy's SIMD exp and CUDA's expf/exp (<= 2 ulp each); episodes whose oracle run saw such a draw
# are excused from bit-exact comparison (and counted).  SURVEY.md section 7, first hard part.
MARGIN_GUARD = 2e-6


def random_positions(map_array, n, rng):
    free = np.argwhere(map_array == 0)
    sel = free[rng.choice(len(free), n, replace=False)]
    return sel.astype(np.int64)


def pack_positions(pos_list, n_max):
    B = len(pos_list)
    out = np.full((B, n_max, 2), -1, dtype=np.int32)
    n = np.zeros((B,), dtype=np.int32)
    for e, p in enumerate(pos_list):
        n[e] = len(p)
        out[e, :len(p)] = p
    return out, n


def oracle_core_episode(map_array, sff, pos0, params, seed, episode, max_steps=None, keep_dff=False):
    o = ffm_numpy.CoreOracle(map_array, sff, pos0, params, PhiloxSource(seed, episode))
    r = o.run(max_steps=max_steps, keep_dff=keep_dff)
    r["final_dff"] = o.dff.copy()
    r["final_positions"] = o.positions.copy()
    return r


def traj_to_cells(traj, width):
    return [(p[:, 0] * width + p[:, 1]).astype(np.int64) for p in traj]


import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CORE_FIXTURES = ["core_12x12_neumann_f32", "core_12x12_moore_f32_full", "core_50x50_moore_f64",
                 "core_50x50_neumann_f64", "core_20x20_moore_params", "core_16x24_moore_kd0"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["params"] = json.loads(str(g["params"]))
    counts = g["traj_counts"]
    offs = np.concatenate([[0], np.cumsum(counts)])
    g["traj_list"] = [g["traj"][offs[t]:offs[t + 1]].astype(np.int64) for t in range(len(counts))]
    return g


UNIFIED_FIXTURES = ["uni_critic_12x12", "uni_critic_moore_bs5", "uni_critic_20x20_f64", "uni_actor_eps",
                    "uni_actor_greedy", "uni_both", "uni_both_moore"]


def load_unified(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["params"] = json.loads(str(g["params"]))
    g["mode"] = str(g["mode"])
    g["v_from"] = str(g["v_from"]) if "v_from" in g else ""
    E = int(g["episodes"]) if "episodes" in g else 1
    g["ep"] = []
    for ep in range(E):
        counts = g[f"counts_{ep}"]
        offs = np.concatenate([[0], np.cumsum(counts)])
        g["ep"].append(dict(pos0=g[f"pos0_{ep}"].astype(np.int64),
                            traj=[g[f"traj_{ep}"][offs[t]:offs[t + 1]].astype(np.int64) for t in range(len(counts))]))
    return g


MCQ_FIXTURES = ["mcq_12x12_penalties", "mcq_12x12_default", "mcq_20x20_kq"]


def load_mcq(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["params"] = json.loads(str(g["params"]))
    g["ep"] = []
    for ep in range(len(g["betas"])):
        counts = g[f"counts_{ep}"]
        offs = np.concatenate([[0], np.cumsum(counts)])
        g["ep"].append(dict(pos0=g[f"pos0_{ep}"].astype(np.int64),
                            traj=[g[f"traj_{ep}"][offs[t]:offs[t + 1]].astype(np.int64) for t in range(len(counts))]))
    return g


PRETRAIN_FIXTURES = ["mcq_pretrain_12x12", "mcq_pretrain_9x14"]


def load_pretrain(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["params"] = json.loads(str(g["params"]))
    return g


LEGACY_AC_FIXTURES = ["legacy_ac_12x12", "legacy_ac_moore_f64", "legacy_ac_12x12_full"]


def load_legacy(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["params"] = json.loads(str(g["params"]))
    g["ep"] = []
    for ep in range(int(g["episodes"])):
        counts = g[f"counts_{ep}"]
        offs = np.concatenate([[0], np.cumsum(counts)])
        g["ep"].append(dict(pos0=g[f"pos0_{ep}"].astype(np.int64),
                            traj=[g[f"traj_{ep}"][offs[t]:offs[t + 1]].astype(np.int64) for t in range(len(counts))]))
    return g
LEGACY_ACTOR_FIXTURES = ["legacy_actor_12x12", "legacy_actor_eps_moore", "legacy_actor_crowded"]
