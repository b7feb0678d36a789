"""Calibration of the KS statistic of tests/test_distribution_gpu.py (c2_dff): GPU under two seeds, the C oracle under
the second seed; same-seed GPU vs oracle episodes should agree except for knife-edge draws."""
import os, sys
sys.path.insert(0, '.')
import numpy as np
from scipy import stats
import bench
from ffm_b200 import BatchSim
from oracle import c_oracle

m = bench.room_map(64, 64); sff = bench.sff_room(m, "moore")
n_ep, N = 2048, 1024
params = {"k_S": 3, "k_D": 1, "neighborhood": "moore"}
n = np.full((n_ep,), N, np.int32)
def gpu(seed, base):
    sim = BatchSim(m, sff, n_ep, N, params, seed=seed, episode_base=base, track_dff=True)
    sim.set_positions(bench.place(m, N, n_ep, base, seed), n)
    sim.rollout(4096)
    s, _ = sim.counters(); sim.close(); return s
def cpu(seed, base):
    r = c_oracle.run_core_batch(m, sff, bench.place(m, N, n_ep, base, seed), n, params, seed=seed, episode_base=base, max_steps=4096,
                                threads=os.cpu_count() or 1, track_dff=True, want_state=True)
    return r["steps"], r["min_margin"]
gA = gpu(0xA11CE, 0); gB = gpu(0xB0B, 50_000); gC = gpu(0xC0C, 100_000)
cB, mB = cpu(0xB0B, 50_000)
cA, mA = cpu(0xA11CE, 0)
ks = lambda a, b: float(stats.ks_2samp(a, b).statistic)
print("means", gA.mean(), gB.mean(), gC.mean(), cA.mean(), cB.mean(), "stds", gA.std(), gB.std(), cB.std())
print("KS gA-gB", ks(gA, gB), "gA-cB", ks(gA, cB), "gB-cB", ks(gB, cB), "gA-gC", ks(gA, gC), "gB-gC", ks(gB, gC), "cA-cB", ks(cA, cB), "gA-cA", ks(gA, cA))
print("same-seed equal fraction B", (gB == cB).mean(), "A", (gA == cA).mean(), "knife B", (mB < 2e-6).mean())
