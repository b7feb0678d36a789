"""bench.py output contract: one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _run(args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-budget", "3"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["metric"] == "pedestrian_steps_per_sec"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                       text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_line(cuda_device):
    d = _run(["--episodes", "296", "--steps", "2", "--warmup", "3", "--cpu-budget", "2"])
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline", "cpu_baseline"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["gpu_launches"] >= 2 * d["steps"] and d["value"] > 1e9 and d["e2e"]["value"] > 1e9
    assert d["e2e"]["h2d_bytes_per_step"] == 296 * 1024 * 8 + 296 * 4 and d["e2e"]["d2h_bytes_per_step"] == 296 * 12
    rf = d["roofline"]
    # C2 keeps its episode state in shared memory: the bound is the measured SMEM bandwidth, the HBM view rides along
    assert rf["bound"] == "smem" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["peak"] > 1e4 and abs(rf["hbm"]["frac"] - rf["achieved"] / rf["hbm"]["peak"]) < 1e-9
    assert "secondary" not in d      # --episodes given: headline only
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])


@pytest.mark.gpu
@pytest.mark.parametrize("workload,metric", [("c2traj", "pedestrian_steps_per_sec"), ("sff", "sff_cells_per_sec"), ("c4", "pedestrian_steps_per_sec"),
                                             ("c4train", "training_episodes_per_sec"), ("c5train", "training_episodes_per_sec"),
                                             ("legacy", "pedestrian_steps_per_sec")])
def test_gpu_arm_secondary_workloads(cuda_device, workload, metric):
    """Each secondary workload of the default run, stand-alone and shortened."""
    extra = {"c2traj": ["--episodes", "32"], "c4": ["--episodes", "512"], "sff": [], "c4train": [], "c5train": [], "legacy": []}[workload]
    d = _run(["--workload", workload, "--steps", "1", "--warmup", "3", "--no-cpu"] + extra)
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline"} <= set(d) and d["metric"] == metric and d["value"] > 0
    assert d["e2e"]["value"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
    if workload == "c2traj":
        assert not d["record"]["overflowed"] and 4.0 <= d["record"]["bytes_per_ped_step"] < 4.1
    if workload == "c4":
        assert d["config"]["syncs_timed"] == 16 and d["sync_ms_blocking"] > 0
    if workload == "c5train":
        assert d["config"]["pretrain_patterns"] == 11328 and d["tables"]["Q_rows_after_training"] > d["tables"]["Q_rows_after_pretrain"] > 1000
    if workload == "legacy":
        assert d["episodes_per_sec"] > 10 and d["actor_only"]["value"] > 0 and d["frozen_batch"]["value"] > 1e8
        assert d["tables"]["V_states"] > 100 and d["tables"]["actor_H_rows"] > 10
    if workload == "c4train":
        assert d["value"] > 1000 and all(v["in_band_2N-1..2N+14"] >= 0.9 for v in d["acceptance"].values()), d["acceptance"]
