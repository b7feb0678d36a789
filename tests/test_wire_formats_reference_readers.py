"""The reference's OWN inspection scripts read files written by the CUDA path.

tests/golden/gpu_wire_* were produced on a B200 by tests/make_gpu_wire_fixtures.py (drop-in classes driven like main.py /
the training drivers).  Here -- in the container that holds /root/reference -- the unmodified inspect_trajectory.py and
q_inspect.py consume them; where the reference is absent (the GPU box) the format checks still run."""
import importlib
import io
import os
import pickle
import sys
from contextlib import redirect_stdout

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REF = "/root/reference"
needs_reference = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")


def _ref_module(name):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    return importlib.import_module(name)


def test_positions_npy_has_the_main_py_shape():
    log = np.load(os.path.join(GOLDEN, "gpu_wire_positions.npy"), allow_pickle=True)     # main.py:52
    assert log.dtype == object and len(log) > 20
    n = [len(p) for p in log]
    assert n[0] <= 25 and n[-1] == 0 and all(a >= b for a, b in zip(n, n[1:]))           # people only ever leave
    assert all(p.dtype == np.int64 and p.ndim == 2 and p.shape[1] == 2 for p in log if len(p))


@needs_reference
def test_reference_inspect_trajectory_reads_the_gpu_npz():
    mod = _ref_module("inspect_trajectory")
    mod.TRAJECTORY_PATH = os.path.join(GOLDEN, "gpu_wire_trajectory_N12_ep00001_total00001.npz")
    out = io.StringIO()
    with redirect_stdout(out):
        mod.main()                                                                       # inspect_trajectory.py:12-76, unmodified
    text = out.getvalue()
    assert "positions" in text and "12" in text and "(12, 2)" in text                    # first step: 12 agents x (row, col)
    z = np.load(mod.TRAJECTORY_PATH, allow_pickle=True)
    assert f"{len(z['positions'])}" in text and int(z["steps"]) == len(z["positions"])


@needs_reference
def test_reference_q_inspect_reads_the_gpu_q_pickle():
    mod = _ref_module("q_inspect")
    summary, per_block, crowd = mod.analyze_q(os.path.join(GOLDEN, "gpu_wire_Q.pkl"))    # q_inspect.py:52-139, unmodified
    with open(os.path.join(GOLDEN, "gpu_wire_Q.pkl"), "rb") as f:
        Q = pickle.load(f)
    assert int(summary["n_states_total"][0]) == len(Q) > 100 and int(summary["action_size"][0]) == 5
    assert mod._schema(next(iter(Q))) == "tuple_block2"                                   # (bytes, (bx, by)) keys
    assert set(zip(per_block["block_x"], per_block["block_y"])) <= {(a, b) for a in range(4) for b in range(4)}
    assert int(per_block["n_states"].sum()) == len(Q)


def test_v_pickle_is_the_driver_format():
    with open(os.path.join(GOLDEN, "gpu_wire_V.pkl"), "rb") as f:                         # run_unified_critic_training.py:290-299
        V = pickle.load(f)
    assert isinstance(V, dict) and len(V) > 20
    for (ranks, blk), v in V.items():
        assert len(ranks) == 4 and all(0 <= r <= 3 for r in ranks) and len(blk) == 2 and isinstance(v, float)


def test_legacy_table_pickles_are_the_driver_format():
    """run_critic_training.py:219-226 / run_actor_only_training.py:293-299: dicts keyed by pickle.dumps((state_13, (bx, by)))."""
    with open(os.path.join(GOLDEN, "gpu_wire_legacy_V.pkl"), "rb") as f:
        V = pickle.load(f)
    with open(os.path.join(GOLDEN, "gpu_wire_legacy_H_actor.pkl"), "rb") as f:
        Ht = pickle.load(f)
    assert len(V) > 100 and len(Ht) > 10
    for k, v in V.items():
        cells, blk = pickle.loads(k)
        assert len(cells) == 13 and all(0 <= c <= 3 for c in cells) and cells[4] == 1 and len(blk) == 2 and isinstance(v, float)
    for k, row in Ht.items():
        cells, blk = pickle.loads(k)
        assert len(cells) == 13 and cells[4] == 1 and isinstance(row, list) and len(row) == 5


@needs_reference
def test_reference_inspect_h_actor_reads_the_gpu_h_pickle(tmp_path):
    """inspect_h_actor_formatted.py (a script with a hard-wired relative path) run unmodified, from a directory that holds the
    GPU-written H table under that path; its report must list every state without a decode error."""
    import runpy
    import shutil
    src = open(os.path.join(REF, "inspect_h_actor_formatted.py")).read()
    date = src.split('date = "')[1].split('"')[0]
    name = src.split('file_name = "')[1].split('"')[0]
    d = os.path.join(tmp_path, "output", "logs", "actor_only_training", f"run_{date}")
    os.makedirs(d)
    shutil.copy(os.path.join(GOLDEN, "gpu_wire_legacy_H_actor.pkl"), os.path.join(d, name))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with redirect_stdout(io.StringIO()):
            runpy.run_path(os.path.join(REF, "inspect_h_actor_formatted.py"), run_name="__main__")
    finally:
        os.chdir(cwd)
    report = open(os.path.join(tmp_path, "H_actor_analysis.txt"), encoding="utf-8").read()
    with open(os.path.join(GOLDEN, "gpu_wire_legacy_H_actor.pkl"), "rb") as f:
        n = len(pickle.load(f))
    assert f"総状態数: {n}" in report and f"ソート済み状態数: {n}" in report and "デコードエラー" not in report


@needs_reference
def test_reference_actor_class_loads_the_gpu_legacy_v_pickle(tmp_path):
    """The UNMODIFIED legacy actor class reads the GPU-written critic table as its pretrained_v_path (ffm_actor_only.py:56-69)."""
    mod = _ref_module("model.ffm_actor_only")
    m = np.zeros((12, 12), np.uint8); m[0, :] = 2; m[-1, :] = 2; m[:, 0] = 2; m[:, -1] = 2; m[0, 6] = 3
    p = os.path.join(tmp_path, "sff.npy")
    np.save(p, np.zeros((12, 12), np.float32))
    with redirect_stdout(io.StringIO()):
        model = mod.FloorFieldModelActorOnly(m, p, 1, pretrained_v_path=os.path.join(GOLDEN, "gpu_wire_legacy_V.pkl"))
    with open(os.path.join(GOLDEN, "gpu_wire_legacy_V.pkl"), "rb") as f:
        assert model.initial_v_size == len(pickle.load(f))
