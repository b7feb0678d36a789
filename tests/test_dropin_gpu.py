"""The drop-in class (ffm_b200/model/ffm_core.py) driven the way the reference's main.py drives
model/ffm_core.py (main.py:39-46), compared with the oracle step by step."""
import os

import numpy as np
import pytest

from helpers import MARGIN_GUARD
from oracle import assets, ffm_numpy
from oracle.inject import PhiloxSource

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nbh,dtype,hw,N", [("neumann", np.float32, 12, 50), ("moore", np.float64, 50, 100)])
def test_main_py_loop(cuda_device, tmp_path, nbh, dtype, hw, N):
    from ffm_b200.model.ffm_core import FloorFieldModel

    m = assets.room_map(hw, hw).astype(np.int64)          # np.load(config["map"]) is int64 (main.py:34)
    sff = assets.sff_norm_min(m, "L1", dtype)
    p = os.path.join(tmp_path, "sff.npy")
    np.save(p, sff)
    np.random.seed(42)
    model = FloorFieldModel(m, p, N, {"neighborhood": nbh, "seed": 777, "k_A": 3, "max_steps": 500})
    assert model.positions.shape == (N, 2) and model.positions.dtype == np.int64
    assert model.dff.shape == m.shape and model.dff.dtype == np.float32 and not model.dff.any()
    assert model.map_array.dtype == np.uint8 and len(model.neighbors) == (4 if nbh == "neumann" else 8)
    # same global-generator placement as the reference constructor (ffm_core.py:24-25)
    np.random.seed(42)
    free = np.argwhere(m == 0)
    assert np.array_equal(model.positions, free[np.random.choice(len(free), N, replace=False)])

    o = ffm_numpy.CoreOracle(m, sff, model.positions, model.params, PhiloxSource(777, 0))
    log, step = [], 0
    while model.positions.shape[0] > 0:                     # main.py:44-46
        model.step()
        log.append(np.copy(model.positions))
        step += 1
        o.step()
        if o.min_margin >= MARGIN_GUARD:
            assert np.array_equal(log[-1], o.positions), f"step {step}"
            if step % 25 == 0:
                assert np.array_equal(model.dff.view(np.uint32), o.dff.view(np.uint32))
    assert o.positions.shape[0] == 0 or o.min_margin < MARGIN_GUARD


def test_run_and_attribute_assignment(cuda_device, tmp_path):
    from ffm_b200.model.ffm_core import FloorFieldModel

    m = assets.room_map(12, 12)
    sff = assets.sff_norm_min(m, "L1", np.float32)
    p = os.path.join(tmp_path, "sff.npy")
    np.save(p, sff)
    model = FloorFieldModel(m, p, 30, {"seed": 5})
    model.step()
    # run_trained_ffm.py:235-236 assigns positions and dff
    model.positions = np.array([[5, 5], [6, 6], [3, 2]])
    model.dff = np.zeros_like(model.map_array, dtype=np.float32)
    assert model.positions.shape == (3, 2) and not model.dff.any()
    model.update_dff()
    assert model.run() is None and model.positions.shape[0] == 0
    with pytest.raises(ValueError):
        model.positions = np.array([[0, 0]])                # a wall
    with pytest.raises(ValueError):
        FloorFieldModel(m, p, 500, {"seed": 5})            # more pedestrians than free cells (ffm_core.py:25)
    with pytest.raises(FileNotFoundError):
        FloorFieldModel(m, os.path.join(tmp_path, "missing.npy"), 3)
