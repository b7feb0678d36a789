"""The C-ABI shared library builds, loads without a GPU and exports exactly what include/ffm_b200.h
declares; the ctypes mirror covers every declared function (no compute calls here)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "ffm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ffm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(ffm_lib):
    names = _declared()
    assert "ffm_rollout" in names and len(names) >= 10
    for n in names:
        assert hasattr(ffm_lib, n), f"{n} declared in include/ffm_b200.h but not exported"


def test_ctypes_mirror_covers_header():
    from ffm_b200 import _abi
    assert sorted(_abi.SIGNATURES) == _declared()


def test_abi_version_and_struct_sizes(ffm_lib):
    from ffm_b200 import _abi
    assert ffm_lib.ffm_abi_version() == _abi.ABI_VERSION
    assert ctypes.sizeof(_abi.Config) == 88 + 16 + 10 * 8 + 24   # core block + 4 x int32 + 10 x double + MCQ block
    assert ctypes.sizeof(_abi.Draws) == 32 and ctypes.sizeof(_abi.RolloutOut) == 80
    assert ctypes.sizeof(_abi.LegacyConfig) == 12 * 4 + 3 * 8 + 4 * 4 + 9 * 8 + 8 + 8      # ffm_legacy_config_t


def test_no_cpu_fallback_without_gpu():
    """Creating a simulation without a CUDA device must fail loudly."""
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ffm_b200 import BatchSim
    m = np.zeros((5, 5), np.uint8)
    with pytest.raises(RuntimeError):
        BatchSim(m, np.zeros((5, 5), np.float32), 1, 1)
    from ffm_b200.legacy import LegacySim
    with pytest.raises(RuntimeError):
        LegacySim(m, np.zeros((5, 5), np.float32), 1, 1)


def test_product_does_not_import_oracle():
    """Nothing under ffm_b200/ may reference oracle/ (the oracle is a checker, not a code path)."""
    pkg = os.path.join(ROOT, "ffm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "oracle/" not in txt or f.endswith(".md"), f
