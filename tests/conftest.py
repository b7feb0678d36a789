import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ffm_lib():
    """Build (if stale) and load libffm_b200.so."""
    from ffm_b200 import _abi, build
    build.build()
    return _abi.lib()


@pytest.fixture(scope="session")
def cuda_device(ffm_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return 0


@pytest.fixture(params=["cell", "ped", "cl2", "cl4", "cl8", "cl4g"])
def core_kernel(request, monkeypatch):
    """Runs a test once per base-model kernel: "cell" = cell-centric (csrc/ffm_cell_kernel.cuh), "ped" =
    pedestrian-centric (csrc/ffm_core_kernel.cuh), "clN" = the cell-centric kernel as a thread-block cluster of N CTAs
    per episode (row bands in distributed shared memory); libffm_b200 reads FFM_KERNEL / FFM_CLUSTER in ffm_create."""
    if request.param.startswith("cl"):
        monkeypatch.setenv("FFM_KERNEL", "cell")
        monkeypatch.setenv("FFM_CLUSTER", request.param[2:3])
        if request.param.endswith("g"):
            monkeypatch.delenv("FFM_FIELDS_SMEM", raising=False)   # "cl4g": score + DFF stay in L2 (what ffm_create picks by itself)
        else:
            monkeypatch.setenv("FFM_FIELDS_SMEM", "1")  # fields in distributed shared memory where they fit (the L2 form is
                                                          # covered by test_fullsize_gpu / test_maps_beyond_one_sm_run_as_clusters)
    else:
        monkeypatch.setenv("FFM_KERNEL", request.param)
        monkeypatch.delenv("FFM_CLUSTER", raising=False)
    return request.param
