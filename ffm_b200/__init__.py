"""ffm_b200 -- B200-native (sm_100a) implementation of the Floor-Field-Model hot path.

Layout
  csrc/            hand-written CUDA kernels + the C ABI of include/ffm_b200.h (libffm_b200.so)
  _abi.py          ctypes mirror of the header
  sim.py           BatchSim / UnifiedSim / McqSim: B independent episodes on one map (the batched rollout API)
  model/           drop-in classes with the reference's model/ffm_*.py interface
  unified_training.py, mcq_training.py   batched forms of the reference's training drivers (curricula, coverage pretrain)
  sharding.py      episode partition over ranks, the NCCL exchange of the learning configurations
  sff.py           static-floor-field generation (Create_SFF.py + geodesic modes)

The CUDA library is the only compute path; importing this package never falls back to a CPU
implementation.
"""
from .sim import BatchSim, McqSim, UnifiedSim  # noqa: F401

__all__ = ["BatchSim", "McqSim", "UnifiedSim"]
