"""Multi-GPU plumbing: episodes shard over ranks with no data-path collective; the learning
configurations add ONE exchange step per sync -- the all-reduce of the table deltas.

One process per GPU (torch.distributed, NCCL over NVLink); every function also works on CPU tensors
with the gloo backend, which is how tests/test_sharding_cpu.py exercises the logic without GPUs.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(total, rank=None, world_size=None):
    """Contiguous range of GLOBAL episode ids owned by `rank`: (first, count).  Draw keys use the global
    id, so per-episode results do not depend on the number of ranks."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(int(total), int(world_size))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def allreduce_deltas(deltas, seen_flags=()):
    """Sum the delta tables (dV [S], dH [S, A]) and OR the key-present flags (uint8) over all ranks,
    in place.  No-op on a single rank.  The flags ride along as one MAX all-reduce."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    handles = [dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True) for t in deltas if t is not None]
    handles += [dist.all_reduce(f, op=dist.ReduceOp.MAX, async_op=True) for f in seen_flags if f is not None]
    for h in handles:
        h.wait()


class BatchedLearner:
    """Synchronous batched TD learning of the unified model (BASELINE config 4).

    Each sync: every rank rolls out its B episodes against the frozen tables (kernel mode
    FFM_LEARN_BATCHED accumulates TD errors, visit counts and alpha_h*delta into dV / dN / dH with atomics), the deltas and key flags are
    all-reduced over ranks, and every rank applies the same update -- so all ranks hold identical
    tables without ever broadcasting them.  This is a different algorithm from the reference's
    sequential per-agent updates ("sequential learning semantics"); it is judged
    statistically, while FFM_LEARN_EXACT reproduces the reference bit for bit on one episode.
    """

    def __init__(self, sim, distributed=True):
        assert sim.learn == "batched"
        self.sim = sim
        self.distributed = distributed      # False: never touch the process group (single-rank reference runs)
        dev = sim.dV.device
        self._vseen = torch.zeros(sim.S, dtype=torch.uint8, device=dev)
        self._hseen = torch.zeros(sim.S, dtype=torch.uint8, device=dev)

    def sync(self):
        """All-reduce deltas + flags, apply.  Call after rollout()."""
        import ctypes as C
        from . import _abi
        from .sim import _ptr, _stream
        s = self.sim
        rank, ws = world() if self.distributed else (0, 1)
        if ws > 1:
            _abi.check(s._lib.ffm_tables_get(s._h, None, _ptr(self._vseen), None, _ptr(self._hseen), _abi.FFM_DEVICE, _stream()))
            allreduce_deltas([s.dV, s.dN, s.dH], [self._vseen, self._hseen])
            _abi.check(s._lib.ffm_tables_set(s._h, None, _ptr(self._vseen), None, _ptr(self._hseen), _abi.FFM_DEVICE, _stream()))
        s.apply_deltas()

    def round(self, pos_rc, n, max_steps, sync_every=None):
        """One batch of episodes.  sync_every = K folds the deltas in every K CA steps (value information then
        travels K-step-wise within an episode, closer to the reference's per-agent updates); None = once,
        at the end.  The number of syncs is fixed by max_steps so that all ranks stay in lock-step."""
        self.sim.set_positions(pos_rc, n)
        if sync_every is None:
            self.sim.rollout(max_steps)
            self.sync()
        else:
            for _ in range(0, max_steps, sync_every):
                self.sim.rollout(min(sync_every, max_steps))
                self.sync()
        return self.sim.counters()
