"""Multi-GPU plumbing: episodes shard over ranks with no data-path collective; the learning
configurations add ONE exchange step per sync -- a single all-reduce of the flat delta buffer.

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


def allreduce_flat(flat, async_op=False):
    """Sum ONE flat delta buffer ([dV | dN | dF | dH], float64) over all ranks, in place.  The key-touched marks dF
    ride along as numbers (> 0 = touched on some rank), so the whole exchange is a single collective.
    Returns the work handle when async_op (None on a single rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    return dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=async_op)


class BatchedLearner:
    """Synchronous batched TD learning of the unified model (BASELINE config 4).

    Each sync: every rank rolls out its B episodes against the frozen tables (kernel mode FFM_LEARN_BATCHED
    accumulates TD errors, visit counts, key marks and alpha_h*delta into the flat delta buffer), the buffer is
    all-reduced over ranks -- ONE collective -- and every rank applies the same update, so all ranks hold identical
    tables without ever broadcasting them.  This is a different algorithm from the reference's sequential per-agent
    updates ("sequential learning semantics"); it is judged statistically, while FFM_LEARN_EXACT
    reproduces the reference bit for bit on one episode.

    overlap=True pipelines the exchange: the all-reduce of chunk k runs on NCCL's stream while chunk k+1 rolls out
    into a second buffer; chunk k's deltas are applied right after rollout k+1 (tables lag by one chunk --
    stale-synchronous with staleness 1, deterministic and identical on every rank).  Call flush() at the end.
    """

    def __init__(self, sim, distributed=True, overlap=False):
        assert sim.learn == "batched"
        self.sim = sim
        self.distributed = distributed      # False: never touch the process group (single-rank reference runs)
        self.overlap = overlap
        self.bufs = [sim.delta, sim.new_delta_buffer()] if overlap else [sim.delta]
        self._cur = 0
        self._pending = None                # (work handle or None, buffer index) of the exchange in flight

    def _apply(self, idx):
        self.sim.bind_deltas(self.bufs[idx])
        self.sim.apply_deltas()

    def sync(self):
        """Exchange + fold in.  Call after rollout().  No host synchronisation: everything is stream-ordered."""
        if not self.overlap:
            if self.distributed:
                allreduce_flat(self.sim.delta)
            self.sim.apply_deltas()
            return
        cur = self._cur
        work = allreduce_flat(self.bufs[cur], async_op=True) if self.distributed else None
        self.flush()                                   # the previous chunk's deltas: reduced while this chunk rolled out
        self._pending = (work, cur)
        self._cur = 1 - cur
        self.sim.bind_deltas(self.bufs[self._cur])     # zeroed by its last apply

    def flush(self):
        """Apply the exchange still in flight (overlap mode)."""
        if self._pending is not None:
            work, idx = self._pending
            if work is not None:
                work.wait()                            # makes the current stream wait; the host does not block
            self._apply(idx)
            self._pending = None
            self.sim.bind_deltas(self.bufs[self._cur])

    def round(self, pos_rc, n, max_steps, sync_every=None):
        """One batch of episodes.  sync_every = K folds the deltas in every K CA steps (value information then
        travels K-step-wise within an episode, closer to the reference's per-agent updates); None = once,
        at the end.  The number of syncs is fixed by max_steps so that all ranks stay in lock-step."""
        self.sim.set_positions(pos_rc, n)
        if sync_every is None:
            self.sim.rollout(max_steps)
            self.sync()
        else:
            done = 0
            while done < max_steps:
                k = min(sync_every, max_steps - done)
                self.sim.rollout(k)
                self.sync()
                done += k
        if self.overlap:
            self.flush()
        return self.sim.counters()
