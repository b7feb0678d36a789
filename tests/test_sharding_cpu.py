"""Multi-rank host logic on CPU (gloo, world_size 2): episode partition and the delta all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ffm_b200 import sharding


def test_shard_range_partitions_exactly():
    for total in (1, 7, 8, 4096, 4099):
        for ws in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (a, ca), (b, _) in zip(spans, spans[1:]):
                assert a + ca == b
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _worker(rank, ws, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        first, count = sharding.shard_range(10, None, None)
        S, A = 64, 5
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.zeros((3 + A) * S, dtype=torch.float64)          # [dV | dN | dF | dH], one collective
        flat[:S] = torch.rand(S, dtype=torch.float64, generator=g)
        flat[3 * S:] = torch.rand(S * A, dtype=torch.float64, generator=g)
        flat[2 * S + rank:3 * S:3] = 1.0                               # key-touched marks of this rank
        work = sharding.allreduce_flat(flat, async_op=True)
        work.wait()
        out.put((rank, first, count, flat[:S].numpy().copy(), flat[3 * S:].numpy().copy(),
                 (flat[2 * S:3 * S] > 0).numpy().astype(np.uint8)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_delta_allreduce_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (res[0][1], res[0][2]) == (0, 5) and (res[1][1], res[1][2]) == (5, 5)
    want_v = sum(torch.rand(64, dtype=torch.float64, generator=torch.Generator().manual_seed(100 + r)).numpy() for r in range(2))
    assert np.allclose(res[0][3], want_v) and np.array_equal(res[0][3], res[1][3])
    assert np.array_equal(res[0][4], res[1][4])
    want_seen = np.zeros(64, np.uint8); want_seen[0::3] = 1; want_seen[1::3] = 1
    assert np.array_equal(res[0][5], want_seen) and np.array_equal(res[1][5], want_seen)


class _FakeMcqSim:
    """Stand-in for McqSim on CPU tensors: the by-key exchange logic of McqBatchedLearner.sync without a GPU."""
    learn = "batched"

    def __init__(self, rank):
        self.rank, self.table, self.folded = rank, {}, 0
        # rank r touched keys {r, r + 1, ..., r + 3 + r}: ragged lists, one key shared between the ranks
        self.local = {100 + k: np.full(10, float(10 * rank + k)) for k in range(rank, rank + 4 + rank)}

    def delta_device(self):
        return "cpu"

    def accumulate(self):
        pass

    def export_deltas(self, capacity, out=None):
        out.zero_()
        count = out[:1].view(torch.int32)[:1]
        keys = out[1:1 + capacity].view(torch.int64)
        rows = out[1 + capacity:].view(capacity, 10)
        for i, (k, v) in enumerate(sorted(self.local.items())):
            keys[i] = k
            rows[i] = torch.from_numpy(v)
        count[0] = len(self.local)
        return keys, rows, count

    def import_deltas(self, keys, rows, count, count_dev=None):
        n = min(int(count), int(count_dev[0])) if count_dev is not None else int(count)
        for i in range(n):
            k = int(keys[i])
            self.table[k] = self.table.get(k, np.zeros(10)) + rows[i].numpy()

    def fold(self):
        self.folded += 1


def _mcq_worker(rank, ws, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from ffm_b200.mcq_training import McqBatchedLearner
        sim = _FakeMcqSim(rank)
        McqBatchedLearner(sim, export_capacity=16).sync()
        out.put((rank, sorted((k, v.tolist()) for k, v in sim.table.items()), sim.folded))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_mcq_by_key_exchange_gloo_world2():
    """Ragged per-rank lists of touched (key, sums) rows: all-gathered padded, imported in rank order on every rank."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_mcq_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2] == 1
    table = dict(res[0][1])
    assert sorted(table) == [100, 101, 102, 103, 104, 105]
    assert table[100] == [0.0] * 10 and table[101] == [1.0 + 11.0] * 10 and table[105] == [15.0] * 10
