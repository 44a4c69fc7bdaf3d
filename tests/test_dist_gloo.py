"""world_size-2 gloo test (CPU) of the multi-GPU host logic: batch sharding and the flat gradient all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpgcn_b200 import dist as mdist


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 64, 1001):
        for world in (1, 2, 3, 8):
            spans = [mdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    r, w = mdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)                        # identical weights everywhere
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    extra = torch.nn.Parameter(torch.ones(3))   # a parameter that receives no gradient on any rank
    data = torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10
    lo, hi = mdist.shard_range(8, rank, world)
    model(data[lo:hi]).sum().backward()         # local gradient of the local shard
    n = mdist.allreduce_mean_gradients(list(model.parameters()) + [extra])
    got = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    q.put((rank, n, got, extra.grad.clone()))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # expected: mean over ranks of the per-shard gradients == (sum over all 8 rows) / 2
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    data = torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10
    model(data).sum().backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()]) / 2
    for rank, n, got, extra_grad in results:
        assert n == want.numel() + 3
        torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)
        assert float(extra_grad.abs().sum()) == 0.0
