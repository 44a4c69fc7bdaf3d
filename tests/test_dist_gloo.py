"""world_size-2 gloo test (CPU) of the multi-GPU host logic: batch sharding and the flat gradient all-reduce."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from mpgcn_b200 import dist as mdist


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 64, 1001):
        for world in (1, 2, 3, 8):
            spans = [mdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_flat_gradient_allreduce_world2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_gloo_worker.py")
    procs = []
    for r in range(2):
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(r), WORLD_SIZE="2")
        procs.append(subprocess.Popen([sys.executable, worker, str(tmp_path / f"r{r}.pt")], env=env))
    for p in procs:
        assert p.wait(timeout=180) == 0
    results = [torch.load(tmp_path / f"r{r}.pt") for r in range(2)]
    # expected: mean over ranks of the per-shard gradients == (sum over all 8 rows) / 2
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    data = torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10
    model(data).sum().backward()
    want = torch.cat([p.grad.reshape(-1) for p in model.parameters()]) / 2
    for res in results:
        assert res["n"] == want.numel() + 3
        torch.testing.assert_close(res["got"], want, rtol=1e-6, atol=1e-6)
        assert float(res["extra"].abs().sum()) == 0.0
