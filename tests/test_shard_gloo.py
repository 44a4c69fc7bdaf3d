"""world_size-2 gloo tests (CPU) of the model-parallel shards (mpgcn_b200/shard.py): the exchange logic -- reduce-scatter of the
partial pre-activation / all-gather of dPre (origin-row shard), all-reduce of pre and dX + row-sharded LSTM (K shard), the
plan-aware gradient reduction -- with the CUDA engine replaced by a torch stand-in (tests/shard_standin.py), against the numpy
oracle of the WHOLE model (reference MPGCN.py:89-112)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch
from torch import nn

from oracle import mpgcn_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_plan_partitions():
    from mpgcn_b200 import shard
    for world in (1, 2, 4, 8):
        rows = [shard.ShardPlan("row", r, world, 1000, 3) for r in range(world)]
        assert rows[0].row_lo == 0 and rows[-1].row_hi == 1000 and all(p.rows == 1000 // world for p in rows)
        ks = [shard.ShardPlan("k", r, world, 1000, 6) for r in range(world)]
        assert sum(p.Kd for p in ks) == 6 and ks[0].d_lo == 0 and ks[-1].d_hi == 6
    with pytest.raises(ValueError):
        shard.ShardPlan("row", 0, 3, 1000, 3)


@pytest.mark.parametrize("kind", ["row", "k", "rowhyb"])
def test_sharded_model_world2_matches_whole_model_oracle(kind, tmp_path):
    """row / k: world 2.  rowhyb: world 4 = 2 batch groups x 2 row ranks (the exchange stays inside a group; gradients are summed
    over all ranks and averaged over the groups)."""
    nproc = 4 if kind == "rowhyb" else 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    worker = os.path.join(HERE, "_shard_worker.py")
    procs = []
    for r in range(nproc):
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(r), WORLD_SIZE=str(nproc))
        procs.append(subprocess.Popen([sys.executable, worker, kind, str(tmp_path / f"r{r}.pt")], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0
    res = [torch.load(tmp_path / f"r{r}.pt") for r in range(nproc)]
    # the same model / inputs as the worker builds, evaluated whole by the oracle
    sys.path.insert(0, os.path.dirname(HERE))
    import MPGCN as shim
    N, K, T, B, hid = 8, 3, 3, 2, 8
    torch.manual_seed(0)
    model = shim.MPGCN(M=2, K=K, input_dim=1, lstm_hidden_dim=hid, lstm_num_layers=1, gcn_hidden_dim=hid, gcn_num_layers=3,
                       num_nodes=N, user_bias=True, activation=nn.ReLU)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.add_(0.05)
    rng = np.random.default_rng(1)
    x = (rng.random((B, T, N, N, 1)) * 4).astype(np.float32)
    y = rng.random((B, 1, N, N, 1)).astype(np.float32)
    G = (rng.random((K, N, N)) / N).astype(np.float32)
    go = (rng.random((B, K, N, N)) / N).astype(np.float32)
    gd = (rng.random((B, K, N, N)) / N).astype(np.float32)
    params = {k: v.detach().numpy().astype(np.float64) for k, v in model.state_dict().items()}
    GL = [G.astype(np.float64), (go.astype(np.float64), gd.astype(np.float64))]
    y_o = orc.mpgcn_forward(params, x.astype(np.float64), GL, M=2, gcn_num_layers=3)
    d_y = 2.0 * (y_o - y) / y.size
    _, grads_o = orc.mpgcn_forward_backward(params, x.astype(np.float64), GL, M=2, gcn_num_layers=3, d_y=d_y)
    loss_o = float(((y_o - y) ** 2).mean())
    if kind == "row":
        pred = np.concatenate([r["pred"].numpy() for r in sorted(res, key=lambda r: r["rank"])], axis=2)
    elif kind == "rowhyb":     # ranks (0,1) hold sample 0's row slabs, ranks (2,3) sample 1's
        by = sorted(res, key=lambda r: r["rank"])
        pred = np.concatenate([np.concatenate([by[2 * gi]["pred"].numpy(), by[2 * gi + 1]["pred"].numpy()], axis=2) for gi in range(2)], axis=0)
    else:
        pred = res[0]["pred"].numpy()
        assert np.array_equal(pred, res[1]["pred"].numpy()), "K shard: the prediction is replicated"
    assert max(orc.rel_errors(pred, y_o)) <= 1e-5
    for r in res:
        assert abs(r["loss"] - loss_o) <= 1e-5 * loss_o
        assert set(r["grads"]) == set(grads_o)
        for k, g in r["grads"].items():
            assert max(orc.rel_errors(g.numpy(), grads_o[k])) <= 2e-5, (kind, k)
