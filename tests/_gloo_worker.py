"""Worker of tests/test_dist_gloo.py (launched once per rank with RANK / WORLD_SIZE / MASTER_* set)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mpgcn_b200 import dist as mdist  # noqa: E402


def main(out_path):
    rank, world = mdist.init_from_env("gloo")
    torch.manual_seed(0)                        # identical weights everywhere
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    extra = torch.nn.Parameter(torch.ones(3))   # a parameter that receives no gradient on any rank
    data = torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6) / 10
    lo, hi = mdist.shard_range(8, rank, world)
    model(data[lo:hi]).sum().backward()         # local gradient of the local shard
    n = mdist.allreduce_mean_gradients(list(model.parameters()) + [extra])
    got = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    torch.save({"rank": rank, "n": n, "got": got, "extra": extra.grad.clone()}, out_path)
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
