#!/usr/bin/env python
"""Launch-bound sizes: eager training step vs the CUDA-graph captured step (mpgcn_b200.graph_step) at the reference's real
N = 47 (Data_Container_OD.py:16) and at N = 200.  Prints one JSON line per size (CUDA events, 50 steps after 10 warm-up).

    python tools/small_n_bench.py > gpurun_out/small_n.jsonl
"""
import json
import os
import sys

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import MPGCN as shim  # noqa: E402
from mpgcn_b200 import _lib  # noqa: E402
from mpgcn_b200.graph_step import GraphedTrainStep  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    lib = _lib.load()
    for N, K, T, B in ((47, 3, 7, 4), (200, 3, 8, 4)):
        torch.manual_seed(0)
        model = shim.MPGCN(M=2, K=K, input_dim=1, lstm_hidden_dim=32, lstm_num_layers=1, gcn_hidden_dim=32, gcn_num_layers=3,
                           num_nodes=N, user_bias=True, activation=nn.ReLU).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True)
        crit = nn.MSELoss()
        x = torch.rand(B, T, N, N, 1, device=dev) * 6
        y = torch.rand(B, 1, N, N, 1, device=dev) * 6
        G = torch.randn(K, N, N, device=dev) / N ** 0.5
        go, gd = torch.randn(B, K, N, N, device=dev) / N ** 0.5, torch.randn(B, K, N, N, device=dev) / N ** 0.5

        def eager():
            loss = crit(model(x_seq=x, G_list=[G, (go, gd)]), y)
            opt.zero_grad(set_to_none=False)
            loss.backward()
            opt.step()

        def timed(fn, n=50, warm=10):
            for _ in range(warm):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n

        lib.mpgcn_profile_reset()
        eager()
        launches = sum(v["launches"] for t, v in _lib.profile_read().items() if t not in _lib.REGION_TAGS)
        t_eager = timed(eager)
        step = GraphedTrainStep(model, crit, opt, example=(x, y, G, (go, gd)))
        t_graph = timed(lambda: step(x, y, go, gd))
        t_par = None
        try:        # the two branches forked onto parallel streams inside the captured graph
            step2 = GraphedTrainStep(model, crit, opt, example=(x, y, G, (go, gd)), branch_streams=True)
            t_par = timed(lambda: step2(x, y, go, gd))
        except Exception as e:
            print(f"branch-parallel capture failed at N={N}: {type(e).__name__}: {str(e)[:200]}", file=sys.stderr)
        model.branch_streams = None
        print(json.dumps({"N": N, "K": K, "T": T, "batch": B, "library_kernels_per_step": launches, "eager_ms_per_step": round(t_eager, 4),
                          "graphed_ms_per_step": round(t_graph, 4), "graphed_branch_parallel_ms_per_step": None if t_par is None else round(t_par, 4),
                          "od_cells_per_s_graphed": B * T * N * N / (min(t_graph, t_par or t_graph) * 1e-3)}), flush=True)


if __name__ == "__main__":
    main()
