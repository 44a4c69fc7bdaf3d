#!/usr/bin/env python
"""A/B of MPGCN_B200_LSTM_POLY = 0 / 1 / 2 (exponentials per unit and step on the FMA pipe instead of the SFU): per-launch
times of the tcgen05 LSTM forward (training) and backward at the bench shape, and the error against the fp32 CUDA-core LSTM.
The knob is read once per process, so every setting runs in a child.   python tools/lstm_poly_probe.py > gpurun_out/lstm_poly.jsonl"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    import torch
    sys.path.insert(0, ROOT)
    from mpgcn_b200 import _lib, ops
    B, T, N = int(os.environ.get("PROBE_B", "8")), 12, 1000
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    lstm = torch.nn.LSTM(1, 32, 1, batch_first=True).to(dev)
    ws = [w.detach().requires_grad_(True) for w in (lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0)]
    x = torch.rand(B, T, N, N, 1, device=dev) * 6
    d = torch.randn(B * N * N, 32, device=dev) * 1e-6
    lib = _lib.load()
    for _ in range(2):
        h = ops.lstm_last(x, *ws, precision="fp16")
        h.backward(d)
    torch.cuda.synchronize()
    lib.mpgcn_profile_reset()
    lib.mpgcn_profile_enable(1)
    for _ in range(5):
        for w in ws:
            w.grad = None
        h = ops.lstm_last(x, *ws, precision="fp16")
        h.backward(d)
    torch.cuda.synchronize()
    lib.mpgcn_profile_enable(0)
    prof = _lib.profile_read()
    g16 = [w.grad.clone() for w in ws]
    h16 = h.detach().clone()
    # reference: fp32 CUDA-core LSTM on a slice (it is ~10x slower)
    xs, ds = x[:1, :, :200].contiguous(), d[:200 * N].contiguous()
    for w in ws:
        w.grad = None
    h32 = ops.lstm_last(xs, *ws, precision="fp32")
    h32.backward(ds)
    g32 = [w.grad.clone() for w in ws]
    for w in ws:
        w.grad = None
    h16s = ops.lstm_last(xs, *ws, precision="fp16")
    h16s.backward(ds)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print(json.dumps({"poly": int(os.environ.get("MPGCN_B200_LSTM_POLY", "0")), "pack": int(os.environ.get("MPGCN_B200_LSTM_PACK", "0")), "batch": B,
                      "fwd_ms": prof["LSTM_FWD"]["ms"] / prof["LSTM_FWD"]["launches"], "bwd_ms": prof["LSTM_BWD"]["ms"] / prof["LSTM_BWD"]["launches"],
                      "hT_rel_linf_vs_fp32": rel(h16s.detach(), h32.detach()),
                      "grad_rel_linf_vs_fp32": [rel(w.grad, g) for w, g in zip(ws, g32)]}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        variants = [(0, 0), (1, 1), (2, 1), (3, 1)] if os.environ.get("PROBE_PACK_ONLY") else [(0, 0), (1, 0), (2, 0), (0, 1), (1, 1)]
        for poly, pack in variants:
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"],
                           env=dict(os.environ, MPGCN_B200_LSTM_POLY=str(poly), MPGCN_B200_LSTM_PACK=str(pack)))
