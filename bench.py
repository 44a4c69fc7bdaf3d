#!/usr/bin/env python
"""bench.py -- OD-cells/s of the MPGCN hot path (per-cell LSTM -> 3 x BDGCN -> FC, M=2 branches,
forward + backward) on synthetic OD tensors, on N GPUs of one node.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps K --warmup W        # CPU arm (oracle/torch_port.py)

Workload (BASELINE.json metric "OD-cells/sec (B*T*N^2) 2D-GCN fwd+bwd at N=1000,K=3"): N=1000 nodes, K=3
supports, T=12, hidden 32, M=2 branches (static + dynamic graph), 3 BDGCN layers per branch, batch 4 per
GPU.  A step = one forward + backward of the whole hot path over one batch (+ the gradient all-reduce for
N>1).  Scaling is WEAK: every rank processes its own batch of independent OD samples (SURVEY.md 8(e)
"batch shard"); the only exchange step is the all-reduce of the (tiny) parameter gradients over NCCL.

One JSON line is printed by rank 0 (see README / DESIGN.md for the field meanings).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "OD-cells/sec (B*T*N^2), MPGCN hot path fwd+bwd"
UNIT = "OD-cells/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--nodes", type=int, default=1000, help="N (OD zones)")
    ap.add_argument("--supports", type=int, default=3, help="K")
    ap.add_argument("--obs", type=int, default=12, help="T")
    ap.add_argument("--batch", type=int, default=4, help="samples per GPU")
    ap.add_argument("--hidden", type=int, default=32)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32", "auto"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile", action="store_true", help="under ncu: honour --warmup < 3, skip the e2e and CPU arms (numbers are not bench values)")
    return ap.parse_args()


def workload_config(a, world):
    return {
        "workload": f"MPGCN hot path N={a.nodes} K={a.supports} T={a.obs} hidden={a.hidden} M=2 L=3, batch {a.batch}/GPU",
        "N": a.nodes, "K": a.supports, "T": a.obs, "hidden": a.hidden, "M": 2, "gcn_layers": 3,
        "batch_per_gpu": a.batch, "global_batch": a.batch * world, "parallelism": f"batch-shard x{world}",
        "precision": a.precision, "l2": "inputs exceed L2 (activations >= 0.5 GB per layer); no explicit flush",
    }


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return dict(tflops=float(d["bf16_tflops_sustained"]), tflops_burst=float(d["bf16_tflops"]), hbm=float(d["hbm_gbs"]),
                        source="MEASURED_PEAKS.json (bf16 sustained; kernel timed inside a long step)")
        except Exception:
            pass
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# CPU arm (reference execution strategy, oracle/torch_port.py)
# ------------------------------------------------------------------------------------------------
def cpu_sample(a, seed=0):
    """One bounded sample of the workload on the host cores -> (estimated seconds for one B=1 model step, detail)."""
    import torch
    from oracle import torch_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    N, K, T, C = a.nodes, a.supports, a.obs, a.hidden
    # keep one sample to a few seconds: if the full-size layer is too slow on this host, time a smaller N and scale by N^3
    x = torch.randn(1536, 1536)
    t0 = time.perf_counter()
    for _ in range(3):
        x @ x
    rate = 3 * 2 * 1536 ** 3 / (time.perf_counter() - t0)          # flop/s of a large sgemm
    est = 3.0 * (4 * K * K * N ** 3 * C) / rate * 1.6
    Ns = N
    while est > 4.0 and Ns > 100:
        Ns = int(Ns * 0.85)
        est = 3.0 * (4 * K * K * Ns ** 3 * C) / rate * 1.6
    t_layer_s = torch_port.time_bdgcn_layer_fwd_bwd(Ns, K, B=1, C=C, H=C, seed=seed)
    t_layer = t_layer_s * (N / Ns) ** 3
    cells = N * N
    sample_cells = min(cells, 20_000)
    t_lstm = torch_port.time_lstm_fwd_bwd(sample_cells, T, C=C, seed=seed) * cells / sample_cells
    total = 2 * (3 * t_layer + t_lstm)
    detail = dict(cores=cores, bdgcn_layer_N_timed=Ns, t_bdgcn_layer_s=round(t_layer_s, 4), t_bdgcn_layer_scaled_s=round(t_layer, 4),
                  lstm_cells_timed=sample_cells, t_lstm_scaled_s=round(t_lstm, 4), sgemm_gflops=round(rate / 1e9, 1))
    return total, detail


def sample_text(d, a):
    scale = "" if d["bdgcn_layer_N_timed"] == a.nodes else f" (timed at N={d['bdgcn_layer_N_timed']}, scaled by N^3)"
    return (f"B=1: one BDGCN layer fwd+bwd at N={a.nodes},K={a.supports},C=H={a.hidden} in the reference's K^2-einsum order{scale}; "
            f"LSTM fwd+bwd on {d['lstm_cells_timed']} of {a.nodes ** 2} cells, T={a.obs}; model step = M*(L*t_layer + t_lstm), M=2, L=3")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    for _ in range(a.warmup):
        cpu_sample(a)
    ts, detail = [], None
    t_all = time.perf_counter()
    for i in range(a.steps):
        t, detail = cpu_sample(a, seed=i)
        ts.append(t)
    wall = time.perf_counter() - t_all
    t_step = sum(ts) / len(ts)
    cells = a.obs * a.nodes ** 2                    # B=1
    value = cells / t_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * t_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(a, 1), note="CPU arm: each step is a bounded sample, extrapolated to one B=1 model step",
                       wall_s_per_sample=round(wall / max(1, a.steps), 2)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": detail["cores"], "kind": "port", "sample": sample_text(detail, a), "detail": detail},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    from torch import nn

    import MPGCN as shim
    from mpgcn_b200 import _lib, dist as mdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the engine has no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    mdist.init_from_env("nccl", device=dev)
    lib = _lib.load()

    N, K, T, B, hid = a.nodes, a.supports, a.obs, a.batch, a.hidden
    torch.manual_seed(1234)                                   # identical weights on every rank
    model = shim.MPGCN(M=2, K=K, input_dim=1, lstm_hidden_dim=hid, lstm_num_layers=1, gcn_hidden_dim=hid, gcn_num_layers=3,
                       num_nodes=N, user_bias=True, activation=nn.ReLU).to(dev)
    model.lstm_precision = a.precision
    for mod in model.modules():
        if isinstance(mod, shim.BDGCN):
            mod.precision = a.precision
    crit = nn.MSELoss()
    g = torch.Generator().manual_seed(4321 + rank)            # a different batch per rank (weak scaling)
    x_host = (torch.rand(B, T, N, N, 1, generator=g) * 8).pin_memory()
    y_host = (torch.rand(B, 1, N, N, 1, generator=g) * 8).pin_memory()
    G_static = (torch.randn(K, N, N, generator=torch.Generator().manual_seed(7)) / N ** 0.5).to(dev)
    go_host = (torch.randn(B, K, N, N, generator=g) / N ** 0.5).pin_memory()
    gd_host = (torch.randn(B, K, N, N, generator=g) / N ** 0.5).pin_memory()
    params = [p for p in model.parameters()]

    def step(x, y, go, gd):
        for p in params:
            p.grad = None
        loss = crit(model(x_seq=x, G_list=[G_static, (go, gd)]), y)
        loss.backward()
        mdist.allreduce_mean_gradients(params)      # the only exchange step of the batch shard (no-op for one rank)
        return loss

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident-input arm -------------------------------------------------------------------
    x, y, go, gd = (t.to(dev) for t in (x_host, y_host, go_host, gd_host))
    n_warm = a.warmup if a.profile else max(3, a.warmup)
    for _ in range(n_warm):
        step(x, y, go, gd)
    sync_all()
    lib.mpgcn_profile_reset()
    lib.mpgcn_profile_enable(1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(a.steps):
        step(x, y, go, gd)
    e1.record()
    sync_all()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    lib.mpgcn_profile_enable(0)
    prof = _lib.profile_read()
    ms_per_step = ms / a.steps
    cells_per_step = world * B * T * N * N
    value = cells_per_step / (ms_per_step * 1e-3)

    # ---- end-to-end arm: host buffers in, loss out, every step -----------------------------------
    e2e = None
    if not (a.no_e2e or a.profile):
        del x, go, gd
        h2d = sum(t.numel() * t.element_size() for t in (x_host, y_host, go_host, gd_host))

        # Double-buffered input pipeline, as a prefetching data loader would do it: step k's inputs are copied from pinned
        # host memory on a side stream while step k-1 computes; every step still pays its own H2D copies and reads its loss
        # back (D2H) inside the timed region.
        copy_stream = torch.cuda.Stream(device=dev)
        hosts = (x_host, y_host, go_host, gd_host)

        def stage():
            with torch.cuda.stream(copy_stream):
                bufs = [t.to(dev, non_blocking=True) for t in hosts]
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return bufs, ev

        def e2e_loop(n):
            nxt = stage()
            losses = []
            for k in range(n):
                bufs, ev = nxt
                torch.cuda.current_stream().wait_event(ev)
                for b in bufs:
                    b.record_stream(torch.cuda.current_stream())
                if k + 1 < n:
                    nxt = stage()
                loss = step(*bufs)
                losses.append(float(loss.item()))            # device -> host read of the loss, every step
            return losses

        e2e_loop(2)
        sync_all()
        t0 = time.perf_counter()
        e0.record()
        e2e_loop(a.steps)
        e1.record()
        sync_all()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), wall_ms)) / a.steps
        e2e = {"value": cells_per_step / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 4 * world,
               "ms_per_step": ms_e2e}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    # ---- roofline of the dominant kernel (the N^3 tcgen05 contractions) -------------------------
    peaks = measured_peaks()
    big = ["FWD_A", "FWD_B", "BWD_V", "BWD_DX"]
    fl = sum(prof[t]["flops"] for t in big)
    tms = sum(prof[t]["ms"] for t in big)
    nl = sum(prof[t]["launches"] for t in big)
    achieved = fl / (tms * 1e-3) / 1e12 if tms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_dram_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "tensor", "kernel": "mpgcn::tc::contract_kernel (FWD_A/FWD_B/BWD_V/BWD_DX launches)", "achieved": achieved, "peak": peaks["tflops"],
        "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "traffic": traffic, "peak_source": peaks["source"],
        "launches": nl, "avg_launch_ms": tms / nl if nl else None, "algorithmic_flops_per_launch": fl / nl if nl else None,
        "share_of_step": tms / ms if ms > 0 else None,
        "per_stage": {t: {"launches": v["launches"], "ms": round(v["ms"], 3),
                          "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 and v["flops"] > 0 else None)} for t, v in prof.items()},
    }
    gpu_launches = sum(v["launches"] for v in prof.values())

    cpu_baseline = None
    if world == 1 and not (a.no_cpu_baseline or a.profile):
        t_cpu, detail = cpu_sample(a)
        cpu_baseline = {"value": (T * N * N) / t_cpu, "unit": UNIT, "cores": detail["cores"], "kind": "port", "sample": sample_text(detail, a),
                        "detail": detail}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": n_warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16" if a.precision != "fp32" else "f32", "data": "synthetic", "config": workload_config(a, world),
        "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
    }
    if world > 1:
        dist.destroy_process_group()
    return line


def main():
    a = parse()
    # Libraries (NCCL's version banner, torchrun notices) may write to fd 1; the contract is ONE JSON line on stdout.
    # Route everything else to stderr and hand the real stdout only to the final print.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    try:
        line = run_reference(a) if a.impl == "reference" else run_ours(a)
    finally:
        sys.stdout = real_stdout
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
