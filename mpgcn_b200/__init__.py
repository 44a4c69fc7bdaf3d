"""mpgcn_b200 -- B200-native engine for the MPGCN hot path (BDGCN 2-D graph convolution + per-cell LSTM).

Layout:  csrc/       CUDA kernels + C ABI (libmpgcn_b200.so; header in include/mpgcn_b200.h)
         _lib        ctypes binding                 ops         autograd operators over the C ABI
         MPGCN, GCN  drop-in `BDGCN` / `MPGCN` nn.Modules and `Adj_Processor` mirroring the reference's module surface
         dist        batch shard (one process per GPU, gradient all-reduce)
         shard       origin-row and K shards of one batch (layer parts; exchange over NVLink peer memory or NCCL)
         rollout     CUDA-graph replay of the trainer's autoregressive test loop      graph_step  captured training step
         dyn_graph   `construct_dyn_G` on the GPU
"""
from . import _lib, ops            # noqa: F401
from .MPGCN import BDGCN, MPGCN    # noqa: F401
from .GCN import Adj_Processor   # noqa: F401

__all__ = ["BDGCN", "MPGCN", "Adj_Processor", "ops"]
