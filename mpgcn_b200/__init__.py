"""mpgcn_b200 -- B200-native engine for the MPGCN hot path (BDGCN 2-D graph convolution + per-cell LSTM).

Layout:  csrc/  CUDA kernels + C ABI (libmpgcn_b200.so; header in include/mpgcn_b200.h)
         _lib   ctypes binding          ops    autograd operators over the C ABI
         MPGCN  drop-in `BDGCN` / `MPGCN` nn.Modules mirroring the reference's module surface
"""
from . import _lib, ops            # noqa: F401
from .MPGCN import BDGCN, MPGCN    # noqa: F401
from .GCN import Adj_Processor   # noqa: F401

__all__ = ["BDGCN", "MPGCN", "Adj_Processor", "ops"]
