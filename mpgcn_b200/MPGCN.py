"""Drop-in module surface of the reference's `MPGCN.py`, backed by the B200 engine.

`Model_Trainer.py:5` does `import GCN, MPGCN` by bare name and builds
`MPGCN.MPGCN(M=..., K=..., ..., user_bias=..., activation=nn.ReLU)` (Model_Trainer.py:47-56); put this
repository ahead of the reference on `sys.path` (the root-level `MPGCN.py` re-exports this module)
and the trainer runs unchanged on our kernels.  What is kept identical to the reference:

  * class names, constructor signatures (including the `user_bias` spelling) and attributes;
  * parameter names / shapes / init: `W [K*K*C_in, H]` Xavier-normal with row order (o, d, l),
    `b [H]` zeros (reference MPGCN.py:16-22); `state_dict` keys `branch_models.{m}.temporal.*`,
    `branch_models.{m}.spatial.{n}.{W,b}`, `branch_models.{m}.fc.0.{weight,bias}` -- checkpoints are
    interchangeable in both directions;
  * call conventions and error behaviour: `forward(X, G)` with G a `[K,N,N]` tensor or a 2-tuple of
    `[B,K,N,N]` tensors, AssertionError on K / shape mismatches, NotImplementedError for any other G
    (reference MPGCN.py:26-42, 95-96).

What differs: the arithmetic.  No einsum / cat / cuDNN -- `BDGCN.forward` is one call of
`ops.bdgcn` (factored 2K-contraction engine) and the per-cell LSTM is `ops.lstm_last`, which reads
`x_seq` in place, assumes the zero initial state the reference always passes (MPGCN.py:80-87,98) and
never materialises the `[B*N*N, T, C]` output sequence.
"""
from __future__ import annotations

import contextlib
import os

import torch
from torch import nn

from . import ops


class BDGCN(nn.Module):
    """2-D (origin x destination) multi-graph convolution.  Reference: MPGCN.py:6-50."""

    def __init__(self, K: int, input_dim: int, hidden_dim: int, use_bias=True, activation=None):
        super().__init__()
        self.K = K
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.use_bias = use_bias
        self.activation = activation() if activation is not None else None     # a class, as in the reference (MPGCN.py:13)
        self.precision = None          # None -> ops.default_precision(); or "auto" / "fp16" / "fp32"
        self.init_params()

    def init_params(self, b_init=0.0):
        self.W = nn.Parameter(torch.empty(self.input_dim * (self.K ** 2), self.hidden_dim), requires_grad=True)
        nn.init.xavier_normal_(self.W)
        if self.use_bias:
            self.b = nn.Parameter(torch.empty(self.hidden_dim), requires_grad=True)
            nn.init.constant_(self.b, val=b_init)

    def extra_repr(self) -> str:
        return f"K={self.K}, {self.input_dim} -> {self.hidden_dim}, bias={self.use_bias}"

    def forward(self, X: torch.Tensor, G):
        if isinstance(G, torch.Tensor):                     # static supports (K, N, N)
            assert self.K == G.shape[-3]
            assert G.dim() == 3, "static graph input must be (K, N, N)"
        elif isinstance(G, tuple):                          # dynamic supports ((B,K,N,N), (B,K,N,N))
            assert (len(G) == 2) & (self.K == G[0].shape[-3] == G[1].shape[-3])
            assert G[0].dim() == 4 and G[1].dim() == 4 and G[0].shape[0] == X.shape[0] == G[1].shape[0]
        else:
            raise NotImplementedError
        assert X.dim() == 4 and X.shape[1] == X.shape[2] == G[0].shape[-1] and X.shape[3] == self.input_dim
        fused_relu = isinstance(self.activation, nn.ReLU)
        out = ops.bdgcn(X, G, self.W, self.b if self.use_bias else None, relu=fused_relu, precision=self.precision)
        if self.activation is not None and not fused_relu:  # any other activation: unfused epilogue
            out = self.activation(out)
        return out


class MPGCN(nn.Module):
    """Multi-perspective model: per branch LSTM -> L x BDGCN -> Linear+ReLU, mean over branches.
    Reference: MPGCN.py:54-112."""

    def __init__(self, M: int, K: int, input_dim: int, lstm_hidden_dim: int, lstm_num_layers: int, gcn_hidden_dim: int,
                 gcn_num_layers: int, num_nodes: int, user_bias: bool, activation=None):
        super().__init__()
        self.M = M
        self.K = K
        self.num_nodes = num_nodes
        self.lstm_hidden_dim = lstm_hidden_dim
        self.lstm_num_layers = lstm_num_layers
        self.gcn_num_layers = gcn_num_layers
        self.lstm_precision = None      # None -> ops.default_precision(); or "auto" / "fp16" / "fp32"
        # True: evaluate the M branches on M CUDA streams (they are independent until the head, reference MPGCN.py:101-110), so that
        # the HBM-bound elementwise kernels of one branch run beside the tensor-bound contractions of the other; None -> env
        # MPGCN_B200_BRANCH_STREAMS (default off)
        self.branch_streams = None
        self._streams = None
        self.branch_models = nn.ModuleList()
        for _ in range(self.M):
            branch = nn.ModuleDict()
            # nn.LSTM is kept as the parameter container so state_dict keys / default init match
            branch['temporal'] = nn.LSTM(input_size=input_dim, hidden_size=lstm_hidden_dim, num_layers=lstm_num_layers, batch_first=True)
            branch['spatial'] = nn.ModuleList(
                BDGCN(K=K, input_dim=lstm_hidden_dim if n == 0 else gcn_hidden_dim, hidden_dim=gcn_hidden_dim,
                      use_bias=user_bias, activation=activation) for n in range(gcn_num_layers))
            branch['fc'] = nn.Sequential(nn.Linear(in_features=gcn_hidden_dim, out_features=input_dim, bias=True), nn.ReLU())
            self.branch_models.append(branch)

    def init_hidden_list(self, batch_size: int):
        """Kept for API compatibility (reference MPGCN.py:80-87); forward() does not need it."""
        weight = next(self.parameters()).data
        shape = (self.lstm_num_layers, batch_size * (self.num_nodes ** 2), self.lstm_hidden_dim)
        return [(weight.new_zeros(shape), weight.new_zeros(shape)) for _ in range(self.M)]

    def _temporal(self, lstm: nn.LSTM, x_seq: torch.Tensor) -> torch.Tensor:
        B, T, N, _, I = x_seq.shape
        if I == 1 and lstm.num_layers == 1 and lstm.hidden_size <= 64:
            return ops.lstm_last(x_seq, lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0, precision=self.lstm_precision)
        # configurations the trainer never builds (Model_Trainer.py:49-51 hard-codes input_dim=1, 1 layer)
        lstm_in = x_seq.permute(0, 2, 3, 1, 4).reshape(B * N * N, T, I)
        return lstm(lstm_in)[0][:, -1, :]

    def forward(self, x_seq: torch.Tensor, G_list: list):
        """x_seq (B, T, N, N, 1); G_list: per branch a static (K,N,N) tensor or a dynamic tuple.  -> (B, 1, N, N, 1)"""
        assert (len(x_seq.shape) == 5) & (self.num_nodes == x_seq.shape[2] == x_seq.shape[3])
        assert len(G_list) == self.M
        B, N = x_seq.shape[0], self.num_nodes
        use_streams = self.branch_streams if self.branch_streams is not None else os.environ.get("MPGCN_B200_BRANCH_STREAMS", "0") == "1"
        use_streams = bool(use_streams) and x_seq.is_cuda and self.M > 1
        capturing = use_streams and torch.cuda.is_current_stream_capturing()
        cur = torch.cuda.current_stream() if use_streams else None
        if use_streams and (self._streams is None or self._streams[0].device != x_seq.device):
            self._streams = [torch.cuda.Stream(device=x_seq.device) for _ in range(self.M)]
        feats = []
        for m in range(self.M):
            branch = self.branch_models[m]
            if use_streams:
                self._streams[m].wait_stream(cur)
            with (torch.cuda.stream(self._streams[m]) if use_streams else contextlib.nullcontext()):
                gcn_in = self._temporal(branch['temporal'], x_seq).reshape(B, N, N, self.lstm_hidden_dim)
                for layer in branch['spatial']:
                    gcn_in = layer(gcn_in, G_list[m])
            feats.append(gcn_in)
        if use_streams:
            for m in range(self.M):
                cur.wait_stream(self._streams[m])
                if not capturing:       # under CUDA-graph capture the join above is a graph dependency: later frees / re-uses are ordered by it
                    feats[m].record_stream(cur)
        fcs = [self.branch_models[m]['fc'][0] for m in range(self.M)]
        if all(fc.out_features == 1 for fc in fcs) and feats[0].shape[-1] % 4 == 0 and self.M <= 8:
            # Linear(C -> 1) + ReLU per branch and the mean over branches in one fused pass (reference MPGCN.py:107,110)
            w = torch.cat([fc.weight for fc in fcs], dim=0)            # [M, C]
            b = torch.cat([fc.bias for fc in fcs], dim=0)              # [M]
            ensemble_out = ops.fc_relu_mean(feats, w, b)               # [B, N, N, 1]
        else:
            branch_out = [self.branch_models[m]['fc'](feats[m]) for m in range(self.M)]
            ensemble_out = torch.mean(torch.stack(branch_out, dim=-1), dim=-1)
        return ensemble_out.unsqueeze(dim=1)
