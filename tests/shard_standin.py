"""CPU stand-in for `mpgcn_b200.shard._ENGINE` (TEST INFRASTRUCTURE): the part semantics of include/mpgcn_b200.h
(`mpgcn_bdgcn_forward_part` / `_backward_part`, `mpgcn_bias_act`, `mpgcn_relu_backward`) restated with torch einsums in
float64, so that the exchange logic of mpgcn_b200/shard.py runs under gloo on a machine without a GPU, and the CUDA part
kernels can be checked against an independent evaluation on the GPU box."""
import torch


def _g(G, dynamic):
    return G if dynamic else G[None]            # [B or 1, K, N, N]


class TorchEngine:
    dtype = torch.float64

    def forward_part(self, X, Go, Gd, dynamic, W, N, row0, Ko, Kd, prec, keep):
        B, rows, _, C = X.shape
        H = W.shape[1]
        X, W = X.to(self.dtype), W.to(self.dtype).view(Ko, Kd, C, H)
        go, gd = _g(Go, dynamic).to(self.dtype), _g(Gd, dynamic).to(self.dtype)
        Z = torch.einsum("bncl,zdce->bdnel" if not dynamic else "bncl,bdce->bdnel", X, gd)            # [B,Kd,rows,N,C]
        U = torch.einsum("bdnel,odlh->boneh", Z, W)                                                   # [B,Ko,rows,N,H]
        gos = go[:, :, row0:row0 + rows, :]                                                           # [*,Ko,rows,N(m)]
        pre = torch.einsum("zonm,boneh->bmeh" if not dynamic else "bonm,boneh->bmeh", gos, U)
        return pre.to(torch.float32), (Z if keep else None)

    def backward_part(self, d_pre, Go, Gd, dynamic, W, saved, N, row0, rows, Ko, Kd, C, prec, need_dx):
        H = W.shape[1]
        W4 = W.to(self.dtype).view(Ko, Kd, C, H)
        go, gd = _g(Go, dynamic).to(self.dtype), _g(Gd, dynamic).to(self.dtype)
        gos = go[:, :, row0:row0 + rows, :]
        V = torch.einsum("zonm,bmeh->boneh" if not dynamic else "bonm,bmeh->boneh", gos, d_pre.to(self.dtype))
        dW = torch.einsum("bdnel,boneh->odlh", saved, V).reshape(Ko * Kd * C, H)
        dX = None
        if need_dx:
            Y = torch.einsum("boneh,odlh->bdnel", V, W4)
            dX = torch.einsum("bdnel,zdce->bncl" if not dynamic else "bdnel,bdce->bncl", Y, gd).to(torch.float32)
        return dX, dW.to(torch.float32)

    def bias_act(self, pre, bias, act):
        if bias is not None:
            pre.add_(bias)
        if act:
            pre.clamp_(min=0)
        return pre

    def relu_backward(self, d_out, out, act, want_db):
        d_pre = d_out * (out > 0) if act else d_out.clone()
        return d_pre, (d_pre.reshape(-1, d_pre.shape[-1]).sum(0) if want_db else None)

    def lstm_last(self, x_seq, lstm, precision):
        B, T, rows, N, I = x_seq.shape
        out, _ = lstm(x_seq.permute(0, 2, 3, 1, 4).reshape(B * rows * N, T, I))
        return out[:, -1, :]

    def head(self, feats, w, b):
        outs = [torch.relu(f @ w[m:m + 1].T + b[m]) for m, f in enumerate(feats)]
        return torch.mean(torch.stack(outs, dim=-1), dim=-1)

    def resolve_precision(self, name, B, N, K, C, H):
        return 0
