"""Direct C-ABI calls with torch tensors (test helper): exposes `saved` and lets a test pass any
`out` (ReLU mask) to backward."""
import torch

from mpgcn_b200 import _lib

PREC = {"fp32": 0, "fp16": 1}


def _p(t):
    return None if t is None else t.data_ptr()


def forward(X, Go, Gd, W, b, relu, prec, want_saved=True):
    lib = _lib.load()
    B, N, _, C = X.shape
    K, H = Go.shape[-3], W.shape[1]
    dyn = int(Go.dim() == 4)
    pc = PREC[prec]
    out = torch.empty(B, N, N, H, device=X.device)
    saved = torch.empty(lib.mpgcn_bdgcn_saved_bytes(B, N, K, C, H, pc), dtype=torch.uint8, device=X.device) if want_saved else None
    ws = torch.empty(lib.mpgcn_bdgcn_fwd_workspace_bytes(B, N, K, C, H, dyn, pc), dtype=torch.uint8, device=X.device)
    _lib.check(lib.mpgcn_bdgcn_forward(_p(X), _p(Go), _p(Gd), dyn, _p(W), _p(b), int(relu), _p(out), _p(saved), _p(ws), ws.numel(),
                                       B, N, K, C, H, pc, torch.cuda.current_stream().cuda_stream), "forward")
    return out, saved


def backward(d_out, out, Go, Gd, W, relu, saved, prec, has_bias=True, C=None):
    lib = _lib.load()
    B, N, _, H = d_out.shape
    K = Go.shape[-3]
    C = W.shape[0] // (K * K) if C is None else C
    dyn = int(Go.dim() == 4)
    pc = PREC[prec]
    dX = torch.empty(B, N, N, C, device=d_out.device)
    dW = torch.empty_like(W)
    db = torch.empty(H, device=d_out.device) if has_bias else None
    ws = torch.empty(lib.mpgcn_bdgcn_bwd_workspace_bytes(B, N, K, C, H, dyn, pc), dtype=torch.uint8, device=d_out.device)
    _lib.check(lib.mpgcn_bdgcn_backward(_p(d_out), _p(out), _p(Go), _p(Gd), dyn, _p(W), int(relu), _p(saved), _p(dX), _p(dW), _p(db),
                                        _p(ws), ws.numel(), B, N, K, C, H, pc, torch.cuda.current_stream().cuda_stream), "backward")
    return dX, dW, db
