"""Autograd-aware operators of the hot path, each a thin call into the C ABI.

    bdgcn(X, G, W, b, activation)  <->  reference BDGCN.forward           (MPGCN.py:24-50)
    lstm_last(x_seq, w_ih, w_hh, b_ih, b_hh)  <->  nn.LSTM(...)[:, -1, :]  (MPGCN.py:69,100-104)

PyTorch supplies device memory, the current stream and the autograd tape; all arithmetic is in
libmpgcn_b200.so.
"""
from __future__ import annotations

import collections
import os

import torch

from . import _lib

_PREC_NAMES = {"fp32": _lib.PREC_FP32, "fp16": _lib.PREC_FP16_TC, "auto": -1}

# bytes of training state (Z stash, LSTM c_t/h_t, head pre-activations) allocated by forward calls since the last .clear():
# stays at zero under torch.no_grad() (tests/test_gpu_at_size.py::test_no_grad_allocates_no_training_state)
STASH_BYTES = collections.Counter()


def default_precision() -> str:
    """'auto' (tensor cores whenever the shape allows), 'fp16' or 'fp32'.  Env: MPGCN_B200_PRECISION."""
    return os.environ.get("MPGCN_B200_PRECISION", "auto")


def resolve_precision(name, B, N, K, C, H) -> int:
    name = default_precision() if name is None else name
    if name not in _PREC_NAMES:
        raise ValueError(f"unknown precision {name!r}; expected one of {sorted(_PREC_NAMES)}")
    lib = _lib.load()
    if name == "auto":
        return _lib.PREC_FP16_TC if lib.mpgcn_bdgcn_precision_supported(B, N, K, C, H, _lib.PREC_FP16_TC) else _lib.PREC_FP32
    code = _PREC_NAMES[name]
    if not lib.mpgcn_bdgcn_precision_supported(B, N, K, C, H, code):
        raise RuntimeError(f"precision {name!r} does not support B={B} N={N} K={K} C={C} H={H} (tensor path needs C == H == 32, K <= 8)")
    return code


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"mpgcn_b200: {what} must be a CUDA tensor (got device {t.device}); the engine has no CPU path")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(dtype=torch.float32).contiguous()


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _scratch(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# Gradient-magnitude hand-over between consecutive backward calls of the fp16 path: the kernel that writes a gradient tensor also
# records max|grad| in a device scalar, and the backward that receives that tensor reuses it for its power-of-two scaling
# instead of re-reading the tensor.  The hand-over follows the autograd graph, not memory addresses: at forward time a consumer
# (BDGCN layer, FC head) notes which of OUR autograd nodes produced its input (looking through pure view nodes only); at
# backward time it leaves (data_ptr, version, numel, scalar) of the gradient it wrote ON that node, and the node's backward
# accepts the hint only if the gradient it receives is that very memory, unmodified (gradient accumulation from a second
# consumer arrives in a different buffer or with a bumped version).  Purely an optimisation: a missed hint costs one pass.
_VIEW_NODES = ("ViewBackward", "UnsafeViewBackward", "ReshapeAliasBackward", "AliasBackward")
_OUR_NODES = ("_BDGCNFnBackward", "_LSTMLastFnBackward")


def _producer_node(t):
    node = getattr(t, "grad_fn", None)
    for _ in range(8):
        if node is None:
            return None
        name = type(node).__name__
        if name in _OUR_NODES:
            return node
        if not name.startswith(_VIEW_NODES) or len(node.next_functions) != 1:
            return None
        node = node.next_functions[0][0]
    return None


def _put_hint(node, grad, absmax_scalar):
    if node is not None and grad is not None and absmax_scalar is not None:
        node._mpgcn_hint = (grad.data_ptr(), grad._version, grad.numel(), absmax_scalar)


def _take_hint(node, grad):
    e = getattr(node, "_mpgcn_hint", None)
    if e is None:
        return None
    node._mpgcn_hint = None
    return e[3] if (e[0] == grad.data_ptr() and e[1] == grad._version and e[2] == grad.numel()) else None


# Supports staged once: the same G_o / G_d serve every layer of a branch, forward and backward (6 calls per training step),
# so their fp16 conversion (+ diagonal remainders) is cached per support TENSOR OBJECT.  An entry is valid only while that
# very object is alive (weak reference), unmodified (version counter), at the same address and used on the same stream.
_SUPPORT_CACHE = {}
_SUPPORT_CACHE_MAX = 8


def _prepared_supports(lib, G, Gc, planes: int, N: int):
    import weakref
    key = id(G)
    stream = _stream()
    e = _SUPPORT_CACHE.get(key)
    if e is not None:
        ref, ver, ptr, st, blob = e
        if ref() is G and ver == G._version and ptr == Gc.data_ptr() and st == stream and blob.device == Gc.device:
            return blob
        del _SUPPORT_CACHE[key]
    for k in [k for k, v in _SUPPORT_CACHE.items() if v[0]() is None]:      # supports that no longer exist: free their staging now
        del _SUPPORT_CACHE[k]
    while len(_SUPPORT_CACHE) >= _SUPPORT_CACHE_MAX:
        _SUPPORT_CACHE.pop(next(iter(_SUPPORT_CACHE)))
    nbytes = lib.mpgcn_bdgcn_supports_prepared_bytes(planes, N)
    blob = torch.empty(nbytes, dtype=torch.uint8, device=Gc.device)
    with torch.cuda.device(Gc.device):
        _lib.check(lib.mpgcn_bdgcn_prepare_supports(_ptr(Gc), _ptr(blob), nbytes, planes, N, stream), "bdgcn_prepare_supports")
    try:
        _SUPPORT_CACHE[key] = (weakref.ref(G), G._version, Gc.data_ptr(), stream, blob)
    except TypeError:        # not weak-referenceable: do not cache
        pass
    return blob


class _BDGCNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, G_o, G_d, W, b, dynamic: bool, act: int, precision, grad_mode: bool):
        import ctypes
        lib = _lib.load()
        B, N, N2, C = X.shape
        K = G_o.shape[-3]
        H = W.shape[1]
        prec = resolve_precision(precision, B, N, K, C, H)
        tc = prec == _lib.PREC_FP16_TC
        Xc, Goc, Wc = _f32c(X), _f32c(G_o), _f32c(W)
        Gdc = Goc if G_d is G_o else _f32c(G_d)
        bc = None if b is None else _f32c(b)
        out = torch.empty((B, N, N, H), dtype=torch.float32, device=X.device)
        # ctx.needs_input_grad ignores the grad mode (it is True under torch.no_grad() too), and inside Function.forward
        # torch.is_grad_enabled() is always False: the wrapper reads the mode and hands it in.  Validation / test / the
        # autoregressive rollout therefore allocate no Z stash.
        need_grad = grad_mode and any(ctx.needs_input_grad)
        saved = _scratch(lib.mpgcn_bdgcn_saved_bytes(B, N, K, C, H, prec), X.device) if need_grad else None
        STASH_BYTES["bdgcn"] += saved.numel() if saved is not None else 0
        ws_bytes = lib.mpgcn_bdgcn_fwd_workspace_bytes(B, N, K, C, H, int(dynamic), prec)
        ws = _scratch(ws_bytes, X.device)
        ex = _lib.BdgcnExtras()
        preps = (None, None)
        if tc:
            # (The extras also allow handing an fp16 copy of the activation from layer to layer -- x_f16 / out_f16.  Measured on
            # B200 that is neutral: the extra 256 MB written by the FWD_B epilogue cost what the separate cast pass costs.)
            planes = (B if dynamic else 1) * K
            with torch.cuda.device(X.device):
                go_p = _prepared_supports(lib, G_o, Goc, planes, N)
                gd_p = go_p if G_d is G_o else _prepared_supports(lib, G_d, Gdc, planes, N)
            preps = (go_p, gd_p)
            ex.go_prepared, ex.gd_prepared = _ptr(go_p), _ptr(gd_p)
        with torch.cuda.device(X.device):
            _lib.check(lib.mpgcn_bdgcn_forward_x(_ptr(Xc), _ptr(Goc), _ptr(Gdc), int(dynamic), _ptr(Wc), _ptr(bc), act, _ptr(out),
                                                 _ptr(saved), _ptr(ws), ws.numel(), B, N, K, C, H, prec, ctypes.addressof(ex), _stream()),
                       "bdgcn_forward")
        ctx.shape = (B, N, K, C, H)
        ctx.meta = (bool(dynamic), act, prec, b is not None)
        ctx.x_producer = _producer_node(X) if need_grad else None
        ctx.preps = preps
        ctx.save_for_backward(out, Goc, Gdc, Wc, saved if saved is not None else torch.empty(0, device=X.device))
        return out

    @staticmethod
    def backward(ctx, d_out):
        import ctypes
        lib = _lib.load()
        out, Goc, Gdc, Wc, saved = ctx.saved_tensors
        B, N, K, C, H = ctx.shape
        dynamic, act, prec, has_bias = ctx.meta
        if saved.numel() == 0:
            raise RuntimeError("mpgcn_b200.bdgcn: backward called but forward ran without requires_grad inputs")
        tc = prec == _lib.PREC_FP16_TC
        hint = _take_hint(ctx, d_out) if (tc and d_out.dtype == torch.float32 and d_out.is_contiguous()) else None
        d_out = _f32c(d_out)
        dev = d_out.device
        need_dx = ctx.needs_input_grad[0]
        dX = torch.empty((B, N, N, C), dtype=torch.float32, device=dev) if need_dx else None
        dx_absmax = torch.empty(1, dtype=torch.float32, device=dev) if (need_dx and tc) else None
        dW = torch.empty_like(Wc)
        db = torch.empty(H, dtype=torch.float32, device=dev) if has_bias else None
        ws = _scratch(lib.mpgcn_bdgcn_bwd_workspace_bytes(B, N, K, C, H, int(dynamic), prec), dev)
        ex = _lib.BdgcnExtras()
        ex.go_prepared, ex.gd_prepared = _ptr(ctx.preps[0]), _ptr(ctx.preps[1])
        ex.d_out_absmax, ex.dX_absmax = _ptr(hint), _ptr(dx_absmax)
        with torch.cuda.device(dev):
            _lib.check(lib.mpgcn_bdgcn_backward_x(_ptr(d_out), _ptr(out), _ptr(Goc), _ptr(Gdc), int(dynamic),
                                                  _ptr(Wc), act, _ptr(saved), _ptr(dX), _ptr(dW), _ptr(db), _ptr(ws), ws.numel(), B, N, K,
                                                  C, H, prec, ctypes.addressof(ex), _stream()), "bdgcn_backward")
        _put_hint(ctx.x_producer, dX, dx_absmax)
        return dX, None, None, dW, db, None, None, None, None


def bdgcn(X: torch.Tensor, G, W: torch.Tensor, b, relu: bool, precision=None) -> torch.Tensor:
    """out = act(cat_{o,d}(G_o^T X G_d) W + b); G is a [K,N,N] tensor or a pair of [B,K,N,N] tensors."""
    _require_cuda(X, "X")
    if isinstance(G, torch.Tensor):
        G_o = G_d = G
        dynamic = False
    else:
        G_o, G_d = G
        dynamic = True
    for g in (G_o, G_d):
        _require_cuda(g, "G")
        if g.device != X.device:
            raise RuntimeError("mpgcn_b200: X and G must be on the same device")
    grad_mode = torch.is_grad_enabled()
    if grad_mode and (G_o.requires_grad or G_d.requires_grad):
        # the reference's einsums would deliver dL/dG through autograd; the trainer never asks for it (static G is a plain
        # tensor, dynamic G comes from the data loader) and the engine has no dG kernels: refuse instead of returning None
        raise NotImplementedError("mpgcn_b200.bdgcn: gradients with respect to the supports G are not implemented "
                                  "(pass G.detach(), or learnable supports through the reference's einsum path)")
    return _BDGCNFn.apply(X, G_o, G_d, W, b, dynamic, 1 if relu else 0, precision, grad_mode)


def resolve_lstm_precision(name, T, C) -> int:
    name = default_precision() if name is None else name
    if name not in _PREC_NAMES:
        raise ValueError(f"unknown precision {name!r}; expected one of {sorted(_PREC_NAMES)}")
    lib = _lib.load()
    if name == "auto":
        return _lib.PREC_FP16_TC if lib.mpgcn_lstm_precision_supported(T, C, _lib.PREC_FP16_TC) else _lib.PREC_FP32
    code = _PREC_NAMES[name]
    if not lib.mpgcn_lstm_precision_supported(T, C, code):
        raise RuntimeError(f"LSTM precision {name!r} does not support T={T}, hidden={C} (tensor path needs hidden == 32)")
    return code


class _LSTMLastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_seq, w_ih, w_hh, b_ih, b_hh, precision, grad_mode: bool):
        lib = _lib.load()
        B, T = x_seq.shape[0], x_seq.shape[1]
        NN = x_seq[0, 0].numel()
        C = w_hh.shape[1]
        prec = resolve_lstm_precision(precision, T, C)
        xc = _f32c(x_seq)
        ws = [_f32c(t) for t in (w_ih, w_hh, b_ih, b_hh)]
        hT = torch.empty((B * NN, C), dtype=torch.float32, device=x_seq.device)
        # training: the forward keeps c_t / h_t of every step (fp16) so that the backward is one reverse walk
        nsave = lib.mpgcn_lstm_saved_bytes(B, T, NN, C, prec) if (grad_mode and any(ctx.needs_input_grad[:5])) else 0
        saved = torch.empty(nsave, dtype=torch.uint8, device=x_seq.device) if nsave else None
        # (precision 0 keeps no state -- its backward recomputes -- so count the request, not the buffer)
        STASH_BYTES["lstm"] += nsave if nsave else (1 if (grad_mode and any(ctx.needs_input_grad[:5])) else 0)
        with torch.cuda.device(x_seq.device):
            _lib.check(lib.mpgcn_lstm_last_forward_train(_ptr(xc), *[_ptr(t) for t in ws], _ptr(hT), _ptr(saved), nsave, B, T, NN, C, prec,
                                                         _stream()), "lstm_last_forward")
        ctx.dims = (B, T, NN, C, prec)
        ctx.lstm_saved = saved
        ctx.save_for_backward(xc, *ws)
        return hT

    @staticmethod
    def backward(ctx, d_hT):
        lib = _lib.load()
        xc, w_ih, w_hh, b_ih, b_hh = ctx.saved_tensors
        B, T, NN, C, prec = ctx.dims
        hint = _take_hint(ctx, d_hT) if (prec == _lib.PREC_FP16_TC and d_hT.dtype == torch.float32 and d_hT.is_contiguous()) else None
        d_hT = _f32c(d_hT)
        dev = xc.device
        g_wih, g_whh = torch.empty_like(w_ih), torch.empty_like(w_hh)
        g_bih, g_bhh = torch.empty_like(b_ih), torch.empty_like(b_hh)
        d_x = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        saved = ctx.lstm_saved
        ws = _scratch(1024 if saved is not None else lib.mpgcn_lstm_bwd_workspace_bytes(B, T, NN, C, prec), dev)
        with torch.cuda.device(dev):
            _lib.check(lib.mpgcn_lstm_last_backward_saved(_ptr(xc), _ptr(w_ih), _ptr(w_hh), _ptr(b_ih), _ptr(b_hh), _ptr(d_hT), _ptr(g_wih),
                                                          _ptr(g_whh), _ptr(g_bih), _ptr(g_bhh), _ptr(d_x), _ptr(saved),
                                                          saved.numel() if saved is not None else 0, _ptr(ws), ws.numel(), B, T, NN, C,
                                                          prec, _ptr(hint), _stream()), "lstm_last_backward")
        ctx.lstm_saved = None
        return d_x, g_wih, g_whh, g_bih, g_bhh, None, None


def lstm_last(x_seq: torch.Tensor, w_ih, w_hh, b_ih, b_hh, precision=None) -> torch.Tensor:
    """h_T of a 1-layer, input-size-1 LSTM run over every OD cell of x_seq [B,T,N,N,1] -> [B*N*N, C]."""
    _require_cuda(x_seq, "x_seq")
    return _LSTMLastFn.apply(x_seq, w_ih, w_hh, b_ih, b_hh, precision, torch.is_grad_enabled())


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad_mode, w, b, *gs):
        import ctypes
        lib = _lib.load()
        M = len(gs)
        C = gs[0].shape[-1]
        cells = gs[0].numel() // C
        gc = [_f32c(g) for g in gs]
        wc, bc = _f32c(w), _f32c(b)
        y = torch.empty(gs[0].shape[:-1] + (1,), dtype=torch.float32, device=gs[0].device)
        need = grad_mode and any(ctx.needs_input_grad)
        pre = torch.empty((M, cells), dtype=torch.float32, device=y.device) if need else None
        STASH_BYTES["head"] += pre.numel() * 4 if pre is not None else 0
        ptrs = (ctypes.c_void_p * M)(*[g.data_ptr() for g in gc])
        with torch.cuda.device(y.device):
            _lib.check(lib.mpgcn_head_forward(ptrs, _ptr(wc), _ptr(bc), _ptr(y), _ptr(pre), cells, C, M, _stream()), "head_forward")
        ctx.dims = (M, C, cells)
        ctx.g_producers = [_producer_node(g) for g in gs] if need else None
        ctx.save_for_backward(wc, pre if pre is not None else torch.empty(0, device=y.device), *gc)
        return y

    @staticmethod
    def backward(ctx, dy):
        import ctypes
        lib = _lib.load()
        wc, pre, *gc = ctx.saved_tensors
        M, C, cells = ctx.dims
        dy = _f32c(dy)
        if pre.numel() == 0:
            raise RuntimeError("mpgcn_b200.fc_relu_mean: backward called but forward ran without requires_grad inputs")
        dgs = [torch.empty_like(g) if ctx.needs_input_grad[3 + m] else None for m, g in enumerate(gc)]
        dw = torch.empty_like(wc)
        db = torch.empty(M, dtype=torch.float32, device=dy.device)
        ptrs = (ctypes.c_void_p * M)(*[g.data_ptr() for g in gc])
        dptrs = (ctypes.c_void_p * M)(*[(d.data_ptr() if d is not None else None) for d in dgs])
        amax = torch.empty(M, dtype=torch.float32, device=dy.device)
        with torch.cuda.device(dy.device):
            _lib.check(lib.mpgcn_head_backward(ptrs, _ptr(wc), _ptr(pre), _ptr(dy), dptrs, _ptr(dw), _ptr(db), _ptr(amax), cells, C, M,
                                               _stream()), "head_backward")
        for m, d in enumerate(dgs):
            _put_hint(ctx.g_producers[m], d, amax[m:m + 1])
        return (None, dw, db) + tuple(dgs)


def fc_relu_mean(gs, w, b) -> torch.Tensor:
    """(1/M) * sum_m relu(g_m @ w[m] + b[m]) for M branch activations g_m [..., C]; w [M, C], b [M] -> [..., 1]."""
    for g in gs:
        _require_cuda(g, "branch activation")
    return _HeadFn.apply(torch.is_grad_enabled(), w, b, *gs)
