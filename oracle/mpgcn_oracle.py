"""CPU oracle for the MPGCN hot path (BDGCN 2D graph convolution, per-cell LSTM, MPGCN branch stack).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (`mpgcn_b200/`) may import this
module; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs use it, and only as the checker / the CPU arm that is timed.

This is a plain-numpy *restatement* of the arithmetic of the reference
(`/root/reference/MPGCN.py`, torch.einsum / nn.LSTM / nn.Linear), written from the
formulas, not a copy of the reference source.  Each function cites the reference lines
it follows.

PARITY PIN: the reference ships no tests or golden vectors, so the pin is
made here: `oracle/gen_golden.py` imports the *unmodified* reference classes from
`/root/reference` in the build container (fp32 CPU, fixed seeds), records their outputs
and autograd gradients into `tests/golden/*.npz`, and `tests/test_oracle_golden.py`
checks every function below against those fixtures (fp32 agreement <= 2e-5 relative, the
level of fp32 summation-order noise).  The fixtures travel to the GPU box; the
reference itself does not.

All arrays are numpy, any float dtype (float64 gives a tie-breaker reference).
Shapes use the reference's names: B batch, T obs length, N nodes, K supports,
C in-channels, H out-channels, M branches.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "bdgcn_forward", "bdgcn_backward", "lstm_last_forward", "lstm_last_backward",
    "fc_relu_forward", "fc_relu_backward", "mpgcn_forward", "mpgcn_forward_backward",
    "bdgcn_forward_factored", "bdgcn_backward_factored", "rel_errors", "adj_process",
]


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def _support_pair(G, b_dim: int):
    """Normalise the two graph-input forms of BDGCN.forward (MPGCN.py:26,34) to a pair
    (G_origin, G_dest) each of shape [B or 1, K, N, N]."""
    if isinstance(G, (tuple, list)):
        assert len(G) == 2
        go, gd = np.asarray(G[0]), np.asarray(G[1])
        assert go.ndim == 4 and gd.ndim == 4
        return go, gd
    G = np.asarray(G)
    assert G.ndim == 3
    return G[None], G[None]


def rel_errors(y, y_ref):
    """The parity metric used everywhere:
    rel_Linf = max|y - y_ref| / max|y_ref| ,  rel_L2 = ||y - y_ref|| / ||y_ref||."""
    y = np.asarray(y, dtype=np.float64)
    y_ref = np.asarray(y_ref, dtype=np.float64)
    den_inf = max(np.abs(y_ref).max(), 1e-30)
    den_l2 = max(np.linalg.norm(y_ref.ravel()), 1e-30)
    return float(np.abs(y - y_ref).max() / den_inf), float(np.linalg.norm((y - y_ref).ravel()) / den_l2)


# --------------------------------------------------------------------------------------
# BDGCN  (reference: MPGCN.py:24-50)
# --------------------------------------------------------------------------------------
def _pair_features(X, go, gd):
    """K*K bilinear graph features, concatenated on the channel axis in (o, d, l) order.

    Follows MPGCN.py:28-32 (static) / :36-40 (dynamic) and the cat at :44:
        F_od[b,m,e,l] = sum_{n,c} G_o[n,m] * X[b,n,c,l] * G_d[c,e]
    (the origin support is contracted over its FIRST index, i.e. G_o^T X G_d).
    """
    Bsz, N, _, C = X.shape
    K = go.shape[1]
    feats = np.empty((Bsz, N, N, K * K * C), dtype=X.dtype)
    for o in range(K):
        # origin-mode product, shared by all d (the reference recomputes it K times)
        Go = np.broadcast_to(go[:, o], (Bsz, N, N))                 # [B,n,m]
        m1 = np.einsum("bnm,bncl->bmcl", Go, X, optimize=True)      # MPGCN.py:30/38
        for d in range(K):
            Gd = np.broadcast_to(gd[:, d], (Bsz, N, N))             # [B,c,e]
            m2 = np.einsum("bmcl,bce->bmel", m1, Gd, optimize=True)  # MPGCN.py:31/39
            j = (o * K + d) * C
            feats[..., j:j + C] = m2
    return feats


def bdgcn_forward(X, G, W, b=None, act="relu", return_cache=False):
    """out = act( cat_{o,d}(G_o^T X G_d) @ W + b )          (MPGCN.py:24-50)

    X [B,N,N,C]; G static [K,N,N] or dynamic tuple ([B,K,N,N],[B,K,N,N]);
    W [K*K*C, H] with row index (o*K+d)*C+l (MPGCN.py:17,44-45); b [H] or None;
    act in {"relu", None} (the trainer passes nn.ReLU, Model_Trainer.py:56).
    """
    X = np.asarray(X)
    go, gd = _support_pair(G, X.shape[0])
    K = go.shape[1]
    assert W.shape[0] == K * K * X.shape[-1], "W rows must be K^2*C (MPGCN.py:17,27,35)"
    feats = _pair_features(X, go, gd)
    pre = feats @ W                                                  # MPGCN.py:45
    if b is not None:
        pre = pre + b                                                # MPGCN.py:47-48
    out = np.maximum(pre, 0) if act == "relu" else pre               # MPGCN.py:49
    if return_cache:
        return out, (feats, pre)
    return out


def bdgcn_backward(X, G, W, b, act, d_out, mask_from=None):
    """Gradients the reference obtains from autograd through MPGCN.py:24-50
    (loss.backward(), Model_Trainer.py:114).  The supports never require grad.
    Returns (dX [B,N,N,C], dW [K*K*C,H], db [H] or None).

    mask_from: optional forward output of ANOTHER implementation; if given, the ReLU mask is
    (mask_from > 0) instead of the oracle's own (pre > 0).  A reduced-precision forward flips
    the sign of pre-activations that lie within its rounding error of zero, so its (exact)
    gradient differs from the fp32 one on those few elements by O(1); passing its output here
    yields the gradient of the function it actually computed ("ReLU mask")."""
    X = np.asarray(X)
    go, gd = _support_pair(G, X.shape[0])
    Bsz, N, _, C = X.shape
    K = go.shape[1]
    out, (feats, pre) = bdgcn_forward(X, G, W, b, act, return_cache=True)
    if act == "relu":
        d_pre = d_out * ((pre if mask_from is None else np.asarray(mask_from)) > 0)
    else:
        d_pre = d_out
    db = d_pre.sum(axis=(0, 1, 2)) if b is not None else None
    dW = np.tensordot(feats, d_pre, axes=([0, 1, 2], [0, 1, 2]))     # [K*K*C, H]
    d_feats = d_pre @ W.T                                            # [B,N,N,K*K*C]
    dX = np.zeros_like(X)
    for o in range(K):
        Go = np.broadcast_to(go[:, o], (Bsz, N, N))
        for d in range(K):
            Gd = np.broadcast_to(gd[:, d], (Bsz, N, N))
            j = (o * K + d) * C
            dF = d_feats[..., j:j + C]                               # [B,m,e,l]
            # dX[b,n,c,l] += sum_{m,e} G_o[n,m] dF[b,m,e,l] G_d[c,e]
            t = np.einsum("bnm,bmel->bnel", Go, dF, optimize=True)
            dX += np.einsum("bnel,bce->bncl", t, Gd, optimize=True)
    return dX, dW, db


def _factored_forward(X, go, gd, W, b):
    """Shared by the factored forward / backward: returns (pre, Z) with Z[d][b] = X[b] x_2 G_d as [n, l, e] arrays.
    Every product is ONE BLAS call per sample (np.tensordot), so the oracle reaches N = 1000..2000 in seconds."""
    Bsz, N, _, C = X.shape
    K = go.shape[1]
    H = W.shape[1]
    W4 = W.reshape(K, K, C, H)
    pre = np.zeros((Bsz, N, N, H), dtype=X.dtype)
    Z = [[None] * Bsz for _ in range(K)]
    for bi in range(Bsz):
        for d in range(K):
            Gd = gd[bi if gd.shape[0] > 1 else 0, d]                       # [c, e]
            Z[d][bi] = np.tensordot(X[bi], Gd, axes=([1], [0]))            # [n, l, e]   (MPGCN.py:31/39, once per d)
        for o in range(K):
            U = sum(np.einsum("nle,lh->neh", Z[d][bi], W4[o, d], optimize=True) for d in range(K))   # [n, e, h]
            Go = go[bi if go.shape[0] > 1 else 0, o]                       # [n, m]
            pre[bi] += np.tensordot(Go, U, axes=([0], [0]))                # [m, e, h]   (MPGCN.py:30/38 + :45)
    if b is not None:
        pre = pre + b
    return pre, Z


def bdgcn_forward_factored(X, G, W, b=None, act="relu"):
    """Algebraically identical 2K-product evaluation order used by the CUDA engine
    (SURVEY.md section 7.1):  Z_d = X x_2 G_d ;  U_o = sum_d Z_d W[o,d] ;  pre = sum_o G_o^T x_1 U_o.
    Checked against the reference-order `bdgcn_forward` in tests/test_oracle_golden.py (1e-12 in float64)."""
    X = np.asarray(X)
    go, gd = _support_pair(G, X.shape[0])
    pre, _ = _factored_forward(X, go, gd, np.asarray(W), b)
    return np.maximum(pre, 0) if act == "relu" else pre


def bdgcn_backward_factored(X, G, W, b, act, d_out, mask_from=None):
    """Gradients of the layer in the factored order -- the same numbers as `bdgcn_backward`
    (checked to 1e-12 in float64 in tests/test_oracle_golden.py), at 2K instead of 2K^2 N^3-products, each one BLAS call:
        dPre = dOut * [pre > 0] ;  V_o = G_o x_1 dPre ;  dW[o,d] = sum Z_d^T V_o ;  Y_d = sum_o V_o W[o,d]^T ;
        dX = sum_d Y_d x_2 G_d^T.
    Returns (out, dX, dW, db)."""
    X = np.asarray(X)
    W = np.asarray(W)
    go, gd = _support_pair(G, X.shape[0])
    Bsz, N, _, C = X.shape
    K = go.shape[1]
    H = W.shape[1]
    W4 = W.reshape(K, K, C, H)
    pre, Z = _factored_forward(X, go, gd, W, b)
    out = np.maximum(pre, 0) if act == "relu" else pre
    d_pre = d_out * ((pre if mask_from is None else np.asarray(mask_from)) > 0) if act == "relu" else np.asarray(d_out)
    db = d_pre.sum(axis=(0, 1, 2)) if b is not None else None
    dW4 = np.zeros((K, K, C, H), dtype=X.dtype)
    dX = np.zeros_like(X)
    for bi in range(Bsz):
        V = [np.tensordot(go[bi if go.shape[0] > 1 else 0, o], d_pre[bi], axes=([1], [0])) for o in range(K)]   # [n, e, h]
        for d in range(K):
            Y = np.zeros((N, N, C), dtype=X.dtype)                                                          # [n, e, l]
            for o in range(K):
                dW4[o, d] += np.tensordot(Z[d][bi], V[o], axes=([0, 2], [0, 1]))                            # [l, h]
                Y += V[o] @ W4[o, d].T
            Gd = gd[bi if gd.shape[0] > 1 else 0, d]                                                        # [c, e]
            dX[bi] += np.transpose(np.tensordot(Y, Gd, axes=([1], [1])), (0, 2, 1))                         # [n, l, c] -> [n, c, l]
    return out, dX, dW4.reshape(K * K * C, H), db


# --------------------------------------------------------------------------------------
# per-cell LSTM, last hidden state only   (reference: MPGCN.py:69,80-87,100-104)
# --------------------------------------------------------------------------------------
def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def lstm_last_forward(x, w_ih, w_hh, b_ih, b_hh, return_cache=False):
    """Single-layer LSTM (torch gate order i,f,g,o), zero initial state (MPGCN.py:80-87),
    input feature size I (=1 in the model, MPGCN.py:69), returns h_T only -- the only
    slice the model uses (`lstm_out[:,-1,:]`, MPGCN.py:104).

    x [S,T,I] sequences (S = B*N*N cells); w_ih [4C,I]; w_hh [4C,C]; b_ih,b_hh [4C].
    """
    S, T, _ = x.shape
    C = w_hh.shape[1]
    h = np.zeros((S, C), dtype=x.dtype)
    c = np.zeros((S, C), dtype=x.dtype)
    cache = []
    bias = b_ih + b_hh
    for t in range(T):
        a = x[:, t, :] @ w_ih.T + h @ w_hh.T + bias
        i = _sigmoid(a[:, 0 * C:1 * C])
        f = _sigmoid(a[:, 1 * C:2 * C])
        g = np.tanh(a[:, 2 * C:3 * C])
        o = _sigmoid(a[:, 3 * C:4 * C])
        c_new = f * c + i * g
        tc = np.tanh(c_new)
        h_new = o * tc
        if return_cache:
            cache.append((h, c, i, f, g, o, tc))
        h, c = h_new, c_new
    if return_cache:
        return h, cache
    return h


def lstm_last_backward(x, w_ih, w_hh, b_ih, b_hh, d_hT):
    """BPTT for lstm_last_forward given dL/dh_T.  Returns (dx, dw_ih, dw_hh, db_ih, db_hh)."""
    S, T, _ = x.shape
    C = w_hh.shape[1]
    _, cache = lstm_last_forward(x, w_ih, w_hh, b_ih, b_hh, return_cache=True)
    dh = d_hT.copy()
    dc = np.zeros((S, C), dtype=x.dtype)
    dx = np.zeros_like(x)
    dw_ih = np.zeros_like(w_ih)
    dw_hh = np.zeros_like(w_hh)
    db = np.zeros_like(b_ih)
    for t in reversed(range(T)):
        h_prev, c_prev, i, f, g, o, tc = cache[t]
        do = dh * tc
        dc = dc + dh * o * (1 - tc * tc)
        di = dc * g
        df = dc * c_prev
        dg = dc * i
        da = np.concatenate([di * i * (1 - i), df * f * (1 - f), dg * (1 - g * g), do * o * (1 - o)], axis=1)
        dw_ih += da.T @ x[:, t, :]
        dw_hh += da.T @ h_prev
        db += da.sum(axis=0)
        dx[:, t, :] = da @ w_ih
        dh = da @ w_hh
        dc = dc * f
    return dx, dw_ih, dw_hh, db, db.copy()


# --------------------------------------------------------------------------------------
# FC head + branch fusion   (reference: MPGCN.py:74-76,107,110,112)
# --------------------------------------------------------------------------------------
def fc_relu_forward(x, w, b):
    """ReLU(Linear(C -> out)) per OD cell (MPGCN.py:74-76,107).  x [...,C]; w [O,C]; b [O]."""
    return np.maximum(x @ w.T + b, 0)


def fc_relu_backward(x, w, b, d_y, mask_from=None):
    pre = x @ w.T + b
    d_pre = d_y * ((pre if mask_from is None else np.asarray(mask_from)) > 0)
    dw = np.tensordot(d_pre, x, axes=(list(range(x.ndim - 1)), list(range(x.ndim - 1))))
    db = d_pre.reshape(-1, d_pre.shape[-1]).sum(axis=0)
    return d_pre @ w, dw, db


# --------------------------------------------------------------------------------------
# full model   (reference: MPGCN.py:54-112)
# --------------------------------------------------------------------------------------
def _p(params, key):
    return np.asarray(params[key])


def mpgcn_forward(params, x_seq, G_list, M, gcn_num_layers, act="relu", factored=False):
    """MPGCN.forward (MPGCN.py:89-112).  `params` maps the reference's state_dict keys
    (branch_models.{m}.temporal.weight_ih_l0 ... ) to arrays.
    x_seq [B,T,N,N,I]; G_list has M entries (static array or dynamic pair).
    Returns [B,1,N,N,I]."""
    x_seq = np.asarray(x_seq)
    Bsz, T, N, N2, I = x_seq.shape
    assert N == N2 and len(G_list) == M                              # MPGCN.py:95-96
    lstm_in = np.transpose(x_seq, (0, 2, 3, 1, 4)).reshape(Bsz * N * N, T, I)   # MPGCN.py:100
    outs = []
    for m in range(M):
        pre = f"branch_models.{m}."
        h = lstm_last_forward(lstm_in, _p(params, pre + "temporal.weight_ih_l0"), _p(params, pre + "temporal.weight_hh_l0"),
                              _p(params, pre + "temporal.bias_ih_l0"), _p(params, pre + "temporal.bias_hh_l0"))
        g = h.reshape(Bsz, N, N, -1)                                 # MPGCN.py:104
        for n in range(gcn_num_layers):                              # MPGCN.py:105-106
            bkey = pre + f"spatial.{n}.b"
            g = (bdgcn_forward_factored if factored else bdgcn_forward)(g, G_list[m], _p(params, pre + f"spatial.{n}.W"),
                                                                        _p(params, bkey) if bkey in params else None, act)
        outs.append(fc_relu_forward(g, _p(params, pre + "fc.0.weight"), _p(params, pre + "fc.0.bias")))  # :107
    y = np.mean(np.stack(outs, axis=-1), axis=-1)                    # MPGCN.py:110
    return y[:, None]                                                # MPGCN.py:112


def mpgcn_forward_backward(params, x_seq, G_list, M, gcn_num_layers, d_y, act="relu", masks=None, factored=False):
    """Forward + gradients of every parameter (what `loss.backward()` produces through
    MPGCN.py:89-112).  d_y [B,1,N,N,I] is dL/d(output).  Returns (y, grads dict keyed like
    the state_dict).

    masks: optional {m: {"layers": [out_0, .., out_{L-1}], "fc": fc_out}} -- forward outputs of another
    implementation whose signs replace the oracle's own ReLU masks in the backward pass (see
    bdgcn_backward's `mask_from`): the gradient of the function that implementation computed.
    factored: evaluate the BDGCN layers in the factored (BLAS-shaped) order -- same numbers, usable at N = 200+."""
    x_seq = np.asarray(x_seq)
    Bsz, T, N, _, I = x_seq.shape
    lstm_in = np.transpose(x_seq, (0, 2, 3, 1, 4)).reshape(Bsz * N * N, T, I)
    grads = {}
    outs = []
    tapes = []
    for m in range(M):
        pre = f"branch_models.{m}."
        lw = [_p(params, pre + "temporal." + k) for k in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
        h = lstm_last_forward(lstm_in, *lw)
        acts = [h.reshape(Bsz, N, N, -1)]
        for n in range(gcn_num_layers):
            bkey = pre + f"spatial.{n}.b"
            acts.append((bdgcn_forward_factored if factored else bdgcn_forward)(
                acts[-1], G_list[m], _p(params, pre + f"spatial.{n}.W"), _p(params, bkey) if bkey in params else None, act))
        outs.append(fc_relu_forward(acts[-1], _p(params, pre + "fc.0.weight"), _p(params, pre + "fc.0.bias")))
        tapes.append((lw, acts))
    y = np.mean(np.stack(outs, axis=-1), axis=-1)[:, None]
    for m in range(M):
        pre = f"branch_models.{m}."
        lw, acts = tapes[m]
        d = d_y[:, 0] / M
        mk = None if masks is None else masks[m]
        d, dw, db = fc_relu_backward(acts[-1], _p(params, pre + "fc.0.weight"), _p(params, pre + "fc.0.bias"), d,
                                     mask_from=None if mk is None else mk["fc"])
        grads[pre + "fc.0.weight"], grads[pre + "fc.0.bias"] = dw, db
        for n in reversed(range(gcn_num_layers)):
            bkey = pre + f"spatial.{n}.b"
            largs = (acts[n], G_list[m], _p(params, pre + f"spatial.{n}.W"), _p(params, bkey) if bkey in params else None, act, d)
            mf = None if mk is None else mk["layers"][n]
            if factored:
                _, d, dW, dbb = bdgcn_backward_factored(*largs, mask_from=mf)
            else:
                d, dW, dbb = bdgcn_backward(*largs, mask_from=mf)
            grads[pre + f"spatial.{n}.W"] = dW
            if dbb is not None:
                grads[bkey] = dbb
        _, dwi, dwh, dbi, dbh = lstm_last_backward(lstm_in, *lw, d.reshape(Bsz * N * N, -1))
        grads[pre + "temporal.weight_ih_l0"], grads[pre + "temporal.weight_hh_l0"] = dwi, dwh
        grads[pre + "temporal.bias_ih_l0"], grads[pre + "temporal.bias_hh_l0"] = dbi, dbh
    return y, grads


# --------------------------------------------------------------------------------------
# support-matrix builder   (reference: GCN.py:49-138 Adj_Processor)
# --------------------------------------------------------------------------------------
def _rw_normalize(A):
    """D^-1 A with 1/0 -> 0 (GCN.py:103-108)."""
    s = A.sum(axis=1)
    with np.errstate(divide="ignore"):
        dinv = np.where(s == 0, 0, 1.0 / s).astype(A.dtype)
    return dinv[:, None] * A


def _sym_normalize(A):
    """D^-1/2 A D^-1/2, no zero-degree guard (GCN.py:111-114)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.power(A.sum(axis=1), -0.5).astype(A.dtype)
    return (d[:, None] * A) * d[None, :]


def _cheb_series(x, K):
    """T_0 = I, T_1 = x, T_k = 2 x T_{k-1} - T_{k-2} for k = 0..K (GCN.py:128-138)."""
    N = x.shape[0]
    T = [np.eye(N, dtype=x.dtype)]
    if K >= 1:
        T.append(x)
    for k in range(2, K + 1):
        T.append(2 * (x @ T[k - 1]) - T[k - 2])
    return T


def adj_process(flow, kernel_type, K):
    """Adj_Processor(kernel_type, K).process(flow): flow [B,N,N] -> supports [B,Ks,N,N] (GCN.py:56-100).
    chebyshev uses lambda_max = 2 -- the branch the reference always takes on torch >= 2 (`torch.eig` was removed, the bare
    `except` at GCN.py:122 catches the error)."""
    flow = np.asarray(flow)
    if kernel_type == "localpool":
        K = 1
    out = []
    for A in flow:
        N = A.shape[0]
        eye = np.eye(N, dtype=A.dtype)
        if kernel_type == "localpool":
            ks = [eye + _sym_normalize(A)]                                   # GCN.py:69-72
        elif kernel_type == "chebyshev":
            L = eye - _sym_normalize(A)                                      # GCN.py:75
            ks = _cheb_series((2 / 2) * L - eye, K)                          # GCN.py:76-77,125
        elif kernel_type == "random_walk_diffusion":
            ks = _cheb_series(_rw_normalize(A).T, K)                         # GCN.py:79-82
        elif kernel_type == "dual_random_walk_diffusion":
            f = _cheb_series(_rw_normalize(A).T, K)                          # GCN.py:84-91
            bk = _cheb_series(_rw_normalize(A.T).T, K)
            ks = f + bk[1:]
        else:
            raise ValueError("Invalid kernel_type. Must be one of [chebyshev, localpool, random_walk_diffusion, dual_random_walk_diffusion].")
        out.append(np.stack(ks, axis=0))
    return np.stack(out, axis=0)


def construct_dyn_g(OD_data, split_ratio, perceived_period=7):
    """DataInput.construct_dyn_G (Data_Container_OD.py:39-59), vectorised in float64.
    OD_data [days, N, N, 1] -> (O_dyn_G, D_dyn_G) [N, N, P];  O[i,j,t] = cos_dist(A_t[i,:], A_t[j,:]),
    D[i,j,t] = cos_dist(A_t[:,i], A_t[j,:]) -- column i against ROW j, the reference's own eq.-(7) indexing at :56 --
    with scipy's cosine distance clip(1 - u.v / sqrt(u.u v.v), 0, 2) (NaN for a zero vector)."""
    OD_data = np.asarray(OD_data, dtype=np.float64)
    train_len = int(OD_data.shape[0] * split_ratio[0] / sum(split_ratio))               # :40
    periods = train_len // perceived_period                                               # :41
    hist = OD_data[:periods * perceived_period]                                           # :42
    Os, Ds = [], []
    with np.errstate(invalid="ignore", divide="ignore"):
        for t in range(perceived_period):
            A = hist[t::perceived_period].mean(axis=0)[..., 0]                            # :45
            rr = (A * A).sum(axis=1)          # |row i|^2
            cc = (A * A).sum(axis=0)          # |col i|^2
            O = 1.0 - (A @ A.T) / np.sqrt(rr[:, None] * rr[None, :])                      # :50-52
            D = 1.0 - (A.T @ A.T) / np.sqrt(cc[:, None] * rr[None, :])                    # :54-56  (col i . row j)
            Os.append(np.clip(O, 0.0, 2.0))
            Ds.append(np.clip(D, 0.0, 2.0))
    return np.stack(Os, axis=-1), np.stack(Ds, axis=-1)                                   # :59

