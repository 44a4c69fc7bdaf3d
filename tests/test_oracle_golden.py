"""Pin the CPU oracle (oracle/mpgcn_oracle.py) against golden vectors produced by the
unmodified reference classes (oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import mpgcn_oracle as orc

TOL = 2e-5      # fp32 summation-order noise between torch (bmm) and numpy (einsum)


def _check(a, ref, tol=TOL, what=""):
    linf, l2 = orc.rel_errors(a, ref)
    assert linf <= tol and l2 <= tol, f"{what}: rel_Linf={linf:.3e} rel_L2={l2:.3e} > {tol}"


def _graph(g):
    return (g["G_o"], g["G_d"]) if int(g["dynamic"]) else g["G"]


@pytest.mark.parametrize("name", golden_names("bdgcn_"))
def test_bdgcn_forward_backward_matches_reference(name):
    g = load_golden(name)
    act = None if str(g["act"]) == "none" else "relu"
    b = g.get("b")
    out = orc.bdgcn_forward(g["X"], _graph(g), g["W"], b, act)
    _check(out, g["out"], what="out")
    dX, dW, db = orc.bdgcn_backward(g["X"], _graph(g), g["W"], b, act, g["d_out"])
    _check(dX, g["dX"], what="dX")
    _check(dW, g["dW"], what="dW")
    if b is not None:
        _check(db, g["db"], what="db")


@pytest.mark.parametrize("name", golden_names("bdgcn_"))
def test_factored_order_equals_unfactored(name):
    g = load_golden(name)
    act = None if str(g["act"]) == "none" else "relu"
    X64 = g["X"].astype(np.float64)
    G = _graph(g)
    G64 = tuple(a.astype(np.float64) for a in G) if isinstance(G, tuple) else G.astype(np.float64)
    b = g.get("b")
    ref = orc.bdgcn_forward(X64, G64, g["W"].astype(np.float64), None if b is None else b.astype(np.float64), act)
    fac = orc.bdgcn_forward_factored(X64, G64, g["W"].astype(np.float64), None if b is None else b.astype(np.float64), act)
    _check(fac, ref, tol=1e-12, what="factored")
    _check(ref, g["out"], what="fp64 vs reference fp32")


@pytest.mark.parametrize("name", golden_names("lstm_"))
def test_lstm_matches_reference(name):
    g = load_golden(name)
    w = (g["w_ih"], g["w_hh"], g["b_ih"], g["b_hh"])
    _check(orc.lstm_last_forward(g["x"], *w), g["hT"], what="hT")
    dx, dwi, dwh, dbi, dbh = orc.lstm_last_backward(g["x"], *w, g["d_hT"])
    for a, k in ((dx, "dx"), (dwi, "dw_ih"), (dwh, "dw_hh"), (dbi, "db_ih"), (dbh, "db_hh")):
        _check(a, g[k], tol=5e-5, what=k)


@pytest.mark.parametrize("name", golden_names("mpgcn_"))
def test_model_matches_reference(name):
    g = load_golden(name)
    params = {k[6:]: v for k, v in g.items() if k.startswith("param:")}
    G_list = [g["G_static"], (g["G_o"], g["G_d"])]
    y, grads = orc.mpgcn_forward_backward(params, g["x_seq"], G_list, M=2, gcn_num_layers=3, d_y=g["d_y"])
    _check(y, g["y"], tol=5e-5, what="y")
    assert set(grads) == {k[5:] for k in g if k.startswith("grad:")}
    for k, v in grads.items():
        _check(v, g["grad:" + k], tol=2e-4, what=k)


def test_properties_linearity_and_identity():
    rng = np.random.default_rng(7)
    B, N, C, H, K = 2, 6, 3, 4, 2
    X1, X2 = rng.standard_normal((2, B, N, N, C))
    G = rng.standard_normal((K, N, N))
    W = rng.standard_normal((K * K * C, H))
    f = lambda X: orc.bdgcn_forward(X, G, W, None, None)
    np.testing.assert_allclose(f(2.0 * X1 - 3.0 * X2), 2.0 * f(X1) - 3.0 * f(X2), rtol=1e-10, atol=1e-10)
    eye = np.stack([np.eye(N)] * K)
    W4 = W.reshape(K, K, C, H)
    np.testing.assert_allclose(orc.bdgcn_forward(X1, eye, W, None, None), X1 @ W4.sum(axis=(0, 1)), rtol=1e-10, atol=1e-10)
    # static == dynamic with the static stack broadcast over the batch
    Gb = np.broadcast_to(G, (B, K, N, N)).copy()
    np.testing.assert_allclose(orc.bdgcn_forward(X1, (Gb, Gb), W, None, "relu"), orc.bdgcn_forward(X1, G, W, None, "relu"), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["bdgcn_s_k3_n12", "bdgcn_d_k3_n10", "bdgcn_s_k3_n12_linear_nobias"])
def test_torch_port_matches_reference_fixture(name):
    """The CPU timing arm (oracle/torch_port.py) computes the same function as the reference."""
    import torch
    from oracle import torch_port
    g = load_golden(name)
    X = torch.from_numpy(g["X"]).requires_grad_(True)
    W = torch.from_numpy(g["W"]).requires_grad_(True)
    b = torch.from_numpy(g["b"]).requires_grad_(True) if "b" in g else None
    G = (torch.from_numpy(g["G_o"]), torch.from_numpy(g["G_d"])) if int(g["dynamic"]) else torch.from_numpy(g["G"])
    out = torch_port.bdgcn_layer(X, G, W, b, relu=str(g["act"]) == "relu")
    out.backward(torch.from_numpy(g["d_out"]))
    _check(out.detach().numpy(), g["out"], what="out")
    _check(X.grad.numpy(), g["dX"], what="dX")
    _check(W.grad.numpy(), g["dW"], what="dW")


@pytest.mark.parametrize("name", golden_names("adj_"))
def test_adj_process_matches_reference(name):
    """Support-matrix builder (reference GCN.Adj_Processor.process) vs fixtures produced by the reference itself."""
    g = load_golden(name)
    sup = orc.adj_process(g["flow"], str(g["kernel_type"]), int(g["K"]))
    assert sup.shape == g["supports"].shape
    _check(sup, g["supports"], tol=2e-5, what="supports")


@pytest.mark.parametrize("name", golden_names("dyn_"))
def test_construct_dyn_g_matches_reference(name):
    """Vectorised restatement vs DataInput.construct_dyn_G of the unmodified reference (per-pair scipy distance.cosine),
    including the NaN pattern of an all-zero origin row and the column-vs-row indexing of the D graph."""
    g = load_golden(name)
    O, D = orc.construct_dyn_g(g["od"].astype(np.float64), list(g["split_ratio"]))
    for a, ref, what in ((O, g["O_dyn_G"], "O"), (D, g["D_dyn_G"], "D")):
        assert a.shape == ref.shape and a.dtype == np.float64
        assert np.array_equal(np.isnan(a), np.isnan(ref)), f"{name}/{what}: NaN pattern"
        assert np.nanmax(np.abs(a - ref)) <= 1e-12, f"{name}/{what}"
    if "zero" not in name:
        assert not np.allclose(D, np.transpose(D, (1, 0, 2))), "D graph must keep the reference's asymmetric (column i, row j) form"



# ---- at-size fixtures (N = 129, 200; C = H = 32): inputs regenerated from the seed, outputs pinned on a row subset ----
def load_big(name):
    """-> (fixture dict, X, d_out, G) with X / d_out regenerated from the stored seed and checked against the stored checksums."""
    from oracle.gen_golden import big_case_inputs
    g = load_golden(name)
    X, d_out = big_case_inputs(int(g["seed"]), int(g["N"]))
    assert abs(float(X.astype(np.float64).sum()) - float(g["x_checksum"])) < 1e-6, "numpy RNG stream changed: regenerate the fixture"
    assert abs(float(d_out.astype(np.float64).sum()) - float(g["d_out_checksum"])) < 1e-6
    rows = g["rows"]
    assert np.array_equal(X[0, rows[:4], 5, :4], g["x_probe"])
    G = (g["G_o"], g["G_d"]) if int(g["dynamic"]) else g["G"]
    return g, X, d_out, G


@pytest.mark.parametrize("name", golden_names("big_bdgcn_"))
def test_factored_oracle_matches_reference_at_size(name):
    """The BLAS-shaped factored oracle (the one the GPU tests use at N = 129 .. 2000) against the unmodified reference at
    N = 129 / 200, C = H = 32, static and dynamic supports -- forward and every gradient."""
    g, X, d_out, G = load_big(name)
    rows = g["rows"]
    out, dX, dW, db = orc.bdgcn_backward_factored(X, G, g["W"], g["b"], "relu", d_out)
    _check(out[:, rows], g["out_rows"], what="out rows")
    _check(dX[:, rows], g["dX_rows"], tol=5e-5, what="dX rows")
    _check(dW, g["dW"], tol=1e-4, what="dW")
    _check(db, g["db"], tol=1e-4, what="db")
    assert abs(np.abs(out).max() - float(g["out_absmax"])) <= 1e-4 * float(g["out_absmax"])
    assert abs(np.linalg.norm(out.astype(np.float64)) - float(g["out_norm"])) <= 1e-5 * float(g["out_norm"])
    assert abs(np.linalg.norm(dX.astype(np.float64)) - float(g["dX_norm"])) <= 1e-4 * float(g["dX_norm"])


@pytest.mark.parametrize("dyn", [False, True])
def test_factored_backward_equals_reference_order_backward(dyn):
    rng = np.random.default_rng(5 + dyn)
    B, N, K, C, H = 2, 9, 3, 4, 5
    X = rng.standard_normal((B, N, N, C))
    W = rng.standard_normal((K * K * C, H))
    b = rng.standard_normal(H)
    d = rng.standard_normal((B, N, N, H))
    G = (rng.standard_normal((B, K, N, N)), rng.standard_normal((B, K, N, N))) if dyn else rng.standard_normal((K, N, N))
    ref = orc.bdgcn_backward(X, G, W, b, "relu", d)
    out, dX, dW, db = orc.bdgcn_backward_factored(X, G, W, b, "relu", d)
    _check(out, orc.bdgcn_forward(X, G, W, b, "relu"), tol=1e-12, what="out")
    for a, r, what in zip((dX, dW, db), ref, ("dX", "dW", "db")):
        _check(a, r, tol=1e-12, what=what)


def test_cfg1_fixture_has_two_live_branches():
    """BASELINE config 1 (N=50, K=1, T=4, B=2): both branches of the reference-generated fixture carry gradient (round 1's
    fixture had a dead dynamic branch -- every branch_models.1.* gradient exactly 0)."""
    g = load_golden("mpgcn_cfg1_n50_k1")
    for k, v in g.items():
        if k.startswith("grad:"):
            assert np.abs(v).max() > 0, k
