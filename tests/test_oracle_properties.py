"""Size-independent properties of the layer, checked on the CPU oracle with hypothesis-generated shapes:
linearity in X, identity supports, K = 1 equals the single pair, static == broadcast dynamic, factored == reference order (forward
and backward).  The GPU suite checks the same properties on the CUDA path (tests/test_gpu_parity.py::test_size_independent_properties)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import mpgcn_oracle as orc

shapes = st.tuples(st.integers(1, 3), st.integers(2, 7), st.integers(1, 3), st.integers(1, 5), st.integers(1, 5), st.integers(0, 10 ** 6))


def _case(B, N, K, C, H, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((B, N, N, C))
    G = rng.standard_normal((K, N, N)) / np.sqrt(N)
    W = rng.standard_normal((K * K * C, H))
    b = rng.standard_normal(H)
    return rng, X, G, W, b


@settings(max_examples=40, deadline=None)
@given(shapes)
def test_linearity_and_identity_supports(s):
    B, N, K, C, H, seed = s
    rng, X, G, W, b = _case(*s)
    X2 = rng.standard_normal(X.shape)
    f = lambda x: orc.bdgcn_forward(x, G, W, None, None)
    np.testing.assert_allclose(f(2.0 * X - 0.5 * X2), 2.0 * f(X) - 0.5 * f(X2), rtol=1e-9, atol=1e-9)
    eye = np.stack([np.eye(N)] * K)
    Wsum = W.reshape(K, K, C, H).sum(axis=(0, 1))
    np.testing.assert_allclose(orc.bdgcn_forward(X, eye, W, b, None), X @ Wsum + b, rtol=1e-9, atol=1e-9)


@settings(max_examples=40, deadline=None)
@given(shapes)
def test_static_equals_broadcast_dynamic_and_factored_order(s):
    B, N, K, C, H, seed = s
    rng, X, G, W, b = _case(*s)
    Gb = np.broadcast_to(G, (B, K, N, N)).copy()
    ref = orc.bdgcn_forward(X, G, W, b, "relu")
    np.testing.assert_allclose(orc.bdgcn_forward(X, (Gb, Gb), W, b, "relu"), ref, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(orc.bdgcn_forward_factored(X, G, W, b, "relu"), ref, rtol=1e-9, atol=1e-9)
    d = rng.standard_normal((B, N, N, H))
    dX, dW, db = orc.bdgcn_backward(X, G, W, b, "relu", d)
    out, dXf, dWf, dbf = orc.bdgcn_backward_factored(X, (Gb, Gb), W, b, "relu", d, mask_from=ref)
    np.testing.assert_allclose(out, ref, rtol=1e-9, atol=1e-9)
    for a, r in ((dXf, dX), (dWf, dW), (dbf, db)):
        np.testing.assert_allclose(a, r, rtol=1e-8, atol=1e-8)


@settings(max_examples=25, deadline=None)
@given(st.tuples(st.integers(1, 2), st.integers(2, 6), st.integers(1, 4), st.integers(1, 4), st.integers(0, 10 ** 6)))
def test_k1_is_the_single_pair_product(s):
    B, N, C, H, seed = s
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((B, N, N, C))
    G = rng.standard_normal((1, N, N))
    W = rng.standard_normal((C, H))
    want = np.einsum("nm,bncl,ce,lh->bmeh", G[0], X, G[0], W)
    np.testing.assert_allclose(orc.bdgcn_forward(X, G, W, None, None), want, rtol=1e-9, atol=1e-9)
