"""Pin the CPU oracle to the reference: the reference's own known answers and per-layer test
assertions, re-expressed on oracle/vv_oracle.cpp.

ref tests restated here:
  src/caffe/test/test_util_blas.cpp:22-130           GEMM / GEMV exact integer known answers
  src/caffe/test/test_max_margin_loss_layer.cpp:55-99 forward L1 recompute (1e-3), gradients L1 / L2
  src/caffe/test/test_normalization_layer.cpp:39-83   sum y^2 = 1 (1e-3), exhaustive gradient
  src/caffe/test/test_sum_layer.cpp:39-117            num_output 1 / 10 forward + gradient
  src/caffe/test/test_eltwise_layer.cpp:54-208        PROD / SUM / SUM-coeff forward (1e-4)
  src/caffe/test/test_inner_product_layer.cpp:43-111  forward sanity, gradient wrt W, b and bottom
  src/caffe/test/test_gradient_based_solver.cpp:58-305 least-squares update closed form (1 %)
plus an independent float64 autograd model of the whole TRAIN net (SURVEY 8a formulas).
"""
import numpy as np
import pytest
import torch


def numeric_grad(f, x, step):
    g = np.zeros_like(x, dtype=np.float64)
    flat = x.reshape(-1)
    for i in range(flat.size):
        old = flat[i]
        flat[i] = old + step; fp = f(x)
        flat[i] = old - step; fm = f(x)
        flat[i] = old
        g.reshape(-1)[i] = (fp - fm) / (2 * step)
    return g


def check_grad(analytic, numeric, threshold):
    # test_gradient_check_util.hpp:148-170: |a - n| <= threshold * max(|a|, |n|, 1)
    scale = np.maximum(np.maximum(np.abs(analytic), np.abs(numeric)), 1.0)
    assert np.all(np.abs(analytic - numeric) <= threshold * scale), float(np.max(np.abs(analytic - numeric) / scale))


# ---- BLAS known answers (test_util_blas.cpp:26-29, 97-99), built-in loops AND OpenBLAS ------------
@pytest.mark.parametrize("blas", ["builtin", "openblas"])
def test_gemm_known_answers(oracle, blas):
    if blas == "openblas":
        if oracle.use_openblas(0) == 1 and oracle.find_openblas() is None:
            pytest.skip("no OpenBLAS in this image")
    else:
        oracle.use_builtin_blas()
    data = np.arange(1, 13, dtype=np.float32)
    A = data[:6]; B = data
    A_t = np.array([1, 4, 2, 5, 3, 6], np.float32)
    B_t = np.array([1, 5, 9, 2, 6, 10, 3, 7, 11, 4, 8, 12], np.float32)
    result = np.array([38, 44, 50, 56, 83, 98, 113, 128], np.float32)
    z = np.zeros(8, np.float32)
    assert np.array_equal(oracle.gemm(0, 0, 2, 4, 3, 1.0, A, B, 0.0, z), result)
    assert np.array_equal(oracle.gemm(1, 0, 2, 4, 3, 1.0, A_t, B, 0.0, z), result)
    assert np.array_equal(oracle.gemm(1, 1, 2, 4, 3, 1.0, A_t, B_t, 0.0, z), result)
    assert np.array_equal(oracle.gemm(0, 1, 2, 4, 3, 1.0, A, B_t, 0.0, z), result)
    # gemv: A [2,3], x = {1,2,3} -> {14,32}; A^T y with y = {14,32}... the reference uses x2 = {1,1} -> {5,7,9}? no:
    # test_util_blas.cpp:97-99 result_2 = {9,12,15} for A^T * {1,2}... restated exactly:
    x = np.array([1, 2, 3], np.float32)
    assert np.array_equal(oracle.gemv(0, 2, 3, 1.0, A, x, 0.0, np.zeros(2)), np.array([14, 32], np.float32))
    y = np.array([1, 2], np.float32)
    assert np.array_equal(oracle.gemv(1, 2, 3, 1.0, A, y, 0.0, np.zeros(3)), np.array([9, 12, 15], np.float32))
    oracle.use_builtin_blas()


# ---- MaxMarginLoss ---------------------------------------------------------------------------------
def _loss_inputs():
    rng = np.random.RandomState(1701)
    return (rng.normal(0, 10, (10, 5)).astype(np.float32), rng.normal(0, 10, (10, 5)).astype(np.float32))


def test_max_margin_forward_l1(oracle):
    t, b = _loss_inputs()
    loss, viol, _ = oracle.max_margin_forward(t, b, margin=1.0, norm=1)
    d = t.astype(np.float64) - b.astype(np.float64)
    ref = np.where(d < 1, 1 - d, 0).sum() / t.size
    assert abs(ref - loss) < 1e-3
    assert viol == float((t - b < 0).sum())


@pytest.mark.parametrize("norm,step,thr", [(1, 1e-2, 2e-3), (2, 1e-2, 1e-2)])
def test_max_margin_gradient(oracle, norm, step, thr):
    t, b = _loss_inputs()
    if norm == 1:  # GradientChecker kink = 1, kink_range = 0.01: skip elements near the hinge kink
        d = 1.0 - (t - b)
        keep = np.abs(d) > 0.02
    else:
        keep = np.ones_like(t, bool)
    dt, dbg = oracle.max_margin_backward(t, b, margin=1.0, norm=norm, loss_weight=1.0)
    f = lambda x: oracle.max_margin_forward(x, b, margin=1.0, norm=norm)[0]
    num = numeric_grad(f, t.copy(), step)
    check_grad(dt[keep].astype(np.float64), num[keep], thr)
    assert np.array_equal(dt, -dbg)


def test_max_margin_weighted_asymmetry(oracle):
    """forward L2 uses sqrt(w)*h, backward uses w*h (max_margin_loss_layer.cpp:87 vs :154)."""
    t, b = _loss_inputs()
    w = np.full_like(t, 4.0)
    loss_w, _, hinge = oracle.max_margin_forward(t, b, margin=1.0, norm=2, weights=w)
    loss_1, _, hinge1 = oracle.max_margin_forward(t, b, margin=1.0, norm=2)
    assert np.allclose(hinge, 2.0 * hinge1)
    _, g_w = oracle.max_margin_backward(t, b, margin=1.0, norm=2, weights=w)
    _, g_1 = oracle.max_margin_backward(t, b, margin=1.0, norm=2)
    assert np.allclose(g_w, 4.0 * g_1)


# ---- Normalization -----------------------------------------------------------------------------------
def test_normalization_forward_unit_norm(oracle):
    rng = np.random.RandomState(1701)
    x = rng.normal(0, 1, (2, 3 * 4 * 5)).astype(np.float32)
    y = oracle.normalization_forward(x)
    assert np.all(np.abs((y.astype(np.float64) ** 2).sum(1) - 1.0) < 1e-3)
    # all-zero row: 0 / (0 + 1e-10) = 0 (dropout 0.9 can produce such rows)
    z = oracle.normalization_forward(np.zeros((1, 8), np.float32))
    assert np.array_equal(z, np.zeros((1, 8), np.float32))


def test_normalization_gradient(oracle):
    rng = np.random.RandomState(1701)
    x = rng.normal(0, 1, (2, 12)).astype(np.float32)
    for r in range(2):
        for c in range(12):   # exhaustive over top elements like CheckGradientExhaustive
            dy = np.zeros_like(x); dy[r, c] = 1.0
            dx = oracle.normalization_backward(x, dy)
            f = lambda v: float(oracle.normalization_forward(v)[r, c])
            num = numeric_grad(f, x.copy(), 1e-2)
            check_grad(dx.astype(np.float64), num, 1e-3 * 3)


# ---- Sum / Eltwise -------------------------------------------------------------------------------------
@pytest.mark.parametrize("nout", [1, 10])
def test_sum_forward_backward(oracle, nout):
    rng = np.random.RandomState(1701)
    x = rng.normal(0, 1, (2, 60)).astype(np.float32)
    y = oracle.sum_forward(x, nout)
    assert y.shape == (2, nout)
    assert np.all(np.abs(y - x.astype(np.float64).sum(1, keepdims=True)) < 1e-3)
    dy = rng.normal(0, 1, (2, nout)).astype(np.float32)
    dx = oracle.sum_backward(dy, 60)
    assert np.allclose(dx, np.repeat(dy.sum(1, keepdims=True), 60, 1), atol=1e-5)


def test_eltwise_sum_coeff(oracle):
    rng = np.random.RandomState(1701)
    bs = [rng.normal(0, 1, (2, 60)).astype(np.float32) for _ in range(3)]
    top = oracle.eltwise_sum_forward(bs, [1.0, -0.5, 2.0])
    assert np.all(np.abs(top - (bs[0] - 0.5 * bs[1] + 2 * bs[2])) < 1e-4)


# ---- InnerProduct ----------------------------------------------------------------------------------------
def test_inner_product_forward_and_gradients(oracle):
    rng = np.random.RandomState(1701)
    X = rng.uniform(0, 1, (2, 60)).astype(np.float32)
    W = rng.uniform(0, 1, (10, 60)).astype(np.float32)
    b = rng.uniform(1, 2, (10,)).astype(np.float32)
    Z = oracle.ip_forward(X, W, b)
    assert np.all(Z >= 1.0)                                   # test_inner_product_layer.cpp:64-86
    assert np.allclose(Z, X.astype(np.float64) @ W.T.astype(np.float64) + b, atol=1e-4)
    dZ = rng.normal(0, 1, Z.shape).astype(np.float32)
    dW, db, dX = oracle.ip_backward(dZ, X, W, want_dx=True)
    f = lambda: None
    numW = numeric_grad(lambda w: float((oracle.ip_forward(X, w, b).astype(np.float64) * dZ).sum()), W.copy(), 1e-2)
    check_grad(dW.astype(np.float64), numW, 1e-3)
    numX = numeric_grad(lambda x: float((oracle.ip_forward(x, W, b).astype(np.float64) * dZ).sum()), X.copy(), 1e-2)
    check_grad(dX.astype(np.float64), numX, 1e-3)
    assert np.allclose(db, dZ.sum(0), atol=1e-5)
    # fork-added regularization scales dW by (1 + reg/2) (inner_product_layer.cpp:80-90)
    dWr, _, _ = oracle.ip_backward(dZ, X, W, regularization=0.5)
    assert np.allclose(dWr, dW * 1.25, rtol=1e-6)


# ---- Solver: least-squares closed form (test_gradient_based_solver.cpp:223-305) -----------------------------
@pytest.mark.parametrize("lr,decay,momentum,iters", [(1.0, 0.0, 0.0, 1), (0.01, 0.5, 0.0, 1), (0.01, 0.0, 0.5, 4),
                                                     (0.01, 0.1, 0.9, 4)])
def test_solver_least_squares_update(oracle, lr, decay, momentum, iters):
    rng = np.random.RandomState(1701)
    n, D = 4, 27
    X = rng.normal(0, 1, (n, D)).astype(np.float32)
    y = rng.normal(0, 1, (n, 1)).astype(np.float32)
    w = rng.normal(0, 1, (1, D)).astype(np.float32); b = np.zeros(1, np.float32)
    hw = np.zeros_like(w); hb = np.zeros_like(b)
    w64, b64, hw64, hb64 = w.astype(np.float64), b.astype(np.float64), hw.astype(np.float64), hb.astype(np.float64)
    for it in range(iters):
        # EuclideanLoss gradient: (Xw + b - y)/n ; dW = that^T X
        r = (oracle.ip_forward(X, w, b) - y) / n
        dW, db, _ = oracle.ip_backward(r, X, w)
        w, _, hw = oracle.sgd_update(w, dW, hw, lr, momentum, decay)
        b, _, hb = oracle.sgd_update(b, db, hb, lr, momentum, decay)
        r64 = (X @ w64.T + b64 - y) / n
        gW = r64.T @ X + decay * w64; gb = r64.sum(0) + decay * b64
        hw64 = momentum * hw64 + lr * gW; hb64 = momentum * hb64 + lr * gb
        w64 = w64 - hw64; b64 = b64 - hb64
    assert np.allclose(w, w64, rtol=1e-2, atol=1e-5) and np.allclose(b, b64, rtol=1e-2, atol=1e-5)
    assert np.allclose(hw, hw64, rtol=1e-2, atol=1e-5) and np.allclose(hb, hb64, rtol=1e-2, atol=1e-5)


def test_dropout_constants(oracle):
    # scale_ = 1./(1.-0.9f) stored as float; uint_thres_ = (unsigned)(UINT_MAX * 0.9f) (dropout_layer.cpp:17-21)
    assert oracle.lib().orc_dropout_scale(np.float32(0.9)) == np.float32(1.0 / (1.0 - float(np.float32(0.9))))
    assert oracle.lib().orc_dropout_uint_thres(np.float32(0.9)) == 3865470464


# ---- whole TRAIN net vs an independent float64 autograd model -----------------------------------------------
def torch_model(data, W, b, mask, B, C, Nn, margin, norm, ratio, lw):
    R = C + Nn
    X = data.permute(1, 0, 2).reshape(R * B, -1)                 # slice dim1 + concat dim0
    Z = X @ W.t() + b
    scale = 1.0 / (1.0 - float(np.float32(ratio)))
    H = torch.relu(Z) * mask * float(np.float32(scale))
    Hs = H.reshape(R, B, -1)
    cbar = sum(Hs[i] * float(np.float32(1.0 / (C - 1))) for i in range(1, C))
    eps = 1e-10

    def l2n(x):
        # all-zero rows (dropout) give 0 forward and 0 backward in the reference
        # (normalization_layer.cpp:36-59, 101-110); keep autograd finite there.
        s = x.pow(2).sum(1, keepdim=True)
        s = torch.where(s > 0, s, torch.ones_like(s))
        return x / (s.sqrt() + eps)

    chat = l2n(cbar)
    that = l2n(Hs[0])
    st = (chat * that).sum(1, keepdim=True)
    sn = torch.stack([(chat * l2n(Hs[C + k])).sum(1) for k in range(Nn)], 1)
    h = torch.clamp(margin - (st - sn), min=0)
    loss = (h.pow(2) if norm == 2 else h.abs()).sum() / (B * Nn)
    return loss * lw, (st - sn < 0).sum()


@pytest.mark.parametrize("B,C,Nn,K,N,norm", [(6, 5, 10, 24, 16, 2), (4, 3, 4, 16, 8, 1), (3, 7, 5, 12, 20, 2)])
def test_net_matches_float64_autograd(oracle, B, C, Nn, K, N, norm):
    rng = np.random.RandomState(7)
    R = C + Nn
    data = np.maximum(rng.normal(0, 1, (B, R, K)), 0).astype(np.float32)
    W = rng.normal(0, 0.2, (N, K)).astype(np.float32)
    b = rng.normal(0, 0.1, (N,)).astype(np.float32)
    mask = (rng.uniform(0, 1, (R * B, N)) < 0.6).astype(np.uint32)
    ratio = 0.4
    out = oracle.net_forward_backward(data, W, b, mask, B, C, Nn, margin=2.0, norm=norm, dropout_ratio=ratio,
                                      want=("loss", "violations", "dW", "db", "dX", "H", "dZ"), want_dx=True)
    td = torch.tensor(data, dtype=torch.float64, requires_grad=True)
    tW = torch.tensor(W, dtype=torch.float64, requires_grad=True)
    tb = torch.tensor(b, dtype=torch.float64, requires_grad=True)
    loss, viol = torch_model(td, tW, tb, torch.tensor(mask.astype(np.float64)), B, C, Nn, 2.0, norm, ratio, 1.0)
    loss.backward()
    assert abs(out["loss"][0] - loss.item()) <= 1e-5 * max(1.0, abs(loss.item()))
    assert out["violations"][0] == float(viol.item())
    for name, g in (("dW", tW.grad.numpy()), ("db", tb.grad.numpy())):
        err = np.abs(out[name] - g).max() / max(np.abs(g).max(), 1e-12)
        assert err < 2e-5, (name, err)
    gX = td.grad.permute(1, 0, 2).reshape(R * B, K).numpy()
    assert np.abs(out["dX"] - gX).max() / max(np.abs(gX).max(), 1e-12) < 2e-5
