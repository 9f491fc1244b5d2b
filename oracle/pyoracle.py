"""ctypes wrapper of the CPU oracle (oracle/vv_oracle.cpp).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never by the product package."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_build", "libvv_oracle.so")
_P, _i, _f, _d = C.c_void_p, C.c_int, C.c_float, C.c_double


class NetCfg(C.Structure):
    _fields_ = [("B", _i), ("C", _i), ("Nn", _i), ("K", _i), ("N", _i), ("margin", _f), ("norm", _i),
                ("dropout_ratio", _f), ("negative_slope", _f), ("loss_weight", _f), ("regularization", _d),
                ("has_bias", _i)]


class NetOut(C.Structure):
    _fields_ = [(n, _P) for n in ("X", "Z", "H", "target_score", "neg_score", "loss", "violations", "dH", "dZ",
                                  "dW", "db", "dX", "phase_seconds")]


_lib = None


def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    src = os.path.join(_HERE, "vv_oracle.cpp")
    if (not os.path.exists(SO)) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-std=c++14", "-Wno-deprecated-declarations", "-fPIC", "-shared",
                               "-o", SO, src, "-ldl"])


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
        _lib.orc_dropout_scale.restype = _f
        _lib.orc_dropout_scale.argtypes = [_f]
        _lib.orc_dropout_uint_thres.restype = C.c_uint
        _lib.orc_dropout_uint_thres.argtypes = [_f]
        _lib.orc_learning_rate.restype = _f
        _lib.orc_learning_rate.argtypes = [_i, _f, _f, _f, _i, _i]
        _lib.orc_sampler_create.restype = _P
        _lib.orc_sampler_create.argtypes = [_i, _i, _P, _P, _P, _P, _i, _i, _i, _i, _i, _i, _i]
        _lib.orc_sampler_create_mode.restype = _P
        _lib.orc_sampler_create_mode.argtypes = [_i, _i, _P, _P, _P, _P, _i, _i, _i, _i, _i, _i, _i, _i]
        _lib.orc_sampler_create_opts.restype = _P
        _lib.orc_sampler_create_opts.argtypes = [_i, _i, _P, _P, _P, _P, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _P, _P, _P, _P, _i]
        _lib.orc_sampler_next.argtypes = [_P, _P, _P, _P]
        _lib.orc_sampler_destroy.argtypes = [_P]
        _lib.orc_sampler_cursor.argtypes = [_P]
        _lib.orc_set_blas.argtypes = [C.c_char_p, _i]
        _lib.orc_net_forward_backward.argtypes = [C.POINTER(NetCfg), _P, _P, _P, _P, C.POINTER(NetOut)]
    return _lib


def find_openblas():
    import scipy
    cands = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so"))
    return cands[0] if cands else None


def use_openblas(threads=0):
    """Route the oracle's gemm/gemv through the OpenBLAS in the SciPy wheel (the reference's
    BLAS := open, Makefile.config:34).  Returns the thread count in use (1 = built-in loops)."""
    path = find_openblas()
    if path is None or lib().orc_set_blas(path.encode(), threads) != 0:
        lib().orc_set_blas(None, 0)
        return 1
    return lib().orc_blas_threads()


def use_builtin_blas():
    lib().orc_set_blas(None, 0)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


# ---- layer-level wrappers (numpy in, numpy out) -------------------------------------------------
def gemm(transA, transB, M, N, K, alpha, A, B, beta, Cm):
    A, B = f32(A), f32(B)
    Cm = f32(Cm).copy()
    lib().orc_gemm(int(transA), int(transB), M, N, K, _f(alpha), _p(A), _p(B), _f(beta), _p(Cm))
    return Cm


def gemv(transA, M, N, alpha, A, x, beta, y):
    A, x = f32(A), f32(x)
    y = f32(y).copy()
    lib().orc_gemv(int(transA), M, N, _f(alpha), _p(A), _p(x), _f(beta), _p(y))
    return y


def ip_forward(X, W, b):
    X, W = f32(X), f32(W)
    M, K = X.shape
    N = W.shape[0]
    Z = np.empty((M, N), np.float32)
    bb = f32(b) if b is not None else None
    lib().orc_ip_forward(M, N, K, _p(X), _p(W), _p(bb), _p(Z))
    return Z


def ip_backward(dZ, X, W, regularization=0.0, want_dx=False):
    dZ, X, W = f32(dZ), f32(X), f32(W)
    M, K = X.shape
    N = W.shape[0]
    dW = np.empty((N, K), np.float32)
    db = np.empty((N,), np.float32)
    dX = np.empty((M, K), np.float32) if want_dx else None
    lib().orc_ip_backward(M, N, K, _p(dZ), _p(X), _p(W), _d(regularization), _p(dW), _p(db), _p(dX))
    return dW, db, dX


def relu_forward(x, slope=0.0):
    x = f32(x); y = np.empty_like(x)
    lib().orc_relu_forward(C.c_size_t(x.size), _p(x), _f(slope), _p(y))
    return y


def relu_backward(x, dy, slope=0.0):
    x, dy = f32(x), f32(dy); dx = np.empty_like(x)
    lib().orc_relu_backward(C.c_size_t(x.size), _p(x), _p(dy), _f(slope), _p(dx))
    return dx


def dropout_forward(x, mask, ratio):
    x = f32(x); m = np.ascontiguousarray(mask, dtype=np.uint32); y = np.empty_like(x)
    lib().orc_dropout_forward(C.c_size_t(x.size), _p(x), _p(m), _f(ratio), _p(y))
    return y


def dropout_backward(dy, mask, ratio):
    dy = f32(dy); m = np.ascontiguousarray(mask, dtype=np.uint32); dx = np.empty_like(dy)
    lib().orc_dropout_backward(C.c_size_t(dy.size), _p(dy), _p(m), _f(ratio), _p(dx))
    return dx


def normalization_forward(x):
    x = f32(x); y = np.empty_like(x)
    lib().orc_normalization_forward(x.shape[0], x.size // x.shape[0], _p(x), _p(y))
    return y


def normalization_backward(x, dy):
    x, dy = f32(x), f32(dy); dx = np.empty_like(x)
    lib().orc_normalization_backward(x.shape[0], x.size // x.shape[0], _p(x), _p(dy), _p(dx))
    return dx


def sum_forward(x, num_output):
    x = f32(x); num = x.shape[0]; y = np.empty((num, num_output), np.float32)
    lib().orc_sum_forward(num, x.size // num, num_output, _p(x), _p(y))
    return y


def sum_backward(dy, dim):
    dy = f32(dy); num, nout = dy.shape; dx = np.empty((num, dim), np.float32)
    lib().orc_sum_backward(num, dim, nout, _p(dy), _p(dx))
    return dx


def eltwise_sum_forward(bottoms, coeffs):
    bs = [f32(b) for b in bottoms]
    ptrs = (C.c_void_p * len(bs))(*[b.ctypes.data for b in bs])
    co = f32(coeffs); top = np.empty_like(bs[0])
    lib().orc_eltwise_sum_forward(C.c_size_t(top.size), len(bs), ptrs, _p(co), _p(top))
    return top


def max_margin_forward(s_true, s_bogus, margin=1.0, norm=1, weights=None):
    a, b = f32(s_true), f32(s_bogus)
    w = f32(weights) if weights is not None else None
    hinge = np.empty(a.size, np.float32); loss = np.zeros(1, np.float32); viol = np.zeros(1, np.float32)
    lib().orc_max_margin_forward(a.size, _p(a), _p(b), _p(w), _f(margin), norm, _p(hinge), _p(loss), _p(viol))
    return float(loss[0]), float(viol[0]), hinge.reshape(a.shape)


def max_margin_backward(s_true, s_bogus, margin=1.0, norm=1, loss_weight=1.0, weights=None):
    a, b = f32(s_true), f32(s_bogus)
    w = f32(weights) if weights is not None else None
    dt = np.empty_like(a); dbg = np.empty_like(a)
    lib().orc_max_margin_backward(a.size, _p(a), _p(b), _p(w), _f(margin), norm, _f(loss_weight), _p(dt), _p(dbg))
    return dt, dbg


POLICY = {"fixed": 0, "step": 1, "exp": 2, "inv": 3}


def learning_rate(policy, base_lr, gamma, power, stepsize, it):
    return float(lib().orc_learning_rate(POLICY[policy], base_lr, gamma, power, stepsize, it))


def sgd_update(data, diff, hist, local_rate, momentum, local_decay, reg_type=2):
    """In place on copies; returns (data, diff, hist) after one ComputeUpdateValue + Update."""
    data, diff, hist = f32(data).copy(), f32(diff).copy(), f32(hist).copy()
    lib().orc_sgd_update(C.c_size_t(data.size), _p(data), _p(diff), _p(hist), _f(local_rate), _f(momentum),
                         _f(local_decay), reg_type)
    return data, diff, hist


def net_forward_backward(data, W, b, mask, B, Cc, Nn, margin=2.0, norm=2, dropout_ratio=0.9, loss_weight=1.0,
                         regularization=0.0, want=("loss", "violations", "dW", "db"), want_dx=False):
    """The whole TRAIN net on the data blob [B,R,K].  Returns a dict of numpy arrays."""
    data, W = f32(data), f32(W)
    K = data.shape[-1]; N = W.shape[0]; R = Cc + Nn; M = R * B
    cfg = NetCfg(B, Cc, Nn, K, N, margin, norm, dropout_ratio, 0.0, loss_weight, regularization, 1 if b is not None else 0)
    shapes = dict(X=(M, K), Z=(M, N), H=(M, N), target_score=(B, Nn), neg_score=(B, Nn), loss=(1,), violations=(1,),
                  dH=(M, N), dZ=(M, N), dW=(N, K), db=(N,), dX=(M, K))
    arrs = {}
    out = NetOut()
    for name in want:
        arrs[name] = np.zeros(shapes[name], np.float32)
        setattr(out, name, arrs[name].ctypes.data)
    if want_dx and "dX" not in arrs:
        arrs["dX"] = np.zeros(shapes["dX"], np.float32); out.dX = arrs["dX"].ctypes.data
    ph = np.zeros(8, np.float64); out.phase_seconds = ph.ctypes.data
    bb = f32(b) if b is not None else None
    mm = np.ascontiguousarray(mask, dtype=np.uint32) if mask is not None else None
    rc = lib().orc_net_forward_backward(C.byref(cfg), _p(data), _p(W), _p(bb), _p(mm), C.byref(out))
    assert rc == 0
    arrs["phase_seconds"] = ph
    return arrs


class Sampler:
    """The reference's sampler on an in-memory dataset; uses the process-global libc rand()."""

    def __init__(self, video_id, shot_off, shot_ids, feat, K, batch_size, context_size=5, num_negative_samples=10,
                 max_buffer_size=5000, negative_swap_percentage=50, max_same_video_negs=6, max_tries_for_negs=100,
                 seed=1, context_type=1, start_skip=0, negative_dataset=None):
        """start_skip: the value of the data layer's rand_skip draw; negative_dataset: (video_id, shot_off, shot_ids, feat,
        row_base) -- indices of its shots come out as row_base + position in that set."""
        self.video_id = np.ascontiguousarray(video_id, np.int32)
        self.shot_off = np.ascontiguousarray(shot_off, np.int32)
        self.shot_ids = np.ascontiguousarray(shot_ids, np.int32)
        self.feat = f32(feat) if feat is not None else None
        self.B, self.R, self.K = batch_size, context_size + num_negative_samples, K
        lib().orc_srand(seed)
        nv, neg, nbase = 0, [None] * 4, 0
        if negative_dataset is not None:
            neg = [np.ascontiguousarray(a, np.int32) for a in negative_dataset[:3]] + [f32(negative_dataset[3]) if negative_dataset[3] is not None else None]
            nv, nbase = len(neg[0]), int(negative_dataset[4])
        self._neg = neg
        self._h = lib().orc_sampler_create_opts(len(self.video_id), K, _p(self.video_id), _p(self.shot_off), _p(self.shot_ids),
                                                _p(self.feat), batch_size, context_size, num_negative_samples, max_buffer_size,
                                                negative_swap_percentage, max_same_video_negs, max_tries_for_negs, context_type,
                                                int(start_skip), nv, _p(neg[0]), _p(neg[1]), _p(neg[2]), _p(neg[3]), nbase)
        if not self._h:
            raise RuntimeError("oracle sampler: could not fill the negative buffer")

    def next(self):
        idx = np.empty((self.B, self.R), np.int32); quirk = np.empty((self.B, self.R), np.int32)
        data = np.empty((self.B, self.R, self.K), np.float32) if self.feat is not None else None
        rc = lib().orc_sampler_next(self._h, _p(idx), _p(quirk), _p(data))
        assert rc == 0
        return idx, quirk, data

    @property
    def cursor(self):
        return lib().orc_sampler_cursor(self._h)

    def close(self):
        if self._h:
            lib().orc_sampler_destroy(self._h); self._h = None


def test_embed(frames, W, bias, coeff=None):
    """TEST-phase embedding of frames [B, F, K]: mean of the F frames -> fc7 -> ReLU -> L2 normalisation.  Returns (xbar, E)."""
    frames = f32(frames); W = f32(W); bias = f32(bias)
    B, F, K = frames.shape; N = W.shape[0]
    coeff = f32(coeff if coeff is not None else np.full(F, 1.0 / F))
    xbar = np.empty((B, K), np.float32); E = np.empty((B, N), np.float32)
    lib().orc_test_embed(B, F, K, N, _p(frames), _p(coeff), _p(W), _p(bias), _p(xbar), _p(E))
    return xbar, E


def retrieval_stats(E, video_ids, labels, exclude_same_video_shots=False, dist=None):
    """RetrievalStatsLayer (shot level).  Returns dict(map, hit1, hit5, per_query [B,3], dist [B,B])."""
    E = f32(E); B, N = E.shape
    vid = np.ascontiguousarray(video_ids, np.int32); lab = np.ascontiguousarray(labels, np.int32)
    D = f32(dist).copy() if dist is not None else np.empty((B, B), np.float32)
    out = np.zeros(3, np.float64); pq = np.empty((B, 3), np.float64)
    lib().orc_retrieval_stats(B, N, _p(E), _p(vid), _p(lab), int(exclude_same_video_shots), _p(D), int(dist is not None), _p(out), _p(pq))
    return dict(map=out[0], hit1=out[1], hit5=out[2], per_query=pq, dist=D)


def id_lookup_forward(table, ids):
    table = f32(table); ids = f32(ids); top = np.empty((ids.size, table.shape[1]), np.float32)
    lib().orc_id_lookup_forward(ids.size, table.shape[1], _p(table), _p(ids), _p(top))
    return top


def id_lookup_backward(top_diff, ids, rows):
    top_diff = f32(top_diff); ids = f32(ids); d = np.empty((rows, top_diff.shape[1]), np.float32)
    lib().orc_id_lookup_backward(ids.size, top_diff.shape[1], rows, _p(top_diff), _p(ids), _p(d))
    return d
