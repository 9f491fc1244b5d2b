"""ctypes wrapper of oracle/_ref/libvv_ref.so: the REFERENCE's own layer classes compiled unmodified against shim
headers (oracle/ref_shim/).  TEST INFRASTRUCTURE ONLY.  available() is False where the library was not built
(e.g. a checkout without /root/reference): callers then fall back to the oracle port and say so."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libvv_ref.so")
_lib = None


def available():
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO)
        _lib.ref_describe.restype = C.c_char_p
    return _lib


DROPIN_SO = os.path.join(_HERE, "_ref", "libvv_dropin.so")
_dropin = None


def dropin_available():
    return os.path.exists(DROPIN_SO)


def dropin_lib():
    """oracle/_ref/libvv_dropin.so: the same reference sources compiled in Caffe GPU mode, their device side (the layers'
    Forward_gpu / Backward_gpu, caffe_gpu_*) supplied by oracle/ref_shim/dropin_gpu.cpp = calls into libvv_b200.so's C-ABI.
    Same C entry points as libvv_ref.so; ref_solver_create runs the reference's Net / SGDSolver in GPU mode."""
    global _dropin
    if _dropin is None:
        _dropin = C.CDLL(DROPIN_SO)
    return _dropin


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def net_forward_backward(data, W, b, B, Cc, Nn, margin=2.0, norm=2, dropout_ratio=0.9, seed=1701):
    """Whole TRAIN net through the reference's layer classes.  Returns a dict incl. the dropout mask it drew."""
    data, W, b = f32(data), f32(W), f32(b)
    K = data.shape[-1]; N = W.shape[0]; R = Cc + Nn; M = R * B
    out = dict(loss=np.zeros(1, np.float32), violations=np.zeros(1, np.float32), dW=np.zeros((N, K), np.float32),
               db=np.zeros(N, np.float32), mask=np.ones((M, N), np.uint32), H=np.zeros((M, N), np.float32),
               dZ=np.zeros((M, N), np.float32), target_score=np.zeros((B, Nn), np.float32),
               neg_score=np.zeros((B, Nn), np.float32), seconds=np.zeros(3, np.float64))
    rc = lib().ref_net_forward_backward(B, Cc, Nn, K, N, C.c_float(margin), norm, C.c_float(dropout_ratio), _p(data), _p(W), _p(b),
                                        C.c_uint(seed), _p(out["loss"]), _p(out["violations"]), _p(out["dW"]), _p(out["db"]),
                                        _p(out["mask"]), _p(out["H"]), _p(out["dZ"]), _p(out["target_score"]), _p(out["neg_score"]),
                                        _p(out["seconds"]))
    if rc != 0:
        raise RuntimeError("reference net failed")
    return out


def normalization(x, dy):
    x, dy = f32(x), f32(dy); y = np.empty_like(x); dx = np.empty_like(x)
    assert lib().ref_normalization(x.shape[0], x.size // x.shape[0], _p(x), _p(dy), _p(y), _p(dx)) == 0
    return y, dx


def max_margin(s_true, s_bogus, margin=1.0, norm=1, loss_weight=1.0):
    a, b = f32(s_true), f32(s_bogus)
    loss = np.zeros(1, np.float32); viol = np.zeros(1, np.float32); dt = np.empty_like(a); dbg = np.empty_like(a)
    assert lib().ref_max_margin(a.shape[0], a.size // a.shape[0], _p(a), _p(b), C.c_float(margin), norm, C.c_float(loss_weight),
                                _p(loss), _p(viol), _p(dt), _p(dbg)) == 0
    return float(loss[0]), float(viol[0]), dt, dbg


def max_margin_weighted(s_true, s_bogus, third, direct, id_to_weight_file="", margin=1.0, norm=1, loss_weight=1.0):
    """The reference layer with its third bottom: per-element weights (direct) or video ids + an "id,weight" file."""
    a, b, c = f32(s_true), f32(s_bogus), f32(third)
    loss = np.zeros(1, np.float32); viol = np.zeros(1, np.float32); dt = np.empty_like(a); dbg = np.empty_like(a)
    assert lib().ref_max_margin_w(a.shape[0], a.size // a.shape[0], _p(a), _p(b), _p(c), int(bool(direct)),
                                  id_to_weight_file.encode(), C.c_float(margin), norm, C.c_float(loss_weight),
                                  _p(loss), _p(viol), _p(dt), _p(dbg)) == 0
    return float(loss[0]), float(viol[0]), dt, dbg


def inner_product(X, W, b, dZ, regularization=0.0):
    X, W, b, dZ = f32(X), f32(W), f32(b), f32(dZ)
    M, K = X.shape; N = W.shape[0]
    Z = np.empty((M, N), np.float32); dW = np.empty((N, K), np.float32); db = np.empty(N, np.float32); dX = np.empty((M, K), np.float32)
    assert lib().ref_inner_product(M, N, K, _p(X), _p(W), _p(b), _p(dZ), C.c_float(regularization), _p(Z), _p(dW), _p(db), _p(dX)) == 0
    return Z, dW, db, dX


def retrieval_stats(E, video_ids, id_to_class_file, exclude_same_video_shots=True):
    """The reference's RetrievalStatsLayer (shot level).  Returns (mAP, hit@1, hit@5)."""
    E = f32(E); ids = f32(video_ids)
    out = np.zeros(3, np.float32)
    rc = lib().ref_retrieval_stats(E.shape[0], E.shape[1], _p(E), _p(ids), str(id_to_class_file).encode(), int(exclude_same_video_shots), _p(out))
    assert rc == 0
    return out


def retrieval_stats_ex(E, video_ids, id_to_class_file, exclude_same_video_shots=True, video_level=False, max_num_videos=0,
                       stats_output_file=""):
    """The reference's RetrievalStatsLayer with video_level_retrieval / stats_output_file.  Returns (mAP, hit@1, hit@5)."""
    E = f32(E); ids = f32(video_ids)
    out = np.zeros(3, np.float32)
    rc = lib().ref_retrieval_stats_ex(E.shape[0], E.shape[1], _p(E), _p(ids), str(id_to_class_file).encode(),
                                      int(exclude_same_video_shots), int(video_level), int(max_num_videos),
                                      str(stats_output_file).encode(), _p(out))
    assert rc == 0
    return out


def id_to_weight(table, ids, top_diff=None):
    """The reference's IdToWeightMappingLayer: (top, table_diff or None)."""
    table = f32(table); ids = f32(ids)
    M = ids.size; rows, N = table.shape
    top = np.empty((M, N), np.float32)
    d = np.empty((rows, N), np.float32) if top_diff is not None else None
    td = f32(top_diff) if top_diff is not None else None
    rc = lib().ref_id_to_weight(M, N, rows, _p(table), _p(ids), _p(td), _p(top), _p(d))
    assert rc == 0
    return top, d


class Sampler:
    """The REFERENCE's VideoSampledShotsDataLayer itself (compiled unmodified) over an in-memory fake LMDB.  Draws from the
    process-global libc rand() like the oracle's sampler does: run one of them to completion before the other."""

    def __init__(self, video_id, shot_off, shot_ids, feat, K, batch_size, context_size=5, num_negative_samples=10,
                 max_buffer_size=5000, negative_swap_percentage=50, max_same_video_negs=6, seed=1, context_type=1,
                 rand_skip=0, caffe_seed=0, negative_dataset=None):
        """rand_skip / caffe_seed: the layer draws its skip from caffe_rng_rand() right after Caffe::set_random_seed(caffe_seed);
        negative_dataset: (video_id, shot_off, shot_ids, feat) served as a second fake LMDB."""
        L = lib()
        L.ref_sampler_create.restype = C.c_void_p
        L.ref_sampler_create_opts.restype = C.c_void_p
        L.ref_sampler_next.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_sampler_rows.argtypes = [C.c_void_p]
        L.ref_sampler_destroy.argtypes = [C.c_void_p]
        self.vid = np.ascontiguousarray(video_id, np.int32); self.off = np.ascontiguousarray(shot_off, np.int32)
        self.sid = np.ascontiguousarray(shot_ids, np.int32); self.feat = f32(feat)
        L.ref_srand(C.c_uint(seed))
        nv, neg = 0, [None] * 4
        if negative_dataset is not None:
            neg = [np.ascontiguousarray(a, np.int32) for a in negative_dataset[:3]] + [f32(negative_dataset[3])]
            nv = len(neg[0])
        self._neg = neg
        self._h = L.ref_sampler_create_opts(len(self.vid), K, _p(self.vid), _p(self.off), _p(self.sid), _p(self.feat), batch_size,
                                            context_size, num_negative_samples, max_buffer_size, negative_swap_percentage,
                                            max_same_video_negs, context_type, int(rand_skip), C.c_uint(caffe_seed), nv,
                                            *[(_p(a) if a is not None else None) for a in neg])
        if not self._h:
            raise RuntimeError("reference data layer failed to set up")
        self.B, self.R, self.K = batch_size, L.ref_sampler_rows(self._h), K

    def next(self):
        data = np.empty((self.B, self.R, self.K), np.float32)
        assert lib().ref_sampler_next(self._h, _p(data)) == 0
        return data

    def close(self):
        if self._h:
            lib().ref_sampler_destroy(self._h); self._h = None


class TestLayer:
    """The REFERENCE's VideoShotWindowTestDataLayer (compiled unmodified) over the fake LMDB: records of
    TestVideoShotWindows built from data [n, ctx+pos+neg, K]; next() -> (data blob [B, rows, K], labels [B])."""

    def __init__(self, data, video_id, pos_id, neg_id, ctx, pos, neg, batch_size, include_positives=True, include_negatives=True):
        L = lib()
        L.ref_testlayer_create.restype = C.c_void_p
        L.ref_testlayer_next.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_testlayer_rows.argtypes = [C.c_void_p]
        L.ref_testlayer_destroy.argtypes = [C.c_void_p]
        self.data = f32(data); n, rows, K = self.data.shape
        assert rows == ctx + pos + neg
        self.vid = np.ascontiguousarray(video_id, np.int32)
        self.pid = np.ascontiguousarray(pos_id, np.int32); self.nid = np.ascontiguousarray(neg_id, np.int32)
        self._h = L.ref_testlayer_create(n, ctx, pos, neg, K, _p(self.data), _p(self.vid), _p(self.pid), _p(self.nid), batch_size,
                                         int(include_positives), int(include_negatives))
        if not self._h:
            raise RuntimeError("reference test data layer failed to set up")
        self.B, self.R, self.K = batch_size, L.ref_testlayer_rows(self._h), K

    def next(self):
        data = np.empty((self.B, self.R, self.K), np.float32); lab = np.empty(self.B, np.float32)
        assert lib().ref_testlayer_next(self._h, _p(data), _p(lab)) == 0
        return data, lab

    def close(self):
        if self._h:
            lib().ref_testlayer_destroy(self._h); self._h = None


class Solver:
    """The REFERENCE's whole training pipeline: its VideoSampledShotsDataLayer (fake LMDB, libc rand()), Net (net.cpp,
    insert_splits.cpp) and SGDSolver (solver.cpp: learning-rate policy, weight decay, momentum, Net::Update), compiled
    unmodified, on the shipped TRAIN graph without the dropout layer.  step() = one iteration of Solver::Solve's loop."""
    POLICY = {"fixed": 0, "inv": 1, "step": 2}

    def __init__(self, video_id, shot_off, shot_ids, feat, W0, b0, batch_size, context_size=5, num_negative_samples=10,
                 max_buffer_size=5000, negative_swap_percentage=50, max_same_video_negs=6, context_type=1, margin=2.0,
                 norm=2, base_lr=0.001, momentum=0.9, weight_decay=0.0005, lr_policy="inv", gamma=0.001, power=0.75,
                 stepsize=1, seed=1, dropout_ratio=0.0, test=None, reg_type=2, library=None):
        """test = dict(data=[n, frames, K], video_id=[n], batch=.., id_to_class_file=path, exclude_same=True): adds the shipped
        file's TEST-phase graph (TestVideoShotWindows records on a second fake LMDB); see test()."""
        L = self._L = library if library is not None else lib()      # library=dropin_lib(): the GPU-mode build
        L.ref_solver_create.restype = C.c_void_p
        L.ref_solver_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_solver_get.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.ref_solver_num_layers.argtypes = [C.c_void_p]
        L.ref_solver_layer_name.argtypes = [C.c_void_p, C.c_int]; L.ref_solver_layer_name.restype = C.c_char_p
        L.ref_solver_destroy.argtypes = [C.c_void_p]
        self.vid = np.ascontiguousarray(video_id, np.int32); self.off = np.ascontiguousarray(shot_off, np.int32)
        self.sid = np.ascontiguousarray(shot_ids, np.int32); self.feat = f32(feat)
        W0, b0 = f32(W0), f32(b0)
        self.N, self.K = W0.shape
        self.B, self.R = batch_size, context_size + num_negative_samples
        L.ref_srand(C.c_uint(seed))
        fl = C.c_float
        if test:
            self.tdata = f32(test["data"]); self.tvid = np.ascontiguousarray(test["video_id"], np.int32)
            n, frames, tk = self.tdata.shape
            assert tk == self.K
            targs = (n, frames, _p(self.tdata), _p(self.tvid), int(test["batch"]), str(test["id_to_class_file"]).encode(),
                     int(test.get("exclude_same", True)))
        else:
            targs = (0, 0, None, None, 0, None, 0)
        L.ref_solver_test.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_solver_test_num_layers.argtypes = [C.c_void_p]
        L.ref_solver_test_layer_name.argtypes = [C.c_void_p, C.c_int]; L.ref_solver_test_layer_name.restype = C.c_char_p
        L.ref_solver_test_output_name.argtypes = [C.c_void_p, C.c_int]; L.ref_solver_test_output_name.restype = C.c_char_p
        self._h = L.ref_solver_create(len(self.vid), self.K, _p(self.vid), _p(self.off), _p(self.sid), _p(self.feat), batch_size,
                                      context_size, num_negative_samples, self.N, max_buffer_size, negative_swap_percentage,
                                      max_same_video_negs, context_type, fl(margin), norm, fl(base_lr), fl(momentum),
                                      fl(weight_decay), self.POLICY[lr_policy], fl(gamma), fl(power), stepsize, _p(W0), _p(b0),
                                      fl(dropout_ratio), *targs, int(reg_type))
        if not self._h:
            raise RuntimeError("reference solver failed to set up")

    def step(self):
        loss, viol = C.c_float(0), C.c_float(0)
        assert self._L.ref_solver_step(self._h, C.addressof(loss), C.addressof(viol)) == 0
        return loss.value, viol.value

    def state(self, want_data=False):
        W = np.empty((self.N, self.K), np.float32); b = np.empty(self.N, np.float32)
        hW = np.empty_like(W); hb = np.empty_like(b)
        data = np.empty((self.B, self.R, self.K), np.float32) if want_data else None
        assert self._L.ref_solver_get(self._h, _p(W), _p(b), _p(hW), _p(hb), _p(data)) == 0
        return dict(W=W, b=b, hW=hW, hb=hb, data=data)

    def layer_names(self):
        return [self._L.ref_solver_layer_name(self._h, i).decode() for i in range(self._L.ref_solver_num_layers(self._h))]

    def blob_names(self):
        self._L.ref_solver_blob_name.restype = C.c_char_p; self._L.ref_solver_blob_name.argtypes = [C.c_void_p, C.c_int]
        self._L.ref_solver_num_blobs.argtypes = [C.c_void_p]
        return [self._L.ref_solver_blob_name(self._h, i).decode() for i in range(self._L.ref_solver_num_blobs(self._h))]

    def blob(self, name, diff=False):
        """A blob of the TRAIN net as the host sees it after the last step (data or diff)."""
        self._L.ref_solver_blob.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int]
        n = self._L.ref_solver_blob(self._h, name.encode(), int(diff), None, 0)
        assert n >= 0, name
        out = np.empty(n, np.float32)
        assert self._L.ref_solver_blob(self._h, name.encode(), int(diff), _p(out), n) == n
        return out

    def test(self, iters):
        """Solver::Test's loop on the TEST net (weights shared with the TRAIN net): mean (mAP, hit@1, hit@5) over iters."""
        out = np.zeros(3, np.float32)
        assert self._L.ref_solver_test(self._h, iters, _p(out)) == 0
        return out

    def test_output_names(self):
        """in the order the reference's Net reports its outputs (Net::Init's std::set of blob names: lexicographic)"""
        return [self._L.ref_solver_test_output_name(self._h, j).decode() for j in range(3)]

    def test_layer_names(self):
        return [self._L.ref_solver_test_layer_name(self._h, i).decode() for i in range(self._L.ref_solver_test_num_layers(self._h))]

    def close(self):
        if self._h:
            self._L.ref_solver_destroy(self._h); self._h = None
