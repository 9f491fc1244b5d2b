"""ctypes wrapper over the caffe_compat host (csrc/host/caffe_compat/): Net / SGDSolver with the reference's
interface, built from prototxt text.  Used by the tests; the C++ CLI is build/vv_caffe."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import VVError, PREC

_P = C.c_void_p


def _l():
    lib = _lib.load()
    if getattr(lib, "_vvc_ready", False):
        return lib
    lib.vvc_last_error.restype = C.c_char_p
    lib.vvc_net_create.restype = _P; lib.vvc_net_create.argtypes = [C.c_char_p, C.c_int]
    lib.vvc_solver_create.restype = _P; lib.vvc_solver_create.argtypes = [C.c_char_p, C.c_char_p]
    lib.vvc_solver_net.restype = _P; lib.vvc_solver_net.argtypes = [_P]
    lib.vvc_net_layer_name.restype = C.c_char_p; lib.vvc_net_layer_name.argtypes = [_P, C.c_int]
    lib.vvc_net_blob_name.restype = C.c_char_p; lib.vvc_net_blob_name.argtypes = [_P, C.c_int]
    lib.vvc_solver_learning_rate.restype = C.c_float; lib.vvc_solver_learning_rate.argtypes = [_P]
    for name in ("vvc_net_destroy", "vvc_solver_destroy"):
        getattr(lib, name).argtypes = [_P]; getattr(lib, name).restype = None
    for name in ("vvc_net_num_layers", "vvc_net_num_blobs", "vvc_net_num_params", "vvc_solver_iter"):
        getattr(lib, name).argtypes = [_P]
    for name in ("vvc_net_layer_type", "vvc_net_layer_need_backward", "vvc_net_param_count"):
        getattr(lib, name).argtypes = [_P, C.c_int]
    lib.vvc_net_blob_shape.argtypes = [_P, C.c_char_p, _P]
    lib.vvc_net_blob_read.argtypes = [_P, C.c_char_p, C.c_int, _P]
    lib.vvc_net_param_read.argtypes = [_P, C.c_int, C.c_int, _P]
    lib.vvc_net_param_write.argtypes = [_P, C.c_int, _P]
    lib.vvc_net_set_dropout_mask.argtypes = [_P, _P]
    lib.vvc_net_enable_fusion.argtypes = [_P, _P, C.c_int]
    lib.vvc_net_forward_backward.argtypes = [_P, _P]
    lib.vvc_net_forward.argtypes = [_P, _P]
    lib.vvc_solver_step.argtypes = [_P, _P]
    lib.vvc_solver_solve.argtypes = [_P, C.c_int]
    lib.vvc_solver_history_read.argtypes = [_P, C.c_int, _P]
    lib.vvc_transform_net.argtypes = [C.c_char_p, C.c_int, _P, C.c_int]
    lib.vvc_set_stream.argtypes = [_P]
    lib.vvc_solver_test.argtypes = [_P, C.c_int, _P, C.c_int]
    lib.vvc_solver_test_net.restype = _P; lib.vvc_solver_test_net.argtypes = [_P, C.c_int]
    lib.vvc_net_share_trained_layers_with.argtypes = [_P, _P]
    lib.vvc_net_save.argtypes = [_P, C.c_char_p, C.c_int]
    lib.vvc_net_copy_trained_from.argtypes = [_P, C.c_char_p]
    lib.vvc_solver_snapshot.argtypes = [_P, _P, C.c_int]
    lib.vvc_solver_restore.argtypes = [_P, C.c_char_p]
    lib.vvc_solver_solve_resume.argtypes = [_P, C.c_int, C.c_char_p]
    lib.vvc_pb_open.restype = _P; lib.vvc_pb_open.argtypes = [C.c_char_p, C.c_char_p]
    lib.vvc_pb_close.argtypes = [_P]; lib.vvc_pb_close.restype = None
    lib.vvc_pb_text.argtypes = [_P, _P, C.c_int]
    lib.vvc_pb_num_arrays.argtypes = [_P]
    lib.vvc_pb_array_info.argtypes = [_P, C.c_int, _P, C.c_int]
    lib.vvc_pb_array_read.argtypes = [_P, C.c_int, _P]
    lib.vvc_pb_write.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, _P, _P, _P]
    lib._vvc_ready = True
    return lib


def _check(rc):
    if rc < 0:
        raise VVError("caffe host: " + _l().vvc_last_error().decode())
    return rc


def set_device(dev=0):
    _check(_l().vvc_set_device(dev))


def set_phase(phase):
    _l().vvc_set_phase(1 if phase == "TEST" else 0)


def set_seed(seed):
    _l().vvc_set_seed(seed)


def set_precision(prec):
    _l().vvc_set_precision(PREC[prec] if isinstance(prec, str) else prec)


def run_layer(layer_prototxt, bottoms, n_top, propagate_down=None, top_diffs=None, cap=1 << 20, forwards=1):
    """One layer by itself, as the reference's per-layer tests drive it (src/caffe/test/test_*_layer.cpp): `bottoms` are
    float32 arrays with up to 4 dims (num, channels, height, width), `layer_prototxt` a `layers { ... }` entry.
    forwards: Forward is run that many times before the tops are read (the n-th batch of a data layer).
    Returns (loss, [top data], [bottom diff or None])."""
    lib = _l()
    nb = len(bottoms)
    arrs = [np.ascontiguousarray(b, np.float32) for b in bottoms]
    shapes = (C.c_int * (4 * nb))(*[d for a in arrs for d in (list(a.shape) + [1, 1, 1, 1])[:4]])
    bptr = (_P * nb)(*[a.ctypes.data for a in arrs])
    tops = [np.zeros(cap, np.float32) for _ in range(n_top)]
    tptr = (_P * n_top)(*[t.ctypes.data for t in tops])
    counts = (C.c_int * n_top)()
    tdiff = None
    if top_diffs is not None:
        td = [None if d is None else np.ascontiguousarray(d, np.float32) for d in top_diffs]
        tdiff = (_P * n_top)(*[None if d is None else d.ctypes.data for d in td])
    pd = bd = None
    bdiffs = [None] * nb
    if propagate_down is not None:
        pd = (C.c_int * nb)(*[int(bool(x)) for x in propagate_down])
        bdiffs = [np.zeros(a.size, np.float32) if propagate_down[i] else None for i, a in enumerate(arrs)]
        bd = (_P * nb)(*[None if d is None else d.ctypes.data for d in bdiffs])
    loss = C.c_float(0)
    lib.vvc_layer_run.argtypes = [C.c_char_p, C.c_int, _P, _P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, C.c_int]
    _check(lib.vvc_layer_run(layer_prototxt.encode(), nb, shapes, bptr, n_top, cap, tptr, counts, tdiff, pd, bd, C.byref(loss),
                             int(forwards)))
    return (loss.value, [t[:counts[i]].copy() for i, t in enumerate(tops)],
            [None if d is None else d.reshape(arrs[i].shape) for i, d in enumerate(bdiffs)])


def transform_net(prototxt, phase="TRAIN"):
    """FilterNet(phase) + InsertSplits on prototxt text (host only, no GPU needed)."""
    buf = C.create_string_buffer(1 << 22)
    _check(_l().vvc_transform_net(prototxt.encode(), 1 if phase == "TEST" else 0, buf, len(buf)))
    return buf.value.decode()


class Net:
    def __init__(self, prototxt=None, phase="TRAIN", handle=None, owner=None):
        self._lib = _l()
        self._owner = owner
        if handle is None:
            handle = self._lib.vvc_net_create(prototxt.encode(), 1 if phase == "TEST" else 0)
            if not handle:
                raise VVError("caffe host: " + self._lib.vvc_last_error().decode())
            self._own = True
        else:
            self._own = False
        self._h = handle

    @property
    def layer_names(self):
        return [self._lib.vvc_net_layer_name(self._h, i).decode() for i in range(self._lib.vvc_net_num_layers(self._h))]

    @property
    def layer_types(self):
        return [self._lib.vvc_net_layer_type(self._h, i) for i in range(self._lib.vvc_net_num_layers(self._h))]

    @property
    def layer_need_backward(self):
        return [bool(self._lib.vvc_net_layer_need_backward(self._h, i)) for i in range(self._lib.vvc_net_num_layers(self._h))]

    @property
    def blob_names(self):
        return [self._lib.vvc_net_blob_name(self._h, i).decode() for i in range(self._lib.vvc_net_num_blobs(self._h))]

    @property
    def num_params(self):
        return self._lib.vvc_net_num_params(self._h)

    def blob(self, name, diff=False):
        shape = (C.c_int * 4)()
        n = _check(self._lib.vvc_net_blob_shape(self._h, name.encode(), shape))
        out = np.empty(n, np.float32)
        _check(self._lib.vvc_net_blob_read(self._h, name.encode(), int(diff), out.ctypes.data))
        return out.reshape([s for s in shape])

    def param(self, i, diff=False):
        out = np.empty(self._lib.vvc_net_param_count(self._h, i), np.float32)
        _check(self._lib.vvc_net_param_read(self._h, i, int(diff), out.ctypes.data))
        return out

    def set_param(self, i, value):
        v = np.ascontiguousarray(value, np.float32).reshape(-1)
        assert v.size == self._lib.vvc_net_param_count(self._h, i)
        _check(self._lib.vvc_net_param_write(self._h, i, v.ctypes.data))

    def set_dropout_mask(self, mask_dev):
        self._mask = mask_dev            # keep the tensor alive
        _check(self._lib.vvc_net_set_dropout_mask(self._h, C.c_void_p(mask_dev.data_ptr())))

    def enable_fusion(self):
        why = C.create_string_buffer(512)
        ok = _check(self._lib.vvc_net_enable_fusion(self._h, why, 512))
        return bool(ok), why.value.decode()

    def forward_backward(self):
        loss = C.c_float(0)
        _check(self._lib.vvc_net_forward_backward(self._h, C.byref(loss)))
        return loss.value

    def forward(self):
        loss = C.c_float(0)
        _check(self._lib.vvc_net_forward(self._h, C.byref(loss)))
        return loss.value

    def save(self, path, write_diff=False):
        """Net::ToProto + WriteProtoToBinaryFile: a .caffemodel the reference's CopyTrainedLayersFrom loads."""
        _check(self._lib.vvc_net_save(self._h, str(path).encode(), int(write_diff)))

    def copy_trained_from(self, path):
        _check(self._lib.vvc_net_copy_trained_from(self._h, str(path).encode()))

    def close(self):
        if self._h and self._own:
            self._lib.vvc_net_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Solver:
    def __init__(self, solver_prototxt, net_prototxt):
        self._lib = _l()
        self._h = self._lib.vvc_solver_create(solver_prototxt.encode(), net_prototxt.encode())
        if not self._h:
            raise VVError("caffe host: " + self._lib.vvc_last_error().decode())
        self.net = Net(handle=self._lib.vvc_solver_net(self._h), owner=self)

    def step(self):
        loss = C.c_float(0)
        _check(self._lib.vvc_solver_step(self._h, C.byref(loss)))
        return loss.value

    def solve(self, max_iter, resume=None):
        if resume is None:
            _check(self._lib.vvc_solver_solve(self._h, max_iter))
        else:
            _check(self._lib.vvc_solver_solve_resume(self._h, max_iter, str(resume).encode()))

    def test(self, test_net_id=0):
        """Solver::Test: test_iter forward passes of the TEST net on the shared weights; mean of every output value."""
        out = np.zeros(16, np.float32)
        n = _check(self._lib.vvc_solver_test(self._h, test_net_id, out.ctypes.data, 16))
        return out[:n].copy()

    def test_net(self, i=0):
        h = self._lib.vvc_solver_test_net(self._h, i)
        return Net(handle=h, owner=self) if h else None

    def snapshot(self):
        """Solver::Snapshot: writes <snapshot_prefix>_iter_N.caffemodel / .solverstate, returns the model path."""
        buf = C.create_string_buffer(1024)
        _check(self._lib.vvc_solver_snapshot(self._h, buf, 1024))
        return buf.value.decode()

    def restore(self, state_path):
        _check(self._lib.vvc_solver_restore(self._h, str(state_path).encode()))

    @property
    def iter(self):
        return self._lib.vvc_solver_iter(self._h)

    def learning_rate(self):
        return float(self._lib.vvc_solver_learning_rate(self._h))

    def history(self, i):
        out = np.empty(self._lib.vvc_net_param_count(self.net._h, i), np.float32)
        _check(self._lib.vvc_solver_history_read(self._h, i, out.ctypes.data))
        return out

    def close(self):
        if self._h:
            self._lib.vvc_solver_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- binary protobuf files (.caffemodel = NetParameter, .solverstate = SolverState), host only ----------------------
def read_binary_proto(path, type_name):
    """Returns (text-format string of the message without float arrays, {dotted path: float32 array})."""
    lib = _l()
    h = lib.vvc_pb_open(str(path).encode(), type_name.encode())
    if not h:
        raise VVError("caffe host: " + lib.vvc_last_error().decode())
    try:
        buf = C.create_string_buffer(1 << 22)
        _check(lib.vvc_pb_text(h, buf, len(buf)))
        arrays = {}
        for i in range(lib.vvc_pb_num_arrays(h)):
            name = C.create_string_buffer(512)
            n = _check(lib.vvc_pb_array_info(h, i, name, 512))
            a = np.empty(n, np.float32)
            _check(lib.vvc_pb_array_read(h, i, a.ctypes.data))
            arrays[name.value.decode()] = a
        return buf.value.decode(), arrays
    finally:
        lib.vvc_pb_close(h)


def write_binary_proto(path, type_name, text, arrays=None):
    """Writes message `type_name` from text format plus float arrays keyed by dotted path ("layers[0].blobs[1].data")."""
    lib = _l()
    arrays = arrays or {}
    keys = list(arrays.keys())
    vals = [np.ascontiguousarray(arrays[k], np.float32).reshape(-1) for k in keys]
    kp = (C.c_char_p * len(keys))(*[k.encode() for k in keys])
    vp = (C.c_void_p * len(keys))(*[v.ctypes.data for v in vals])
    cn = (C.c_int * len(keys))(*[v.size for v in vals])
    _check(lib.vvc_pb_write(str(path).encode(), type_name.encode(), text.encode(), len(keys), kp, vp, cn))
