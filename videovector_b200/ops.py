"""Thin torch-tensor wrappers over the C-ABI (PyTorch supplies device memory and
streams only; every op below is one call into libvv_b200.so)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (Act, Operand, RankCfg, TrainerCfg, PREC, DROPOUT_NONE, DROPOUT_MASK01,
                   DROPOUT_MASK_U32, DROPOUT_PHILOX, DROPOUT_HASH, VVError, check)


def _ptr(t):
    if t is None:
        return None
    assert t.is_contiguous(), "tensor must be contiguous"
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _prec(p):
    return PREC[p] if isinstance(p, str) else int(p)


class OperandT:
    """GEMM operand copies of an fp32 tensor for a precision (see vv_operand_t)."""

    def __init__(self, hi, lo=None, block=None):
        self.hi, self.lo, self.block = hi, lo, block       # block: the f16x3 allocation (header + planes)

    @property
    def scale(self):
        """f16x3: the power-of-two scale in the operand's header (device -> host read)."""
        return float(self.block[:8].view(torch.float32)[0].item())

    @property
    def absmax(self):
        return float(self.block[8:12].view(torch.float32)[0].item())

    def dequant(self):
        """f16x3: (h0 + h1) / scale as fp32 -- what the GEMM effectively multiplies."""
        return (self.hi.float() + self.lo.float()) / self.scale

    def c(self):
        return Operand(_ptr(self.hi), _ptr(self.lo))


def alloc_operand(shape, prec, device="cuda"):
    p = _prec(prec)
    if p == PREC["tf32x3"]:
        return OperandT(torch.empty(shape, dtype=torch.float32, device=device),
                        torch.empty(shape, dtype=torch.float32, device=device))
    if p == PREC["bf16"]:
        return OperandT(torch.empty(shape, dtype=torch.bfloat16, device=device))
    if p == PREC["f16x3"]:
        n = 1
        for d in shape:
            n *= int(d)
        ho, lo_ = C.c_size_t(0), C.c_size_t(0)
        nbytes = _lib.load().vv_operand_bytes(n, p, C.byref(ho), C.byref(lo_))
        block = torch.zeros(nbytes, dtype=torch.uint8, device=device)            # zero header = "scale not set"
        hi = block[ho.value:ho.value + 2 * n].view(torch.float16).view(shape)
        lo = block[lo_.value:lo_.value + 2 * n].view(torch.float16).view(shape)
        return OperandT(hi, lo, block)
    return None


def operand_rescale(op, prec, target_log2=10):
    check(_lib.load().vv_operand_rescale(_ptr(op.hi), _prec(prec), target_log2, _stream()))


def operand_measure(op, prec, src):
    check(_lib.load().vv_operand_measure(_ptr(op.hi), _prec(prec), _ptr(src), src.numel(), _stream()))


def prepare_operand(x, prec):
    """fp32 tensor -> OperandT for `prec` (FP32_SIMT / TF32 use the tensor itself)."""
    p = _prec(prec)
    if p in (PREC["fp32_simt"], PREC["tf32"]):
        return OperandT(x)
    op = alloc_operand(x.shape, p, x.device)
    check(_lib.load().vv_prepare_operand(_ptr(x), x.numel(), p, _ptr(op.hi), _ptr(op.lo), _stream()))
    return op


def gather_rows(bank, idx, quirk, prec="fp32_simt", want_x=True, want_blob=False):
    """K0.  bank [rows,K] fp32, idx/quirk [B,R] int32 (device).  Returns (X, operand, blob)."""
    lib = _lib.load()
    B, R = idx.shape
    K = bank.shape[1]
    p = _prec(prec)
    X = torch.empty((R * B, K), dtype=torch.float32, device=bank.device) if want_x else None
    op = alloc_operand((R * B, K), p, bank.device)
    blob = torch.empty((B, R, K), dtype=torch.float32, device=bank.device) if want_blob else None
    if p == PREC["f16x3"]:                       # the X scale comes from max|bank|
        operand_measure(op, p, bank); operand_rescale(op, p, 12)
    check(lib.vv_gather_rows(_ptr(bank), bank.shape[0], K, _ptr(idx), _ptr(quirk), B, R, _ptr(X),
                             _ptr(op.hi) if op else None, _ptr(op.lo) if op else None, p, _ptr(blob), _stream()))
    if op is None and X is not None:
        op = OperandT(X)
    return X, op, blob


def make_act(relu=True, negative_slope=0.0, dropout_mode=DROPOUT_NONE, ratio=0.0, mask=None, mask_out=None,
             seed=0, step=0):
    a = Act()
    a.relu = int(relu); a.negative_slope = negative_slope; a.dropout_mode = dropout_mode
    a.dropout_ratio = ratio
    a.mask = mask.data_ptr() if mask is not None else None
    a.mask_out = mask_out.data_ptr() if mask_out is not None else None
    a.seed = seed; a.step = step
    return a


def ip_forward(X, W, bias, M, N, K, prec, act=None, want_z=False):
    """K1 forward.  X, W are OperandT.  Returns (H, Z)."""
    dev = W.hi.device
    H = torch.empty((M, N), dtype=torch.float32, device=dev)
    Z = torch.empty((M, N), dtype=torch.float32, device=dev) if (want_z and act is not None) else None
    check(_lib.load().vv_ip_forward(X.c(), W.c(), _ptr(bias), M, N, K, _prec(prec),
                                    C.byref(act) if act is not None else None, _ptr(Z), _ptr(H), _stream()))
    return H, Z


def ip_wgrad(dZ, X, M, N, K, prec, regularization=0.0, nsplit=0):
    """K1 wgrad.  Returns dW [N,K] (slabs reduced)."""
    lib = _lib.load()
    dev = X.hi.device
    p = _prec(prec)
    if nsplit == 0:
        ws_bytes = lib.vv_ip_wgrad_workspace_bytes(M, N, K, p)
        ws = torch.empty((max(ws_bytes, 4) // 4,), dtype=torch.float32, device=dev)
        dW = torch.empty((N, K), dtype=torch.float32, device=dev)
        check(lib.vv_ip_wgrad(dZ.c(), X.c(), M, N, K, p, regularization, _ptr(dW), 0, _ptr(ws), ws_bytes, _stream()))
        return dW
    parts = torch.empty((nsplit, N, K), dtype=torch.float32, device=dev)
    check(lib.vv_ip_wgrad(dZ.c(), X.c(), M, N, K, p, regularization, _ptr(parts), nsplit, None, 0, _stream()))
    return parts.sum(0) if nsplit > 1 else parts[0]


def ip_bias_grad(dZ):
    M, N = dZ.shape
    db = torch.empty((N,), dtype=torch.float32, device=dZ.device)
    check(_lib.load().vv_ip_bias_grad(_ptr(dZ), M, N, _ptr(db), _stream()))
    return db


def ip_dgrad(dZ, W, M, N, K, prec):
    dX = torch.empty((M, K), dtype=torch.float32, device=W.hi.device)
    check(_lib.load().vv_ip_dgrad(dZ.c(), W.c(), M, N, K, _prec(prec), _ptr(dX), _stream()))
    return dX


def rank_cfg(B, Cc, Nn, N, margin=2.0, norm=2, coeff=None, eps=1e-10):
    c = RankCfg()
    c.B, c.C, c.Nn, c.N = B, Cc, Nn, N
    co = coeff if coeff is not None else [1.0 / (Cc - 1)] * (Cc - 1)
    for i, v in enumerate(co):
        c.coeff[i] = np.float32(v)
    c.margin = margin; c.norm = norm; c.eps = eps
    return c


def rank_loss_forward(H, cfg):
    """K2.  Returns dict(stats, target_score, neg_score, loss, violations)."""
    dev = H.device
    B, Nn = cfg.B, cfg.Nn
    stride = 1 + 2 * (1 + Nn)
    out = dict(stats=torch.empty((B, stride), dtype=torch.float32, device=dev),
               target_score=torch.empty((B, Nn), dtype=torch.float32, device=dev),
               neg_score=torch.empty((B, Nn), dtype=torch.float32, device=dev),
               item_loss=torch.empty((B,), dtype=torch.float32, device=dev),
               item_viol=torch.empty((B,), dtype=torch.float32, device=dev),
               loss=torch.empty((1,), dtype=torch.float32, device=dev),
               violations=torch.empty((1,), dtype=torch.float32, device=dev))
    check(_lib.load().vv_rank_loss_forward(_ptr(H), C.byref(cfg), _ptr(out["stats"]), _ptr(out["target_score"]),
                                           _ptr(out["neg_score"]), _ptr(out["item_loss"]), _ptr(out["item_viol"]),
                                           _ptr(out["loss"]), _ptr(out["violations"]), _stream()))
    return out


def rank_loss_backward(H, cfg, stats, loss_weight=1.0, act_fused=True, dropout_scale=1.0, prec="fp32_simt",
                       want_db=True):
    """K3.  Returns (dZ fp32, operand copies or None, db or None)."""
    dev = H.device
    p = _prec(prec)
    dZ = torch.empty_like(H)
    op = alloc_operand(H.shape, p, dev)
    db = torch.zeros((cfg.N,), dtype=torch.float32, device=dev) if want_db else None
    for _pass in range(2 if p == PREC["f16x3"] else 1):     # f16x3: a first pass measures max|dZ| for the scale
        if db is not None:
            db.zero_()
        if _pass == 1:
            operand_rescale(op, p)
        check(_lib.load().vv_rank_loss_backward(_ptr(H), C.byref(cfg), _ptr(stats), loss_weight, int(act_fused),
                                                dropout_scale, _ptr(dZ), _ptr(op.hi) if op else None,
                                                _ptr(op.lo) if op else None, p, _ptr(db), _stream()))
    return dZ, (op if op is not None else OperandT(dZ)), db


def rank_loss_fused_supported(cfg):
    return bool(_lib.load().vv_rank_loss_fused_supported(C.byref(cfg)))


def rank_loss_fused(H, cfg, loss_weight=1.0, act_fused=True, dropout_scale=1.0, prec="fp32_simt", want_db=True, want_dz=True,
                    want_scores=True):
    """K2 + K3 in one pass.  Returns (forward dict as rank_loss_forward, dZ fp32, operand copies, db).
    want_dz=False, want_scores=False: operand-only output -- the trainer's form, served by the second-generation kernel."""
    dev = H.device
    B, Nn = cfg.B, cfg.Nn
    p = _prec(prec)
    out = dict(stats=torch.empty((B, 1 + 2 * (1 + Nn)), dtype=torch.float32, device=dev),
               target_score=torch.empty((B, Nn), dtype=torch.float32, device=dev),
               neg_score=torch.empty((B, Nn), dtype=torch.float32, device=dev),
               item_loss=torch.empty((B,), dtype=torch.float32, device=dev),
               item_viol=torch.empty((B,), dtype=torch.float32, device=dev),
               loss=torch.empty((1,), dtype=torch.float32, device=dev),
               violations=torch.empty((1,), dtype=torch.float32, device=dev))
    dZ = torch.empty_like(H) if want_dz else None
    if not want_scores:
        out["target_score"] = out["neg_score"] = None
    op = alloc_operand(H.shape, p, dev)
    db = torch.zeros((cfg.N,), dtype=torch.float32, device=dev) if want_db else None
    for _pass in range(2 if p == PREC["f16x3"] else 1):     # f16x3: a first pass measures max|dZ| for the scale
        if db is not None:
            db.zero_()
        if _pass == 1:
            operand_rescale(op, p)
        check(_lib.load().vv_rank_loss_fused(_ptr(H), C.byref(cfg), loss_weight, int(act_fused), dropout_scale,
                                             _ptr(out["stats"]), _ptr(out["target_score"]), _ptr(out["neg_score"]),
                                             _ptr(out["item_loss"]), _ptr(out["item_viol"]), _ptr(out["loss"]),
                                             _ptr(out["violations"]), _ptr(dZ), _ptr(op.hi) if op else None,
                                             _ptr(op.lo) if op else None, p, _ptr(db), None, None, _stream()))
    return out, dZ, (op if op is not None else OperandT(dZ)), db


def sgd_update(W, grad_parts, hist, local_rate, momentum, local_decay, reg_type=2, grad_scale=1.0, prec="fp32_simt",
               Wop=None, diff_out=None):
    """K4 (in place on W, hist).  grad_parts [S, ...] or [...]."""
    count = W.numel()
    nparts = grad_parts.numel() // count
    if Wop is not None and _prec(prec) == PREC["f16x3"]:
        operand_rescale(Wop, prec)
    check(_lib.load().vv_sgd_update(_ptr(W), _ptr(grad_parts), nparts, count, _ptr(hist), _ptr(diff_out), count,
                                    local_rate, momentum, local_decay, reg_type, grad_scale,
                                    _ptr(Wop.hi) if Wop is not None else None,
                                    _ptr(Wop.lo) if (Wop is not None and Wop.lo is not None) else None,
                                    _prec(prec), _stream()))


def learning_rate(policy, base_lr, gamma, power, stepsize, it):
    return float(_lib.load().vv_learning_rate(policy.encode(), base_lr, gamma, power, stepsize, it))


def dropout_make_mask(rows, cols, ratio, seed, step, device="cuda", mode=DROPOUT_PHILOX):
    m = torch.empty((rows, cols), dtype=torch.int32, device=device)
    check(_lib.load().vv_dropout_make_mask_mode(_ptr(m), rows, cols, ratio, seed, step, mode, _stream()))
    return m


def fill_bank(rows, K, seed, device="cuda"):
    bank = torch.empty((rows, K), dtype=torch.float32, device=device)
    check(_lib.load().vv_fill_bank(_ptr(bank), rows, K, seed, _stream()))
    return bank


def bank_host(rows, K, seed):
    """numpy restatement of the synthetic bank hash (bit-identical to vv_fill_bank)."""
    e = np.arange(rows * K, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = np.uint64(seed) * np.uint64(0xD1342543DE82EF95) + e
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    m = np.uint64(0xFFFF)
    t = ((x & m) + ((x >> np.uint64(16)) & m) + ((x >> np.uint64(32)) & m) + ((x >> np.uint64(48)) & m)).astype(np.int64) - 131070
    v = t.astype(np.float32) * np.float32(np.float32(1.0) / np.float32(37837.0))
    return np.maximum(v, np.float32(0)).reshape(rows, K)


# ---------------------------------------------------------------------------------
# host sampler + synthetic dataset description
# ---------------------------------------------------------------------------------
def synthetic_videos(V, S):
    """V videos of S shots each: video_id = v, shot_ids = 0..S-1, bank row = v*S + s."""
    video_id = np.arange(V, dtype=np.int32)
    shot_off = (np.arange(V + 1, dtype=np.int64) * S).astype(np.int32)
    shot_ids = np.tile(np.arange(S, dtype=np.int32), V)
    return video_id, shot_off, shot_ids


class RecordSet:
    """The DB values the reference's data layers parse (VideoShots / TestVideoShotWindows protobuf records, in DB key
    order) -> host feature bank [rows, K] + the sampler's tables (vv_record_set_*, include/vv_b200.h)."""

    def __init__(self, kind="video_shots", include_positives=True, include_negatives=True):
        self._lib = _lib.load()
        k = {"video_shots": _lib.RECORD_VIDEO_SHOTS, "test_windows": _lib.RECORD_TEST_WINDOWS}[kind]
        self._h = self._lib.vv_record_set_create(k, int(include_positives), int(include_negatives))
        if not self._h:
            raise VVError(_lib.last_error())

    def add(self, value):
        check(self._lib.vv_record_set_add(self._h, bytes(value), len(value)))

    def load_file(self, path):
        check(self._lib.vv_record_set_load_file(self._h, str(path).encode()))
        return self

    def info(self):
        import ctypes as C
        rec, rows, K, rpr = C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
        check(self._lib.vv_record_set_info(self._h, C.addressof(rec), C.addressof(rows), C.addressof(K), C.addressof(rpr)))
        return dict(records=rec.value, rows=rows.value, feature_size=K.value, rows_per_record=rpr.value)

    def tables(self):
        i = self.info()
        vid = np.empty(i["records"], np.int32); off = np.empty(i["records"] + 1, np.int32); ids = np.empty(i["rows"], np.int32)
        check(self._lib.vv_record_set_tables(self._h, vid.ctypes.data, off.ctypes.data, ids.ctypes.data))
        return vid, off, ids

    def bank_host(self):
        import ctypes as C
        i = self.info()
        p = self._lib.vv_record_set_bank(self._h)
        if not p:
            return np.empty((0, i["feature_size"]), np.float32)
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(i["rows"], i["feature_size"]))
        return a.copy()

    def bank_device(self, device="cuda"):
        i = self.info()
        bank = torch.empty((i["rows"], i["feature_size"]), dtype=torch.float32, device=device)
        check(self._lib.vv_record_set_upload(self._h, bank.data_ptr(), _stream()))
        return bank

    def close(self):
        if self._h:
            self._lib.vv_record_set_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Sampler:
    """VideoSampledShotsDataLayer's sampler (all five context types) as an index stream (host, C++)."""

    def __init__(self, video_id, shot_off, shot_ids, batch_size, context_size=5, num_negative_samples=10,
                 max_buffer_size=5000, negative_swap_percentage=50, max_same_video_negs=6,
                 max_tries_for_negs=100, rand_seed=1, context_type="window", start_skip=0, negative_dataset=None):
        """start_skip: the data layer's rand_skip draw (the record cursor starts that many records in);
        negative_dataset: (video_id, shot_off, shot_ids, row_base) of a second record set whose shots seed the
        negative buffer, its rows living at [row_base, ...) of the bank."""
        self._lib = _lib.load()
        self.video_id = np.ascontiguousarray(video_id, dtype=np.int32)
        self.shot_off = np.ascontiguousarray(shot_off, dtype=np.int32)
        self.shot_ids = np.ascontiguousarray(shot_ids, dtype=np.int32)
        self.B, self.R = batch_size, context_size + num_negative_samples
        ct = _lib.CONTEXT[context_type] if isinstance(context_type, str) else int(context_type)
        nv, nvid, noff, nsid, nbase = 0, None, None, None, 0
        if negative_dataset is not None:
            self._neg = [np.ascontiguousarray(a, dtype=np.int32) for a in negative_dataset[:3]]
            nv, nvid, noff, nsid, nbase = len(self._neg[0]), self._neg[0].ctypes.data, self._neg[1].ctypes.data, self._neg[2].ctypes.data, int(negative_dataset[3])
        self._h = self._lib.vv_sampler_create_ex2(len(self.video_id), self.video_id.ctypes.data, self.shot_off.ctypes.data,
                                                  self.shot_ids.ctypes.data, batch_size, context_size, num_negative_samples,
                                                  max_buffer_size, negative_swap_percentage, max_same_video_negs,
                                                  max_tries_for_negs, rand_seed, ct, int(start_skip), nv, nvid, noff, nsid, nbase)
        if not self._h:
            raise VVError("vv_sampler_create failed (bad parameters, or could not fill the negative buffer)")

    def next(self):
        idx = np.empty((self.B, self.R), dtype=np.int32)
        quirk = np.empty((self.B, self.R), dtype=np.int32)
        check(self._lib.vv_sampler_next(self._h, idx.ctypes.data, quirk.ctypes.data))
        return idx, quirk

    def next_into(self, idx, quirk):
        check(self._lib.vv_sampler_next(self._h, idx.ctypes.data, quirk.ctypes.data))

    def prefetch(self, depth):
        """Start (depth > 0) / stop the native prefetch thread; next() then pops batches drawn ahead, same stream."""
        check(self._lib.vv_sampler_prefetch(self._h, int(depth)))

    @property
    def ready(self):
        """Batches the prefetch thread has drawn ahead and not yet handed out."""
        return self._lib.vv_sampler_prefetch_ready(self._h)

    @property
    def cursor(self):
        return self._lib.vv_sampler_cursor(self._h)

    def close(self):
        if self._h:
            self._lib.vv_sampler_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiSampler:
    """k reference-exact sampler streams over k disjoint sub-shards of the videos, one native prefetch thread each;
    batch i comes from stream i % k.  Each stream is the reference's sampler on its own videos (own rand() stream seeded
    rand_seed + s, own negative buffer); a single sequential stream tops out near 3.3 M items/s per host core, k streams
    feed a GPU k times faster.  streams=1 is exactly Sampler."""

    def __init__(self, video_id, shot_off, shot_ids, batch_size, context_size=5, num_negative_samples=10,
                 max_buffer_size=5000, negative_swap_percentage=50, max_same_video_negs=6,
                 max_tries_for_negs=100, rand_seed=1, context_type="window", streams=2, row_base=0):
        video_id = np.ascontiguousarray(video_id, dtype=np.int32)
        shot_off = np.ascontiguousarray(shot_off, dtype=np.int32)
        shot_ids = np.ascontiguousarray(shot_ids, dtype=np.int32)
        V = len(video_id)
        assert 1 <= streams <= V
        self.B, self.R = batch_size, context_size + num_negative_samples
        self.parts = []
        for s in range(streams):
            v0, v1 = s * V // streams, (s + 1) * V // streams
            sm = Sampler(video_id[v0:v1], shot_off[v0:v1 + 1] - shot_off[v0], shot_ids[shot_off[v0]:shot_off[v1]], batch_size,
                         context_size, num_negative_samples, max_buffer_size, negative_swap_percentage, max_same_video_negs,
                         max_tries_for_negs, rand_seed + s, context_type)
            check(sm._lib.vv_sampler_set_row_base(sm._h, int(row_base + shot_off[v0])))
            self.parts.append(sm)
        self._i = 0

    def next_into(self, idx, quirk):
        self.parts[self._i % len(self.parts)].next_into(idx, quirk)
        self._i += 1

    def next(self):
        idx = np.empty((self.B, self.R), dtype=np.int32); quirk = np.empty((self.B, self.R), dtype=np.int32)
        self.next_into(idx, quirk)
        return idx, quirk

    def prefetch(self, depth):
        """depth = batches drawn ahead over ALL streams (rounded up to a multiple of the stream count); 0 stops."""
        k = len(self.parts)
        for sm in self.parts:
            sm.prefetch((depth + k - 1) // k if depth > 0 else 0)

    @property
    def ready(self):
        return sum(sm.ready for sm in self.parts)

    def close(self):
        for sm in self.parts:
            sm.close()


# ---------------------------------------------------------------------------------
# trainer
# ---------------------------------------------------------------------------------
SOLVER_DEFAULTS = dict(lr_policy="inv", base_lr=1e-3, gamma=1e-3, power=0.75, stepsize=1,
                       momentum=0.9, weight_decay=5e-4, reg_type=2, lr_mult=(1.0, 2.0), decay_mult=(1.0, 0.0))


def trainer_cfg(B, C_=5, Nn=10, K=4096, N=512, margin=2.0, norm=2, dropout_ratio=0.9, dropout_mode=DROPOUT_PHILOX,
                dropout_seed=7, loss_weight=1.0, regularization=0.0, prec="f16x3", world_size=1, rank=0,
                compute_dgrad=False, keep_blobs=False, coeff=None, split_rank_loss=False, **solver):
    s = dict(SOLVER_DEFAULTS); s.update(solver)
    c = TrainerCfg()
    c.B, c.C, c.Nn, c.K, c.N = B, C_, Nn, K, N
    if coeff is not None:
        for i, v in enumerate(coeff):
            c.coeff[i] = v
    c.margin, c.norm = margin, norm
    c.dropout_ratio, c.dropout_mode, c.dropout_seed = dropout_ratio, dropout_mode, dropout_seed
    c.loss_weight, c.regularization = loss_weight, regularization
    c.lr_policy = s["lr_policy"].encode()
    c.base_lr, c.gamma, c.power, c.stepsize = s["base_lr"], s["gamma"], s["power"], s["stepsize"]
    c.momentum, c.weight_decay, c.reg_type = s["momentum"], s["weight_decay"], s["reg_type"]
    c.lr_mult[0], c.lr_mult[1] = s["lr_mult"]
    c.decay_mult[0], c.decay_mult[1] = s["decay_mult"]
    c.prec = _prec(prec); c.world_size, c.rank = world_size, rank
    c.compute_dgrad, c.keep_blobs = int(compute_dgrad), int(keep_blobs)
    c.split_rank_loss = int(split_rank_loss)
    return c


class Trainer:
    """One data-parallel rank of the fused training step (C++ object behind the C-ABI)."""

    def __init__(self, cfg, stream=None):
        self._lib = _lib.load()
        self.cfg = cfg
        self.stream = stream if stream is not None else torch.cuda.current_stream()
        self._h = self._lib.vv_trainer_create(C.byref(cfg), C.c_void_p(self.stream.cuda_stream))
        if not self._h:
            raise VVError("vv_trainer_create failed: " + _lib.last_error())
        self.R = cfg.C + cfg.Nn
        self.M = self.R * cfg.B

    def tensor(self, which):
        N, K, B = self.cfg.N, self.cfg.K, self.cfg.B
        L = self._lib
        table = {
            "W": (L.vv_trainer_weight, (N, K)), "b": (L.vv_trainer_bias, (N,)),
            "W_hist": (L.vv_trainer_weight_hist, (N, K)), "b_hist": (L.vv_trainer_bias_hist, (N,)),
            "W_diff": (L.vv_trainer_weight_diff, (N, K)), "b_diff": (L.vv_trainer_bias_diff, (N,)),
        }
        if which in table:
            fn, shape = table[which]
            return _device_view(fn(self._h), shape)
        shapes = {"X": (self.M, K), "Z": (self.M, N), "H": (self.M, N), "dZ": (self.M, N),
                  "stats": (B, 1 + 2 * (1 + self.cfg.Nn)), "loss": (1,), "violations": (1,),
                  "dW_raw": (N, K), "db_raw": (N,), "dX": (self.M, K), "db_raw_ext": (N + 2,), "wlast": (N,)}
        # raw operand planes of W as float32 words (f16x3 / bf16: 2-byte elements; tf32x3: hi fp32, lo two bf16 planes)
        pw = {PREC["f16x3"]: (N * K // 2, N * K // 2), PREC["bf16"]: (N * K // 2, 0), PREC["tf32x3"]: (N * K, N * K)}.get(self.cfg.prec, (0, 0))
        shapes["Wop_hi"], shapes["Wop_lo"] = (pw[0],), (pw[1],)
        MN = self.M * N
        pz = {PREC["f16x3"]: (MN // 2, MN // 2), PREC["bf16"]: (MN // 2, 0), PREC["tf32x3"]: (MN, MN)}.get(self.cfg.prec, (0, 0))
        shapes["dZop_hi"], shapes["dZop_lo"] = (pz[0],), (pz[1],)
        ptr = L.vv_trainer_blob(self._h, (b"db_raw" if which == "db_raw_ext" else which.encode()))
        if not ptr:
            raise VVError("trainer blob %s is not allocated in this configuration" % which)
        return _device_view(ptr, shapes[which])

    def dZ_from_operand(self):
        """fp32 dZ [M,N] reconstructed from the operand copy the wgrad multiplies (f16x3: (h0 + h1) / scale; bf16: the
        bf16 values) -- the gather-fused path keeps no fp32 dZ."""
        N = self.cfg.N
        hi = self.tensor("dZop_hi")
        if self.cfg.prec == PREC["f16x3"]:
            ptr = self._lib.vv_trainer_blob(self._h, b"dZop_hi")
            scale = _device_view(ptr - _lib.F16X3_HEADER_BYTES, (4,))[0]
            h0 = hi.view(torch.float16).float(); h1 = self.tensor("dZop_lo").view(torch.float16).float()
            return ((h0 + h1) / scale).view(self.M, N)
        if self.cfg.prec == PREC["bf16"]:
            return hi.view(torch.bfloat16).float().view(self.M, N)
        raise VVError("dZ_from_operand: f16x3 / bf16 only")

    def set_weights(self, W, b):
        self.tensor("W").copy_(W)
        self.tensor("b").copy_(b)
        check(self._lib.vv_trainer_sync_weights(self._h))

    def set_bank(self, bank):
        """Register the resident bank: later steps on it use the gather-fused GEMMs (no materialised X)."""
        self._bank = bank
        check(self._lib.vv_trainer_set_bank(self._h, _ptr(bank), bank.shape[0]))

    def step(self, bank, idx, quirk, mask=None, it=0, do_update=True):
        check(self._lib.vv_trainer_step(self._h, _ptr(bank), bank.shape[0], _ptr(idx), _ptr(quirk), _ptr(mask),
                                        it, int(do_update)))

    def extract(self, F):
        out = torch.empty((F.shape[0], self.cfg.N), dtype=torch.float32, device=F.device)
        check(self._lib.vv_trainer_extract(self._h, _ptr(F), F.shape[0], _ptr(out)))
        return out

    def extract_into(self, F, out):
        check(self._lib.vv_trainer_extract(self._h, _ptr(F), F.shape[0], _ptr(out)))

    @property
    def last_launches(self):
        return self._lib.vv_trainer_last_launches(self._h)

    PHASES = ("gather", "fc7_forward", "rank_loss_forward", "rank_loss_backward", "wgrad", "dgrad", "allreduce",
              "sgd_update")

    def set_timing(self, enable):
        check(self._lib.vv_trainer_set_timing(self._h, int(enable)))

    def phase_ms(self):
        """(dict phase -> mean ms per step, number of timed steps) since the last call."""
        ms = (C.c_float * 8)()
        n = C.c_int(0)
        check(self._lib.vv_trainer_phase_ms(self._h, C.addressof(ms), C.addressof(n)))
        steps = max(n.value, 1)
        return {name: ms[i] / steps for i, name in enumerate(self.PHASES)}, n.value

    def dp_init(self, id_bytes):
        buf = (C.c_char * 128).from_buffer_copy(id_bytes)
        check(self._lib.vv_dp_init(self._h, C.addressof(buf)))

    @property
    def dp_mode(self):
        """'single', 'nccl' (all-reduce between wgrad and update) or 'p2p' (exchange inside the update kernel over peer memory)."""
        return ("single", "nccl", "p2p")[self._lib.vv_dp_mode(self._h)]

    @property
    def dp_mode_reason(self):
        return self._lib.vv_dp_mode_reason(self._h).decode()

    def dp_gather_state(self):
        """Collective (every rank): all-gather the owner-sharded master weights / history of the p2p mode."""
        check(self._lib.vv_dp_gather_state(self._h))

    def close(self):
        if self._h:
            self._lib.vv_trainer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def dp_unique_id():
    buf = (C.c_char * 128)()
    check(_lib.load().vv_dp_unique_id(C.addressof(buf)))
    return bytes(buf.raw)


class _CudaArrayHolder:
    def __init__(self, ptr, shape):
        n = int(np.prod(shape))
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def _device_view(ptr, shape):
    """Zero-copy torch view of trainer-owned device memory."""
    t = torch.as_tensor(_CudaArrayHolder(ptr, shape), device="cuda")
    return t.view(*shape)


# ---- TEST-phase evaluation ------------------------------------------------------------------------------------
def gather_mean_rows(bank, idx, coeff=None):
    """average_for_test: mean (or coeff-weighted sum) of the F bank rows of every item.  idx [B,F] int32 (device)."""
    B, F = idx.shape
    K = bank.shape[1]
    out = torch.empty((B, K), dtype=torch.float32, device=bank.device)
    c = np.ascontiguousarray(coeff, np.float32) if coeff is not None else None
    check(_lib.load().vv_gather_mean_rows(_ptr(bank), bank.shape[0], K, _ptr(idx), B, F, c.ctypes.data if c is not None else None,
                                          _ptr(out), _stream()))
    return out


def test_embed(bank, idx, W, bias, prec="f16x3", coeff=None):
    """The TEST graph up to `ip2_norm`: frame mean -> fc7 + ReLU (dropout is a copy in TEST) -> L2 normalisation."""
    xbar = gather_mean_rows(bank, idx, coeff)
    M, K = xbar.shape
    N = W.shape[0]
    H, _ = ip_forward(prepare_operand(xbar, prec), prepare_operand(W, prec), bias, M, N, K, prec, act=make_act(True, 0.0, DROPOUT_NONE))
    E = torch.empty_like(H)
    check(_lib.load().vv_l2norm_forward(_ptr(H), M, N, _ptr(E), _stream()))
    return xbar, E


def retrieval_stats(E, video_ids, labels, exclude_same_video_shots=False, gram=None):
    """RetrievalStatsLayer on the device.  Returns dict(map, hit1, hit5, per_query [B,3] float64 tensor)."""
    lib = _lib.load()
    B, N = (E.shape if E is not None else (gram.shape[0], 0))
    dev = (E if E is not None else gram).device
    ws_bytes = lib.vv_retrieval_stats_workspace_bytes(B)
    ws = torch.empty((ws_bytes + 7) // 8, dtype=torch.float64, device=dev)
    out = torch.zeros(3, dtype=torch.float64, device=dev)
    pq = torch.empty((B, 3), dtype=torch.float64, device=dev)
    vid = torch.as_tensor(np.ascontiguousarray(video_ids, np.int32)).to(dev)
    lab = torch.as_tensor(np.ascontiguousarray(labels, np.int32)).to(dev)
    check(lib.vv_retrieval_stats(_ptr(E), B, max(N, 1), _ptr(vid), _ptr(lab), int(exclude_same_video_shots), _ptr(gram), _ptr(ws),
                                 ws_bytes, _ptr(out), _ptr(pq), _stream()))
    o = out.cpu().numpy()
    return dict(map=float(o[0]), hit1=float(o[1]), hit5=float(o[2]), per_query=pq)


def id_lookup_forward(table, ids):
    """IdToWeightMapping forward: rows of `table` [rows, N] named by the float ids [M]."""
    M = ids.numel(); rows, N = table.shape
    top = torch.empty((M, N), dtype=torch.float32, device=table.device)
    check(_lib.load().vv_id_lookup_forward(_ptr(table), rows, N, _ptr(ids), M, _ptr(top), _stream()))
    return top


def id_lookup_backward(top_diff, ids, rows):
    M, N = top_diff.shape
    d = torch.empty((rows, N), dtype=torch.float32, device=top_diff.device)
    check(_lib.load().vv_id_lookup_backward(_ptr(top_diff), _ptr(ids), M, N, rows, _ptr(d), _stream()))
    return d
