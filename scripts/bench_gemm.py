"""Times the projection GEMMs alone (CUDA events, L2 flushed between launches) at the BASELINE shapes, to separate
main-loop, epilogue and gather effects.  Not a test; prints a table."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from videovector_b200 import ops, _lib
from videovector_b200._lib import DROPOUT_PHILOX, DROPOUT_HASH, DROPOUT_NONE

torch.cuda.set_device(0)
prec = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
M, N, K = 61440, 512, 4096
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.relu(torch.randn(M, K, device="cuda", generator=g))
W = torch.randn(N, K, device="cuda", generator=g) * 0.01
dZ = torch.randn(M, N, device="cuda", generator=g) * 1e-4
b = torch.zeros(N, device="cuda")
Xo, Wo, dZo = ops.prepare_operand(X, prec), ops.prepare_operand(W, prec), ops.prepare_operand(dZ, prec)
H = torch.empty(M, N, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flops = 2.0 * M * N * K


def timeit(name, fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    print("%-44s %.4f ms  %7.1f TFLOP/s (algorithmic)" % (name, ms, flops / ms / 1e9))


def fwd(act):
    ops.check(lib.vv_ip_forward(Xo.c(), Wo.c(), ops._ptr(b), M, N, K, ops.PREC[prec], C.byref(act) if act is not None else None,
                                None, ops._ptr(H), ops._stream()))


timeit("fwd  plain (no activation)", lambda: fwd(None))
timeit("fwd  relu", lambda: fwd(ops.make_act(True, 0.0, DROPOUT_NONE)))
timeit("fwd  relu + philox dropout", lambda: fwd(ops.make_act(True, 0.0, DROPOUT_PHILOX, 0.9, seed=7, step=1)))
timeit("fwd  relu + hash dropout", lambda: fwd(ops.make_act(True, 0.0, DROPOUT_HASH, 0.9, seed=7, step=1)))
nsplit = lib.vv_ip_wgrad_auto_nsplit(M, N, K, ops.PREC[prec])
parts = torch.empty(nsplit, N, K, device="cuda")
timeit("wgrad (nsplit %d)" % nsplit, lambda: ops.check(lib.vv_ip_wgrad(dZo.c(), Xo.c(), M, N, K, ops.PREC[prec], 0.0, ops._ptr(parts), nsplit, None, 0, ops._stream())))
dX = torch.empty(8192, K, device="cuda")
if prec in ("f16x3", "bf16"):
    rows = 65536
    bank = torch.relu(torch.randn(rows, K, device="cuda", generator=g))
    Bo = ops.alloc_operand(bank.shape, prec)        # the bank's operand copy in the gather producers' preferred layout
    ops.check(lib.vv_prepare_bank_operand(ops._ptr(bank), rows, K, ops.PREC[prec], ops._ptr(Bo.hi), ops._ptr(Bo.lo), ops._stream()))
    rowmap = torch.randint(0, rows, (M,), device="cuda", dtype=torch.int32)
    act = ops.make_act(True, 0.0, DROPOUT_HASH, 0.9, seed=7, step=1)
    timeit("fwd  gathered, relu + hash dropout", lambda: ops.check(lib.vv_ip_forward_gathered(
        Bo.c(), rows, ops._ptr(rowmap), None, None, Wo.c(), ops._ptr(b), M, N, K, ops.PREC[prec], C.byref(act), None, ops._ptr(H), ops._stream())))
    timeit("fwd  gathered, plain", lambda: ops.check(lib.vv_ip_forward_gathered(
        Bo.c(), rows, ops._ptr(rowmap), None, None, Wo.c(), ops._ptr(b), M, N, K, ops.PREC[prec], None, None, ops._ptr(H), ops._stream())))
    seq = torch.arange(M, device="cuda", dtype=torch.int32) % rows
    timeit("fwd  gathered (sequential rows), plain", lambda: ops.check(lib.vv_ip_forward_gathered(
        Bo.c(), rows, ops._ptr(seq), None, None, Wo.c(), ops._ptr(b), M, N, K, ops.PREC[prec], None, None, ops._ptr(H), ops._stream())))
    timeit("wgrad gathered (transposed), nsplit %d" % nsplit, lambda: ops.check(lib.vv_ip_wgrad_gathered(
        dZo.c(), Bo.c(), rows, ops._ptr(rowmap), M, N, K, ops.PREC[prec], 0.0, ops._ptr(parts), nsplit, ops._stream())))
