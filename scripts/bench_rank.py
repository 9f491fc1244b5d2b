"""Times vv_rank_loss_fused alone at the trainer's mode (B=4096, C=5, Nn=10, N=512, f16x3 operand output only) for the
kernel variants selected by VV_RANK_RING / VV_RANK_RING_PER_SM, and checks the variants against each other."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from videovector_b200 import ops, _lib
from videovector_b200._lib import PREC, check

B, Cc, Nn, N = int(os.environ.get("B", 4096)), 5, 10, int(os.environ.get("N", 512))
R = Cc + Nn
torch.manual_seed(0)
H = torch.relu(torch.randn(R * B, N, device="cuda")) * (torch.rand(R * B, N, device="cuda") > 0.9).float() * 10
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
L = _lib.load()
cfg = ops.rank_cfg(B, Cc, Nn, N)
p = PREC["f16x3"]


def call(op, out, db):
    check(L.vv_rank_loss_fused(ops._ptr(H), C.byref(cfg), 1.0, 1, 10.0, ops._ptr(out["stats"]), None, None,
                               ops._ptr(out["item_loss"]), ops._ptr(out["item_viol"]), ops._ptr(out["loss"]), ops._ptr(out["viol"]),
                               None, ops._ptr(op.hi), ops._ptr(op.lo), p, ops._ptr(db), None, None, ops._stream()))


def run(tag, env):
    for k in ("VV_RANK_RING", "VV_RANK_RING_PER_SM"):
        os.environ.pop(k, None)
    os.environ.update(env)
    out = dict(stats=torch.zeros((B, 1 + 2 * (1 + Nn)), device="cuda"), item_loss=torch.zeros(B, device="cuda"),
               item_viol=torch.zeros(B, device="cuda"), loss=torch.zeros(1, device="cuda"), viol=torch.zeros(1, device="cuda"))
    op = ops.alloc_operand(H.shape, p, "cuda")
    db = torch.zeros(N, device="cuda")
    call(op, out, db); ops.operand_rescale(op, p); db.zero_(); call(op, out, db)
    torch.cuda.synchronize()
    res = (out["stats"].clone(), out["loss"].clone(), op.dequant().clone(), db.clone())
    ts = []
    for _ in range(14):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); call(op, out, db); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[2:]))
    gb = (R * B * N * 4 * 2) / 1e9
    return ms, gb / (ms * 1e-3), res


if __name__ == "__main__":
    if len(sys.argv) > 2:                       # one variant only (for ncu): bench_rank.py <stages> <per_sm>
        ms, bw, _ = run("ring", {"VV_RANK_RING": sys.argv[1], "VV_RANK_RING_PER_SM": sys.argv[2]})
        print("stages=%s per_sm=%s %.4f ms %.0f GB/s" % (sys.argv[1], sys.argv[2], ms, bw)); sys.exit(0)
    ms, bw, ref = run("reg", {"VV_RANK_RING": "0"})
    print("%-28s %.4f ms  %.0f GB/s" % ("register-resident", ms, bw), flush=True)
    for st in (2, 3, 4, 5, 6):
        for per in (1, 2, 3, 4, 6):
            try:
                ms, bw, r = run("ring", {"VV_RANK_RING": str(st), "VV_RANK_RING_PER_SM": str(per)})
            except Exception as e:
                print("ring", st, per, "failed", e, flush=True); continue
            same = [bool(torch.equal(a, b)) for a, b in zip(r[:3], ref[:3])]
            dberr = float((r[3] - ref[3]).abs().max() / ref[3].abs().max())
            print("%-28s %.4f ms  %.0f GB/s  bit-identical stats/loss/dZ %s db rel %.1e" % ("ring stages=%d per_sm=%d" % (st, per), ms, bw, same, dberr), flush=True)
