"""1 000 iterations of the REFERENCE's own training pipeline (oracle/_ref: its data layer, Net and SGDSolver sources compiled
unmodified, Caffe CPU mode) at a well-conditioned shape -> tests/golden/curve_ref.npz: per-iteration loss and violation
count.  The dataset, W0 and b0 are regenerated from seeds by the tests (ops.bank_host is bit-identical to vv_fill_bank),
so the fixture holds only the curve.  No dropout layer (ratio 0): the curve is then a deterministic function of the
sampler stream, which the product reproduces bit for bit.
    python tests/golden/make_curve_golden.py        (needs oracle/_ref, i.e. /root/reference at build time)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import pyoracle as orc, pyref
from videovector_b200 import ops

CURVE = dict(B=128, C=5, Nn=10, K=1024, N=256, V=256, S=16, P=1000, swap=50, max_same=6, steps=1000,
             base_lr=0.01, gamma=1e-3, power=0.75, momentum=0.9, weight_decay=5e-4, bank_seed=1234, w_seed=1701, w_std=0.1, b_std=0.01)


def problem(c=CURVE):
    vid, off, sid = ops.synthetic_videos(c["V"], c["S"])
    feat = ops.bank_host(c["V"] * c["S"], c["K"], c["bank_seed"])
    rng = np.random.RandomState(c["w_seed"])
    W0 = rng.normal(0, c["w_std"], (c["N"], c["K"])).astype(np.float32)
    b0 = rng.normal(0, c["b_std"], (c["N"],)).astype(np.float32)
    return vid, off, sid, feat, W0, b0


if __name__ == "__main__":
    c = CURVE
    vid, off, sid, feat, W0, b0 = problem()
    orc.use_openblas(0)
    sol = pyref.Solver(vid, off, sid, feat, W0, b0, c["B"], c["C"], c["Nn"], c["P"], c["swap"], c["max_same"], margin=2.0, norm=2,
                       base_lr=c["base_lr"], momentum=c["momentum"], weight_decay=c["weight_decay"], lr_policy="inv", gamma=c["gamma"],
                       power=c["power"], dropout_ratio=0.0)
    t0 = time.time()
    loss = np.zeros(c["steps"], np.float32); viol = np.zeros(c["steps"], np.float32)
    for it in range(c["steps"]):
        loss[it], viol[it] = sol.step()
    st = sol.state()
    sol.close()
    print("reference: %d steps in %.1f s; loss %.5f -> %.5f; violations %d -> %d" % (c["steps"], time.time() - t0, loss[:20].mean(), loss[-20:].mean(), viol[0], viol[-1]))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "curve_ref.npz"), loss=loss, viol=viol,
                        W_final_rowsum=st["W"].astype(np.float64).sum(1).astype(np.float32), b_final=st["b"],
                        cfg=np.array([repr(sorted(c.items()))]))
    # how well conditioned is it?  the oracle restatement (built-in GEMM loops, another summation order) on the same stream
    orc.use_builtin_blas()
    smp = orc.Sampler(vid, off, sid, feat, c["K"], c["B"], c["C"], c["Nn"], c["P"], c["swap"], c["max_same"], 100, seed=1)
    W, b = W0.copy(), b0.copy(); hW = np.zeros_like(W); hb = np.zeros_like(b)
    ol = np.zeros(c["steps"], np.float32)
    orc.use_openblas(0)
    for it in range(c["steps"]):
        idx, quirk, data = smp.next()
        out = orc.net_forward_backward(data, W, b, None, c["B"], c["C"], c["Nn"], margin=2.0, norm=2, dropout_ratio=0.0,
                                       want=("loss", "dW", "db"))
        ol[it] = out["loss"][0]
        rate = orc.learning_rate("inv", c["base_lr"], c["gamma"], c["power"], 1, it)
        W, _, hW = orc.sgd_update(W, out["dW"], hW, rate, c["momentum"], c["weight_decay"])
        b, _, hb = orc.sgd_update(b, out["db"], hb, rate * 2, c["momentum"], 0.0)
    d = np.abs(ol - loss) / np.abs(loss).max()
    print("oracle vs reference, step by step: max rel %.2e (first 100: %.2e, last 100: %.2e)" % (d.max(), d[:100].max(), d[-100:].max()))
