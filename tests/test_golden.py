"""Golden vectors produced by the REFERENCE's own layer code (tests/golden/make_golden.py, via oracle/_ref):
the oracle must reproduce them; where oracle/_ref is present (build container) the reference is also run live."""
import glob
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NETS = sorted(glob.glob(os.path.join(GOLD, "net_*.npz")))


def rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def test_fixtures_present():
    assert len(NETS) >= 3 and os.path.exists(os.path.join(GOLD, "layers.npz")) and os.path.exists(os.path.join(GOLD, "make_golden.py"))


@pytest.mark.parametrize("path", NETS, ids=[os.path.basename(p)[:-4] for p in NETS])
def test_oracle_reproduces_reference_net(oracle, path):
    g = np.load(path)
    B, C, Nn = int(g["B"]), int(g["C"]), int(g["Nn"])
    ratio = float(g["dropout_ratio"])
    mask = g["mask"].astype(np.uint32) if ratio > 0 else None
    for blas in ("builtin", "openblas"):
        if blas == "openblas":
            oracle.use_openblas(0)
        o = oracle.net_forward_backward(g["data"], g["W"], g["b"], mask, B, C, Nn, margin=float(g["margin"]), norm=int(g["norm"]),
                                        dropout_ratio=ratio if ratio > 0 else 0.5,
                                        want=("loss", "violations", "dW", "db", "H", "dZ", "target_score", "neg_score"))
        oracle.use_builtin_blas()
        assert abs(o["loss"][0] - g["loss"][0]) <= 2e-6 * max(1, abs(g["loss"][0]))
        assert o["violations"][0] == g["violations"][0]
        for k in ("H", "target_score", "neg_score", "dZ", "dW", "db"):
            assert rel(o[k], g[k]) < 2e-6, (k, rel(o[k], g[k]))


def test_oracle_reproduces_reference_layers(oracle):
    g = np.load(os.path.join(GOLD, "layers.npz"))
    assert rel(oracle.normalization_forward(g["norm_x"]), g["norm_y"]) < 1e-6
    assert rel(oracle.normalization_backward(g["norm_x"], g["norm_dy"]), g["norm_dx"]) < 2e-6
    assert np.array_equal(oracle.normalization_forward(g["norm_x"])[2], np.zeros(40, np.float32))      # zero row -> zeros
    for norm in (1, 2):
        loss, viol, _ = oracle.max_margin_forward(g["mm_t"], g["mm_s"], margin=1.0, norm=norm)
        dt, dbg = oracle.max_margin_backward(g["mm_t"], g["mm_s"], margin=1.0, norm=norm, loss_weight=1.0)
        assert abs(loss - float(g["loss%d" % norm])) < 1e-6 * max(1, loss) and viol == float(g["viol%d" % norm])
        assert rel(dt, g["dt%d" % norm]) < 1e-6 and rel(dbg, g["db%d" % norm]) < 1e-6
    Z = oracle.ip_forward(g["ip_X"], g["ip_W"], g["ip_b"])
    dW, db, dX = oracle.ip_backward(g["ip_dZ"], g["ip_X"], g["ip_W"], regularization=0.5, want_dx=True)
    assert rel(Z, g["ip_Z"]) < 1e-6 and rel(dW, g["ip_dW"]) < 1e-6 and rel(db, g["ip_db"]) < 1e-6 and rel(dX, g["ip_dX"]) < 1e-6


def mm_table_weights(g):
    """id -> weight as the reference's map holds it: first line of an id wins, unknown ids weigh 0."""
    table = {}
    for i, w in zip(g["table_ids"], g["table_w"]):
        table.setdefault(int(i), float(w))
    return np.vectorize(lambda i: table.get(int(i), 0.0))(g["ids"]).astype(np.float32)


def test_oracle_reproduces_reference_weighted_max_margin(oracle):
    """MaxMarginLoss with its third bottom (per-video weights), both forms, against the compiled reference layer
    (tests/golden/mm_weights.npz <- make_mm_weights_golden.py; ref: max_margin_loss_layer.cpp:18-39,79-97,150-186)."""
    g = np.load(os.path.join(GOLD, "mm_weights.npz"))
    margin, lw = float(g["margin"]), float(g["loss_weight"])
    for form, w in (("direct", g["w"]), ("table", mm_table_weights(g))):
        for norm in (1, 2):
            loss, viol, _ = oracle.max_margin_forward(g["t"], g["s"], margin=margin, norm=norm, weights=w)
            dt, dbg = oracle.max_margin_backward(g["t"], g["s"], margin=margin, norm=norm, loss_weight=lw, weights=w)
            assert abs(loss - float(g["%s_loss%d" % (form, norm)])) < 1e-6 * max(1, loss) and viol == float(g["%s_viol%d" % (form, norm)])
            assert rel(dt, g["%s_dt%d" % (form, norm)]) < 1e-6 and rel(dbg, g["%s_db%d" % (form, norm)]) < 1e-6


def test_live_reference_agrees_with_fixtures_and_oracle(oracle):
    """Only where oracle/_ref was built (needs /root/reference at build time)."""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref not built on this machine")
    g = np.load(NETS[0])
    r = pyref.net_forward_backward(g["data"], g["W"], g["b"], int(g["B"]), int(g["C"]), int(g["Nn"]), margin=2.0, norm=int(g["norm"]),
                                   dropout_ratio=float(g["dropout_ratio"]), seed=1701)
    assert np.array_equal(r["mask"].astype(np.uint8), g["mask"]) and rel(r["dW"], g["dW"]) < 1e-6   # fixtures are reproducible
    # fresh random cases at the shipped widths: reference vs oracle
    rng = np.random.RandomState(11)
    for (B, C, Nn, K, N) in [(16, 5, 10, 4096, 512), (4, 17, 50, 512, 1024)]:
        data = np.maximum(rng.normal(0, 1, (B, C + Nn, K)), 0).astype(np.float32)
        W = rng.normal(0, 0.01, (N, K)).astype(np.float32); b = rng.normal(0, 0.01, N).astype(np.float32)
        r = pyref.net_forward_backward(data, W, b, B, C, Nn, dropout_ratio=0.9, seed=3)
        oracle.use_openblas(0)
        o = oracle.net_forward_backward(data, W, b, r["mask"], B, C, Nn, dropout_ratio=0.9, want=("loss", "violations", "dW", "db", "dZ"))
        oracle.use_builtin_blas()
        assert abs(o["loss"][0] - r["loss"][0]) < 2e-6 * max(1, r["loss"][0]) and o["violations"][0] == r["violations"][0]
        assert rel(o["dW"], r["dW"]) < 5e-6 and rel(o["db"], r["db"]) < 5e-6 and rel(o["dZ"], r["dZ"]) < 5e-6


def test_oracle_reproduces_reference_eval_layers(oracle):
    """RetrievalStatsLayer / IdToWeightMappingLayer outputs of the compiled reference (tests/golden/eval_layers.npz)."""
    g = np.load(os.path.join(GOLD, "eval_layers.npz"))
    idmap = dict(zip(g["rs_map_keys"].tolist(), g["rs_map_vals"].tolist()))
    vids = g["rs_vids"].astype(np.int32)
    labels = np.array([idmap[int(v)] for v in vids], np.int32)
    assert (labels < 0).any()
    for excl in (0, 1):
        o = oracle.retrieval_stats(g["rs_E"], vids, labels, bool(excl))
        ref = g["rs_out_%d" % excl]
        assert abs(o["map"] - ref[0]) < 1e-6 and abs(o["hit1"] - ref[1]) < 1e-6 and abs(o["hit5"] - ref[2]) < 1e-6, (excl, o, ref)
    assert np.array_equal(oracle.id_lookup_forward(g["id_table"], g["id_ids"]), g["id_top"])
    assert np.array_equal(oracle.id_lookup_backward(g["id_tdiff"], g["id_ids"], g["id_table"].shape[0]), g["id_tgrad"])


def test_oracle_reproduces_reference_solver_trajectory(oracle):
    """tests/golden/solver_ref.npz: 8 iterations of the REFERENCE's whole pipeline -- its data layer, Net (net.cpp,
    insert_splits.cpp) and SGDSolver (solver.cpp), compiled unmodified in oracle/_ref (make_golden.py).  The oracle's
    sampler + net + solver-update restatements must follow the same trajectory; where oracle/_ref exists the reference is
    also run live with another policy / norm."""
    g = np.load(os.path.join(GOLD, "solver_ref.npz"))
    B, C, Nn, P, swap, max_same = [int(x) for x in g["cfg"]]
    base_lr, mom, wd, gamma, power = [float(x) for x in g["hyper"]]
    K = g["feat"].shape[1]

    def run(vid, off, sid, feat, W0, b0, steps, policy, norm, stepsize=1, reg_type=2, hyper=None):
        lr0, mo, dec, gam, pw = hyper if hyper is not None else (base_lr, mom, wd, gamma, power)
        smp = oracle.Sampler(vid, off, sid, feat, K, B, C, Nn, P, swap, max_same, 100, seed=1)
        W, b = W0.copy(), b0.copy(); hW = np.zeros_like(W); hb = np.zeros_like(b)
        out = []
        for it in range(steps):
            data = smp.next()[2]
            r = oracle.net_forward_backward(data, W, b, None, B, C, Nn, margin=2.0, norm=norm, dropout_ratio=0.0)
            rate = oracle.learning_rate(policy, lr0, gam, pw, stepsize, it)
            W, _, hW = oracle.sgd_update(W, r["dW"], hW, rate * 1.0, mo, dec * 1.0, reg_type)      # blobs_lr 1 / 2, weight_decay 1 / 0
            b, _, hb = oracle.sgd_update(b, r["db"], hb, rate * 2.0, mo, 0.0, reg_type)
            out.append((float(r["loss"][0]), float(r["violations"][0])))
        smp.close()
        return out, dict(W=W, b=b, hW=hW, hb=hb)

    traj, st = run(g["vid"], g["off"], g["sid"], g["feat"], g["W0"], g["b0"], len(g["loss"]), "inv", 2)
    for it, (loss, viol) in enumerate(traj):
        assert abs(loss - g["loss"][it]) < 1e-5 * max(1, abs(g["loss"][it])) and viol == g["violations"][it], it
    for k in ("W", "b", "hW", "hb"):
        assert rel(st[k], g[k]) < 1e-5, k
    # the second fixture trajectory: L1 hinge, "step" policy, L1 weight regularisation
    a_lr, a_mom, a_wd, a_gamma, a_step = [float(x) for x in g["alt_hyper"]]
    traj, sa = run(g["vid"], g["off"], g["sid"], g["feat"], g["W0"], g["b0"], len(g["alt_loss"]), "step", 1, stepsize=int(a_step),
                   reg_type=1, hyper=(a_lr, a_mom, a_wd, a_gamma, 0.0))
    for it, (loss, viol) in enumerate(traj):
        assert abs(loss - g["alt_loss"][it]) < 1e-5 * max(1, abs(g["alt_loss"][it])) and viol == g["alt_violations"][it], it
    for k in ("W", "b", "hW", "hb"):
        assert rel(sa[k], g["alt_" + k]) < 1e-5, k
    # Solver::Test's loop on the reference's TEST net (shared weights), 2 iterations before and after the 8 steps
    tdata, tvid, TB = g["test_data"], g["test_vid"], int(g["test_batch"])
    cls = dict(zip(g["id_keys"].tolist(), g["id_vals"].tolist()))

    def test_scores(W, b, start):
        acc = np.zeros(3)
        for i in range(2):
            item = (np.arange(TB) + (start + i) * TB) % len(tdata)
            E = oracle.test_embed(tdata[item], W, b)[1]
            labels = np.array([cls.get(int(v), 0) for v in tvid[item]], np.int32)
            r = oracle.retrieval_stats(E, tvid[item], labels, True)
            acc += [r["map"], r["hit1"], r["hit5"]]
        return acc / 2
    assert np.abs(test_scores(g["W0"], g["b0"], 0) - g["test_before"]).max() < 1e-6
    assert np.abs(test_scores(st["W"], st["b"], 2) - g["test_after"]).max() < 1e-5
    from oracle import pyref
    if pyref.available():
        gamma, power = 0.5, 0.0          # "step": rate = base_lr * gamma^(iter / stepsize)
        ref = pyref.Solver(g["vid"], g["off"], g["sid"], g["feat"], g["W0"], g["b0"], B, C, Nn, P, swap, max_same, norm=1,
                           base_lr=base_lr, momentum=mom, weight_decay=wd, lr_policy="step", gamma=gamma, power=power, stepsize=2)
        rt = [ref.step() for _ in range(5)]
        rs = ref.state(); ref.close()
        traj, st = run(g["vid"], g["off"], g["sid"], g["feat"], g["W0"], g["b0"], 5, "step", 1, stepsize=2)
        for it in range(5):
            assert abs(traj[it][0] - rt[it][0]) < 1e-5 * max(1, abs(rt[it][0])) and traj[it][1] == rt[it][1], it
        for k in ("W", "b", "hW", "hb"):
            assert rel(st[k], rs[k]) < 1e-5, k


def test_oracle_follows_reference_loss_curve(oracle):
    """tests/golden/curve_ref.npz = 1 000 iterations of the compiled reference pipeline (make_curve_golden.py) at a
    well-conditioned shape (no dropout layer: the curve is a function of the bit-exact sampler stream alone).  The oracle's
    sampler + net + update restatements follow it step by step (first 150 iterations here; the generator checks all 1 000)."""
    import sys
    sys.path.insert(0, GOLD)
    from make_curve_golden import CURVE as c, problem
    g = np.load(os.path.join(GOLD, "curve_ref.npz"))
    assert len(g["loss"]) == c["steps"] == 1000
    vid, off, sid, feat, W0, b0 = problem()
    smp = oracle.Sampler(vid, off, sid, feat, c["K"], c["B"], c["C"], c["Nn"], c["P"], c["swap"], c["max_same"], 100, seed=1)
    W, b = W0.copy(), b0.copy(); hW = np.zeros_like(W); hb = np.zeros_like(b)
    oracle.use_openblas(0)
    n = 150
    for it in range(n):
        idx, quirk, data = smp.next()
        out = oracle.net_forward_backward(data, W, b, None, c["B"], c["C"], c["Nn"], margin=2.0, norm=2, dropout_ratio=0.0,
                                          want=("loss", "violations", "dW", "db"))
        assert abs(out["loss"][0] - g["loss"][it]) <= 1e-5 * g["loss"][it], it
        assert out["violations"][0] == g["viol"][it], it
        rate = oracle.learning_rate("inv", c["base_lr"], c["gamma"], c["power"], 1, it)
        W, _, hW = oracle.sgd_update(W, out["dW"], hW, rate, c["momentum"], c["weight_decay"])
        b, _, hb = oracle.sgd_update(b, out["db"], hb, rate * 2, c["momentum"], 0.0)
    oracle.use_builtin_blas()
    smp.close()
