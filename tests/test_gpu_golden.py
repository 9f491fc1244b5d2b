"""The CUDA path (through the C-ABI) against golden vectors produced by the reference's own layer code."""
import glob
import os

import numpy as np
import pytest
import torch

from videovector_b200 import ops
from videovector_b200._lib import DROPOUT_MASK01

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NETS = sorted(glob.glob(os.path.join(GOLD, "net_*.npz")))


def rel(a, b):
    a = torch.as_tensor(a).double().cpu().numpy(); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("prec,tol", [("fp32_simt", 1e-5), ("tf32x3", 1e-5), ("f16x3", 1e-5)])
@pytest.mark.parametrize("path", NETS, ids=[os.path.basename(p)[:-4] for p in NETS])
def test_fused_step_reproduces_reference(path, prec, tol):
    g = np.load(path)
    B, C, Nn, K, N = (int(g[k]) for k in ("B", "C", "Nn", "K", "N"))
    R = C + Nn
    ratio = float(g["dropout_ratio"])
    bank = torch.as_tensor(g["data"].reshape(B * R, K)).cuda()                 # every slot is its own bank row
    idx = torch.arange(B * R, dtype=torch.int32).reshape(B, R).cuda()
    quirk = torch.full((B, R), -2, dtype=torch.int32).cuda()
    cfg = ops.trainer_cfg(B, C, Nn, K, N, margin=float(g["margin"]), norm=int(g["norm"]), dropout_ratio=ratio,
                          dropout_mode=DROPOUT_MASK01, prec=prec, keep_blobs=True)
    tr = ops.Trainer(cfg)
    tr.set_weights(torch.as_tensor(g["W"]).cuda(), torch.as_tensor(g["b"]).cuda())
    mask = torch.as_tensor(g["mask"].astype(np.int32)).cuda() if ratio > 0 else None
    tr.step(bank, idx, quirk, mask, it=0, do_update=False)
    assert abs(tr.tensor("loss").item() - g["loss"][0]) <= tol * max(1, abs(g["loss"][0]))
    assert tr.tensor("violations").item() == g["violations"][0]
    assert rel(tr.tensor("H"), g["H"]) < tol
    assert rel(tr.tensor("dZ"), g["dZ"]) < 2 * tol
    assert rel(tr.tensor("dW_raw"), g["dW"]) < 2 * tol
    assert rel(tr.tensor("db_raw"), g["db"]) < 2 * tol
    tr.close()


def test_layer_kernels_reproduce_reference(vvlib):
    from videovector_b200.ops import _ptr, _stream
    from videovector_b200._lib import check
    g = np.load(os.path.join(GOLD, "layers.npz"))
    x = torch.as_tensor(g["norm_x"]).cuda(); dy = torch.as_tensor(g["norm_dy"]).cuda(); y = torch.empty_like(x)
    check(vvlib.vv_l2norm_forward(_ptr(x), x.shape[0], x.shape[1], _ptr(y), _stream()))
    assert rel(y, g["norm_y"]) < 1e-6
    check(vvlib.vv_l2norm_backward(_ptr(x), _ptr(dy), x.shape[0], x.shape[1], _ptr(y), _stream()))
    assert rel(y, g["norm_dx"]) < 1e-5
    for prec, tol in (("fp32_simt", 1e-6), ("tf32x3", 1e-5), ("f16x3", 1e-5)):
        X = torch.as_tensor(np.pad(g["ip_X"], ((0, 0), (0, 4)))).cuda().contiguous()     # K 60 -> 64 (tensor-core path needs K % 8)
        W = torch.as_tensor(np.pad(g["ip_W"], ((0, 6), (0, 4)))).cuda().contiguous()     # N 10 -> 16
        b = torch.as_tensor(np.pad(g["ip_b"], (0, 6))).cuda()
        H, _ = ops.ip_forward(ops.prepare_operand(X, prec), ops.prepare_operand(W, prec), b, 7, 16, 64, prec)
        assert rel(H[:, :10], g["ip_Z"]) < tol


def test_weighted_max_margin_reproduces_reference(vvlib, tmp_path):
    """MaxMarginLoss with per-video weights (third bottom): the C-ABI kernels and the reference-interface layer
    (MaxMarginLossLayer<float> of caffe_compat, built from a `layers { }` entry and driven like the reference's per-layer
    tests) against the compiled reference layer's outputs -- tests/golden/mm_weights.npz, both forms: direct weights and
    video ids through an id_to_weight_file (ref: max_margin_loss_layer.cpp:18-39,79-97,150-186)."""
    from videovector_b200 import caffe_host
    from videovector_b200.ops import _ptr, _stream
    from videovector_b200._lib import check
    g = np.load(os.path.join(GOLD, "mm_weights.npz"))
    margin, lw = float(g["margin"]), float(g["loss_weight"])
    t = torch.as_tensor(g["t"]).cuda(); s = torch.as_tensor(g["s"]).cuda()
    count = t.numel()
    # the table as the layer uploads it: ascending ids, first line of an id wins
    table = {}
    for i, w in zip(g["table_ids"], g["table_w"]):
        table.setdefault(int(i), float(w))
    tid = torch.as_tensor(np.array(sorted(table), np.int32)).cuda()
    tw = torch.as_tensor(np.array([table[i] for i in sorted(table)], np.float32)).cuda()
    w_table = torch.empty_like(t)
    check(vvlib.vv_id_to_weight(_ptr(torch.as_tensor(g["ids"]).cuda()), count, _ptr(tid), _ptr(tw), tid.numel(), _ptr(w_table), _stream()))
    assert (w_table[3] == 0).all()                                     # id 1000 is not in the table
    path = tmp_path / "id2w.txt"
    path.write_text("".join("%d,%.9g\n" % (int(i), float(w)) for i, w in zip(g["table_ids"], g["table_w"])))
    caffe_host.set_device(0)
    for form, w, third, extra in (("direct", torch.as_tensor(g["w"]).cuda(), g["w"], "use_direct_weight: true"),
                                  ("table", w_table, g["ids"], 'id_to_weight_file: "%s"' % path)):
        for norm in (1, 2):
            hinge = torch.empty_like(t); loss = torch.zeros(1, device="cuda"); viol = torch.zeros(1, device="cuda")
            dt = torch.empty_like(t); dbg = torch.empty_like(t)
            check(vvlib.vv_max_margin_forward_w(_ptr(t), _ptr(s), _ptr(w), count, margin, norm, _ptr(hinge), _ptr(loss), _ptr(viol), _stream()))
            check(vvlib.vv_max_margin_backward_w(_ptr(t), _ptr(s), _ptr(w), count, margin, norm, lw, _ptr(dt), _ptr(dbg), _stream()))
            want = float(g["%s_loss%d" % (form, norm)])
            assert abs(loss.item() - want) < 2e-6 * max(1, want) and viol.item() == float(g["%s_viol%d" % (form, norm)])
            assert rel(dt, g["%s_dt%d" % (form, norm)]) < 1e-6 and rel(dbg, g["%s_db%d" % (form, norm)]) < 1e-6
            # the layer, through the reference's interface
            text = ('layers { name: "loss" type: MAX_MARGIN_LOSS bottom: "t" bottom: "s" bottom: "v" top: "l" top: "nv" '
                    'loss_weight: %g loss_weight: 0 max_margin_loss_param { norm: %s margin: %g %s } }' % (lw, "L2" if norm == 2 else "L1", margin, extra))
            lval, tops, diffs = caffe_host.run_layer(text, [g["t"], g["s"], third], 2, propagate_down=[True, True, False])
            assert abs(tops[0][0] - want) < 2e-6 * max(1, want) and tops[1][0] == float(g["%s_viol%d" % (form, norm)])
            assert abs(lval - lw * want) < 2e-6 * max(1, want)
            assert rel(diffs[0], g["%s_dt%d" % (form, norm)]) < 1e-6 and rel(diffs[1], g["%s_db%d" % (form, norm)]) < 1e-6 and diffs[2] is None


def _csv_rows(text):
    rows = []
    for line in text.strip().splitlines():
        if line.startswith("#"):
            rows.append(line)
        else:
            f = line.split(",")
            rows.append((int(f[0]), int(f[1]), float(f[2]), float(f[3]), float(f[4])) + tuple(int(x) for x in f[5:]))
    return rows


def _csv_equal(a, b):
    if len(a) != len(b):
        return False
    for x, y in zip(a, b):
        if isinstance(x, str) or isinstance(y, str):
            if x != y:
                return False
        elif x[:2] != y[:2] or x[5:] != y[5:] or any(abs(p - q) > 2e-6 for p, q in zip(x[2:5], y[2:5])):
            return False
    return True


@pytest.mark.parametrize("excl", [0, 1])
def test_retrieval_stats_video_level_and_csv_reproduce_reference(vvlib, tmp_path, excl):
    """RetrievalStatsLayer's `video_level_retrieval` and `stats_output_file` (retrieval_stats_layer.cpp:146-206, 306-340)
    through the reference's interface (caffe_compat's layer, driven alone) against the compiled reference layer's outputs
    and CSV files (tests/golden/retrieval_opts.npz <- make_retrieval_opts_golden.py).  Video-level CSV lines come in the
    reference's hash-map order there and in ascending video id here: compared as sets."""
    from videovector_b200 import caffe_host
    from videovector_b200.ops import _ptr, _stream
    from videovector_b200._lib import check
    g = np.load(os.path.join(GOLD, "retrieval_opts.npz"))
    E, vids = g["E"], g["vids"]
    B, N = E.shape
    idf = tmp_path / "id2class.txt"
    idf.write_text("".join("%d,%d\n" % (int(a), int(b)) for a, b in zip(g["map_ids"], g["map_cls"])))
    # the mean embeddings themselves
    uniq = sorted(set(int(v) for v in vids))
    group = torch.as_tensor(np.array([uniq.index(int(v)) for v in vids], np.int32)).cuda()
    Ed = torch.as_tensor(E).cuda()
    mean = torch.empty((len(uniq), N), device="cuda")
    check(vvlib.vv_video_mean_rows(_ptr(Ed), B, N, _ptr(group), len(uniq), _ptr(mean), _stream()))
    want = np.stack([E[vids == v].astype(np.float64).mean(0) for v in uniq])
    assert rel(mean, want) < 1e-6
    caffe_host.set_device(0)
    for level, key in ((False, "shot"), (True, "video")):
        csv = tmp_path / ("%s_%d.csv" % (key, excl))
        text = ('layers { name: "retrieval_stats" type: RETRIEVAL_STATS bottom: "e" bottom: "ids" top: "map" top: "hit1" top: "hit5" '
                'retrieval_stats_param { id_to_class_file: "%s" stats_output_file: "%s" exclude_same_video_shots: %s%s } }'
                % (idf, csv, "true" if excl else "false",
                   (" video_level_retrieval: true max_num_videos: %d" % int(g["max_num_videos"])) if level else ""))
        _, tops, _ = caffe_host.run_layer(text, [E.reshape(B, N, 1, 1), vids.reshape(B, 1, 1, 1)], 3)
        got = np.array([t[0] for t in tops])
        assert np.abs(got - g["%s_out_%d" % (key, excl)]).max() < 2e-6, (key, got, g["%s_out_%d" % (key, excl)])
        ours = _csv_rows(csv.read_text())
        ref = _csv_rows(bytes(g["%s_csv_%d" % (key, excl)]).decode())
        if level:
            ours = [ours[0]] + sorted(ours[1:]); ref = [ref[0]] + sorted(ref[1:])
        assert _csv_equal(ours, ref), (key, ours[:3], ref[:3])


def test_device_eval_kernels_reproduce_reference_fixtures():
    """vv_retrieval_stats / vv_id_lookup_* against the compiled reference's outputs (tests/golden/eval_layers.npz)."""
    import torch
    from videovector_b200 import ops
    g = np.load(os.path.join(GOLD, "eval_layers.npz"))
    idmap = dict(zip(g["rs_map_keys"].tolist(), g["rs_map_vals"].tolist()))
    vids = g["rs_vids"].astype(np.int32)
    labels = np.array([idmap[int(v)] for v in vids], np.int32)
    for excl in (0, 1):
        o = ops.retrieval_stats(torch.as_tensor(g["rs_E"]).cuda(), vids, labels, bool(excl))
        ref = g["rs_out_%d" % excl]
        assert abs(o["map"] - ref[0]) < 1e-5 and abs(o["hit1"] - ref[1]) < 1e-6 and abs(o["hit5"] - ref[2]) < 1e-6, (excl, o, ref)
    top = ops.id_lookup_forward(torch.as_tensor(g["id_table"]).cuda(), torch.as_tensor(g["id_ids"]).cuda())
    assert np.array_equal(top.cpu().numpy(), g["id_top"])
    d = ops.id_lookup_backward(torch.as_tensor(g["id_tdiff"]).cuda(), torch.as_tensor(g["id_ids"]).cuda(), g["id_table"].shape[0])
    assert np.array_equal(d.cpu().numpy(), g["id_tgrad"])


def _solver_fixture():
    g = np.load(os.path.join(GOLD, "solver_ref.npz"))
    B, C, Nn, P, swap, max_same = [int(x) for x in g["cfg"]]
    base_lr, mom, wd, gamma, power = [float(x) for x in g["hyper"]]
    return g, (B, C, Nn, P, swap, max_same), dict(base_lr=base_lr, momentum=mom, weight_decay=wd, gamma=gamma, power=power, lr_policy="inv")


@pytest.mark.parametrize("prec,fused_gather", [("fp32_simt", False), ("tf32x3", False), ("f16x3", False), ("f16x3", True)])
def test_trainer_follows_reference_solver_trajectory(prec, fused_gather):
    """tests/golden/solver_ref.npz = the REFERENCE's whole pipeline (its data layer + Net + SGDSolver, compiled unmodified)
    over 8 iterations.  The product -- host sampler + fused trainer step on the device -- must follow it: loss (1e-5) and
    violation count (exact) every iteration, weights, bias and both momentum histories at the end (1e-5)."""
    g, (B, C, Nn, P, swap, max_same), hyper = _solver_fixture()
    N, K = g["W0"].shape
    bank = torch.as_tensor(g["feat"]).cuda()
    smp = ops.Sampler(g["vid"], g["off"], g["sid"], B, C, Nn, P, swap, max_same, 100, rand_seed=1)
    tr = ops.Trainer(ops.trainer_cfg(B, C, Nn, K, N, dropout_ratio=0.0, prec=prec, **hyper))
    tr.set_weights(torch.as_tensor(g["W0"]).cuda(), torch.as_tensor(g["b0"]).cuda())
    if fused_gather:
        tr.set_bank(bank)
    for it in range(len(g["loss"])):
        idx, quirk = smp.next()
        tr.step(bank, torch.as_tensor(idx).cuda(), torch.as_tensor(quirk).cuda(), None, it=it)
        assert abs(tr.tensor("loss").item() - g["loss"][it]) < 1e-5 * max(1, abs(g["loss"][it])), it
        assert tr.tensor("violations").item() == g["violations"][it], it
    for name, key in (("W", "W"), ("b", "b"), ("W_hist", "hW"), ("b_hist", "hb")):
        assert rel(tr.tensor(name), g[key]) < 1e-5, name
    tr.close(); smp.close()


@pytest.mark.parametrize("fuse", [False, True])
def test_caffe_host_solver_follows_reference_solver_trajectory(tmp_path, monkeypatch, fuse):
    """The same fixture through the drop-in boundary end to end: VideoShots records (protobuf wire bytes) -> the compat
    data layer -> Net built from the prototxt -> SGDSolver, layer by layer and fused.  The net must also contain the
    layers (names, order, the inserted split) the reference's Net::Init produced."""
    import records_util
    from videovector_b200 import caffe_host, prototxt
    g, (B, C, Nn, P, swap, max_same), hyper = _solver_fixture()
    N, K = g["W0"].shape
    src = records_util.write_vvrs(tmp_path / "train.vvrs", records_util.video_shots_records(g["vid"], g["off"], g["sid"], g["feat"]))
    monkeypatch.setenv("VV_FUSE", "1" if fuse else "0")
    caffe_host.set_device(0); caffe_host.set_precision("f16x3")
    tsrc = records_util.write_vvrs(tmp_path / "test.vvrs", records_util.test_window_records(g["test_data"], g["test_vid"]))
    idfile = tmp_path / "id_to_class.txt"
    idfile.write_text("".join("%d,%d\n" % (k, v) for k, v in zip(g["id_keys"].tolist(), g["id_vals"].tolist())))
    net_txt = prototxt.train_net(B=B, C=C, Nn=Nn, K=K, N=N, dropout=0, max_buffer_size=P, swap=swap, max_same=max_same, source=src,
                                 test=dict(batch=int(g["test_batch"]), frames=g["test_data"].shape[1], source=tsrc,
                                           id_to_class_file=str(idfile), exclude_same=True))
    sol = caffe_host.Solver(prototxt.solver(base_lr=hyper["base_lr"], momentum=hyper["momentum"], weight_decay=hyper["weight_decay"],
                                            gamma=hyper["gamma"], power=hyper["power"], display=0, snapshot=0, test_iter=2,
                                            test_interval=1000000), net_txt)
    assert sol.net.layer_names == [str(x) for x in g["layer_names"]]
    assert sol.test_net(0).layer_names == [str(x) for x in g["test_layer_names"]]
    sol.net.set_param(0, g["W0"]); sol.net.set_param(1, g["b0"])
    # Solver::Test (2 iterations of the TEST net on the shared weights) against the reference's TEST net; the reference
    # reports its outputs in Net::Init's lexicographic order, the scores are matched by name
    order = [str(x) for x in g["test_output_names"]]
    assert sol.test_net(0).blob_names[-3:] == ["test_map", "test_hit_at_1", "test_hit_at_5"]
    before = dict(zip(order, sol.test(0)))
    assert abs(before["test_map"] - g["test_before"][0]) < 2e-3 and abs(before["test_hit_at_1"] - g["test_before"][1]) < 1e-6 \
        and abs(before["test_hit_at_5"] - g["test_before"][2]) < 1e-6, before
    for it in range(len(g["loss"])):
        loss = sol.step()
        assert abs(loss - g["loss"][it]) < 1e-5 * max(1, abs(g["loss"][it])), it
    after = dict(zip(order, sol.test(0)))
    assert abs(after["test_map"] - g["test_after"][0]) < 2e-3 and abs(after["test_hit_at_1"] - g["test_after"][1]) < 1e-6 \
        and abs(after["test_hit_at_5"] - g["test_after"][2]) < 1e-6, after
    assert rel(sol.net.param(0), g["W"].reshape(-1)) < 1e-5 and rel(sol.net.param(1), g["b"]) < 1e-5
    assert rel(sol.history(0), g["hW"].reshape(-1)) < 1e-5 and rel(sol.history(1), g["hb"]) < 1e-5
    sol.close()



@pytest.mark.parametrize("prec,fused_gather", [("fp32_simt", False), ("f16x3", True)])
def test_trainer_follows_reference_trajectory_l1_step_policy(prec, fused_gather):
    """The fixture's second reference trajectory: L1 hinge (max_margin_loss norm: L1), "step" learning-rate policy and L1
    weight regularisation -- the other branches of the loss gradient and of SGDSolver::ComputeUpdateValue."""
    g, (B, C, Nn, P, swap, max_same), _ = _solver_fixture()
    lr, mom, wd, gamma, stepsize = [float(x) for x in g["alt_hyper"]]
    N, K = g["W0"].shape
    bank = torch.as_tensor(g["feat"]).cuda()
    smp = ops.Sampler(g["vid"], g["off"], g["sid"], B, C, Nn, P, swap, max_same, 100, rand_seed=1)
    tr = ops.Trainer(ops.trainer_cfg(B, C, Nn, K, N, norm=1, dropout_ratio=0.0, prec=prec, base_lr=lr, momentum=mom, weight_decay=wd,
                                     lr_policy="step", gamma=gamma, power=0.0, stepsize=int(stepsize), reg_type=1))
    tr.set_weights(torch.as_tensor(g["W0"]).cuda(), torch.as_tensor(g["b0"]).cuda())
    if fused_gather:
        tr.set_bank(bank)
    for it in range(len(g["alt_loss"])):
        idx, quirk = smp.next()
        tr.step(bank, torch.as_tensor(idx).cuda(), torch.as_tensor(quirk).cuda(), None, it=it)
        assert abs(tr.tensor("loss").item() - g["alt_loss"][it]) < 1e-5 * max(1, abs(g["alt_loss"][it])), it
        assert tr.tensor("violations").item() == g["alt_violations"][it], it
    for name, key in (("W", "alt_W"), ("b", "alt_b"), ("W_hist", "alt_hW"), ("b_hist", "alt_hb")):
        assert rel(tr.tensor(name), g[key]) < 1e-5, name
    tr.close(); smp.close()


# ---- north_star: loss curves over 1k steps against the REFERENCE's own pipeline -------------------------------------
def _product_curve(prec, fused_gather):
    import sys
    sys.path.insert(0, GOLD)
    from make_curve_golden import CURVE as c, problem
    vid, off, sid, _, W0, b0 = problem()
    bank = ops.fill_bank(c["V"] * c["S"], c["K"], c["bank_seed"])         # bit-identical to the fixture's ops.bank_host
    smp = ops.Sampler(vid, off, sid, c["B"], c["C"], c["Nn"], c["P"], c["swap"], c["max_same"], 100, rand_seed=1)
    tr = ops.Trainer(ops.trainer_cfg(c["B"], c["C"], c["Nn"], c["K"], c["N"], dropout_ratio=0.0, prec=prec, base_lr=c["base_lr"],
                                     gamma=c["gamma"], power=c["power"], momentum=c["momentum"], weight_decay=c["weight_decay"]))
    tr.set_weights(torch.as_tensor(W0).cuda(), torch.as_tensor(b0).cuda())
    if fused_gather:
        tr.set_bank(bank)
    loss = torch.zeros(c["steps"], device="cuda"); viol = torch.zeros(c["steps"], device="cuda")
    for it in range(c["steps"]):
        idx, quirk = smp.next()
        tr.step(bank, torch.as_tensor(idx).cuda(), torch.as_tensor(quirk).cuda(), None, it=it)
        loss[it] = tr.tensor("loss")[0]; viol[it] = tr.tensor("violations")[0]
    out = loss.cpu().numpy(), viol.cpu().numpy()
    tr.close(); smp.close()
    return out


@pytest.mark.parametrize("prec,fused_gather,tol,smooth_tol", [
    ("f16x3", True, 1e-4, 1e-4), ("tf32x3", False, 1e-4, 1e-4), ("fp32_simt", False, 1e-4, 1e-4),
    ("tf32", False, 1e-2, 1e-2), ("bf16", True, 1e-2, 1e-2)])
def test_loss_curve_1k_steps_against_reference_pipeline(prec, fused_gather, tol, smooth_tol):
    """tests/golden/curve_ref.npz: 1 000 iterations of the reference's own data layer + Net + SGDSolver (compiled from its
    sources, make_curve_golden.py).  Same sampler stream (bit-exact), same W0, no dropout layer.  The fp32-parity modes
    follow the reference curve step by step (1e-4 relative at every one of the 1 000 steps and on the 20-step running
    mean; the oracle restatement itself is within 6e-6); TF32 / bf16 stay within the north-star's 1e-2 at every step."""
    g = np.load(os.path.join(GOLD, "curve_ref.npz"))
    loss, viol = _product_curve(prec, fused_gather)
    assert np.isfinite(loss).all()
    ref = g["loss"]
    assert ref[-50:].mean() < ref[:50].mean() - 0.03                     # the reference run trains
    step_err = np.abs(loss - ref).max() / np.abs(ref).max()
    k = np.ones(20) / 20
    smooth_err = np.abs(np.convolve(loss, k, "valid") - np.convolve(ref, k, "valid")).max() / np.abs(ref).max()
    print("curve vs reference pipeline: %s step %.2e smoothed %.2e violations differ at %d steps (max %d)" % (
        prec, step_err, smooth_err, int((viol != g["viol"]).sum()), int(np.abs(viol - g["viol"]).max())))
    assert step_err < tol, (prec, step_err)
    assert smooth_err < smooth_tol, (prec, smooth_err)
    if tol <= 1e-4:
        # violation counts (of 1 280 hinge terms per step) are integers decided by the sign of score differences; a few
        # of them sit within fp32 rounding of zero at this near-collapsed start (all cosine scores ~ 1), where the GEMM's
        # summation order decides: a handful of flips per step, no drift
        dv = np.abs(viol - g["viol"])
        assert dv.max() <= 4 and dv.mean() < 0.6, (prec, dv.max(), dv.mean())
