"""The drop-in boundary: the reference's operator interface (Blob / Layer<Dtype> / Net / SGDSolver, built from a
prototxt with the shipped net's structure) running layer by layer on the C-ABI kernels, and the same net fused;
both against the oracle.  These read like the reference's net / solver tests (test_net.cpp, test_gradient_based_solver.cpp)."""
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT
from videovector_b200 import caffe_host, prototxt

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


CFG = dict(B=16, C=5, Nn=10, K=256, N=64, dropout=0.5, videos=64, shots=24, max_buffer_size=200)


def make_net(prec, fuse, W0, b0, mask):
    caffe_host.set_device(0)
    caffe_host.set_precision(prec)
    net = caffe_host.Net(prototxt.train_net(**CFG))
    net.set_param(0, W0); net.set_param(1, b0)
    net.set_dropout_mask(mask)
    if fuse:
        ok, why = net.enable_fusion()
        assert ok, why
    return net


def problem():
    rng = np.random.RandomState(1701)
    R = CFG["C"] + CFG["Nn"]
    W0 = rng.normal(0, 0.02, (CFG["N"], CFG["K"])).astype(np.float32)
    b0 = rng.normal(0, 0.01, CFG["N"]).astype(np.float32)
    mask = (rng.uniform(0, 1, (R * CFG["B"], CFG["N"])) > 0.5).astype(np.uint32)
    return W0, b0, mask, torch.as_tensor(mask.astype(np.int32)).cuda()


@pytest.mark.parametrize("prec,tol", [("fp32_simt", 1e-5), ("tf32x3", 1e-5), ("f16x3", 1e-5)])
def test_layer_by_layer_net_matches_oracle(oracle, prec, tol):
    W0, b0, mask, mask_dev = problem()
    net = make_net(prec, False, W0, b0, mask_dev)
    B, C, Nn, K, N = CFG["B"], CFG["C"], CFG["Nn"], CFG["K"], CFG["N"]
    # net structure the reference's Net::Init would produce
    assert net.num_params == 2
    nb = dict(zip(net.layer_names, net.layer_need_backward))
    assert not nb["shot_windows"] and not nb["slice_input_data"] and not nb["batch_concat_input"] and not nb["flatten_input"]
    assert nb["fc7"] and nb["max_margin_loss"] and nb["context_feature_word_embedding_norm_0_split"]
    for it in range(2):
        loss = net.forward_backward()
        data = net.blob("data").reshape(B, C + Nn, K)
        ref = oracle.net_forward_backward(data, W0, b0, mask, B, C, Nn, margin=2.0, norm=2, dropout_ratio=0.5,
                                          want=("loss", "violations", "dW", "db", "X", "Z", "H", "target_score", "neg_score", "dZ"))
        assert abs(loss - ref["loss"][0]) < tol * max(1, ref["loss"][0])
        assert np.array_equal(net.blob("original_feature").reshape(-1, K), ref["X"])          # slice + concat + flatten: bit-exact
        assert rel(net.blob("ip1_nonorm").reshape(-1, N), ref["Z"]) < tol
        assert rel(net.blob("ip2").reshape(-1, N), ref["H"]) < tol
        assert rel(net.blob("target_score").reshape(B, Nn), ref["target_score"]) < tol
        assert rel(net.blob("negative_score").reshape(B, Nn), ref["neg_score"]) < tol
        assert net.blob("train_violations").item() == ref["violations"][0]
        assert net.blob("loss_output", diff=True).item() == 1.0                              # the loss weight lives in top.diff
        assert rel(net.blob("ip1_nonorm", diff=True).reshape(-1, N), ref["dZ"]) < 2 * tol
        assert rel(net.param(0, diff=True).reshape(N, K), ref["dW"]) < 2 * tol
        assert rel(net.param(1, diff=True), ref["db"]) < 2 * tol
    net.close()


@pytest.mark.parametrize("prec,tol", [("tf32x3", 1e-5), ("f16x3", 1e-5), ("bf16", 5e-2)])
def test_fused_net_matches_layer_by_layer_and_oracle(oracle, prec, tol):
    W0, b0, mask, mask_dev = problem()
    B, C, Nn, K, N = CFG["B"], CFG["C"], CFG["Nn"], CFG["K"], CFG["N"]
    plain = make_net("fp32_simt", False, W0, b0, mask_dev)
    fused = make_net(prec, True, W0, b0, mask_dev)
    for it in range(2):                     # both data layers replay the same sampler stream
        l_plain = plain.forward_backward()
        l_fused = fused.forward_backward()
        data = plain.blob("data").reshape(B, C + Nn, K)
        ref = oracle.net_forward_backward(data, W0, b0, mask, B, C, Nn, dropout_ratio=0.5)
        assert abs(l_fused - ref["loss"][0]) < tol * max(1, ref["loss"][0]) and abs(l_fused - l_plain) < tol * max(1, l_plain)
        assert fused.blob("train_violations").item() == plain.blob("train_violations").item() or tol > 1e-5
        if tol <= 1e-5:
            assert rel(fused.param(0, diff=True), ref["dW"].reshape(-1)) < 2 * tol
            assert rel(fused.param(1, diff=True), ref["db"]) < 2 * tol
        else:
            d = fused.param(0, diff=True).astype(np.float64) - ref["dW"].reshape(-1)
            assert np.linalg.norm(d) / np.linalg.norm(ref["dW"]) < 10 * tol
    plain.close(); fused.close()


def test_fusion_refuses_other_graphs():
    caffe_host.set_device(0)
    txt = prototxt.train_net(**CFG).replace("slice_dim: 0", "slice_dim: 1", 1) if False else prototxt.train_net(**CFG)
    # a leaky ReLU is not what the fused kernels compute: the pass must decline and say why
    leaky = txt.replace('type: RELU\n  top: "ip2"\n  bottom: "ip1_nonorm"\n', 'type: RELU\n  top: "ip2"\n  bottom: "ip1_nonorm"\n  relu_param { negative_slope: 0.1 }\n')
    net = caffe_host.Net(leaky)
    ok, why = net.enable_fusion()
    assert not ok and "ReLU" in why
    net.close()


@pytest.mark.parametrize("fuse", [False, True])
def test_solver_trajectory(oracle, fuse, monkeypatch):
    """SGDSolver::Solve's loop body through the reference interface: 4 iterations, layer by layer and fused,
    against the oracle's update (ref: test_gradient_based_solver.cpp checks weights AND history)."""
    monkeypatch.setenv("VV_FUSE", "1" if fuse else "0")
    W0, b0, mask, mask_dev = problem()
    B, C, Nn, K, N = CFG["B"], CFG["C"], CFG["Nn"], CFG["K"], CFG["N"]
    caffe_host.set_device(0); caffe_host.set_precision("tf32x3")
    sol = caffe_host.Solver(prototxt.solver(base_lr=0.05, display=0), prototxt.train_net(**CFG))
    sol.net.set_param(0, W0); sol.net.set_param(1, b0); sol.net.set_dropout_mask(mask_dev)
    twin = make_net("fp32_simt", False, W0, b0, mask_dev)         # same sampler stream, only used to read the data blob
    W, b = W0.copy(), b0.copy(); hW = np.zeros_like(W); hb = np.zeros_like(b)
    for it in range(4):
        assert abs(sol.learning_rate() - oracle.learning_rate("inv", 0.05, 1e-3, 0.75, 1, it)) < 1e-12
        loss = sol.step()
        twin.forward()
        data = twin.blob("data").reshape(B, C + Nn, K)
        ref = oracle.net_forward_backward(data, W, b, mask, B, C, Nn, dropout_ratio=0.5)
        rate = oracle.learning_rate("inv", 0.05, 1e-3, 0.75, 1, it)
        W, dWo, hW = oracle.sgd_update(W, ref["dW"], hW, rate * 1.0, 0.9, 5e-4 * 1.0)
        b, dbo, hb = oracle.sgd_update(b, ref["db"], hb, rate * 2.0, 0.9, 0.0)
        assert abs(loss - ref["loss"][0]) < 1e-5 * max(1, ref["loss"][0])
        assert rel(sol.net.param(0), W.reshape(-1)) < 1e-5 and rel(sol.net.param(1), b) < 1e-5
        assert rel(sol.history(0), hW.reshape(-1)) < 3e-5 and rel(sol.history(1), hb) < 3e-5
        assert rel(sol.net.param(0, diff=True), dWo.reshape(-1)) < 3e-5                    # diff := history
    assert sol.iter == 4
    sol.close(); twin.close()


def test_cli_train_log_format(tmp_path):
    """`vv_caffe train` prints the fork's log lines (solver.cpp:195-217) that parse_log.sh / plot_training_stats.py scrape."""
    exe = os.path.join(ROOT, "build", "vv_caffe")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "tools"], cwd=ROOT)
    netp = tmp_path / "net.prototxt"; solp = tmp_path / "solver.prototxt"
    netp.write_text(prototxt.train_net(**CFG))
    solp.write_text(prototxt.solver(str(netp), display=1, max_iter=3))
    r = subprocess.run([exe, "train", "--solver=%s" % solp], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "Iteration 0, loss = " in r.stderr and "Iteration 2, lr = " in r.stderr
    assert "Train net output #0: loss_output = iter = 0 value = " in r.stderr
    assert "Train net output #1: train_violations = iter = 2 value = " in r.stderr
    # a CPU solver must fail loudly: there is no CPU fallback
    solp.write_text(prototxt.solver(str(netp), display=1, max_iter=1).replace("solver_mode: GPU", "solver_mode: CPU"))
    r = subprocess.run([exe, "train", "--solver=%s" % solp], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


# ---- .caffemodel / .solverstate ---------------------------------------------------------------
@pytest.mark.parametrize("fuse", [False, True])
def test_snapshot_restore_roundtrip(tmp_path, monkeypatch, fuse):
    """Solver::Snapshot writes <prefix>_iter_N.caffemodel/.solverstate that (a) the real protobuf runtime parses as the
    reference's NetParameter / SolverState, (b) a fresh solver restores: weights, momentum history and iter continue
    (ref: solver.cpp:320-341, 418-429; net.cpp:692-727)."""
    from test_caffemodel_io import _classes
    monkeypatch.setenv("VV_FUSE", "1" if fuse else "0")
    W0, b0, mask, mask_dev = problem()
    N, K = CFG["N"], CFG["K"]
    caffe_host.set_device(0); caffe_host.set_precision("f16x3")
    prefix = str(tmp_path / "snap" / "videovec")
    os.makedirs(os.path.dirname(prefix))
    sol = caffe_host.Solver(prototxt.solver(base_lr=0.05, display=0, snapshot=0, snapshot_prefix=prefix), prototxt.train_net(**CFG))
    sol.net.set_param(0, W0); sol.net.set_param(1, b0); sol.net.set_dropout_mask(mask_dev)
    for _ in range(3):
        sol.step()
    model = sol.snapshot()
    assert model == prefix + "_iter_3.caffemodel" and os.path.exists(model) and os.path.exists(prefix + "_iter_3.solverstate")
    W3, b3, hW3, hb3 = sol.net.param(0), sol.net.param(1), sol.history(0), sol.history(1)
    assert np.abs(hW3).max() > 0
    # (a) the reference's own schema, real protobuf runtime
    get = _classes()
    net = get("NetParameter")(); net.ParseFromString(open(model, "rb").read())
    fc7 = [l for l in net.layers if l.name == "fc7"][0]
    assert fc7.type == 14 and len(fc7.blobs) == 2 and fc7.inner_product_param.num_output == N
    assert (fc7.blobs[0].num, fc7.blobs[0].channels, fc7.blobs[0].height, fc7.blobs[0].width) == (1, 1, N, K)
    assert (fc7.blobs[1].num, fc7.blobs[1].channels, fc7.blobs[1].height, fc7.blobs[1].width) == (1, 1, 1, N)
    assert np.array_equal(np.array(fc7.blobs[0].data, np.float32), W3) and np.array_equal(np.array(fc7.blobs[1].data, np.float32), b3)
    assert len(fc7.blobs[0].diff) == 0                                      # snapshot_diff defaults to false
    assert [l.name for l in net.layers] == sol.net.layer_names              # every layer is listed, split layers included
    st = get("SolverState")(); st.ParseFromString(open(prefix + "_iter_3.solverstate", "rb").read())
    assert st.iter == 3 and st.learned_net == model and len(st.history) == 2
    assert np.array_equal(np.array(st.history[0].data, np.float32), hW3) and np.array_equal(np.array(st.history[1].data, np.float32), hb3)
    # (b) a fresh solver picks everything up
    sol2 = caffe_host.Solver(prototxt.solver(base_lr=0.05, display=0, snapshot=0), prototxt.train_net(**CFG))
    sol2.net.set_dropout_mask(mask_dev)
    sol2.restore(prefix + "_iter_3.solverstate")
    assert sol2.iter == 3 and abs(sol2.learning_rate() - sol.learning_rate()) < 1e-12
    assert np.array_equal(sol2.net.param(0), W3) and np.array_equal(sol2.net.param(1), b3)
    l2 = sol2.step()
    assert np.isfinite(l2) and sol2.iter == 4
    # the restored momentum took part in the first step after the restore: W4 - W3 = -(momentum * h3 + rate * grad)
    step = sol2.net.param(0) - W3
    assert rel(sol2.history(0), -step) < 1e-6
    assert np.abs(sol2.history(0) - 0.9 * hW3).max() < np.abs(sol2.history(0)).max()      # history carried, not restarted from 0
    # restoring into a net whose fused trainer already exists also lands in the trainer's buffers
    sol2.restore(prefix + "_iter_3.solverstate")
    assert sol2.iter == 3 and np.array_equal(sol2.net.param(0), W3) and np.array_equal(sol2.history(0), hW3)
    sol2.step()
    sol.close(); sol2.close()


def test_copy_trained_layers_semantics(tmp_path):
    """Net::CopyTrainedLayersFrom: layers matched by name, unknown source layers ignored, shape mismatch fatal
    (ref: net.cpp:692-722)."""
    caffe_host.set_device(0); caffe_host.set_precision("f16x3")
    N, K = CFG["N"], CFG["K"]
    rng = np.random.RandomState(5)
    W = rng.normal(0, 0.02, (N, K)).astype(np.float32); b = rng.normal(0, 0.01, N).astype(np.float32)
    good = tmp_path / "good.caffemodel"
    caffe_host.write_binary_proto(good, "NetParameter",
        'name: "x"\nlayers { name: "somebody_elses_conv" type: CONVOLUTION blobs { num: 1 channels: 1 height: 1 width: 2 } }\n'
        'layers { name: "fc7" type: INNER_PRODUCT blobs { num: 1 channels: 1 height: %d width: %d } blobs { num: 1 channels: 1 height: 1 width: %d } }\n' % (N, K, N),
        {"layers[0].blobs[0].data": np.zeros(2, np.float32), "layers[1].blobs[0].data": W, "layers[1].blobs[1].data": b})
    net = caffe_host.Net(prototxt.train_net(**CFG))
    net.copy_trained_from(good)
    assert np.array_equal(net.param(0), W.reshape(-1)) and np.array_equal(net.param(1), b)
    out = tmp_path / "resaved.caffemodel"
    net.save(out, write_diff=True)
    _, arrays = caffe_host.read_binary_proto(out, "NetParameter")
    keys = sorted(arrays)
    assert len(keys) == 4 and all(".blobs[" in k for k in keys) and sum(k.endswith(".diff") for k in keys) == 2
    bad = tmp_path / "bad.caffemodel"
    caffe_host.write_binary_proto(bad, "NetParameter",
        'layers { name: "fc7" type: INNER_PRODUCT blobs { num: 1 channels: 1 height: %d width: %d } blobs { num: 1 channels: 1 height: 1 width: %d } }\n' % (N + 1, K, N),
        {"layers[0].blobs[0].data": np.zeros((N + 1) * K, np.float32), "layers[0].blobs[1].data": b})
    with pytest.raises(Exception):
        net.copy_trained_from(bad)
    net.close()


def test_cli_snapshot_and_resume(tmp_path):
    """`vv_caffe train` honours snapshot / snapshot_prefix / snapshot_after_train and --snapshot / --weights
    (ref: solver.cpp:180-184, 225-227; tools/caffe.cpp:106-118)."""
    exe = os.path.join(ROOT, "build", "vv_caffe")
    netp = tmp_path / "net.prototxt"; solp = tmp_path / "solver.prototxt"
    prefix = str(tmp_path / "vv")
    netp.write_text(prototxt.train_net(**CFG))
    solp.write_text(prototxt.solver(str(netp), display=1, max_iter=4, snapshot=2, snapshot_prefix=prefix))
    r = subprocess.run([exe, "train", "--solver=%s" % solp], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    for it in (2, 4):                                  # periodic at iter 2, final after training at iter 4
        assert os.path.exists("%s_iter_%d.caffemodel" % (prefix, it)) and os.path.exists("%s_iter_%d.solverstate" % (prefix, it))
    assert not os.path.exists(prefix + "_iter_0.caffemodel")
    text, arrays = caffe_host.read_binary_proto(prefix + "_iter_4.solverstate", "SolverState")
    assert "iter: 4" in text and len(arrays) == 2
    solp.write_text(prototxt.solver(str(netp), display=1, max_iter=6, snapshot=0, snapshot_prefix=prefix))
    r = subprocess.run([exe, "train", "--solver=%s" % solp, "--snapshot=%s_iter_4.solverstate" % prefix], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "Resuming from" in r.stderr and "Iteration 4, loss = " in r.stderr and "Iteration 5, lr = " in r.stderr
    assert "Iteration 0, loss" not in r.stderr and os.path.exists(prefix + "_iter_6.caffemodel")
    r = subprocess.run([exe, "train", "--solver=%s" % solp, "--weights=%s_iter_4.caffemodel" % prefix, "--iterations=1"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "Finetuning from" in r.stderr and "Iteration 0, loss" in r.stderr
    r = subprocess.run([exe, "train", "--solver=%s" % solp, "--weights=a", "--snapshot=b"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "not both" in r.stderr


# ---- TEST-phase net ---------------------------------------------------------------------------
def test_test_phase_net_and_solver_test(tmp_path, oracle, monkeypatch):
    """The shipped file's TEST graph through the reference interface: data -> average_for_test -> shared fc7 ->
    test_norm -> retrieval_stats, run by Solver::Test with the TRAIN net's (fused) weights, against the oracle's
    restatement of the TEST graph + RetrievalStatsLayer (ref: solver.cpp:243-317, retrieval_stats_layer.cpp)."""
    monkeypatch.setenv("VV_FUSE", "1")
    caffe_host.set_device(0); caffe_host.set_precision("f16x3")
    tv, ts, F, TB = 24, 10, 4, 60
    idfile = tmp_path / "id_to_class.txt"
    classes = {v: (v * 7) % 5 for v in range(tv)}
    idfile.write_text("".join("%d,%d\n" % (v, c) for v, c in classes.items() if v != 3))      # video 3 unlisted -> class 0 (map operator[])
    classes[3] = 0
    net_txt = prototxt.train_net(test=dict(batch=TB, frames=F, videos=tv, shots=ts, seed=99, id_to_class_file=str(idfile), exclude_same=True), **CFG)
    sol = caffe_host.Solver(prototxt.solver(base_lr=0.05, display=0, snapshot=0, test_iter=2, test_interval=1000), net_txt)
    W0, b0, mask, mask_dev = problem()
    sol.net.set_param(0, W0); sol.net.set_param(1, b0); sol.net.set_dropout_mask(mask_dev)
    tnet = sol.test_net(0)
    assert tnet.layer_names == ["shot_windows", "slice_input_data", "batch_concat_input_test", "flatten_input", "slice_test",
                                "average_for_test", "fc7", "fc7_relu", "test_norm", "retrieval_stats"]
    for rounds in range(2):                     # before any training step, then after two fused steps (weights moved)
        scores = sol.test(0)
        assert scores.shape == (3,)
        W, b = sol.net.param(0).reshape(CFG["N"], CFG["K"]), sol.net.param(1)
        # the last of the two test iterations is still in the test net's blobs
        data = tnet.blob("data").reshape(TB, F, CFG["K"])
        vids = tnet.blob("video_ids").reshape(-1).astype(np.int32)
        xbar_ref, E_ref = oracle.test_embed(data, W, b)
        assert rel(tnet.blob("original_feature").reshape(TB, -1), xbar_ref) < 1e-6
        assert rel(tnet.blob("ip2_norm").reshape(TB, -1), E_ref) < 1e-5
        labels = np.array([classes[int(v)] for v in vids], np.int32)
        ref = oracle.retrieval_stats(E_ref, vids, labels, True)
        last = np.array([tnet.blob("test_map")[0, 0, 0, 0], tnet.blob("test_hit_at_1")[0, 0, 0, 0], tnet.blob("test_hit_at_5")[0, 0, 0, 0]])
        assert np.abs(last - np.array([ref["map"], ref["hit1"], ref["hit5"]])).max() < 2e-3
        assert (scores >= 0).all() and (scores <= 1).all()
        assert tnet.blob_names[-3:] == ["test_map", "test_hit_at_1", "test_hit_at_5"]
        # windows are served in order: consecutive batches differ, the pass after next repeats nothing yet
        assert vids[0] == ((2 * (2 * rounds + 1) * TB) - TB) % tv
        if rounds == 0:
            sol.step(); sol.step()
    sol.close()


def test_cli_runs_test_phase(tmp_path):
    """`vv_caffe train` with test_iter / test_interval prints the reference's test log lines (solver.cpp:252-253, 306-313)."""
    exe = os.path.join(ROOT, "build", "vv_caffe")
    idfile = tmp_path / "id.txt"; idfile.write_text("".join("%d,%d\n" % (v, v % 3) for v in range(16)))
    netp = tmp_path / "net.prototxt"; solp = tmp_path / "solver.prototxt"
    netp.write_text(prototxt.train_net(test=dict(batch=32, videos=16, shots=8, id_to_class_file=str(idfile)), **CFG))
    solp.write_text(prototxt.solver(str(netp), display=1, max_iter=4, snapshot=0, test_iter=1, test_interval=2))
    r = subprocess.run([exe, "train", "--solver=%s" % solp], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stderr.count("Testing net (#0)") == 3                               # iterations 0 (test_initialization), 2 and 4
    for it in (0, 2, 4):
        assert "Iteration %d, Testing net (#0)" % it in r.stderr
    # net outputs come out of a std::set of blob names (net.cpp:158-165), i.e. in lexicographic order
    assert "    Test net output #0: test_hit_at_1 = " in r.stderr and "    Test net output #2: test_map = " in r.stderr


def _record_dataset(rng, V=70, K=256):
    counts = rng.randint(2, 21, size=V)
    vid = (rng.permutation(V) + 100).astype(np.int32)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    sid = np.concatenate([np.sort(rng.choice(300, c, replace=False)) for c in counts]).astype(np.int32)
    feat = np.maximum(rng.normal(0, 1, (off[-1], K)), 0).astype(np.float32)
    return vid, off, sid, feat


@pytest.mark.parametrize("container", ["vvrs", "lmdb"])
def test_net_trains_from_video_shots_records(tmp_path, oracle, container):
    """`source:` naming real VideoShots records: the data layer decodes them into the resident bank,
    its data blob must equal records -> oracle sampler -> copy, and the fused step on that bank must match the oracle."""
    import records_util
    from videovector_b200 import ops
    rng = np.random.RandomState(11)
    vid, off, sid, feat = _record_dataset(rng, K=CFG["K"])
    recs = records_util.video_shots_records(vid, off, sid, feat)
    if container == "vvrs":
        src = records_util.write_vvrs(tmp_path / "train.vvrs", recs)
    else:
        from lmdb_writer import write_lmdb
        src = str(tmp_path / "train_lmdb"); write_lmdb(src, recs)
    cfg = dict(CFG, max_buffer_size=60); cfg.pop("videos"); cfg.pop("shots")
    W0, b0, mask, mask_dev = problem()
    caffe_host.set_device(0); caffe_host.set_precision("f16x3")
    B, C, Nn, K, N = cfg["B"], cfg["C"], cfg["Nn"], cfg["K"], cfg["N"]
    osmp = oracle.Sampler(vid, off, sid, feat, K, B, C, Nn, 60, 50, 6, 100, seed=1)
    blobs = [osmp.next()[2] for _ in range(3)]
    osmp.close()
    net = caffe_host.Net(prototxt.train_net(source=src, **cfg))
    net.set_param(0, W0); net.set_param(1, b0); net.set_dropout_mask(mask_dev)
    for it in range(2):                               # layer by layer: the data blob itself is materialised
        loss = net.forward_backward()
        assert np.array_equal(net.blob("data").reshape(B, C + Nn, K), blobs[it])
        ref = oracle.net_forward_backward(blobs[it], W0, b0, mask, B, C, Nn, margin=2.0, norm=2, dropout_ratio=0.5, want=("loss", "dW"))
        assert abs(loss - ref["loss"]) < 1e-5 * max(1, abs(ref["loss"]))
        assert rel(net.param(0, diff=True).reshape(N, K), ref["dW"]) < 1e-5
    net2 = caffe_host.Net(prototxt.train_net(source=src, **cfg))
    net2.set_param(0, W0); net2.set_param(1, b0); net2.set_dropout_mask(mask_dev)
    ok, why = net2.enable_fusion()
    assert ok, why
    loss = net2.forward_backward()                    # fused: gather folded into the GEMMs, same first batch
    ref = oracle.net_forward_backward(blobs[0], W0, b0, mask, B, C, Nn, margin=2.0, norm=2, dropout_ratio=0.5, want=("loss", "dW"))
    assert abs(loss - ref["loss"]) < 1e-5 * max(1, abs(ref["loss"]))
    assert rel(net2.param(0, diff=True).reshape(N, K), ref["dW"]) < 1e-5


@pytest.mark.parametrize("name", ["skip_window", "neg_window", "skip_neg_past"])
def test_data_layer_rand_skip_and_negative_dataset(tmp_path, name):
    """The data layer's `rand_skip` and `negative_dataset` (video_sampled_shots_data_layer.cpp:104-180, 253-338) through the
    reference's interface: VideoSampledShotsDataLayer<float> of caffe_compat, built from a `layers { }` entry over two
    record files, must serve the data blobs the compiled reference layer served (tests/golden/sampler_opts_ref.npz)."""
    import records_util
    g = np.load(os.path.join(ROOT, "tests", "golden", "sampler_opts_ref.npz"))
    mode, B, C, Nn, P, swap, max_same, rand_skip, cseed, with_neg = [int(x) for x in g["cfg_" + name]]
    ref = g["blobs_" + name]
    src = records_util.write_vvrs(tmp_path / "main.vvrs", records_util.video_shots_records(g["vid"], g["off"], g["sid"], g["feat"]))
    extra = ""
    if with_neg:
        extra += ' negative_dataset: "%s"' % records_util.write_vvrs(
            tmp_path / "neg.vvrs", records_util.video_shots_records(g["nvid"], g["noff"], g["nsid"], g["nfeat"]))
    if rand_skip:
        extra += " rand_skip: %d" % rand_skip
    ctx = ["PAIRWISE", "WINDOW", "PAST", "PAST_CONTINUOUS", "PAST_CONTINUOUS_FIXED"][mode]
    text = ('layers { name: "shot_windows" type: VIDEO_SAMPLED_SHOTS_DATA top: "data" video_sampled_shots_data_param { '
            'source: "%s" backend: LMDB batch_size: %d num_negative_samples: %d max_buffer_size: %d negative_swap_percentage: %d '
            'max_same_video_negs: %d context_type: %s context_size: %d%s } }' % (src, B, Nn, P, swap, max_same, ctx, C, extra))
    caffe_host.set_device(0)
    caffe_host.set_seed(cseed)
    for batch in (0, 2, 7):
        _, tops, _ = caffe_host.run_layer(text, [], 1, forwards=batch + 1)
        assert np.array_equal(tops[0].reshape(ref[batch].shape), ref[batch]), "batch %d differs from the reference data layer" % batch


def test_test_net_reads_test_window_records(tmp_path):
    """TEST-phase data layer on TestVideoShotWindows records: items in DB order, label = video_id, wrap at the end
    (video_shot_window_test_data_layer.cpp:160-262)."""
    import records_util
    rng = np.random.RandomState(12)
    n, F, K, TB = 37, 4, CFG["K"], 16
    data = rng.normal(0, 1, (n, F, K)).astype(np.float32); vids = rng.randint(0, 12, n).astype(np.int32)
    src = records_util.write_vvrs(tmp_path / "test.vvrs", records_util.test_window_records(data, vids))
    idfile = tmp_path / "id_to_class.txt"; idfile.write_text("".join("%d,%d\n" % (v, v % 3) for v in range(12)))
    caffe_host.set_device(0); caffe_host.set_precision("f16x3")
    try:
        tnet = caffe_host.Net(prototxt.train_net(test=dict(batch=TB, frames=F, source=src, id_to_class_file=str(idfile)), **CFG), phase="TEST")
        for it in range(4):                            # 64 items over 37 records: wraps once
            tnet.forward()
            want = (np.arange(TB) + it * TB) % n
            assert np.array_equal(tnet.blob("data").reshape(TB, F, K), data[want])
            assert np.array_equal(tnet.blob("video_ids").reshape(-1).astype(np.int32), vids[want])
            assert 0 <= tnet.blob("test_map").reshape(-1)[0] <= 1
    finally:
        caffe_host.set_phase("TRAIN")
