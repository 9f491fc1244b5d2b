"""Full-size parity of the path bench.py times (BASELINE configs[1]: B = 4096, K = 4096, N = 512, R = 15; gather fused
into the GEMMs, persistent multi-wave forward: 480 cluster tiles on 74 clusters) against the oracle's whole TRAIN net
(ref: inner_product_layer.cpp:61-106, relu_layer.cpp:10-36, dropout_layer.cpp:13-68, eltwise / normalization / sum /
split layers, max_margin_loss_layer.cpp:54-214) on identical inputs with an explicit dropout mask, and of configs[3]'s
shape (C = 17, Nn = 50, N = 1024: the wide rank-loss kernel, two phases per item) at B = 512."""
import numpy as np
import pytest
import torch

from videovector_b200 import ops
from videovector_b200._lib import DROPOUT_MASK01

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = torch.as_tensor(a).double().cpu(); b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    a = torch.as_tensor(a).double().cpu(); b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def data_blob(bank_np, idx, quirk):
    K = bank_np.shape[1]
    g = bank_np[idx]
    g[..., K - 1] = np.where(quirk >= 0, bank_np[np.maximum(quirk, 0), K - 1], g[..., K - 1])
    g[..., K - 1] = np.where(quirk == -1, 0.0, g[..., K - 1])
    return g


def run_case(oracle, prec, B, C, Nn, K, N, V, S, tol, update_check):
    R = C + Nn
    vid, off, sid = ops.synthetic_videos(V, S)
    bank = ops.fill_bank(V * S, K, 1234)
    smp = ops.Sampler(vid, off, sid, B, C, Nn, 5000, 50, 6, 100, rand_seed=1)
    rng = np.random.RandomState(1701)
    W0 = rng.normal(0, 0.01, (N, K)).astype(np.float32)
    b0 = rng.normal(0, 0.01, N).astype(np.float32)
    ratio = 0.9
    tr = ops.Trainer(ops.trainer_cfg(B, C, Nn, K, N, dropout_ratio=ratio, dropout_mode=DROPOUT_MASK01, prec=prec))
    tr.set_weights(torch.as_tensor(W0).cuda(), torch.as_tensor(b0).cuda())
    tr.set_bank(bank)                                        # the bench's path: the GEMMs gather the bank rows themselves
    smp.next()                                               # second batch: the K-1 copy quirk has history to show
    idx, quirk = smp.next()
    assert (quirk != -2).any()
    mask = (np.random.RandomState(3).uniform(0, 1, (R * B, N)) > ratio).astype(np.uint32)
    dm = torch.as_tensor(mask.astype(np.int32)).cuda()
    di, dq = torch.as_tensor(idx).cuda(), torch.as_tensor(quirk).cuda()
    tr.step(bank, di, dq, dm, it=0, do_update=False)
    if prec == "f16x3":                                      # steady state: the dZ operand scale trails the previous step
        tr.step(bank, di, dq, dm, it=0, do_update=False)
    torch.cuda.synchronize()
    oracle.use_openblas(0)
    ref = oracle.net_forward_backward(data_blob(bank.cpu().numpy(), idx, quirk), W0, b0, mask, B, C, Nn, margin=2.0, norm=2,
                                      dropout_ratio=ratio, want=("loss", "violations", "dW", "db", "H", "dZ"))
    oracle.use_builtin_blas()
    loss = tr.tensor("loss").item()
    assert abs(loss - ref["loss"][0]) <= tol * max(1.0, abs(ref["loss"][0])), (loss, ref["loss"][0])
    if tol <= 1e-5:
        assert tr.tensor("violations").item() == ref["violations"][0]
        assert rel(tr.tensor("H"), ref["H"]) < 2e-5
        assert rel(tr.dZ_from_operand(), ref["dZ"]) < 2e-5
        assert rel(tr.tensor("dW_raw"), ref["dW"]) < 2e-5, rel(tr.tensor("dW_raw"), ref["dW"])
        assert rel(tr.tensor("dW_raw")[:, K - 1], ref["dW"][:, K - 1]) < 2e-5       # the quirk column
        assert rel(tr.tensor("db_raw"), ref["db"]) < 2e-5
    else:
        # bf16: pre-activations within rounding of 0 flip their ReLU gate; gradients compared in the L2 norm
        assert rel(tr.tensor("H"), ref["H"]) < tol
        assert rel_l2(tr.dZ_from_operand(), ref["dZ"]) < 10 * tol
        assert rel_l2(tr.tensor("dW_raw"), ref["dW"]) < 10 * tol
        assert rel_l2(tr.tensor("db_raw"), ref["db"]) < 10 * tol
    # the same step again: loss and the bias gradient (column sums of dZ) are summed in a fixed order -> bit-identical
    db_first, loss_first = tr.tensor("db_raw").clone(), tr.tensor("loss").clone()
    tr.step(bank, di, dq, dm, it=0, do_update=False)
    assert torch.equal(tr.tensor("db_raw"), db_first) and torch.equal(tr.tensor("loss"), loss_first)
    if update_check:
        # one update on top (K4 / the fused update): W, b and both histories against the oracle's SGD step
        tr.step(bank, di, dq, dm, it=0, do_update=True)
        rate = oracle.learning_rate("inv", 1e-3, 1e-3, 0.75, 1, 0)
        W1, _, hW = oracle.sgd_update(W0, ref["dW"], np.zeros_like(W0), rate, 0.9, 5e-4)
        b1, _, hb = oracle.sgd_update(b0, ref["db"], np.zeros_like(b0), rate * 2.0, 0.9, 0.0)
        assert rel(tr.tensor("W"), W1) < 1e-5 and rel(tr.tensor("b"), b1) < 1e-5
        assert rel(tr.tensor("W_hist"), hW) < 2e-5 and rel(tr.tensor("b_hist"), hb) < 2e-5
        assert rel(tr.tensor("wlast"), W1[:, K - 1]) < 1e-5
    tr.close(); smp.close()


@pytest.mark.parametrize("prec,tol", [("f16x3", 1e-5), ("bf16", 5e-2)])
def test_bench_configuration_one_step_matches_oracle(oracle, prec, tol):
    run_case(oracle, prec, 4096, 5, 10, 4096, 512, 2048, 32, tol, update_check=(prec == "f16x3"))


def test_large_window_configuration_matches_oracle(oracle):
    """BASELINE configs[3] shape (R = 67 rows per item, N = 1024) at B = 512: 34 304 rows, 268 m-tiles x 4 n-tiles."""
    run_case(oracle, "f16x3", 512, 17, 50, 4096, 1024, 1024, 32, 1e-5, update_check=False)
