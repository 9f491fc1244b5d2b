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
