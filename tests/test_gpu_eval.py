"""TEST-phase evaluation on the device against the oracle's restatement of the shipped TEST graph
and RetrievalStatsLayer (retrieval_stats_layer.cpp:98-140, 143-359)."""
import numpy as np
import pytest
import torch

from videovector_b200 import ops

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a.cpu() if torch.is_tensor(a) else a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def make_eval_set(B, F, K, N, classes, seed, rows=4000):
    """Class-structured synthetic features so that retrieval is neither trivial nor random."""
    rng = np.random.RandomState(seed)
    centers = rng.normal(0, 1, (classes, K)).astype(np.float32)
    video_of_row = rng.randint(0, B // 3 + 1, rows)                       # several shots per video
    class_of_video = rng.randint(0, classes, B // 3 + 1)
    bank = np.maximum(0, centers[class_of_video[video_of_row]] * 0.6 + rng.normal(0, 1, (rows, K))).astype(np.float32)
    base = rng.randint(0, rows - F, B)
    idx = (base[:, None] + np.arange(F)[None, :]).astype(np.int32)        # F consecutive frames of a window
    video_ids = (video_of_row[base] + 1000).astype(np.int32)
    labels = class_of_video[video_of_row[base]].astype(np.int32)
    labels[rng.rand(B) < 0.1] = -1                                         # unscored samples (:256-258)
    W = rng.normal(0, 0.02, (N, K)).astype(np.float32); b = rng.normal(0, 0.01, N).astype(np.float32)
    return bank, idx, video_ids, labels, W, b


@pytest.mark.parametrize("B,F,K,N", [(673, 4, 4096, 512), (100, 4, 256, 64), (37, 3, 64, 32)])
@pytest.mark.parametrize("prec,tol", [("f16x3", 1e-5), ("fp32_simt", 1e-5), ("bf16", 2e-2)])
def test_test_phase_embedding_matches_oracle(oracle, B, F, K, N, prec, tol):
    if prec != "fp32_simt" and (N % 8 or K % 8):
        pytest.skip("tensor-core path needs N, K % 8 == 0")
    bank, idx, _, _, W, b = make_eval_set(B, F, K, N, 7, B)
    oracle.use_openblas(0)
    xbar_ref, E_ref = oracle.test_embed(bank[idx], W, b)
    oracle.use_builtin_blas()
    xbar, E = ops.test_embed(torch.as_tensor(bank).cuda(), torch.as_tensor(idx).cuda(), torch.as_tensor(W).cuda(),
                             torch.as_tensor(b).cuda(), prec=prec)
    assert rel(xbar, xbar_ref) < 1e-6
    assert rel(E, E_ref) < tol, rel(E, E_ref)
    assert np.allclose((E.double() ** 2).sum(1).cpu().numpy(), 1.0, atol=1e-5)


@pytest.mark.parametrize("B", [673, 128, 1024, 5])
@pytest.mark.parametrize("exclude", [False, True])
def test_retrieval_stats_exact_on_shared_distances(oracle, B, exclude):
    """Same Gram matrix on both sides -> identical ranking -> AP / hit@1 / hit@5 equal to the last bit of a double
    ratio sum (ComputeStats, :98-140), including unscored queries and exclude_same_video_shots."""
    N = 64
    bank, idx, video_ids, labels, W, b = make_eval_set(B, 2, 64, N, 5, B + 1, rows=3000)
    _, E = oracle.test_embed(bank[idx], W, b)
    G = (E.astype(np.float32) @ E.astype(np.float32).T).astype(np.float32)
    ref = oracle.retrieval_stats(E, video_ids, labels, exclude, dist=-2.0 * G)
    got = ops.retrieval_stats(None, video_ids, labels, exclude, gram=torch.as_tensor(G).cuda())
    pq = got["per_query"].cpu().numpy()
    assert np.array_equal(pq[:, 1:], ref["per_query"][:, 1:])                 # hit counts: exact
    assert np.abs(pq[:, 0] - ref["per_query"][:, 0]).max() < 1e-12              # AP: double sums in a different order
    for k in ("map", "hit1", "hit5"):
        assert abs(got[k] - ref[k]) < 1e-12, k
    assert (pq[labels < 0] == -1).all() and (pq[labels >= 0] >= 0).all()


def test_retrieval_stats_end_to_end(oracle):
    """Embeddings and distances computed on each side independently (GPU fp32 FMA vs CPU BLAS): near-ties may swap,
    the metrics agree to 1e-3; and they are above chance (1/6) on the class-structured set even with random weights."""
    B, F, K, N = 673, 4, 1024, 128
    bank, idx, video_ids, labels, W, b = make_eval_set(B, F, K, N, 6, 99)
    oracle.use_openblas(0)
    _, E_ref = oracle.test_embed(bank[idx], W, b)
    ref = oracle.retrieval_stats(E_ref, video_ids, labels, True)
    oracle.use_builtin_blas()
    _, E = ops.test_embed(torch.as_tensor(bank).cuda(), torch.as_tensor(idx).cuda(), torch.as_tensor(W).cuda(), torch.as_tensor(b).cuda())
    got = ops.retrieval_stats(E, video_ids, labels, True)
    for k in ("map", "hit1", "hit5"):
        assert abs(got[k] - ref[k]) < 1e-3, (k, got[k], ref[k])
    assert got["map"] > 1.1 / 6 and 0 <= got["hit5"] <= 1 and 0 <= got["hit1"] <= 1


def test_retrieval_stats_argument_checks():
    E = torch.zeros(4, 8, device="cuda")
    with pytest.raises(Exception):
        ops.retrieval_stats(torch.zeros(9000, 8, device="cuda"), np.zeros(9000), np.zeros(9000))      # > 8192 items
    out = ops.retrieval_stats(E, [1, 1, 2, 2], [-1, -1, -1, -1])
    assert np.isnan(out["map"])                                                                       # nothing scored: 0/0 as in the reference


# ---- IdToWeightMapping ---------------------------------------------------------------------
@pytest.mark.parametrize("M,N,rows", [(4096, 512, 300), (64, 64, 1000), (7, 5, 3), (1, 33, 9)])
def test_id_lookup_matches_oracle_bit_exact(oracle, M, N, rows):
    """Forward = row gather; backward = scatter-add in increasing item order (id_to_weight_mapping_layer.cpp:61-106):
    the device result is bit-identical to the reference's sequential axpy loop, duplicates included."""
    rng = np.random.RandomState(M + N)
    table = rng.normal(0, 1, (rows, N)).astype(np.float32)
    ids = rng.randint(0, rows, M).astype(np.float32)                         # ids travel as floats in a Caffe blob
    ids[: M // 3] = ids[0]                                                     # heavy duplication
    dtop = rng.normal(0, 1, (M, N)).astype(np.float32)
    top = ops.id_lookup_forward(torch.as_tensor(table).cuda(), torch.as_tensor(ids).cuda())
    assert np.array_equal(top.cpu().numpy(), oracle.id_lookup_forward(table, ids))
    d = ops.id_lookup_backward(torch.as_tensor(dtop).cuda(), torch.as_tensor(ids).cuda(), rows)
    ref = oracle.id_lookup_backward(dtop, ids, rows)
    assert np.array_equal(d.cpu().numpy(), ref)
    untouched = np.setdiff1d(np.arange(rows), ids.astype(int))
    assert (d.cpu().numpy()[untouched] == 0).all()
