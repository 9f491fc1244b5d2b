"""Row S (sampler): the product's host index sampler reproduces the oracle's (= the reference's
state machine on the real libc rand()) index stream bit-exactly, and gathering bank rows with
(idx, quirk) reproduces the materialised data blob, K-1 copy quirk included.
ref: video_sampled_shots_data_layer.cpp:25-44,245-344,372-507,769-909; util/rng.hpp:43-54."""
import ctypes

import numpy as np
import pytest

from videovector_b200 import ops


def make_dataset(rng, V, smin, smax, K=None):
    counts = rng.randint(smin, smax + 1, size=V)
    video_id = (rng.permutation(V) + 100).astype(np.int32)           # arbitrary, unique ids
    shot_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    shot_ids = np.concatenate([np.sort(rng.choice(1000, c, replace=False)) for c in counts]).astype(np.int32)
    feat = rng.normal(0, 1, (shot_off[-1], K)).astype(np.float32) if K else None
    return video_id, shot_off, shot_ids, feat


def test_glibc_rand_matches_libc(vvlib):
    libc = ctypes.CDLL("libc.so.6")
    for seed in (1, 2, 12345, 0, 4294967295):
        libc.srand(ctypes.c_uint(seed))
        g = vvlib.vv_glibc_rand_create(seed)
        a = [libc.rand() for _ in range(5000)]
        b = [vvlib.vv_glibc_rand_next(g) for _ in range(5000)]
        vvlib.vv_glibc_rand_destroy(g)
        assert a == b, "seed %d" % seed


CASES = [
    # V, smin, smax, B, C, Nn, P, swap, max_same
    (300, 1, 12, 16, 5, 10, 50, 50, 6),     # many records skipped (n < C), ragged shot counts
    (200, 6, 40, 32, 5, 10, 100, 50, 6),    # the shipped parameters at small buffer
    (120, 18, 30, 8, 17, 50, 400, 50, 6),   # cfg-4: window +-8, 50 negatives
    (150, 3, 9, 8, 3, 4, 30, 0, 2),         # no swapping, max_same < default
    (150, 5, 9, 8, 5, 6, 40, 99, 6),        # n == C records (no same-video negatives), heavy swapping
    (64, 8, 8, 128, 5, 10, 60, 50, 6),      # cursor wraps several times per batch
]


@pytest.mark.parametrize("V,smin,smax,B,C,Nn,P,swap,max_same", CASES)
def test_index_stream_matches_oracle(vvlib, oracle, V, smin, smax, B, C, Nn, P, swap, max_same):
    rng = np.random.RandomState(V + B)
    video_id, shot_off, shot_ids, _ = make_dataset(rng, V, smin, smax)
    for seed in (1, 77):
        osmp = oracle.Sampler(video_id, shot_off, shot_ids, None, 4, B, C, Nn, P, swap, max_same, 100, seed=seed)
        ostream = [osmp.next()[:2] for _ in range(12)]      # consumes libc rand(): run the oracle to completion first
        ocur = osmp.cursor
        osmp.close()
        psmp = ops.Sampler(video_id, shot_off, shot_ids, B, C, Nn, P, swap, max_same, 100, rand_seed=seed)
        for it, (oi, oq) in enumerate(ostream):
            pi, pq = psmp.next()
            assert np.array_equal(pi, oi), "idx differs at batch %d" % it
            assert np.array_equal(pq, oq), "quirk differs at batch %d" % it
        assert psmp.cursor == ocur
        psmp.close()


def test_gather_reproduces_materialised_blob(oracle):
    """numpy gather of (idx, quirk) == the data blob the reference would have built (incl. K-1 quirk)."""
    rng = np.random.RandomState(3)
    K = 12
    video_id, shot_off, shot_ids, feat = make_dataset(rng, 80, 4, 20, K)
    smp = oracle.Sampler(video_id, shot_off, shot_ids, feat, K, 16, 5, 10, 60, 50, 6, 100, seed=1)
    saw_quirk_shot = saw_quirk_zero = False
    for _ in range(10):
        idx, quirk, data = smp.next()
        g = feat[idx]                                           # [B,R,K]
        q = quirk >= 0
        g[..., K - 1] = np.where(q, feat[np.maximum(quirk, 0), K - 1], g[..., K - 1])
        g[..., K - 1] = np.where(quirk == -1, 0.0, g[..., K - 1])
        assert np.array_equal(g, data)
        saw_quirk_shot |= bool(q.any()); saw_quirk_zero |= bool((quirk == -1).any())
    assert saw_quirk_shot and saw_quirk_zero
    smp.close()


def test_sampler_rejects_bad_parameters(vvlib):
    video_id, shot_off, shot_ids = ops.synthetic_videos(10, 8)
    with pytest.raises(Exception):
        ops.Sampler(video_id, shot_off, shot_ids, 4, context_size=4)              # even context (CHECK :435)
    with pytest.raises(Exception):
        ops.Sampler(video_id, shot_off, shot_ids, 4, max_buffer_size=5000)        # cannot find 5000 unique shots (CHECK_EQ :346)
    with pytest.raises(Exception):
        ops.Sampler(video_id, shot_off, shot_ids, 4, negative_swap_percentage=100, max_buffer_size=20)
    with pytest.raises(Exception):   # more same-video negatives than negative slots: slot overflow in the reference
        ops.Sampler(video_id, shot_off, shot_ids, 4, num_negative_samples=4, max_same_video_negs=6, max_buffer_size=20)


def test_synthetic_bank_hash_host_matches_numpy(vvlib):
    b = ops.bank_host(5, 16, 1234)
    for r in range(5):
        for c in range(16):
            assert b[r, c] == vvlib.vv_bank_value_host(1234, r, c, 16)
    big = ops.bank_host(64, 256, 1234)
    assert 0.3 < (big > 0).mean() < 0.7 and 0.4 < big.std() < 0.8


# ---- the other context types of the data layer -----------------------------------------------
MODE_CASES = [
    # mode, V, smin, smax, B, C, Nn, P, swap, max_same
    ("pairwise", 200, 1, 12, 16, 2, 10, 50, 50, 0),            # two shots per record, every negative from the buffer
    ("pairwise", 100, 2, 5, 8, 2, 0, 0, 0, 0),                 # no negatives at all
    ("past", 200, 3, 30, 16, 5, 10, 100, 50, 6),               # target = last frame of the sorted window
    ("past", 150, 4, 9, 8, 4, 6, 40, 99, 6),                   # even context size, n == C records
    ("past", 120, 18, 30, 8, 17, 50, 400, 50, 6),
    ("past_continuous", 200, 3, 40, 16, 5, 10, 100, 50, 6),    # random stride, negatives = frames before the window
    ("past_continuous", 150, 2, 8, 8, 2, 4, 30, 50, 2),
    ("past_continuous_fixed", 200, 3, 40, 16, 5, 10, 100, 50, 6),
    ("past_continuous_fixed", 64, 9, 9, 64, 3, 6, 60, 0, 6),
]


@pytest.mark.parametrize("mode,V,smin,smax,B,C,Nn,P,swap,max_same", MODE_CASES)
def test_other_context_types_match_oracle(vvlib, oracle, mode, V, smin, smax, B, C, Nn, P, swap, max_same):
    """PAIRWISE / PAST / PAST_CONTINUOUS / PAST_CONTINUOUS_FIXED (video_sampled_shots_data_layer.cpp:396-422, 509-757):
    the same bit-exact index stream contract as WINDOW, against the oracle on the real libc rand()."""
    from videovector_b200._lib import CONTEXT
    rng = np.random.RandomState(V + B + C)
    video_id, shot_off, shot_ids, feat = make_dataset(rng, V, smin, smax, K=6)
    R = C + Nn
    for seed in (1, 5):
        osmp = oracle.Sampler(video_id, shot_off, shot_ids, feat, 6, B, C, Nn, P, swap, max_same, 100, seed=seed, context_type=CONTEXT[mode])
        ostream = [osmp.next() for _ in range(10)]
        ocur = osmp.cursor
        osmp.close()
        psmp = ops.Sampler(video_id, shot_off, shot_ids, B, C, Nn, P, swap, max_same, 100, rand_seed=seed, context_type=mode)
        prev = np.zeros((B, R, 6), np.float32)                      # the prefetch buffer persists between batches
        for it, (oi, oq, data) in enumerate(ostream):
            pi, pq = psmp.next()
            assert np.array_equal(pi, oi), "%s: idx differs at batch %d" % (mode, it)
            assert np.array_equal(pq, oq), "%s: quirk differs at batch %d" % (mode, it)
            # gathering bank rows with (idx, quirk) rebuilds the reference's data blob, K-1 copy quirk included
            blob = feat[pi]
            last = np.where(pq >= 0, feat[np.maximum(pq, 0), 5], np.where(pq == -1, 0.0, blob[..., 5]))
            blob[..., 5] = last
            assert np.array_equal(blob, data), "%s: data blob differs at batch %d" % (mode, it)
            # structure: the target is the LAST frame of its window in the PAST modes, one video per item
            vid_of = np.searchsorted(shot_off, pi[:, :C], side="right") - 1
            assert (vid_of == vid_of[:, :1]).all()
            if mode.startswith("past"):
                assert (pi[:, 0:1] > pi[:, 1:C]).all() and (np.diff(pi[:, 1:C], axis=1) > 0).all()
            if mode == "past_continuous_fixed" and C > 2:
                d = np.diff(np.concatenate([pi[:, 1:C], pi[:, :1]], 1), axis=1)
                assert (d == d[:, :1]).all()                             # evenly spaced
        assert psmp.cursor == ocur
        psmp.close()


def test_context_type_argument_checks(vvlib):
    rng = np.random.RandomState(0)
    video_id, shot_off, shot_ids, _ = make_dataset(rng, 50, 6, 12)
    with pytest.raises(Exception):
        ops.Sampler(video_id, shot_off, shot_ids, 4, 3, 4, 20, 50, 2, 100, context_type="pairwise")    # PAIRWISE needs C == 2
    with pytest.raises(Exception):
        ops.Sampler(video_id, shot_off, shot_ids, 4, 4, 4, 20, 50, 2, 100, context_type="window")      # WINDOW needs an odd C
    ops.Sampler(video_id, shot_off, shot_ids, 4, 4, 4, 20, 50, 2, 100, context_type="past").close()   # PAST does not
    with pytest.raises(Exception):
        ops.Sampler(video_id, shot_off, shot_ids, 4, 5, 4, 20, 50, 2, 100, context_type=7)


# ---- pinned to the REFERENCE's own data layer ----------------------------------------------------------------------
GOLD_SAMPLER = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden", "sampler_ref.npz")
FIXTURE_CASES = ["window", "window_c3", "past", "past_continuous", "past_continuous_fixed", "pairwise"]


def _blob_from_indices(feat, idx, quirk):
    blob = feat[idx]
    K = feat.shape[1]
    blob[..., K - 1] = np.where(quirk >= 0, feat[np.maximum(quirk, 0), K - 1], np.where(quirk == -1, 0.0, blob[..., K - 1]))
    return blob


@pytest.mark.parametrize("name", FIXTURE_CASES)
def test_sampler_reproduces_reference_data_layer_fixtures(vvlib, oracle, name):
    """tests/golden/sampler_ref.npz holds the data blobs the reference's VideoSampledShotsDataLayer ITSELF produced
    (video_sampled_shots_data_layer.cpp compiled unmodified, fake in-memory LMDB, real libc rand(), seed 1; see
    make_golden.py).  The oracle's sampler and the product's index stream (+ gather) must rebuild them bit for bit."""
    g = np.load(GOLD_SAMPLER)
    mode, B, C, Nn, P, swap, max_same = [int(x) for x in g["cfg_" + name]]
    ref = g["blobs_" + name]
    osmp = oracle.Sampler(g["vid"], g["off"], g["sid"], g["feat"], 5, B, C, Nn, P, swap, max_same, 100, seed=1, context_type=mode)
    for i in range(ref.shape[0]):
        assert np.array_equal(osmp.next()[2], ref[i]), "oracle differs from the reference data layer at batch %d" % i
    osmp.close()
    psmp = ops.Sampler(g["vid"], g["off"], g["sid"], B, C, Nn, P, swap, max_same, 100, rand_seed=1, context_type=mode)
    for i in range(ref.shape[0]):
        idx, quirk = psmp.next()
        assert np.array_equal(_blob_from_indices(g["feat"], idx, quirk), ref[i]), "product differs at batch %d" % i
    psmp.close()


def mt19937_first(seed):
    """First 32-bit output of mt19937 seeded with an integer (boost::mt19937 = std::mt19937): what caffe_rng_rand() returns
    right after Caffe::set_random_seed(seed) (ref: common.cpp:47-50, math_functions.cpp:265-267)."""
    mt = [0] * 624
    mt[0] = seed & 0xFFFFFFFF
    for i in range(1, 624):
        mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
    for i in range(624):                                   # one twist
        y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
        mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
    y = mt[0]
    y ^= y >> 11; y ^= (y << 7) & 0x9D2C5680; y ^= (y << 15) & 0xEFC60000; y ^= y >> 18
    return y & 0xFFFFFFFF


GOLD_SAMPLER_OPTS = GOLD_SAMPLER.replace("sampler_ref.npz", "sampler_opts_ref.npz")


@pytest.mark.parametrize("name", ["skip_window", "neg_window", "skip_neg_past"])
def test_sampler_options_reproduce_reference_data_layer(vvlib, oracle, name):
    """rand_skip and negative_dataset (video_sampled_shots_data_layer.cpp:137-153, 157-180, 273-284, 324-338): the data blobs
    the reference's data layer produced with them (tests/golden/sampler_opts_ref.npz <- make_sampler_opts_golden.py: two
    fake LMDBs, skip drawn from caffe_rng_rand()) rebuilt bit for bit by the oracle's sampler and by the product's index
    stream over a bank that holds the negative set's rows behind the main set's."""
    g = np.load(GOLD_SAMPLER_OPTS)
    mode, B, C, Nn, P, swap, max_same, rand_skip, cseed, with_neg = [int(x) for x in g["cfg_" + name]]
    ref = g["blobs_" + name]
    skip = mt19937_first(cseed) % rand_skip if rand_skip else 0
    bank = np.concatenate([g["feat"], g["nfeat"]]) if with_neg else g["feat"]
    base = len(g["feat"])
    oneg = (g["nvid"], g["noff"], g["nsid"], g["nfeat"], base) if with_neg else None
    osmp = oracle.Sampler(g["vid"], g["off"], g["sid"], g["feat"], g["feat"].shape[1], B, C, Nn, P, swap, max_same, 100, seed=1,
                          context_type=mode, start_skip=skip, negative_dataset=oneg)
    for i in range(ref.shape[0]):
        idx, quirk, data = osmp.next()
        assert np.array_equal(data, ref[i]), "oracle differs from the reference data layer at batch %d" % i
        assert np.array_equal(_blob_from_indices(bank, idx, quirk), ref[i])
    osmp.close()
    pneg = (g["nvid"], g["noff"], g["nsid"], base) if with_neg else None
    psmp = ops.Sampler(g["vid"], g["off"], g["sid"], B, C, Nn, P, swap, max_same, 100, rand_seed=1, context_type=mode,
                       start_skip=skip, negative_dataset=pneg)
    for i in range(ref.shape[0]):
        idx, quirk = psmp.next()
        assert np.array_equal(_blob_from_indices(bank, idx, quirk), ref[i]), "product differs at batch %d" % i
    psmp.close()
    if with_neg:
        # the reference CHECKs that the buffer fills exactly at a record boundary (:346): one slot less is an error
        with pytest.raises(Exception):
            ops.Sampler(g["vid"], g["off"], g["sid"], B, C, Nn, P - 1, swap, max_same, 100, rand_seed=1, context_type=mode,
                        negative_dataset=pneg)
        with pytest.raises(Exception):
            oracle.Sampler(g["vid"], g["off"], g["sid"], g["feat"], g["feat"].shape[1], B, C, Nn, P - 1, swap, max_same, 100, seed=1,
                           context_type=mode, negative_dataset=oneg)


def test_sampler_option_arguments_are_checked(vvlib):
    """start_skip wraps around the record count; a negative dataset excludes sub-shard row offsets (its rows are absolute
    bank rows); malformed option arguments are refused."""
    g = np.load(GOLD_SAMPLER_OPTS)
    mode, B, C, Nn, P = [int(x) for x in g["cfg_neg_window"][:5]]
    V = len(g["vid"])
    a = ops.Sampler(g["vid"], g["off"], g["sid"], B, C, Nn, 40, 50, 4, 100, rand_seed=1, start_skip=3)
    b = ops.Sampler(g["vid"], g["off"], g["sid"], B, C, Nn, 40, 50, 4, 100, rand_seed=1, start_skip=3 + 2 * V)
    for _ in range(3):
        ia, qa = a.next(); ib, qb = b.next()
        assert np.array_equal(ia, ib) and np.array_equal(qa, qb)
    a.close(); b.close()
    with pytest.raises(Exception):
        ops.Sampler(g["vid"], g["off"], g["sid"], B, C, Nn, 40, 50, 4, 100, rand_seed=1, start_skip=-1)
    neg = (g["nvid"], g["noff"], g["nsid"], len(g["feat"]))
    s = ops.Sampler(g["vid"], g["off"], g["sid"], B, C, Nn, P, 50, 4, 100, rand_seed=1, negative_dataset=neg)
    assert vvlib.vv_sampler_set_row_base(s._h, 128) != 0 and vvlib.vv_sampler_set_row_base(s._h, 0) == 0
    idx, _ = s.next()
    assert (idx[:, C:] >= len(g["feat"])).any()          # buffer negatives of the first batch come from the negative set's rows
    s.close()


@pytest.mark.parametrize("mode,C", [(1, 5), (2, 6), (3, 4), (4, 5), (0, 2)])
def test_live_reference_data_layer(vvlib, oracle, mode, C):
    """Where oracle/_ref was built: the reference's data layer run live on a fresh dataset against the product sampler."""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref not built on this machine")
    rng = np.random.RandomState(100 + mode)
    video_id, shot_off, shot_ids, feat = make_dataset(rng, 150, 2, 25, K=4)
    B, Nn, P, max_same = 24, 10, 90, (0 if mode == 0 else 6)
    r = pyref.Sampler(video_id, shot_off, shot_ids, feat, 4, B, C, Nn, P, 50, max_same, seed=3, context_type=mode)
    ref = [r.next() for _ in range(10)]
    r.close()
    psmp = ops.Sampler(video_id, shot_off, shot_ids, B, C, Nn, P, 50, max_same, 100, rand_seed=3, context_type=mode)
    for i, blob in enumerate(ref):
        idx, quirk = psmp.next()
        assert np.array_equal(_blob_from_indices(feat, idx, quirk), blob), "batch %d" % i
    psmp.close()


def test_prefetch_thread_serves_the_same_stream(vvlib):
    """vv_sampler_prefetch (the reference's prefetch thread, base_data_layer.cpp:53-95): same index stream, same cursor per
    batch, stop/restart in the middle loses nothing, destroy with a full ring does not hang."""
    rng = np.random.RandomState(21)
    video_id, shot_off, shot_ids, _ = make_dataset(rng, 120, 2, 30, K=2)
    args = (video_id, shot_off, shot_ids, 32, 5, 10, 80, 50, 6, 100)
    plain = ops.Sampler(*args, rand_seed=1)
    want = []
    for _ in range(40):
        i, q = plain.next(); want.append((i, q, plain.cursor))
    plain.close()
    pf = ops.Sampler(*args, rand_seed=1)
    pf.prefetch(4)
    for k in range(40):
        if k == 13:
            pf.prefetch(0)          # stop: the batches drawn ahead are served first, then inline generation continues
        if k == 22:
            pf.prefetch(3)
        i, q = pf.next()
        assert np.array_equal(i, want[k][0]) and np.array_equal(q, want[k][1]), k
        assert pf.cursor == want[k][2], k
    pf.close()                      # ring full, producer blocked
    with pytest.raises(Exception):
        ops.Sampler(*args, rand_seed=1).prefetch(5000)


def test_multi_stream_sampler_is_k_reference_exact_streams(vvlib):
    """ops.MultiSampler: k sub-shard samplers (each the reference's sampler on its own videos, own rand() stream and
    negative buffer, own prefetch thread), batches round-robin; indices are offset into the common bank."""
    V, S, B, C, Nn = 96, 20, 24, 5, 10
    vid, off, sid = ops.synthetic_videos(V, S)
    for k in (1, 2, 3):
        ms = ops.MultiSampler(vid, off, sid, B, C, Nn, 300, 50, 6, 100, rand_seed=5, streams=k, row_base=1000)
        ms.prefetch(4)
        parts = []
        for s in range(k):
            v0, v1 = s * V // k, (s + 1) * V // k
            parts.append((ops.Sampler(vid[v0:v1], off[v0:v1 + 1] - off[v0], sid[off[v0]:off[v1]], B, C, Nn, 300, 50, 6, 100,
                                      rand_seed=5 + s), 1000 + int(off[v0]), int(off[v0]), int(off[v1])))
        for it in range(3 * k + 1):
            idx, quirk = ms.next()
            smp, base, lo, hi = parts[it % k]
            ri, rq = smp.next()
            assert np.array_equal(idx, ri + base) and np.array_equal(quirk, np.where(rq >= 0, rq + base, rq))
            assert idx.min() >= 1000 + lo and idx.max() < 1000 + hi          # stays inside its sub-shard of the bank
        ms.prefetch(0)
        assert ms.ready >= 0
        ms.close()
        for p in parts:
            p[0].close()
