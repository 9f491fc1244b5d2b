"""Record reader (vv_record_set_*): the reference's VideoShots / TestVideoShotWindows DB values ->
feature bank + sampler tables.  The golden records were serialised by the real protobuf runtime from the reference's
schemas (tests/golden/make_records_golden.py); the dataset is the one the reference's compiled data layer produced
tests/golden/sampler_ref.npz from, so records -> bank -> sampler -> data blob is pinned end to end."""
import os
import numpy as np
import pytest
from videovector_b200 import ops
from videovector_b200._lib import VVError

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(autouse=True)
def _built(vvlib):
    return vvlib


def _classes():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fds = descriptor_pb2.FileDescriptorSet()
    fds.ParseFromString(open(os.path.join(GOLD, "records_schema.desc"), "rb").read())
    pool = descriptor_pool.DescriptorPool()
    for f in fds.file:
        pool.Add(f)
    return lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("video_shot_sentences." + n))


@pytest.mark.parametrize("fname", ["video_shots.vvrs", "video_shots_packed.vvrs", "video_shots.mdbdump", "video_shots_p.mdbdump"])
def test_video_shots_records_decode_to_the_fixture_dataset(fname):
    g = np.load(os.path.join(GOLD, "sampler_ref.npz"))
    rs = ops.RecordSet("video_shots").load_file(os.path.join(GOLD, fname))
    info = rs.info()
    assert info == dict(records=len(g["vid"]), rows=int(g["off"][-1]), feature_size=g["feat"].shape[1], rows_per_record=0)
    vid, off, sid = rs.tables()
    assert np.array_equal(vid, g["vid"]) and np.array_equal(off, g["off"]) and np.array_equal(sid, g["sid"])
    assert np.array_equal(rs.bank_host().view(np.uint32), g["feat"].view(np.uint32))
    rs.close()


@pytest.mark.parametrize("name", ["window", "past", "past_continuous", "pairwise"])
def test_records_to_data_blob_matches_reference_data_layer(name):
    """wire bytes -> RecordSet -> Sampler -> gathered blob == the blob the reference's VideoSampledShotsDataLayer built."""
    g = np.load(os.path.join(GOLD, "sampler_ref.npz"))
    mode, B, C, Nn, P, swap, max_same = [int(x) for x in g["cfg_" + name]]
    rs = ops.RecordSet().load_file(os.path.join(GOLD, "video_shots.vvrs"))
    vid, off, sid = rs.tables(); feat = rs.bank_host()
    smp = ops.Sampler(vid, off, sid, B, C, Nn, P, swap, max_same, 100, rand_seed=1, context_type=mode)
    K = feat.shape[1]
    for i, ref in enumerate(g["blobs_" + name]):
        idx, quirk = smp.next()
        blob = feat[idx]
        blob[..., K - 1] = np.where(quirk >= 0, feat[np.maximum(quirk, 0), K - 1], np.where(quirk == -1, 0.0, blob[..., K - 1]))
        assert np.array_equal(blob, ref), "batch %d" % i
    smp.close(); rs.close()


@pytest.mark.parametrize("pos,neg", [(True, True), (True, False), (False, False)])
def test_test_windows_records(pos, neg):
    t = np.load(os.path.join(GOLD, "test_windows.npz"))
    data, F = t["data"], 4
    rs = ops.RecordSet("test_windows", include_positives=pos, include_negatives=neg).load_file(os.path.join(GOLD, "test_windows.vvrs"))
    rows = F + (1 if pos else 0) + (2 if neg else 0)
    assert rs.info() == dict(records=len(data), rows=len(data) * rows, feature_size=data.shape[2], rows_per_record=rows)
    vid, off, sid = rs.tables()
    assert np.array_equal(vid, t["vids"]) and np.array_equal(off, np.arange(len(data) + 1) * rows)
    keep = list(range(F)) + ([F] if pos else []) + ([F + 1, F + 2] if neg else [])
    assert np.array_equal(rs.bank_host().reshape(len(data), rows, -1), data[:, keep])
    want_ids = np.concatenate([np.full((len(data), F), -1, np.int32)] + ([t["pos_id"]] if pos else []) + ([t["neg_id"]] if neg else []), axis=1)
    assert np.array_equal(sid.reshape(len(data), rows), want_ids)


def test_live_records_unknown_fields_negative_ids_and_cut_datums():
    get = _classes()
    VideoShots = get("VideoShots")
    rng = np.random.RandomState(3)
    rs = ops.RecordSet()
    want = []
    for v, n in enumerate([3, 1, 6, 0, 2]):
        m = VideoShots(); m.video_id = -7 if v == 2 else v * 1000003       # negative int32 = 10-byte varint
        for s in range(n):
            m.shot_ids.append(-s if v == 2 else s * 5)
            d = m.shot_words.add()
            d.channels, d.height, d.width, d.label = 4, 1, 1, 9; d.data = b"\x00\x01\x02"; d.mean.extend([1.0, 2.0])
            f = rng.normal(0, 1, 4 + (v == 1)).astype(np.float32)          # video 1 carries an extra float: cut to feature_size
            d.float_data.extend(f.tolist()); want.append(f[:4])
        m.shot_ids.extend([77] * (v == 4))                                  # more ids than shots: the surplus is ignored
        raw = m.SerializeToString()
        if v == 0:
            raw += bytes([0x9b, 0x06, 0x08, 0x01, 0x9c, 0x06]) + bytes([0xa1, 0x06]) + b"\x00" * 8   # unknown group 99 and fixed64 100
        rs.add(raw)
    info = rs.info()
    assert info["records"] == 5 and info["rows"] == 12 and info["feature_size"] == 4
    vid, off, sid = rs.tables()
    assert list(vid) == [0, 1000003, -7, 3000009, 4000012] and list(off) == [0, 3, 4, 10, 10, 12]
    assert list(sid[4:10]) == [0, -1, -2, -3, -4, -5]
    assert np.array_equal(rs.bank_host(), np.stack(want))


def test_malformed_records_fail_loudly():
    get = _classes()
    VideoShots = get("VideoShots")
    m = VideoShots(); m.video_id = 1
    for s in range(3):
        m.shot_ids.append(s); m.shot_words.add().float_data.extend([1.0, 2.0, 3.0])
    good = m.SerializeToString()
    rs = ops.RecordSet()
    with pytest.raises(VVError, match="no shot_words"):
        rs.add(VideoShots(video_id=4).SerializeToString())
    rs.add(good)
    for cut in (len(good) - 1, len(good) - 7, 5):
        with pytest.raises(VVError, match="malformed|truncated"):
            rs.add(good[:cut])
    short = VideoShots(); short.shot_ids.append(0); short.shot_words.add().float_data.extend([1.0, 2.0])
    with pytest.raises(VVError, match="feature_size is 3"):
        rs.add(short.SerializeToString())
    few = VideoShots(); few.shot_words.add().float_data.extend([1.0, 2.0, 3.0])
    with pytest.raises(VVError, match="1 shot_words but 0 shot_ids"):
        rs.add(few.SerializeToString())
    assert rs.info()["records"] == 1 and rs.info()["rows"] == 3            # failed adds leave the set unchanged
    assert np.array_equal(rs.bank_host(), np.tile(np.array([1, 2, 3], np.float32), (3, 1)))
    Test = get("TestVideoShotWindows")
    ts = ops.RecordSet("test_windows")
    t = Test(); t.context_shot_words.add().float_data.extend([1.0])
    with pytest.raises(VVError, match="No video id"):
        ts.add(t.SerializeToString())
    t.video_id = 3; ts.add(t.SerializeToString())
    t.context_shot_words.add().float_data.extend([1.0])
    with pytest.raises(VVError, match="2 context words, expected 1"):
        ts.add(t.SerializeToString())
    # sizes come from the first ACCEPTED record: a rejected first record leaves nothing behind
    ts2 = ops.RecordSet("test_windows")
    bad = Test(); bad.video_id = 1
    bad.context_shot_words.add().float_data.extend([1.0, 2.0]); bad.context_shot_words.add().float_data.extend([1.0])
    with pytest.raises(VVError, match="feature_size is 2"):
        ts2.add(bad.SerializeToString())
    good3 = Test(); good3.video_id = 2
    for _ in range(3):
        good3.context_shot_words.add().float_data.extend([4.0, 5.0, 6.0])
    ts2.add(good3.SerializeToString())
    assert ts2.info() == dict(records=1, rows=3, feature_size=3, rows_per_record=3)


def test_load_file_errors(tmp_path):
    rs = ops.RecordSet()
    with pytest.raises(VVError, match="cannot open"):
        rs.load_file(tmp_path / "missing.vvrs")
    p = tmp_path / "junk.bin"; p.write_bytes(b"not a record file at all")
    with pytest.raises(VVError, match="not an LMDB environment"):
        rs.load_file(p)
    raw = open(os.path.join(GOLD, "video_shots.vvrs"), "rb").read()
    q = tmp_path / "cut.vvrs"; q.write_bytes(raw[:len(raw) - 11])
    with pytest.raises(VVError, match="truncated"):
        ops.RecordSet().load_file(q)


def _read_vvrs(path):
    import struct
    raw = open(path, "rb").read(); at = 8; out = []
    while at < len(raw):
        kl, = struct.unpack_from("<I", raw, at); key = raw[at + 4:at + 4 + kl]; at += 4 + kl
        vl, = struct.unpack_from("<Q", raw, at); out.append((key, raw[at + 8:at + 8 + vl])); at += 8 + vl
    return out


def test_lmdb_environment_directory(tmp_path):
    """`source:` naming an LMDB directory: data.mdb's B+tree is walked in key order (container layout unpinned: no liblmdb
    here, the file comes from tests/lmdb_writer.py)."""
    from lmdb_writer import write_lmdb
    g = np.load(os.path.join(GOLD, "sampler_ref.npz"))
    recs = _read_vvrs(os.path.join(GOLD, "video_shots.vvrs"))
    st = write_lmdb(str(tmp_path / "small_lmdb"), reversed(recs))
    assert st["leaf"] > 1 and st["branch"] >= 1
    rs = ops.RecordSet().load_file(tmp_path / "small_lmdb")
    vid, off, sid = rs.tables()
    assert np.array_equal(vid, g["vid"]) and np.array_equal(off, g["off"]) and np.array_equal(sid, g["sid"])
    assert np.array_equal(rs.bank_host(), g["feat"])
    # many records, three tree levels, values on overflow pages
    VideoShots = _classes()("VideoShots")
    rng = np.random.RandomState(9)
    big, want = [], []
    for v in range(1500):
        m = VideoShots(); m.video_id = v
        for s_ in range(int(rng.randint(1, 4)) if v % 50 else 40):
            f = rng.normal(0, 1, 24).astype(np.float32)
            m.shot_ids.append(s_); m.shot_words.add().float_data.extend(f.tolist()); want.append(f)
        big.append((b"%08d_some_longer_key_to_fill_branch_pages_%04d" % (v, v), m.SerializeToString()))
    st = write_lmdb(str(tmp_path / "big_lmdb"), big)
    assert st["depth"] >= 3 and st["overflow"] > 0
    rs2 = ops.RecordSet().load_file(tmp_path / "big_lmdb")
    assert rs2.info()["records"] == 1500 and np.array_equal(rs2.tables()[0], np.arange(1500))
    assert np.array_equal(rs2.bank_host(), np.stack(want))
    with pytest.raises(VVError, match="no data.mdb"):
        ops.RecordSet().load_file(tmp_path)
    raw = open(tmp_path / "big_lmdb" / "data.mdb", "rb").read()
    (tmp_path / "cut_lmdb").mkdir(); open(tmp_path / "cut_lmdb" / "data.mdb", "wb").write(raw[:len(raw) // 2])
    with pytest.raises(VVError, match="shorter than its last page"):
        ops.RecordSet().load_file(tmp_path / "cut_lmdb")


@pytest.mark.parametrize("tag,pos,neg", [("pn", True, True), ("p", True, False), ("ctx", False, False)])
def test_test_windows_match_reference_test_data_layer(tag, pos, neg):
    """tests/golden/test_windows.npz also holds what the reference's VideoShotWindowTestDataLayer ITSELF (compiled
    unmodified into oracle/_ref, fake LMDB) served from these records: item c of the stream = record c mod n, channels =
    its rows, label = its video_id.  The product's decoded record set must reproduce blobs and labels bit for bit; where
    oracle/_ref exists the layer is also run live."""
    t = np.load(os.path.join(GOLD, "test_windows.npz"))
    rs = ops.RecordSet("test_windows", include_positives=pos, include_negatives=neg).load_file(os.path.join(GOLD, "test_windows.vvrs"))
    info = rs.info(); n, rows = info["records"], info["rows_per_record"]
    bank = rs.bank_host().reshape(n, rows, -1); vid = rs.tables()[0]
    blobs, labels = t["blob_" + tag], t["label_" + tag]
    B = blobs.shape[1]
    assert blobs.shape[2] == rows
    for it in range(blobs.shape[0]):
        item = (np.arange(B) + it * B) % n
        assert np.array_equal(bank[item], blobs[it]) and np.array_equal(vid[item].astype(np.float32), labels[it])
    from oracle import pyref
    if pyref.available():
        lay = pyref.TestLayer(t["data"], t["vids"], t["pos_id"], t["neg_id"], 4, 1, 2, B, pos, neg)
        for it in range(3):
            d, l = lay.next()
            assert np.array_equal(d, blobs[it]) and np.array_equal(l, labels[it])
        lay.close()
