"""Test helpers: build VideoShots / TestVideoShotWindows records with the real protobuf runtime (schema from
tests/golden/records_schema.desc, derived from the reference's .proto files) and write them as a VVRS stream."""
import os, struct

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def classes():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fds = descriptor_pb2.FileDescriptorSet()
    fds.ParseFromString(open(os.path.join(GOLD, "records_schema.desc"), "rb").read())
    pool = descriptor_pool.DescriptorPool()
    for f in fds.file:
        pool.Add(f)
    return lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("video_shot_sentences." + n))


def video_shots_records(vid, off, sid, feat):
    VideoShots = classes()("VideoShots")
    recs = []
    for v in range(len(vid)):
        m = VideoShots(); m.video_id = int(vid[v]); m.video_name = "video_%d" % vid[v]
        for r in range(off[v], off[v + 1]):
            m.shot_ids.append(int(sid[r])); m.shot_words.add().float_data.extend(feat[r].tolist())
        recs.append((b"%08d" % v, m.SerializeToString()))
    return recs


def test_window_records(data, vids):
    """data [n, frames, K] -> one record per item, context frames only (what the shipped TEST graph slices)."""
    Test = classes()("TestVideoShotWindows")
    recs = []
    for i in range(len(data)):
        m = Test(); m.video_id = int(vids[i])
        for fr in data[i]:
            m.context_shot_words.add().float_data.extend(fr.tolist())
        recs.append((b"%08d" % i, m.SerializeToString()))
    return recs


def write_vvrs(path, records):
    with open(path, "wb") as f:
        f.write(b"VVRS0001")
        for key, val in records:
            f.write(struct.pack("<I", len(key))); f.write(key); f.write(struct.pack("<Q", len(val))); f.write(val)
    return str(path)
