#!/usr/bin/env python
"""Golden DB records for the record reader (vv_record_set_*), serialised by the REAL protobuf runtime (python
`google.protobuf`) from the reference's own schemas (caffe.proto + video_shot_sentences.proto parsed into a
FileDescriptorSet, committed as records_schema.desc so tests need no reference tree):
  * video_shots.vvrs        the sampler fixture dataset of sampler_ref.npz (vid/off/sid/feat) as one VideoShots record
                            per video, in DB key order -> together with sampler_ref.npz's data blobs (produced by the
                            reference's compiled data layer) this pins records -> bank -> sampler -> data blob end to end
  * video_shots_packed.vvrs the same with float_data / shot_ids declared [packed = true] (a parser must take both)
  * video_shots.mdbdump / video_shots_p.mdbdump   the same records in the text `mdb_dump` / `mdb_dump -p` prints
                            (written here from the format's description: liblmdb is not in this image)
  * test_windows.vvrs + test_windows.npz          TestVideoShotWindows records and the arrays they hold
Run here (needs /root/reference)."""
import os, struct, sys
import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_caffemodel_golden import build_descriptor        # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def schema(packed=False):
    caffe = build_descriptor(open(os.path.join(REF, "src/caffe/proto/caffe.proto")).read())
    vss = build_descriptor(open(os.path.join(REF, "src/caffe/proto/video_shot_sentences.proto")).read().replace('import "caffe.proto";', ""),
                           package="video_shot_sentences", name="video_shot_sentences.proto", deps=["caffe.proto"],
                           external=("caffe.Datum",))
    if packed:
        for f in [m for m in caffe.message_type if m.name == "Datum"][0].field:
            if f.name == "float_data":
                f.options.packed = True
        for m in vss.message_type:
            for f in m.field:
                if f.label == f.LABEL_REPEATED and f.type == f.TYPE_INT32:
                    f.options.packed = True
    return caffe, vss


def classes(caffe, vss):
    pool = descriptor_pool.DescriptorPool()
    pool.Add(caffe); pool.Add(vss)
    return lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("video_shot_sentences." + n))


def write_vvrs(path, records):
    with open(path, "wb") as f:
        f.write(b"VVRS0001")
        for key, val in records:
            f.write(struct.pack("<I", len(key))); f.write(key); f.write(struct.pack("<Q", len(val))); f.write(val)


def write_mdb_dump(path, records, printable):
    def enc(b):
        if not printable:
            return b.hex()
        return "".join("\\\\" if c == 0x5c else chr(c) if 0x20 <= c < 0x7f else "\\%02x" % c for c in b)
    with open(path, "w") as f:
        f.write("VERSION=3\nformat=%s\ntype=btree\nmapsize=1099511627776\nmaxreaders=126\ndb_pagesize=4096\nHEADER=END\n" % ("print" if printable else "bytevalue"))
        for key, val in records:
            f.write(" " + enc(key) + "\n " + enc(val) + "\n")
        f.write("DATA=END\n")


def main():
    caffe, vss = schema()
    fds = descriptor_pb2.FileDescriptorSet(); fds.file.add().CopyFrom(caffe); fds.file.add().CopyFrom(vss)
    open(os.path.join(HERE, "records_schema.desc"), "wb").write(fds.SerializeToString())
    g = np.load(os.path.join(HERE, "sampler_ref.npz"))
    vid, off, sid, feat = g["vid"], g["off"], g["sid"], g["feat"]
    for tag, packed in (("", False), ("_packed", True)):
        VideoShots = classes(*schema(packed))("VideoShots")
        recs = []
        for v in range(len(vid)):
            m = VideoShots(); m.video_id = int(vid[v]); m.video_name = "video_%d.mp4" % vid[v]
            for r in range(off[v], off[v + 1]):
                m.shot_ids.append(int(sid[r]))
                d = m.shot_words.add(); d.channels, d.height, d.width = feat.shape[1], 1, 1      # as the creation tools set them
                d.float_data.extend(feat[r].tolist())
            recs.append((b"%08d" % v, m.SerializeToString()))
        write_vvrs(os.path.join(HERE, "video_shots%s.vvrs" % tag), recs)
        if not packed:
            write_mdb_dump(os.path.join(HERE, "video_shots.mdbdump"), recs, False)
            write_mdb_dump(os.path.join(HERE, "video_shots_p.mdbdump"), recs, True)
    # TEST-phase records: 4 context frames, 1 positive, 2 negatives per item
    Test = classes(*schema())("TestVideoShotWindows")
    rng = np.random.RandomState(5)
    n, F, P, Ng, K = 23, 4, 1, 2, 6
    data = rng.normal(0, 1, (n, F + P + Ng, K)).astype(np.float32)
    vids = rng.randint(0, 9, n).astype(np.int32); pos_id = rng.randint(0, 50, (n, P)).astype(np.int32); neg_id = rng.randint(0, 50, (n, Ng)).astype(np.int32)
    recs = []
    for i in range(n):
        m = Test(); m.video_id = int(vids[i]); m.video_name = "v%d" % vids[i]
        for j in range(F):
            m.context_shot_words.add().float_data.extend(data[i, j].tolist())
        for j in range(P):
            m.positive_shot_id.append(int(pos_id[i, j])); m.positive_shot_words.add().float_data.extend(data[i, F + j].tolist())
        for j in range(Ng):
            m.negative_shot_id.append(int(neg_id[i, j])); m.negative_shot_words.add().float_data.extend(data[i, F + P + j].tolist())
        recs.append((b"%08d" % i, m.SerializeToString()))
    write_vvrs(os.path.join(HERE, "test_windows.vvrs"), recs)
    # what the reference's own VideoShotWindowTestDataLayer (oracle/_ref, compiled unmodified, fake LMDB) serves from these
    # records: 4 batches of 10 (wraps after 23 records) for the three include_positives / include_negatives settings
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import pyref
    ref = {}
    for tag, ip, ineg in (("pn", True, True), ("p", True, False), ("ctx", False, False)):
        lay = pyref.TestLayer(data, vids, pos_id, neg_id, F, P, Ng, 10, ip, ineg)
        out = [lay.next() for _ in range(4)]
        ref["blob_" + tag] = np.stack([o[0] for o in out]); ref["label_" + tag] = np.stack([o[1] for o in out])
        lay.close()
    np.savez_compressed(os.path.join(HERE, "test_windows.npz"), data=data, vids=vids, pos_id=pos_id, neg_id=neg_id, **ref)
    print("wrote", [f for f in sorted(os.listdir(HERE)) if "video_shots" in f or "test_windows" in f or f == "records_schema.desc"])


if __name__ == "__main__":
    main()
