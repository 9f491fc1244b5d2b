#!/usr/bin/env python
"""Golden outputs of the REFERENCE's RetrievalStatsLayer (oracle/_ref/libvv_ref.so <- retrieval_stats_layer.cpp compiled
unmodified) for its two remaining options: `video_level_retrieval` (the shots of a video are averaged first) and
`stats_output_file` (the per-query CSV).  Run in the build container:
    python tests/golden/make_retrieval_opts_golden.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402

assert pyref.available(), "build oracle/_ref first: bash oracle/ref_shim/build_ref.sh"
OUT = os.path.dirname(os.path.abspath(__file__))
rng = np.random.RandomState(4711)
B, N, ncls, nvid = 83, 24, 4, 17
cls_of_video = rng.randint(0, ncls, nvid)
vids = rng.randint(0, nvid, B).astype(np.float32)
vids[:nvid] = np.arange(nvid)                                    # every video present: max_num_videos = 17
centers = rng.normal(0, 1, (ncls, N)).astype(np.float32)
E = centers[cls_of_video[vids.astype(int)]] * 0.7 + rng.normal(0, 1, (B, N)).astype(np.float32)
E = (E / np.sqrt((E ** 2).sum(1, keepdims=True))).astype(np.float32)
idmap = {v: int(cls_of_video[v]) for v in range(nvid)}
idmap[3] = -1                                                    # an unscored video (label < 0): no CSV line, not in the means
out = dict(E=E, vids=vids, map_ids=np.array(list(idmap.keys()), np.int32), map_cls=np.array(list(idmap.values()), np.int32),
           max_num_videos=nvid)
with tempfile.TemporaryDirectory() as d:
    idf = os.path.join(d, "id2class.txt")
    open(idf, "w").write("".join("%d,%d\n" % kv for kv in idmap.items()))
    for excl in (0, 1):
        csv_shot = os.path.join(d, "shot_%d.csv" % excl); csv_video = os.path.join(d, "video_%d.csv" % excl)
        out["shot_out_%d" % excl] = pyref.retrieval_stats_ex(E, vids, idf, bool(excl), stats_output_file=csv_shot)
        out["video_out_%d" % excl] = pyref.retrieval_stats_ex(E, vids, idf, bool(excl), video_level=True, max_num_videos=nvid,
                                                              stats_output_file=csv_video)
        out["shot_csv_%d" % excl] = np.frombuffer(open(csv_shot, "rb").read(), np.uint8)
        out["video_csv_%d" % excl] = np.frombuffer(open(csv_video, "rb").read(), np.uint8)
        print("exclude", excl, "shot", out["shot_out_%d" % excl], "video", out["video_out_%d" % excl])
print(bytes(out["shot_csv_1"][:300]).decode())
print(bytes(out["video_csv_1"][:200]).decode())
np.savez_compressed(os.path.join(OUT, "retrieval_opts.npz"), **out)
print("retrieval_opts.npz written")
