#!/usr/bin/env python
"""Golden data blobs for the data layer's `rand_skip` and `negative_dataset` options, produced by the REFERENCE's own
VideoSampledShotsDataLayer (oracle/_ref/libvv_ref.so <- video_sampled_shots_data_layer.cpp compiled unmodified, two fake
in-memory LMDBs, real libc rand(), caffe_rng_rand() for the skip).  Run in the build container:
    python tests/golden/make_sampler_opts_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402

assert pyref.available(), "build oracle/_ref first: bash oracle/ref_shim/build_ref.sh"
OUT = os.path.dirname(os.path.abspath(__file__))
rng = np.random.RandomState(77)
K = 8                                            # the device gather takes feature sizes that are multiples of 4


def dataset(V, lo, hi, vid0):
    counts = rng.randint(lo, hi, V)
    vid = (rng.permutation(V) + vid0).astype(np.int32)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    sid = np.concatenate([np.sort(rng.choice(60, c, replace=False)) for c in counts]).astype(np.int32)
    feat = rng.normal(0, 1, (off[-1], K)).astype(np.float32)
    return vid, off, sid, feat


vid, off, sid, feat = dataset(40, 6, 20, 100)
nvid, noff, nsid, nfeat = dataset(12, 4, 10, 500)
# one record of the negative set carries the SAME keys as a record of the main set (video id and shot ids): those keys
# are then in the buffer's key set from the start and the main data's swap step must see them as present
nvid[2] = vid[5]
n2 = noff[3] - noff[2]
nsid[noff[2]:noff[3]] = sid[off[5]:off[5] + n2]
P = int(noff[7])                                 # the buffer fills exactly at the end of the 7th negative record
out = dict(vid=vid, off=off, sid=sid, feat=feat, nvid=nvid, noff=noff, nsid=nsid, nfeat=nfeat)
B, Nn = 10, 6
CASES = {   # name: context_type, C, max_same, rand_skip, caffe_seed, with negative_dataset, max_buffer_size
    "skip_window": (1, 5, 4, 7, 5, False, 50),
    "neg_window": (1, 5, 4, 0, 0, True, P),
    "skip_neg_past": (2, 4, 3, 11, 9, True, P),
}
for name, (mode, C, max_same, rand_skip, cseed, with_neg, bufsize) in CASES.items():
    r = pyref.Sampler(vid, off, sid, feat, K, B, C, Nn, bufsize, 50, max_same, seed=1, context_type=mode,
                      rand_skip=rand_skip, caffe_seed=cseed, negative_dataset=(nvid, noff, nsid, nfeat) if with_neg else None)
    out["blobs_" + name] = np.stack([r.next() for _ in range(8)])
    out["cfg_" + name] = np.array([mode, B, C, Nn, bufsize, 50, max_same, rand_skip, cseed, int(with_neg)], np.int64)
    r.close()
    print(name, out["blobs_" + name].shape)
np.savez_compressed(os.path.join(OUT, "sampler_opts_ref.npz"), **out)
print("sampler_opts_ref.npz written; buffer size for the negative-dataset cases:", P)
