#!/usr/bin/env python
"""Golden vectors for MaxMarginLoss with per-video weights (its optional third bottom), produced by the REFERENCE's own
layer (oracle/_ref/libvv_ref.so <- /root/reference/src/caffe/layers/max_margin_loss_layer.cpp) in both forms: direct
weights (use_direct_weight) and video ids looked up in an id_to_weight_file.  Run in the build container:
    python tests/golden/make_mm_weights_golden.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402

assert pyref.available(), "build oracle/_ref first: bash oracle/ref_shim/build_ref.sh"
OUT = os.path.dirname(os.path.abspath(__file__))
rng = np.random.RandomState(424242)
num, ch = 12, 10
t = rng.normal(0, 1, (num, ch)).astype(np.float32)
s = rng.normal(0, 1, (num, ch)).astype(np.float32)
w = rng.uniform(0, 3, (num, ch)).astype(np.float32)
w[0, :3] = 0.0                                           # zero weights: no loss, no gradient
# video ids per score (every column of an item carries the item's video id, as a tiled id blob would); id 1000 is not
# in the table (the reference's map default-inserts weight 0), id 7 appears twice in the file (first line wins)
ids = np.repeat(rng.randint(0, 9, (num, 1)), ch, axis=1).astype(np.float32)
ids[3] = 1000
table_lines = [(int(i), float(np.float32(rng.uniform(0, 2)))) for i in range(9)] + [(7, 5.0)]
out = dict(t=t, s=s, w=w, ids=ids, table_ids=np.array([a for a, _ in table_lines], np.int32),
           table_w=np.array([b for _, b in table_lines], np.float32), margin=np.float32(1.5), loss_weight=np.float32(0.7))
with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
    f.write("".join("%d,%.9g\n" % kv for kv in table_lines))
for norm in (1, 2):
    loss, viol, dt, dbg = pyref.max_margin_weighted(t, s, w, True, "", margin=1.5, norm=norm, loss_weight=0.7)
    out.update({"direct_loss%d" % norm: loss, "direct_viol%d" % norm: viol, "direct_dt%d" % norm: dt, "direct_db%d" % norm: dbg})
    loss, viol, dt, dbg = pyref.max_margin_weighted(t, s, ids, False, f.name, margin=1.5, norm=norm, loss_weight=0.7)
    out.update({"table_loss%d" % norm: loss, "table_viol%d" % norm: viol, "table_dt%d" % norm: dt, "table_db%d" % norm: dbg})
    print("norm", norm, "direct loss", out["direct_loss%d" % norm], "table loss", loss, "violations", viol)
os.unlink(f.name)
np.savez_compressed(os.path.join(OUT, "mm_weights.npz"), **out)
print("mm_weights.npz written")
