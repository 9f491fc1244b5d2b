#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own layer code
(oracle/_ref/libvv_ref.so, built by oracle/ref_shim/build_ref.sh from /root/reference) on seeded inputs.
Run in the build container (the reference cannot travel): python tests/golden/make_golden.py
Inputs are stored with the outputs so every consumer (oracle, CUDA path) replays exactly the same case;
the dropout mask is the one the reference's DropoutLayer drew."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402

assert pyref.available(), "build oracle/_ref first: bash oracle/ref_shim/build_ref.sh"
OUT = os.path.dirname(os.path.abspath(__file__))
CASES = {            # name: B, C, Nn, K, N, norm, dropout
    "net_cfg1_small": (8, 5, 10, 256, 64, 2, 0.9),      # the shipped structure (window +-2, 10 negatives, dropout 0.9, L2 margin 2)
    "net_cfg4_small": (3, 17, 50, 128, 128, 2, 0.9),    # large-window variant
    "net_l1_nodrop": (5, 3, 4, 64, 32, 1, 0.0),         # L1 hinge, no dropout layer
}
for name, (B, C, Nn, K, N, norm, ratio) in CASES.items():
    rng = np.random.RandomState(abs(hash(name)) % (2 ** 31))
    rng = np.random.RandomState(sum(map(ord, name)))
    R = C + Nn
    data = np.maximum(rng.normal(0, 1, (B, R, K)), 0).astype(np.float32)
    data[0, 0] = 0.0                                   # an all-zero target row (dropout can do this too)
    W = rng.normal(0, 0.05, (N, K)).astype(np.float32)
    b = rng.normal(0, 0.05, N).astype(np.float32)
    r = pyref.net_forward_backward(data, W, b, B, C, Nn, margin=2.0, norm=norm, dropout_ratio=ratio, seed=1701)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), B=B, C=C, Nn=Nn, K=K, N=N, norm=norm, dropout_ratio=ratio, margin=2.0,
                        data=data, W=W, b=b, mask=r["mask"].astype(np.uint8), loss=r["loss"], violations=r["violations"],
                        dW=r["dW"], db=r["db"], H=r["H"], dZ=r["dZ"], target_score=r["target_score"], neg_score=r["neg_score"])
    print(name, "loss", r["loss"][0], "violations", r["violations"][0], "kept", r["mask"].mean())
# single layers
rng = np.random.RandomState(1701)
x = rng.normal(0, 1, (6, 40)).astype(np.float32); x[2] = 0; dy = rng.normal(0, 1, (6, 40)).astype(np.float32)
y, dx = pyref.normalization(x, dy)
t = rng.normal(0, 10, (10, 5)).astype(np.float32); s = rng.normal(0, 10, (10, 5)).astype(np.float32)
mm = {}
for norm in (1, 2):
    loss, viol, dt, dbg = pyref.max_margin(t, s, margin=1.0, norm=norm, loss_weight=1.0)
    mm["loss%d" % norm] = loss; mm["viol%d" % norm] = viol; mm["dt%d" % norm] = dt; mm["db%d" % norm] = dbg
X = rng.uniform(0, 1, (7, 60)).astype(np.float32); W = rng.uniform(-1, 1, (10, 60)).astype(np.float32)
bb = rng.uniform(1, 2, 10).astype(np.float32); dZ = rng.normal(0, 1, (7, 10)).astype(np.float32)
Z, dW, db, dX = pyref.inner_product(X, W, bb, dZ, regularization=0.5)
np.savez_compressed(os.path.join(OUT, "layers.npz"), norm_x=x, norm_dy=dy, norm_y=y, norm_dx=dx, mm_t=t, mm_s=s,
                    ip_X=X, ip_W=W, ip_b=bb, ip_dZ=dZ, ip_Z=Z, ip_dW=dW, ip_db=db, ip_dX=dX, **mm)
print("layers.npz written")

# TEST-phase + embedding-table layers: RetrievalStatsLayer and IdToWeightMappingLayer of the reference
import tempfile
rng = np.random.RandomState(2015)
Bq, Nq, ncls = 97, 24, 4
cls_of_video = rng.randint(0, ncls, 40)
vids = rng.randint(0, 40, Bq).astype(np.float32)
centers = rng.normal(0, 1, (ncls, Nq)).astype(np.float32)
E = centers[cls_of_video[vids.astype(int)]] * 0.8 + rng.normal(0, 1, (Bq, Nq)).astype(np.float32)
E = (E / np.sqrt((E ** 2).sum(1, keepdims=True))).astype(np.float32)
idmap = {v: int(cls_of_video[v]) for v in range(40)}
idmap[7] = -1                                                   # an unscored video (label < 0)
with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
    f.write("".join("%d,%d\n" % kv for kv in idmap.items()))
rs = {}
for excl in (0, 1):
    rs["rs_out_%d" % excl] = pyref.retrieval_stats(E, vids, f.name, bool(excl))
os.unlink(f.name)
table = rng.normal(0, 1, (11, 9)).astype(np.float32)
ids = rng.randint(0, 11, 30).astype(np.float32); ids[:8] = 3
tdiff = rng.normal(0, 1, (30, 9)).astype(np.float32)
top, tgrad = pyref.id_to_weight(table, ids, tdiff)
np.savez_compressed(os.path.join(OUT, "eval_layers.npz"), rs_E=E, rs_vids=vids, rs_map_keys=np.array(list(idmap.keys()), np.int32),
                    rs_map_vals=np.array(list(idmap.values()), np.int32), id_table=table, id_ids=ids, id_tdiff=tdiff, id_top=top,
                    id_tgrad=tgrad, **rs)
print("eval_layers.npz written", rs)

# The data layer itself: the reference's VideoSampledShotsDataLayer (compiled unmodified, fake in-memory LMDB, real libc
# rand()) for every context type -> data blobs of the first batches.  Pins row S (the sampler).
rng = np.random.RandomState(77)
counts = rng.randint(2, 19, size=90)
s_vid = (rng.permutation(90) + 500).astype(np.int32)
s_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
s_sid = np.concatenate([np.sort(rng.choice(400, c, replace=False)) for c in counts]).astype(np.int32)
s_feat = rng.normal(0, 1, (s_off[-1], 5)).astype(np.float32)
samp = dict(vid=s_vid, off=s_off, sid=s_sid, feat=s_feat)
CASES = [("window", 1, 5, 10, 6), ("window_c3", 1, 3, 4, 2), ("past", 2, 4, 6, 4), ("past_continuous", 3, 5, 8, 6),
         ("past_continuous_fixed", 4, 3, 6, 6), ("pairwise", 0, 2, 6, 0)]
for name, mode, C, Nn, max_same in CASES:
    r = pyref.Sampler(s_vid, s_off, s_sid, s_feat, 5, 12, C, Nn, 70, 50, max_same, seed=1, context_type=mode)
    samp["blobs_" + name] = np.stack([r.next() for _ in range(8)])
    samp["cfg_" + name] = np.array([mode, 12, C, Nn, 70, 50, max_same], np.int32)
    r.close()
np.savez_compressed(os.path.join(OUT, "sampler_ref.npz"), **samp)
print("sampler_ref.npz written:", [k for k in samp if k.startswith("blobs_")])


# The whole reference training pipeline: VideoSampledShotsDataLayer + Net (net.cpp, insert_splits.cpp) + SGDSolver
# (solver.cpp), compiled unmodified, 8 iterations of Solver::Solve's loop body on the shipped TRAIN graph (no dropout).
# Pins the trajectory: per-iteration loss / violations, final weights, bias and momentum histories.
rng = np.random.RandomState(91)
V = 64; counts = rng.randint(8, 22, V)
t_vid = (rng.permutation(V) + 40).astype(np.int32); t_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
t_sid = np.concatenate([np.sort(rng.choice(100, c, replace=False)) for c in counts]).astype(np.int32)
TK, TN, TB = 128, 64, 16
t_feat = np.maximum(rng.normal(0, 1, (t_off[-1], TK)), 0).astype(np.float32)
t_W0 = rng.normal(0, 0.03, (TN, TK)).astype(np.float32); t_b0 = rng.normal(0, 0.01, TN).astype(np.float32)
hyper = dict(base_lr=0.05, momentum=0.9, weight_decay=5e-4, lr_policy="inv", gamma=1e-3, power=0.75)
sol = pyref.Solver(t_vid, t_off, t_sid, t_feat, t_W0, t_b0, TB, 5, 10, 60, 50, 6, **hyper)
layer_names = sol.layer_names()
traj = [sol.step() for _ in range(8)]
st = sol.state(); sol.close()
# the same run with the shipped file's TEST-phase graph in the NetParameter: the TRAIN trajectory must not move, and
# Solver::Test's loop (2 iterations of the TEST net on shared weights) is recorded before and after the 8 steps
tn, tF, tTB = 50, 4, 20
t_tdata = np.maximum(rng.normal(0, 1, (tn, tF, TK)), 0).astype(np.float32); t_tvid = rng.randint(0, 9, tn).astype(np.int32)
idmap = {v: v % 3 for v in range(9) if v != 4}                  # video 4 unlisted -> class 0 (std::map operator[])
idf = os.path.join(OUT, "_id2class.tmp"); open(idf, "w").write("".join("%d,%d\n" % kv for kv in idmap.items()))
sol = pyref.Solver(t_vid, t_off, t_sid, t_feat, t_W0, t_b0, TB, 5, 10, 60, 50, 6,
                   test=dict(data=t_tdata, video_id=t_tvid, batch=tTB, id_to_class_file=idf, exclude_same=True), **hyper)
test_before = sol.test(2)
traj2 = [sol.step() for _ in range(8)]
test_after = sol.test(2)
assert traj2 == traj and sol.layer_names() == layer_names
test_layer_names, test_output_names = sol.test_layer_names(), sol.test_output_names()
sol.close(); os.remove(idf)
# a second trajectory with the other branches of the path: L1 hinge (max_margin_loss norm: L1), "step" learning-rate
# policy, L1 weight regularisation
alt_hyper = dict(base_lr=0.02, momentum=0.8, weight_decay=1e-3, lr_policy="step", gamma=0.5, power=0.0, stepsize=2)
sol = pyref.Solver(t_vid, t_off, t_sid, t_feat, t_W0, t_b0, TB, 5, 10, 60, 50, 6, norm=1, reg_type=1, **alt_hyper)
alt_traj = [sol.step() for _ in range(6)]
alt_st = sol.state(); sol.close()
alt = dict(alt_loss=np.array([t[0] for t in alt_traj], np.float32), alt_violations=np.array([t[1] for t in alt_traj], np.float32),
           alt_W=alt_st["W"], alt_b=alt_st["b"], alt_hW=alt_st["hW"], alt_hb=alt_st["hb"],
           alt_hyper=np.array([0.02, 0.8, 1e-3, 0.5, 2], np.float64))
np.savez_compressed(os.path.join(OUT, "solver_ref.npz"), **alt, vid=t_vid, off=t_off, sid=t_sid, feat=t_feat, W0=t_W0, b0=t_b0,
                    cfg=np.array([TB, 5, 10, 60, 50, 6], np.int32), hyper=np.array([0.05, 0.9, 5e-4, 1e-3, 0.75], np.float64),
                    loss=np.array([t[0] for t in traj], np.float32), violations=np.array([t[1] for t in traj], np.float32),
                    W=st["W"], b=st["b"], hW=st["hW"], hb=st["hb"], layer_names=np.array(layer_names),
                    test_data=t_tdata, test_vid=t_tvid, test_batch=np.int32(tTB), id_keys=np.array(list(idmap.keys()), np.int32),
                    id_vals=np.array(list(idmap.values()), np.int32), test_before=test_before, test_after=test_after,
                    test_layer_names=np.array(test_layer_names), test_output_names=np.array(test_output_names))
print("solver_ref.npz written: losses", [round(t[0], 5) for t in traj], "test", test_before, test_after)
