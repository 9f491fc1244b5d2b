"""Writer of V1 `layers {}` prototxts with the structure and layer/blob names of the reference project
(projects/videovec_embedding/mednet_embedding_train.prototxt:2-671 and ..._solver.prototxt), parameterised on
batch, window, negatives and dimensions (BASELINE configs 1-4).  The data layer's `source` is a synthetic://
feature bank because the LMDB reader is out of scope."""


def train_net(B=128, C=5, Nn=10, K=4096, N=512, dropout=0.9, margin=2.0, norm="L2", videos=2048, shots=32, seed=1234,
              max_buffer_size=5000, swap=50, max_same=6, name="med_embedding", test=None, source=None,
              context_type="WINDOW"):
    """test = dict(batch=673, frames=4, videos=.., shots=.., seed=.., id_to_class_file=.., exclude_same=True) adds the shipped
    file's TEST-phase graph (data -> frame average -> shared fc7 -> test_norm -> retrieval_stats).  `source` (and
    test["source"]): a record source in place of the synthetic bank -- an LMDB directory, a VVRS stream or an mdb_dump text
    file of VideoShots (TEST: TestVideoShotWindows) records."""
    L = []

    def layer(s, both_phases=False):
        # fc7 / fc7_relu carry no include rule in the shipped file (used by TRAIN and TEST), all others are TRAIN-only
        L.append("layers {\n" + s.rstrip() + ("\n" if both_phases else "\n  include: { phase: TRAIN }\n") + "}\n")
    ctx = ["context_window_emb_%d_nonorm" % i for i in range(1, C)]
    neg = ["negative_emb_%d" % k for k in range(1, Nn + 1)]
    raw = ["target"] + ["context_window_%d" % i for i in range(1, C)] + ["negative_%d" % k for k in range(1, Nn + 1)]

    def tops(names, key="top"):
        return "".join('  %s: "%s"\n' % (key, n) for n in names)
    layer('  name: "shot_windows"\n  type: VIDEO_SAMPLED_SHOTS_DATA\n  top: "data"\n  video_sampled_shots_data_param {\n'
          '    source: "%s"\n    backend: LMDB\n    batch_size: %d\n'
          '    num_negative_samples: %d\n    max_buffer_size: %d\n    negative_swap_percentage: %d\n    max_same_video_negs: %d\n'
          '    context_type: %s\n    context_size: %d\n  }\n' % (source or "synthetic://videos=%d&shots=%d&dim=%d&seed=%d" % (videos, shots, K, seed),
                                                                B, Nn, max_buffer_size, swap, max_same, context_type, C))
    layer('  name: "slice_input_data"\n  type: SLICE\n  bottom: "data"\n' + tops(raw) + '  slice_param {\n    slice_dim: 1\n  }\n')
    layer('  name: "batch_concat_input"\n  type: CONCAT\n  top: "batch_concat"\n' + tops(raw, "bottom") + '  concat_param {\n    concat_dim: 0\n  }\n')
    layer('  name: "flatten_input"\n  type: FLATTEN\n  bottom: "batch_concat"\n  top: "original_feature"\n')
    layer('  name: "fc7"\n  type: INNER_PRODUCT\n  bottom: "original_feature"\n  top: "ip1_nonorm"\n  blobs_lr: 1\n  blobs_lr: 2\n'
          '  weight_decay: 1\n  weight_decay: 0\n  inner_product_param {\n    num_output: %d\n    weight_filler {\n      type: "gaussian"\n'
          '      std: 0.001\n    }\n    bias_filler {\n      type: "constant"\n    }\n  }\n' % N, both_phases=True)
    layer('  name: "fc7_relu"\n  type: RELU\n  top: "ip2"\n  bottom: "ip1_nonorm"\n', both_phases=True)
    if dropout and dropout > 0:
        layer('  name: "drop2"\n  type: DROPOUT\n  bottom: "ip2"\n  top: "ip2"\n  dropout_param {\n    dropout_ratio: %g\n  }\n' % dropout)
    layer('  name: "slice_emb"\n  type: SLICE\n  bottom: "ip2"\n' + tops(["target_emb_nonorm"] + ctx + [n + "_nonorm" for n in neg]) +
          '  slice_param {\n    slice_dim: 0\n  }\n')
    layer('  name: "context_average"\n  type: ELTWISE\n' + tops(ctx, "bottom") + '  top: "context_feature_nonorm"\n  eltwise_param {\n'
          '    operation: SUM\n' + "".join("    coeff: %.10g\n" % (1.0 / (C - 1)) for _ in ctx) + '  }\n')
    layer('  name: "word_embedding_norm"\n  type: NORMALIZATION\n  bottom: "context_feature_nonorm"\n  top: "context_feature"\n')
    layer('  name: "concat_pos_neg_nonorm"\n  type: CONCAT\n  top: "pos_neg_nonorm"\n' + tops(["target_emb_nonorm"] + [n + "_nonorm" for n in neg], "bottom") +
          '  concat_param {\n    concat_dim: 0\n  }\n')
    layer('  name: "pos_neg_normalize"\n  type: NORMALIZATION\n  bottom: "pos_neg_nonorm"\n  top: "pos_neg_norm"\n')
    layer('  name: "slice_pos_neg_norm"\n  type: SLICE\n  bottom: "pos_neg_norm"\n' + tops(["target_emb"] + neg) + '  slice_param {\n    slice_dim: 0\n  }\n')
    layer('  name: "prod_true"\n  type: ELTWISE\n  bottom: "context_feature"\n  bottom: "target_emb"\n  top: "target_prod"\n  eltwise_param {\n    operation: PROD\n  }\n')
    layer('  name: "sum_true"\n  type: SUM\n  bottom: "target_prod"\n  top: "target_score"\n  sum_param {\n    num_output: %d\n  }\n' % Nn)
    for k in range(1, Nn + 1):
        layer('  name: "prod_neg_%d"\n  type: ELTWISE\n  bottom: "context_feature"\n  bottom: "negative_emb_%d"\n  top: "negative_emb_%d_prod"\n'
              '  eltwise_param {\n    operation: PROD\n  }\n' % (k, k, k))
        layer('  name: "sum_neg_%d"\n  type: SUM\n  bottom: "negative_emb_%d_prod"\n  top: "neg_score_%d"\n' % (k, k, k))
    layer('  name: "concat_negative_scores"\n  type: CONCAT\n' + tops(["neg_score_%d" % k for k in range(1, Nn + 1)], "bottom") +
          '  top: "negative_score"\n  concat_param {\n    concat_dim: 1\n  }\n')
    layer('  name: "max_margin_loss"\n  type: MAX_MARGIN_LOSS\n  bottom: "target_score"\n  bottom: "negative_score"\n  top: "loss_output"\n'
          '  top: "train_violations"\n  loss_weight: 1\n  loss_weight: 0\n  max_margin_loss_param {\n    norm: %s\n    margin: %g\n  }\n' % (norm, margin))
    # the TEST-only layers of the shipped file (:29-45, 75-177, 345-352, 673-689); without `test` only test_norm, to
    # exercise phase filtering
    norm_l = 'layers {\n  name: "test_norm"\n  type: NORMALIZATION\n  bottom: "ip2"\n  top: "ip2_norm"\n  include: { phase: TEST }\n}\n'
    head, tail = "", ""
    if test:
        F = test.get("frames", 4)
        fr = ["context_datum_%d" % i for i in range(1, F + 1)]
        sl = ["test_sample_frame_%d" % i for i in range(1, F + 1)]

        def tl(s):
            return "layers {\n" + s.rstrip() + "\n  include: { phase: TEST }\n}\n"
        head = (tl('  name: "shot_windows"\n  type: VIDEO_SHOT_WINDOW_TEST_DATA\n  top: "data"\n  top: "video_ids"\n  video_shot_window_test_data_param {\n'
                   '    source: "%s"\n    backend: LMDB\n    batch_size: %d\n  }\n'
                   % (test.get("source") or "synthetic://videos=%d&shots=%d&dim=%d&seed=%d&frames=%d" % (
                       test.get("videos", 128), test.get("shots", 16), K, test.get("seed", 4321), F), test.get("batch", 673))) +
                tl('  name: "slice_input_data"\n  type: SLICE\n  bottom: "data"\n' + tops(fr) + '  slice_param {\n    slice_dim: 1\n  }\n') +
                tl('  name: "batch_concat_input_test"\n  type: CONCAT\n' + tops(fr, "bottom") + '  top: "concat_input_datums"\n  concat_param {\n    concat_dim: 0\n  }\n') +
                tl('  name: "flatten_input"\n  type: FLATTEN\n  bottom: "concat_input_datums"\n  top: "concat_input_datums_flat"\n') +
                tl('  name: "slice_test"\n  type: SLICE\n  bottom: "concat_input_datums_flat"\n' + tops(sl) + '  slice_param {\n    slice_dim: 0\n  }\n') +
                tl('  name: "average_for_test"\n  type: ELTWISE\n' + tops(sl, "bottom") + '  top: "original_feature"\n  eltwise_param {\n    operation: SUM\n' +
                   "".join("    coeff: %.10g\n" % (1.0 / F) for _ in sl) + '  }\n'))
        tail = tl('  name: "retrieval_stats"\n  type: RETRIEVAL_STATS\n  bottom: "ip2_norm"\n  bottom: "video_ids"\n  top: "test_map"\n  top: "test_hit_at_1"\n'
                  '  top: "test_hit_at_5"\n  retrieval_stats_param {\n    id_to_class_file: "%s"\n    exclude_same_video_shots: %s\n  }\n'
                  % (test["id_to_class_file"], "true" if test.get("exclude_same", True) else "false"))
    return 'name: "%s"\n' % name + "".join(L[:4]) + head + "".join(L[4:7]) + norm_l + "".join(L[7:]) + tail


def solver(net_path="", base_lr=0.001, momentum=0.9, weight_decay=0.0005, lr_policy="inv", gamma=0.001, power=0.75,
           max_iter=200000, display=10, random_seed=None, snapshot=2000, snapshot_prefix=None, test_iter=None, test_interval=50,
           test_initialization=True):
    s = ('net: "%s"\n' % net_path if net_path else "") + (
        "base_lr: %g\nmomentum: %g\nweight_decay: %g\nlr_policy: \"%s\"\ngamma: %g\npower: %g\n"
        "display: %d\nmax_iter: %d\nsnapshot: %d\nsolver_mode: GPU\n" % (base_lr, momentum, weight_decay, lr_policy, gamma, power, display, max_iter, snapshot))
    if test_iter is not None:
        s += "test_iter: %d\ntest_interval: %d\ntest_initialization: %s\n" % (test_iter, test_interval, "true" if test_initialization else "false")
    if snapshot_prefix is not None:
        s += 'snapshot_prefix: "%s"\n' % snapshot_prefix
    if random_seed is not None:
        s += "random_seed: %d\n" % random_seed
    return s
