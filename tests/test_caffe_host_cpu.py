"""Host logic of the caffe_compat layer that needs no GPU: prototxt text parsing, FilterNet, InsertSplits
(ref: src/caffe/net.cpp:227-268, src/caffe/util/insert_splits.cpp:12-142)."""
import re

import pytest

from videovector_b200 import caffe_host, prototxt


def names(txt):
    return re.findall(r'^  name: "(.*)"', txt, flags=re.M)


def test_train_graph_after_filter_and_splits(vvlib):
    C, Nn = 5, 10
    out = caffe_host.transform_net(prototxt.train_net(B=8, C=C, Nn=Nn, K=64, N=32), "TRAIN")
    n = names(out)
    expect = (["shot_windows", "slice_input_data", "batch_concat_input", "flatten_input", "fc7", "fc7_relu", "drop2", "slice_emb",
               "context_average", "word_embedding_norm", "context_feature_word_embedding_norm_0_split", "concat_pos_neg_nonorm",
               "pos_neg_normalize", "slice_pos_neg_norm", "prod_true", "sum_true"] +
              [x for k in range(1, Nn + 1) for x in ("prod_neg_%d" % k, "sum_neg_%d" % k)] + ["concat_negative_scores", "max_margin_loss"])
    assert n == expect                                      # exactly one SPLIT, TEST-only layer removed
    tops = re.findall(r'top: "(context_feature_word_embedding_norm_0_split_\d+)"', out)
    assert tops == ["context_feature_word_embedding_norm_0_split_%d" % k for k in range(1 + Nn)]
    # consumers are rewired in layer order: split_0 -> prod_true, split_k -> prod_neg_k
    blocks = out.split("layers {")
    for blk in blocks:
        m = re.search(r'name: "prod_(true|neg_(\d+))"', blk)
        if m:
            k = 0 if m.group(1) == "true" else int(m.group(2))
            assert 'bottom: "context_feature_word_embedding_norm_0_split_%d"' % k in blk
    # loss_output has a loss weight but a single use: no split (insert_splits.cpp:48-57)
    assert "loss_output_" not in out


def test_test_phase_filter(vvlib):
    # TEST keeps fc7, fc7_relu (no include rule) and test_norm; the TEST data layers of the shipped file are out of
    # scope, so the filtered graph has no producer for fc7's bottom -- the reference's CHECK text is reproduced
    with pytest.raises(Exception, match="Unknown blob input original_feature to layer fc7"):
        caffe_host.transform_net(prototxt.train_net(B=8, K=64, N=32), "TEST")


def test_cfg4_graph(vvlib):
    out = caffe_host.transform_net(prototxt.train_net(B=4, C=17, Nn=50, K=64, N=32), "TRAIN")
    assert len(names(out)) == 16 + 2 * 50 + 2
    assert out.count("coeff: 0.0625") == 16


def test_parser_accepts_reference_style_text(vvlib):
    txt = '''name: "x"   # trailing comment
####################################################################
layers { name: "a" type: RELU bottom: "in" top: "out" include: { phase: TRAIN } }
layers {
  name: 'b'
  type: DROPOUT
  bottom: "out"
  top: "out"
  dropout_param { dropout_ratio: 0.9 }
}
'''
    # "in" has no producer: the reference aborts with "Unknown blob input" -- same text here
    with pytest.raises(Exception, match="Unknown blob input in"):
        caffe_host.transform_net(txt, "TRAIN")
    ok = txt.replace('layers { name: "a"', 'layers { name: "src" type: RELU top: "in" }\nlayers { name: "a"')
    out = caffe_host.transform_net(ok, "TRAIN")
    assert names(out) == ["src", "a", "b"] and "dropout_ratio: 0.9" in out
    with pytest.raises(Exception, match="missing|unbalanced|parse error|unexpected"):
        caffe_host.transform_net("layers { name: \"a\" ", "TRAIN")
