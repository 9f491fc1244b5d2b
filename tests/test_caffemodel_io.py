"""`.caffemodel` / `.solverstate` binary protobuf reader-writer, host only.

Oracle = the real protobuf runtime: tests/golden/ref_small.{caffemodel,solverstate} were serialised by python
`google.protobuf` from message classes built out of the reference's own caffe.proto (make_caffemodel_golden.py;
the schema travels as tests/golden/caffe_schema.desc).  The hand-rolled codec (caffe_compat/wire.cpp) must
(1) decode them exactly, (2) re-encode them byte for byte, (3) produce files the protobuf runtime parses back."""
import os

import numpy as np
import pytest

from videovector_b200 import caffe_host as ch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module", autouse=True)
def _built(vvlib):            # builds libvv_b200.so on first use if the tree is fresh
    return vvlib


def _classes():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fds = descriptor_pb2.FileDescriptorSet()
    fds.ParseFromString(open(os.path.join(GOLD, "caffe_schema.desc"), "rb").read())
    pool = descriptor_pool.DescriptorPool()
    for f in fds.file:
        pool.Add(f)
    return lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("caffe." + n))


def test_decode_reference_caffemodel():
    ref = np.load(os.path.join(GOLD, "ref_small_arrays.npz"))
    text, arrays = ch.read_binary_proto(os.path.join(GOLD, "ref_small.caffemodel"), "NetParameter")
    assert np.array_equal(arrays["layers[1].blobs[0].data"], ref["W"].reshape(-1))
    assert np.array_equal(arrays["layers[1].blobs[0].diff"], ref["dW"].reshape(-1))
    assert np.array_equal(arrays["layers[1].blobs[1].data"], ref["b"])
    assert set(arrays) == {"layers[1].blobs[0].data", "layers[1].blobs[0].diff", "layers[1].blobs[1].data"}
    # the parameter messages come back as text format with enum names and the original values
    for needle in ('name: "videovec_small"', "type: VIDEO_SAMPLED_SHOTS_DATA", "type: INNER_PRODUCT", "num_output: 6",
                   "operation: SUM", "norm: L2", "phase: TRAIN", "context_type: WINDOW", 'source: "synthetic"',
                   "dropout_ratio: 0.899999976", "height: 6", "width: 8", "max_buffer_size: 5000"):
        assert needle in text, needle
    assert text.count("layers {") == 5 and text.count("blobs {") == 2
    assert text.count("coeff: 0.25") == 2 and text.count("loss_weight:") == 2 and text.count("blobs_lr:") == 2


def test_decode_reference_solverstate():
    ref = np.load(os.path.join(GOLD, "ref_small_arrays.npz"))
    text, arrays = ch.read_binary_proto(os.path.join(GOLD, "ref_small.solverstate"), "SolverState")
    assert "iter: 12345" in text and 'learned_net: "snap/videovec_iter_12345.caffemodel"' in text
    assert np.array_equal(arrays["history[0].data"], ref["hW"].reshape(-1))
    assert np.array_equal(arrays["history[1].data"], ref["hb"])


@pytest.mark.parametrize("name,typ", [("ref_small.caffemodel", "NetParameter"), ("ref_small.solverstate", "SolverState")])
def test_reencode_is_byte_identical(tmp_path, name, typ):
    """decode -> encode reproduces libprotobuf's serialisation (field-number order, packed floats, enum numbers)."""
    src = os.path.join(GOLD, name)
    text, arrays = ch.read_binary_proto(src, typ)
    out = tmp_path / name
    ch.write_binary_proto(out, typ, text, arrays)
    assert open(out, "rb").read() == open(src, "rb").read()


def test_written_model_parses_with_protobuf_runtime(tmp_path):
    """Our writer -> the real protobuf parser (what the reference's ReadProtoFromBinaryFile runs)."""
    get = _classes()
    rng = np.random.RandomState(3)
    W = rng.normal(0, 1, (3, 5)).astype(np.float32); b = rng.normal(0, 1, 3).astype(np.float32)
    W[0, 0] = np.float32(-0.0); W[1, 1] = np.float32(1e-38); W[2, 2] = np.float32(3.4e38)      # edge values survive bit-exactly
    text = ('name: "t"\nlayers { name: "fc7" type: INNER_PRODUCT bottom: "x" top: "y" blobs_lr: 1 blobs_lr: 2 '
            'inner_product_param { num_output: 3 regularization: 0.5 } '
            'blobs { num: 1 channels: 1 height: 3 width: 5 } blobs { num: 1 channels: 1 height: 1 width: 3 } }\n'
            'layers { name: "s" type: SLICE bottom: "a" top: "b" top: "c" slice_param { slice_dim: 1 slice_point: 1 slice_point: 5 } '
            'include { phase: TEST } }\n')
    p = tmp_path / "m.caffemodel"
    ch.write_binary_proto(p, "NetParameter", text, {"layers[0].blobs[0].data": W, "layers[0].blobs[1].data": b})
    net = get("NetParameter")()
    net.ParseFromString(open(p, "rb").read())
    assert net.name == "t" and len(net.layers) == 2
    fc7 = net.layers[0]
    assert fc7.name == "fc7" and fc7.type == 14 and list(fc7.bottom) == ["x"] and list(fc7.blobs_lr) == [1.0, 2.0]
    assert fc7.inner_product_param.num_output == 3 and fc7.inner_product_param.regularization == 0.5
    assert (fc7.blobs[0].num, fc7.blobs[0].channels, fc7.blobs[0].height, fc7.blobs[0].width) == (1, 1, 3, 5)
    got = np.array(fc7.blobs[0].data, np.float32)
    assert np.array_equal(got.view(np.uint32), W.reshape(-1).view(np.uint32))
    assert np.array_equal(np.array(fc7.blobs[1].data, np.float32), b)
    assert list(net.layers[1].slice_param.slice_point) == [1, 5] and net.layers[1].include[0].phase == 1
    # and the runtime's re-serialisation of what it parsed equals our bytes
    assert net.SerializeToString() == open(p, "rb").read()


def test_solverstate_roundtrip_and_negative_ints(tmp_path):
    get = _classes()
    h = np.arange(12, dtype=np.float32)
    p = tmp_path / "s.solverstate"
    ch.write_binary_proto(p, "SolverState", 'iter: 7\nlearned_net: "x.caffemodel"\nhistory { num: 1 channels: 1 height: 3 width: 4 }\n',
                          {"history[0].data": h})
    st = get("SolverState")(); st.ParseFromString(open(p, "rb").read())
    assert st.iter == 7 and st.learned_net == "x.caffemodel" and list(st.history[0].data) == list(h)
    # negative int32 -> 10-byte varint, as libprotobuf writes it
    q = tmp_path / "n.bin"
    ch.write_binary_proto(q, "SolverParameter", "random_seed: -1\nmax_iter: 300000\nbase_lr: 0.001\n", {})
    sp = get("SolverParameter")(); sp.ParseFromString(open(q, "rb").read())
    assert sp.random_seed == -1 and sp.max_iter == 300000 and abs(sp.base_lr - 0.001) < 1e-9
    assert sp.SerializeToString() == open(q, "rb").read()
    text, _ = ch.read_binary_proto(q, "SolverParameter")
    assert "random_seed: -1" in text


def test_unknown_fields_are_skipped_and_errors_are_loud(tmp_path):
    # a NetParameter with an extra unknown field (number 900, varint) in front: the decoder must skip it
    src = open(os.path.join(GOLD, "ref_small.solverstate"), "rb").read()
    p = tmp_path / "u.bin"
    open(p, "wb").write(bytes([0xA0, 0x38, 0x05]) + src)          # tag (900<<3|0) = 7200 -> varint a0 38, value 5
    text, arrays = ch.read_binary_proto(p, "SolverState")
    assert "iter: 12345" in text and len(arrays) == 2
    open(p, "wb").write(src[:-3])                                   # truncated file
    with pytest.raises(Exception):
        ch.read_binary_proto(p, "SolverState")
    with pytest.raises(Exception):
        ch.write_binary_proto(tmp_path / "x", "NetParameter", "no_such_field: 1\n", {})
