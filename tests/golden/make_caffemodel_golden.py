#!/usr/bin/env python
"""Golden binary protobuf files for the .caffemodel / .solverstate reader-writer, produced by the REAL protobuf
runtime (python `google.protobuf`) from the reference's own schema: the reference's caffe.proto is parsed into a
FileDescriptorSet (committed as caffe_schema.desc so the tests need no reference tree), message classes are
built from it, and a NetParameter (the shipped fc7 + loss layers with blobs) and a SolverState are serialised.
Run here (needs /root/reference); outputs: tests/golden/{caffe_schema.desc, ref_small.caffemodel,
ref_small.solverstate, ref_small_arrays.npz}."""
import os, re, sys
import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory, text_format

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
T = descriptor_pb2.FieldDescriptorProto
SCALAR = {"double": T.TYPE_DOUBLE, "float": T.TYPE_FLOAT, "int32": T.TYPE_INT32, "int64": T.TYPE_INT64, "uint32": T.TYPE_UINT32,
          "uint64": T.TYPE_UINT64, "bool": T.TYPE_BOOL, "string": T.TYPE_STRING, "bytes": T.TYPE_BYTES}


def build_descriptor(proto_text, package="caffe", name="caffe.proto", deps=(), external=()):
    """external: fully qualified message types of imported files (e.g. "caffe.Datum")."""
    text = re.sub(r"//[^\n]*", "", proto_text)
    fdp = descriptor_pb2.FileDescriptorProto(name=name, package=package, syntax="proto2")
    fdp.dependency.extend(deps)
    names = {}        # fully scoped name -> "message" / "enum"

    def scan(body, scope):
        i = 0
        while i < len(body):
            m = re.compile(r"\s*(message|enum)\s+(\w+)\s*\{").match(body, i)
            if m:
                depth, j = 1, m.end()
                while depth:
                    depth += {"{": 1, "}": -1}.get(body[j], 0); j += 1
                full = (scope + "." if scope else "") + m.group(2)
                names[full] = m.group(1)
                if m.group(1) == "message":
                    scan(body[m.end():j - 1], full)
                i = j; continue
            i += 1
    scan(text, "")

    def resolve(typ, scope):
        parts = scope.split(".") if scope else []
        for k in range(len(parts), -1, -1):
            cand = ".".join(parts[:k] + [typ])
            if cand in names:
                return cand
        raise KeyError((typ, scope))

    def fill(body, scope, add_msg, add_enum):
        i = 0
        out_fields = []
        while i < len(body):
            m = re.compile(r"\s*(message|enum)\s+(\w+)\s*\{").match(body, i)
            if m:
                depth, j = 1, m.end()
                while depth:
                    depth += {"{": 1, "}": -1}.get(body[j], 0); j += 1
                inner = body[m.end():j - 1]
                full = (scope + "." if scope else "") + m.group(2)
                if m.group(1) == "enum":
                    e = add_enum(); e.name = m.group(2)
                    for v, n in re.findall(r"(\w+)\s*=\s*(-?\d+)", inner):
                        ev = e.value.add(); ev.name = v; ev.number = int(n)
                else:
                    d = add_msg(); d.name = m.group(2)
                    for f in fill(inner, full, d.nested_type.add, d.enum_type.add):
                        d.field.add().CopyFrom(f)
                i = j; continue
            m = re.compile(r"\s*(optional|repeated|required)\s+([\w.]+)\s+(\w+)\s*=\s*(\d+)\s*(\[[^\]]*\])?\s*;").match(body, i)
            if m:
                label, typ, name, num, opts = m.groups()
                f = descriptor_pb2.FieldDescriptorProto(name=name, number=int(num))
                f.label = {"optional": T.LABEL_OPTIONAL, "repeated": T.LABEL_REPEATED, "required": T.LABEL_REQUIRED}[label]
                if typ in SCALAR:
                    f.type = SCALAR[typ]
                elif typ in external:
                    f.type = T.TYPE_MESSAGE
                    f.type_name = "." + typ
                else:
                    full = resolve(typ, scope)
                    f.type = T.TYPE_ENUM if names[full] == "enum" else T.TYPE_MESSAGE
                    f.type_name = "." + package + "." + full
                if opts and re.search(r"packed\s*=\s*true", opts):
                    f.options.packed = True
                d = re.search(r"default\s*=\s*([^,\]]+)", opts or "")
                if d:
                    f.default_value = d.group(1).strip().strip("\"'")
                out_fields.append(f)
                i = m.end(); continue
            i += 1
        return out_fields

    fill(text, "", fdp.message_type.add, fdp.enum_type.add)
    return fdp


def classes(fdp):
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("caffe." + n))
    return get


def main():
    fdp = build_descriptor(open(os.path.join(REF, "src", "caffe", "proto", "caffe.proto")).read())
    fds = descriptor_pb2.FileDescriptorSet(); fds.file.add().CopyFrom(fdp)
    open(os.path.join(HERE, "caffe_schema.desc"), "wb").write(fds.SerializeToString())
    get = classes(fdp)
    NetParameter, SolverState = get("NetParameter"), get("SolverState")
    rng = np.random.RandomState(1701)
    N, K = 6, 8
    W = rng.normal(0, 0.01, (N, K)).astype(np.float32); b = rng.normal(0, 0.1, N).astype(np.float32)
    dW = rng.normal(0, 1e-3, (N, K)).astype(np.float32)
    net = NetParameter()
    text_format.Parse('''
      name: "videovec_small"
      layers { name: "data" type: VIDEO_SAMPLED_SHOTS_DATA top: "data"
               video_sampled_shots_data_param { source: "synthetic" batch_size: 4 num_negative_samples: 10 context_size: 5
                                                max_buffer_size: 5000 negative_swap_percentage: 50 max_same_video_negs: 6 context_type: WINDOW }
               include { phase: TRAIN } }
      layers { name: "fc7" type: INNER_PRODUCT bottom: "original_feature" top: "ip1_nonorm" blobs_lr: 1 blobs_lr: 2 weight_decay: 1 weight_decay: 0
               inner_product_param { num_output: 6 weight_filler { type: "gaussian" std: 0.001 } bias_filler { type: "constant" value: 0 } } }
      layers { name: "drop" type: DROPOUT bottom: "ip1_nonorm" top: "ip2" dropout_param { dropout_ratio: 0.9 } }
      layers { name: "context_average" type: ELTWISE bottom: "c1" bottom: "c2" top: "ctx" eltwise_param { operation: SUM coeff: 0.25 coeff: 0.25 } }
      layers { name: "max_margin_loss" type: MAX_MARGIN_LOSS bottom: "s1" bottom: "s2" top: "loss_output" top: "train_violations"
               loss_weight: 1 loss_weight: 0 max_margin_loss_param { norm: L2 margin: 2 } }
    ''', net)
    fc7 = net.layers[1]
    bw = fc7.blobs.add(); bw.num, bw.channels, bw.height, bw.width = 1, 1, N, K; bw.data.extend(W.reshape(-1).tolist()); bw.diff.extend(dW.reshape(-1).tolist())
    bb = fc7.blobs.add(); bb.num, bb.channels, bb.height, bb.width = 1, 1, 1, N; bb.data.extend(b.tolist())
    open(os.path.join(HERE, "ref_small.caffemodel"), "wb").write(net.SerializeToString())
    hW = rng.normal(0, 1e-4, (N, K)).astype(np.float32); hb = rng.normal(0, 1e-4, N).astype(np.float32)
    st = SolverState(); st.iter = 12345; st.learned_net = "snap/videovec_iter_12345.caffemodel"
    h0 = st.history.add(); h0.num, h0.channels, h0.height, h0.width = 1, 1, N, K; h0.data.extend(hW.reshape(-1).tolist())
    h1 = st.history.add(); h1.num, h1.channels, h1.height, h1.width = 1, 1, 1, N; h1.data.extend(hb.tolist())
    open(os.path.join(HERE, "ref_small.solverstate"), "wb").write(st.SerializeToString())
    np.savez(os.path.join(HERE, "ref_small_arrays.npz"), W=W, b=b, dW=dW, hW=hW, hb=hb)
    print("wrote caffe_schema.desc (%d B), ref_small.caffemodel (%d B), ref_small.solverstate (%d B)" % (
        len(fds.SerializeToString()), len(net.SerializeToString()), len(st.SerializeToString())))


if __name__ == "__main__":
    main()
