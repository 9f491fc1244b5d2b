#!/usr/bin/env python
"""ref_shim: generate accessor-only C++ classes from the reference's .proto files (a tiny stand-in for protoc:
no wire format, no text format -- just the generated-code API surface the reference's headers and layer sources
use).  Reads /root/reference/src/caffe/proto/*.proto, writes headers into oracle/_ref/gen/caffe/proto/.
Test infrastructure only; runs in the build container where /root/reference exists."""
import os
import re
import sys

SCALARS = {"int32": "int32_t", "int64": "int64_t", "uint32": "uint32_t", "uint64": "uint64_t", "float": "float",
           "double": "double", "bool": "bool", "string": "std::string", "bytes": "std::string"}


def strip_comments(s):
    return re.sub(r"//[^\n]*", "", s)


def parse_block(body, prefix, messages, enums):
    """body: text inside a message {...}; returns list of fields. Nested messages / enums are hoisted with a prefix."""
    fields = []
    i = 0
    while i < len(body):
        m = re.compile(r"\s*(message|enum)\s+(\w+)\s*\{").match(body, i)
        if m:
            depth, j = 1, m.end()
            while depth:
                depth += {"{": 1, "}": -1}.get(body[j], 0)
                j += 1
            inner = body[m.end():j - 1]
            name = prefix + "_" + m.group(2) if prefix else m.group(2)
            if m.group(1) == "enum":
                vals = re.findall(r"(\w+)\s*=\s*(-?\d+)", inner)
                enums[name] = vals
                fields.append(("__enum__", m.group(2), name, vals))
            else:
                messages[name] = None      # reserve order
                messages[name] = parse_block(inner, name, messages, enums)
            i = j
            continue
        m = re.compile(r"\s*(optional|repeated|required)\s+([\w.]+)\s+(\w+)\s*=\s*\d+\s*(\[[^\]]*\])?\s*;").match(body, i)
        if m:
            label, typ, name, opts = m.groups()
            dflt = None
            if opts:
                d = re.search(r"default\s*=\s*([^,\]]+)", opts)
                if d:
                    dflt = d.group(1).strip()
            fields.append((label, typ, name, dflt))
            i = m.end()
            continue
        m = re.compile(r"\s*(extensions|option|reserved)[^;]*;").match(body, i)
        if m:
            i = m.end(); continue
        i += 1
    return fields


def resolve(typ, scope, messages, enums):
    """Find the C++ name of a message/enum type referenced from `scope` (innermost first)."""
    if "." in typ and typ.split(".")[0] not in messages:      # cross-package reference, e.g. caffe.Datum
        return "::" + typ.replace(".", "::")
    typ = typ.replace(".", "_")
    parts = scope.split("_") if scope else []
    for k in range(len(parts), -1, -1):
        cand = "_".join(parts[:k] + [typ]) if k else typ
        if cand in messages or cand in enums:
            return cand
    return typ


def gen(proto_text, package):
    text = strip_comments(proto_text)
    messages, enums = {}, {}
    parse_block(text, "", messages, enums)
    out = []
    # enums first
    for name, vals in enums.items():
        plain = ["%s = %s" % (v, n) for v, n in vals] if "_" not in name else []     # protoc: top-level enum values are plain names
        out.append("enum %s { %s };" % (name, ", ".join(["%s_%s = %s" % (name, v, n) for v, n in vals] + plain)))
        out.append("inline const std::string& %s_Name(%s v) { static std::map<int, std::string> m = {%s}; static std::string e; auto it = m.find(int(v)); return it == m.end() ? e : it->second; }" %
                   (name, name, ", ".join('{%s, "%s"}' % (n, v) for v, n in vals)))
        out.append("inline bool %s_IsValid(int v) { return %s; }" % (name, " || ".join("v == %s" % n for _, n in vals) or "false"))
    for name in messages:
        out.append("class %s;" % name)
    # messages in dependency-safe form: message fields are held through std::shared_ptr, so order is free
    merges = {}
    for name, fields in messages.items():
        pub, priv, clear, copy, merge = [], [], [], [], []
        merges[name] = merge
        for f in fields:
            if f[0] == "__enum__":
                _, short, full, vals = f
                pub.append("  typedef %s %s;" % (full, short))
                for v, _n in vals:
                    pub.append("  static const %s %s = %s_%s;" % (full, v, full, v))
                pub.append("  static const std::string& %s_Name(%s v) { return %s_Name(v); }" % (short, full, full))
                continue
            label, typ, fname, dflt = f
            is_scalar = typ in SCALARS
            ctype = SCALARS[typ] if is_scalar else resolve(typ, name, messages, enums)
            is_enum = (not is_scalar) and ctype in enums
            is_msg = (not is_scalar) and not is_enum
            if label == "repeated":
                if is_msg:
                    pub += ["  int %s_size() const { return int(%s_.size()); }" % (fname, fname),
                            "  const %s& %s(int i) const { return *%s_[i]; }" % (ctype, fname, fname),
                            "  %s* mutable_%s(int i) { return %s_[i].get(); }" % (ctype, fname, fname),
                            "  %s* add_%s();" % (ctype, fname),
                            "  void clear_%s() { %s_.clear(); }" % (fname, fname)]
                    priv.append("  std::vector<std::shared_ptr<%s> > %s_;" % (ctype, fname))
                    merge.append("  for (const auto& p : o.%s_) add_%s()->MergeFrom(*p);" % (fname, fname))
                else:
                    arg = "const std::string&" if ctype == "std::string" else ctype
                    ret = "const std::string&" if ctype == "std::string" else ctype
                    pub += ["  int %s_size() const { return int(%s_.size()); }" % (fname, fname),
                            "  %s %s(int i) const { return %s_[i]; }" % (ret, fname, fname),
                            "  void set_%s(int i, %s v) { %s_[i] = v; }" % (fname, arg, fname),
                            "  void add_%s(%s v) { %s_.push_back(v); }" % (fname, arg, fname),
                            "  void clear_%s() { %s_.clear(); }" % (fname, fname),
                            "  const ::google::protobuf::RepeatedField<%s>& %s() const { return %s_; }" % (ctype, fname, fname),
                            "  ::google::protobuf::RepeatedField<%s>* mutable_%s() { return &%s_; }" % (ctype, fname, fname)]
                    priv.append("  ::google::protobuf::RepeatedField<%s> %s_;" % (ctype, fname))
                    merge.append("  %s_.insert(%s_.end(), o.%s_.begin(), o.%s_.end());" % (fname, fname, fname, fname))
                clear.append("    %s_.clear();" % fname)
                continue
            if is_msg:
                pub += ["  bool has_%s() const { return bool(%s_); }" % (fname, fname),
                        "  const %s& %s() const;" % (ctype, fname),
                        "  %s* mutable_%s();" % (ctype, fname),
                        "  void clear_%s() { %s_.reset(); }" % (fname, fname)]
                priv.append("  std::shared_ptr<%s> %s_;" % (ctype, fname))
                clear.append("    %s_.reset();" % fname)
                merge.append("  if (o.%s_) mutable_%s()->MergeFrom(*o.%s_);" % (fname, fname, fname))
                continue
            if dflt is None:
                d = '""' if ctype == "std::string" else ("%s(0)" % ctype)
                if is_enum:
                    d = "%s_%s" % (ctype, enums[ctype][0][0])
            elif is_enum:
                d = "%s_%s" % (ctype, dflt)
            elif ctype == "std::string":
                d = '"' + dflt.strip("'\"") + '"'
            elif ctype == "float":
                d = dflt if re.search(r"[.eE]", dflt) else dflt + ".0"
                d = d + "f" if not d.lower().endswith(("inf", "nan")) else d
            else:
                d = dflt
            if ctype == "std::string":
                pub += ["  const std::string& %s() const { return %s_; }" % (fname, fname),
                        "  void set_%s(const std::string& v) { %s_ = v; has_%s_ = true; }" % (fname, fname, fname),
                        "  void set_%s(const char* v) { %s_ = v; has_%s_ = true; }" % (fname, fname, fname),
                        "  std::string* mutable_%s() { has_%s_ = true; return &%s_; }" % (fname, fname, fname)]
            else:
                pub += ["  %s %s() const { return %s_; }" % (ctype, fname, fname),
                        "  void set_%s(%s v) { %s_ = v; has_%s_ = true; }" % (fname, ctype, fname, fname)]
            pub += ["  bool has_%s() const { return has_%s_; }" % (fname, fname),
                    "  void clear_%s() { %s_ = %s; has_%s_ = false; }" % (fname, fname, d, fname)]
            priv += ["  %s %s_ = %s;" % (ctype, fname, d), "  bool has_%s_ = false;" % fname]
            clear.append("    clear_%s();" % fname)
            merge.append("  if (o.has_%s_) { %s_ = o.%s_; has_%s_ = true; }" % (fname, fname, fname, fname))
        out.append("class %s : public ::google::protobuf::Message {\n public:\n  %s() {}\n" % (name, name) + "\n".join(pub) +
                   "\n  void Clear() {\n" + "\n".join(clear) + "\n  }\n"
                   "  // protobuf semantics: MergeFrom takes the fields that are set in o (repeated fields append, sub-messages merge\n"
                   "  // recursively); CopyFrom is a deep copy.  operator= stays the shallow member-wise copy (sub-messages shared)\n"
                   "  void MergeFrom(const %s& o);\n  void CopyFrom(const %s& o) { if (&o == this) return; Clear(); MergeFrom(o); }\n"
                   "  std::string DebugString() const { return \"<%s>\"; }\n"
                   "  bool ParseFromString(const std::string&) { return false; }\n"
                   "  // no wire format in the shim: the in-memory fake LMDB of ref_driver.cpp hands out a pointer to a live message\n"
                   "  // (size = VV_SHIM_LIVE_OBJECT) and parsing is a copy\n"
                   "  bool ParseFromArray(const void* p, int n) { if (n != -0x5EED) return false; *this = *static_cast<const %s*>(p); return true; }\n"
                   "  bool SerializeToString(std::string*) const { return false; }\n"
                   "  static const %s& default_instance() { static %s d; return d; }\n private:\n" % (name, name, name, name, name, name) +
                   "\n".join(priv) + "\n};")
    # out-of-line bodies that need complete types
    for name in messages:
        out.append("inline void %s::MergeFrom(const %s& o) {\n%s\n}" % (name, name, "\n".join(merges[name])))
    for name, fields in messages.items():
        for f in fields:
            if f[0] == "__enum__":
                continue
            label, typ, fname, dflt = f
            if typ in SCALARS:
                continue
            ctype = resolve(typ, name, messages, enums)
            if ctype in enums:
                continue
            if label == "repeated":
                out.append("inline %s* %s::add_%s() { %s_.push_back(std::make_shared<%s>()); return %s_.back().get(); }" % (ctype, name, fname, fname, ctype, fname))
            else:
                out.append("inline const %s& %s::%s() const { return %s_ ? *%s_ : %s::default_instance(); }" % (ctype, name, fname, fname, fname, ctype))
                out.append("inline %s* %s::mutable_%s() { if (!%s_) %s_ = std::make_shared<%s>(); return %s_.get(); }" % (ctype, name, fname, fname, fname, ctype, fname))
    ns_open = "".join("namespace %s {\n" % p for p in package.split(".")) if package else ""
    ns_close = "".join("}\n" for _ in package.split(".")) if package else ""
    imports = "".join('#include "%s"\n' % i.replace(".proto", ".pb.h") for i in re.findall(r'import\s+"([^"]+)"', proto_text))
    return ("// GENERATED by oracle/ref_shim/gen_pb_shim.py from the reference's .proto (accessor-only, no wire format)\n"
            "#pragma once\n" + imports + "#include <cstdint>\n#include <map>\n#include <memory>\n#include <string>\n#include <vector>\n"
            "#include \"google/protobuf/message.h\"\n#include \"google/protobuf/repeated_field.h\"\n" + ns_open + "\n".join(out) + "\n" + ns_close)


if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    outdir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "_ref", "gen")
    os.makedirs(os.path.join(outdir, "caffe", "proto"), exist_ok=True)
    pdir = os.path.join(ref, "src", "caffe", "proto")
    for fn in sorted(os.listdir(pdir)):
        if not fn.endswith(".proto"):
            continue
        text = open(os.path.join(pdir, fn)).read()
        pkg = re.search(r"package\s+([\w.]+)\s*;", text)
        hdr = gen(text, pkg.group(1) if pkg else "")
        open(os.path.join(outdir, "caffe", "proto", fn.replace(".proto", ".pb.h")), "w").write(hdr)
        print("generated", fn.replace(".proto", ".pb.h"))
