// wire.cpp -- protobuf binary wire format for the reference's messages, over the PbMsg tree.
//
// The reference writes `.caffemodel` (NetParameter) and `.solverstate` (SolverState) files with
// WriteProtoToBinaryFile (ref: src/caffe/util/io.cpp:49-67, solver.cpp:320-341) and reads them back in
// Net::CopyTrainedLayersFrom / Solver::Restore (net.cpp:692-727, solver.cpp:418-429).  libprotobuf is not
// available here, so this is a hand-rolled encoder/decoder driven by a schema table (field numbers, wire
// types, enum numbers = the on-disk contract, caffe/proto/caffe_schema.inc).
//
// Wire format (protobuf encoding spec): a message is a sequence of (tag, value) records, tag = varint
// (field_number << 3 | wire_type); wire types 0 varint, 1 64-bit, 2 length-delimited, 5 32-bit.  Negative
// int32/int64/enum values are 10-byte varints; packed repeated scalars are one length-delimited record.
#include <algorithm>
#include <cstring>
#include <fstream>
#include <map>
#include "caffe/proto/caffe_params.hpp"

namespace caffe {
namespace {

enum Kind { kVarint = 0, kBool = 1, kEnum = 2, kFloat = 3, kDouble = 4, kString = 5, kMessage = 6, kFixed32 = 7, kFixed64 = 8, kZigzag = 9 };
struct FieldDef { const char* msg; const char* name; int number; int kind; int repeated; int packed; const char* type; };
struct EnumDef { const char* name; const char* value; int number; };

#define VV_PB_FIELD(m, f, n, k, r, p, t) {m, f, n, k, r, p, t},
#define VV_PB_ENUM(e, v, n)
const FieldDef kFields[] = {
#include "caffe/proto/caffe_schema.inc"
};
#undef VV_PB_FIELD
#undef VV_PB_ENUM
#define VV_PB_FIELD(m, f, n, k, r, p, t)
#define VV_PB_ENUM(e, v, n) {e, v, n},
const EnumDef kEnums[] = {
#include "caffe/proto/caffe_schema.inc"
};
#undef VV_PB_FIELD
#undef VV_PB_ENUM

struct Schema {
  std::map<string, std::map<string, const FieldDef*> > by_name;   // message -> field name -> def
  std::map<string, std::map<int, const FieldDef*> > by_number;
  std::map<string, std::map<string, int> > enum_number;
  std::map<string, std::map<int, string> > enum_name;
  Schema() {
    for (const FieldDef& f : kFields) { by_name[f.msg][f.name] = &f; by_number[f.msg][f.number] = &f; }
    for (const EnumDef& e : kEnums) { enum_number[e.name][e.value] = e.number; if (!enum_name[e.name].count(e.number)) enum_name[e.name][e.number] = e.value; }
  }
};
const Schema& schema() { static Schema s; return s; }

// ---- encoder ----------------------------------------------------------------------------------------------
void put_varint(string* out, uint64_t v) {
  while (v >= 0x80) { out->push_back(char((v & 0x7F) | 0x80)); v >>= 7; }
  out->push_back(char(v));
}
void put_tag(string* out, int number, int wire) { put_varint(out, (uint64_t(number) << 3) | uint64_t(wire)); }
void put_fixed32(string* out, uint32_t v) { char b[4]; memcpy(b, &v, 4); out->append(b, 4); }   // little-endian hosts only
void put_fixed64(string* out, uint64_t v) { char b[8]; memcpy(b, &v, 8); out->append(b, 8); }

long long parse_int(const string& s, const string& what) {
  char* end = nullptr;
  const long long v = strtoll(s.c_str(), &end, 0);
  if (end == s.c_str() || *end != 0) {           // "1e3"-style values of integer fields in hand-written prototxt
    const double d = strtod(s.c_str(), &end);
    CHECK(end != s.c_str() && *end == 0) << "field " << what << " is not a number: " << s;
    return (long long)d;
  }
  return v;
}

void encode_scalar(string* out, const FieldDef& d, const string& v, bool with_tag) {
  switch (d.kind) {
    case kVarint: {
      if (with_tag) put_tag(out, d.number, 0);
      put_varint(out, uint64_t(parse_int(v, d.name)));              // negative values sign-extend to 64 bits = 10 bytes
      break;
    }
    case kZigzag: {
      if (with_tag) put_tag(out, d.number, 0);
      const long long x = parse_int(v, d.name);
      put_varint(out, (uint64_t(x) << 1) ^ uint64_t(x >> 63));
      break;
    }
    case kBool: if (with_tag) put_tag(out, d.number, 0); put_varint(out, (v == "true" || v == "1") ? 1 : 0); break;
    case kEnum: {
      if (with_tag) put_tag(out, d.number, 0);
      const auto& tbl = schema().enum_number.at(d.type);
      auto it = tbl.find(v);
      long long n;
      if (it != tbl.end()) n = it->second; else n = parse_int(v, d.name);   // numeric form is legal text format too
      put_varint(out, uint64_t(n));
      break;
    }
    case kFloat: { if (with_tag) put_tag(out, d.number, 5); const float f = strtof(v.c_str(), nullptr); uint32_t u; memcpy(&u, &f, 4); put_fixed32(out, u); break; }
    case kDouble: { if (with_tag) put_tag(out, d.number, 1); const double f = strtod(v.c_str(), nullptr); uint64_t u; memcpy(&u, &f, 8); put_fixed64(out, u); break; }
    case kFixed32: if (with_tag) put_tag(out, d.number, 5); put_fixed32(out, uint32_t(parse_int(v, d.name))); break;
    case kFixed64: if (with_tag) put_tag(out, d.number, 1); put_fixed64(out, uint64_t(parse_int(v, d.name))); break;
    case kString: if (with_tag) put_tag(out, d.number, 2); put_varint(out, v.size()); out->append(v); break;
    default: LOG_FATAL << "encode_scalar: bad kind";
  }
}

void encode_msg(string* out, const PbMsg& m, const string& type) {
  auto mt = schema().by_name.find(type);
  CHECK(mt != schema().by_name.end()) << "unknown protobuf message type " << type;
  // stable order: by field number, repeated occurrences in stored order (what libprotobuf emits)
  vector<std::pair<int, const PbField*> > order;
  for (const PbField& f : m.fields) {
    auto it = mt->second.find(f.key);
    CHECK(it != mt->second.end()) << "message " << type << " has no field '" << f.key << "'";
    order.push_back(std::make_pair(it->second->number, &f));
  }
  std::stable_sort(order.begin(), order.end(), [](const std::pair<int, const PbField*>& a, const std::pair<int, const PbField*>& b) { return a.first < b.first; });
  for (size_t i = 0; i < order.size(); ++i) {
    const PbField& f = *order[i].second;
    const FieldDef& d = *mt->second.at(f.key);
    if (d.kind == kMessage) {
      CHECK(f.msg) << "field " << f.key << " of " << type << " must be a message";
      string body; encode_msg(&body, *f.msg, d.type);
      put_tag(out, d.number, 2); put_varint(out, body.size()); out->append(body);
    } else if (f.floats) {
      CHECK(d.kind == kFloat && d.repeated) << "float array on a non-float field " << f.key;
      if (f.floats->empty()) continue;                              // an empty packed field is not emitted
      if (d.packed) {
        put_tag(out, d.number, 2); put_varint(out, f.floats->size() * 4);
        out->append(reinterpret_cast<const char*>(f.floats->data()), f.floats->size() * 4);
      } else {
        for (float v : *f.floats) { put_tag(out, d.number, 5); uint32_t u; memcpy(&u, &v, 4); put_fixed32(out, u); }
      }
    } else if (d.repeated && d.packed) {
      // consecutive scalar occurrences of a packed field form one record
      string body; size_t j = i;
      while (j < order.size() && order[j].first == d.number && !order[j].second->floats) { encode_scalar(&body, d, order[j].second->scalar, false); ++j; }
      put_tag(out, d.number, 2); put_varint(out, body.size()); out->append(body);
      i = j - 1;
    } else {
      encode_scalar(out, d, f.scalar, true);
    }
  }
}

// ---- decoder ----------------------------------------------------------------------------------------------
struct Reader {
  const unsigned char* p; const unsigned char* end;
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0; int shift = 0;
    for (;;) {
      CHECK(p < end) << "truncated varint";
      const unsigned char b = *p++;
      v |= uint64_t(b & 0x7F) << shift;
      if (!(b & 0x80)) break;
      shift += 7;
      CHECK_LT(shift, 70) << "malformed varint";
    }
    return v;
  }
  uint32_t fixed32() { CHECK(p + 4 <= end) << "truncated fixed32"; uint32_t v; memcpy(&v, p, 4); p += 4; return v; }
  uint64_t fixed64() { CHECK(p + 8 <= end) << "truncated fixed64"; uint64_t v; memcpy(&v, p, 8); p += 8; return v; }
  Reader sub() { const uint64_t n = varint(); CHECK(uint64_t(end - p) >= n) << "truncated length-delimited field"; Reader r{p, p + n}; p += n; return r; }
};

string fmt_float(float f) { char b[64]; snprintf(b, sizeof(b), "%.9g", double(f)); return b; }
string fmt_double(double f) { char b[64]; snprintf(b, sizeof(b), "%.17g", f); return b; }

void decode_scalar(PbMsg* m, const FieldDef& d, int wire, Reader& r) {
  switch (d.kind) {
    case kVarint: { CHECK_EQ(wire, 0); m->add_scalar(d.name, std::to_string((long long)r.varint())); break; }
    case kZigzag: { CHECK_EQ(wire, 0); const uint64_t u = r.varint(); m->add_scalar(d.name, std::to_string((long long)((u >> 1) ^ (~(u & 1) + 1)))); break; }
    case kBool: { CHECK_EQ(wire, 0); m->add_scalar(d.name, r.varint() ? "true" : "false"); break; }
    case kEnum: {
      CHECK_EQ(wire, 0);
      const int n = int((long long)r.varint());
      const auto& tbl = schema().enum_name.at(d.type);
      auto it = tbl.find(n);
      m->add_scalar(d.name, it != tbl.end() ? it->second : std::to_string(n));
      break;
    }
    case kFloat: { CHECK_EQ(wire, 5); const uint32_t u = r.fixed32(); float f; memcpy(&f, &u, 4); m->add_scalar(d.name, fmt_float(f)); break; }
    case kDouble: { CHECK_EQ(wire, 1); const uint64_t u = r.fixed64(); double f; memcpy(&f, &u, 8); m->add_scalar(d.name, fmt_double(f)); break; }
    case kFixed32: { CHECK_EQ(wire, 5); m->add_scalar(d.name, std::to_string(r.fixed32())); break; }
    case kFixed64: { CHECK_EQ(wire, 1); m->add_scalar(d.name, std::to_string(r.fixed64())); break; }
    default: LOG_FATAL << "decode_scalar: bad kind";
  }
}

shared_ptr<PbMsg> decode_msg(Reader r, const string& type) {
  auto mt = schema().by_number.find(type);
  CHECK(mt != schema().by_number.end()) << "unknown protobuf message type " << type;
  auto m = std::make_shared<PbMsg>();
  while (!r.done()) {
    const uint64_t tag = r.varint();
    const int number = int(tag >> 3), wire = int(tag & 7);
    auto it = mt->second.find(number);
    if (it == mt->second.end()) {                       // unknown field: skip by wire type
      switch (wire) {
        case 0: r.varint(); break;
        case 1: r.fixed64(); break;
        case 2: r.sub(); break;
        case 5: r.fixed32(); break;
        default: LOG_FATAL << "unsupported wire type " << wire << " in " << type;
      }
      continue;
    }
    const FieldDef& d = *it->second;
    if (d.kind == kMessage) {
      CHECK_EQ(wire, 2) << "message field " << d.name << " with wire type " << wire;
      m->fields.push_back(PbField{d.name, "", decode_msg(r.sub(), d.type), nullptr});
    } else if (d.kind == kString) {
      CHECK_EQ(wire, 2);
      Reader s = r.sub();
      m->add_scalar(d.name, string(reinterpret_cast<const char*>(s.p), size_t(s.end - s.p)));
    } else if (wire == 2 && d.repeated) {
      // packed encoding (parsers must accept it for any repeated scalar field)
      Reader s = r.sub();
      if (d.kind == kFloat && d.packed) {
        // one array per field: append to an existing one (a field may legally be split into several records)
        PbField* dst = nullptr;
        for (auto& f : m->fields) if (f.key == d.name && f.floats) dst = &f;
        if (!dst) { m->fields.push_back(PbField{d.name, "", nullptr, std::make_shared<vector<float> >()}); dst = &m->fields.back(); }
        const size_t n = size_t(s.end - s.p) / 4, old = dst->floats->size();
        CHECK_EQ(size_t(s.end - s.p), n * 4) << "packed float field with a ragged length";
        dst->floats->resize(old + n);
        memcpy(dst->floats->data() + old, s.p, n * 4);
      } else {
        const int w = (d.kind == kDouble || d.kind == kFixed64) ? 1 : (d.kind == kFixed32 ? 5 : 0);
        while (!s.done()) decode_scalar(m.get(), d, w, s);
      }
    } else {
      decode_scalar(m.get(), d, wire, r);
    }
  }
  return m;
}

}  // namespace

string SerializeBinary(const PbMsg& m, const string& type) { string out; encode_msg(&out, m, type); return out; }
shared_ptr<PbMsg> ParseBinary(const string& bytes, const string& type) {
  Reader r{reinterpret_cast<const unsigned char*>(bytes.data()), reinterpret_cast<const unsigned char*>(bytes.data()) + bytes.size()};
  return decode_msg(r, type);
}
void WriteProtoToBinaryFile(const PbMsg& m, const string& type, const string& path) {
  const string bytes = SerializeBinary(m, type);
  std::ofstream f(path.c_str(), std::ios::binary | std::ios::trunc);
  CHECK(f.good()) << "cannot open " << path << " for writing";
  f.write(bytes.data(), std::streamsize(bytes.size()));
  CHECK(f.good()) << "write failed: " << path;
}
shared_ptr<PbMsg> ReadProtoFromBinaryFile(const string& path, const string& type) { return ParseBinary(ReadFileOrDie(path), type); }

}  // namespace caffe
