// protobuf text-format subset parser + parameter helpers (see caffe/proto/caffe_params.hpp)
#include <cctype>
#include <fstream>
#include "caffe/proto/caffe_params.hpp"

namespace caffe {

void LogInfo(const string& msg) {
  static const bool verbose = getenv("VV_CAFFE_VERBOSE") != nullptr;
  if (verbose) fprintf(stderr, "I %s\n", msg.c_str());
}

int PbMsg::count(const string& key) const { int n = 0; for (auto& f : fields) n += f.key == key; return n; }
const PbField* PbMsg::nth(const string& key, int i) const {
  for (auto& f : fields) if (f.key == key && i-- == 0) return &f;
  return nullptr;
}
string PbMsg::str(const string& key, const string& dflt, int i) const { auto f = nth(key, i); return (f && !f->msg) ? f->scalar : dflt; }
double PbMsg::num(const string& key, double dflt, int i) const {
  auto f = nth(key, i);
  if (!f || f->msg) return dflt;
  char* end = nullptr; const double v = strtod(f->scalar.c_str(), &end);
  CHECK(end != f->scalar.c_str()) << "field '" << key << "' is not a number: " << f->scalar;
  return v;
}
bool PbMsg::boolean(const string& key, bool dflt) const {
  auto f = nth(key, 0);
  if (!f || f->msg) return dflt;
  return f->scalar == "true" || f->scalar == "1";
}
shared_ptr<PbMsg> PbMsg::sub(const string& key, int i) const { auto f = nth(key, i); return (f && f->msg) ? f->msg : std::make_shared<PbMsg>(); }
void PbMsg::set_scalar(const string& key, const string& v) {
  for (auto& f : fields) if (f.key == key && !f.msg) { f.scalar = v; return; }
  add_scalar(key, v);
}
void PbMsg::add_scalar(const string& key, const string& v) { fields.push_back(PbField{key, v, nullptr}); }

namespace {
struct Lexer {
  const string& s; size_t p = 0; int line = 1;
  explicit Lexer(const string& t) : s(t) {}
  void skip() {
    for (;;) {
      while (p < s.size() && isspace((unsigned char)s[p])) { if (s[p] == '\n') ++line; ++p; }
      if (p < s.size() && s[p] == '#') { while (p < s.size() && s[p] != '\n') ++p; continue; }
      break;
    }
  }
  bool eof() { skip(); return p >= s.size(); }
  char peek() { skip(); return p < s.size() ? s[p] : '\0'; }
  string token() {
    skip();
    CHECK(p < s.size()) << "unexpected end of prototxt";
    if (s[p] == '"' || s[p] == '\'') {
      const char q = s[p++]; string out;
      while (p < s.size() && s[p] != q) { if (s[p] == '\\' && p + 1 < s.size()) ++p; out += s[p++]; }
      CHECK(p < s.size()) << "unterminated string at line " << line;
      ++p; return out;
    }
    const size_t b = p;
    while (p < s.size() && !isspace((unsigned char)s[p]) && s[p] != ':' && s[p] != '{' && s[p] != '}' && s[p] != '#') ++p;
    CHECK(p > b) << "parse error at line " << line << " near '" << s.substr(b, 10) << "'";
    return s.substr(b, p - b);
  }
};
shared_ptr<PbMsg> parse_msg(Lexer& lx, bool top) {
  auto m = std::make_shared<PbMsg>();
  for (;;) {
    if (lx.eof()) { CHECK(top) << "missing '}'"; return m; }
    if (lx.peek() == '}') { CHECK(!top) << "unbalanced '}' at line " << lx.line; ++lx.p; return m; }
    const string key = lx.token();
    char c = lx.peek();
    if (c == ':') { ++lx.p; c = lx.peek(); }
    if (c == '{') { ++lx.p; m->fields.push_back(PbField{key, "", parse_msg(lx, false)}); }
    else m->fields.push_back(PbField{key, lx.token(), nullptr});
  }
}
}  // namespace

static void print_msg(const PbMsg& m, int indent, std::ostringstream& os) {
  const string pad(indent, ' ');
  for (const PbField& f : m.fields) {
    if (f.msg) { os << pad << f.key << " {\n"; print_msg(*f.msg, indent + 2, os); os << pad << "}\n"; }
    else {
      char* end = nullptr; strtod(f.scalar.c_str(), &end);
      const bool bare = (!f.scalar.empty() && end && *end == 0) || f.key == "type" || f.key == "operation" || f.key == "norm" ||
                        f.key == "phase" || f.key == "backend" || f.key == "context_type" || f.key == "solver_mode" ||
                        f.scalar == "true" || f.scalar == "false";
      // `type` is an enum identifier in LayerParameter but a string in FillerParameter
      const bool quoted_type = f.key == "type" && LayerTypeFromName(f.scalar) == LayerParameter_LayerType_NONE;
      if (bare && !quoted_type) os << pad << f.key << ": " << f.scalar << "\n";
      else os << pad << f.key << ": \"" << f.scalar << "\"\n";
    }
  }
}
string PrintTextFormat(const PbMsg& m) { std::ostringstream os; print_msg(m, 0, os); return os.str(); }

shared_ptr<PbMsg> ParseTextFormat(const string& text) { Lexer lx(text); return parse_msg(lx, true); }
string ReadFileOrDie(const string& path) {
  std::ifstream f(path.c_str(), std::ios::binary);
  CHECK(f.good()) << "File not found: " << path;
  std::ostringstream ss; ss << f.rdbuf(); return ss.str();
}
NetParameter ReadNetParamsFromTextFileOrDie(const string& path) { return NetParameter(ParseTextFormat(ReadFileOrDie(path))); }
SolverParameter ReadSolverParamsFromTextFileOrDie(const string& path) { return SolverParameter(ParseTextFormat(ReadFileOrDie(path))); }

static const struct { const char* name; LayerParameter_LayerType t; } kTypes[] = {
  {"CONCAT", LayerParameter_LayerType_CONCAT}, {"DROPOUT", LayerParameter_LayerType_DROPOUT},
  {"FLATTEN", LayerParameter_LayerType_FLATTEN}, {"INNER_PRODUCT", LayerParameter_LayerType_INNER_PRODUCT},
  {"RELU", LayerParameter_LayerType_RELU}, {"SPLIT", LayerParameter_LayerType_SPLIT},
  {"ELTWISE", LayerParameter_LayerType_ELTWISE}, {"SLICE", LayerParameter_LayerType_SLICE},
  {"NORMALIZATION", LayerParameter_LayerType_NORMALIZATION}, {"MAX_MARGIN_LOSS", LayerParameter_LayerType_MAX_MARGIN_LOSS},
  {"SUM", LayerParameter_LayerType_SUM}, {"VIDEO_SAMPLED_SHOTS_DATA", LayerParameter_LayerType_VIDEO_SAMPLED_SHOTS_DATA},
  {"RETRIEVAL_STATS", LayerParameter_LayerType_RETRIEVAL_STATS}, {"ID_TO_WEIGHT_MAPPING", LayerParameter_LayerType_ID_TO_WEIGHT_MAPPING}, {"VIDEO_SHOT_WINDOW_TEST_DATA", LayerParameter_LayerType_VIDEO_SHOT_WINDOW_TEST_DATA}};
LayerParameter_LayerType LayerTypeFromName(const string& name) {
  for (auto& e : kTypes) if (name == e.name) return e.t;
  return LayerParameter_LayerType_NONE;
}
const char* LayerTypeName(LayerParameter_LayerType t) {
  for (auto& e : kTypes) if (t == e.t) return e.name;
  return "NONE";
}
EltwiseParameter_EltwiseOp EltwiseParameter::operation() const {
  const string op = m->str("operation", "SUM");
  if (op == "PROD") return EltwiseParameter_EltwiseOp_PROD;
  if (op == "MAX") return EltwiseParameter_EltwiseOp_MAX;
  return EltwiseParameter_EltwiseOp_SUM;
}
VideoSampledShotsDataParameter_ContextType VideoSampledShotsDataParameter::context_type() const {
  const string t = m->str("context_type", "PAIRWISE");            // [default = PAIRWISE], caffe.proto:606
  if (t == "PAIRWISE") return VideoSampledShotsDataParameter_CONTEXT_PAIRWISE;
  if (t == "WINDOW") return VideoSampledShotsDataParameter_CONTEXT_WINDOW;
  if (t == "PAST") return VideoSampledShotsDataParameter_CONTEXT_PAST;
  if (t == "PAST_CONTINUOUS") return VideoSampledShotsDataParameter_CONTEXT_PAST_CONTINUOUS;
  if (t == "PAST_CONTINUOUS_FIXED") return VideoSampledShotsDataParameter_CONTEXT_PAST_CONTINUOUS_FIXED;
  LOG_FATAL << "Unknown context type " << t;
  return VideoSampledShotsDataParameter_CONTEXT_PAIRWISE;
}

}  // namespace caffe
