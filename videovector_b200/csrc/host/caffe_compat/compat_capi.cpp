// C entry points over the caffe_compat host for the Python tests: build a net / solver from prototxt text,
// run it layer by layer or fused, read blobs by name.  Errors (failed CHECKs) come back as -1 + message.
#include <cuda_runtime_api.h>
#include <cstring>
#include <map>
#include "caffe/solver.hpp"

using namespace caffe;

static thread_local string g_err;
#define GUARD(body) try { body } catch (const std::exception& e) { g_err = e.what(); return -1; }
#define GUARDP(body) try { body } catch (const std::exception& e) { g_err = e.what(); return nullptr; }

extern "C" {

const char* vvc_last_error() { return g_err.c_str(); }
int vvc_set_device(int dev) { GUARD(Caffe::SetDevice(dev); return 0;) }
int vvc_device_synchronize() { return int(cudaDeviceSynchronize()); }
int vvc_set_seed(unsigned seed) { Caffe::set_random_seed(seed); return 0; }
int vvc_set_precision(int prec) { Caffe::set_precision(prec); return 0; }
int vvc_set_stream(void* s) { Caffe::set_stream(reinterpret_cast<vv_stream_t>(s)); return 0; }

// FilterNet(phase) + InsertSplits on prototxt text (no device needed); result written to out (NUL-terminated)
int vvc_transform_net(const char* prototxt_text, int phase, char* out, int out_len) {
  GUARD(const NetParameter p = InsertSplits(FilterNet(NetParameter(ParseTextFormat(prototxt_text)), phase == 1 ? Caffe::TEST : Caffe::TRAIN));
        const string s = PrintTextFormat(*p.m);
        CHECK_LT(int(s.size()), out_len) << "output buffer too small";
        memcpy(out, s.c_str(), s.size() + 1); return int(s.size());)
}

// One layer by itself, the way the reference's per-layer tests drive a layer (src/caffe/test/test_*_layer.cpp): bottoms
// filled from host arrays, SetUp, Forward, Backward.  net_text: a NetParameter with one `layers { }` entry.
// forwards: how many times Forward runs before the tops are read (a data layer's n-th batch).
// bottom_shapes: 4 ints per bottom; top_data / bottom_diff: host buffers of at least `cap` floats each (entries may be
// NULL); top_diff: what to seed the top diffs with before Backward (NULL: what SetUp left there, i.e. the loss weights).
int vvc_layer_run(const char* net_text, int n_bottom, const int* bottom_shapes, const float* const* bottom_data, int n_top,
                  int cap, float* const* top_data, int* top_counts, const float* const* top_diff, const int* propagate_down,
                  float* const* bottom_diff, float* loss, int forwards) {
  try {
    const NetParameter np(ParseTextFormat(net_text));
    CHECK_EQ(np.layers_size(), 1) << "vvc_layer_run takes exactly one layer";
    shared_ptr<Layer<float> > layer(GetLayer<float>(np.layers(0)));
    vector<shared_ptr<Blob<float> > > hold;
    vector<Blob<float>*> bottom, top;
    for (int i = 0; i < n_bottom; ++i) {
      const int* sh = bottom_shapes + 4 * i;
      hold.push_back(shared_ptr<Blob<float> >(new Blob<float>(sh[0], sh[1], sh[2], sh[3])));
      memcpy(hold.back()->mutable_cpu_data(), bottom_data[i], sizeof(float) * hold.back()->count());
      bottom.push_back(hold.back().get());
    }
    for (int i = 0; i < n_top; ++i) { hold.push_back(shared_ptr<Blob<float> >(new Blob<float>())); top.push_back(hold.back().get()); }
    layer->SetUp(bottom, &top);
    float l = 0.f;
    for (int f = 0; f < (forwards < 1 ? 1 : forwards); ++f) l = layer->Forward(bottom, &top);   // data layers: the f-th batch
    if (loss) *loss = l;
    for (int i = 0; i < n_top; ++i) {
      CHECK_LE(top[i]->count(), cap) << "top " << i << " does not fit the output buffer";
      if (top_counts) top_counts[i] = top[i]->count();
      if (top_data && top_data[i]) memcpy(top_data[i], top[i]->cpu_data(), sizeof(float) * top[i]->count());
      if (top_diff && top_diff[i]) memcpy(top[i]->mutable_cpu_diff(), top_diff[i], sizeof(float) * top[i]->count());
    }
    if (propagate_down) {
      vector<bool> pd(n_bottom);
      for (int i = 0; i < n_bottom; ++i) pd[i] = propagate_down[i] != 0;
      layer->Backward(top, pd, &bottom);
      for (int i = 0; i < n_bottom; ++i) {
        if (!(bottom_diff && bottom_diff[i] && pd[i])) continue;
        CHECK_LE(bottom[i]->count(), cap) << "bottom " << i << " does not fit the output buffer";
        memcpy(bottom_diff[i], bottom[i]->cpu_diff(), sizeof(float) * bottom[i]->count());
      }
    }
    cudaDeviceSynchronize();
    return 0;
  } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

void* vvc_net_create(const char* prototxt_text, int phase) {
  GUARDP(return new Net<float>(NetParameter(ParseTextFormat(prototxt_text)), phase == 1 ? Caffe::TEST : Caffe::TRAIN);)
}
void vvc_net_destroy(void* n) { delete static_cast<Net<float>*>(n); }
int vvc_net_num_layers(void* n) { return static_cast<Net<float>*>(n)->layers().size(); }
const char* vvc_net_layer_name(void* n, int i) { return static_cast<Net<float>*>(n)->layer_names()[i].c_str(); }
int vvc_net_layer_type(void* n, int i) { return int(static_cast<Net<float>*>(n)->layers()[i]->type()); }
int vvc_net_layer_need_backward(void* n, int i) { return static_cast<Net<float>*>(n)->layer_need_backward()[i] ? 1 : 0; }
int vvc_net_num_blobs(void* n) { return static_cast<Net<float>*>(n)->blobs().size(); }
const char* vvc_net_blob_name(void* n, int i) { return static_cast<Net<float>*>(n)->blob_names()[i].c_str(); }
int vvc_net_num_params(void* n) { return static_cast<Net<float>*>(n)->params().size(); }

// shape[4] out; returns count or -1
int vvc_net_blob_shape(void* n, const char* name, int* shape) {
  GUARD(auto b = static_cast<Net<float>*>(n)->blob_by_name(name);
        shape[0] = b->num(); shape[1] = b->channels(); shape[2] = b->height(); shape[3] = b->width(); return b->count();)
}
// copies data (which = 0) or diff (which = 1) of a named blob to host memory
int vvc_net_blob_read(void* n, const char* name, int which, float* out) {
  GUARD(auto b = static_cast<Net<float>*>(n)->blob_by_name(name);
        memcpy(out, which ? b->cpu_diff() : b->cpu_data(), sizeof(float) * b->count()); return 0;)
}
int vvc_net_param_count(void* n, int i) { return static_cast<Net<float>*>(n)->params()[i]->count(); }
int vvc_net_param_read(void* n, int i, int which, float* out) {
  GUARD(auto b = static_cast<Net<float>*>(n)->params()[i];
        memcpy(out, which ? b->cpu_diff() : b->cpu_data(), sizeof(float) * b->count()); return 0;)
}
int vvc_net_param_write(void* n, int i, const float* in) {
  GUARD(auto b = static_cast<Net<float>*>(n)->params()[i];
        CHECK(!static_cast<Net<float>*>(n)->trainer()) << "write parameters before the first fused step";
        memcpy(b->mutable_cpu_data(), in, sizeof(float) * b->count()); return 0;)
}
// explicit dropout mask (device, 0/1 uint32 [R*B, N]) for parity runs: applies to the DROPOUT layer and the fused path
int vvc_net_set_dropout_mask(void* n, const void* device_mask01) {
  GUARD(Net<float>* net = static_cast<Net<float>*>(n);
        for (auto& l : net->layers())
          if (auto* d = dynamic_cast<DropoutLayer<float>*>(l.get())) d->set_fixed_mask(static_cast<const uint32_t*>(device_mask01));
        net->set_fixed_dropout_mask(static_cast<const uint32_t*>(device_mask01)); return 0;)
}
int vvc_net_enable_fusion(void* n, char* why, int why_len) {
  GUARD(string w; const bool ok = static_cast<Net<float>*>(n)->EnableFusion(&w);
        if (why && why_len > 0) { strncpy(why, w.c_str(), why_len - 1); why[why_len - 1] = 0; }
        return ok ? 1 : 0;)
}
int vvc_net_forward_backward(void* n, float* loss) {
  GUARD(*loss = static_cast<Net<float>*>(n)->ForwardBackward(); cudaDeviceSynchronize(); return 0;)
}
int vvc_net_forward(void* n, float* loss) {
  GUARD(static_cast<Net<float>*>(n)->ForwardPrefilled(loss); cudaDeviceSynchronize(); return 0;)
}

void* vvc_solver_create(const char* solver_text, const char* net_text) {
  // `net:` may name a file, or the net prototxt is given inline (written to a temp file next to nothing: kept in memory)
  GUARDP(
    auto sp = ParseTextFormat(solver_text);
    if (net_text && *net_text) {
      char path[] = "/tmp/vvc_net_XXXXXX";
      int fd = mkstemp(path); CHECK_GE(fd, 0) << "mkstemp failed";
      FILE* f = fdopen(fd, "w"); fputs(net_text, f); fclose(f);
      sp->set_scalar("net", path);
      Solver<float>* s = GetSolver<float>(SolverParameter(sp));
      remove(path);
      return s;
    }
    return GetSolver<float>(SolverParameter(sp));)
}
void vvc_solver_destroy(void* s) { delete static_cast<Solver<float>*>(s); }
void* vvc_solver_net(void* s) { return static_cast<Solver<float>*>(s)->net().get(); }
int vvc_solver_step(void* s, float* loss) { GUARD(*loss = static_cast<Solver<float>*>(s)->Step(); cudaDeviceSynchronize(); return 0;) }
int vvc_solver_solve(void* s, int max_iter) { GUARD(static_cast<Solver<float>*>(s)->Solve(max_iter); cudaDeviceSynchronize(); return 0;) }
int vvc_solver_solve_resume(void* s, int max_iter, const char* state_path) {
  GUARD(static_cast<Solver<float>*>(s)->Solve(max_iter, state_path); cudaDeviceSynchronize(); return 0;)
}
// runs Solver::Test(test_net_id): returns the number of output values, copies up to cap of their means to out
int vvc_solver_test(void* s, int test_net_id, float* out, int cap) {
  GUARD(Solver<float>* sol = static_cast<Solver<float>*>(s);
        CHECK_LT(test_net_id, int(sol->test_nets().size())) << "no such test net";
        const vector<float> r = sol->Test(test_net_id); cudaDeviceSynchronize();
        for (int i = 0; i < int(r.size()) && i < cap; ++i) out[i] = r[i];
        return int(r.size());)
}
void* vvc_solver_test_net(void* s, int i) {
  Solver<float>* sol = static_cast<Solver<float>*>(s);
  return i < int(sol->test_nets().size()) ? sol->test_nets()[i].get() : nullptr;
}
int vvc_net_share_trained_layers_with(void* n, void* other) {
  GUARD(static_cast<Net<float>*>(n)->ShareTrainedLayersWith(static_cast<Net<float>*>(other)); return 0;)
}
int vvc_set_phase(int phase) { Caffe::set_phase(phase == 1 ? Caffe::TEST : Caffe::TRAIN); return 0; }
int vvc_solver_iter(void* s) { return static_cast<Solver<float>*>(s)->iter(); }
int vvc_solver_history_read(void* s, int i, float* out) {
  GUARD(auto* sg = dynamic_cast<SGDSolver<float>*>(static_cast<Solver<float>*>(s)); CHECK(sg);
        CHECK_LT(i, int(sg->history().size())) << "history is allocated by the first step";
        memcpy(out, sg->history()[i]->cpu_data(), sizeof(float) * sg->history()[i]->count()); return 0;)
}
// ---- trained-parameter IO: .caffemodel / .solverstate (binary protobuf, caffe_compat/wire.cpp) ----------------------
int vvc_net_save(void* n, const char* path, int write_diff) {
  GUARD(WriteProtoToBinaryFile(*static_cast<Net<float>*>(n)->ToProto(write_diff != 0), "NetParameter", path); return 0;)
}
int vvc_net_copy_trained_from(void* n, const char* path) { GUARD(static_cast<Net<float>*>(n)->CopyTrainedLayersFrom(string(path)); return 0;) }
int vvc_solver_snapshot(void* s, char* model_path_out, int cap) {
  GUARD(const string m = static_cast<Solver<float>*>(s)->Snapshot();
        if (model_path_out && cap > 0) { strncpy(model_path_out, m.c_str(), cap - 1); model_path_out[cap - 1] = 0; }
        return 0;)
}
int vvc_solver_restore(void* s, const char* state_path) { GUARD(static_cast<Solver<float>*>(s)->Restore(state_path); cudaDeviceSynchronize(); return 0;) }

// Generic access to a binary message file for tests and tools (no device needed):
//   open -> handle; text = the message in text format without the float arrays; blobs = every packed float array in
//   document order with the dotted path of its owner ("layers[3].blobs[0].data", "history[1].data").
struct PbFile { shared_ptr<PbMsg> root; string type; vector<std::pair<string, shared_ptr<vector<float> > > > arrays; };
static void collect_arrays(const PbMsg& m, const string& prefix, PbFile* f) {
  std::map<string, int> seen;
  for (const PbField& fld : m.fields) {
    const int k = seen[fld.key]++;
    const string here = prefix + (prefix.empty() ? "" : ".") + fld.key;
    if (fld.floats) f->arrays.push_back(std::make_pair(here, fld.floats));
    else if (fld.msg) collect_arrays(*fld.msg, here + "[" + std::to_string(k) + "]", f);
  }
}
static shared_ptr<PbMsg> strip_arrays(const PbMsg& m) {
  auto out = std::make_shared<PbMsg>();
  for (const PbField& fld : m.fields) {
    if (fld.floats) continue;
    out->fields.push_back(PbField{fld.key, fld.scalar, fld.msg ? strip_arrays(*fld.msg) : nullptr, nullptr});
  }
  return out;
}
void* vvc_pb_open(const char* path, const char* type) {
  GUARDP(PbFile* f = new PbFile; f->type = type; f->root = ReadProtoFromBinaryFile(path, type); collect_arrays(*f->root, "", f); return f;)
}
void vvc_pb_close(void* h) { delete static_cast<PbFile*>(h); }
int vvc_pb_text(void* h, char* out, int cap) {
  GUARD(const string s = PrintTextFormat(*strip_arrays(*static_cast<PbFile*>(h)->root));
        CHECK_LT(int(s.size()), cap) << "output buffer too small"; memcpy(out, s.c_str(), s.size() + 1); return int(s.size());)
}
int vvc_pb_num_arrays(void* h) { return int(static_cast<PbFile*>(h)->arrays.size()); }
int vvc_pb_array_info(void* h, int i, char* path_out, int cap) {
  GUARD(PbFile* f = static_cast<PbFile*>(h); CHECK_LT(i, int(f->arrays.size()));
        strncpy(path_out, f->arrays[i].first.c_str(), cap - 1); path_out[cap - 1] = 0; return int(f->arrays[i].second->size());)
}
int vvc_pb_array_read(void* h, int i, float* out) {
  GUARD(PbFile* f = static_cast<PbFile*>(h); CHECK_LT(i, int(f->arrays.size()));
        memcpy(out, f->arrays[i].second->data(), sizeof(float) * f->arrays[i].second->size()); return 0;)
}
// Writes message `type` from its text form plus float arrays attached by dotted path (the parent message must exist in the text).
int vvc_pb_write(const char* path, const char* type, const char* text, int n_arrays, const char** array_paths,
                 const float** arrays, const int* counts) {
  GUARD(
    shared_ptr<PbMsg> root = ParseTextFormat(text);
    for (int a = 0; a < n_arrays; ++a) {
      // walk "layers[3].blobs[0].data"
      PbMsg* cur = root.get();
      string p = array_paths[a];
      size_t pos = 0;
      for (;;) {
        const size_t dot = p.find('.', pos);
        const string part = p.substr(pos, dot == string::npos ? string::npos : dot - pos);
        if (dot == string::npos) {
          cur->fields.push_back(PbField{part, "", nullptr, std::make_shared<vector<float> >(arrays[a], arrays[a] + counts[a])});
          break;
        }
        const size_t lb = part.find('[');
        CHECK(lb != string::npos) << "bad array path " << p;
        const string key = part.substr(0, lb);
        const int idx = atoi(part.c_str() + lb + 1);
        const PbField* f = cur->nth(key, idx);
        CHECK(f && f->msg) << "array path " << p << ": no message " << part;
        cur = f->msg.get();
        pos = dot + 1;
      }
    }
    WriteProtoToBinaryFile(*root, type, path);
    return 0;)
}

float vvc_solver_learning_rate(void* s) {
  try { auto* sg = dynamic_cast<SGDSolver<float>*>(static_cast<Solver<float>*>(s)); return sg->GetLearningRate(); }
  catch (const std::exception& e) { g_err = e.what(); return -1.f; }
}

}  // extern "C"
