// ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.  Drives the REFERENCE's own layer classes (compiled unmodified from
// /root/reference/src/caffe/layers/*.cpp against the shim headers in oracle/ref_shim/include) through the TRAIN
// graph of projects/videovec_embedding/mednet_embedding_train.prototxt, in Caffe CPU mode, wired in the order
// Net::ForwardFromTo / BackwardFromTo would run it.  Used to (a) validate oracle/vv_oracle.cpp and generate the
// golden vectors under tests/golden/, (b) serve as cpu_baseline kind "reference" in bench.py.
// Nothing here is reference source: only its public class interface is used.
#include <chrono>
#include <cstring>
#include <memory>
#include <vector>
#include "caffe/blob.hpp"
#include "caffe/common.hpp"
#include "caffe/vision_layers.hpp"
#include "caffe/data_layers.hpp"
#include "caffe/util/io.hpp"
#include "caffe/net.hpp"
#include "caffe/solver.hpp"
#include "caffe/util/upgrade_proto.hpp"
#include "caffe/util/pb2json.h"

using namespace caffe;  // NOLINT
typedef std::vector<Blob<float>*> BV;

namespace {
struct DropoutPeek : public DropoutLayer<float> {
  explicit DropoutPeek(const LayerParameter& p) : DropoutLayer<float>(p) {}
  const unsigned int* mask() { return this->rand_vec_.cpu_data(); }
};
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
std::unique_ptr<Blob<float> > nb() { return std::unique_ptr<Blob<float> >(new Blob<float>()); }
}  // namespace

#define REF_API __attribute__((visibility("default")))
extern "C" {

REF_API const char* ref_describe() { return "reference layer classes (eevignesh/videovector src/caffe/layers/*.cpp), CPU mode, shim-compiled"; }

// Whole TRAIN net forward + backward.  data [B,R,K]; outputs may be NULL.  mask_out receives the 0/1 mask the
// reference's DropoutLayer drew (caffe_rng_bernoulli) so the other implementations can replay it.
REF_API int ref_net_forward_backward(int B, int C, int Nn, int K, int N, float margin, int norm, float dropout_ratio,
                             const float* data, const float* W, const float* bias, unsigned seed,
                             float* loss_out, float* viol_out, float* dW, float* db, unsigned* mask_out,
                             float* H_out, float* dZ_out, float* tscore_out, float* nscore_out, double* seconds) {
  try {
    Caffe::set_mode(Caffe::CPU);
    Caffe::set_phase(Caffe::TRAIN);
    Caffe::set_random_seed(seed);
    const int R = C + Nn;
    const double t0 = now();
    std::vector<std::unique_ptr<Blob<float> > > keep;
    auto blob = [&]() { keep.push_back(nb()); return keep.back().get(); };
    std::vector<std::unique_ptr<Layer<float> > > layers;
    struct Node { Layer<float>* l; BV bottom, top; std::vector<bool> pd; };
    std::vector<Node> g;
    auto add = [&](Layer<float>* l, BV bottom, BV top, std::vector<bool> pd) {
      layers.emplace_back(l);
      g.push_back(Node{l, bottom, top, pd});
      l->SetUp(g.back().bottom, &g.back().top);
    };
    Blob<float>* d = blob(); d->Reshape(B, R, K, 1);
    memcpy(d->mutable_cpu_data(), data, sizeof(float) * size_t(B) * R * K);
    // slice_input_data (dim 1), batch_concat_input (dim 0), flatten_input
    BV raw; for (int j = 0; j < R; ++j) raw.push_back(blob());
    { LayerParameter p; p.mutable_slice_param()->set_slice_dim(1); add(new SliceLayer<float>(p), BV{d}, raw, {false}); }
    Blob<float>* cat = blob();
    { LayerParameter p; p.mutable_concat_param()->set_concat_dim(0); add(new ConcatLayer<float>(p), raw, BV{cat}, std::vector<bool>(R, false)); }
    Blob<float>* X = blob();
    { LayerParameter p; add(new FlattenLayer<float>(p), BV{cat}, BV{X}, {false}); }
    // fc7
    Blob<float>* Z = blob();
    InnerProductLayer<float>* ip;
    { LayerParameter p; p.mutable_inner_product_param()->set_num_output(N);
      p.mutable_inner_product_param()->mutable_weight_filler()->set_type("constant");
      p.mutable_inner_product_param()->mutable_bias_filler()->set_type("constant");
      ip = new InnerProductLayer<float>(p); add(ip, BV{X}, BV{Z}, {false}); }
    memcpy(ip->blobs()[0]->mutable_cpu_data(), W, sizeof(float) * size_t(N) * K);
    memcpy(ip->blobs()[1]->mutable_cpu_data(), bias, sizeof(float) * N);
    // fc7_relu, drop2 (in place)
    Blob<float>* H = blob();
    { LayerParameter p; add(new ReLULayer<float>(p), BV{Z}, BV{H}, {true}); }
    DropoutPeek* drop = nullptr;
    if (dropout_ratio > 0.f) {
      LayerParameter p; p.mutable_dropout_param()->set_dropout_ratio(dropout_ratio);
      drop = new DropoutPeek(p); add(drop, BV{H}, BV{H}, {true});
    }
    // slice_emb (dim 0)
    BV emb; for (int j = 0; j < R; ++j) emb.push_back(blob());
    { LayerParameter p; p.mutable_slice_param()->set_slice_dim(0); add(new SliceLayer<float>(p), BV{H}, emb, {true}); }
    // context_average, word_embedding_norm, split
    BV ctx(emb.begin() + 1, emb.begin() + C);
    Blob<float>* cbar = blob(); Blob<float>* chat = blob();
    { LayerParameter p; p.mutable_eltwise_param()->set_operation(EltwiseParameter_EltwiseOp_SUM);
      for (int i = 0; i < C - 1; ++i) p.mutable_eltwise_param()->add_coeff(1.f / float(C - 1));
      add(new EltwiseLayer<float>(p), ctx, BV{cbar}, std::vector<bool>(C - 1, true)); }
    { LayerParameter p; add(new NormalizationLayer<float>(p), BV{cbar}, BV{chat}, {true}); }
    BV sp; for (int k = 0; k <= Nn; ++k) sp.push_back(blob());
    { LayerParameter p; add(new SplitLayer<float>(p), BV{chat}, sp, {true}); }
    // concat_pos_neg_nonorm, pos_neg_normalize, slice_pos_neg_norm
    BV pn; pn.push_back(emb[0]); for (int k = 0; k < Nn; ++k) pn.push_back(emb[C + k]);
    Blob<float>* P = blob(); Blob<float>* Ph = blob();
    { LayerParameter p; p.mutable_concat_param()->set_concat_dim(0); add(new ConcatLayer<float>(p), pn, BV{P}, std::vector<bool>(1 + Nn, true)); }
    { LayerParameter p; add(new NormalizationLayer<float>(p), BV{P}, BV{Ph}, {true}); }
    BV pnn; for (int k = 0; k <= Nn; ++k) pnn.push_back(blob());
    { LayerParameter p; p.mutable_slice_param()->set_slice_dim(0); add(new SliceLayer<float>(p), BV{Ph}, pnn, {true}); }
    // prod_true/sum_true, prod_neg_k/sum_neg_k
    Blob<float>* tscore = blob(); BV nsc;
    for (int k = 0; k <= Nn; ++k) {
      Blob<float>* prod = blob();
      { LayerParameter p; p.mutable_eltwise_param()->set_operation(EltwiseParameter_EltwiseOp_PROD);
        add(new EltwiseLayer<float>(p), BV{sp[k], pnn[k]}, BV{prod}, {true, true}); }
      Blob<float>* sc = (k == 0) ? tscore : blob();
      if (k) nsc.push_back(sc);
      { LayerParameter p; p.mutable_sum_param()->set_num_output(k == 0 ? Nn : 1); add(new SumLayer<float>(p), BV{prod}, BV{sc}, {true}); }
    }
    Blob<float>* nscore = blob();
    { LayerParameter p; p.mutable_concat_param()->set_concat_dim(1); add(new ConcatLayer<float>(p), nsc, BV{nscore}, std::vector<bool>(Nn, true)); }
    Blob<float>* loss = blob(); Blob<float>* viol = blob();
    { LayerParameter p; p.add_loss_weight(1.f); p.add_loss_weight(0.f);
      p.mutable_max_margin_loss_param()->set_margin(margin);
      p.mutable_max_margin_loss_param()->set_norm(norm == 2 ? MaxMarginLossParameter_Norm_L2 : MaxMarginLossParameter_Norm_L1);
      add(new MaxMarginLossLayer<float>(p), BV{tscore, nscore}, BV{loss, viol}, {true, true}); }
    const double t1 = now();
    // forward (Reshape + Forward per layer, net.cpp:508-509), backward in reverse; the first three layers need none
    float total = 0.f;
    for (auto& n : g) { n.l->Reshape(n.bottom, &n.top); total += n.l->Forward(n.bottom, &n.top); }
    const double t2 = now();
    for (int i = int(g.size()) - 1; i >= 3; --i) g[i].l->Backward(g[i].top, g[i].pd, &g[i].bottom);
    const double t3 = now();
    if (loss_out) *loss_out = total;
    if (viol_out) *viol_out = viol->cpu_data()[0];
    if (dW) memcpy(dW, ip->blobs()[0]->cpu_diff(), sizeof(float) * size_t(N) * K);
    if (db) memcpy(db, ip->blobs()[1]->cpu_diff(), sizeof(float) * N);
    if (mask_out && drop) memcpy(mask_out, drop->mask(), sizeof(unsigned) * size_t(R) * B * N);
    if (H_out) memcpy(H_out, H->cpu_data(), sizeof(float) * size_t(R) * B * N);
    if (dZ_out) memcpy(dZ_out, Z->cpu_diff(), sizeof(float) * size_t(R) * B * N);
    if (tscore_out) memcpy(tscore_out, tscore->cpu_data(), sizeof(float) * size_t(B) * Nn);
    if (nscore_out) memcpy(nscore_out, nscore->cpu_data(), sizeof(float) * size_t(B) * Nn);
    if (seconds) { seconds[0] = t1 - t0; seconds[1] = t2 - t1; seconds[2] = t3 - t2; }
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "ref_driver: %s\n", e.what());
    return -1;
  }
}

// Single layers, for pinning the oracle's restatements one by one.
REF_API int ref_normalization(int num, int dim, const float* x, const float* dy, float* y, float* dx) {
  try {
    Caffe::set_mode(Caffe::CPU);
    Blob<float> b(num, dim, 1, 1), t;
    memcpy(b.mutable_cpu_data(), x, sizeof(float) * num * dim);
    LayerParameter p; NormalizationLayer<float> l(p);
    BV bv{&b}, tv{&t};
    l.SetUp(bv, &tv); l.Forward(bv, &tv);
    memcpy(y, t.cpu_data(), sizeof(float) * num * dim);
    if (dy && dx) {
      memcpy(t.mutable_cpu_diff(), dy, sizeof(float) * num * dim);
      l.Backward(tv, {true}, &bv);
      memcpy(dx, b.cpu_diff(), sizeof(float) * num * dim);
    }
    return 0;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); return -1; }
}
REF_API int ref_max_margin(int num, int ch, const float* s_true, const float* s_bogus, float margin, int norm, float loss_weight,
                   float* loss, float* viol, float* d_true, float* d_bogus) {
  try {
    Caffe::set_mode(Caffe::CPU);
    Blob<float> a(num, ch, 1, 1), b(num, ch, 1, 1), l, v;
    memcpy(a.mutable_cpu_data(), s_true, sizeof(float) * num * ch);
    memcpy(b.mutable_cpu_data(), s_bogus, sizeof(float) * num * ch);
    LayerParameter p; p.add_loss_weight(loss_weight); p.add_loss_weight(0.f);
    p.mutable_max_margin_loss_param()->set_margin(margin);
    p.mutable_max_margin_loss_param()->set_norm(norm == 2 ? MaxMarginLossParameter_Norm_L2 : MaxMarginLossParameter_Norm_L1);
    MaxMarginLossLayer<float> layer(p);
    BV bv{&a, &b}, tv{&l, &v};
    layer.SetUp(bv, &tv);
    layer.Forward(bv, &tv);
    *loss = l.cpu_data()[0]; *viol = v.cpu_data()[0];
    layer.Backward(tv, {true, true}, &bv);
    memcpy(d_true, a.cpu_diff(), sizeof(float) * num * ch);
    memcpy(d_bogus, b.cpu_diff(), sizeof(float) * num * ch);
    return 0;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); return -1; }
}
// MaxMarginLossLayer with its optional third bottom (max_margin_loss_layer.cpp:18-39,79-97,150-186): `third` holds the
// per-element weights (direct != 0) or video ids looked up in the "id,weight" file
REF_API int ref_max_margin_w(int num, int ch, const float* s_true, const float* s_bogus, const float* third, int direct,
                             const char* id_to_weight_file, float margin, int norm, float loss_weight,
                             float* loss, float* viol, float* d_true, float* d_bogus) {
  try {
    Caffe::set_mode(Caffe::CPU);
    Blob<float> a(num, ch, 1, 1), b(num, ch, 1, 1), c(num, ch, 1, 1), l, v;
    memcpy(a.mutable_cpu_data(), s_true, sizeof(float) * num * ch);
    memcpy(b.mutable_cpu_data(), s_bogus, sizeof(float) * num * ch);
    memcpy(c.mutable_cpu_data(), third, sizeof(float) * num * ch);
    LayerParameter p; p.add_loss_weight(loss_weight); p.add_loss_weight(0.f);
    p.mutable_max_margin_loss_param()->set_margin(margin);
    p.mutable_max_margin_loss_param()->set_norm(norm == 2 ? MaxMarginLossParameter_Norm_L2 : MaxMarginLossParameter_Norm_L1);
    p.mutable_max_margin_loss_param()->set_use_direct_weight(direct != 0);
    if (id_to_weight_file && *id_to_weight_file) p.mutable_max_margin_loss_param()->set_id_to_weight_file(id_to_weight_file);
    MaxMarginLossLayer<float> layer(p);
    BV bv{&a, &b, &c}, tv{&l, &v};
    layer.SetUp(bv, &tv);
    layer.Forward(bv, &tv);
    *loss = l.cpu_data()[0]; *viol = v.cpu_data()[0];
    layer.Backward(tv, {true, true, false}, &bv);
    memcpy(d_true, a.cpu_diff(), sizeof(float) * num * ch);
    memcpy(d_bogus, b.cpu_diff(), sizeof(float) * num * ch);
    return 0;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); return -1; }
}
REF_API int ref_inner_product(int M, int N, int K, const float* X, const float* W, const float* bias, const float* dZ, float reg,
                      float* Z, float* dW, float* db, float* dX) {
  try {
    Caffe::set_mode(Caffe::CPU);
    Blob<float> b(M, K, 1, 1), t;
    memcpy(b.mutable_cpu_data(), X, sizeof(float) * size_t(M) * K);
    LayerParameter p; p.mutable_inner_product_param()->set_num_output(N); p.mutable_inner_product_param()->set_regularization(reg);
    p.mutable_inner_product_param()->mutable_weight_filler()->set_type("constant");
    p.mutable_inner_product_param()->mutable_bias_filler()->set_type("constant");
    InnerProductLayer<float> l(p);
    BV bv{&b}, tv{&t};
    l.SetUp(bv, &tv);
    memcpy(l.blobs()[0]->mutable_cpu_data(), W, sizeof(float) * size_t(N) * K);
    memcpy(l.blobs()[1]->mutable_cpu_data(), bias, sizeof(float) * N);
    l.Forward(bv, &tv);
    memcpy(Z, t.cpu_data(), sizeof(float) * size_t(M) * N);
    if (dZ) {
      memcpy(t.mutable_cpu_diff(), dZ, sizeof(float) * size_t(M) * N);
      l.Backward(tv, {true}, &bv);
      memcpy(dW, l.blobs()[0]->cpu_diff(), sizeof(float) * size_t(N) * K);
      memcpy(db, l.blobs()[1]->cpu_diff(), sizeof(float) * N);
      memcpy(dX, b.cpu_diff(), sizeof(float) * size_t(M) * K);
    }
    return 0;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); return -1; }
}

// RetrievalStatsLayer (retrieval_stats_layer.cpp), shot level: E [B,N], video ids and an id->class text file; out = {mAP, hit@1, hit@5}
REF_API int ref_retrieval_stats(int B, int N, const float* E, const float* video_ids, const char* id_to_class_file,
                                int exclude_same_video_shots, float* out3) {
  try {
    Caffe::set_mode(Caffe::CPU);
    Blob<float> e(B, N, 1, 1), ids(B, 1, 1, 1), t0, t1, t2;
    memcpy(e.mutable_cpu_data(), E, sizeof(float) * size_t(B) * N);
    memcpy(ids.mutable_cpu_data(), video_ids, sizeof(float) * B);
    LayerParameter p;
    p.mutable_retrieval_stats_param()->set_id_to_class_file(id_to_class_file);
    p.mutable_retrieval_stats_param()->set_exclude_same_video_shots(exclude_same_video_shots != 0);
    RetrievalStatsLayer<float> l(p);
    BV bv{&e, &ids}, tv{&t0, &t1, &t2};
    l.SetUp(bv, &tv); l.Forward(bv, &tv);
    out3[0] = t0.cpu_data()[0]; out3[1] = t1.cpu_data()[0]; out3[2] = t2.cpu_data()[0];
    return 0;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); return -1; }
}
// + video_level_retrieval (shots of a video averaged first, :160-206; max_num_videos must equal the number of distinct
// videos in the batch) and stats_output_file (the per-query CSV, :146-151, 306-340)
REF_API int ref_retrieval_stats_ex(int B, int N, const float* E, const float* video_ids, const char* id_to_class_file,
                                   int exclude_same_video_shots, int video_level, int max_num_videos,
                                   const char* stats_output_file, float* out3) {
  try {
    Caffe::set_mode(Caffe::CPU);
    Blob<float> e(B, N, 1, 1), ids(B, 1, 1, 1), t0, t1, t2;
    memcpy(e.mutable_cpu_data(), E, sizeof(float) * size_t(B) * N);
    memcpy(ids.mutable_cpu_data(), video_ids, sizeof(float) * B);
    LayerParameter p;
    p.mutable_retrieval_stats_param()->set_id_to_class_file(id_to_class_file);
    p.mutable_retrieval_stats_param()->set_exclude_same_video_shots(exclude_same_video_shots != 0);
    if (video_level) {
      p.mutable_retrieval_stats_param()->set_video_level_retrieval(true);
      p.mutable_retrieval_stats_param()->set_max_num_videos(max_num_videos);
    }
    if (stats_output_file && *stats_output_file) p.mutable_retrieval_stats_param()->set_stats_output_file(stats_output_file);
    RetrievalStatsLayer<float> l(p);
    BV bv{&e, &ids}, tv{&t0, &t1, &t2};
    l.SetUp(bv, &tv); l.Forward(bv, &tv);
    out3[0] = t0.cpu_data()[0]; out3[1] = t1.cpu_data()[0]; out3[2] = t2.cpu_data()[0];
    return 0;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); return -1; }
}
// IdToWeightMappingLayer (id_to_weight_mapping_layer.cpp): table [rows,N], ids [M]; top [M,N] and the table gradient for top_diff
REF_API int ref_id_to_weight(int M, int N, int rows, const float* table, const float* ids, const float* top_diff,
                             float* top, float* table_diff) {
  try {
    Caffe::set_mode(Caffe::CPU);
    Blob<float> b(M, 1, 1, 1), t;
    memcpy(b.mutable_cpu_data(), ids, sizeof(float) * M);
    LayerParameter p;
    p.mutable_id_to_weight_mapping_param()->set_num_output(N);
    p.mutable_id_to_weight_mapping_param()->set_max_ids(rows);
    p.mutable_id_to_weight_mapping_param()->mutable_weight_filler()->set_type("constant");
    IdToWeightMappingLayer<float> l(p);
    BV bv{&b}, tv{&t};
    l.SetUp(bv, &tv);
    memcpy(l.blobs()[0]->mutable_cpu_data(), table, sizeof(float) * size_t(rows) * N);
    l.Forward(bv, &tv);
    memcpy(top, t.cpu_data(), sizeof(float) * size_t(M) * N);
    if (top_diff && table_diff) {
      memcpy(t.mutable_cpu_diff(), top_diff, sizeof(float) * size_t(M) * N);
      l.Backward(tv, {false}, &bv);
      memcpy(table_diff, l.blobs()[0]->cpu_diff(), sizeof(float) * size_t(rows) * N);
    }
    return 0;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); return -1; }
}

}  // extern "C"

// ---- the reference's own data layer over an in-memory fake LMDB --------------------------------------------------
// VideoSampledShotsDataLayer (video_sampled_shots_data_layer.cpp) is compiled unmodified; the LMDB C functions it calls
// are implemented here over a vector of live VideoShots messages (the shim's ParseFromArray copies the object a value
// points to).  This runs the reference's DataLayerSetUp (negative-buffer initialisation) and InternalThreadEntry
// (AddSamplesToTop, RandomShuffleTopids, AddToBuffer) -- i.e. the whole sampler state machine on the real libc rand().
leveldb::Options caffe::GetLevelDBOptions() { return leveldb::Options(); }
// data_transformer.cpp reads a mean file through this when transform_param.mean_file is set (never here)
bool caffe::ReadProtoFromBinaryFile(const char*, google::protobuf::Message*) { return false; }

namespace {
typedef video_shot_sentences::VideoShots VideoShots;
typedef video_shot_sentences::TestVideoShotWindows TestVideoShotWindows;
struct FakeDb { std::vector<std::shared_ptr<void> > records; };      // live VideoShots / TestVideoShotWindows messages
FakeDb* g_fake_db = nullptr;          // the dataset the next mdb_env_open serves ...
FakeDb* g_fake_db_test = nullptr;     // ... unless the source names the TEST database
FakeDb* g_fake_db_neg = nullptr;      // ... or the negative_dataset
}  // namespace
struct MDB_env { FakeDb* db; };
struct MDB_txn { MDB_env* env; };
struct MDB_cursor { FakeDb* db; size_t pos; };
extern "C" {
int mdb_env_create(MDB_env** env) { *env = new MDB_env{nullptr}; return MDB_SUCCESS; }
int mdb_env_set_mapsize(MDB_env*, size_t) { return MDB_SUCCESS; }
int mdb_env_open(MDB_env* env, const char* path, unsigned int, mdb_mode_t) {
  env->db = (path && strstr(path, "fake-lmdb-test")) ? g_fake_db_test : (path && strstr(path, "fake-lmdb-neg")) ? g_fake_db_neg : g_fake_db;
  return env->db ? MDB_SUCCESS : -1;
}
int mdb_txn_begin(MDB_env* env, MDB_txn*, unsigned int, MDB_txn** txn) { *txn = new MDB_txn{env}; return MDB_SUCCESS; }
int mdb_open(MDB_txn*, const char*, unsigned int, MDB_dbi* dbi) { *dbi = 1; return MDB_SUCCESS; }
int mdb_cursor_open(MDB_txn* txn, MDB_dbi, MDB_cursor** cursor) { *cursor = new MDB_cursor{txn->env->db, 0}; return MDB_SUCCESS; }
int mdb_cursor_get(MDB_cursor* c, MDB_val* key, MDB_val* data, MDB_cursor_op op) {
  if (op == MDB_FIRST) c->pos = 0;
  else if (op == MDB_NEXT) { if (c->pos + 1 >= c->db->records.size()) return MDB_NOTFOUND; ++c->pos; }
  if (c->db->records.empty()) return MDB_NOTFOUND;
  if (key) { key->mv_size = 0; key->mv_data = nullptr; }
  data->mv_data = c->db->records[c->pos].get();
  data->mv_size = size_t(VV_SHIM_LIVE_OBJECT);          // ParseFromArray(void*, int) sees -0x5EED
  return MDB_SUCCESS;
}
void mdb_cursor_close(MDB_cursor* c) { delete c; }
void mdb_close(MDB_env*, MDB_dbi) {}
void mdb_txn_abort(MDB_txn* t) { delete t; }
void mdb_env_close(MDB_env* e) { delete e; }
}  // extern "C"

namespace {
struct RefSampler {
  FakeDb db, neg_db;
  std::unique_ptr<VideoSampledShotsDataLayer<float> > layer;
  Blob<float> top;
  int B, R, K;
};
}  // namespace

extern "C" {
// Dataset as in the oracle's sampler: videos [V] with shots [shot_off[v], shot_off[v+1]) of `feat` [total, K].
// context_type: 0 PAIRWISE, 1 WINDOW, 2 PAST, 3 PAST_CONTINUOUS, 4 PAST_CONTINUOUS_FIXED.  Call srand(seed) first: the
// layer draws from the process-global rand() (from its prefetch thread; one batch is always prefetched ahead).
static void FillFakeDb(FakeDb* db, int V, int K, const int* video_id, const int* shot_off, const int* shot_ids, const float* feat) {
  for (int v = 0; v < V; ++v) {
    std::shared_ptr<VideoShots> rec(new VideoShots());
    rec->set_video_id(video_id[v]);
    for (int g = shot_off[v]; g < shot_off[v + 1]; ++g) {
      rec->add_shot_ids(shot_ids[g]);
      Datum* d = rec->add_shot_words();
      for (int k = 0; k < K; ++k) d->add_float_data(feat[size_t(g) * K + k]);
    }
    db->records.push_back(rec);
  }
}
REF_API void* ref_sampler_create_opts(int V, int K, const int* video_id, const int* shot_off, const int* shot_ids, const float* feat,
                                      int batch_size, int context_size, int num_negative_samples, int max_buffer_size,
                                      int negative_swap_percentage, int max_same_video_negs, int context_type,
                                      int rand_skip, unsigned caffe_seed, int negV, const int* neg_video_id,
                                      const int* neg_shot_off, const int* neg_shot_ids, const float* neg_feat);
REF_API void* ref_sampler_create(int V, int K, const int* video_id, const int* shot_off, const int* shot_ids, const float* feat,
                                 int batch_size, int context_size, int num_negative_samples, int max_buffer_size,
                                 int negative_swap_percentage, int max_same_video_negs, int context_type) {
  return ref_sampler_create_opts(V, K, video_id, shot_off, shot_ids, feat, batch_size, context_size, num_negative_samples,
                                 max_buffer_size, negative_swap_percentage, max_same_video_negs, context_type,
                                 0, 0, 0, nullptr, nullptr, nullptr, nullptr);
}
// + rand_skip (drawn by the layer from caffe_rng_rand() right after Caffe::set_random_seed(caffe_seed)) and
// negative_dataset (a second fake LMDB) -- video_sampled_shots_data_layer.cpp:137-153, 157-180
REF_API void* ref_sampler_create_opts(int V, int K, const int* video_id, const int* shot_off, const int* shot_ids, const float* feat,
                                      int batch_size, int context_size, int num_negative_samples, int max_buffer_size,
                                      int negative_swap_percentage, int max_same_video_negs, int context_type,
                                      int rand_skip, unsigned caffe_seed, int negV, const int* neg_video_id,
                                      const int* neg_shot_off, const int* neg_shot_ids, const float* neg_feat) {
  try {
    Caffe::set_mode(Caffe::CPU);
    RefSampler* s = new RefSampler();
    FillFakeDb(&s->db, V, K, video_id, shot_off, shot_ids, feat);
    if (negV > 0) FillFakeDb(&s->neg_db, negV, K, neg_video_id, neg_shot_off, neg_shot_ids, neg_feat);
    LayerParameter p;
    VideoSampledShotsDataParameter* vp = p.mutable_video_sampled_shots_data_param();
    vp->set_source("mem://fake-lmdb"); vp->set_backend(VideoSampledShotsDataParameter_DB_LMDB);
    vp->set_batch_size(batch_size); vp->set_context_size(context_size); vp->set_num_negative_samples(num_negative_samples);
    vp->set_max_buffer_size(max_buffer_size); vp->set_negative_swap_percentage(negative_swap_percentage);
    vp->set_max_same_video_negs(max_same_video_negs);
    vp->set_context_type(VideoSampledShotsDataParameter_CONTEXT(context_type));
    if (rand_skip > 0) { vp->set_rand_skip(rand_skip); Caffe::set_random_seed(caffe_seed); }
    if (negV > 0) vp->set_negative_dataset("mem://fake-lmdb-neg");
    g_fake_db = &s->db; g_fake_db_neg = negV > 0 ? &s->neg_db : nullptr;
    s->layer.reset(new VideoSampledShotsDataLayer<float>(p));
    BV bottom, tv{&s->top};
    s->layer->SetUp(bottom, &tv);            // DataLayerSetUp (negative buffer init) + first prefetch
    g_fake_db = nullptr; g_fake_db_neg = nullptr;
    s->B = batch_size; s->R = s->top.channels(); s->K = K;
    return s;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); return nullptr; }
}
// next batch: data [B, R, K] as the reference's Forward_cpu hands it to the net
REF_API int ref_sampler_next(void* h, float* data) {
  try {
    RefSampler* s = static_cast<RefSampler*>(h);
    BV bottom, tv{&s->top};
    s->layer->Forward(bottom, &tv);
    memcpy(data, s->top.cpu_data(), sizeof(float) * size_t(s->B) * s->R * s->K);
    return 0;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); return -1; }
}
REF_API int ref_sampler_rows(void* h) { return static_cast<RefSampler*>(h)->R; }

// The reference's TEST-phase data layer (video_shot_window_test_data_layer.cpp, compiled unmodified) over the same fake
// LMDB: n records of TestVideoShotWindows with ctx context / pos positive / neg negative datums each (data [n, ctx+pos+neg, K]).
struct RefTestLayer {
  FakeDb db;
  std::unique_ptr<VideoShotWindowTestDataLayer<float> > layer;
  Blob<float> top, label;
  int B, R, K;
};
REF_API void* ref_testlayer_create(int n, int ctx, int pos, int neg, int K, const float* data, const int* video_id,
                                   const int* pos_id, const int* neg_id, int batch_size, int include_pos, int include_neg) {
  try {
    Caffe::set_mode(Caffe::CPU);
    RefTestLayer* s = new RefTestLayer();
    const int rows = ctx + pos + neg;
    for (int i = 0; i < n; ++i) {
      std::shared_ptr<TestVideoShotWindows> rec(new TestVideoShotWindows());
      rec->set_video_id(video_id[i]);
      for (int r = 0; r < rows; ++r) {
        Datum* d = r < ctx ? rec->add_context_shot_words() : r < ctx + pos ? rec->add_positive_shot_words() : rec->add_negative_shot_words();
        for (int k = 0; k < K; ++k) d->add_float_data(data[(size_t(i) * rows + r) * K + k]);
        if (r >= ctx && r < ctx + pos) rec->add_positive_shot_id(pos_id[i * pos + (r - ctx)]);
        if (r >= ctx + pos) rec->add_negative_shot_id(neg_id[i * neg + (r - ctx - pos)]);
      }
      s->db.records.push_back(rec);
    }
    LayerParameter p;
    VideoShotWindowTestDataParameter* vp = p.mutable_video_shot_window_test_data_param();
    vp->set_source("mem://fake-lmdb"); vp->set_backend(VideoShotWindowTestDataParameter_DB_LMDB);
    vp->set_batch_size(batch_size); vp->set_include_positives(include_pos != 0); vp->set_include_negatives(include_neg != 0);
    g_fake_db = &s->db;
    s->layer.reset(new VideoShotWindowTestDataLayer<float>(p));
    BV bottom, tv{&s->top, &s->label};
    s->layer->SetUp(bottom, &tv);
    g_fake_db = nullptr;
    s->B = batch_size; s->R = s->top.channels(); s->K = K;
    return s;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); return nullptr; }
}
REF_API int ref_testlayer_rows(void* h) { return static_cast<RefTestLayer*>(h)->R; }
REF_API int ref_testlayer_next(void* h, float* data, float* labels) {
  try {
    RefTestLayer* s = static_cast<RefTestLayer*>(h);
    BV bottom, tv{&s->top, &s->label};
    s->layer->Forward(bottom, &tv);
    memcpy(data, s->top.cpu_data(), sizeof(float) * size_t(s->B) * s->R * s->K);
    memcpy(labels, s->label.cpu_data(), sizeof(float) * size_t(s->B));
    return 0;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); return -1; }
}
REF_API void ref_testlayer_destroy(void* h) { delete static_cast<RefTestLayer*>(h); }
REF_API void ref_sampler_destroy(void* h) { delete static_cast<RefSampler*>(h); }
REF_API void ref_srand(unsigned seed) { srand(seed); }

}  // extern "C"


// ---- the reference's own Net + SGDSolver (net.cpp, solver.cpp, util/insert_splits.cpp compiled unmodified) ------------
// The whole training pipeline of the reference -- VideoSampledShotsDataLayer (fake LMDB, libc rand()), Net::Init with
// split insertion and loss weights, Net::ForwardBackward, SGDSolver::ComputeUpdateValue (learning-rate policy, weight
// decay, momentum history), Net::Update -- on a NetParameter built here with the structure of
// projects/videovec_embedding/mednet_embedding_train.prototxt (no text parser in the shim).  Not compiled from the
// reference and replaced by the few stand-ins below: layer_factory.cpp (it names every layer type of the fork),
// util/io.cpp / upgrade_proto.cpp (protobuf text + binary IO), pb2json.
namespace caffe {
template <typename Dtype>
Layer<Dtype>* GetLayer(const LayerParameter& param) {
  switch (param.type()) {
    case LayerParameter_LayerType_VIDEO_SAMPLED_SHOTS_DATA: return new VideoSampledShotsDataLayer<Dtype>(param);
    case LayerParameter_LayerType_SLICE: return new SliceLayer<Dtype>(param);
    case LayerParameter_LayerType_CONCAT: return new ConcatLayer<Dtype>(param);
    case LayerParameter_LayerType_FLATTEN: return new FlattenLayer<Dtype>(param);
    case LayerParameter_LayerType_INNER_PRODUCT: return new InnerProductLayer<Dtype>(param);
    case LayerParameter_LayerType_RELU: return new ReLULayer<Dtype>(param);
    case LayerParameter_LayerType_DROPOUT: return new DropoutLayer<Dtype>(param);
    case LayerParameter_LayerType_ELTWISE: return new EltwiseLayer<Dtype>(param);
    case LayerParameter_LayerType_NORMALIZATION: return new NormalizationLayer<Dtype>(param);
    case LayerParameter_LayerType_SUM: return new SumLayer<Dtype>(param);
    case LayerParameter_LayerType_SPLIT: return new SplitLayer<Dtype>(param);
    case LayerParameter_LayerType_MAX_MARGIN_LOSS: return new MaxMarginLossLayer<Dtype>(param);
    case LayerParameter_LayerType_VIDEO_SHOT_WINDOW_TEST_DATA: return new VideoShotWindowTestDataLayer<Dtype>(param);
    case LayerParameter_LayerType_RETRIEVAL_STATS: return new RetrievalStatsLayer<Dtype>(param);
    default: LOG(FATAL) << "layer type " << param.type() << " is not part of the shim factory"; return NULL;
  }
}
template Layer<float>* GetLayer(const LayerParameter& param);
template Layer<double>* GetLayer(const LayerParameter& param);
bool NetNeedsUpgrade(const NetParameter&) { return false; }
bool UpgradeV0Net(const NetParameter&, NetParameter*) { return false; }
bool NetNeedsDataUpgrade(const NetParameter&) { return false; }
void UpgradeNetDataTransformation(NetParameter*) {}
void ReadNetParamsFromTextFileOrDie(const string&, NetParameter*) { LOG(FATAL) << "no text parser in the shim"; }
void ReadNetParamsFromBinaryFileOrDie(const string&, NetParameter*) { LOG(FATAL) << "no binary parser in the shim"; }
bool ReadProtoFromTextFile(const char*, google::protobuf::Message*) { return false; }
void WriteProtoToTextFile(const google::protobuf::Message&, const char*) {}
void WriteProtoToBinaryFile(const google::protobuf::Message&, const char*) {}
}  // namespace caffe
char* pb2json(const google::protobuf::Message&) { return strdup("{}"); }

namespace {
struct StepSolver : public SGDSolver<float> {
  explicit StepSolver(const SolverParameter& p) : SGDSolver<float>(p) {}
  void pre() { Caffe::set_phase(Caffe::TRAIN); PreSolve(); iter_ = 0; }
  // the body of Solver::Solve's loop (solver.cpp:195-220) without display / snapshot / test
  float step() {
    vector<Blob<float>*> bottom_vec;
    const float loss = net_->ForwardBackward(bottom_vec);
    ComputeUpdateValue();
    net_->Update();
    ++iter_;
    return loss;
  }
  vector<shared_ptr<Blob<float> > >& hist() { return history_; }
};
struct RefSolver { FakeDb db, test_db; std::unique_ptr<StepSolver> solver; };
void phase_rule(LayerParameter* l, Phase p) { l->add_include()->set_phase(p); }
LayerParameter* add_layer(NetParameter* np, const char* name, LayerParameter_LayerType type,
                          const std::vector<std::string>& bottoms, const std::vector<std::string>& tops) {
  LayerParameter* l = np->add_layers();
  l->set_name(name); l->set_type(type);
  for (const auto& b : bottoms) l->add_bottom(b);
  for (const auto& t : tops) l->add_top(t);
  return l;
}
std::string nm(const char* fmt, int i) { char buf[96]; snprintf(buf, sizeof(buf), fmt, i); return buf; }
}  // namespace

extern "C" {
// lr_policy: 0 fixed, 1 inv, 2 step.  Parity runs pass dropout_ratio = 0 (no dropout layer: its mask comes from boost's
// generator in the reference; DropoutLayer itself is pinned at the layer level).  Call ref_srand(seed) first (the data layer's rand()).
REF_API void* ref_solver_create(int V, int K, const int* video_id, const int* shot_off, const int* shot_ids, const float* feat,
                                int B, int C, int Nn, int N, int max_buffer_size, int negative_swap_percentage,
                                int max_same_video_negs, int context_type, float margin, int norm,
                                float base_lr, float momentum, float weight_decay, int lr_policy, float gamma, float power,
                                int stepsize, const float* W0, const float* b0, float dropout_ratio,
                                int n_test, int frames, const float* test_data /*[n_test, frames, K]*/, const int* test_video_id,
                                int test_batch, const char* id_to_class_file, int exclude_same_video_shots, int reg_type) {
  try {
#ifdef VV_DROPIN_GPU      // the drop-in build (dropin_gpu.cpp): the same solver in Caffe GPU mode, device bodies = the C-ABI
    Caffe::set_mode(Caffe::GPU);
#else
    Caffe::set_mode(Caffe::CPU);
#endif
    Caffe::set_phase(Caffe::TRAIN);
    RefSolver* s = new RefSolver();
    for (int v = 0; v < V; ++v) {
      std::shared_ptr<VideoShots> rec(new VideoShots());
      rec->set_video_id(video_id[v]);
      for (int g = shot_off[v]; g < shot_off[v + 1]; ++g) {
        rec->add_shot_ids(shot_ids[g]);
        Datum* d = rec->add_shot_words();
        for (int k = 0; k < K; ++k) d->add_float_data(feat[size_t(g) * K + k]);
      }
      s->db.records.push_back(rec);
    }
    const bool with_test = n_test > 0 && test_data && id_to_class_file;
    for (int i = 0; with_test && i < n_test; ++i) {
      std::shared_ptr<TestVideoShotWindows> rec(new TestVideoShotWindows());
      rec->set_video_id(test_video_id[i]);
      for (int f = 0; f < frames; ++f) {
        Datum* d = rec->add_context_shot_words();
        for (int k = 0; k < K; ++k) d->add_float_data(test_data[(size_t(i) * frames + f) * K + k]);
      }
      s->test_db.records.push_back(rec);
    }
    SolverParameter sp;
    if (with_test) { sp.add_test_iter(1); sp.set_test_interval(1 << 30); sp.set_test_initialization(false); }
    sp.set_base_lr(base_lr); sp.set_momentum(momentum); sp.set_weight_decay(weight_decay);
    if (reg_type == 1) sp.set_regularization_type("L1");
    sp.set_lr_policy(lr_policy == 1 ? "inv" : lr_policy == 2 ? "step" : "fixed");
    sp.set_gamma(gamma); sp.set_power(power); sp.set_stepsize(stepsize);
    sp.set_max_iter(1 << 30); sp.set_display(0); sp.set_snapshot(0); sp.set_snapshot_after_train(false);
    sp.set_solver_mode(SolverParameter_SolverMode_CPU);
    NetParameter* np = sp.mutable_net_param();
    np->set_name("med_embedding");
    typedef std::vector<std::string> SV;
    const size_t first_layer = 0;
    LayerParameter* l = add_layer(np, "shot_windows", LayerParameter_LayerType_VIDEO_SAMPLED_SHOTS_DATA, {}, {"data"});
    VideoSampledShotsDataParameter* vp = l->mutable_video_sampled_shots_data_param();
    vp->set_source("mem://fake-lmdb"); vp->set_backend(VideoSampledShotsDataParameter_DB_LMDB);
    vp->set_batch_size(B); vp->set_context_size(C); vp->set_num_negative_samples(Nn);
    vp->set_max_buffer_size(max_buffer_size); vp->set_negative_swap_percentage(negative_swap_percentage);
    vp->set_max_same_video_negs(max_same_video_negs);
    vp->set_context_type(VideoSampledShotsDataParameter_CONTEXT(context_type));
    SV raw{"target"}, ctx_e, neg_e, neg_n;
    for (int i = 1; i < C; ++i) { raw.push_back(nm("context_window_%d", i)); ctx_e.push_back(nm("context_window_emb_%d_nonorm", i)); }
    for (int k = 1; k <= Nn; ++k) { raw.push_back(nm("negative_%d", k)); neg_e.push_back(nm("negative_emb_%d", k)); neg_n.push_back(nm("negative_emb_%d_nonorm", k)); }
    add_layer(np, "slice_input_data", LayerParameter_LayerType_SLICE, {"data"}, raw)->mutable_slice_param()->set_slice_dim(1);
    add_layer(np, "batch_concat_input", LayerParameter_LayerType_CONCAT, raw, {"batch_concat"})->mutable_concat_param()->set_concat_dim(0);
    add_layer(np, "flatten_input", LayerParameter_LayerType_FLATTEN, {"batch_concat"}, {"original_feature"});
    l = add_layer(np, "fc7", LayerParameter_LayerType_INNER_PRODUCT, {"original_feature"}, {"ip1_nonorm"});
    l->add_blobs_lr(1); l->add_blobs_lr(2); l->add_weight_decay(1); l->add_weight_decay(0);
    l->mutable_inner_product_param()->set_num_output(N);
    l->mutable_inner_product_param()->mutable_weight_filler()->set_type("constant");
    l->mutable_inner_product_param()->mutable_bias_filler()->set_type("constant");
    add_layer(np, "fc7_relu", LayerParameter_LayerType_RELU, {"ip1_nonorm"}, {"ip2"});
    if (dropout_ratio > 0.f)      // timing runs only: the mask comes from the shim's stand-in for boost's generator
      add_layer(np, "drop2", LayerParameter_LayerType_DROPOUT, {"ip2"}, {"ip2"})->mutable_dropout_param()->set_dropout_ratio(dropout_ratio);
    SV emb{"target_emb_nonorm"}; emb.insert(emb.end(), ctx_e.begin(), ctx_e.end()); emb.insert(emb.end(), neg_n.begin(), neg_n.end());
    add_layer(np, "slice_emb", LayerParameter_LayerType_SLICE, {"ip2"}, emb)->mutable_slice_param()->set_slice_dim(0);
    l = add_layer(np, "context_average", LayerParameter_LayerType_ELTWISE, ctx_e, {"context_feature_nonorm"});
    l->mutable_eltwise_param()->set_operation(EltwiseParameter_EltwiseOp_SUM);
    for (int i = 1; i < C; ++i) l->mutable_eltwise_param()->add_coeff(1.0f / float(C - 1));
    add_layer(np, "word_embedding_norm", LayerParameter_LayerType_NORMALIZATION, {"context_feature_nonorm"}, {"context_feature"});
    SV pn{"target_emb_nonorm"}; pn.insert(pn.end(), neg_n.begin(), neg_n.end());
    add_layer(np, "concat_pos_neg_nonorm", LayerParameter_LayerType_CONCAT, pn, {"pos_neg_nonorm"})->mutable_concat_param()->set_concat_dim(0);
    add_layer(np, "pos_neg_normalize", LayerParameter_LayerType_NORMALIZATION, {"pos_neg_nonorm"}, {"pos_neg_norm"});
    SV pe{"target_emb"}; pe.insert(pe.end(), neg_e.begin(), neg_e.end());
    add_layer(np, "slice_pos_neg_norm", LayerParameter_LayerType_SLICE, {"pos_neg_norm"}, pe)->mutable_slice_param()->set_slice_dim(0);
    add_layer(np, "prod_true", LayerParameter_LayerType_ELTWISE, {"context_feature", "target_emb"}, {"target_prod"})
        ->mutable_eltwise_param()->set_operation(EltwiseParameter_EltwiseOp_PROD);
    add_layer(np, "sum_true", LayerParameter_LayerType_SUM, {"target_prod"}, {"target_score"})->mutable_sum_param()->set_num_output(Nn);
    SV nsc;
    for (int k = 1; k <= Nn; ++k) {
      add_layer(np, nm("prod_neg_%d", k).c_str(), LayerParameter_LayerType_ELTWISE, {"context_feature", nm("negative_emb_%d", k)},
                {nm("negative_emb_%d_prod", k)})->mutable_eltwise_param()->set_operation(EltwiseParameter_EltwiseOp_PROD);
      add_layer(np, nm("sum_neg_%d", k).c_str(), LayerParameter_LayerType_SUM, {nm("negative_emb_%d_prod", k)}, {nm("neg_score_%d", k)});
      nsc.push_back(nm("neg_score_%d", k));
    }
    add_layer(np, "concat_negative_scores", LayerParameter_LayerType_CONCAT, nsc, {"negative_score"})->mutable_concat_param()->set_concat_dim(1);
    l = add_layer(np, "max_margin_loss", LayerParameter_LayerType_MAX_MARGIN_LOSS, {"target_score", "negative_score"},
                  {"loss_output", "train_violations"});
    l->add_loss_weight(1); l->add_loss_weight(0);
    l->mutable_max_margin_loss_param()->set_margin(margin);
    l->mutable_max_margin_loss_param()->set_norm(norm == 2 ? MaxMarginLossParameter_Norm_L2 : MaxMarginLossParameter_Norm_L1);
    if (with_test) {
      // every layer so far is TRAIN-only except fc7 / fc7_relu (no include rule in the shipped file), then the TEST graph
      for (int i = int(first_layer); i < np->layers_size(); ++i) {
        const std::string& n = np->layers(i).name();
        if (n != "fc7" && n != "fc7_relu") phase_rule(np->mutable_layers(i), TRAIN);
      }
      SV fr, sl;
      for (int f = 1; f <= frames; ++f) { fr.push_back(nm("context_datum_%d", f)); sl.push_back(nm("test_sample_frame_%d", f)); }
      LayerParameter* t = add_layer(np, "shot_windows", LayerParameter_LayerType_VIDEO_SHOT_WINDOW_TEST_DATA, {}, {"data", "video_ids"});
      t->mutable_video_shot_window_test_data_param()->set_source("mem://fake-lmdb-test");
      t->mutable_video_shot_window_test_data_param()->set_backend(VideoShotWindowTestDataParameter_DB_LMDB);
      t->mutable_video_shot_window_test_data_param()->set_batch_size(test_batch);
      phase_rule(t, TEST);
      t = add_layer(np, "slice_input_data", LayerParameter_LayerType_SLICE, {"data"}, fr); t->mutable_slice_param()->set_slice_dim(1); phase_rule(t, TEST);
      t = add_layer(np, "batch_concat_input_test", LayerParameter_LayerType_CONCAT, fr, {"concat_input_datums"}); t->mutable_concat_param()->set_concat_dim(0); phase_rule(t, TEST);
      t = add_layer(np, "flatten_input", LayerParameter_LayerType_FLATTEN, {"concat_input_datums"}, {"concat_input_datums_flat"}); phase_rule(t, TEST);
      t = add_layer(np, "slice_test", LayerParameter_LayerType_SLICE, {"concat_input_datums_flat"}, sl); t->mutable_slice_param()->set_slice_dim(0); phase_rule(t, TEST);
      t = add_layer(np, "average_for_test", LayerParameter_LayerType_ELTWISE, sl, {"original_feature"});
      t->mutable_eltwise_param()->set_operation(EltwiseParameter_EltwiseOp_SUM);
      for (int f = 0; f < frames; ++f) t->mutable_eltwise_param()->add_coeff(1.0f / float(frames));
      phase_rule(t, TEST);
      t = add_layer(np, "test_norm", LayerParameter_LayerType_NORMALIZATION, {"ip2"}, {"ip2_norm"}); phase_rule(t, TEST);
      t = add_layer(np, "retrieval_stats", LayerParameter_LayerType_RETRIEVAL_STATS, {"ip2_norm", "video_ids"},
                    {"test_map", "test_hit_at_1", "test_hit_at_5"});
      t->mutable_retrieval_stats_param()->set_id_to_class_file(id_to_class_file);
      t->mutable_retrieval_stats_param()->set_exclude_same_video_shots(exclude_same_video_shots != 0);
      phase_rule(t, TEST);
      // Net::Init appends layers in file order and a blob must be produced before it is consumed: the TEST data path has to
      // precede fc7, as in the shipped file (its TEST layers sit between the TRAIN input layers and fc7)
      NetParameter ordered; ordered.set_name(np->name());
      auto is_test_input = [&](const std::string& n, const LayerParameter& lp) {
        return lp.include_size() == 1 && lp.include(0).phase() == TEST && n != "test_norm" && n != "retrieval_stats";
      };
      for (int i = 0; i < np->layers_size(); ++i) { const LayerParameter& lp = np->layers(i); if (lp.name() == "fc7") break; ordered.add_layers()->CopyFrom(lp); }
      for (int i = 0; i < np->layers_size(); ++i) { const LayerParameter& lp = np->layers(i); if (is_test_input(lp.name(), lp)) ordered.add_layers()->CopyFrom(lp); }
      bool after = false;
      for (int i = 0; i < np->layers_size(); ++i) {
        const LayerParameter& lp = np->layers(i);
        if (lp.name() == "fc7") after = true;
        if (after && !is_test_input(lp.name(), lp)) ordered.add_layers()->CopyFrom(lp);
      }
      np->CopyFrom(ordered);
    }
    g_fake_db = &s->db; g_fake_db_test = with_test ? &s->test_db : nullptr;
    s->solver.reset(new StepSolver(sp));           // Solver::Init -> Net::Init -> every layer's SetUp (data layer: buffer init)
    g_fake_db = nullptr; g_fake_db_test = nullptr;
    s->solver->pre();
    const vector<shared_ptr<Blob<float> > >& params = s->solver->net()->params();
    if (params.size() != 2) { fprintf(stderr, "ref_driver: expected 2 parameter blobs, got %d\n", int(params.size())); delete s; return nullptr; }
    memcpy(params[0]->mutable_cpu_data(), W0, sizeof(float) * size_t(N) * K);
    memcpy(params[1]->mutable_cpu_data(), b0, sizeof(float) * N);
    return s;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); return nullptr; }
}
REF_API int ref_solver_step(void* h, float* loss, float* violations) {
  try {
    RefSolver* s = static_cast<RefSolver*>(h);
    *loss = s->solver->step();
    if (violations) *violations = s->solver->net()->blob_by_name("train_violations")->cpu_data()[0];
    return 0;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); return -1; }
}
// W [N,K], b [N], their momentum histories, and the data blob [B,R,K] of the last step (any may be NULL)
REF_API int ref_solver_get(void* h, float* W, float* b, float* hW, float* hb, float* data) {
  try {
    RefSolver* s = static_cast<RefSolver*>(h);
    const vector<shared_ptr<Blob<float> > >& params = s->solver->net()->params();
    if (W) memcpy(W, params[0]->cpu_data(), sizeof(float) * params[0]->count());
    if (b) memcpy(b, params[1]->cpu_data(), sizeof(float) * params[1]->count());
    if (hW) memcpy(hW, s->solver->hist()[0]->cpu_data(), sizeof(float) * params[0]->count());
    if (hb) memcpy(hb, s->solver->hist()[1]->cpu_data(), sizeof(float) * params[1]->count());
    if (data) { const shared_ptr<Blob<float> > d = s->solver->net()->blob_by_name("data"); memcpy(data, d->cpu_data(), sizeof(float) * d->count()); }
    return 0;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); return -1; }
}
// any blob of the TRAIN net by name (data, or diff with diff != 0), as the host sees it; returns its count (or < 0)
REF_API int ref_solver_blob(void* h, const char* name, int diff, float* out, int cap) {
  try {
    RefSolver* s = static_cast<RefSolver*>(h);
    if (!s->solver->net()->has_blob(name)) return -2;
    const shared_ptr<Blob<float> > b = s->solver->net()->blob_by_name(name);
    const int n = b->count();
    if (out && cap >= n) memcpy(out, diff ? b->cpu_diff() : b->cpu_data(), sizeof(float) * n);
    return n;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); return -1; }
}
REF_API int ref_solver_num_blobs(void* h) { return int(static_cast<RefSolver*>(h)->solver->net()->blob_names().size()); }
REF_API const char* ref_solver_blob_name(void* h, int i) { return static_cast<RefSolver*>(h)->solver->net()->blob_names()[i].c_str(); }
// Solver::Test's loop (solver.cpp:252-317) on test net 0: weights shared with the train net, `iters` forward passes,
// mean of the three outputs (test_map, test_hit_at_1, test_hit_at_5)
REF_API int ref_solver_test(void* h, int iters, float* out3) {
  try {
    RefSolver* s = static_cast<RefSolver*>(h);
    if (s->solver->test_nets().empty()) return -2;
    Net<float>* tn = s->solver->test_nets()[0].get();
    Caffe::set_phase(Caffe::TEST);
    tn->ShareTrainedLayersWith(s->solver->net().get());
    double acc[3] = {0, 0, 0};
    vector<Blob<float>*> bottom_vec;
    for (int i = 0; i < iters; ++i) {
      float loss = 0;
      const vector<Blob<float>*>& res = tn->Forward(bottom_vec, &loss);
      if (res.size() != 3) { Caffe::set_phase(Caffe::TRAIN); return -3; }
      // output blobs come in the order of Net::Init's std::set of blob names (lexicographic): place them by name
      for (int j = 0; j < 3; ++j) {
        const std::string& nme = tn->blob_names()[tn->output_blob_indices()[j]];
        const int slot = nme == "test_map" ? 0 : nme == "test_hit_at_1" ? 1 : nme == "test_hit_at_5" ? 2 : -1;
        if (slot < 0) { Caffe::set_phase(Caffe::TRAIN); return -4; }
        acc[slot] += res[j]->cpu_data()[0];
      }
    }
    Caffe::set_phase(Caffe::TRAIN);
    for (int j = 0; j < 3; ++j) out3[j] = float(acc[j] / iters);
    return 0;
  } catch (const std::exception& e) { fprintf(stderr, "ref_driver: %s\n", e.what()); Caffe::set_phase(Caffe::TRAIN); return -1; }
}
REF_API int ref_solver_test_num_layers(void* h) { RefSolver* s = static_cast<RefSolver*>(h); return s->solver->test_nets().empty() ? 0 : int(s->solver->test_nets()[0]->layers().size()); }
REF_API const char* ref_solver_test_layer_name(void* h, int i) { return static_cast<RefSolver*>(h)->solver->test_nets()[0]->layer_names()[i].c_str(); }
REF_API const char* ref_solver_test_output_name(void* h, int j) {
  Net<float>* tn = static_cast<RefSolver*>(h)->solver->test_nets()[0].get();
  return tn->blob_names()[tn->output_blob_indices()[j]].c_str();
}
REF_API int ref_solver_num_layers(void* h) { return int(static_cast<RefSolver*>(h)->solver->net()->layers().size()); }
REF_API const char* ref_solver_layer_name(void* h, int i) { return static_cast<RefSolver*>(h)->solver->net()->layer_names()[i].c_str(); }
REF_API void ref_solver_destroy(void* h) { delete static_cast<RefSolver*>(h); }
}  // extern "C"
