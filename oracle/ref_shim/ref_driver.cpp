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
FakeDb* g_fake_db = nullptr;          // the dataset the next mdb_env_open serves
}  // namespace
struct MDB_env { FakeDb* db; };
struct MDB_txn { MDB_env* env; };
struct MDB_cursor { FakeDb* db; size_t pos; };
extern "C" {
int mdb_env_create(MDB_env** env) { *env = new MDB_env{nullptr}; return MDB_SUCCESS; }
int mdb_env_set_mapsize(MDB_env*, size_t) { return MDB_SUCCESS; }
int mdb_env_open(MDB_env* env, const char*, unsigned int, mdb_mode_t) { env->db = g_fake_db; return env->db ? MDB_SUCCESS : -1; }
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
  FakeDb db;
  std::unique_ptr<VideoSampledShotsDataLayer<float> > layer;
  Blob<float> top;
  int B, R, K;
};
}  // namespace

extern "C" {
// Dataset as in the oracle's sampler: videos [V] with shots [shot_off[v], shot_off[v+1]) of `feat` [total, K].
// context_type: 0 PAIRWISE, 1 WINDOW, 2 PAST, 3 PAST_CONTINUOUS, 4 PAST_CONTINUOUS_FIXED.  Call srand(seed) first: the
// layer draws from the process-global rand() (from its prefetch thread; one batch is always prefetched ahead).
REF_API void* ref_sampler_create(int V, int K, const int* video_id, const int* shot_off, const int* shot_ids, const float* feat,
                                 int batch_size, int context_size, int num_negative_samples, int max_buffer_size,
                                 int negative_swap_percentage, int max_same_video_negs, int context_type) {
  try {
    Caffe::set_mode(Caffe::CPU);
    RefSampler* s = new RefSampler();
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
    LayerParameter p;
    VideoSampledShotsDataParameter* vp = p.mutable_video_sampled_shots_data_param();
    vp->set_source("mem://fake-lmdb"); vp->set_backend(VideoSampledShotsDataParameter_DB_LMDB);
    vp->set_batch_size(batch_size); vp->set_context_size(context_size); vp->set_num_negative_samples(num_negative_samples);
    vp->set_max_buffer_size(max_buffer_size); vp->set_negative_swap_percentage(negative_swap_percentage);
    vp->set_max_same_video_negs(max_same_video_negs);
    vp->set_context_type(VideoSampledShotsDataParameter_CONTEXT(context_type));
    g_fake_db = &s->db;
    s->layer.reset(new VideoSampledShotsDataLayer<float>(p));
    BV bottom, tv{&s->top};
    s->layer->SetUp(bottom, &tv);            // DataLayerSetUp (negative buffer init) + first prefetch
    g_fake_db = nullptr;
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
