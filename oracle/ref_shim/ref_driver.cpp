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
