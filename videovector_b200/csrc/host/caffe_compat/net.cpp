// Net<Dtype> (see caffe/net.hpp for the reference lines mirrored here).
#include <cuda_runtime_api.h>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include "caffe/net.hpp"

namespace caffe {

// ---- FilterNet: keep a layer if it has no include rule or one rule matches the phase (ref: net.cpp:227-268)
NetParameter FilterNet(const NetParameter& param, Caffe::Phase phase) {
  auto out = std::make_shared<PbMsg>();
  for (const PbField& f : param.m->fields) {
    if (f.key != "layers") { out->fields.push_back(f); continue; }
    LayerParameter lp(f.msg);
    bool keep = lp.include_size() == 0;
    for (int i = 0; i < lp.include_size() && !keep; ++i) {
      NetStateRule r = lp.include(i);
      keep = !r.has_phase() || r.phase() == phase;
    }
    if (keep) out->fields.push_back(f);
  }
  return NetParameter(out);
}

// ---- InsertSplits (ref: util/insert_splits.cpp:12-142): every top consumed by more than one bottom (or consumed
// and carrying a loss weight) gets a SPLIT layer right after its producer; consumers are rewired in layer order.
static string SplitLayerName(const string& layer, const string& blob, int blob_idx) {
  return blob + "_" + layer + "_" + std::to_string(blob_idx) + "_split";
}
static string SplitBlobName(const string& layer, const string& blob, int blob_idx, int split_idx) {
  return blob + "_" + layer + "_" + std::to_string(blob_idx) + "_split_" + std::to_string(split_idx);
}
NetParameter InsertSplits(const NetParameter& param) {
  typedef std::pair<int, int> Idx;                       // (layer, top index)
  std::map<string, Idx> producer;                        // blob name -> latest producer
  std::map<Idx, Idx> bottom_to_top;                      // (layer, bottom) -> producing (layer, top)
  std::map<Idx, int> use_count, seen;
  std::map<Idx, float> loss_weight;
  const int n = param.layers_size();
  vector<LayerParameter> L;
  for (int i = 0; i < n; ++i) L.push_back(param.layers(i));
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < L[i].bottom_size(); ++j) {
      auto it = producer.find(L[i].bottom(j));
      CHECK(it != producer.end()) << "Unknown blob input " << L[i].bottom(j) << " to layer " << L[i].name();
      bottom_to_top[Idx(i, j)] = it->second;
      ++use_count[it->second];
    }
    for (int j = 0; j < L[i].top_size(); ++j) producer[L[i].top(j)] = Idx(i, j);
    const int nw = std::min(L[i].loss_weight_size(), L[i].top_size());
    for (int j = 0; j < nw; ++j) {
      const float w = L[i].loss_weight(j);
      loss_weight[Idx(i, j)] = w;
      if (w) ++use_count[Idx(i, j)];
    }
  }
  auto out = std::make_shared<PbMsg>();
  for (const PbField& f : param.m->fields) if (f.key != "layers") out->fields.push_back(f);
  for (int i = 0; i < n; ++i) {
    // deep-copy the layer message so renaming bottoms does not touch the caller's parameter
    auto copy = std::make_shared<PbMsg>(*L[i].m);
    int bidx = 0;
    for (PbField& f : copy->fields) {
      if (f.key != "bottom") continue;
      const Idx top = bottom_to_top[Idx(i, bidx)];
      if (use_count[top] > 1) f.scalar = SplitBlobName(L[top.first].name(), L[top.first].top(top.second), top.second, seen[top]++);
      ++bidx;
    }
    out->fields.push_back(PbField{"layers", "", copy});
    for (int j = 0; j < L[i].top_size(); ++j) {
      const Idx top(i, j);
      const int cnt = use_count[top];
      if (cnt <= 1) continue;
      LayerParameter split;
      split.set_name(SplitLayerName(L[i].name(), L[i].top(j), j));
      split.set_type(LayerParameter_LayerType_SPLIT);
      split.add_bottom(L[i].top(j));
      const float w = loss_weight.count(top) ? loss_weight[top] : 0.f;
      for (int k = 0; k < cnt; ++k) {
        split.add_top(SplitBlobName(L[i].name(), L[i].top(j), j, k));
        if (w) { if (k == 0) seen[top]++; split.add_loss_weight(k == 0 ? w : 0.f); }
      }
      out->fields.push_back(PbField{"layers", "", split.m});
    }
  }
  return NetParameter(out);
}

template <typename Dtype>
Net<Dtype>::~Net() { if (trainer_) vv_trainer_destroy(trainer_); }

template <typename Dtype>
int Net<Dtype>::AppendBottom(const NetParameter& param, int layer_id, int bottom_id, std::set<string>* available,
                             std::map<string, int>* name_to_idx) {
  const LayerParameter lp = param.layers(layer_id);
  const string blob_name = lp.bottom(bottom_id);
  CHECK(available->count(blob_name)) << "Unknown blob input " << blob_name << " (at index " << bottom_id << ") to layer " << layer_id;
  const int blob_id = (*name_to_idx)[blob_name];
  bottom_vecs_[layer_id].push_back(blobs_[blob_id].get());
  bottom_id_vecs_[layer_id].push_back(blob_id);
  available->erase(blob_name);
  bottom_need_backward_[layer_id].push_back(blob_need_backward_[blob_id]);
  return blob_id;
}
template <typename Dtype>
void Net<Dtype>::AppendTop(const NetParameter& param, int layer_id, int top_id, std::set<string>* available,
                           std::map<string, int>* name_to_idx) {
  const LayerParameter lp = param.layers(layer_id);
  const string blob_name = lp.top(top_id);
  if (lp.bottom_size() > top_id && blob_name == lp.bottom(top_id)) {
    // in-place computation (e.g. drop2: bottom ip2, top ip2)
    top_vecs_[layer_id].push_back(blobs_[(*name_to_idx)[blob_name]].get());
    top_id_vecs_[layer_id].push_back((*name_to_idx)[blob_name]);
  } else {
    CHECK(!name_to_idx->count(blob_name)) << "Duplicate blobs produced by multiple sources: " << blob_name;
    blobs_.push_back(shared_ptr<Blob<Dtype> >(new Blob<Dtype>()));
    const int blob_id = blobs_.size() - 1;
    blob_names_.push_back(blob_name);
    blob_need_backward_.push_back(false);
    (*name_to_idx)[blob_name] = blob_id;
    top_id_vecs_[layer_id].push_back(blob_id);
    top_vecs_[layer_id].push_back(blobs_[blob_id].get());
  }
  available->insert(blob_name);
}

template <typename Dtype>
void Net<Dtype>::Init(const NetParameter& in_param, Caffe::Phase phase) {
  Caffe::set_phase(phase);
  const NetParameter param = InsertSplits(FilterNet(in_param, phase));
  name_ = param.name();
  std::map<string, int> blob_name_to_idx;
  std::set<string> available_blobs;
  const int n = param.layers_size();
  bottom_vecs_.resize(n); top_vecs_.resize(n); bottom_id_vecs_.resize(n); top_id_vecs_.resize(n); bottom_need_backward_.resize(n);
  for (int layer_id = 0; layer_id < n; ++layer_id) {
    const LayerParameter lp = param.layers(layer_id);
    layers_.push_back(shared_ptr<Layer<Dtype> >(GetLayer<Dtype>(lp)));
    layer_names_.push_back(lp.name());
    LogInfo("Creating Layer " + lp.name());
    bool need_backward = false;
    for (int b = 0; b < lp.bottom_size(); ++b)
      need_backward |= blob_need_backward_[AppendBottom(param, layer_id, b, &available_blobs, &blob_name_to_idx)];
    for (int t = 0; t < lp.top_size(); ++t) AppendTop(param, layer_id, t, &available_blobs, &blob_name_to_idx);
    layers_[layer_id]->SetUp(bottom_vecs_[layer_id], &top_vecs_[layer_id]);
    for (size_t t = 0; t < top_vecs_[layer_id].size(); ++t) {
      const int id = top_id_vecs_[layer_id][t];
      if (int(blob_loss_weights_.size()) <= id) blob_loss_weights_.resize(id + 1, Dtype(0));
      blob_loss_weights_[id] = layers_[layer_id]->loss(t);
    }
    const int blobs_lr_size = lp.blobs_lr_size();
    const int num_param_blobs = layers_[layer_id]->blobs().size();
    CHECK(blobs_lr_size == num_param_blobs || blobs_lr_size == 0)
        << "Incorrect blobs lr size: should be either 0 or the same as the number of the layer's parameter blobs.";
    if (blobs_lr_size) {
      for (int p = 0; p < blobs_lr_size; ++p) {
        const bool pnb = lp.blobs_lr(p) > 0;
        need_backward |= pnb;
        layers_[layer_id]->set_param_propagate_down(p, pnb);
      }
    } else if (num_param_blobs) {
      need_backward = true;
    }
    CHECK(lp.weight_decay_size() == num_param_blobs || lp.weight_decay_size() == 0)
        << "Incorrect weight decay size: should be either 0 or the same as the number of the layer's parameter blobs";
    for (int p = 0; p < num_param_blobs; ++p) {       // AppendParam + GetLearningRateAndWeightDecay (net.cpp:467-499)
      params_.push_back(layers_[layer_id]->blobs()[p]);
      params_lr_.push_back(blobs_lr_size ? lp.blobs_lr(p) : 1.f);
      params_weight_decay_.push_back(lp.weight_decay_size() ? lp.weight_decay(p) : 1.f);
    }
    layer_need_backward_.push_back(need_backward);
    if (need_backward) for (int id : top_id_vecs_[layer_id]) blob_need_backward_[id] = true;
  }
  // which blobs contribute to the loss (net.cpp:146-186)
  std::set<string> under_loss;
  for (int layer_id = n - 1; layer_id >= 0; --layer_id) {
    bool contributes = false;
    for (size_t t = 0; t < top_vecs_[layer_id].size() && !contributes; ++t)
      contributes = layers_[layer_id]->loss(t) || under_loss.count(blob_names_[top_id_vecs_[layer_id][t]]);
    if (!contributes) layer_need_backward_[layer_id] = false;
    for (size_t b = 0; b < bottom_vecs_[layer_id].size(); ++b) {
      if (contributes) under_loss.insert(blob_names_[bottom_id_vecs_[layer_id][b]]);
      else bottom_need_backward_[layer_id][b] = false;
    }
  }
  for (const string& nm : available_blobs) {
    net_output_blobs_.push_back(blobs_[blob_name_to_idx[nm]].get());
    net_output_blob_indices_.push_back(blob_name_to_idx[nm]);
  }
  for (size_t i = 0; i < blob_names_.size(); ++i) blob_names_index_[blob_names_[i]] = i;
  for (size_t i = 0; i < layer_names_.size(); ++i) layer_names_index_[layer_names_[i]] = i;
  LogInfo("Network initialization done.");
}

template <typename Dtype>
Dtype Net<Dtype>::ForwardFromTo(int start, int end) {
  CHECK_GE(start, 0); CHECK_LT(end, int(layers_.size()));
  Dtype loss = 0;
  for (int i = start; i <= end; ++i) {
    layers_[i]->Reshape(bottom_vecs_[i], &top_vecs_[i]);       // Reshape before every Forward (net.cpp:508)
    loss += layers_[i]->Forward(bottom_vecs_[i], &top_vecs_[i]);
  }
  return loss;
}
template <typename Dtype>
const vector<Blob<Dtype>*>& Net<Dtype>::ForwardPrefilled(Dtype* loss) {
  const Dtype l = ForwardFromTo(0, layers_.size() - 1);
  if (loss) *loss = l;
  return net_output_blobs_;
}
template <typename Dtype>
void Net<Dtype>::BackwardFromTo(int start, int end) {
  CHECK_GE(end, 0); CHECK_LT(start, int(layers_.size()));
  for (int i = start; i >= end; --i)
    if (layer_need_backward_[i]) layers_[i]->Backward(top_vecs_[i], bottom_need_backward_[i], &bottom_vecs_[i]);
}
template <typename Dtype> void Net<Dtype>::Backward() { BackwardFromTo(layers_.size() - 1, 0); }
template <typename Dtype> void Net<Dtype>::Update() { for (auto& p : params_) p->Update(); }   // no shared params on this path
template <typename Dtype>
const shared_ptr<Blob<Dtype> > Net<Dtype>::blob_by_name(const string& blob_name) {
  CHECK(has_blob(blob_name)) << "Unknown blob name " << blob_name;
  return blobs_[blob_names_index_[blob_name]];
}
template <typename Dtype>
const shared_ptr<Layer<Dtype> > Net<Dtype>::layer_by_name(const string& layer_name) {
  CHECK(has_layer(layer_name)) << "Unknown layer name " << layer_name;
  return layers_[layer_names_index_[layer_name]];
}

// ---------------------------------------------------------------------------------------------------------
// Fusion pass: recognise the shipped TRAIN graph by layer types, key parameters and wiring.
// ---------------------------------------------------------------------------------------------------------
template <typename Dtype>
bool Net<Dtype>::EnableFusion(string* why) {
  string dummy; string& w = why ? *why : dummy;
#define FUSE_REQUIRE(cond, msg) if (!(cond)) { w = msg; return false; }
  typedef LayerParameter_LayerType T;
  auto type = [&](size_t i) { return i < layers_.size() ? layers_[i]->type() : LayerParameter_LayerType_NONE; };
  size_t i = 0;
  FUSE_REQUIRE(type(0) == LayerParameter_LayerType_VIDEO_SAMPLED_SHOTS_DATA, "layer 0 is not VIDEO_SAMPLED_SHOTS_DATA");
  auto* data = dynamic_cast<VideoSampledShotsDataLayer<Dtype>*>(layers_[0].get());
  const int B = data->batch_size(), C = data->context_size(), Nn = data->num_negative_samples(), K = data->feature_size(), R = C + Nn;
  FUSE_REQUIRE(top_vecs_[0].size() == 1, "data layer with a label top is not fused");
  FUSE_REQUIRE(Nn >= 1 && (C % 2) == 1 && C >= 3, "needs an odd context_size >= 3 and negatives");
  // expected type sequence
  vector<T> expect = {LayerParameter_LayerType_VIDEO_SAMPLED_SHOTS_DATA, LayerParameter_LayerType_SLICE, LayerParameter_LayerType_CONCAT,
                      LayerParameter_LayerType_FLATTEN, LayerParameter_LayerType_INNER_PRODUCT, LayerParameter_LayerType_RELU};
  bool has_dropout = type(6) == LayerParameter_LayerType_DROPOUT;
  if (has_dropout) expect.push_back(LayerParameter_LayerType_DROPOUT);
  for (T t : {LayerParameter_LayerType_SLICE, LayerParameter_LayerType_ELTWISE, LayerParameter_LayerType_NORMALIZATION,
              LayerParameter_LayerType_SPLIT, LayerParameter_LayerType_CONCAT, LayerParameter_LayerType_NORMALIZATION,
              LayerParameter_LayerType_SLICE}) expect.push_back(t);
  for (int k = 0; k <= Nn; ++k) { expect.push_back(LayerParameter_LayerType_ELTWISE); expect.push_back(LayerParameter_LayerType_SUM); }
  expect.push_back(LayerParameter_LayerType_CONCAT); expect.push_back(LayerParameter_LayerType_MAX_MARGIN_LOSS);
  FUSE_REQUIRE(layers_.size() == expect.size(), "layer count differs from the temporal-embedding TRAIN graph");
  for (i = 0; i < expect.size(); ++i) FUSE_REQUIRE(type(i) == expect[i], "layer " + layer_names_[i] + " breaks the expected sequence");
  // parameters and wiring that the kernels assume
  size_t li = 1;
  const LayerParameter slice_in = layers_[li]->layer_param();
  FUSE_REQUIRE(slice_in.slice_param().slice_dim() == 1 && int(top_vecs_[li].size()) == R && slice_in.slice_param().slice_point_size() == 0, "input slice must cut dim 1 into R equal parts");
  ++li; FUSE_REQUIRE(layers_[li]->layer_param().concat_param().concat_dim() == 0 && int(bottom_vecs_[li].size()) == R, "batch concat must be along dim 0 over R blobs");
  for (int j = 0; j < R; ++j) FUSE_REQUIRE(bottom_vecs_[li][j] == top_vecs_[li - 1][j], "batch concat order differs from slice order");
  ++li; ++li;   // flatten, fc7
  auto* ip = dynamic_cast<InnerProductLayer<Dtype>*>(layers_[li].get());
  const LayerParameter ipp = layers_[li]->layer_param();
  FUSE_REQUIRE(ipp.inner_product_param().bias_term(), "fc7 without bias is not fused");
  const int N = ipp.inner_product_param().num_output();
  FUSE_REQUIRE(N % 4 == 0 && K % 4 == 0 && N <= 4096, "embedding / feature dims must be multiples of 4 (N <= 4096)");
  // the tensor-core GEMMs (and the row gather) need N % 8 == 0 and K % 8 == 0 (gemm_tc_supported); other shapes stay on
  // the layer-by-layer path, whose InnerProductLayer falls back to the exact fp32 kernel
  FUSE_REQUIRE(Caffe::precision() == VV_PREC_FP32_SIMT || (N % 8 == 0 && K % 8 == 0),
               "the tensor-core precisions need embedding / feature dims that are multiples of 8");
  FUSE_REQUIRE(!bottom_need_backward_[li][0], "fc7 bottom needs a gradient (dgrad) -- not the shipped net");
  ++li; FUSE_REQUIRE(layers_[li]->layer_param().relu_param().negative_slope() == 0.f, "leaky ReLU is not fused");
  float ratio = 0.f;
  if (has_dropout) { ++li; ratio = layers_[li]->layer_param().dropout_param().dropout_ratio(); FUSE_REQUIRE(top_vecs_[li][0] == bottom_vecs_[li][0], "dropout must be in place"); }
  ++li; const size_t slice_emb = li;
  FUSE_REQUIRE(layers_[li]->layer_param().slice_param().slice_dim() == 0 && int(top_vecs_[li].size()) == R, "slice_emb must cut dim 0 into R parts");
  ++li; const LayerParameter ctx = layers_[li]->layer_param();
  FUSE_REQUIRE(ctx.eltwise_param().operation() == EltwiseParameter_EltwiseOp_SUM && int(bottom_vecs_[li].size()) == C - 1, "context_average must SUM the C-1 context rows");
  for (int c = 0; c < C - 1; ++c) FUSE_REQUIRE(bottom_vecs_[li][c] == top_vecs_[slice_emb][1 + c], "context rows must be slots 1..C-1 in order");
  memset(&fused_cfg_, 0, sizeof(fused_cfg_));
  for (int c = 0; c < C - 1; ++c) fused_cfg_.coeff[c] = ctx.eltwise_param().coeff_size() ? ctx.eltwise_param().coeff(c) : 1.f;
  li += 2; FUSE_REQUIRE(int(top_vecs_[li].size()) == 1 + Nn, "context_feature must fan out to 1+Nn consumers");
  ++li; FUSE_REQUIRE(layers_[li]->layer_param().concat_param().concat_dim() == 0 && int(bottom_vecs_[li].size()) == 1 + Nn, "pos/neg concat must be dim 0 over 1+Nn blobs");
  FUSE_REQUIRE(bottom_vecs_[li][0] == top_vecs_[slice_emb][0], "pos/neg concat must start with the target row");
  for (int k = 0; k < Nn; ++k) FUSE_REQUIRE(bottom_vecs_[li][1 + k] == top_vecs_[slice_emb][C + k], "negatives must be slots C.. in order");
  li += 3;   // normalization, slice_pos_neg_norm -> first prod
  for (int k = 0; k <= Nn; ++k, li += 2) {
    FUSE_REQUIRE(layers_[li]->layer_param().eltwise_param().operation() == EltwiseParameter_EltwiseOp_PROD, "score branch must be ELTWISE PROD + SUM");
    const int nout = int(layers_[li + 1]->layer_param().sum_param().num_output());
    FUSE_REQUIRE(nout == (k == 0 ? Nn : 1), "sum_true must replicate to Nn columns, sum_neg_k to 1");
  }
  FUSE_REQUIRE(layers_[li]->layer_param().concat_param().concat_dim() == 1, "negative scores must be concatenated along dim 1");
  ++li; const LayerParameter lossp = layers_[li]->layer_param();
  FUSE_REQUIRE(bottom_vecs_[li].size() == 2, "weighted max-margin loss is not fused");
  fused_cfg_.B = B; fused_cfg_.C = C; fused_cfg_.Nn = Nn; fused_cfg_.K = K; fused_cfg_.N = N;
  fused_cfg_.margin = lossp.max_margin_loss_param().margin();
  fused_cfg_.norm = lossp.max_margin_loss_param().norm() == MaxMarginLossParameter_Norm_L2 ? 2 : 1;
  fused_cfg_.dropout_ratio = ratio; fused_cfg_.dropout_mode = VV_DROPOUT_HASH; fused_cfg_.dropout_seed = Caffe::rng_seed();
  fused_cfg_.loss_weight = layers_[li]->loss(0);
  fused_cfg_.regularization = ipp.inner_product_param().regularization();
  strncpy(fused_cfg_.lr_policy, "fixed", sizeof(fused_cfg_.lr_policy));
  fused_cfg_.reg_type = 2; fused_cfg_.lr_mult[0] = params_lr_[0]; fused_cfg_.lr_mult[1] = params_lr_[1];
  fused_cfg_.decay_mult[0] = params_weight_decay_[0]; fused_cfg_.decay_mult[1] = params_weight_decay_[1];
  fused_cfg_.prec = Caffe::precision(); fused_cfg_.world_size = 1; fused_cfg_.rank = 0;
  fused_data_ = data; fused_ip_ = ip;
  fused_loss_ = top_vecs_[li][0]; fused_viol_ = top_vecs_[li].size() > 1 ? top_vecs_[li][1] : nullptr;
#undef FUSE_REQUIRE
  return true;
}

// ref: net.cpp:774-801 (Net::ToProto) + layer.hpp:462-469 (Layer::ToProto: the layer's own parameter message, then its blobs)
template <typename Dtype>
shared_ptr<PbMsg> Net<Dtype>::ToProto(bool write_diff) const {
  auto out = std::make_shared<PbMsg>();
  out->add_scalar("name", name_);
  for (size_t i = 0; i < layers_.size(); ++i) {
    auto lp = std::make_shared<PbMsg>(*layers_[i]->layer_param().m);       // copy of the parameter message
    lp->fields.erase(std::remove_if(lp->fields.begin(), lp->fields.end(), [](const PbField& f) { return f.key == "blobs"; }), lp->fields.end());
    for (auto& b : layers_[i]->blobs()) {
      auto bp = std::make_shared<PbMsg>();
      b->ToProto(bp.get(), write_diff);
      lp->fields.push_back(PbField{"blobs", "", bp, nullptr});
    }
    out->fields.push_back(PbField{"layers", "", lp, nullptr});
  }
  return out;
}
// ref: net.cpp:692-727
template <typename Dtype>
void Net<Dtype>::CopyTrainedLayersFrom(const PbMsg& param) {
  const int num_source_layers = param.count("layers");
  for (int i = 0; i < num_source_layers; ++i) {
    const shared_ptr<PbMsg> src = param.sub("layers", i);
    const string source_layer_name = src->str("name");
    size_t target = 0;
    while (target != layer_names_.size() && layer_names_[target] != source_layer_name) ++target;
    if (target == layer_names_.size()) { LogInfo("Ignoring source layer " + source_layer_name); continue; }
    LogInfo("Copying source layer " + source_layer_name);
    vector<shared_ptr<Blob<Dtype> > >& target_blobs = layers_[target]->blobs();
    CHECK_EQ(int(target_blobs.size()), src->count("blobs")) << "Incompatible number of blobs for layer " << source_layer_name;
    for (size_t j = 0; j < target_blobs.size(); ++j) {
      const shared_ptr<PbMsg> sb = src->sub("blobs", int(j));
      CHECK_EQ(target_blobs[j]->num(), int(sb->num("num", 0)));
      CHECK_EQ(target_blobs[j]->channels(), int(sb->num("channels", 0)));
      CHECK_EQ(target_blobs[j]->height(), int(sb->num("height", 0)));
      CHECK_EQ(target_blobs[j]->width(), int(sb->num("width", 0)));
      const PbField* d = sb->nth("data", 0);
      CHECK(d && d->floats && int(d->floats->size()) == target_blobs[j]->count()) << "blob data size mismatch in layer " << source_layer_name;
      // written through the existing storage (not FromProto's Reshape) so that a parameter blob aliasing the fused
      // trainer's buffer keeps aliasing it
      memcpy(target_blobs[j]->mutable_cpu_data(), d->floats->data(), sizeof(Dtype) * target_blobs[j]->count());
      const PbField* g = sb->nth("diff", 0);
      if (g && g->floats && int(g->floats->size()) == target_blobs[j]->count())
        memcpy(target_blobs[j]->mutable_cpu_diff(), g->floats->data(), sizeof(Dtype) * target_blobs[j]->count());
      target_blobs[j]->gpu_data();                        // push to the device copy (the trainer's buffer when fused)
    }
  }
  if (trainer_) VV_CHECK(vv_trainer_sync_weights(trainer_));   // refresh the GEMM operand copy of W
}
template <typename Dtype>
void Net<Dtype>::ShareTrainedLayersWith(Net* other) {
  for (size_t i = 0; i < other->layers().size(); ++i) {
    Layer<Dtype>* source_layer = other->layers()[i].get();
    const string& source_layer_name = other->layer_names()[i];
    size_t target = 0;
    while (target != layer_names_.size() && layer_names_[target] != source_layer_name) ++target;
    if (target == layer_names_.size()) continue;
    vector<shared_ptr<Blob<Dtype> > >& target_blobs = layers_[target]->blobs();
    CHECK_EQ(target_blobs.size(), source_layer->blobs().size()) << "Incompatible number of blobs for layer " << source_layer_name;
    for (size_t j = 0; j < target_blobs.size(); ++j) {
      Blob<Dtype>* source_blob = source_layer->blobs()[j].get();
      CHECK_EQ(target_blobs[j]->num(), source_blob->num()); CHECK_EQ(target_blobs[j]->channels(), source_blob->channels());
      CHECK_EQ(target_blobs[j]->height(), source_blob->height()); CHECK_EQ(target_blobs[j]->width(), source_blob->width());
      target_blobs[j]->ShareData(*source_blob);
    }
  }
}
template <typename Dtype>
void Net<Dtype>::CopyTrainedLayersFrom(const string& trained_filename) {
  CopyTrainedLayersFrom(*ReadProtoFromBinaryFile(trained_filename, "NetParameter"));
}

template <typename Dtype>
void Net<Dtype>::CreateTrainer(const vv_trainer_cfg_t* solver_cfg) {
  CHECK(fused_data_) << "EnableFusion() did not match";
  if (!trainer_) {
    // first use: create the trainer with the solver constants, hand the current parameters over and make the
    // net's parameter blobs alias the trainer's buffers so params()/snapshots/debug keep reading live values
    if (solver_cfg) {
      vv_trainer_cfg_t c = fused_cfg_;
      memcpy(c.lr_policy, solver_cfg->lr_policy, sizeof(c.lr_policy));
      c.base_lr = solver_cfg->base_lr; c.gamma = solver_cfg->gamma; c.power = solver_cfg->power; c.stepsize = solver_cfg->stepsize;
      c.momentum = solver_cfg->momentum; c.weight_decay = solver_cfg->weight_decay; c.reg_type = solver_cfg->reg_type;
      fused_cfg_ = c;
    }
    if (fixed_mask_) fused_cfg_.dropout_mode = VV_DROPOUT_MASK01;
    trainer_ = vv_trainer_create(&fused_cfg_, Caffe::stream());
    CHECK(trainer_) << vv_last_error();
    Blob<Dtype>* W = fused_ip_->blobs()[0].get(); Blob<Dtype>* b = fused_ip_->blobs()[1].get();
    cudaStream_t s = reinterpret_cast<cudaStream_t>(Caffe::stream());
    CHECK_EQ(int(cudaMemcpyAsync(vv_trainer_weight(trainer_), W->gpu_data(), sizeof(Dtype) * W->count(), cudaMemcpyDeviceToDevice, s)), 0);
    CHECK_EQ(int(cudaMemcpyAsync(vv_trainer_bias(trainer_), b->gpu_data(), sizeof(Dtype) * b->count(), cudaMemcpyDeviceToDevice, s)), 0);
    CHECK_EQ(int(cudaStreamSynchronize(s)), 0);
    VV_CHECK(vv_trainer_sync_weights(trainer_));
    W->set_gpu_data(vv_trainer_weight(trainer_)); W->set_gpu_diff(vv_trainer_weight_diff(trainer_));
    b->set_gpu_data(vv_trainer_bias(trainer_)); b->set_gpu_diff(vv_trainer_bias_diff(trainer_));
    // the data layer's resident feature bank: registered once so the GEMMs gather its rows themselves
    // (a no-op for the precisions without gather producers); VV_MATERIALISE=1 keeps the K0 kernel
    const char* mat = getenv("VV_MATERIALISE");
    if (!(mat && mat[0] == '1')) VV_CHECK(vv_trainer_set_bank(trainer_, fused_data_->bank(), fused_data_->bank_rows()));
  }
}

template <typename Dtype>
Dtype Net<Dtype>::FusedStep(int iter, bool do_update, const vv_trainer_cfg_t* solver_cfg) {
  CreateTrainer(solver_cfg);
  const int32_t* dq = nullptr;
  const int32_t* di = fused_data_->NextIndices(&dq);
  VV_CHECK(vv_trainer_step(trainer_, fused_data_->bank(), fused_data_->bank_rows(), di, dq, fixed_mask_, iter, do_update ? 1 : 0));
  // the trainer wrote the aliased parameter buffers behind the blobs' back: move their heads to the device
  for (auto& pb : fused_ip_->blobs()) { pb->mutable_gpu_data(); pb->mutable_gpu_diff(); }
  // populate the net outputs other code reads (display, tests)
  cudaStream_t s = reinterpret_cast<cudaStream_t>(Caffe::stream());
  CHECK_EQ(int(cudaMemcpyAsync(fused_loss_->mutable_gpu_data(), vv_trainer_blob(trainer_, "loss"), sizeof(Dtype), cudaMemcpyDeviceToDevice, s)), 0);
  if (fused_viol_) CHECK_EQ(int(cudaMemcpyAsync(fused_viol_->mutable_gpu_data(), vv_trainer_blob(trainer_, "violations"), sizeof(Dtype), cudaMemcpyDeviceToDevice, s)), 0);
  return fused_loss_->cpu_data()[0] * fused_cfg_.loss_weight;
}

template class Net<float>;

}  // namespace caffe
