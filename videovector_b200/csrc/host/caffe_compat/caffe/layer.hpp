// Layer<Dtype>: the operator ABI of the reference (ref: include/caffe/layer.hpp:25-458), pointer-style
// `vector<Blob*>* top` signatures included.  GPU only: Forward_cpu / Backward_cpu abort (no CPU fallback).
#pragma once
#include "caffe/blob.hpp"
#include "caffe/proto/caffe_params.hpp"

namespace caffe {

template <typename Dtype>
class Layer {
 public:
  explicit Layer(const LayerParameter& param) : layer_param_(param) {}
  virtual ~Layer() {}
  void SetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
    CheckBlobCounts(bottom, *top);
    LayerSetUp(bottom, top);
    Reshape(bottom, top);
    SetLossWeights(top);
  }
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) = 0;
  inline Dtype Forward(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  inline void Backward(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom);
  vector<shared_ptr<Blob<Dtype> > >& blobs() { return blobs_; }
  const LayerParameter& layer_param() const { return layer_param_; }
  inline Dtype loss(const int top_index) const { return (int(loss_.size()) > top_index) ? loss_[top_index] : Dtype(0); }
  inline void set_loss(const int top_index, const Dtype value) {
    if (int(loss_.size()) <= top_index) loss_.resize(top_index + 1, Dtype(0));
    loss_[top_index] = value;
  }
  virtual inline LayerParameter_LayerType type() const { return LayerParameter_LayerType_NONE; }
  virtual inline const string& type_name() const { static const string n = ""; return n; }
  virtual inline int ExactNumBottomBlobs() const { return -1; }
  virtual inline int MinBottomBlobs() const { return -1; }
  virtual inline int MaxBottomBlobs() const { return -1; }
  virtual inline int ExactNumTopBlobs() const { return -1; }
  virtual inline int MinTopBlobs() const { return -1; }
  virtual inline int MaxTopBlobs() const { return -1; }
  virtual inline bool EqualNumBottomTopBlobs() const { return false; }
  virtual inline bool AutoTopBlobs() const { return false; }
  virtual inline bool AllowForceBackward(const int bottom_index) const { return true; }
  inline bool param_propagate_down(const int param_id) { return (int(param_propagate_down_.size()) > param_id) ? param_propagate_down_[param_id] : false; }
  inline void set_param_propagate_down(const int param_id, const bool value) {
    if (int(param_propagate_down_.size()) <= param_id) param_propagate_down_.resize(param_id + 1, true);
    param_propagate_down_[param_id] = value;
  }

 protected:
  LayerParameter layer_param_;
  vector<shared_ptr<Blob<Dtype> > > blobs_;
  vector<bool> param_propagate_down_;
  vector<Dtype> loss_;

  virtual void Forward_cpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) { NO_CPU; }
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) = 0;
  virtual void Backward_cpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) { NO_CPU; }
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) = 0;

  virtual void CheckBlobCounts(const vector<Blob<Dtype>*>& bottom, const vector<Blob<Dtype>*>& top) {
    if (ExactNumBottomBlobs() >= 0) CHECK_EQ(ExactNumBottomBlobs(), int(bottom.size())) << type_name() << " Layer takes " << ExactNumBottomBlobs() << " bottom blob(s) as input.";
    if (MinBottomBlobs() >= 0) CHECK_LE(MinBottomBlobs(), int(bottom.size())) << type_name() << " Layer takes at least " << MinBottomBlobs() << " bottom blob(s) as input.";
    if (MaxBottomBlobs() >= 0) CHECK_GE(MaxBottomBlobs(), int(bottom.size())) << type_name() << " Layer takes at most " << MaxBottomBlobs() << " bottom blob(s) as input.";
    if (ExactNumTopBlobs() >= 0) CHECK_EQ(ExactNumTopBlobs(), int(top.size())) << type_name() << " Layer produces " << ExactNumTopBlobs() << " top blob(s) as output.";
    if (MinTopBlobs() >= 0) CHECK_LE(MinTopBlobs(), int(top.size())) << type_name() << " Layer produces at least " << MinTopBlobs() << " top blob(s) as output.";
    if (MaxTopBlobs() >= 0) CHECK_GE(MaxTopBlobs(), int(top.size())) << type_name() << " Layer produces at most " << MaxTopBlobs() << " top blob(s) as output.";
    if (EqualNumBottomTopBlobs()) CHECK_EQ(bottom.size(), top.size()) << type_name() << " Layer produces one top blob as output for each bottom blob input.";
  }
  // loss weights live in the top blob's diff (ref: layer.hpp:387-404); the loss layer reads top[0]->diff[0]
  inline void SetLossWeights(vector<Blob<Dtype>*>* top) {
    const int num_loss_weights = layer_param_.loss_weight_size();
    if (num_loss_weights) {
      CHECK_EQ(int(top->size()), num_loss_weights) << "loss_weight must be unspecified or specified once per top blob.";
      for (int top_id = 0; top_id < int(top->size()); ++top_id) {
        const Dtype loss_weight = layer_param_.loss_weight(top_id);
        if (loss_weight == Dtype(0)) continue;
        this->set_loss(top_id, loss_weight);
        Dtype* loss_multiplier = (*top)[top_id]->mutable_cpu_diff();
        for (int i = 0; i < (*top)[top_id]->count(); ++i) loss_multiplier[i] = loss_weight;
      }
    }
  }
};

// Forward: run the device implementation, then loss = sum_top dot(data, diff) for tops with a loss weight
// (ref: layer.hpp:410-442; the dot is over 1-element blobs here, read back like caffe_gpu_dot does).
template <typename Dtype>
inline Dtype Layer<Dtype>::Forward(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  Dtype loss = 0;
  CHECK(Caffe::mode() == Caffe::GPU);
  Forward_gpu(bottom, top);
  for (int top_id = 0; top_id < int(top->size()); ++top_id) {
    if (!this->loss(top_id)) continue;
    const int count = (*top)[top_id]->count();
    const Dtype* data = (*top)[top_id]->cpu_data();
    const Dtype* loss_weights = (*top)[top_id]->cpu_diff();
    for (int i = 0; i < count; ++i) loss += data[i] * loss_weights[i];
  }
  return loss;
}
template <typename Dtype>
inline void Layer<Dtype>::Backward(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  CHECK(Caffe::mode() == Caffe::GPU);
  Backward_gpu(top, propagate_down, bottom);
}

template <typename Dtype> Layer<Dtype>* GetLayer(const LayerParameter& param);   // ref: src/caffe/layer_factory.cpp:177-309

}  // namespace caffe
