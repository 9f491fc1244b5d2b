// The 12 layer types of the shipped TRAIN net (ref: include/caffe/{common,neuron,loss,data}_layers.hpp),
// same class names, same virtuals, same parameters; every Forward_gpu / Backward_gpu body is one or a few
// calls into the C-ABI (include/vv_b200.h).  These are the bodies INTEGRATION.md pastes over the
// reference's .cu files.
#pragma once
#include <map>
#include "caffe/layer.hpp"

namespace caffe {

#define VV_LAYER_COMMON(Name, TYPE)                                                                        \
  virtual inline LayerParameter_LayerType type() const { return LayerParameter_LayerType_##TYPE; }          \
  virtual inline const string& type_name() const { static const string n = #TYPE; return n; }

// raw device scratch owned by a layer (operand copies, masks, workspaces)
class DeviceBuffer {
 public:
  DeviceBuffer() : p_(nullptr), bytes_(0) {}
  ~DeviceBuffer();
  void* get(size_t bytes);   // grows, never shrinks
 private:
  void* p_; size_t bytes_;
  DeviceBuffer(const DeviceBuffer&) = delete; DeviceBuffer& operator=(const DeviceBuffer&) = delete;
};

// ref: common_layers.hpp:309-339, inner_product_layer.{cpp,cu}
template <typename Dtype>
class InnerProductLayer : public Layer<Dtype> {
 public:
  explicit InnerProductLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  VV_LAYER_COMMON(InnerProduct, INNER_PRODUCT)
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom);
  vv_operand_t Operand(const Dtype* src, int64_t count, DeviceBuffer* hi, DeviceBuffer* lo);
  int M_, K_, N_;
  bool bias_term_;
  DeviceBuffer x_hi_, x_lo_, w_hi_, w_lo_, dz_hi_, dz_lo_, workspace_;
};

// ref: neuron_layers.hpp:23-45
template <typename Dtype>
class NeuronLayer : public Layer<Dtype> {
 public:
  explicit NeuronLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) { (*top)[0]->ReshapeLike(*bottom[0]); }
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
};
// ref: neuron_layers.hpp:295, relu_layer.{cpp,cu}
template <typename Dtype>
class ReLULayer : public NeuronLayer<Dtype> {
 public:
  explicit ReLULayer(const LayerParameter& param) : NeuronLayer<Dtype>(param) {}
  VV_LAYER_COMMON(ReLU, RELU)
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom);
};
// ref: neuron_layers.hpp:161-213, dropout_layer.{cpp,cu}
template <typename Dtype>
class DropoutLayer : public NeuronLayer<Dtype> {
 public:
  explicit DropoutLayer(const LayerParameter& param) : NeuronLayer<Dtype>(param) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  VV_LAYER_COMMON(Dropout, DROPOUT)
  // parity tests inject an explicit 0/1 mask (the reference's boost / curand masks are unpinned)
  void set_fixed_mask(const uint32_t* device_mask01) { fixed_mask_ = device_mask01; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom);
  DeviceBuffer rand_vec_;
  const uint32_t* fixed_mask_ = nullptr;
  const uint32_t* mask_in_use_ = nullptr;
  Dtype threshold_, scale_;
};

// ref: common_layers.hpp:519-558, slice_layer.{cpp,cu}
template <typename Dtype>
class SliceLayer : public Layer<Dtype> {
 public:
  explicit SliceLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  VV_LAYER_COMMON(Slice, SLICE)
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int MinTopBlobs() const { return 2; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom);
  int count_, num_, channels_, height_, width_, slice_dim_;
  vector<int> slice_point_;
};
// ref: common_layers.hpp:97-143, concat_layer.{cpp,cu}
template <typename Dtype>
class ConcatLayer : public Layer<Dtype> {
 public:
  explicit ConcatLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  VV_LAYER_COMMON(Concat, CONCAT)
  virtual inline int MinBottomBlobs() const { return 2; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom);
  int count_, num_, channels_, height_, width_, concat_dim_;
};
// ref: common_layers.hpp:220-250, flatten_layer.{cpp,cu}
template <typename Dtype>
class FlattenLayer : public Layer<Dtype> {
 public:
  explicit FlattenLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  VV_LAYER_COMMON(Flatten, FLATTEN)
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) { (*top)[0]->ShareData(*bottom[0]); }
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) { (*bottom)[0]->ShareDiff(*top[0]); }
};
// ref: common_layers.hpp:488-513, split_layer.{cpp,cu}
template <typename Dtype>
class SplitLayer : public Layer<Dtype> {
 public:
  explicit SplitLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  VV_LAYER_COMMON(Split, SPLIT)
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int MinTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom);
  int count_;
};
// ref: common_layers.hpp:160-196, eltwise_layer.{cpp,cu}
template <typename Dtype>
class EltwiseLayer : public Layer<Dtype> {
 public:
  explicit EltwiseLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  VV_LAYER_COMMON(Eltwise, ELTWISE)
  virtual inline int MinBottomBlobs() const { return 2; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom);
  EltwiseParameter_EltwiseOp op_;
  vector<Dtype> coeffs_;
  bool stable_prod_grad_;
};
// ref: common_layers.hpp:385-412, normalization_layer.{cpp,cu}
template <typename Dtype>
class NormalizationLayer : public Layer<Dtype> {
 public:
  explicit NormalizationLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) { (*top)[0]->ReshapeLike(*bottom[0]); }
  VV_LAYER_COMMON(Normalization, NORMALIZATION)
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom);
};
// ref: common_layers.hpp:626-654, sum_layer.{cpp,cu}
template <typename Dtype>
class SumLayer : public Layer<Dtype> {
 public:
  explicit SumLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  VV_LAYER_COMMON(Sum, SUM)
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom);
  int num_output_;
};
// ref: loss_layers.hpp:1142-1184, max_margin_loss_layer.cpp (CPU only in the reference; device here)
template <typename Dtype>
class MaxMarginLossLayer : public Layer<Dtype> {
 public:
  explicit MaxMarginLossLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  VV_LAYER_COMMON(MaxMarginLoss, MAX_MARGIN_LOSS)
  virtual inline int MinBottomBlobs() const { return 2; }
  virtual inline int MaxBottomBlobs() const { return 3; }
  virtual inline int MinTopBlobs() const { return 1; }
  virtual inline int MaxTopBlobs() const { return 2; }
  virtual inline bool AllowForceBackward(const int bottom_index) const { return bottom_index != 2; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom);
  const Dtype* Weights(const vector<Blob<Dtype>*>& bottom);
  float margin_;
  bool use_direct_weight_ = false;
  int table_size_ = 0;
  Blob<Dtype> scratch_;   // loss, violations on the device
  Blob<Dtype> table_ids_, table_w_, weights_;   // id_to_weight_file (ids as raw ints, ascending) and the looked-up weights
};
// ref: data_layers.hpp:223-286, video_sampled_shots_data_layer.cpp.  The DB reader is out of scope (no lmdb
// here): `source` is "synthetic://videos=V&shots=S&dim=K&seed=s", a resident feature bank filled on the
// device; the sampler is the reference's state machine emitting bank-row indices.
template <typename Dtype>
class VideoSampledShotsDataLayer : public Layer<Dtype> {
 public:
  explicit VideoSampledShotsDataLayer(const LayerParameter& param) : Layer<Dtype>(param), sampler_(nullptr) {}
  virtual ~VideoSampledShotsDataLayer();
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {}
  VV_LAYER_COMMON(VideoSampledShotsData, VIDEO_SAMPLED_SHOTS_DATA)
  virtual inline int ExactNumBottomBlobs() const { return 0; }
  virtual inline int MinTopBlobs() const { return 1; }
  virtual inline int MaxTopBlobs() const { return 2; }
  // access for the net-level fusion pass
  const float* bank() const { return bank_ptr_; }
  int64_t bank_rows() const { return bank_rows_; }
  const int32_t* NextIndices(const int32_t** quirk);   // draws a batch, returns device idx [B,R]
  int batch_size() const { return batch_size_; }
  int context_size() const { return context_size_; }
  int num_negative_samples() const { return num_negative_samples_; }
  int feature_size() const { return feature_size_; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {}
  vv_sampler_t* sampler_;
  DeviceBuffer bank_;            // [bank_rows_, feature_size_] fp32 (size_t bytes: real banks exceed a Blob's int count)
  float* bank_ptr_ = nullptr;
  int64_t bank_rows_;
  int batch_size_, context_size_, num_negative_samples_, feature_size_;
  vector<int32_t> idx_host_, quirk_host_;
  DeviceBuffer idx_dev_, quirk_dev_;
};

// ref: vision_layers.hpp (IdToWeightMappingLayer), id_to_weight_mapping_layer.cpp: a [max_ids, num_output] table indexed by
// the bottom's ids (one per item); the gradient flows to the table only.  Host loops in the reference, kernels here.
template <typename Dtype>
class IdToWeightMappingLayer : public Layer<Dtype> {
 public:
  explicit IdToWeightMappingLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  VV_LAYER_COMMON(IdToWeightMapping, ID_TO_WEIGHT_MAPPING)
  virtual inline int ExactNumBottomBlobs() const { return 1; }
  virtual inline int ExactNumTopBlobs() const { return 1; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom);
  int M_ = 0, K_ = 0, N_ = 0;
};
// ref: data_layers.hpp (VideoShotWindowTestDataLayer), video_shot_window_test_data_layer.cpp: the TEST-phase data layer,
// tops = data [B, F, K, 1] (the F context frames of a shot window) and video_ids [B].  As for the TRAIN layer the
// DB reader is out of scope: `source` is "synthetic://videos=V&shots=S&dim=K&seed=s&frames=F"; windows of F
// consecutive shots are served in order (video by video, wrapping), so every test pass sees the same data.
template <typename Dtype>
class VideoShotWindowTestDataLayer : public Layer<Dtype> {
 public:
  explicit VideoShotWindowTestDataLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {}
  VV_LAYER_COMMON(VideoShotWindowTestData, VIDEO_SHOT_WINDOW_TEST_DATA)
  virtual inline int ExactNumBottomBlobs() const { return 0; }
  virtual inline int ExactNumTopBlobs() const { return 2; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {}
  DeviceBuffer bank_;
  float* bank_ptr_ = nullptr;
  int64_t bank_rows_ = 0;
  int videos_ = 0, shots_ = 0, frames_ = 4, batch_size_ = 0, feature_size_ = 0;
  long cursor_ = 0;
  bool from_records_ = false;    // records: item c = rows [c*frames_, (c+1)*frames_) of the bank, label record_video_id_[c]
  vector<int32_t> record_video_id_;
  vector<int32_t> idx_host_;
  DeviceBuffer idx_dev_;
};
// ref: retrieval_stats_layer.cpp (CPU only in the reference; vv_retrieval_stats on the device here), shot level:
// bottoms = L2-normalised embeddings [B, N] and video_ids [B]; tops = mean AP, hit@1, hit@5.
template <typename Dtype>
class RetrievalStatsLayer : public Layer<Dtype> {
 public:
  explicit RetrievalStatsLayer(const LayerParameter& param) : Layer<Dtype>(param) {}
  virtual void LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  VV_LAYER_COMMON(RetrievalStats, RETRIEVAL_STATS)
  virtual inline int ExactNumBottomBlobs() const { return 2; }
  virtual inline int ExactNumTopBlobs() const { return 3; }
 protected:
  virtual void Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top);
  virtual void Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {}
  std::map<int, int> video_id_to_class_;
  bool exclude_same_video_shots_ = true;
  bool video_level_ = false;            // video_level_retrieval: rank per-video mean embeddings (ref: :160-206)
  int max_num_videos_ = 0;
  string stats_output_file_;            // per-query CSV (ref: :146-151, 306-340)
  DeviceBuffer ids_dev_, labels_dev_, work_, out_dev_, group_dev_, mean_dev_, top5_dev_;
};

}  // namespace caffe
