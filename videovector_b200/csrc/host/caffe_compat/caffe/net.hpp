// Net<Dtype>: builds the layer graph from a NetParameter and runs it (ref: include/caffe/net.hpp:23-229,
// src/caffe/net.cpp:34-224 Init, :227-268 FilterNet, :502-578 Forward/Backward, :804-839 Update;
// src/caffe/util/insert_splits.cpp:12-142).  Two execution modes for the shipped TRAIN graph:
//   layer-by-layer : every layer runs its own drop-in body (exact per-layer blobs);
//   fused          : EnableFusion() recognises the graph (data -> slice/concat/flatten -> fc7 -> relu ->
//                    dropout -> slice_emb ... max_margin_loss) and runs it as the K0..K3 + wgrad kernel
//                    sequence of vv_trainer, keeping loss_output / train_violations / param data+diff populated.
#pragma once
#include <map>
#include <set>
#include "caffe/layers.hpp"

namespace caffe {

NetParameter FilterNet(const NetParameter& param, Caffe::Phase phase);
NetParameter InsertSplits(const NetParameter& param);

template <typename Dtype>
class Net {
 public:
  Net(const NetParameter& param, Caffe::Phase phase = Caffe::TRAIN) : trainer_(nullptr) { Init(param, phase); }
  Net(const string& param_file, Caffe::Phase phase = Caffe::TRAIN) : trainer_(nullptr) { Init(ReadNetParamsFromTextFileOrDie(param_file), phase); }
  virtual ~Net();
  void Init(const NetParameter& param, Caffe::Phase phase);
  const vector<Blob<Dtype>*>& ForwardPrefilled(Dtype* loss = NULL);
  Dtype ForwardFromTo(int start, int end);
  void Backward();
  void BackwardFromTo(int start, int end);
  Dtype ForwardBackward(const vector<Blob<Dtype>*>& bottom = vector<Blob<Dtype>*>()) {
    Dtype loss;
    if (fused()) return FusedForwardBackward();
    ForwardPrefilled(&loss);
    Backward();
    return loss;
  }
  void Update();
  inline const string& name() const { return name_; }
  inline const vector<string>& layer_names() const { return layer_names_; }
  inline const vector<string>& blob_names() const { return blob_names_; }
  inline const vector<shared_ptr<Blob<Dtype> > >& blobs() const { return blobs_; }
  inline const vector<shared_ptr<Layer<Dtype> > >& layers() const { return layers_; }
  inline vector<shared_ptr<Blob<Dtype> > >& params() { return params_; }
  inline vector<float>& params_lr() { return params_lr_; }
  inline vector<float>& params_weight_decay() { return params_weight_decay_; }
  inline const vector<bool>& layer_need_backward() const { return layer_need_backward_; }
  inline const vector<vector<bool> >& bottom_need_backward() const { return bottom_need_backward_; }
  inline const vector<Blob<Dtype>*>& output_blobs() const { return net_output_blobs_; }
  inline const vector<int>& output_blob_indices() const { return net_output_blob_indices_; }
  bool has_blob(const string& blob_name) { return blob_names_index_.count(blob_name) > 0; }
  const shared_ptr<Blob<Dtype> > blob_by_name(const string& blob_name);
  bool has_layer(const string& layer_name) { return layer_names_index_.count(layer_name) > 0; }
  const shared_ptr<Layer<Dtype> > layer_by_name(const string& layer_name);

  // ---- trained-parameter IO (ref: net.cpp:692-801) ----
  // NetParameter message tree: name + every layer's parameter message + its blobs (data, optionally diff)
  shared_ptr<PbMsg> ToProto(bool write_diff = false) const;
  // copies blobs into layers of the same name; unknown source layers are ignored, shapes must match exactly
  void CopyTrainedLayersFrom(const PbMsg& net_param);
  void CopyTrainedLayersFrom(const string& trained_filename);
  // parameter blobs of same-named layers become views of `other`'s (ref: net.cpp:638-668): the TEST net reads the
  // weights the TRAIN net (or its fused trainer) is updating
  void ShareTrainedLayersWith(Net* other);
  inline const vector<Dtype>& blob_loss_weights() const { return blob_loss_weights_; }

  // ---- fusion ----
  // Returns true if the graph matched and the fused path is active.  `why` receives the reason if not.
  bool EnableFusion(string* why = nullptr);
  bool fused() const { return fused_data_ != nullptr; }
  vv_trainer_t* trainer() { return trainer_; }
  // fused step used by the solver: forward, backward and (optionally) the SGD update in one kernel sequence
  Dtype FusedStep(int iter, bool do_update, const vv_trainer_cfg_t* solver_cfg);
  // first use of the fused path: builds the trainer with the solver constants and hands the parameters over
  void CreateTrainer(const vv_trainer_cfg_t* solver_cfg);
  void set_fixed_dropout_mask(const uint32_t* device_mask01) { fixed_mask_ = device_mask01; }

 protected:
  Dtype FusedForwardBackward() { return FusedStep(0, false, nullptr); }
  int AppendBottom(const NetParameter& param, int layer_id, int bottom_id, std::set<string>* available, std::map<string, int>* name_to_idx);
  void AppendTop(const NetParameter& param, int layer_id, int top_id, std::set<string>* available, std::map<string, int>* name_to_idx);

  string name_;
  vector<shared_ptr<Layer<Dtype> > > layers_;
  vector<string> layer_names_;
  std::map<string, int> layer_names_index_, blob_names_index_;
  vector<bool> layer_need_backward_;
  vector<shared_ptr<Blob<Dtype> > > blobs_;
  vector<string> blob_names_;
  vector<bool> blob_need_backward_;
  vector<vector<Blob<Dtype>*> > bottom_vecs_, top_vecs_;
  vector<vector<int> > bottom_id_vecs_, top_id_vecs_;
  vector<vector<bool> > bottom_need_backward_;
  vector<Dtype> blob_loss_weights_;
  vector<Blob<Dtype>*> net_output_blobs_;
  vector<int> net_output_blob_indices_;
  vector<shared_ptr<Blob<Dtype> > > params_;
  vector<float> params_lr_, params_weight_decay_;
  // fused path
  vv_trainer_t* trainer_;
  vv_trainer_cfg_t fused_cfg_;
  VideoSampledShotsDataLayer<Dtype>* fused_data_ = nullptr;
  InnerProductLayer<Dtype>* fused_ip_ = nullptr;
  Blob<Dtype>* fused_loss_ = nullptr; Blob<Dtype>* fused_viol_ = nullptr;
  const uint32_t* fixed_mask_ = nullptr;
};

}  // namespace caffe
