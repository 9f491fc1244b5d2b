// Layer bodies: each Forward_gpu / Backward_gpu is a thin call sequence into the C-ABI.
// ref files are named per class in caffe/layers.hpp; semantics (what is overwritten, what is accumulated,
// which blob a mask or gate is read from) follow the reference's .cpp/.cu bodies.
#include <cuda_runtime_api.h>
#include <cmath>
#include <random>
#include <fstream>
#include "caffe/layers.hpp"

namespace caffe {

#define CUDA_CHECK(call) do { cudaError_t _e = (call); CHECK_EQ(int(_e), 0) << cudaGetErrorString(_e); } while (0)
static inline cudaStream_t cs() { return reinterpret_cast<cudaStream_t>(Caffe::stream()); }

DeviceBuffer::~DeviceBuffer() { if (p_) cudaFree(p_); }
void* DeviceBuffer::get(size_t bytes) {
  if (bytes > bytes_) {
    if (p_) { CUDA_CHECK(cudaStreamSynchronize(cs())); cudaFree(p_); }
    CUDA_CHECK(cudaMalloc(&p_, bytes));
    bytes_ = bytes;
  }
  return p_;
}

// ---- fillers (ref: include/caffe/filler.hpp:66-97 gaussian, :41-63 constant/uniform).  The reference draws
// from boost::mt19937 (values unpinned); std::mt19937 seeded from the Caffe seed is used here.
template <typename Dtype>
static void Fill(const FillerParameter& p, Blob<Dtype>* blob, unsigned salt) {
  Dtype* d = blob->mutable_cpu_data();
  const int n = blob->count();
  const string t = p.type();
  std::mt19937 gen(unsigned(Caffe::rng_seed()) * 2654435761u + salt);
  if (t == "constant") { for (int i = 0; i < n; ++i) d[i] = p.value(); }
  else if (t == "gaussian") { std::normal_distribution<float> nd(p.mean(), p.std()); for (int i = 0; i < n; ++i) d[i] = nd(gen); }
  else if (t == "uniform") { std::uniform_real_distribution<float> ud(p.min(), p.max()); for (int i = 0; i < n; ++i) d[i] = ud(gen); }
  else LOG_FATAL << "Unknown filler name: " << t;
}

// =================================== InnerProduct ===================================
template <typename Dtype>
void InnerProductLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const int num_output = this->layer_param_.inner_product_param().num_output();
  bias_term_ = this->layer_param_.inner_product_param().bias_term();
  N_ = num_output;
  K_ = bottom[0]->count() / bottom[0]->num();
  if (this->blobs_.size() > 0) {
    LogInfo("Skipping parameter initialization");
  } else {
    this->blobs_.resize(bias_term_ ? 2 : 1);
    this->blobs_[0].reset(new Blob<Dtype>(1, 1, N_, K_));
    Fill(this->layer_param_.inner_product_param().weight_filler(), this->blobs_[0].get(), 1);
    if (bias_term_) {
      this->blobs_[1].reset(new Blob<Dtype>(1, 1, 1, N_));
      Fill(this->layer_param_.inner_product_param().bias_filler(), this->blobs_[1].get(), 2);
    }
  }
  this->param_propagate_down_.resize(this->blobs_.size(), true);
}
template <typename Dtype>
void InnerProductLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  M_ = bottom[0]->num();
  CHECK_EQ(bottom[0]->count() / bottom[0]->num(), K_) << "Input size incompatible with inner product parameters.";
  (*top)[0]->Reshape(bottom[0]->num(), N_, 1, 1);
}
// fp32 blob -> operand copies for the configured precision (the fused trainer gets these for free from the
// producing kernels; the standalone layer converts per call)
template <typename Dtype>
vv_operand_t InnerProductLayer<Dtype>::Operand(const Dtype* src, int64_t count, DeviceBuffer* hi, DeviceBuffer* lo) {
  vv_operand_t o; o.hi = src; o.lo = nullptr;
  const int prec = Caffe::precision();
  if (prec == VV_PREC_TF32X3) {
    void* h = hi->get(count * 4); void* l = lo->get(count * 4);
    VV_CHECK(vv_prepare_operand(src, count, prec, h, l, Caffe::stream()));
    o.hi = h; o.lo = l;
  } else if (prec == VV_PREC_BF16) {
    void* h = hi->get(count * 2);
    VV_CHECK(vv_prepare_operand(src, count, prec, h, nullptr, Caffe::stream()));
    o.hi = h;
  } else if (prec == VV_PREC_F16X3) {
    // one block: header + two fp16 planes; prepare_operand measures max|src| and sets the scale
    size_t ho = 0, lo_off = 0;
    const size_t bytes = vv_operand_bytes(count, prec, &ho, &lo_off);
    char* base = static_cast<char*>(hi->get(bytes));
    VV_CHECK(vv_prepare_operand(src, count, prec, base + ho, base + lo_off, Caffe::stream()));
    o.hi = base + ho; o.lo = base + lo_off;
  }
  return o;
}
static bool TcShapeOk(int N, int K) { return (N % 8) == 0 && (K % 8) == 0; }
template <typename Dtype>
void InnerProductLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const int prec = TcShapeOk(N_, K_) ? Caffe::precision() : VV_PREC_FP32_SIMT;   // odd shapes: exact fp32 kernel
  const int saved = Caffe::precision(); Caffe::set_precision(prec);
  vv_operand_t X = Operand(bottom[0]->gpu_data(), int64_t(M_) * K_, &x_hi_, &x_lo_);
  vv_operand_t W = Operand(this->blobs_[0]->gpu_data(), int64_t(N_) * K_, &w_hi_, &w_lo_);
  Caffe::set_precision(saved);
  VV_CHECK(vv_ip_forward(X, W, bias_term_ ? this->blobs_[1]->gpu_data() : nullptr, M_, N_, K_, prec, nullptr, nullptr,
                         (*top)[0]->mutable_gpu_data(), Caffe::stream()));
}
template <typename Dtype>
void InnerProductLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down,
                                            vector<Blob<Dtype>*>* bottom) {
  const int prec = TcShapeOk(N_, K_) ? Caffe::precision() : VV_PREC_FP32_SIMT;
  const int saved = Caffe::precision(); Caffe::set_precision(prec);
  vv_operand_t dZ = Operand(top[0]->gpu_diff(), int64_t(M_) * N_, &dz_hi_, &dz_lo_);
  if (this->param_propagate_down_[0]) {
    vv_operand_t X = Operand((*bottom)[0]->gpu_data(), int64_t(M_) * K_, &x_hi_, &x_lo_);
    const size_t ws = vv_ip_wgrad_workspace_bytes(M_, N_, K_, prec);
    // dW is overwritten (beta = 0), then scaled by 1 + regularization/2 (ref: inner_product_layer.cpp:80-90)
    VV_CHECK(vv_ip_wgrad(dZ, X, M_, N_, K_, prec, this->layer_param_.inner_product_param().regularization(),
                         this->blobs_[0]->mutable_gpu_diff(), 0, ws ? workspace_.get(ws) : nullptr, ws, Caffe::stream()));
  }
  if (bias_term_ && this->param_propagate_down_[1])
    VV_CHECK(vv_ip_bias_grad(top[0]->gpu_diff(), M_, N_, this->blobs_[1]->mutable_gpu_diff(), Caffe::stream()));
  if (propagate_down[0]) {
    vv_operand_t W = Operand(this->blobs_[0]->gpu_data(), int64_t(N_) * K_, &w_hi_, &w_lo_);
    VV_CHECK(vv_ip_dgrad(dZ, W, M_, N_, K_, prec, (*bottom)[0]->mutable_gpu_diff(), Caffe::stream()));
  }
  Caffe::set_precision(saved);
}

// =================================== ReLU / Dropout ===================================
template <typename Dtype>
void ReLULayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  VV_CHECK(vv_relu_forward(bottom[0]->gpu_data(), bottom[0]->count(), this->layer_param_.relu_param().negative_slope(),
                           (*top)[0]->mutable_gpu_data(), Caffe::stream()));
}
template <typename Dtype>
void ReLULayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  if (!propagate_down[0]) return;
  // gate on the PRE-activation bottom data (ref: relu_layer.cu:35-42)
  VV_CHECK(vv_relu_backward((*bottom)[0]->gpu_data(), top[0]->gpu_diff(), (*bottom)[0]->count(),
                            this->layer_param_.relu_param().negative_slope(), (*bottom)[0]->mutable_gpu_diff(), Caffe::stream()));
}
template <typename Dtype>
void DropoutLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  threshold_ = this->layer_param_.dropout_param().dropout_ratio();
  CHECK(threshold_ > 0.) << "dropout_ratio must be in (0,1)";
  CHECK(threshold_ < 1.) << "dropout_ratio must be in (0,1)";
  scale_ = 1. / (1. - threshold_);
}
template <typename Dtype>
void DropoutLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const int count = bottom[0]->count();
  if (Caffe::phase() == Caffe::TRAIN) {
    if (fixed_mask_) {
      mask_in_use_ = fixed_mask_;
    } else {
      // one Philox sub-stream per Forward call, keep iff u32 > UINT_MAX*ratio (ref: dropout_layer.cu:14-36)
      uint32_t* m = static_cast<uint32_t*>(rand_vec_.get(size_t(count) * 4));
      const int rows = bottom[0]->num(), cols = count / rows;
      CHECK_EQ(cols % 4, 0) << "dropout mask generation needs an inner size that is a multiple of 4";
      VV_CHECK(vv_dropout_make_mask(m, rows, cols, threshold_, Caffe::rng_seed(), Caffe::next_rng_draw(), Caffe::stream()));
      mask_in_use_ = m;
    }
    VV_CHECK(vv_dropout_forward(bottom[0]->gpu_data(), mask_in_use_, 1 /*MASK01*/, count, threshold_,
                                (*top)[0]->mutable_gpu_data(), Caffe::stream()));
  } else if ((*top)[0]->gpu_data() != bottom[0]->gpu_data()) {
    (*top)[0]->CopyFrom(*bottom[0]);
  }
}
template <typename Dtype>
void DropoutLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  if (!propagate_down[0]) return;
  if (Caffe::phase() == Caffe::TRAIN) {
    CHECK(mask_in_use_) << "Dropout Backward before Forward";
    VV_CHECK(vv_dropout_backward(top[0]->gpu_diff(), mask_in_use_, 1, (*bottom)[0]->count(), threshold_,
                                 (*bottom)[0]->mutable_gpu_diff(), Caffe::stream()));
  } else if ((*bottom)[0]->gpu_diff() != top[0]->gpu_diff()) {
    (*bottom)[0]->CopyFrom(*top[0], true);
  }
}

// =================================== Slice / Concat / Flatten / Split ===================================
template <typename Dtype>
void SliceLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const SliceParameter sp = this->layer_param_.slice_param();
  slice_dim_ = sp.slice_dim();
  CHECK_GE(slice_dim_, 0); CHECK_LE(slice_dim_, 1) << "Can only slice num and channels";
  slice_point_.clear();
  for (int i = 0; i < sp.slice_point_size(); ++i) slice_point_.push_back(sp.slice_point(i));
}
template <typename Dtype>
void SliceLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  count_ = 0; num_ = bottom[0]->num(); channels_ = bottom[0]->channels(); height_ = bottom[0]->height(); width_ = bottom[0]->width();
  const int ntop = top->size();
  const int extent = slice_dim_ == 0 ? num_ : channels_;
  vector<int> sizes;
  if (!slice_point_.empty()) {
    CHECK_EQ(int(slice_point_.size()), ntop - 1);
    int prev = 0;
    for (size_t i = 0; i < slice_point_.size(); ++i) { CHECK_GT(slice_point_[i], prev); sizes.push_back(slice_point_[i] - prev); prev = slice_point_[i]; }
    sizes.push_back(extent - prev);
  } else {
    CHECK_EQ(extent % ntop, 0) << "Number of top blobs (" << ntop << ") should evenly divide the input slice dimension (" << extent << ")";
    sizes.assign(ntop, extent / ntop);
  }
  for (int i = 0; i < ntop; ++i) {
    if (slice_dim_ == 0) (*top)[i]->Reshape(sizes[i], channels_, height_, width_);
    else (*top)[i]->Reshape(num_, sizes[i], height_, width_);
    count_ += (*top)[i]->count();
  }
  CHECK_EQ(count_, bottom[0]->count());
}
template <typename Dtype>
void SliceLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const Dtype* src = bottom[0]->gpu_data();
  const int64_t inner = int64_t(height_) * width_;
  int64_t off = 0;
  for (size_t i = 0; i < top->size(); ++i) {
    Blob<Dtype>* t = (*top)[i];
    if (slice_dim_ == 0) {
      VV_CHECK(vv_copy_strided(src + off, t->count(), t->mutable_gpu_data(), t->count(), 1, t->count(), Caffe::stream()));
      off += t->count();
    } else {   // one strided copy instead of one cudaMemcpy per sample (ref: slice_layer.cu:22-33)
      const int64_t cols = t->channels() * inner;
      VV_CHECK(vv_copy_strided(src + off, channels_ * inner, t->mutable_gpu_data(), cols, num_, cols, Caffe::stream()));
      off += cols;
    }
  }
}
template <typename Dtype>
void SliceLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  if (!propagate_down[0]) return;
  Dtype* dst = (*bottom)[0]->mutable_gpu_diff();
  const int64_t inner = int64_t(height_) * width_;
  int64_t off = 0;
  for (size_t i = 0; i < top.size(); ++i) {
    Blob<Dtype>* t = top[i];
    if (slice_dim_ == 0) {
      VV_CHECK(vv_copy_strided(t->gpu_diff(), t->count(), dst + off, t->count(), 1, t->count(), Caffe::stream()));
      off += t->count();
    } else {
      const int64_t cols = t->channels() * inner;
      VV_CHECK(vv_copy_strided(t->gpu_diff(), cols, dst + off, channels_ * inner, num_, cols, Caffe::stream()));
      off += cols;
    }
  }
}
template <typename Dtype>
void ConcatLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  concat_dim_ = this->layer_param_.concat_param().concat_dim();
  CHECK_GE(concat_dim_, 0) << "concat_dim should be >= 0";
  CHECK_LE(concat_dim_, 1) << "For now concat_dim <=1, it can only concat num and channels";
}
template <typename Dtype>
void ConcatLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  count_ = bottom[0]->count(); num_ = bottom[0]->num(); channels_ = bottom[0]->channels();
  height_ = bottom[0]->height(); width_ = bottom[0]->width();
  for (size_t i = 1; i < bottom.size(); ++i) {
    count_ += bottom[i]->count();
    if (concat_dim_ == 0) num_ += bottom[i]->num();
    else { channels_ += bottom[i]->channels(); CHECK_EQ(bottom[i]->num(), num_); }
    CHECK_EQ(bottom[i]->height(), height_); CHECK_EQ(bottom[i]->width(), width_);
  }
  (*top)[0]->Reshape(num_, channels_, height_, width_);
  CHECK_EQ(count_, (*top)[0]->count());
}
template <typename Dtype>
void ConcatLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  Dtype* dst = (*top)[0]->mutable_gpu_data();
  const int64_t inner = int64_t(height_) * width_;
  int64_t off = 0;
  for (size_t i = 0; i < bottom.size(); ++i) {
    if (concat_dim_ == 0) {
      VV_CHECK(vv_copy_strided(bottom[i]->gpu_data(), bottom[i]->count(), dst + off, bottom[i]->count(), 1, bottom[i]->count(), Caffe::stream()));
      off += bottom[i]->count();
    } else {
      const int64_t cols = bottom[i]->channels() * inner;
      VV_CHECK(vv_copy_strided(bottom[i]->gpu_data(), cols, dst + off, channels_ * inner, num_, cols, Caffe::stream()));
      off += cols;
    }
  }
}
template <typename Dtype>
void ConcatLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  const Dtype* src = top[0]->gpu_diff();
  const int64_t inner = int64_t(height_) * width_;
  int64_t off = 0;
  for (size_t i = 0; i < bottom->size(); ++i) {
    Blob<Dtype>* b = (*bottom)[i];
    const int64_t cols = concat_dim_ == 0 ? b->count() : b->channels() * inner;
    if (propagate_down[i]) {
      if (concat_dim_ == 0) VV_CHECK(vv_copy_strided(src + off, cols, b->mutable_gpu_diff(), cols, 1, cols, Caffe::stream()));
      else VV_CHECK(vv_copy_strided(src + off, channels_ * inner, b->mutable_gpu_diff(), cols, num_, cols, Caffe::stream()));
    }
    off += cols;
  }
}
template <typename Dtype>
void FlattenLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  (*top)[0]->Reshape(bottom[0]->num(), bottom[0]->channels() * bottom[0]->height() * bottom[0]->width(), 1, 1);
}
template <typename Dtype>
void SplitLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  count_ = bottom[0]->count();
  for (size_t i = 0; i < top->size(); ++i) {
    CHECK((*top)[i] != bottom[0]) << "Layer does not allow in-place computation.";
    (*top)[i]->ReshapeLike(*bottom[0]);
  }
}
template <typename Dtype>
void SplitLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  for (size_t i = 0; i < top->size(); ++i) (*top)[i]->ShareData(*bottom[0]);
}
template <typename Dtype>
void SplitLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  if (!propagate_down[0]) return;
  Dtype* d = (*bottom)[0]->mutable_gpu_diff();
  // d = top0 + top1 ; d += top_k  (ref: split_layer.cu:17-33) -- same association order
  VV_CHECK(vv_axpby(count_, 1.f, top[0]->gpu_diff(), 0.f, d, Caffe::stream()));
  for (size_t i = 1; i < top.size(); ++i) VV_CHECK(vv_axpby(count_, 1.f, top[i]->gpu_diff(), 1.f, d, Caffe::stream()));
}

// =================================== Eltwise / Normalization / Sum ===================================
template <typename Dtype>
void EltwiseLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const EltwiseParameter ep = this->layer_param_.eltwise_param();
  CHECK(ep.coeff_size() == 0 || ep.coeff_size() == int(bottom.size())) << "Eltwise Layer takes one coefficient per bottom blob.";
  CHECK(!(ep.operation() == EltwiseParameter_EltwiseOp_PROD && ep.coeff_size())) << "Eltwise layer only takes coefficients for summation.";
  op_ = ep.operation();
  coeffs_ = vector<Dtype>(bottom.size(), 1);
  for (int i = 0; i < ep.coeff_size(); ++i) coeffs_[i] = ep.coeff(i);
  stable_prod_grad_ = ep.stable_prod_grad();
  CHECK_LE(int(bottom.size()), VV_MAX_CONTEXT);
}
template <typename Dtype>
void EltwiseLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  for (size_t i = 1; i < bottom.size(); ++i) {
    CHECK_EQ(bottom[0]->num(), bottom[i]->num()); CHECK_EQ(bottom[0]->channels(), bottom[i]->channels());
    CHECK_EQ(bottom[0]->height(), bottom[i]->height()); CHECK_EQ(bottom[0]->width(), bottom[i]->width());
  }
  (*top)[0]->ReshapeLike(*bottom[0]);
}
template <typename Dtype>
void EltwiseLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const int count = (*top)[0]->count();
  Dtype* t = (*top)[0]->mutable_gpu_data();
  switch (op_) {
    case EltwiseParameter_EltwiseOp_PROD:
      VV_CHECK(vv_eltwise_prod_forward(bottom[0]->gpu_data(), bottom[1]->gpu_data(), count, t, Caffe::stream()));
      for (size_t i = 2; i < bottom.size(); ++i) VV_CHECK(vv_mul(count, t, bottom[i]->gpu_data(), t, Caffe::stream()));
      break;
    case EltwiseParameter_EltwiseOp_SUM: {
      vector<const float*> ptrs(bottom.size());
      for (size_t i = 0; i < bottom.size(); ++i) ptrs[i] = bottom[i]->gpu_data();
      VV_CHECK(vv_eltwise_sum_forward(ptrs.data(), coeffs_.data(), int(bottom.size()), count, t, Caffe::stream()));
      break;
    }
    default: LOG_FATAL << "Eltwise MAX is not on the temporal-embedding path (not built for B200)";
  }
}
template <typename Dtype>
void EltwiseLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  const int count = top[0]->count();
  const Dtype* top_diff = top[0]->gpu_diff();
  for (size_t i = 0; i < bottom->size(); ++i) {
    if (!propagate_down[i]) continue;
    Dtype* bd = (*bottom)[i]->mutable_gpu_diff();
    if (op_ == EltwiseParameter_EltwiseOp_PROD) {
      CHECK(stable_prod_grad_) << "unstable PROD gradient (top / bottom) is not built; stable_prod_grad defaults to true";
      bool initialized = false;      // product of the OTHER bottoms, then times top_diff (ref: eltwise_layer.cu:96-119)
      for (size_t j = 0; j < bottom->size(); ++j) {
        if (i == j) continue;
        if (!initialized) { VV_CHECK(vv_axpby(count, 1.f, (*bottom)[j]->gpu_data(), 0.f, bd, Caffe::stream())); initialized = true; }
        else VV_CHECK(vv_mul(count, (*bottom)[j]->gpu_data(), bd, bd, Caffe::stream()));
      }
      VV_CHECK(vv_mul(count, bd, top_diff, bd, Caffe::stream()));
    } else if (op_ == EltwiseParameter_EltwiseOp_SUM) {
      VV_CHECK(vv_axpby(count, coeffs_[i], top_diff, 0.f, bd, Caffe::stream()));
    } else {
      LOG_FATAL << "Eltwise MAX is not on the temporal-embedding path";
    }
  }
}
template <typename Dtype>
void NormalizationLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const int num = bottom[0]->num();
  VV_CHECK(vv_l2norm_forward(bottom[0]->gpu_data(), num, bottom[0]->count() / num, (*top)[0]->mutable_gpu_data(), Caffe::stream()));
}
template <typename Dtype>
void NormalizationLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  const int num = (*bottom)[0]->num();
  VV_CHECK(vv_l2norm_backward((*bottom)[0]->gpu_data(), top[0]->gpu_diff(), num, (*bottom)[0]->count() / num,
                              (*bottom)[0]->mutable_gpu_diff(), Caffe::stream()));
}
template <typename Dtype>
void SumLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  num_output_ = int(this->layer_param_.sum_param().num_output());   // declared `float` in the proto (:742-744)
  (*top)[0]->Reshape(bottom[0]->num(), num_output_, 1, 1);
}
template <typename Dtype>
void SumLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const int num = bottom[0]->num();
  VV_CHECK(vv_rowsum_forward(bottom[0]->gpu_data(), num, bottom[0]->count() / num, num_output_, (*top)[0]->mutable_gpu_data(), Caffe::stream()));
}
template <typename Dtype>
void SumLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  const int num = (*bottom)[0]->num();
  VV_CHECK(vv_rowsum_backward(top[0]->gpu_diff(), num, (*bottom)[0]->count() / num, num_output_, (*bottom)[0]->mutable_gpu_diff(), Caffe::stream()));
}

// =================================== MaxMarginLoss ===================================
template <typename Dtype>
void MaxMarginLossLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  // LossLayer::LayerSetUp: a loss layer's first top has loss_weight 1 unless specified (ref: loss_layer.cpp:13-20)
  if (this->layer_param_.loss_weight_size() == 0) this->layer_param_.add_loss_weight(Dtype(1));
  // per-video loss weights (ref: max_margin_loss_layer.cpp:18-39): "video_id,weight" lines; the table goes to the device
  // sorted by id, the lookup runs there (vv_id_to_weight)
  const string file = this->layer_param_.max_margin_loss_param().id_to_weight_file();
  if (file != "") {
    std::ifstream in(file.c_str());
    std::map<int, float> table;
    string line;
    while (std::getline(in, line)) {
      const size_t comma = line.find(',');
      CHECK(comma != string::npos && line.find(',', comma + 1) == string::npos) << "Line: " << line;
      char* end = nullptr;
      const string a = line.substr(0, comma), b = line.substr(comma + 1);
      const long id = strtol(a.c_str(), &end, 10);
      CHECK(end == a.c_str() + a.size() && !a.empty()) << "Line: " << line;
      const float w = strtof(b.c_str(), &end);
      CHECK(end == b.c_str() + b.size() && !b.empty()) << "Line: " << line;
      CHECK_GE(w, 0) << "All weights should be greater than 0";
      table.insert(std::make_pair(int(id), w));            // first occurrence wins, as map::insert does
    }
    if (!table.empty()) {
      table_ids_.Reshape(1, 1, 1, int(table.size())); table_w_.Reshape(1, 1, 1, int(table.size()));
      // ids travel as raw ints inside a float blob's storage (never used as floats)
      int* ids = reinterpret_cast<int*>(table_ids_.mutable_cpu_data());
      float* ws = reinterpret_cast<float*>(table_w_.mutable_cpu_data());
      size_t i = 0;
      for (const auto& kv : table) { ids[i] = kv.first; ws[i] = kv.second; ++i; }
    }
    table_size_ = int(table.size());
  }
  use_direct_weight_ = this->layer_param_.max_margin_loss_param().use_direct_weight();
  margin_ = this->layer_param_.max_margin_loss_param().margin();
}
template <typename Dtype>
void MaxMarginLossLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  CHECK_EQ(bottom[0]->num(), bottom[1]->num()) << "The data and label should have the same number.";
  CHECK_EQ(bottom[0]->count(), bottom[1]->count());
  (*top)[0]->Reshape(1, 1, 1, 1);
  if (top->size() >= 2) (*top)[1]->Reshape(1, 1, 1, 1);
  scratch_.Reshape(1, 1, 1, 2);
  if (bottom.size() == 3) {
    CHECK_EQ(bottom[2]->count(), bottom[0]->count()) << "one weight / video id per score";
    if (!use_direct_weight_) weights_.Reshape(1, 1, 1, bottom[0]->count());
  }
}
// the third bottom as per-element weights: itself (use_direct_weight) or its video ids through the table
template <typename Dtype>
const Dtype* MaxMarginLossLayer<Dtype>::Weights(const vector<Blob<Dtype>*>& bottom) {
  if (bottom.size() != 3) return nullptr;
  if (use_direct_weight_) return bottom[2]->gpu_data();
  VV_CHECK(vv_id_to_weight(bottom[2]->gpu_data(), bottom[2]->count(),
                           table_size_ ? reinterpret_cast<const int*>(table_ids_.gpu_data()) : nullptr,
                           table_size_ ? table_w_.gpu_data() : nullptr, table_size_, weights_.mutable_gpu_data(), Caffe::stream()));
  return weights_.gpu_data();
}
template <typename Dtype>
void MaxMarginLossLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const int count = bottom[0]->count();
  const int norm = this->layer_param_.max_margin_loss_param().norm() == MaxMarginLossParameter_Norm_L2 ? 2 : 1;
  // the reference uses bottom[0].diff as the hinge scratch (max_margin_loss_layer.cpp:62-69); kept for blob parity
  VV_CHECK(vv_max_margin_forward_w(bottom[0]->gpu_data(), bottom[1]->gpu_data(), Weights(bottom), count, margin_, norm,
                                   bottom[0]->mutable_gpu_diff(), (*top)[0]->mutable_gpu_data(),
                                   top->size() > 1 ? (*top)[1]->mutable_gpu_data() : scratch_.mutable_gpu_data() + 1, Caffe::stream()));
}
template <typename Dtype>
void MaxMarginLossLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  if (!(propagate_down[0] || propagate_down[1])) return;
  const int count = (*bottom)[0]->count();
  const int norm = this->layer_param_.max_margin_loss_param().norm() == MaxMarginLossParameter_Norm_L2 ? 2 : 1;
  const Dtype loss_weight = top[0]->cpu_diff()[0];
  VV_CHECK(vv_max_margin_backward_w((*bottom)[0]->gpu_data(), (*bottom)[1]->gpu_data(), Weights(*bottom), count, margin_, norm, loss_weight,
                                    propagate_down[0] ? (*bottom)[0]->mutable_gpu_diff() : nullptr,
                                    (*bottom)[1]->mutable_gpu_diff(), Caffe::stream()));
}

// =================================== VideoSampledShotsData ===================================
static long UrlParam(const string& src, const string& key, long dflt) {
  const size_t p = src.find(key + "=");
  if (p == string::npos) return dflt;
  return strtol(src.c_str() + p + key.size() + 1, nullptr, 10);
}
template <typename Dtype>
VideoSampledShotsDataLayer<Dtype>::~VideoSampledShotsDataLayer() { if (sampler_) vv_sampler_destroy(sampler_); }
namespace {
// One `source` / `negative_dataset` of the data layer: its tables now, its rows into the bank later (one allocation holds
// the main set's rows followed by the negative set's).
struct ShotSource {
  int V = 0, K = 0; int64_t rows = 0;
  vector<int32_t> vid, off, ids;
  vv_record_set_t* rs = nullptr;        // record-file sources
  uint64_t seed = 0;                    // synthetic:// sources
  ~ShotSource() { if (rs) vv_record_set_destroy(rs); }
  void Open(const string& src) {
    if (src.compare(0, 12, "synthetic://") == 0) {
      V = int(UrlParam(src, "videos", 2048));
      const int S = int(UrlParam(src, "shots", 32));
      K = int(UrlParam(src, "dim", 4096));
      seed = uint64_t(UrlParam(src, "seed", 1234));
      CHECK_GE(K, 1);
      rows = int64_t(V) * S;
      vid.resize(V); off.resize(V + 1); ids.resize(size_t(V) * S);
      for (int v = 0; v < V; ++v) { vid[v] = v; off[v] = v * S; for (int s2 = 0; s2 < S; ++s2) ids[size_t(v) * S + s2] = s2; }
      off[V] = V * S;
      return;
    }
    // the reference's `source`: an LMDB environment of VideoShots records (:121-135), or a VVRS / mdb_dump file of them;
    // every record is decoded once into the resident device bank, the cursor loop becomes the sampler's index stream
    rs = vv_record_set_create(VV_RECORD_VIDEO_SHOTS, 1, 1);
    CHECK(rs) << vv_last_error();
    const int rc = vv_record_set_load_file(rs, src.c_str());
    if (rc != 0) { const string e = vv_last_error(); LOG_FATAL << "cannot read VideoShots records from '" << src << "': " << e; }
    int64_t records = 0; int32_t k = 0;
    VV_CHECK(vv_record_set_info(rs, &records, &rows, &k, nullptr));
    if (records < 1 || rows < 1) LOG_FATAL << "no VideoShots records in '" << src << "'";
    V = int(records); K = k;
    vid.resize(V); off.resize(V + 1); ids.resize(size_t(rows));
    VV_CHECK(vv_record_set_tables(rs, vid.data(), off.data(), ids.data()));
    LogInfo("VideoShots records: " + std::to_string(records) + " videos, " + std::to_string(rows) + " shots, feature size " + std::to_string(K));
  }
  void Upload(float* dst) {
    if (rs) { VV_CHECK(vv_record_set_upload(rs, dst, Caffe::stream())); }
    else VV_CHECK(vv_fill_bank(dst, rows, K, seed, Caffe::stream()));
  }
};
}  // namespace
template <typename Dtype>
void VideoSampledShotsDataLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const VideoSampledShotsDataParameter p = this->layer_param_.video_sampled_shots_data_param();
  batch_size_ = p.batch_size(); context_size_ = p.context_size(); num_negative_samples_ = p.num_negative_samples();
  ShotSource main_src, neg_src;
  main_src.Open(p.source());
  // negative_dataset (ref: :104-153): a second record set that seeds the negative buffer; its rows follow the main set's
  const bool has_neg = !p.negative_dataset().empty();
  if (has_neg) {
    neg_src.Open(p.negative_dataset());
    CHECK_EQ(neg_src.K, main_src.K) << "negative_dataset has another feature size";
  }
  feature_size_ = main_src.K;
  bank_rows_ = main_src.rows + (has_neg ? neg_src.rows : 0);
  bank_ptr_ = static_cast<float*>(bank_.get(size_t(bank_rows_) * feature_size_ * sizeof(float)));
  main_src.Upload(bank_ptr_);
  if (has_neg) neg_src.Upload(bank_ptr_ + size_t(main_src.rows) * feature_size_);
  CHECK_GE(feature_size_, 1); CHECK_GE(context_size_, 2); CHECK_GE(batch_size_, 1);
  // rand_skip (ref: :157-180): skip = caffe_rng_rand() % rand_skip records, caffe_rng_rand() being the next output of the
  // mt19937 that Caffe::set_random_seed seeded -- its first output here, nothing else has drawn from it at set-up
  int skip = 0;
  if (p.rand_skip() > 0) {
    std::mt19937 gen(unsigned(Caffe::rng_seed()));
    skip = int(gen() % unsigned(p.rand_skip()));
    LogInfo("Skipping first " + std::to_string(skip) + " data points.");
  }
  sampler_ = vv_sampler_create_ex2(main_src.V, main_src.vid.data(), main_src.off.data(), main_src.ids.data(), batch_size_,
                                   context_size_, num_negative_samples_, p.max_buffer_size(), p.negative_swap_percentage(),
                                   p.max_same_video_negs(), 100, 1 /* rand() is never seeded */, int(p.context_type()), skip,
                                   has_neg ? neg_src.V : 0, has_neg ? neg_src.vid.data() : nullptr,
                                   has_neg ? neg_src.off.data() : nullptr, has_neg ? neg_src.ids.data() : nullptr,
                                   int32_t(main_src.rows));
  CHECK(sampler_) << "Could not add requested number of negatives (or an invalid context_size for this context_type)";
  VV_CHECK(vv_sampler_prefetch(sampler_, 3));      // BasePrefetchingDataLayer: the next batches are drawn on a thread
  const int R = context_size_ + num_negative_samples_;
  (*top)[0]->Reshape(batch_size_, R, feature_size_, 1);     // channels = slots, height = feature (ref: :215-220)
  if (top->size() > 1) (*top)[1]->Reshape(batch_size_, 1, 1, 1);
  idx_host_.resize(size_t(batch_size_) * R); quirk_host_.resize(size_t(batch_size_) * R);
}
template <typename Dtype>
const int32_t* VideoSampledShotsDataLayer<Dtype>::NextIndices(const int32_t** quirk) {
  VV_CHECK(vv_sampler_next(sampler_, idx_host_.data(), quirk_host_.data()));
  const size_t bytes = idx_host_.size() * sizeof(int32_t);
  int32_t* di = static_cast<int32_t*>(idx_dev_.get(bytes));
  int32_t* dq = static_cast<int32_t*>(quirk_dev_.get(bytes));
  CUDA_CHECK(cudaMemcpyAsync(di, idx_host_.data(), bytes, cudaMemcpyHostToDevice, cs()));
  CUDA_CHECK(cudaMemcpyAsync(dq, quirk_host_.data(), bytes, cudaMemcpyHostToDevice, cs()));
  CUDA_CHECK(cudaStreamSynchronize(cs()));   // the host vectors are reused by the next draw
  *quirk = dq;
  return di;
}
template <typename Dtype>
void VideoSampledShotsDataLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const int32_t* dq = nullptr;
  const int32_t* di = NextIndices(&dq);
  // the data blob [B, R, K, 1] itself (what the reference's prefetch thread assembles on the host)
  VV_CHECK(vv_gather_rows(bank(), bank_rows_, feature_size_, di, dq, batch_size_, context_size_ + num_negative_samples_,
                          nullptr, nullptr, nullptr, VV_PREC_FP32_SIMT, (*top)[0]->mutable_gpu_data(), Caffe::stream()));
}

// =================================== IdToWeightMapping ===================================
template <typename Dtype>
void IdToWeightMappingLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const IdToWeightMappingParameter p = this->layer_param_.id_to_weight_mapping_param();
  K_ = p.max_ids(); N_ = p.num_output();
  CHECK_GE(K_, 1); CHECK_GE(N_, 1);
  CHECK_EQ(bottom[0]->count(), bottom[0]->num());
  if (this->blobs_.size() > 0) {
    LogInfo("Skipping parameter initialization");
  } else {
    this->blobs_.resize(1);
    this->blobs_[0].reset(new Blob<Dtype>(K_, N_, 1, 1));
    Fill(p.weight_filler(), this->blobs_[0].get(), 3);
  }
  this->param_propagate_down_.resize(this->blobs_.size(), true);
}
template <typename Dtype>
void IdToWeightMappingLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  M_ = bottom[0]->num();
  CHECK_EQ(bottom[0]->count(), M_) << "Input size incompatible with inner product parameters.";
  (*top)[0]->Reshape(M_, N_, 1, 1);
}
template <typename Dtype>
void IdToWeightMappingLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const Dtype* ids = bottom[0]->cpu_data();
  for (int i = 0; i < M_; ++i) CHECK(int(ids[i]) >= 0 && int(ids[i]) < K_) << "id " << ids[i] << " outside [0, max_ids)";
  VV_CHECK(vv_id_lookup_forward(this->blobs_[0]->gpu_data(), K_, N_, bottom[0]->gpu_data(), M_, (*top)[0]->mutable_gpu_data(), Caffe::stream()));
}
template <typename Dtype>
void IdToWeightMappingLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down,
                                                 vector<Blob<Dtype>*>* bottom) {
  CHECK(!propagate_down[0]) << this->type_name() << "Layer cannot backpropogate to input ids.";
  if (this->param_propagate_down_[0])
    VV_CHECK(vv_id_lookup_backward(top[0]->gpu_diff(), (*bottom)[0]->gpu_data(), M_, N_, K_, this->blobs_[0]->mutable_gpu_diff(), Caffe::stream()));
}

// =================================== TEST-phase data + retrieval stats ===================================
template <typename Dtype>
void VideoShotWindowTestDataLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const VideoShotWindowTestDataParameter p = this->layer_param_.video_shot_window_test_data_param();
  const string src = p.source();
  batch_size_ = p.batch_size();
  CHECK_GE(batch_size_, 1);
  if (src.compare(0, 12, "synthetic://") == 0) {
    videos_ = int(UrlParam(src, "videos", 256)); shots_ = int(UrlParam(src, "shots", 32));
    feature_size_ = int(UrlParam(src, "dim", 4096)); frames_ = int(UrlParam(src, "frames", 4));
    const uint64_t seed = uint64_t(UrlParam(src, "seed", 4321));
    CHECK_GE(frames_, 1); CHECK_GE(shots_, frames_);
    bank_rows_ = int64_t(videos_) * shots_;
    bank_ptr_ = static_cast<float*>(bank_.get(size_t(bank_rows_) * feature_size_ * sizeof(float)));
    VV_CHECK(vv_fill_bank(bank_ptr_, bank_rows_, feature_size_, seed, Caffe::stream()));
  } else {
    // TestVideoShotWindows records (video_shot_window_test_data_layer.cpp:76-135): channels = context (+ positive
    // + negative) datums of one record, label = its video_id, the cursor wraps at the end (:241-262)
    vv_record_set_t* rs = vv_record_set_create(VV_RECORD_TEST_WINDOWS, p.include_positives(), p.include_negatives());
    CHECK(rs) << vv_last_error();
    const int rc = vv_record_set_load_file(rs, src.c_str());
    if (rc != 0) { const string e = vv_last_error(); vv_record_set_destroy(rs); LOG_FATAL << "cannot read TestVideoShotWindows records from '" << src << "': " << e; }
    int64_t records = 0, rows = 0; int32_t K = 0, per = 0;
    VV_CHECK(vv_record_set_info(rs, &records, &rows, &K, &per));
    if (records < 1) { vv_record_set_destroy(rs); LOG_FATAL << "no TestVideoShotWindows records in '" << src << "'"; }
    from_records_ = true; videos_ = int(records); frames_ = per; feature_size_ = K; bank_rows_ = rows;
    record_video_id_.resize(size_t(records));
    VV_CHECK(vv_record_set_tables(rs, record_video_id_.data(), nullptr, nullptr));
    bank_ptr_ = static_cast<float*>(bank_.get(size_t(rows) * K * sizeof(float)));
    const int urc = vv_record_set_upload(rs, bank_ptr_, Caffe::stream());
    vv_record_set_destroy(rs);
    VV_CHECK(urc);
  }
  (*top)[0]->Reshape(batch_size_, frames_, feature_size_, 1);     // channels = frames, height = feature
  (*top)[1]->Reshape(batch_size_, 1, 1, 1);
  idx_host_.resize(size_t(batch_size_) * frames_);
}
template <typename Dtype>
void VideoShotWindowTestDataLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  // record c = window number: video c % V, first shot ((c / V) * frames) % (S - frames + 1)
  Dtype* ids = (*top)[1]->mutable_cpu_data();
  for (int b = 0; b < batch_size_ && from_records_; ++b, ++cursor_) {
    const long c = cursor_ % videos_;
    for (int f = 0; f < frames_; ++f) idx_host_[size_t(b) * frames_ + f] = int32_t(c * frames_ + f);
    ids[b] = Dtype(record_video_id_[size_t(c)]);
  }
  for (int b = 0; b < batch_size_ && !from_records_; ++b, ++cursor_) {
    const int v = int(cursor_ % videos_);
    const int start = int(((cursor_ / videos_) * frames_) % (shots_ - frames_ + 1));
    for (int f = 0; f < frames_; ++f) idx_host_[size_t(b) * frames_ + f] = v * shots_ + start + f;
    ids[b] = Dtype(v);
  }
  const size_t bytes = idx_host_.size() * sizeof(int32_t);
  int32_t* di = static_cast<int32_t*>(idx_dev_.get(bytes));
  CHECK_EQ(int(cudaMemcpyAsync(di, idx_host_.data(), bytes, cudaMemcpyHostToDevice, reinterpret_cast<cudaStream_t>(Caffe::stream()))), 0);
  VV_CHECK(vv_gather_rows(bank_ptr_, bank_rows_, feature_size_, di, nullptr, batch_size_, frames_, nullptr, nullptr, nullptr,
                          VV_PREC_FP32_SIMT, (*top)[0]->mutable_gpu_data(), Caffe::stream()));
  CHECK_EQ(int(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(Caffe::stream()))), 0);   // idx_host_ is reused by the next call
}

// ref: retrieval_stats_layer.cpp:20-66 (id,class text file), :71-84 Reshape, :143-359 Forward_cpu
template <typename Dtype>
void RetrievalStatsLayer<Dtype>::LayerSetUp(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const RetrievalStatsParameter p = this->layer_param_.retrieval_stats_param();
  video_level_ = p.video_level_retrieval();
  max_num_videos_ = p.max_num_videos();
  if (video_level_) CHECK_GE(max_num_videos_, 1) << "To do video level retrieval ... need min 1 video";
  stats_output_file_ = p.stats_output_file();
  exclude_same_video_shots_ = p.exclude_same_video_shots();
  std::ifstream f(p.id_to_class_file().c_str());
  CHECK(f.good()) << "cannot open id_to_class_file " << p.id_to_class_file();
  string line;
  while (std::getline(f, line)) {
    if (line.empty()) continue;
    const size_t comma = line.find(',');
    CHECK(comma != string::npos && line.find(',', comma + 1) == string::npos) << "id_to_class_file: expected 'video_id,class_id', got '" << line << "'";
    char* e1 = nullptr; char* e2 = nullptr;
    const long vid = strtol(line.c_str(), &e1, 10), cls = strtol(line.c_str() + comma + 1, &e2, 10);
    CHECK(e1 == line.c_str() + comma && *e2 == 0) << "id_to_class_file: bad line '" << line << "'";
    video_id_to_class_.insert(std::make_pair(int(vid), int(cls)));
  }
  CHECK_GE(int(video_id_to_class_.size()), 1) << "need atleast one entry in id-to-class map!";
}
template <typename Dtype>
void RetrievalStatsLayer<Dtype>::Reshape(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  CHECK_EQ(bottom[0]->num(), bottom[1]->num()) << "The data and label should have the same number.";
  CHECK_EQ(bottom[1]->channels(), 1); CHECK_EQ(bottom[1]->height(), 1); CHECK_EQ(bottom[1]->width(), 1);
  for (int i = 0; i < 3; ++i) (*top)[i]->Reshape(1, 1, 1, 1);      // mean AP, hit@1, hit@5
}
template <typename Dtype>
void RetrievalStatsLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  const int B = bottom[0]->num(), N = bottom[0]->count() / bottom[0]->num();
  const Dtype* vids = bottom[1]->cpu_data();
  cudaStream_t s = reinterpret_cast<cudaStream_t>(Caffe::stream());
  // the items that are ranked: the batch's shots, or -- video_level_retrieval (ref: :160-206) -- one mean embedding per
  // video (the reference enumerates the videos in its hash map's order; ascending id here: the statistics do not depend
  // on it, the CSV's line order does)
  int Q = B;
  vector<int32_t> ids(B), labels;
  for (int b = 0; b < B; ++b) ids[b] = static_cast<int>(vids[b]);
  const Dtype* E = bottom[0]->gpu_data();
  if (video_level_) {
    std::map<int, int> index;
    for (int b = 0; b < B; ++b) index.insert(std::make_pair(ids[b], 0));
    CHECK_EQ(int(index.size()), max_num_videos_);
    vector<int32_t> vid_of;
    for (auto& kv : index) { kv.second = int(vid_of.size()); vid_of.push_back(kv.first); }
    vector<int32_t> group(B);
    for (int b = 0; b < B; ++b) group[b] = index[ids[b]];
    Q = int(vid_of.size());
    int32_t* d_group = static_cast<int32_t*>(group_dev_.get(sizeof(int32_t) * B));
    CHECK_EQ(int(cudaMemcpyAsync(d_group, group.data(), sizeof(int32_t) * B, cudaMemcpyHostToDevice, s)), 0);
    float* mean = static_cast<float*>(mean_dev_.get(sizeof(float) * size_t(Q) * N));
    VV_CHECK(vv_video_mean_rows(E, B, N, d_group, Q, mean, Caffe::stream()));
    CHECK_EQ(int(cudaStreamSynchronize(s)), 0);          // `group` is a local
    E = mean;
    ids = vid_of;
  }
  CHECK_GT(Q, 1) << "retrieval statistics need at least two items";
  labels.resize(Q);
  for (int q = 0; q < Q; ++q) labels[q] = video_id_to_class_[ids[q]];   // operator[]: an unlisted video gets class 0, as in the reference (:249-253)
  int32_t* d_ids = static_cast<int32_t*>(ids_dev_.get(sizeof(int32_t) * Q));
  int32_t* d_lab = static_cast<int32_t*>(labels_dev_.get(sizeof(int32_t) * Q));
  CHECK_EQ(int(cudaMemcpyAsync(d_ids, ids.data(), sizeof(int32_t) * Q, cudaMemcpyHostToDevice, s)), 0);
  CHECK_EQ(int(cudaMemcpyAsync(d_lab, labels.data(), sizeof(int32_t) * Q, cudaMemcpyHostToDevice, s)), 0);
  const size_t ws = vv_retrieval_stats_workspace_bytes(Q);
  const bool csv = !stats_output_file_.empty();
  double* out = static_cast<double*>(out_dev_.get((3 + (csv ? 3 * size_t(Q) : 0)) * sizeof(double)));
  double* d_pq = csv ? out + 3 : nullptr;
  int32_t* d_top5 = csv ? static_cast<int32_t*>(top5_dev_.get(sizeof(int32_t) * 5 * Q)) : nullptr;
  VV_CHECK(vv_retrieval_stats_ex(E, Q, N, d_ids, d_lab, exclude_same_video_shots_ ? 1 : 0, nullptr, work_.get(ws), ws,
                                 out, d_pq, d_top5, Caffe::stream()));
  vector<double> host(3 + (csv ? 3 * size_t(Q) : 0));
  vector<int32_t> top5(csv ? 5 * size_t(Q) : 0);
  CHECK_EQ(int(cudaMemcpyAsync(host.data(), out, host.size() * sizeof(double), cudaMemcpyDeviceToHost, s)), 0);
  if (csv) CHECK_EQ(int(cudaMemcpyAsync(top5.data(), d_top5, top5.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, s)), 0);
  CHECK_EQ(int(cudaStreamSynchronize(s)), 0);
  for (int i = 0; i < 3; ++i) (*top)[i]->mutable_cpu_data()[0] = Dtype(host[i]);
  if (csv) {
    // the per-query CSV (ref: :146-151, 306-340): scored queries only, in item order; default ostream formatting
    std::ofstream f(stats_output_file_.c_str());
    f << "#video_id,class_id,ap,acc@1,acc@5" << ",ret_id_1,ret_id_2,ret_id_3,ret_id_4,ret_id_5"
      << ",class_id_1,class_id_2,class_id_3,class_id_4,class_id_5" << std::endl;
    int t5[5] = {0, 0, 0, 0, 0};              // the reference's vector outlives the loop: a missing entry keeps its last value
    for (int q = 0; q < Q; ++q) {
      if (labels[q] < 0) continue;
      const double ap = host[3 + 3 * q], a1 = host[3 + 3 * q + 1], a5 = host[3 + 3 * q + 2];
      f << ids[q] << "," << labels[q] << "," << ap << "," << a1 << "," << a5;
      if (!video_level_) {
        for (int k = 0; k < 5; ++k) if (top5[5 * size_t(q) + k] >= 0) t5[k] = top5[5 * size_t(q) + k];
        for (int k = 0; k < 5; ++k) f << "," << t5[k];
        for (int k = 0; k < 5; ++k) f << "," << video_id_to_class_[ids[t5[k]]];
      }
      f << std::endl;
    }
  }
}

// =================================== factory ===================================
template <typename Dtype>
Layer<Dtype>* GetLayer(const LayerParameter& param) {
  const string& name = param.name();
  switch (param.type()) {
    case LayerParameter_LayerType_CONCAT: return new ConcatLayer<Dtype>(param);
    case LayerParameter_LayerType_DROPOUT: return new DropoutLayer<Dtype>(param);
    case LayerParameter_LayerType_ELTWISE: return new EltwiseLayer<Dtype>(param);
    case LayerParameter_LayerType_FLATTEN: return new FlattenLayer<Dtype>(param);
    case LayerParameter_LayerType_INNER_PRODUCT: return new InnerProductLayer<Dtype>(param);
    case LayerParameter_LayerType_MAX_MARGIN_LOSS: return new MaxMarginLossLayer<Dtype>(param);
    case LayerParameter_LayerType_NORMALIZATION: return new NormalizationLayer<Dtype>(param);
    case LayerParameter_LayerType_RELU: return new ReLULayer<Dtype>(param);
    case LayerParameter_LayerType_SLICE: return new SliceLayer<Dtype>(param);
    case LayerParameter_LayerType_SPLIT: return new SplitLayer<Dtype>(param);
    case LayerParameter_LayerType_SUM: return new SumLayer<Dtype>(param);
    case LayerParameter_LayerType_VIDEO_SAMPLED_SHOTS_DATA: return new VideoSampledShotsDataLayer<Dtype>(param);
    case LayerParameter_LayerType_VIDEO_SHOT_WINDOW_TEST_DATA: return new VideoShotWindowTestDataLayer<Dtype>(param);
    case LayerParameter_LayerType_RETRIEVAL_STATS: return new RetrievalStatsLayer<Dtype>(param);
    case LayerParameter_LayerType_ID_TO_WEIGHT_MAPPING: return new IdToWeightMappingLayer<Dtype>(param);
    case LayerParameter_LayerType_NONE: LOG_FATAL << "Layer " << name << " has unspecified or unsupported type '" << param.m->str("type") << "'.";
    default: LOG_FATAL << "Layer " << name << " has unknown type " << param.type();
  }
  return nullptr;
}
template Layer<float>* GetLayer(const LayerParameter& param);
template class InnerProductLayer<float>; template class ReLULayer<float>; template class DropoutLayer<float>;
template class SliceLayer<float>; template class ConcatLayer<float>; template class FlattenLayer<float>;
template class SplitLayer<float>; template class EltwiseLayer<float>; template class NormalizationLayer<float>;
template class SumLayer<float>; template class MaxMarginLossLayer<float>; template class VideoSampledShotsDataLayer<float>;
template class VideoShotWindowTestDataLayer<float>; template class RetrievalStatsLayer<float>; template class IdToWeightMappingLayer<float>;

}  // namespace caffe
