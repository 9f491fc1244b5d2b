// SyncedMemory / Blob / Caffe singleton (ref: src/caffe/syncedmem.cpp:21-109, src/caffe/blob.cpp, src/caffe/common.cpp)
#include <cuda_runtime_api.h>
#include <cstring>
#include "caffe/blob.hpp"

namespace caffe {

#define CUDA_CHECK(call) do { cudaError_t _e = (call); CHECK_EQ(int(_e), 0) << cudaGetErrorString(_e); } while (0)

Caffe::Caffe() : mode_(GPU), phase_(TRAIN), stream_(nullptr), seed_(1701), draws_(0), prec_(VV_PREC_F16X3) {
  if (const char* e = getenv("VV_PRECISION")) {
    const string v(e);
    if (v == "fp32_simt") prec_ = VV_PREC_FP32_SIMT; else if (v == "tf32x3") prec_ = VV_PREC_TF32X3;
    else if (v == "tf32") prec_ = VV_PREC_TF32; else if (v == "bf16") prec_ = VV_PREC_BF16;
    else if (v == "f16x3") prec_ = VV_PREC_F16X3;
    else LOG_FATAL << "VV_PRECISION must be fp32_simt|tf32x3|f16x3|tf32|bf16, got " << v;
  }
}
}  // namespace caffe
#include "caffe/proto/caffe_params.hpp"
namespace caffe {
template <typename Dtype>
void Blob<Dtype>::FromProto(const PbMsg& proto) {
  Reshape(int(proto.num("num", 0)), int(proto.num("channels", 0)), int(proto.num("height", 0)), int(proto.num("width", 0)));
  const PbField* d = proto.nth("data", 0);
  CHECK(d && d->floats && int(d->floats->size()) == count_) << "BlobProto data has " << (d && d->floats ? d->floats->size() : 0) << " values, shape needs " << count_;
  memcpy(mutable_cpu_data(), d->floats->data(), sizeof(Dtype) * count_);
  const PbField* g = proto.nth("diff", 0);
  if (g && g->floats && !g->floats->empty()) {
    CHECK_EQ(int(g->floats->size()), count_);
    memcpy(mutable_cpu_diff(), g->floats->data(), sizeof(Dtype) * count_);
  }
}
template <typename Dtype>
void Blob<Dtype>::ToProto(PbMsg* proto, bool write_diff) const {
  proto->fields.clear();
  proto->add_scalar("num", std::to_string(num_)); proto->add_scalar("channels", std::to_string(channels_));
  proto->add_scalar("height", std::to_string(height_)); proto->add_scalar("width", std::to_string(width_));
  auto data = std::make_shared<vector<float> >(cpu_data(), cpu_data() + count_);
  proto->fields.push_back(PbField{"data", "", nullptr, data});
  if (write_diff) {
    auto diff = std::make_shared<vector<float> >(cpu_diff(), cpu_diff() + count_);
    proto->fields.push_back(PbField{"diff", "", nullptr, diff});
  }
}

Caffe& Caffe::Get() { static Caffe c; return c; }
void Caffe::SetDevice(int device_id) { CUDA_CHECK(cudaSetDevice(device_id)); VV_CHECK(vv_device_check()); }

SyncedMemory::~SyncedMemory() {
  if (cpu_ptr_) free(cpu_ptr_);
  if (gpu_ptr_ && own_gpu_) cudaFree(gpu_ptr_);
}
void SyncedMemory::to_cpu() {
  switch (head_) {
    case UNINITIALIZED:
      cpu_ptr_ = malloc(size_); CHECK(cpu_ptr_) << "host allocation of " << size_ << " bytes failed";
      memset(cpu_ptr_, 0, size_); head_ = HEAD_AT_CPU; break;
    case HEAD_AT_GPU:
      if (!cpu_ptr_) { cpu_ptr_ = malloc(size_); CHECK(cpu_ptr_); }
      CUDA_CHECK(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(Caffe::stream())));
      CUDA_CHECK(cudaMemcpy(cpu_ptr_, gpu_ptr_, size_, cudaMemcpyDeviceToHost));
      head_ = SYNCED; break;
    default: break;
  }
}
void SyncedMemory::to_gpu() {
  switch (head_) {
    case UNINITIALIZED:
      CUDA_CHECK(cudaMalloc(&gpu_ptr_, size_)); CUDA_CHECK(cudaMemset(gpu_ptr_, 0, size_)); head_ = HEAD_AT_GPU; break;
    case HEAD_AT_CPU:
      if (!gpu_ptr_) CUDA_CHECK(cudaMalloc(&gpu_ptr_, size_));
      CUDA_CHECK(cudaMemcpy(gpu_ptr_, cpu_ptr_, size_, cudaMemcpyHostToDevice));
      head_ = SYNCED; break;
    default: break;
  }
}
const void* SyncedMemory::cpu_data() { to_cpu(); return cpu_ptr_; }
const void* SyncedMemory::gpu_data() { to_gpu(); return gpu_ptr_; }
void* SyncedMemory::mutable_cpu_data() { to_cpu(); head_ = HEAD_AT_CPU; return cpu_ptr_; }
void* SyncedMemory::mutable_gpu_data() { to_gpu(); head_ = HEAD_AT_GPU; return gpu_ptr_; }
void SyncedMemory::set_gpu_data(void* data) {
  CHECK(data);
  if (gpu_ptr_ && own_gpu_) cudaFree(gpu_ptr_);
  gpu_ptr_ = data; own_gpu_ = false; head_ = HEAD_AT_GPU;
}

template <typename Dtype>
void Blob<Dtype>::Reshape(const int num, const int channels, const int height, const int width) {
  CHECK_GE(num, 0); CHECK_GE(channels, 0); CHECK_GE(height, 0); CHECK_GE(width, 0);
  num_ = num; channels_ = channels; height_ = height; width_ = width;
  count_ = num_ * channels_ * height_ * width_;
  if (count_ > capacity_) {
    capacity_ = count_;
    data_.reset(new SyncedMemory(capacity_ * sizeof(Dtype)));
    diff_.reset(new SyncedMemory(capacity_ * sizeof(Dtype)));
  }
}
template <typename Dtype> const Dtype* Blob<Dtype>::cpu_data() const { CHECK(data_); return (const Dtype*)data_->cpu_data(); }
template <typename Dtype> const Dtype* Blob<Dtype>::gpu_data() const { CHECK(data_); return (const Dtype*)data_->gpu_data(); }
template <typename Dtype> const Dtype* Blob<Dtype>::cpu_diff() const { CHECK(diff_); return (const Dtype*)diff_->cpu_data(); }
template <typename Dtype> const Dtype* Blob<Dtype>::gpu_diff() const { CHECK(diff_); return (const Dtype*)diff_->gpu_data(); }
template <typename Dtype> Dtype* Blob<Dtype>::mutable_cpu_data() { CHECK(data_); return (Dtype*)data_->mutable_cpu_data(); }
template <typename Dtype> Dtype* Blob<Dtype>::mutable_gpu_data() { CHECK(data_); return (Dtype*)data_->mutable_gpu_data(); }
template <typename Dtype> Dtype* Blob<Dtype>::mutable_cpu_diff() { CHECK(diff_); return (Dtype*)diff_->mutable_cpu_data(); }
template <typename Dtype> Dtype* Blob<Dtype>::mutable_gpu_diff() { CHECK(diff_); return (Dtype*)diff_->mutable_gpu_data(); }

template <> void Blob<float>::Update() {
  // data -= diff on the device wherever the head is (no CPU arithmetic in this build)
  VV_CHECK(vv_axpby(count_, -1.f, (const float*)diff_->gpu_data(), 1.f, (float*)data_->mutable_gpu_data(), Caffe::stream()));
}
// debug_info sums (net.cpp's ForwardDebugInfo): read back and summed on the host, not part of any step
static float abs_sum(const float* p, int n) {
  double s = 0;
  for (int i = 0; i < n; ++i) s += p[i] < 0 ? -p[i] : p[i];
  return float(s);
}
template <> float Blob<float>::asum_data() const { return data_ ? abs_sum(cpu_data(), count_) : 0.f; }
template <> float Blob<float>::asum_diff() const { return diff_ ? abs_sum(cpu_diff(), count_) : 0.f; }
template <typename Dtype>
void Blob<Dtype>::CopyFrom(const Blob<Dtype>& source, bool copy_diff, bool reshape) {
  if (num_ != source.num() || channels_ != source.channels() || height_ != source.height() || width_ != source.width()) {
    if (reshape) Reshape(source.num(), source.channels(), source.height(), source.width());
    else LOG_FATAL << "Trying to copy blobs of different sizes.";
  }
  const void* src = copy_diff ? (const void*)source.gpu_diff() : (const void*)source.gpu_data();
  void* dst = copy_diff ? (void*)mutable_gpu_diff() : (void*)mutable_gpu_data();
  CUDA_CHECK(cudaMemcpyAsync(dst, src, sizeof(Dtype) * count_, cudaMemcpyDeviceToDevice, reinterpret_cast<cudaStream_t>(Caffe::stream())));
}
template class Blob<float>;

}  // namespace caffe
