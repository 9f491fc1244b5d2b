// dropin_gpu.cpp -- TEST INFRASTRUCTURE (oracle/): the drop-in proof.
//
// The reference's OWN headers and host sources -- include/caffe/*.hpp, blob.cpp, syncedmem.cpp, common.cpp, net.cpp,
// solver.cpp, insert_splits.cpp, the layers' .cpp files (LayerSetUp / Reshape / Forward_cpu) and its data layer -- are
// compiled from /root/reference in GPU mode (no -DCPU_ONLY).  Every device-side symbol they then need, i.e. what the
// reference's .cu files define (the layers' Forward_gpu / Backward_gpu, util/math_functions.cu's caffe_gpu_*), is defined
// HERE as a call sequence into libvv_b200.so's C-ABI (include/vv_b200.h) -- the bodies INTEGRATION.md asks a maintainer to
// write, against the real class declarations (members M_, K_, N_, slice_dim_, coeffs_, ... of common_layers.hpp /
// neuron_layers.hpp / data_layers.hpp).  The reference's unmodified Net::ForwardBackward, SGDSolver::ComputeUpdateValue and
// Net::Update then drive the B200 kernels; tests/test_gpu_dropin.py steps that build over the 8-iteration fixture the CPU
// build of the same sources produced (tests/golden/solver_ref.npz).
//
// cuBLAS / cuRAND are only ever created and destroyed by the reference's Caffe singleton on this path (common.cpp:85-100);
// the handles are stand-ins (nothing here calls the libraries).
#include <cuda_runtime_api.h>
#include <map>
#include <utility>
#include <vector>

#include "caffe/common.hpp"
#include "caffe/common_layers.hpp"
#include "caffe/data_layers.hpp"
#include "caffe/neuron_layers.hpp"
#include "caffe/util/math_functions.hpp"
#include "vv_b200.h"

#define VV(call) CHECK_EQ(int(call), 0) << vv_last_error()
#define CU(call) do { cudaError_t e_ = (call); CHECK_EQ(int(e_), 0) << cudaGetErrorString(e_); } while (0)

namespace {
const vv_stream_t kStream = nullptr;             // stock Caffe runs everything on the legacy default stream

int precision() {                                // VV_DROPIN_PREC=fp32_simt|tf32x3|tf32|bf16|f16x3 (default f16x3: tcgen05, fp32 parity)
  static const int p = [] {
    const char* e = getenv("VV_DROPIN_PREC");
    const std::string s(e ? e : "f16x3");
    if (s == "fp32_simt") return int(VV_PREC_FP32_SIMT);
    if (s == "tf32x3") return int(VV_PREC_TF32X3);
    if (s == "tf32") return int(VV_PREC_TF32);
    if (s == "bf16") return int(VV_PREC_BF16);
    return int(VV_PREC_F16X3);
  }();
  return p;
}
// operand staging buffers of a layer instance (the reference's class has no members for them)
struct Staging { void* p = nullptr; size_t bytes = 0; };
void* staging(const void* layer, int slot, size_t bytes) {
  static std::map<std::pair<const void*, int>, Staging> pool;
  Staging& s = pool[std::make_pair(layer, slot)];
  if (bytes > s.bytes) {
    if (s.p) { CU(cudaDeviceSynchronize()); cudaFree(s.p); }
    CU(cudaMalloc(&s.p, bytes));
    s.bytes = bytes;
  }
  return s.p;
}
// fp32 blob -> operand copies for the configured precision
vv_operand_t operand(const void* layer, int slot, const float* src, int64_t count, int prec) {
  vv_operand_t o; o.hi = src; o.lo = nullptr;
  if (prec == VV_PREC_FP32_SIMT || prec == VV_PREC_TF32) return o;
  size_t ho = 0, lo_off = 0;
  const size_t bytes = vv_operand_bytes(count, prec, &ho, &lo_off);
  char* base = static_cast<char*>(staging(layer, slot, bytes));
  VV(vv_prepare_operand(src, count, prec, base + ho, lo_off ? base + lo_off : nullptr, kStream));
  o.hi = base + ho; o.lo = lo_off ? base + lo_off : nullptr;
  return o;
}
bool tc_shape(int N, int K) { return (N % 8) == 0 && (K % 8) == 0; }
}  // namespace

// ---- the CUDA libraries the reference's singleton creates (never used on this path) --------------------------------
extern "C" {
__attribute__((visibility("hidden"))) cublasStatus_t cublasCreate_v2(cublasHandle_t* h) { *h = reinterpret_cast<cublasHandle_t>(new int(1)); return CUBLAS_STATUS_SUCCESS; }
__attribute__((visibility("hidden"))) cublasStatus_t cublasDestroy_v2(cublasHandle_t h) { delete reinterpret_cast<int*>(h); return CUBLAS_STATUS_SUCCESS; }
__attribute__((visibility("hidden"))) curandStatus_t curandCreateGenerator(curandGenerator_t* g, curandRngType_t) { *g = reinterpret_cast<curandGenerator_t>(new int(1)); return CURAND_STATUS_SUCCESS; }
__attribute__((visibility("hidden"))) curandStatus_t curandDestroyGenerator(curandGenerator_t g) { delete reinterpret_cast<int*>(g); return CURAND_STATUS_SUCCESS; }
__attribute__((visibility("hidden"))) curandStatus_t curandSetPseudoRandomGeneratorSeed(curandGenerator_t, unsigned long long) { return CURAND_STATUS_SUCCESS; }
__attribute__((visibility("hidden"))) curandStatus_t curandSetGeneratorOffset(curandGenerator_t, unsigned long long) { return CURAND_STATUS_SUCCESS; }
}

// ---- the data layer's rand() / srand() / std::random_shuffle stream, private to this library.  libc's generator is
// process-global state, and in a GPU process other libraries draw from it too (the sampled rows then differ from run to
// run); the product's bit-exact glibc generator (vv_glibc_rand_*, tested equal to libc's) serves this library's calls.
extern "C" {
static vv_glibc_rand_t* g_rand = nullptr;
__attribute__((visibility("hidden"))) int rand(void) { if (!g_rand) g_rand = vv_glibc_rand_create(1u); return vv_glibc_rand_next(g_rand); }
__attribute__((visibility("hidden"))) void srand(unsigned int seed) { if (g_rand) vv_glibc_rand_destroy(g_rand); g_rand = vv_glibc_rand_create(seed); }
}

namespace caffe {

// ---- util/math_functions.cu, the subset Blob / Layer / SGDSolver use (ref: math_functions.cu:15-143, 470-512) --------
void caffe_gpu_memcpy(const size_t N, const void* X, void* Y) { if (X != Y) CU(cudaMemcpy(Y, X, N, cudaMemcpyDefault)); }
template <> void caffe_gpu_axpy<float>(const int N, const float alpha, const float* X, float* Y) { VV(vv_axpby(N, alpha, X, 1.f, Y, kStream)); }
template <> void caffe_gpu_axpby<float>(const int N, const float alpha, const float* X, const float beta, float* Y) {
  // cuBLAS form: scal(beta, Y) then axpy(alpha, X, Y) -- two roundings, like the reference (math_functions.cu:128-138)
  VV(vv_axpby(N, beta, Y, 0.f, Y, kStream));
  VV(vv_axpby(N, alpha, X, 1.f, Y, kStream));
}
template <> void caffe_gpu_sign<float>(const int N, const float* x, float* y) {
  CU(cudaMemsetAsync(y, 0, sizeof(float) * N, nullptr));
  VV(vv_sign_axpy(N, 1.f, x, y, kStream));
}
template <> void caffe_gpu_add<float>(const int N, const float* a, const float* b, float* y) {
  if (y != a) VV(vv_axpby(N, 1.f, a, 0.f, y, kStream));
  VV(vv_axpby(N, 1.f, b, 1.f, y, kStream));
}
template <> void caffe_gpu_dot<float>(const int n, const float* x, const float* y, float* out) {
  // Layer::Forward's loss = dot(top.data, top.diff) over the (1-element) loss blobs (layer.hpp:427-435)
  CHECK_LE(n, 4096) << "caffe_gpu_dot: only the loss-weight dot products of the path are served";
  std::vector<float> hx(n), hy(n);
  CU(cudaMemcpy(hx.data(), x, sizeof(float) * n, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(hy.data(), y, sizeof(float) * n, cudaMemcpyDeviceToHost));
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += hx[i] * hy[i];
  *out = s;
}
template <> void caffe_gpu_asum<float>(const int, const float*, float*) { LOG(FATAL) << "caffe_gpu_asum (debug_info) is not on the path"; }
template <> void caffe_gpu_div<float>(const int, const float*, const float*, float*) { LOG(FATAL) << "caffe_gpu_div is not on the path"; }
template <> void caffe_gpu_powx<float>(const int, const float*, const float, float*) { LOG(FATAL) << "caffe_gpu_powx is not on the path"; }
template <> void caffe_gpu_add_scalar<float>(const int, const float, float*) { LOG(FATAL) << "caffe_gpu_add_scalar is not on the path"; }
#define NOT_BUILT_DOUBLE(sig) template <> sig { LOG(FATAL) << "the B200 path is built for float (tools/caffe.cpp:107 runs float only)"; }
NOT_BUILT_DOUBLE(void caffe_gpu_axpy<double>(const int, const double, const double*, double*))
NOT_BUILT_DOUBLE(void caffe_gpu_axpby<double>(const int, const double, const double*, const double, double*))
NOT_BUILT_DOUBLE(void caffe_gpu_sign<double>(const int, const double*, double*))
NOT_BUILT_DOUBLE(void caffe_gpu_add<double>(const int, const double*, const double*, double*))
NOT_BUILT_DOUBLE(void caffe_gpu_dot<double>(const int, const double*, const double*, double*))
NOT_BUILT_DOUBLE(void caffe_gpu_asum<double>(const int, const double*, double*))
NOT_BUILT_DOUBLE(void caffe_gpu_div<double>(const int, const double*, const double*, double*))
NOT_BUILT_DOUBLE(void caffe_gpu_powx<double>(const int, const double*, const double, double*))
NOT_BUILT_DOUBLE(void caffe_gpu_add_scalar<double>(const int, const double, double*))

// ---- the layers' device bodies.  One generic template per member (float does the work; the double instantiation the
// reference's INSTANTIATE_CLASS expects aborts).
template <typename Dtype> struct F { static const bool ok = false; };
template <> struct F<float> { static const bool ok = true; };
#define FLOAT_ONLY() do { if (!F<Dtype>::ok) LOG(FATAL) << "the B200 path is built for float"; } while (0)
#define FP(x) reinterpret_cast<const float*>(x)
#define FPM(x) reinterpret_cast<float*>(x)

// D: base_data_layer.cu:7-21 -- join the prefetch thread, hand the batch to the device, restart the thread
template <typename Dtype>
void BasePrefetchingDataLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  JoinPrefetchThread();
  CU(cudaMemcpy((*top)[0]->mutable_gpu_data(), prefetch_data_.cpu_data(), sizeof(Dtype) * prefetch_data_.count(), cudaMemcpyHostToDevice));
  if (this->output_labels_)
    CU(cudaMemcpy((*top)[1]->mutable_gpu_data(), prefetch_label_.cpu_data(), sizeof(Dtype) * prefetch_label_.count(), cudaMemcpyHostToDevice));
  CreatePrefetchThread();
}

// P1 / P1': inner_product_layer.cu:12-59
template <typename Dtype>
void InnerProductLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  FLOAT_ONLY();
  const int prec = tc_shape(N_, K_) ? precision() : int(VV_PREC_FP32_SIMT);
  const vv_operand_t X = operand(this, 0, FP(bottom[0]->gpu_data()), int64_t(M_) * K_, prec);
  const vv_operand_t W = operand(this, 1, FP(this->blobs_[0]->gpu_data()), int64_t(N_) * K_, prec);
  VV(vv_ip_forward(X, W, bias_term_ ? FP(this->blobs_[1]->gpu_data()) : nullptr, M_, N_, K_, prec, nullptr, nullptr,
                   FPM((*top)[0]->mutable_gpu_data()), kStream));
}
template <typename Dtype>
void InnerProductLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down,
                                            vector<Blob<Dtype>*>* bottom) {
  FLOAT_ONLY();
  const int prec = tc_shape(N_, K_) ? precision() : int(VV_PREC_FP32_SIMT);
  const vv_operand_t dZ = operand(this, 2, FP(top[0]->gpu_diff()), int64_t(M_) * N_, prec);
  if (this->param_propagate_down_[0]) {
    const vv_operand_t X = operand(this, 0, FP((*bottom)[0]->gpu_data()), int64_t(M_) * K_, prec);
    const size_t ws = vv_ip_wgrad_workspace_bytes(M_, N_, K_, prec);
    VV(vv_ip_wgrad(dZ, X, M_, N_, K_, prec, this->layer_param_.inner_product_param().regularization(),
                   FPM(this->blobs_[0]->mutable_gpu_diff()), 0, ws ? staging(this, 3, ws) : nullptr, ws, kStream));
  }
  if (bias_term_ && this->param_propagate_down_[1])
    VV(vv_ip_bias_grad(FP(top[0]->gpu_diff()), M_, N_, FPM(this->blobs_[1]->mutable_gpu_diff()), kStream));
  if (propagate_down[0]) {
    const vv_operand_t W = operand(this, 1, FP(this->blobs_[0]->gpu_data()), int64_t(N_) * K_, prec);
    VV(vv_ip_dgrad(dZ, W, M_, N_, K_, prec, FPM((*bottom)[0]->mutable_gpu_diff()), kStream));
  }
}

// P2 / P2': relu_layer.cu:9-60 (the backward gates on the bottom DATA)
template <typename Dtype>
void ReLULayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  FLOAT_ONLY();
  VV(vv_relu_forward(FP(bottom[0]->gpu_data()), bottom[0]->count(), this->layer_param_.relu_param().negative_slope(),
                     FPM((*top)[0]->mutable_gpu_data()), kStream));
}
template <typename Dtype>
void ReLULayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  FLOAT_ONLY();
  if (!propagate_down[0]) return;
  VV(vv_relu_backward(FP((*bottom)[0]->gpu_data()), FP(top[0]->gpu_diff()), (*bottom)[0]->count(),
                      this->layer_param_.relu_param().negative_slope(), FPM((*bottom)[0]->mutable_gpu_diff()), kStream));
}

// P3 / P3': dropout_layer.cu:14-70.  The mask blob holds 0/1 words drawn on the device (Philox, one sub-stream per call)
template <typename Dtype>
void DropoutLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  FLOAT_ONLY();
  const int count = bottom[0]->count();
  if (Caffe::phase() == Caffe::TRAIN) {
    static uint64_t draw = 0;
    unsigned int* mask = static_cast<unsigned int*>(rand_vec_.mutable_gpu_data());
    const int rows = bottom[0]->num(), cols = count / rows;
    VV(vv_dropout_make_mask(mask, rows, cols, threshold_, 1701, ++draw, kStream));
    VV(vv_dropout_forward(FP(bottom[0]->gpu_data()), mask, VV_DROPOUT_MASK01, count, threshold_, FPM((*top)[0]->mutable_gpu_data()), kStream));
  } else {
    caffe_gpu_memcpy(sizeof(Dtype) * count, bottom[0]->gpu_data(), (*top)[0]->mutable_gpu_data());
  }
}
template <typename Dtype>
void DropoutLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  FLOAT_ONLY();
  if (!propagate_down[0]) return;
  const int count = (*bottom)[0]->count();
  if (Caffe::phase() == Caffe::TRAIN) {
    VV(vv_dropout_backward(FP(top[0]->gpu_diff()), static_cast<const unsigned int*>(rand_vec_.gpu_data()), VV_DROPOUT_MASK01, count,
                           threshold_, FPM((*bottom)[0]->mutable_gpu_diff()), kStream));
  } else {
    caffe_gpu_memcpy(sizeof(Dtype) * count, top[0]->gpu_diff(), (*bottom)[0]->mutable_gpu_diff());
  }
}

// G0 / G1: slice_layer.cu, concat_layer.cu -- one strided copy per blob instead of one cudaMemcpy per sample
template <typename Dtype>
void SliceLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  FLOAT_ONLY();
  // (the reference's Reshape leaves num_ / channels_ holding the PER-TOP extent when no slice points are given,
  // slice_layer.cpp:59-71: the bottom's own shape gives the source pitch)
  const float* src = FP(bottom[0]->gpu_data());
  const int64_t inner = int64_t(bottom[0]->height()) * bottom[0]->width();
  const int64_t src_pitch = bottom[0]->channels() * inner;
  const int rows = bottom[0]->num();
  int64_t off = 0;
  for (size_t i = 0; i < top->size(); ++i) {
    Blob<Dtype>* t = (*top)[i];
    if (slice_dim_ == 0) {
      VV(vv_copy_strided(src + off, t->count(), FPM(t->mutable_gpu_data()), t->count(), 1, t->count(), kStream));
      off += t->count();
    } else {
      const int64_t cols = t->channels() * inner;
      VV(vv_copy_strided(src + off, src_pitch, FPM(t->mutable_gpu_data()), cols, rows, cols, kStream));
      off += cols;
    }
  }
}
template <typename Dtype>
void SliceLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  FLOAT_ONLY();
  if (!propagate_down[0]) return;
  float* dst = FPM((*bottom)[0]->mutable_gpu_diff());
  const int64_t inner = int64_t((*bottom)[0]->height()) * (*bottom)[0]->width();
  const int64_t dst_pitch = (*bottom)[0]->channels() * inner;
  const int rows = (*bottom)[0]->num();
  int64_t off = 0;
  for (size_t i = 0; i < top.size(); ++i) {
    Blob<Dtype>* t = top[i];
    if (slice_dim_ == 0) {
      VV(vv_copy_strided(FP(t->gpu_diff()), t->count(), dst + off, t->count(), 1, t->count(), kStream));
      off += t->count();
    } else {
      const int64_t cols = t->channels() * inner;
      VV(vv_copy_strided(FP(t->gpu_diff()), cols, dst + off, dst_pitch, rows, cols, kStream));
      off += cols;
    }
  }
}
template <typename Dtype>
void ConcatLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  FLOAT_ONLY();
  float* dst = FPM((*top)[0]->mutable_gpu_data());
  const int64_t inner = int64_t(height_) * width_;
  int64_t off = 0;
  for (size_t i = 0; i < bottom.size(); ++i) {
    if (concat_dim_ == 0) {
      VV(vv_copy_strided(FP(bottom[i]->gpu_data()), bottom[i]->count(), dst + off, bottom[i]->count(), 1, bottom[i]->count(), kStream));
      off += bottom[i]->count();
    } else {
      const int64_t cols = bottom[i]->channels() * inner;
      VV(vv_copy_strided(FP(bottom[i]->gpu_data()), cols, dst + off, channels_ * inner, num_, cols, kStream));
      off += cols;
    }
  }
}
template <typename Dtype>
void ConcatLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  FLOAT_ONLY();
  const float* src = FP(top[0]->gpu_diff());
  const int64_t inner = int64_t(height_) * width_;
  int64_t off = 0;
  for (size_t i = 0; i < bottom->size(); ++i) {
    Blob<Dtype>* b = (*bottom)[i];
    const int64_t cols = concat_dim_ == 0 ? b->count() : b->channels() * inner;
    if (propagate_down[i]) {
      if (concat_dim_ == 0) VV(vv_copy_strided(src + off, cols, FPM(b->mutable_gpu_diff()), cols, 1, cols, kStream));
      else VV(vv_copy_strided(src + off, channels_ * inner, FPM(b->mutable_gpu_diff()), cols, num_, cols, kStream));
    }
    off += cols;
  }
}
// flatten_layer.cu: the top shares the bottom's memory in both directions
template <typename Dtype>
void FlattenLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) { (*top)[0]->ShareData(*bottom[0]); }
template <typename Dtype>
void FlattenLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  (*bottom)[0]->ShareDiff(*top[0]);
}
// C3: split_layer.cu:17-33 -- d = top0 + top1 ; d += top_k, the same association order
template <typename Dtype>
void SplitLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  for (size_t i = 0; i < top->size(); ++i) (*top)[i]->ShareData(*bottom[0]);
}
template <typename Dtype>
void SplitLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  FLOAT_ONLY();
  if (!propagate_down[0]) return;
  float* d = FPM((*bottom)[0]->mutable_gpu_diff());
  if (top.size() == 1) { caffe_gpu_memcpy(sizeof(Dtype) * count_, top[0]->gpu_diff(), d); return; }
  VV(vv_axpby(count_, 1.f, FP(top[0]->gpu_diff()), 0.f, d, kStream));
  for (size_t i = 1; i < top.size(); ++i) VV(vv_axpby(count_, 1.f, FP(top[i]->gpu_diff()), 1.f, d, kStream));
}

// C1 / C4: eltwise_layer.cu:41-54, 96-119
template <typename Dtype>
void EltwiseLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  FLOAT_ONLY();
  const int count = (*top)[0]->count();
  float* t = FPM((*top)[0]->mutable_gpu_data());
  if (op_ == EltwiseParameter_EltwiseOp_PROD) {
    VV(vv_eltwise_prod_forward(FP(bottom[0]->gpu_data()), FP(bottom[1]->gpu_data()), count, t, kStream));
    for (size_t i = 2; i < bottom.size(); ++i) VV(vv_mul(count, t, FP(bottom[i]->gpu_data()), t, kStream));
  } else if (op_ == EltwiseParameter_EltwiseOp_SUM) {
    std::vector<const float*> ptrs(bottom.size());
    std::vector<float> co(bottom.size());
    for (size_t i = 0; i < bottom.size(); ++i) { ptrs[i] = FP(bottom[i]->gpu_data()); co[i] = float(coeffs_[i]); }
    VV(vv_eltwise_sum_forward(ptrs.data(), co.data(), int(bottom.size()), count, t, kStream));
  } else {
    LOG(FATAL) << "Eltwise MAX is not on the temporal-embedding path";
  }
}
template <typename Dtype>
void EltwiseLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  FLOAT_ONLY();
  const int count = top[0]->count();
  const float* top_diff = FP(top[0]->gpu_diff());
  for (size_t i = 0; i < bottom->size(); ++i) {
    if (!propagate_down[i]) continue;
    float* bd = FPM((*bottom)[i]->mutable_gpu_diff());
    if (op_ == EltwiseParameter_EltwiseOp_PROD) {
      CHECK(stable_prod_grad_) << "the unstable PROD gradient is not built";
      bool initialized = false;
      for (size_t j = 0; j < bottom->size(); ++j) {
        if (i == j) continue;
        if (!initialized) { VV(vv_axpby(count, 1.f, FP((*bottom)[j]->gpu_data()), 0.f, bd, kStream)); initialized = true; }
        else VV(vv_mul(count, FP((*bottom)[j]->gpu_data()), bd, bd, kStream));
      }
      VV(vv_mul(count, bd, top_diff, bd, kStream));
    } else {
      VV(vv_axpby(count, float(coeffs_[i]), top_diff, 0.f, bd, kStream));
    }
  }
}

// C2 / N': normalization_layer.cu:10-97 -- one warp per row instead of 6 / 12 launches
template <typename Dtype>
void NormalizationLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  FLOAT_ONLY();
  const int num = bottom[0]->num();
  VV(vv_l2norm_forward(FP(bottom[0]->gpu_data()), num, bottom[0]->count() / num, FPM((*top)[0]->mutable_gpu_data()), kStream));
}
template <typename Dtype>
void NormalizationLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  FLOAT_ONLY();
  const int num = (*bottom)[0]->num();
  VV(vv_l2norm_backward(FP((*bottom)[0]->gpu_data()), FP(top[0]->gpu_diff()), num, (*bottom)[0]->count() / num,
                        FPM((*bottom)[0]->mutable_gpu_diff()), kStream));
}
// C4: sum_layer.cu:10-55
template <typename Dtype>
void SumLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  FLOAT_ONLY();
  const int num = bottom[0]->num();
  VV(vv_rowsum_forward(FP(bottom[0]->gpu_data()), num, bottom[0]->count() / num, num_output_, FPM((*top)[0]->mutable_gpu_data()), kStream));
}
template <typename Dtype>
void SumLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  FLOAT_ONLY();
  const int num = (*bottom)[0]->num();
  VV(vv_rowsum_backward(FP(top[0]->gpu_diff()), num, (*bottom)[0]->count() / num, num_output_, FPM((*bottom)[0]->mutable_gpu_diff()), kStream));
}

// f-4: id_to_weight_mapping_layer.cu -- the per-id embedding table: row gather forward, deterministic scatter-add backward
// (K_ = table rows, N_ = embedding width, M_ = ids in the batch)
template <typename Dtype>
void IdToWeightMappingLayer<Dtype>::Forward_gpu(const vector<Blob<Dtype>*>& bottom, vector<Blob<Dtype>*>* top) {
  FLOAT_ONLY();
  VV(vv_id_lookup_forward(FP(this->blobs_[0]->gpu_data()), this->blobs_[0]->num(), N_, FP(bottom[0]->gpu_data()), M_,
                          FPM((*top)[0]->mutable_gpu_data()), kStream));
}
template <typename Dtype>
void IdToWeightMappingLayer<Dtype>::Backward_gpu(const vector<Blob<Dtype>*>& top, const vector<bool>& propagate_down, vector<Blob<Dtype>*>* bottom) {
  FLOAT_ONLY();
  if (!this->param_propagate_down_[0]) return;
  VV(vv_id_lookup_backward(FP(top[0]->gpu_diff()), FP((*bottom)[0]->gpu_data()), M_, N_, this->blobs_[0]->num(),
                           FPM(this->blobs_[0]->mutable_gpu_diff()), kStream));
}

#define INST_FB(cls)                                                                                                     \
  template void cls<float>::Forward_gpu(const vector<Blob<float>*>&, vector<Blob<float>*>*);                            \
  template void cls<double>::Forward_gpu(const vector<Blob<double>*>&, vector<Blob<double>*>*);                         \
  template void cls<float>::Backward_gpu(const vector<Blob<float>*>&, const vector<bool>&, vector<Blob<float>*>*);      \
  template void cls<double>::Backward_gpu(const vector<Blob<double>*>&, const vector<bool>&, vector<Blob<double>*>*)
INST_FB(InnerProductLayer); INST_FB(ReLULayer); INST_FB(DropoutLayer); INST_FB(SliceLayer); INST_FB(ConcatLayer);
INST_FB(FlattenLayer); INST_FB(SplitLayer); INST_FB(EltwiseLayer); INST_FB(NormalizationLayer); INST_FB(SumLayer);
INST_FB(IdToWeightMappingLayer);
template void BasePrefetchingDataLayer<float>::Forward_gpu(const vector<Blob<float>*>&, vector<Blob<float>*>*);
template void BasePrefetchingDataLayer<double>::Forward_gpu(const vector<Blob<double>*>&, vector<Blob<double>*>*);

}  // namespace caffe
