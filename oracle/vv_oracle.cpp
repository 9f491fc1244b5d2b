// =============================================================================
// vv_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain C++ restatement of the reference's (eevignesh/videovector, a 2014
// Caffe fork) CPU algorithm for the temporal-context embedding training path.
// Nothing in the product (videovector_b200/, include/, host/) links, loads or
// calls this file.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it, and only as the checker or
// the reported CPU baseline.
//
// Every function cites the reference file:line (relative to /root/reference)
// whose arithmetic and operation ORDER it follows.  Floating point is IEEE
// fp32 like the reference's `float` instantiation (tools/caffe.cpp:107).
//
// Parity pinning (see DESIGN.md "Oracle"):
//   * BLAS conventions: reference GemmTest/GemvTest known answers
//     (src/caffe/test/test_util_blas.cpp:22-130) -> tests/test_oracle_pins.py
//   * layer semantics: the reference's own per-layer test assertions
//      re-expressed in tests/test_oracle_pins.py
//   * optional stronger pin: oracle/_ref (reference layer sources compiled
//     unmodified against shim headers, see oracle/ref_shim/)
//   * sampler: the reference has NO test for it -> "parity unpinned by
//     reference tests"; pinned by construction (real glibc rand(), real
//     std::random_shuffle, same call order).
//
// Third-party arithmetic the reference delegates and that is NOT under
// /root/reference: CPU BLAS (OpenBLAS, version unpinned, Makefile.config:34).
// Here: the OpenBLAS inside the SciPy wheel when orc_set_blas() is given its
// path (dlopen, symbols scipy_cblas_*), else the built-in loops below.  BLAS
// summation order is unspecified, so float comparisons are tolerance based.
// =============================================================================
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <dlfcn.h>
#include <numeric>
#include <set>
#include <string>
#include <vector>

namespace {

// ----------------------------------------------------------------------------
// BLAS layer (reference: src/caffe/util/math_functions.cpp:12-54,136-156)
// ----------------------------------------------------------------------------
enum { kRowMajor = 101, kNoTrans = 111, kTrans = 112 };
typedef void (*sgemm_fn)(int, int, int, int, int, int, float, const float*, int,
                         const float*, int, float, float*, int);
typedef void (*sgemv_fn)(int, int, int, int, float, const float*, int,
                         const float*, int, float, float*, int);
typedef void (*set_threads_fn)(int);
typedef int (*get_threads_fn)(void);

void* g_blas = nullptr;
sgemm_fn g_sgemm = nullptr;
sgemv_fn g_sgemv = nullptr;
get_threads_fn g_get_threads = nullptr;
int g_threads = 1;

// C[M,N] = alpha*op(A)*op(B) + beta*C, row-major, lda/ldb chosen exactly as
// caffe_cpu_gemm does (math_functions.cpp:12-21): lda = (TransA==NoTrans)?K:M,
// ldb = (TransB==NoTrans)?N:K.
void builtin_gemm(bool tA, bool tB, int M, int N, int K, float alpha,
                  const float* A, const float* B, float beta, float* C) {
  const int lda = tA ? M : K, ldb = tB ? K : N;
  for (int i = 0; i < M; ++i) {
    for (int j = 0; j < N; ++j) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) {
        const float a = tA ? A[(size_t)k * lda + i] : A[(size_t)i * lda + k];
        const float b = tB ? B[(size_t)j * ldb + k] : B[(size_t)k * ldb + j];
        acc += a * b;
      }
      float& c = C[(size_t)i * N + j];
      c = (beta == 0.f) ? alpha * acc : alpha * acc + beta * c;
    }
  }
}

void cpu_gemm(bool tA, bool tB, int M, int N, int K, float alpha,
              const float* A, const float* B, float beta, float* C) {
  if (g_sgemm) {
    const int lda = tA ? M : K, ldb = tB ? K : N;
    g_sgemm(kRowMajor, tA ? kTrans : kNoTrans, tB ? kTrans : kNoTrans, M, N, K,
            alpha, A, lda, B, ldb, beta, C, N);
  } else {
    builtin_gemm(tA, tB, M, N, K, alpha, A, B, beta, C);
  }
}

// y = alpha*op(A)*x + beta*y, A is M x N row-major (math_functions.cpp:36-41).
void cpu_gemv(bool tA, int M, int N, float alpha, const float* A,
              const float* x, float beta, float* y) {
  if (g_sgemv) {
    g_sgemv(kRowMajor, tA ? kTrans : kNoTrans, M, N, alpha, A, N, x, 1, beta, y, 1);
    return;
  }
  if (!tA) {
    for (int i = 0; i < M; ++i) {
      float acc = 0.f;
      for (int j = 0; j < N; ++j) acc += A[(size_t)i * N + j] * x[j];
      y[i] = (beta == 0.f) ? alpha * acc : alpha * acc + beta * y[i];
    }
  } else {
    for (int j = 0; j < N; ++j) {
      float acc = 0.f;
      for (int i = 0; i < M; ++i) acc += A[(size_t)i * N + j] * x[i];
      y[j] = (beta == 0.f) ? alpha * acc : alpha * acc + beta * y[j];
    }
  }
}

// Level-1 pieces are plain loops: same arithmetic as saxpy/sscal/sdot up to
// the (unspecified) summation order of sdot.
inline void cpu_axpy(size_t n, float a, const float* x, float* y) {
  for (size_t i = 0; i < n; ++i) y[i] += a * x[i];
}
inline void cpu_scal(size_t n, float a, float* x) {
  for (size_t i = 0; i < n; ++i) x[i] *= a;
}
inline float cpu_dot(size_t n, const float* x, const float* y) {
  float acc = 0.f;
  for (size_t i = 0; i < n; ++i) acc += x[i] * y[i];
  return acc;
}
inline float cpu_asum(size_t n, const float* x) {
  float acc = 0.f;
  for (size_t i = 0; i < n; ++i) acc += std::fabs(x[i]);
  return acc;
}
// cblas_saxpby = sscal(beta) then saxpy(alpha) (util/mkl_alternate.hpp:82-87)
inline void cpu_axpby(size_t n, float alpha, const float* x, float beta, float* y) {
  cpu_scal(n, beta, y);
  cpu_axpy(n, alpha, x, y);
}
// vsPowx: y = pow(a, b) evaluated in float (util/mkl_alternate.hpp:54)
inline void cpu_powx(size_t n, const float* a, float b, float* y) {
  for (size_t i = 0; i < n; ++i) y[i] = std::pow(a[i], b);
}

}  // namespace

extern "C" {

// ----------------------------------------------------------------------------
// BLAS selection
// ----------------------------------------------------------------------------
int orc_set_blas(const char* path, int threads) {
  if (!path || !*path) { g_sgemm = nullptr; g_sgemv = nullptr; g_threads = 1; return 0; }
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return -1;
  sgemm_fn f = (sgemm_fn)dlsym(h, "scipy_cblas_sgemm");
  sgemv_fn v = (sgemv_fn)dlsym(h, "scipy_cblas_sgemv");
  if (!f) f = (sgemm_fn)dlsym(h, "cblas_sgemm");
  if (!v) v = (sgemv_fn)dlsym(h, "cblas_sgemv");
  if (!f || !v) return -2;
  set_threads_fn st = (set_threads_fn)dlsym(h, "scipy_openblas_set_num_threads");
  if (!st) st = (set_threads_fn)dlsym(h, "openblas_set_num_threads");
  g_get_threads = (get_threads_fn)dlsym(h, "scipy_openblas_get_num_threads");
  if (!g_get_threads) g_get_threads = (get_threads_fn)dlsym(h, "openblas_get_num_threads");
  if (st && threads > 0) st(threads);
  g_blas = h; g_sgemm = f; g_sgemv = v;
  g_threads = g_get_threads ? g_get_threads() : (threads > 0 ? threads : 1);
  return 0;
}
int orc_blas_threads(void) { return g_sgemm ? g_threads : 1; }

void orc_gemm(int transA, int transB, int M, int N, int K, float alpha,
              const float* A, const float* B, float beta, float* C) {
  cpu_gemm(transA != 0, transB != 0, M, N, K, alpha, A, B, beta, C);
}
void orc_gemv(int transA, int M, int N, float alpha, const float* A,
              const float* x, float beta, float* y) {
  cpu_gemv(transA != 0, M, N, alpha, A, x, beta, y);
}

// ----------------------------------------------------------------------------
// InnerProduct  (src/caffe/layers/inner_product_layer.cpp:61-106)
// ----------------------------------------------------------------------------
// Forward :61-73  Z = X W^T (NoTrans,Trans,M,N,K) ; Z += 1_M b^T (rank-1 gemm)
void orc_ip_forward(int M, int N, int K, const float* X, const float* W,
                    const float* bias, float* Z) {
  cpu_gemm(false, true, M, N, K, 1.f, X, W, 0.f, Z);
  if (bias) {
    std::vector<float> ones(M, 1.f);  // bias_multiplier_ :54-57
    cpu_gemm(false, false, M, N, 1, 1.f, ones.data(), bias, 1.f, Z);
  }
}
// Backward :76-106  dW = dZ^T X (Trans,NoTrans,N,K,M, beta=0 overwrite);
// if regularization/2 > 0: dW *= (1 + regularization/2) (:80,88-91);
// db = dZ^T 1_M (gemv Trans); dX = dZ W (NoTrans,NoTrans,M,K,N).
void orc_ip_backward(int M, int N, int K, const float* dZ, const float* X,
                     const float* W, double regularization_param,
                     float* dW, float* db, float* dX) {
  const double regularization = regularization_param / 2;
  if (dW) {
    cpu_gemm(true, false, N, K, M, 1.f, dZ, X, 0.f, dW);
    if (regularization > 0) cpu_scal((size_t)N * K, float(1.0 + regularization), dW);
  }
  if (db) {
    std::vector<float> ones(M, 1.f);
    cpu_gemv(true, M, N, 1.f, dZ, ones.data(), 0.f, db);
  }
  if (dX) cpu_gemm(false, false, M, K, N, 1.f, dZ, W, 0.f, dX);
}

// ----------------------------------------------------------------------------
// ReLU  (src/caffe/layers/relu_layer.cpp:10-36)
// ----------------------------------------------------------------------------
void orc_relu_forward(size_t n, const float* x, float negative_slope, float* y) {
  for (size_t i = 0; i < n; ++i)
    y[i] = std::max(x[i], 0.f) + negative_slope * std::min(x[i], 0.f);
}
// uses the PRE-activation bottom data (:27-35)
void orc_relu_backward(size_t n, const float* x, const float* dy,
                       float negative_slope, float* dx) {
  for (size_t i = 0; i < n; ++i)
    dx[i] = dy[i] * ((x[i] > 0) + negative_slope * (x[i] <= 0));
}

// ----------------------------------------------------------------------------
// Dropout  (src/caffe/layers/dropout_layer.cpp:13-68), TRAIN phase.
// The mask (0/1 unsigned, caffe_rng_bernoulli on boost mt19937, :41) is an
// explicit input: boost-derived values are "parity unpinned" (SURVEY 8c).
// scale_ = 1./(1.-threshold_) with threshold_ a float, stored as float (:17-20)
// ----------------------------------------------------------------------------
float orc_dropout_scale(float ratio) {
  return (float)(1. / (1. - ratio));
}
unsigned int orc_dropout_uint_thres(float ratio) {   // :21, used by the GPU path
  return static_cast<unsigned int>(UINT_MAX * ratio);
}
void orc_dropout_forward(size_t n, const float* x, const unsigned int* mask,
                         float ratio, float* y) {
  const float scale = orc_dropout_scale(ratio);
  for (size_t i = 0; i < n; ++i) y[i] = x[i] * mask[i] * scale;
}
void orc_dropout_backward(size_t n, const float* dy, const unsigned int* mask,
                          float ratio, float* dx) {
  const float scale = orc_dropout_scale(ratio);
  for (size_t i = 0; i < n; ++i) dx[i] = dy[i] * mask[i] * scale;
}

// ----------------------------------------------------------------------------
// Slice / Concat  (src/caffe/layers/slice_layer.cpp:79-133,
//                  src/caffe/layers/concat_layer.cpp:45-131)
// Blob [num, channels, inner]; equal slices (no slice_point in the shipped net)
// dim 0: contiguous blocks of num/ntop samples; dim 1: per-sample channel copy.
// ----------------------------------------------------------------------------
void orc_slice(int num, int channels, int inner, int dim, int ntop,
               const float* bottom, float** tops) {
  if (dim == 0) {
    const int n = num / ntop;
    const size_t blk = (size_t)n * channels * inner;
    for (int t = 0; t < ntop; ++t) memcpy(tops[t], bottom + t * blk, blk * sizeof(float));
  } else {
    const int c = channels / ntop;
    const size_t blk = (size_t)c * inner;
    for (int t = 0; t < ntop; ++t)
      for (int s = 0; s < num; ++s)
        memcpy(tops[t] + s * blk, bottom + ((size_t)s * channels + (size_t)t * c) * inner,
               blk * sizeof(float));
  }
}
// Concat of nb bottoms each [num_i, ch_i, inner] (equal sizes here).
void orc_concat(int num, int channels, int inner, int dim, int nb,
                float* const* bottoms, float* top) {
  // num/channels describe ONE bottom.
  if (dim == 0) {
    const size_t blk = (size_t)num * channels * inner;
    for (int t = 0; t < nb; ++t) memcpy(top + t * blk, bottoms[t], blk * sizeof(float));
  } else {
    const size_t blk = (size_t)channels * inner;
    for (int t = 0; t < nb; ++t)
      for (int s = 0; s < num; ++s)
        memcpy(top + ((size_t)s * nb + t) * blk, bottoms[t] + s * blk, blk * sizeof(float));
  }
}

// ----------------------------------------------------------------------------
// Eltwise  (src/caffe/layers/eltwise_layer.cpp:53-159)
// ----------------------------------------------------------------------------
// SUM forward :67-73 : top = 0; for i: axpy(coeff_i, bottom_i, top)
void orc_eltwise_sum_forward(size_t count, int nb, const float* const* bottoms,
                             const float* coeffs, float* top) {
  memset(top, 0, count * sizeof(float));
  for (int i = 0; i < nb; ++i) cpu_axpy(count, coeffs[i], bottoms[i], top);
}
// SUM backward :137-143 : copy if coeff==1 else caffe_cpu_scale (copy + scal)
void orc_eltwise_sum_backward(size_t count, float coeff, const float* top_diff,
                              float* bottom_diff) {
  memcpy(bottom_diff, top_diff, count * sizeof(float));
  if (coeff != 1.f) cpu_scal(count, coeff, bottom_diff);
}
// PROD forward :61-66 (two bottoms)
void orc_eltwise_prod_forward(size_t count, const float* a, const float* b, float* top) {
  for (size_t i = 0; i < count; ++i) top[i] = a[i] * b[i];
}
// PROD backward, stable_prod_grad=true (default) :119-136 with two bottoms:
// bottom_diff_i = other_bottom_data ; bottom_diff_i *= top_diff
void orc_eltwise_prod_backward(size_t count, const float* other, const float* top_diff,
                               float* bottom_diff) {
  for (size_t i = 0; i < count; ++i) bottom_diff[i] = other[i] * top_diff[i];
}

// ----------------------------------------------------------------------------
// Normalization (L2 per row)  (src/caffe/layers/normalization_layer.cpp)
// ----------------------------------------------------------------------------
// Forward :30-61 : t=pow(x,2); s=gemv(t,1); s=pow(s,.5); s+=1e-10;
//                  temp = s (x) 1^T (rank-1 gemm) ; y = x / temp
void orc_normalization_forward(int num, int dim, const float* x, float* y) {
  const size_t count = (size_t)num * dim;
  const float eps = 1e-10;
  std::vector<float> temp(count), ssq(num), ones(dim, 1.f);
  cpu_powx(count, x, 2.f, temp.data());
  cpu_gemv(false, num, dim, 1.f, temp.data(), ones.data(), 0.f, ssq.data());
  cpu_powx(num, ssq.data(), 0.5f, ssq.data());
  for (int i = 0; i < num; ++i) ssq[i] += eps;
  cpu_gemm(false, false, num, dim, 1, 1.f, ssq.data(), ones.data(), 0.f, temp.data());
  for (size_t i = 0; i < count; ++i) y[i] = x[i] / temp[i];
}
// Backward :64-112 : a = gemv(x.*dy,1); dx = x .* (a (x) 1); temp=pow(x,2);
//  s = gemv(temp,1); temp = s (x) 1; temp2 = temp .* dy; dx = temp2 - dx;
//  temp = pow(temp,1.5); temp += 1e-10; dx = dx / temp
void orc_normalization_backward(int num, int dim, const float* x, const float* dy,
                                float* dx) {
  const size_t count = (size_t)num * dim;
  const float eps = 1e-10;
  std::vector<float> temp(count), temp2(count), ssq(num), ones(dim, 1.f);
  for (size_t i = 0; i < count; ++i) temp[i] = x[i] * dy[i];
  cpu_gemv(false, num, dim, 1.f, temp.data(), ones.data(), 0.f, ssq.data());
  cpu_gemm(false, false, num, dim, 1, 1.f, ssq.data(), ones.data(), 0.f, dx);
  for (size_t i = 0; i < count; ++i) dx[i] = x[i] * dx[i];
  cpu_powx(count, x, 2.f, temp.data());
  cpu_gemv(false, num, dim, 1.f, temp.data(), ones.data(), 0.f, ssq.data());
  cpu_gemm(false, false, num, dim, 1, 1.f, ssq.data(), ones.data(), 0.f, temp.data());
  for (size_t i = 0; i < count; ++i) temp2[i] = temp[i] * dy[i];
  for (size_t i = 0; i < count; ++i) dx[i] = temp2[i] - dx[i];
  cpu_powx(count, temp.data(), 1.5f, temp.data());
  for (size_t i = 0; i < count; ++i) temp[i] += eps;
  for (size_t i = 0; i < count; ++i) dx[i] = dx[i] / temp[i];
}

// ----------------------------------------------------------------------------
// Sum  (src/caffe/layers/sum_layer.cpp:32-82)
// ----------------------------------------------------------------------------
void orc_sum_forward(int num, int dim, int num_output, const float* x, float* y) {
  std::vector<float> ones(dim, 1.f);
  if (num_output == 1) {
    cpu_gemv(false, num, dim, 1.f, x, ones.data(), 0.f, y);
  } else {
    std::vector<float> temp(num), ones2(num_output, 1.f);
    cpu_gemv(false, num, dim, 1.f, x, ones.data(), 0.f, temp.data());
    cpu_gemm(false, false, num, num_output, 1, 1.f, temp.data(), ones2.data(), 0.f, y);
  }
}
void orc_sum_backward(int num, int dim, int num_output, const float* dy, float* dx) {
  std::vector<float> ones(dim, 1.f);
  if (num_output == 1) {
    cpu_gemm(false, false, num, dim, 1, 1.f, dy, ones.data(), 0.f, dx);
  } else {
    std::vector<float> temp(num), ones2(num_output, 1.f);
    cpu_gemv(false, num, num_output, 1.f, dy, ones2.data(), 0.f, temp.data());
    cpu_gemm(false, false, num, dim, 1, 1.f, temp.data(), ones.data(), 0.f, dx);
  }
}

// ----------------------------------------------------------------------------
// Split backward  (src/caffe/layers/split_layer.cpp:36-51)
// d = top0 + top1 ; d += top_k for k = 2..
// ----------------------------------------------------------------------------
void orc_split_backward(size_t count, int ntop, const float* const* top_diffs, float* bottom_diff) {
  if (ntop == 1) { memcpy(bottom_diff, top_diffs[0], count * sizeof(float)); return; }
  for (size_t i = 0; i < count; ++i) bottom_diff[i] = top_diffs[0][i] + top_diffs[1][i];
  for (int k = 2; k < ntop; ++k) cpu_axpy(count, 1.f, top_diffs[k], bottom_diff);
}

// ----------------------------------------------------------------------------
// MaxMarginLoss  (src/caffe/layers/max_margin_loss_layer.cpp:54-127 fwd,
//                 :130-214 bwd).  norm: 1 = L1, 2 = L2.
// weights: optional per-element direct weights (3rd bottom, use_direct_weight)
// NOTE the reference's asymmetry: forward L2 uses sqrt(w)*h (:87), backward
// uses w*h (:154).  margin_ is float (loss_layers.hpp:1183).
// `hinge` (count floats) receives the forward temp (bottom[0].diff).
// ----------------------------------------------------------------------------
void orc_max_margin_forward(int count, const float* s_true, const float* s_bogus,
                            const float* weights, float margin, int norm,
                            float* hinge, float* loss, float* violations) {
  for (int i = 0; i < count; ++i) hinge[i] = s_true[i] - s_bogus[i];  // caffe_sub :69
  float num_violations = 0;
  for (int i = 0; i < count; ++i) {
    if (hinge[i] < 0) num_violations++;
    if (weights) {
      if (norm == 2) hinge[i] = std::sqrt(weights[i]) * std::max(0.f, margin - hinge[i]);
      else           hinge[i] = weights[i] * std::max(0.f, margin - hinge[i]);
    } else {
      hinge[i] = std::max(0.f, margin - hinge[i]);
    }
  }
  if (norm == 1) *loss = cpu_asum(count, hinge) / count;
  else           *loss = cpu_dot(count, hinge, hinge) / count;
  if (violations) *violations = num_violations;
}
void orc_max_margin_backward(int count, const float* s_true, const float* s_bogus,
                             const float* weights, float margin, int norm,
                             float loss_weight, float* d_true, float* d_bogus) {
  for (int i = 0; i < count; ++i) d_bogus[i] = s_true[i] - s_bogus[i];
  for (int i = 0; i < count; ++i) {
    if (weights) d_bogus[i] = weights[i] * std::max(0.f, margin - d_bogus[i]);
    else         d_bogus[i] = std::max(0.f, margin - d_bogus[i]);
  }
  if (norm == 1) {
    for (int i = 0; i < count; ++i)
      if (d_bogus[i] > 0.f) d_bogus[i] = weights ? weights[i] : 1.f;
    cpu_scal(count, loss_weight / count, d_bogus);
  } else {
    cpu_scal(count, loss_weight * 2 / count, d_bogus);
  }
  if (d_true) cpu_axpby(count, -1.f, d_bogus, 0.f, d_true);   // :210-212
}

// ----------------------------------------------------------------------------
// Solver  (src/caffe/solver.cpp:441-460 GetLearningRate, :486-576
// ComputeUpdateValue CPU branch, src/caffe/net.cpp:804-839 Net::Update,
// src/caffe/blob.cpp:113-136 Blob::Update)
// policy: 0 fixed, 1 step, 2 exp, 3 inv.  Evaluated in float like Dtype=float.
// ----------------------------------------------------------------------------
float orc_learning_rate(int policy, float base_lr, float gamma, float power,
                        int stepsize, int iter) {
  float rate;
  switch (policy) {
    case 0: rate = base_lr; break;
    case 1: { int current_step = iter / stepsize; rate = base_lr * std::pow(gamma, current_step); break; }
    case 2: rate = base_lr * std::pow(gamma, iter); break;
    default: rate = base_lr * std::pow(float(1) + gamma * iter, -power); break;
  }
  return rate;
}
// One param blob: diff += decay*data (L2) or decay*sign(data) (L1);
// hist = momentum*hist (scal) + local_rate*diff (axpy); diff = hist; data -= diff
// reg_type: 2 = L2, 1 = L1.
void orc_sgd_update(size_t count, float* data, float* diff, float* hist,
                    float local_rate, float momentum, float local_decay, int reg_type) {
  if (local_decay) {
    if (reg_type == 2) {
      cpu_axpy(count, local_decay, data, diff);
    } else {
      for (size_t i = 0; i < count; ++i) {
        const float s = (0.f < data[i]) - (data[i] < 0.f);   // caffe_cpu_sign
        diff[i] += local_decay * s;
      }
    }
  }
  cpu_axpby(count, local_rate, diff, momentum, hist);
  memcpy(diff, hist, count * sizeof(float));
  cpu_axpy(count, -1.f, diff, data);
}

// ----------------------------------------------------------------------------
// Whole TRAIN net forward+backward in the layer order of
// projects/videovec_embedding/mednet_embedding_train.prototxt (:2-671) after
// FilterNet + InsertSplits.  data is the data-layer blob
// [B, R, K]; mask is the dropout mask [R*B, N] over ip2 (NULL = no dropout
// layer, i.e. TEST-style copy).  All out pointers may be NULL.
// ----------------------------------------------------------------------------
struct orc_net_cfg {
  int B, C, Nn, K, N;
  float margin; int norm;          // max_margin_loss_param
  float dropout_ratio;             // dropout_param
  float negative_slope;            // relu_param
  float loss_weight;               // loss_weight of loss_output
  double regularization;           // inner_product_param.regularization
  int has_bias;
};
struct orc_net_out {
  float* X;            // [R*B, K]  original_feature
  float* Z;            // [R*B, N]  ip1_nonorm
  float* H;            // [R*B, N]  ip2 (after relu + dropout)
  float* target_score; // [B, Nn]
  float* neg_score;    // [B, Nn]
  float* loss;         // [1]
  float* violations;   // [1]
  float* dH;           // [R*B, N]  ip2.diff after slice_emb backward
  float* dZ;           // [R*B, N]  ip1_nonorm.diff
  float* dW;           // [N, K]
  float* db;           // [N]
  float* dX;           // [R*B, K]  (only if requested; the shipped net skips it)
  double* phase_seconds; // [8] wall-clock per phase (caffe-time style), optional
};

static double now_s() {
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int orc_net_forward_backward(const orc_net_cfg* cfg, const float* data,
                             const float* W, const float* bias,
                             const unsigned int* mask, orc_net_out* out) {
  const int B = cfg->B, C = cfg->C, Nn = cfg->Nn, K = cfg->K, N = cfg->N;
  const int R = C + Nn, M = R * B;
  const size_t BN = (size_t)B * N;
  double t0 = now_s(), t1;
  double ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  // slice_input_data (dim 1) + batch_concat_input (dim 0) + flatten_input
  std::vector<float> X((size_t)M * K);
  {
    std::vector<std::vector<float> > parts(R, std::vector<float>((size_t)B * K));
    std::vector<float*> ptr(R);
    for (int j = 0; j < R; ++j) ptr[j] = parts[j].data();
    orc_slice(B, R, K, 1, R, data, ptr.data());
    orc_concat(B, 1, K, 0, R, ptr.data(), X.data());
  }
  t1 = now_s(); ph[0] = t1 - t0; t0 = t1;

  // fc7, fc7_relu, drop2
  std::vector<float> Z((size_t)M * N), H((size_t)M * N);
  orc_ip_forward(M, N, K, X.data(), W, cfg->has_bias ? bias : nullptr, Z.data());
  t1 = now_s(); ph[1] = t1 - t0; t0 = t1;
  orc_relu_forward((size_t)M * N, Z.data(), cfg->negative_slope, H.data());
  if (mask) orc_dropout_forward((size_t)M * N, H.data(), mask, cfg->dropout_ratio, H.data());

  // slice_emb (dim 0): rows [j*B, (j+1)*B) ; 0 target, 1..C-1 context, C.. negatives
  const float* t_nonorm = H.data();
  std::vector<const float*> ctx(C - 1);
  for (int i = 0; i < C - 1; ++i) ctx[i] = H.data() + (size_t)(1 + i) * BN;
  // context_average: coeff 1/(C-1) each, written explicitly in the prototxt
  std::vector<float> coeffs(C - 1, 1.f / float(C - 1));
  std::vector<float> cbar(BN), chat(BN);
  orc_eltwise_sum_forward(BN, C - 1, ctx.data(), coeffs.data(), cbar.data());
  orc_normalization_forward(B, N, cbar.data(), chat.data());       // word_embedding_norm
  // concat_pos_neg_nonorm (dim 0): target, neg_1..neg_Nn ; pos_neg_normalize
  std::vector<float> P((size_t)(1 + Nn) * BN), Phat((size_t)(1 + Nn) * BN);
  memcpy(P.data(), t_nonorm, BN * sizeof(float));
  memcpy(P.data() + BN, H.data() + (size_t)C * BN, (size_t)Nn * BN * sizeof(float));
  orc_normalization_forward((1 + Nn) * B, N, P.data(), Phat.data());
  // prod_true + sum_true(num_output = Nn); prod_neg_k + sum_neg_k; concat dim 1
  std::vector<float> prod(BN), tscore((size_t)B * Nn), nscore((size_t)B * Nn), col(B);
  orc_eltwise_prod_forward(BN, chat.data(), Phat.data(), prod.data());
  orc_sum_forward(B, N, Nn, prod.data(), tscore.data());
  for (int k = 0; k < Nn; ++k) {
    orc_eltwise_prod_forward(BN, chat.data(), Phat.data() + (size_t)(1 + k) * BN, prod.data());
    orc_sum_forward(B, N, 1, prod.data(), col.data());
    for (int b = 0; b < B; ++b) nscore[(size_t)b * Nn + k] = col[b];   // concat dim 1
  }
  // max_margin_loss
  std::vector<float> hinge((size_t)B * Nn);
  float loss = 0.f, viol = 0.f;
  orc_max_margin_forward(B * Nn, tscore.data(), nscore.data(), nullptr, cfg->margin,
                         cfg->norm, hinge.data(), &loss, &viol);
  t1 = now_s(); ph[2] = t1 - t0; t0 = t1;

  // ---------------- backward ----------------
  std::vector<float> d_true((size_t)B * Nn), d_neg((size_t)B * Nn);
  orc_max_margin_backward(B * Nn, tscore.data(), nscore.data(), nullptr, cfg->margin,
                          cfg->norm, cfg->loss_weight, d_true.data(), d_neg.data());
  // per-branch diffs on the split tops of context_feature, and on Phat
  std::vector<std::vector<float> > dsplit(1 + Nn, std::vector<float>(BN));
  std::vector<float> dPhat((size_t)(1 + Nn) * BN), dprod(BN);
  // reverse layer order: neg_Nn .. neg_1, then true (order only matters for
  // which buffers are written, the split sum order below is fixed)
  for (int k = Nn - 1; k >= 0; --k) {
    for (int b = 0; b < B; ++b) col[b] = d_neg[(size_t)b * Nn + k];   // concat dim 1 bwd
    orc_sum_backward(B, N, 1, col.data(), dprod.data());
    orc_eltwise_prod_backward(BN, Phat.data() + (size_t)(1 + k) * BN, dprod.data(), dsplit[1 + k].data());
    orc_eltwise_prod_backward(BN, chat.data(), dprod.data(), dPhat.data() + (size_t)(1 + k) * BN);
  }
  orc_sum_backward(B, N, Nn, d_true.data(), dprod.data());
  orc_eltwise_prod_backward(BN, Phat.data(), dprod.data(), dsplit[0].data());
  orc_eltwise_prod_backward(BN, chat.data(), dprod.data(), dPhat.data());
  // slice_pos_neg_norm bwd (= concat), pos_neg_normalize bwd, concat bwd (= slice)
  std::vector<float> dP((size_t)(1 + Nn) * BN);
  orc_normalization_backward((1 + Nn) * B, N, P.data(), dPhat.data(), dP.data());
  // split backward on context_feature
  std::vector<float> dchat(BN), dcbar(BN);
  {
    std::vector<const float*> tops(1 + Nn);
    for (int k = 0; k <= Nn; ++k) tops[k] = dsplit[k].data();
    orc_split_backward(BN, 1 + Nn, tops.data(), dchat.data());
  }
  orc_normalization_backward(B, N, cbar.data(), dchat.data(), dcbar.data());
  // slice_emb backward: assemble ip2.diff
  std::vector<float> dH((size_t)M * N);
  memcpy(dH.data(), dP.data(), BN * sizeof(float));                               // target
  for (int i = 0; i < C - 1; ++i)
    orc_eltwise_sum_backward(BN, coeffs[i], dcbar.data(), dH.data() + (size_t)(1 + i) * BN);
  memcpy(dH.data() + (size_t)C * BN, dP.data() + BN, (size_t)Nn * BN * sizeof(float));
  // drop2 bwd (in place on ip2.diff), fc7_relu bwd
  std::vector<float> dZ((size_t)M * N);
  if (out && out->dH) memcpy(out->dH, dH.data(), dH.size() * sizeof(float));
  if (mask) orc_dropout_backward((size_t)M * N, dH.data(), mask, cfg->dropout_ratio, dH.data());
  orc_relu_backward((size_t)M * N, Z.data(), dH.data(), cfg->negative_slope, dZ.data());
  t1 = now_s(); ph[3] = t1 - t0; t0 = t1;
  // fc7 bwd
  std::vector<float> dW((size_t)N * K), db(N);
  orc_ip_backward(M, N, K, dZ.data(), X.data(), W, cfg->regularization, dW.data(),
                  cfg->has_bias ? db.data() : nullptr, (out && out->dX) ? out->dX : nullptr);
  t1 = now_s(); ph[4] = t1 - t0; t0 = t1;

  if (out) {
    if (out->X) memcpy(out->X, X.data(), X.size() * sizeof(float));
    if (out->Z) memcpy(out->Z, Z.data(), Z.size() * sizeof(float));
    if (out->H) memcpy(out->H, H.data(), H.size() * sizeof(float));
    if (out->target_score) memcpy(out->target_score, tscore.data(), tscore.size() * sizeof(float));
    if (out->neg_score) memcpy(out->neg_score, nscore.data(), nscore.size() * sizeof(float));
    if (out->loss) *out->loss = loss;
    if (out->violations) *out->violations = viol;
    if (out->dZ) memcpy(out->dZ, dZ.data(), dZ.size() * sizeof(float));
    if (out->dW) memcpy(out->dW, dW.data(), dW.size() * sizeof(float));
    if (out->db && cfg->has_bias) memcpy(out->db, db.data(), db.size() * sizeof(float));
    if (out->phase_seconds) memcpy(out->phase_seconds, ph, sizeof(ph));
  }
  return 0;
}

// ----------------------------------------------------------------------------
// Sampler: VideoSampledShotsDataLayer, context_type WINDOW
// (src/caffe/layers/video_sampled_shots_data_layer.cpp:25-44 AddToBuffer /
//  RandomShuffleTopids, :65-369 DataLayerSetUp, :372-507 AddSamplesToTop,
//  :769-909 InternalThreadEntry; include/caffe/util/rng.hpp:43-54
//  random_unique).  Uses the process-global glibc rand() exactly like the
// reference (never seeded there -> seed 1; callers srand(1) before create).
//
// Dataset = V records in DB key order; record v has video_id[v], shots
// [shot_off[v], shot_off[v+1]) with shot_ids[] and (optionally) features
// feat[shot, K].  Outputs per batch:
//   idx   [B, R] int32 : global shot index whose features fill each slot
//   quirk [B, R] int32 : -2 = full-row copy; otherwise the slot received only
//                        K-1 floats (:492) and element K-1 keeps the value of
//                        global shot `quirk` (>=0) or 0.0 (-1: never written,
//                        syncedmem.cpp:24-25 zero-fill)
//   data  [B, R, K]    : the materialised prefetch_data_ blob (if feat given)
// ----------------------------------------------------------------------------
}  // extern "C"

// random_unique (util/rng.hpp:43-54)
template <class It> static void orc_random_unique(It first, It last, int num_random) {
  int left = std::distance(first, last);
  while (num_random--) {
    It r = first;
    std::advance(r, rand() % left);
    std::swap(*first, *r);
    ++first; --left;
  }
}

struct OrcSampler {
  int mode = 1;
  int V, K, B, C, Nn, P, swap_pct, max_same;
  const int* video_id; const int* shot_off; const int* shot_ids; const float* feat;
  int cursor;
  std::vector<float> buffer_ids;          // vector<float> like the reference (:81-83)
  std::vector<int> neg_shot;              // buffer slot -> global shot index
  std::vector<float> negatives;           // [P, K] copy (only if feat)
  std::vector<std::string> id_to_key;     // negative_id_to_key_
  std::set<std::string> key_set;          // negative_keys_set_
  std::vector<float> prefetch;            // [B, R, K] persistent (only if feat)
  std::vector<int> last_full;             // [B, R] shot of the last full-row write, -1 none
};

static std::string orc_key(int vid, int shot_id) {
  char buf[64]; snprintf(buf, sizeof(buf), "%d:%d", vid, shot_id);
  return std::string(buf);
}

extern "C" {

void* orc_sampler_create_mode(int V, int K, const int* video_id, const int* shot_off,
                              const int* shot_ids, const float* feat,
                              int batch_size, int context_size, int num_negative_samples,
                              int max_buffer_size, int negative_swap_percentage,
                              int max_same_video_negs, int max_tries_for_negs, int context_type);
void* orc_sampler_create(int V, int K, const int* video_id, const int* shot_off,
                         const int* shot_ids, const float* feat,
                         int batch_size, int context_size, int num_negative_samples,
                         int max_buffer_size, int negative_swap_percentage,
                         int max_same_video_negs, int max_tries_for_negs) {
  return orc_sampler_create_mode(V, K, video_id, shot_off, shot_ids, feat, batch_size, context_size, num_negative_samples,
                                 max_buffer_size, negative_swap_percentage, max_same_video_negs, max_tries_for_negs, 1);
}
void* orc_sampler_create_opts(int V, int K, const int* video_id, const int* shot_off,
                              const int* shot_ids, const float* feat,
                              int batch_size, int context_size, int num_negative_samples,
                              int max_buffer_size, int negative_swap_percentage,
                              int max_same_video_negs, int max_tries_for_negs, int context_type,
                              int start_skip, int negV, const int* neg_video_id, const int* neg_shot_off,
                              const int* neg_shot_ids, const float* neg_feat, int neg_row_base);
// context_type: VideoSampledShotsDataParameter.CONTEXT (caffe.proto:598-604): 0 PAIRWISE, 1 WINDOW, 2 PAST,
// 3 PAST_CONTINUOUS, 4 PAST_CONTINUOUS_FIXED
void* orc_sampler_create_mode(int V, int K, const int* video_id, const int* shot_off,
                              const int* shot_ids, const float* feat,
                              int batch_size, int context_size, int num_negative_samples,
                              int max_buffer_size, int negative_swap_percentage,
                              int max_same_video_negs, int max_tries_for_negs, int context_type) {
  return orc_sampler_create_opts(V, K, video_id, shot_off, shot_ids, feat, batch_size, context_size, num_negative_samples,
                                 max_buffer_size, negative_swap_percentage, max_same_video_negs, max_tries_for_negs, context_type,
                                 0, 0, nullptr, nullptr, nullptr, nullptr, 0);
}
// + the data layer's rand_skip (:157-180; start_skip = the value drawn, caffe_rng_rand() % rand_skip) and
// negative_dataset (:137-153, 273-284, 324-338): a second record set whose shots -- ALL of them, record after record, no
// rand() -- seed the negative buffer; emitted indices of those shots are neg_row_base + their index in that set
void* orc_sampler_create_opts(int V, int K, const int* video_id, const int* shot_off,
                              const int* shot_ids, const float* feat,
                              int batch_size, int context_size, int num_negative_samples,
                              int max_buffer_size, int negative_swap_percentage,
                              int max_same_video_negs, int max_tries_for_negs, int context_type,
                              int start_skip, int negV, const int* neg_video_id, const int* neg_shot_off,
                              const int* neg_shot_ids, const float* neg_feat, int neg_row_base) {
  if (max_same_video_negs > num_negative_samples) return nullptr;   // undefined in the reference (slot overflow)
  if (context_type < 0 || context_type > 4) return nullptr;
  if (context_type == 0 && context_size != 2) return nullptr;       // PAIRWISE fills channels 0 and 1 only (:396-404)
  OrcSampler* s = new OrcSampler();
  s->mode = context_type;
  s->V = V; s->K = K; s->B = batch_size; s->C = context_size; s->Nn = num_negative_samples;
  s->P = num_negative_samples > 0 ? max_buffer_size : 0;
  s->swap_pct = negative_swap_percentage; s->max_same = max_same_video_negs;
  s->video_id = video_id; s->shot_off = shot_off; s->shot_ids = shot_ids; s->feat = feat;
  s->cursor = start_skip % V;                                           // rand_skip: MDB_NEXT x skip, wrapping (:161-178)
  const int R = s->C + s->Nn;
  for (int i = 0; i < s->P; ++i) s->buffer_ids.push_back(i);            // :81-83
  if (feat) { s->prefetch.assign((size_t)s->B * R * K, 0.f); s->negatives.assign((size_t)s->P * K, 0.f); }
  s->last_full.assign((size_t)s->B * R, -1);
  s->neg_shot.assign(s->P, -1);
  // negative buffer init :245-344 : one rand()%num_shots per record visited
  int added = 0;
  int ncur = 0;
  for (long nid = 0; negV > 0 && nid < (long)max_tries_for_negs * s->P; ++nid) {
    // negative_dataset: every shot of the record at the negative cursor (:324-338); the reference copies without a
    // bound check, so running past max_buffer_size inside a record is a buffer overflow there -- an error here
    const int v = ncur;
    ncur = (ncur + 1) % negV;
    for (int g = neg_shot_off[v]; g < neg_shot_off[v + 1]; ++g) {
      const std::string key = orc_key(neg_video_id[v], neg_shot_ids[g]);
      if (s->key_set.find(key) == s->key_set.end()) {
        if (added >= s->P) { delete s; return nullptr; }
        if (feat && neg_feat) memcpy(&s->negatives[(size_t)added * K], neg_feat + (size_t)g * K, K * sizeof(float));
        s->neg_shot[added] = neg_row_base + g;
        s->id_to_key.push_back(key);
        s->key_set.insert(key);
        added++;
      }
    }
    if (added >= s->P) break;
  }
  for (long nid = 0; negV == 0 && nid < (long)max_tries_for_negs * s->P; ++nid) {
    const int v = s->cursor;
    s->cursor = (s->cursor + 1) % V;                                     // MDB_NEXT / wrap
    const int num_shots = shot_off[v + 1] - shot_off[v];
    const int sample_shot = rand() % num_shots;                          // :302
    const std::string key = orc_key(video_id[v], shot_ids[shot_off[v] + sample_shot]);
    if (s->key_set.find(key) == s->key_set.end()) {
      const int g = shot_off[v] + sample_shot;
      if (feat) memcpy(&s->negatives[(size_t)added * K], feat + (size_t)g * K, K * sizeof(float));
      s->neg_shot[added] = g;
      s->id_to_key.push_back(key);
      s->key_set.insert(key);
      added++;
    }
    if (added >= s->P) break;
  }
  if (added != s->P) { delete s; return nullptr; }                        // CHECK_EQ :346
  return s;
}

void orc_sampler_destroy(void* h) { delete (OrcSampler*)h; }
int orc_sampler_cursor(void* h) { return ((OrcSampler*)h)->cursor; }
void orc_sampler_buffer_shots(void* h, int* out) {
  OrcSampler* s = (OrcSampler*)h; memcpy(out, s->neg_shot.data(), s->P * sizeof(int));
}


int orc_sampler_next(void* h, int* idx, int* quirk, float* data) {
  OrcSampler* s = (OrcSampler*)h;
  const int K = s->K, B = s->B, C = s->C, Nn = s->Nn, R = C + Nn;
  float* top = s->feat ? s->prefetch.data() : nullptr;
  int item_id = 0;
  long guard = 0;
  while (item_id < B) {
    if (++guard > 100L * (B + s->V)) return -1;    // dataset has no usable record
    const int v = s->cursor;
    s->cursor = (s->cursor + 1) % s->V;            // advanced before knowing if used (:826-846)
    const int off = s->shot_off[v], n = s->shot_off[v + 1] - off;
    // ---- AddSamplesToTop (:372-757)
    if (n < 2) continue;                           // :387-389
    std::vector<int> ids(n);
    std::iota(ids.begin(), ids.end(), 0);
    int added = 0;
    auto put_full = [&](int slot, int shot) {      // a full K-float copy into a slot
      const int g = off + shot;
      if (top) memcpy(top + ((size_t)item_id * R + slot) * K, s->feat + (size_t)g * K, K * sizeof(float));
      idx[item_id * R + slot] = g; quirk[item_id * R + slot] = -2;
      s->last_full[item_id * R + slot] = g;
    };
    auto put_negative = [&](int shot) {            // same-video negative: datum_height_-1 floats only (:492, :567, :655, :742)
      const int slot = C + added;
      const int g = off + shot;
      if (top) memcpy(top + ((size_t)item_id * R + slot) * K, s->feat + (size_t)g * K, (K - 1) * sizeof(float));
      idx[item_id * R + slot] = g;
      quirk[item_id * R + slot] = s->last_full[item_id * R + slot];   // -1 or a shot
      added++;
    };
    if (s->mode == 0) {                            // PAIRWISE :396-422
      orc_random_unique(ids.begin(), ids.end(), 2);
      put_full(0, ids[0]); put_full(1, ids[1]);
    } else if (s->mode == 1) {                     // WINDOW :425-506
      if ((int)ids.size() < C) continue;           // :427-429
      orc_random_unique(ids.begin(), ids.end(), C);  // :432
      std::sort(ids.begin(), ids.begin() + C);       // :437
      const int half = C / 2;
      int context_id = 0;
      for (int i = 0; i < C; ++i) put_full((i == half) ? 0 : (context_id++ + 1), ids[i]);
      if (Nn > 0 && n > C) {                          // :479-503
        std::random_shuffle(ids.begin() + C, ids.end());
        for (int nid = C; nid < n && added < s->max_same; ++nid)
          if (ids[nid] < ids[half - 1] || ids[nid] > ids[half + 1]) put_negative(ids[nid]);
      }
    } else if (s->mode == 2) {                     // PAST :509-583: the target is the LAST of the sorted window
      if ((int)ids.size() < C) continue;
      orc_random_unique(ids.begin(), ids.end(), C);
      std::sort(ids.begin(), ids.begin() + C);
      int context_id = 0;
      for (int i = 0; i < C; ++i) put_full((i == C - 1) ? 0 : (context_id++ + 1), ids[i]);
      if (Nn > 0 && n > C) {
        std::random_shuffle(ids.begin() + C, ids.end());
        for (int nid = C; nid < n && added < s->max_same; ++nid)
          if (ids[nid] < ids[1]) put_negative(ids[nid]);          // :562
      }
    } else {                                       // PAST_CONTINUOUS :586-671 / PAST_CONTINUOUS_FIXED :674-757
      if ((int)ids.size() < C) continue;
      const int max_sample_length = (n - C) / (C - 1);
      int sample_length, begin_frame;
      if (s->mode == 3) {
        sample_length = rand() % (max_sample_length + 1);                                   // :596
        begin_frame = rand() % (n - (C - 1) * sample_length - C + 1);                       // :598-599
      } else {
        sample_length = (max_sample_length >= 1) ? (max_sample_length - 1) : 0;             // :685
        begin_frame = n - (C - 1) * sample_length - C;                                      // :687-688
      }
      int context_id = 0;
      for (int i = 0; i < C; ++i) put_full((i == C - 1) ? 0 : (context_id++ + 1), begin_frame + i * (sample_length + 1));
      if (Nn > 0 && begin_frame > 0)
        for (int nid = begin_frame - 1; nid >= 0 && added < s->max_same; --nid) put_negative(nid);   // :645-660
    }
    // ---- remaining negatives from the buffer (:852-874)
    if (Nn > 0) {
      orc_random_unique(s->buffer_ids.begin(), s->buffer_ids.end(), Nn - added);   // :42-44,855
      for (int negative_id = C + added; negative_id < C + Nn; ++negative_id) {
        const int neg_id = static_cast<int>(s->buffer_ids[negative_id - C - added]);
        if (top) memcpy(top + ((size_t)item_id * R + negative_id) * K,
                        &s->negatives[(size_t)neg_id * K], K * sizeof(float));
        idx[item_id * R + negative_id] = s->neg_shot[neg_id];
        quirk[item_id * R + negative_id] = -2;
        s->last_full[item_id * R + negative_id] = s->neg_shot[neg_id];
      }
    }
    item_id++;
    // ---- swap this record's shots into the buffer (:888-906, AddToBuffer :25-37)
    if (Nn > 0 && s->swap_pct > 0) {
      for (int j = 0; j < n; ++j) {
        const std::string key = orc_key(s->video_id[v], s->shot_ids[off + j]);
        if (s->key_set.find(key) == s->key_set.end()) {
          int pos = -1;
          if ((rand() % 100) < s->swap_pct) {
            pos = rand() % s->P;
            if (s->feat) memcpy(&s->negatives[(size_t)pos * K], s->feat + (size_t)(off + j) * K, K * sizeof(float));
          }
          if (pos >= 0) {
            s->neg_shot[pos] = off + j;
            const std::string old_key = s->id_to_key[pos];
            s->id_to_key[pos] = key;
            s->key_set.erase(old_key);
            s->key_set.insert(key);
          }
        }
      }
    }
  }
  if (data && top) memcpy(data, top, (size_t)B * R * K * sizeof(float));
  return 0;
}

// ----------------------------------------------------------------------------
// TEST phase: mednet_embedding_train.prototxt TEST graph + RetrievalStatsLayer
// ----------------------------------------------------------------------------
// Embedding of a test batch: frames [B, F, K] (F = 4 sampled frames per shot window) ->
//   slice dim 1 + concat dim 0 + flatten + slice dim 0 + ELTWISE SUM coeff 1/F ("average_for_test", :85-102 of the
//   prototxt; eltwise_layer.cpp:67-73: top = 0, then axpy per bottom) -> fc7 -> ReLU -> Dropout(TEST = copy,
//   dropout_layer.cpp:46-48) -> NORMALIZATION ("test_norm").
void orc_test_embed(int B, int F, int K, int N, const float* frames, const float* coeff, const float* W,
                    const float* bias, float* xbar_out, float* E) {
  std::vector<float> xbar((size_t)B * K, 0.f), Z((size_t)B * N), H((size_t)B * N);
  std::vector<float> plane((size_t)B * K);
  for (int f = 0; f < F; ++f) {
    for (int b = 0; b < B; ++b) memcpy(&plane[(size_t)b * K], frames + ((size_t)b * F + f) * K, K * sizeof(float));
    cpu_axpy((size_t)B * K, coeff[f], plane.data(), xbar.data());
  }
  if (xbar_out) memcpy(xbar_out, xbar.data(), xbar.size() * sizeof(float));
  orc_ip_forward(B, N, K, xbar.data(), W, bias, Z.data());
  orc_relu_forward((size_t)B * N, Z.data(), 0.f, H.data());
  orc_normalization_forward(B, N, H.data(), E);
}

// RetrievalStatsLayer::Forward_cpu, shot-level branch (retrieval_stats_layer.cpp:143-359) with ComputeStats (:98-140).
// E [B, N] embeddings, video_ids [B], labels [B] = video_id_to_class_[video_id] (< 0: sample not scored, :256-258).
// dist_io: if non-NULL and *use_dist != 0 the given [B,B] matrix is used instead of -2 E E^T (to compare rank
// statistics on identical distances); it always receives the matrix used (diagonal set to -1e15, :240-241).
// out = {mean AP, hit@1, hit@5}; per_query (optional) [B,3].
void orc_retrieval_stats(int B, int N, const float* E, const int* video_ids, const int* labels,
                         int exclude_same_video_shots, float* dist_io, int use_dist, double* out, double* per_query) {
  std::vector<float> D((size_t)B * B);
  if (dist_io && use_dist) memcpy(D.data(), dist_io, D.size() * sizeof(float));
  else cpu_gemm(false, true, B, B, N, -2.f, E, E, 0.f, D.data());                       // :226-228
  double mean_ap = 0, mean_acc_1 = 0, mean_acc_5 = 0, num_positives = 0;
  std::vector<int> sort_ids(B);
  for (int i = 0; i < B; ++i) {
    std::iota(sort_ids.begin(), sort_ids.end(), 0);
    D[(size_t)i * B + i] = -1e15f;                                                      // :240-241
    const float* row = &D[(size_t)i * B];
    // ties broken by index: the reference's std::sort leaves their order unspecified
    std::sort(sort_ids.begin(), sort_ids.end(), [row](int a, int b) { return row[a] < row[b] || (row[a] == row[b] && a < b); });
    if (per_query) per_query[3 * i] = per_query[3 * i + 1] = per_query[3 * i + 2] = -1;
    if (labels[i] < 0) continue;                                                        // :256-258
    // ComputeStats :98-140
    double ap = 0, acc_1 = 0, acc_5 = 0, val = 0, ret = 0;
    for (int k = 1; k < B; ++k) {                                                       // the first hit is the query itself
      const int j = sort_ids[k];
      if (video_ids[j] != video_ids[i] || !exclude_same_video_shots) {
        val++;
        if (labels[j] == labels[i]) {
          if (val <= 1) acc_1++;
          if (val <= 5) acc_5++;
          ret++;
          ap += ret / val;
        }
      }
    }
    if (ret > 0) ap /= ret;
    acc_5 /= 5;
    mean_ap += ap; mean_acc_1 += acc_1; mean_acc_5 += acc_5; num_positives++;
    if (per_query) { per_query[3 * i] = ap; per_query[3 * i + 1] = acc_1; per_query[3 * i + 2] = acc_5; }
  }
  if (dist_io) memcpy(dist_io, D.data(), D.size() * sizeof(float));
  out[0] = mean_ap / num_positives; out[1] = mean_acc_1 / num_positives; out[2] = mean_acc_5 / num_positives;   // :352-354
}

// IdToWeightMappingLayer (id_to_weight_mapping_layer.cpp:61-77 forward, :80-107 backward)
void orc_id_lookup_forward(int M, int N, const float* table, const float* ids, float* top) {
  for (int i = 0; i < M; ++i) memcpy(top + (size_t)i * N, table + (size_t)static_cast<int>(ids[i]) * N, N * sizeof(float));
}
void orc_id_lookup_backward(int M, int N, int rows, const float* top_diff, const float* ids, float* table_diff) {
  memset(table_diff, 0, (size_t)rows * N * sizeof(float));                                   // caffe_set :98
  for (int i = 0; i < M; ++i) cpu_axpy(N, 1.f, top_diff + (size_t)i * N, table_diff + (size_t)static_cast<int>(ids[i]) * N);   // :100-106
}

void orc_srand(unsigned seed) { srand(seed); }
int orc_rand(void) { return rand(); }

}  // extern "C"
