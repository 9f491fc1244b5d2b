// vv_capi.cu -- C-ABI glue: error reporting, device check, and the fc7 entry points
// that pick between the tcgen05 kernels and the exact fp32 kernel by precision.
#include <stdarg.h>
#include <stdio.h>
#include "vv_gemm.cuh"

namespace vv {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", int(e), cudaGetErrorString(e), file, line, what);
  return VV_ERR_CUDA;
}
int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = kNumSMsB200;
  }
  return n;
}
void count_launch(int n) { g_launches += n; }
int launches_reset() { const int n = g_launches; g_launches = 0; return n; }

static int fill_epilogue(const vv_act_t* act, const float* bias, float* Z, GemmEpilogue* e) {
  e->bias = bias; e->Z = Z; e->has_act = act ? 1 : 0; e->out_scale = 1.f;
  e->relu = 0; e->negative_slope = 0.f; e->dropout_mode = VV_DROPOUT_NONE; e->dropout_scale = 1.f;
  e->dropout_thres = 0; e->mask = nullptr; e->mask_out = nullptr; e->seed = 0; e->step = 0; e->hash_base = 0;
  e->delta = nullptr; e->wlast = nullptr;
  if (!act) return VV_OK;
  e->relu = act->relu; e->negative_slope = act->negative_slope;
  e->dropout_mode = act->dropout_mode;
  if (act->dropout_mode != VV_DROPOUT_NONE) {
    VV_REQUIRE(act->dropout_ratio > 0.f && act->dropout_ratio < 1.f, "dropout_ratio must be in (0,1)");
    e->dropout_scale = dropout_scale(act->dropout_ratio);
    e->dropout_thres = dropout_uint_thres(act->dropout_ratio);
    if (act->dropout_mode == VV_DROPOUT_MASK01 || act->dropout_mode == VV_DROPOUT_MASK_U32)
      VV_REQUIRE(act->mask, "dropout mask mode needs a mask pointer");
    e->mask = act->mask; e->mask_out = act->mask_out; e->seed = act->seed; e->step = act->step;
    e->hash_base = dropout_hash_base(act->seed, act->step);
  }
  return VV_OK;
}

static int run_gemm(const GemmProblem& g, cudaStream_t stream) {
  if (g.prec == VV_PREC_FP32_SIMT) return gemm_simt_launch(g, stream);
  return gemm_tc_launch(g, stream);
}

}  // namespace vv

using namespace vv;

extern "C" const char* vv_last_error(void) { return g_err; }
extern "C" int vv_version(void) { return 100; }

extern "C" int vv_device_check(void) {
  int dev = 0; cudaDeviceProp prop;
  VV_CUDA(cudaGetDevice(&dev));
  VV_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) { set_error("device %s is sm_%d%d, libvv_b200 is built for sm_100a only", prop.name, prop.major, prop.minor); return VV_ERR_UNSUPPORTED; }
  return VV_OK;
}

extern "C" int vv_ip_forward(vv_operand_t X, vv_operand_t W, const float* bias, int M, int N, int K, int prec,
                             const vv_act_t* act, float* Z, float* H, vv_stream_t stream) {
  return vv::ip_forward_ex(X, W, bias, M, N, K, prec, act, Z, H, nullptr, stream);
}
int vv::ip_forward_ex(vv_operand_t X, vv_operand_t W, const float* bias, int M, int N, int K, int prec,
                      const vv_act_t* act, float* Z, float* H, const DpWait* wait, vv_stream_t stream, const FwdTail* tail) {
  VV_REQUIRE(X.hi && W.hi && H && M > 0 && N > 0 && K > 0, "ip_forward: bad arguments");
  GemmProblem g;
  if (wait) g.wait = *wait;
  if (tail) { g.tail_ws = tail->ws; g.tail_ws_bytes = tail->ws_bytes; g.tail_flags = tail->flags; g.tail_flags_count = tail->flags_count; g.tail_epoch = tail->epoch; }
  g.kind = GEMM_FWD; g.prec = prec; g.A = X; g.B = W; g.M = M; g.N = N; g.K = K;
  g.D = H; g.slab_stride = 0; g.nsplit = 1; g.rowmap = nullptr; g.bank_rows = 0;
  int rc = fill_epilogue(act, bias, Z, &g.epi);
  if (rc) return rc;
  if (!act && Z && Z != H) { set_error("ip_forward: without an activation pass H only (Z == NULL)"); return VV_ERR_INVALID; }
  if (!act) g.epi.Z = nullptr;
  return run_gemm(g, stream);
}

extern "C" int vv_ip_wgrad_auto_nsplit(int M, int N, int K, int prec) {
  if (prec == VV_PREC_FP32_SIMT) return 1;
  // tiles of 128 x 256 over the [N, K] output; split the M reduction so that the
  // persistent grid of num_sms CTAs is filled in (almost) whole waves.
  const int tiles = ((N + 127) / 128) * ((K + 255) / 256);
  const int bk = (prec == VV_PREC_BF16 || prec == VV_PREC_F16X3) ? 64 : 32;
  const int num_kb = (M + bk - 1) / bk;
  const int sms = num_sms();
  int best = 1; double best_cost = 1e30;
  for (int s = 1; s <= 16 && s <= num_kb; ++s) {
    const int per = (num_kb + s - 1) / s;
    if ((s - 1) * per >= num_kb) continue;
    const int units = tiles * s;
    const int waves = (units + sms - 1) / sms;
    // cost ~ waves * k-blocks per unit (+ a small per-slab epilogue/reduction charge)
    const double cost = double(waves) * per + 0.02 * num_kb * s / 16.0 + 8.0 * waves;
    if (cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

extern "C" size_t vv_ip_wgrad_workspace_bytes(int M, int N, int K, int prec) {
  const int s = vv_ip_wgrad_auto_nsplit(M, N, K, prec);
  return s > 1 ? size_t(s) * N * K * sizeof(float) : 0;
}

extern "C" int vv_ip_wgrad(vv_operand_t dZ, vv_operand_t X, int M, int N, int K, int prec, float regularization,
                           float* dW_parts, int nsplit, void* workspace, size_t workspace_bytes, vv_stream_t stream) {
  return vv::ip_wgrad_ex(dZ, X, M, N, K, prec, regularization, dW_parts, nsplit, workspace, workspace_bytes, nullptr, stream);
}
int vv::ip_wgrad_ex(vv_operand_t dZ, vv_operand_t X, int M, int N, int K, int prec, float regularization,
                    float* dW_parts, int nsplit, void* workspace, size_t workspace_bytes, const WgradFinish* finish, vv_stream_t stream) {
  VV_REQUIRE(dZ.hi && X.hi && dW_parts && M > 0 && N > 0 && K > 0 && nsplit >= 0, "ip_wgrad: bad arguments");
  VV_REQUIRE(!finish || (nsplit >= 1 && prec != VV_PREC_FP32_SIMT), "ip_wgrad: the fused split-K finish needs explicit slabs and a tensor-core precision");
  GemmProblem g;
  g.finish = finish;
  g.kind = GEMM_WGRAD; g.prec = prec; g.A = dZ; g.B = X; g.M = M; g.N = N; g.K = K; g.rowmap = nullptr; g.bank_rows = 0;
  int rc = fill_epilogue(nullptr, nullptr, nullptr, &g.epi);
  if (rc) return rc;
  // ref: inner_product_layer.cpp:80,88-91  dW *= (1 + regularization/2) when regularization/2 > 0
  const double reg = double(regularization) / 2;
  g.epi.out_scale = reg > 0 ? float(1.0 + reg) : 1.f;
  g.slab_stride = (long long)N * K;
  if (nsplit >= 1) {
    g.D = dW_parts; g.nsplit = nsplit;
    return run_gemm(g, stream);
  }
  const int s = vv_ip_wgrad_auto_nsplit(M, N, K, prec);
  if (s == 1) { g.D = dW_parts; g.nsplit = 1; return run_gemm(g, stream); }
  VV_REQUIRE(workspace && workspace_bytes >= size_t(s) * N * K * sizeof(float),
             "ip_wgrad: workspace too small (%zu bytes, need %zu)", workspace_bytes, size_t(s) * N * K * sizeof(float));
  g.D = static_cast<float*>(workspace); g.nsplit = s;
  rc = run_gemm(g, stream);
  if (rc) return rc;
  return vv_reduce_parts(static_cast<const float*>(workspace), s, (long long)N * K, (long long)N * K, dW_parts, stream);
}

extern "C" int vv_ip_dgrad(vv_operand_t dZ, vv_operand_t W, int M, int N, int K, int prec, float* dX, vv_stream_t stream) {
  VV_REQUIRE(dZ.hi && W.hi && dX && M > 0 && N > 0 && K > 0, "ip_dgrad: bad arguments");
  GemmProblem g;
  g.kind = GEMM_DGRAD; g.prec = prec; g.A = dZ; g.B = W; g.M = M; g.N = N; g.K = K;
  g.D = dX; g.slab_stride = 0; g.nsplit = 1; g.rowmap = nullptr; g.bank_rows = 0;
  int rc = fill_epilogue(nullptr, nullptr, nullptr, &g.epi);
  if (rc) return rc;
  return run_gemm(g, stream);
}

// ---- gather-fused variants (K0 folded into K1's TMA producer) --------------------------------------------
extern "C" int vv_ip_forward_gathered(vv_operand_t bank, int64_t bank_rows, const int32_t* rowmap, const float* delta,
                                      const float* wlast, vv_operand_t W, const float* bias, int M, int N, int K, int prec,
                                      const vv_act_t* act, float* Z, float* H, vv_stream_t stream) {
  return vv::ip_forward_gathered_ex(bank, bank_rows, rowmap, delta, wlast, W, bias, M, N, K, prec, act, Z, H, nullptr, stream);
}
int vv::ip_forward_gathered_ex(vv_operand_t bank, int64_t bank_rows, const int32_t* rowmap, const float* delta,
                               const float* wlast, vv_operand_t W, const float* bias, int M, int N, int K, int prec,
                               const vv_act_t* act, float* Z, float* H, const DpWait* wait, vv_stream_t stream, const FwdTail* tail) {
  VV_REQUIRE(bank.hi && rowmap && W.hi && H && M > 0 && N > 0 && K > 0 && bank_rows > 0, "ip_forward_gathered: bad arguments");
  VV_REQUIRE(prec != VV_PREC_FP32_SIMT, "ip_forward_gathered needs a tensor-core precision");
  VV_REQUIRE(!delta || wlast, "ip_forward_gathered: delta needs wlast");
  GemmProblem g;
  g.kind = GEMM_FWD; g.prec = prec; g.A = bank; g.B = W; g.M = M; g.N = N; g.K = K;
  g.D = H; g.slab_stride = 0; g.nsplit = 1; g.rowmap = rowmap; g.bank_rows = bank_rows;
  if (wait) g.wait = *wait;
  if (tail) { g.tail_ws = tail->ws; g.tail_ws_bytes = tail->ws_bytes; g.tail_flags = tail->flags; g.tail_flags_count = tail->flags_count; g.tail_epoch = tail->epoch; }
  int rc = fill_epilogue(act, bias, Z, &g.epi);
  if (rc) return rc;
  if (!act) g.epi.Z = nullptr;
  g.epi.delta = delta; g.epi.wlast = wlast;
  return gemm_tc_launch(g, stream);
}

extern "C" int vv_ip_wgrad_gathered(vv_operand_t dZ, vv_operand_t bank, int64_t bank_rows, const int32_t* rowmap, int M, int N,
                                    int K, int prec, float regularization, float* dW_parts, int nsplit, vv_stream_t stream) {
  return vv_ip_wgrad_gathered_part(dZ, bank, bank_rows, rowmap, M, N, K, prec, regularization, dW_parts, nsplit, 0, N, stream);
}
extern "C" int vv_ip_wgrad_gathered_part(vv_operand_t dZ, vv_operand_t bank, int64_t bank_rows, const int32_t* rowmap, int M, int N,
                                         int K, int prec, float regularization, float* dW_parts, int nsplit, int n0, int ncols,
                                         vv_stream_t stream) {
  return vv::ip_wgrad_gathered_ex(dZ, bank, bank_rows, rowmap, M, N, K, prec, regularization, dW_parts, nsplit, n0, ncols, nullptr, stream);
}
int vv::ip_wgrad_gathered_ex(vv_operand_t dZ, vv_operand_t bank, int64_t bank_rows, const int32_t* rowmap, int M, int N,
                             int K, int prec, float regularization, float* dW_parts, int nsplit, int n0, int ncols,
                             const WgradFinish* finish, vv_stream_t stream) {
  VV_REQUIRE(dZ.hi && bank.hi && rowmap && dW_parts && M > 0 && N > 0 && K > 0 && nsplit >= 1 && bank_rows > 0,
             "ip_wgrad_gathered: bad arguments");
  VV_REQUIRE(prec != VV_PREC_FP32_SIMT, "ip_wgrad_gathered needs a tensor-core precision");
  GemmProblem g;
  // computed as (X^T dZ)^T so that the gathered operand is A (see GEMM_WGRAD_T)
  g.kind = GEMM_WGRAD_T; g.prec = prec; g.A = bank; g.B = dZ; g.M = M; g.N = N; g.K = K;
  g.rowmap = rowmap; g.bank_rows = bank_rows;
  int rc = fill_epilogue(nullptr, nullptr, nullptr, &g.epi);
  if (rc) return rc;
  const double reg = double(regularization) / 2;
  g.epi.out_scale = reg > 0 ? float(1.0 + reg) : 1.f;
  g.slab_stride = (long long)N * K; g.D = dW_parts; g.nsplit = nsplit;
  g.n0 = n0; g.ncols = ncols;
  g.finish = finish;
  return gemm_tc_launch(g, stream);
}
