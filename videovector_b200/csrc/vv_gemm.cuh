// vv_gemm.cuh -- internal interface of the projection GEMMs (K1).
#pragma once
#include "vv_common.cuh"
#include "vv_dp_exchange.cuh"

namespace vv {

// Which contraction of the fc7 layer (ref: inner_product_layer.cu:12-59)
enum GemmKind {
  GEMM_FWD = 0,    // D[M,N]  = X[M,K]  . W[N,K]^T     A K-major,  B K-major,  reduce K
  GEMM_WGRAD = 1,  // D[N,K]  = dZ[M,N]^T . X[M,K]     A MN-major, B MN-major, reduce M
  GEMM_DGRAD = 2,  // D[M,K]  = dZ[M,N] . W[N,K]       A K-major,  B MN-major, reduce N
  // gathered wgrad: the same dW computed as (X^T dZ)^T so that the gathered operand is A:
  GEMM_WGRAD_T = 3 // D[N,K]^T = X[M,K]^T . dZ[M,N]    A = X MN-major (gathered), B = dZ MN-major, reduce M, D stored transposed
};

// Epilogue description (device side copy of vv_act_t plus bias / scaling)
struct GemmEpilogue {
  const float* bias;      // per output column, or NULL (FWD only)
  float* Z;               // optional pre-activation output (FWD only)
  int has_act;
  int relu; float negative_slope;
  int dropout_mode; float dropout_scale; uint32_t dropout_thres;
  const uint32_t* mask; uint32_t* mask_out;
  uint64_t seed, step;
  uint32_t hash_base;     // VV_DROPOUT_HASH: dropout_hash_base(seed, step)
  float out_scale;        // multiplies the accumulator (wgrad regularization), 1 otherwise
  // gather-fused forward: Z[m,:] += delta[m] * wlast[:]  (the K-1 copy quirk as a rank-1 correction)
  const float* delta;     // [M] or NULL
  const float* wlast;     // [N] = W[:, K-1]
};

// Split-K finish fused into the weight-gradient kernels (north star: "the momentum update is fused into the wgrad
// epilogue").  Every (tile, split) unit leaves its partial tile in its slab and counts its arrival on the tile's ticket.
// When a CTA has run out of units its epilogue warps DRAIN: the output tiles are cut into chunks that are dealt out over
// all CTAs; a chunk waits for the nsplit arrivals of its tile, re-reads the S partial tiles in slab order -- a fixed order,
// so the result does not depend on which split finished last -- and finishes its elements:
//   mode 1 (one GPU): g = sum_s slab_s (+ the K-1 quirk column); g += decay * W; h = momentum * h + rate * g; W -= h;
//           diff = h; operand copy of W and W[:, K-1] refreshed -- the element-wise arithmetic of sgd_update_tail_kernel
//           (ref: solver.cpp:534-568, net.cpp:837, blob.cpp:126-128); chunks of feature tile 0 also update their rows of
//           the bias.  No update launch.
//   mode 2 (data parallel): the summed rows go straight to their owner rank's receive buffer (vv_dp_exchange.cuh); the CTA
//           finishing the LAST chunk pushes (db, loss, violations) and raises dw_ready on every rank.
struct WgradFinish {
  unsigned int* tickets = nullptr;      // one counter per output tile (+ one for the chunks finished); they only grow:
  unsigned int epoch = 0;               // 1, 2, ... = number of finishing launches on these tickets including this one
  int mode = 0;
  UpdateTail u;                          // mode 1 (u.parts / u.stride / u.nparts are taken from the GEMM problem)
  // mode 2
  int G = 0, rank = 0, rows_per = 0; unsigned int seq = 0;
  const float* col_add = nullptr;       // [N] the quirk column's share (mode 2; mode 1 uses u.col_add)
  const float* small_src = nullptr; int nsmall = 0, small_stride = 0;
  DpPeers peers;
};

// D_rows x D_cols output (row pitch ldd), reduction length red.
// nsplit slabs (split over the reduction) are written at D + s*slab_stride.
struct GemmProblem {
  GemmKind kind;
  int prec;               // vv_precision
  vv_operand_t A, B;      // see GemmKind for which arrays these are
  int M, N, K;            // the fc7 dims (rows, outputs, inputs) -- NOT the tile dims
  float* D; int64_t slab_stride; int nsplit;
  GemmEpilogue epi;
  // gather-fused variants (FWD, WGRAD_T): the X operand (= A) is the resident bank's operand copy and its rows
  // are fetched by index with cp.async by two producer warps; rowmap[m] = bank row of X row m (padded to a multiple of 128)
  const int32_t* rowmap;  // NULL = X is materialised
  int64_t bank_rows;
  int n0 = 0, ncols = 0;  // WGRAD_T only: compute outputs [n0, n0 + ncols) (ncols = 0: all)
  // data parallel: W (operand B of FWD / DGRAD) is being written by the owner ranks' update kernels; the TMA producer
  // lane waits for their w_ready flags right before its first W tile (flags == NULL: nothing to wait for)
  DpWait wait = {nullptr, 0, 0u, nullptr, 0ull};
  const WgradFinish* finish = nullptr;   // WGRAD / WGRAD_T: fused split-K finish (NULL: slabs are left for the caller)
  // FWD: workspace for splitting the units of the last, partial wave of the persistent kernel in two along K (the first
  // half's fp32 partial tile travels through it; NULL: whole units only).  tail_flags: one word per tail CTA, zero at
  // first use; tail_epoch: 1, 2, ... per launch on these flags.
  float* tail_ws = nullptr; size_t tail_ws_bytes = 0; unsigned int* tail_flags = nullptr; size_t tail_flags_count = 0;
  unsigned int tail_epoch = 0;
};
struct FwdTail { float* ws; size_t ws_bytes; unsigned int* flags; size_t flags_count; unsigned int epoch; };

// vv_ip_forward / vv_ip_forward_gathered with the data-parallel wait (wait == NULL: the C-ABI entry points)
int ip_forward_ex(vv_operand_t X, vv_operand_t W, const float* bias, int M, int N, int K, int prec,
                  const vv_act_t* act, float* Z, float* H, const DpWait* wait, vv_stream_t stream, const FwdTail* tail = nullptr);
int ip_forward_gathered_ex(vv_operand_t bank, int64_t bank_rows, const int32_t* rowmap, const float* delta,
                           const float* wlast, vv_operand_t W, const float* bias, int M, int N, int K, int prec,
                           const vv_act_t* act, float* Z, float* H, const DpWait* wait, vv_stream_t stream,
                           const FwdTail* tail = nullptr);
// vv_ip_wgrad / vv_ip_wgrad_gathered_part with the fused split-K finish (finish == NULL: the C-ABI entry points)
int ip_wgrad_ex(vv_operand_t dZ, vv_operand_t X, int M, int N, int K, int prec, float regularization,
                float* dW_parts, int nsplit, void* workspace, size_t workspace_bytes, const WgradFinish* finish, vv_stream_t stream);
int ip_wgrad_gathered_ex(vv_operand_t dZ, vv_operand_t bank, int64_t bank_rows, const int32_t* rowmap, int M, int N,
                         int K, int prec, float regularization, float* dW_parts, int nsplit, int n0, int ncols,
                         const WgradFinish* finish, vv_stream_t stream);
int gemm_tc_launch(const GemmProblem& p, cudaStream_t stream);     // tcgen05 path (TF32X3 / TF32 / BF16)
int gemm_simt_launch(const GemmProblem& p, cudaStream_t stream);   // exact fp32 path
bool gemm_tc_supported(const GemmProblem& p, const char** why);

// ---- shared epilogue element math (used by both paths) ----------------------
// Keep bits (bit k = keep column col + k) of 4 consecutive columns of row `row`, for every dropout mode; writes the
// 0/1 mask to mask_out in the generated modes.  Kept apart from the activation so that callers with a long unrolled
// epilogue can run it in a ROLLED loop first (32 inlined Philox blocks put the promoting forward kernels past the
// instruction cache and cost 30 % of their speed).
__device__ __forceinline__ uint32_t dropout_keep_bits(const GemmEpilogue& e, int N, int row, int col) {
  uint32_t keep[4];
  if (e.dropout_mode == VV_DROPOUT_PHILOX || e.dropout_mode == VV_DROPOUT_HASH) {
    uint32_t w[4];
    if (e.dropout_mode == VV_DROPOUT_PHILOX) dropout_words(e.seed, e.step, uint32_t(row), uint32_t(col >> 2), w);
    else dropout_words_hash(e.hash_base, uint32_t(row), uint32_t(col >> 2), w);
#pragma unroll
    for (int j = 0; j < 4; ++j) keep[j] = (w[j] > e.dropout_thres) ? 1u : 0u;
    if (e.mask_out) {
      *reinterpret_cast<uint4*>(e.mask_out + size_t(row) * N + col) = make_uint4(keep[0], keep[1], keep[2], keep[3]);
    }
  } else {
    const uint4 m = *reinterpret_cast<const uint4*>(e.mask + size_t(row) * N + col);
    if (e.dropout_mode == VV_DROPOUT_MASK_U32) {
      keep[0] = m.x > e.dropout_thres; keep[1] = m.y > e.dropout_thres;
      keep[2] = m.z > e.dropout_thres; keep[3] = m.w > e.dropout_thres;
    } else {
      keep[0] = m.x; keep[1] = m.y; keep[2] = m.z; keep[3] = m.w;
    }
  }
  return (keep[0] & 1u) | ((keep[1] & 1u) << 1) | ((keep[2] & 1u) << 2) | ((keep[3] & 1u) << 3);
}
__device__ __forceinline__ bool epilogue_has_dropout(const GemmEpilogue& e) { return e.has_act && e.dropout_mode != VV_DROPOUT_NONE; }

// Applies the quirk correction, bias, ReLU and dropout (keep bits from dropout_keep_bits) to 4 consecutive columns
// [col, col+4) of row `row`.
__device__ __forceinline__ void epilogue_act4(const GemmEpilogue& e, int N, int row, int col,
                                              float4& v, float4& z_out, uint32_t keep_bits) {
  if (e.delta) {
    const float d = e.delta[row];
    if (d != 0.f) {
      const float4 w = *reinterpret_cast<const float4*>(e.wlast + col);
      v.x = fmaf(d, w.x, v.x); v.y = fmaf(d, w.y, v.y); v.z = fmaf(d, w.z, v.z); v.w = fmaf(d, w.w, v.w);
    }
  }
  if (e.bias) {
    const float4 b = *reinterpret_cast<const float4*>(e.bias + col);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  z_out = v;
  if (!e.has_act) return;
  if (e.relu) {
    const float s = e.negative_slope;
    v.x = fmaxf(v.x, 0.f) + s * fminf(v.x, 0.f);
    v.y = fmaxf(v.y, 0.f) + s * fminf(v.y, 0.f);
    v.z = fmaxf(v.z, 0.f) + s * fminf(v.z, 0.f);
    v.w = fmaxf(v.w, 0.f) + s * fminf(v.w, 0.f);
  }
  if (e.dropout_mode == VV_DROPOUT_NONE) return;
  // reference order: x * mask * scale (dropout_layer.cpp:44)
  v.x = v.x * float(keep_bits & 1u) * e.dropout_scale;
  v.y = v.y * float((keep_bits >> 1) & 1u) * e.dropout_scale;
  v.z = v.z * float((keep_bits >> 2) & 1u) * e.dropout_scale;
  v.w = v.w * float((keep_bits >> 3) & 1u) * e.dropout_scale;
}

}  // namespace vv
