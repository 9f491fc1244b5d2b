/*
 * vv_b200.h -- C-ABI of libvv_b200.so: the B200 (sm_100a) implementation of the
 * temporal-context embedding training hot path of eevignesh/videovector.
 *
 * This is the drop-in boundary.  Every entry point takes plain device pointers,
 * sizes and a CUDA stream, returns an int status (0 = VV_OK, <0 = error, text via
 * vv_last_error()), and replaces the body of one reference operator.  The
 * reference has no FFI of its own (it is one C++ binary); the functions here are
 * what a maintainer would call from the reference's Layer<Dtype>::Forward_gpu /
 * Backward_gpu / SGDSolver::ComputeUpdateValue bodies, with
 * `CHECK_EQ(rc, 0)` standing in for the reference's CUDA_CHECK error convention
 * (include/caffe/util/device_alternate.hpp:48-67).
 *
 * Citations "ref:" are file:line in the reference tree.
 *
 * Layout conventions (all row-major, fp32 unless stated, like Blob<float>,
 * ref: include/caffe/blob.hpp:48-59):
 *   X  [M, K]   bottom of fc7 ("original_feature"), M = R*B rows, row j*B+b is
 *               slot j of batch item b (block-major, the result of SLICE dim 1 +
 *               CONCAT dim 0, ref: slice_layer.cu:22-33, concat_layer.cu:13-20)
 *   W  [N, K]   fc7 weight blob [1,1,N,K] (ref: inner_product_layer.cpp:29)
 *   Z/H [M, N]  ip1_nonorm / ip2
 *   slot order inside an item: 0 target, 1..C-1 context, C..C+Nn-1 negatives
 *               (ref: video_sampled_shots_data_layer.cpp:439-453, 862)
 *
 * There is NO CPU fallback anywhere behind this header: every function either
 * launches sm_100a kernels or returns an error.
 */
#ifndef VV_B200_H_
#define VV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* vv_stream_t; /* == cudaStream_t; 0 = legacy default stream (stock Caffe) */

enum {
  VV_OK = 0,
  VV_ERR_INVALID = -1,     /* bad argument (shape/alignment/null) */
  VV_ERR_CUDA = -2,        /* CUDA runtime / driver error, see vv_last_error() */
  VV_ERR_UNSUPPORTED = -3, /* shape not supported by the sm_100a kernels */
  VV_ERR_NCCL = -4
};

/* How the projection GEMMs compute (the `dtype` of the path).
 *   FP32_SIMT : exact fp32 FMA on CUDA cores (validation / odd shapes)
 *   TF32X3    : split operands x = hi + lo, three products per element pair: hi*hi on the tf32 tensor
 *               pipe, the cross terms bf16(x)*bf16(lo) on the bf16 pipe, fp32 accumulate in TMEM with
 *               periodic promotion to fp32 registers -> fp32-level accuracy (the 1e-5 parity mode)
 *   TF32      : tcgen05 kind::tf32, 1 MMA per product (1e-2 loss-curve mode)
 *   BF16      : tcgen05 kind::f16 with bf16 operands, fp32 accumulate (1e-2 mode)
 *   F16X3     : split operands s*x = h0 + h1 in fp16 (22 significant bits) under a per-tensor power-of-two
 *               scale s, three fp16 products h1*h0' + h0*h1' + h0*h0' on the full-rate f16 pipe, the same
 *               chunked fp32 promotion as TF32X3 -> fp32-level accuracy at 3 MMA units and 4 bytes/element
 *               (TF32X3: 4 units, 8 bytes/element).  The scale follows the tensor's largest magnitude
 *               (vv_operand_rescale); conversions saturate instead of overflowing. */
enum vv_precision {
  VV_PREC_FP32_SIMT = 0,
  VV_PREC_TF32X3 = 1,
  VV_PREC_TF32 = 2,
  VV_PREC_BF16 = 3,
  VV_PREC_F16X3 = 4
};

/* A GEMM operand as the kernels read it.
 *   FP32_SIMT / TF32 : hi = the fp32 array itself, lo = NULL
 *   TF32X3           : hi = round-to-nearest tf32 part (fp32 array, `count` elements); lo = an array of
 *                      the same byte size holding two bf16 planes of `count` elements: bf16(x), bf16(x - hi)
 *   BF16             : hi = bf16 array (uint16 storage), lo = NULL
 *   F16X3            : hi = fp16 plane h0, lo = fp16 plane h1 (`count` elements each), and the
 *                      VV_F16X3_HEADER_BYTES bytes immediately BEFORE hi are the operand's header
 *                      {float scale, float 1/scale, uint32 bits of max|x| seen by the last producer, uint32 layout}:
 *                      allocate header + planes as one block (vv_operand_bytes), zero the header or call
 *                      vv_operand_set_scale once, and call vv_operand_rescale before each re-production.
 * vv_prepare_operand() produces these from an fp32 array; the producer kernels
 * (gather, rank-loss backward, sgd update) can emit them directly. */
#define VV_F16X3_HEADER_BYTES 128
typedef struct vv_operand {
  const void* hi;
  const void* lo;
} vv_operand_t;

/* ReLU + Dropout spec for the fc7 epilogue.
 * ref: relu_layer.cu:9-33 (max(x,0)+slope*min(x,0)), dropout_layer.cu:14-41 */
/* PHILOX: Philox4x32-10 keyed (seed, step, row, col/4); HASH: a 32-bit integer hash per element of the same counters,
 * one third of the arithmetic (what the perf runs use; the reference's curand stream is unspecified either way) */
enum { VV_DROPOUT_NONE = 0, VV_DROPOUT_MASK01 = 1, VV_DROPOUT_MASK_U32 = 2, VV_DROPOUT_PHILOX = 3, VV_DROPOUT_HASH = 4 };
typedef struct vv_act {
  int relu;              /* 0/1 */
  float negative_slope;  /* relu_param.negative_slope */
  int dropout_mode;      /* VV_DROPOUT_* */
  float dropout_ratio;   /* dropout_param.dropout_ratio (threshold_) */
  const uint32_t* mask;  /* [M,N]: MASK01 -> 0/1 (CPU layer's rand_vec_, dropout_layer.cpp:41);
                            MASK_U32 -> raw u32, keep iff mask > uint_thres_ (dropout_layer.cu:19) */
  uint32_t* mask_out;    /* optional [M,N] 0/1 mask written in PHILOX / HASH mode (else NULL) */
  uint64_t seed;         /* PHILOX / HASH: key */
  uint64_t step;         /* PHILOX / HASH: iteration, mixed into the counter */
} vv_act_t;

const char* vv_last_error(void);
int vv_version(void);
/* 0 if the current device is sm_100 and the kernels can launch. */
int vv_device_check(void);

/* ------------------------------------------------------------------------- */
/* K0: data layer + SLICE(dim1) + CONCAT(dim0) + FLATTEN as one index gather. */
/* ref: video_sampled_shots_data_layer.cpp:439-453,492-496,856-864 (what is    */
/* copied where), base_data_layer.cu:7-21, slice_layer.cu:22-33,               */
/* concat_layer.cu:13-20, flatten_layer.cpp:20-24.                             */
/* idx   [B,R] int32 bank row per slot (item-major like the data blob [B,R,K]) */
/* quirk [B,R] int32 or NULL: -2 = full row; >=0 = element K-1 comes from that */
/*       bank row; -1 = element K-1 is 0.0 (the reference's K-1 copy quirk)    */
/* Outputs (any may be NULL): X fp32 [R*B,K] bit-exact; Xop.hi/lo operand       */
/* copies in `prec` layout.  Xblob [B,R,K] = the data blob itself if wanted.    */
/* ------------------------------------------------------------------------- */
int vv_gather_rows(const float* bank, int64_t bank_rows, int K,
                   const int32_t* idx, const int32_t* quirk, int B, int R,
                   float* X, void* Xop_hi, void* Xop_lo, int prec,
                   float* Xblob, vv_stream_t stream);

/* fp32 array -> operand copies for `prec` (no-op for FP32_SIMT/TF32).  F16X3: measures max|src| first and
 * sets the header's scale from it. */
int vv_prepare_operand(const float* src, int64_t count, int prec,
                       void* hi, void* lo, vv_stream_t stream);
/* Operand copy of a resident feature bank [rows, K] for the gather-fused GEMMs (vv_ip_forward_gathered /
 * vv_ip_wgrad_gathered ONLY).  Same allocation as vv_prepare_operand; for F16X3 with K % 64 == 0 the two fp16 planes
 * are interleaved per row in blocks of 64 elements ([64 x h0 | 64 x h1] = 256 contiguous bytes, flagged in the
 * operand header), so that a gathered k-block of a row is one DRAM access instead of two. */
int vv_prepare_bank_operand(const float* bank, int64_t rows, int K, int prec, void* hi, void* lo, vv_stream_t stream);
/* Bytes of one operand allocation for `count` elements and the offsets of hi / lo inside it (lo_offset = 0
 * when the format has no lo array; FP32_SIMT / TF32 read the fp32 array itself: returns 0). */
size_t vv_operand_bytes(int64_t count, int prec, size_t* hi_offset, size_t* lo_offset);
/* F16X3 header maintenance (no-ops for the other precisions).
 *  set_scale : scale = 2^log2_scale, forget the recorded maximum
 *  rescale   : scale = 2^(target_log2 - exponent of the recorded max|x|) when a maximum was recorded (else
 *              unchanged; 1 if never set), then forget the maximum.  target_log2 = 10 leaves a factor 64 of
 *              growth before fp16 saturates.
 *  measure   : record max|src| (use before rescale to initialise from data, e.g. the feature bank) */
int vv_operand_set_scale(void* hi, int prec, int log2_scale, vv_stream_t stream);
int vv_operand_rescale(void* hi, int prec, int target_log2, vv_stream_t stream);
int vv_operand_measure(void* hi, int prec, const float* src, int64_t count, vv_stream_t stream);

/* ------------------------------------------------------------------------- */
/* K1: InnerProduct (fc7).  ref: inner_product_layer.cu:12-25 forward,          */
/* :27-59 backward; CPU restatement inner_product_layer.cpp:61-106.            */
/* ------------------------------------------------------------------------- */
/* Forward: Z = X W^T + 1 b^T ; H = dropout(relu(Z)).  act == NULL -> H = Z.    */
/* Z may be NULL when act != NULL (pre-activation not stored).                  */
int vv_ip_forward(vv_operand_t X, vv_operand_t W, const float* bias,
                  int M, int N, int K, int prec, const vv_act_t* act,
                  float* Z, float* H, vv_stream_t stream);

/* Weight gradient: dW = dZ^T X (overwrite, beta = 0), then dW *= (1+reg/2) if
 * regularization > 0 (ref: inner_product_layer.cpp:80-90).
 * nsplit >= 1 slabs are written at dW_parts + s*N*K (split over M); the caller
 * sums them (vv_sgd_update does, or vv_reduce_parts).  nsplit = 0 lets the
 * library pick and reduce into dW_parts[0] using `workspace`. */
int vv_ip_wgrad(vv_operand_t dZ, vv_operand_t X, int M, int N, int K, int prec,
                float regularization, float* dW_parts, int nsplit,
                void* workspace, size_t workspace_bytes, vv_stream_t stream);
size_t vv_ip_wgrad_workspace_bytes(int M, int N, int K, int prec);
int vv_ip_wgrad_auto_nsplit(int M, int N, int K, int prec);

/* Bias gradient db = dZ^T 1_M (ref: inner_product_layer.cu:48-50). */
int vv_ip_bias_grad(const float* dZ, int M, int N, float* db, vv_stream_t stream);

/* Bottom gradient dX = dZ W (ref: inner_product_layer.cu:52-58). */
int vv_ip_dgrad(vv_operand_t dZ, vv_operand_t W, int M, int N, int K, int prec,
                float* dX, vv_stream_t stream);

/* ---- gather-fused variants: K0 folded into K1 (two producer warps fetch the rows with 16-byte cp.async). ----
 * The X operand is never materialised: `bank` is the operand copy of the whole resident feature bank
 * (vv_prepare_bank_operand, made once; BF16 or F16X3) and rowmap/delta come from vv_gather_plan:
 *   rowmap [round_up(R*B,128)] int32 : bank row of X row j*B+b
 *   delta  [round_up(R*B,128)] fp32  : correction of feature K-1 for rows hit by the K-1 copy quirk
 * forward : Z = bank[rowmap] W^T + delta (x) W[:,K-1] + bias      (wlast [N] = W[:,K-1], fp32)
 * wgrad   : dW = dZ^T bank[rowmap]; the quirk's contribution to dW[:,K-1] is sum_m delta[m] dZ[m,:], which
 *           vv_rank_loss_backward_ex / vv_rank_loss_fused accumulate (dq_accum) and vv_add_column adds; computed as
 *           (X^T dZ)^T so that the gathered operand is again operand A.  2-byte operand formats only (BF16, F16X3). */
int vv_gather_plan(const float* bank, int K, const int32_t* idx, const int32_t* quirk, int B, int R,
                   int32_t* rowmap, float* delta, vv_stream_t stream);
/* Same, with the indices checked against the bank: a row (or quirk row) outside [0, bank_rows) is planned as row 0 and
 * *bad_flag (device-visible word, e.g. host-mapped; may be NULL) gets bit 0 set -- vv_trainer_step uses this and reports
 * the error on the next call.  bank_rows = 0 skips the check. */
int vv_gather_plan_checked(const float* bank, int64_t bank_rows, int K, const int32_t* idx, const int32_t* quirk, int B, int R,
                           int32_t* rowmap, float* delta, uint32_t* bad_flag, vv_stream_t stream);
int vv_ip_forward_gathered(vv_operand_t bank, int64_t bank_rows, const int32_t* rowmap, const float* delta,
                           const float* wlast, vv_operand_t W, const float* bias, int M, int N, int K, int prec,
                           const vv_act_t* act, float* Z, float* H, vv_stream_t stream);
int vv_ip_wgrad_gathered(vv_operand_t dZ, vv_operand_t bank, int64_t bank_rows, const int32_t* rowmap,
                         int M, int N, int K, int prec, float regularization, float* dW_parts, int nsplit,
                         vv_stream_t stream);
/* the rows [n0, n0 + ncols) of dW only (n0 % 8 == 0), written in place inside the [nsplit, N, K] slabs: lets a
 * data-parallel caller all-reduce one slice while the next one is computed */
int vv_ip_wgrad_gathered_part(vv_operand_t dZ, vv_operand_t bank, int64_t bank_rows, const int32_t* rowmap,
                              int M, int N, int K, int prec, float regularization, float* dW_parts, int nsplit,
                              int n0, int ncols, vv_stream_t stream);
/* A[i*ld + col] += v[i], i < n */
int vv_add_column(float* A, int64_t ld, int col, const float* v, int n, vv_stream_t stream);

/* out[i] = sum_s parts[s*stride + i] */
int vv_reduce_parts(const float* parts, int nparts, int64_t stride, int64_t count,
                    float* out, vv_stream_t stream);

/* ------------------------------------------------------------------------- */
/* K2/K3: slice_emb .. max_margin_loss fused (context mean, 3 L2 normalisations,*/
/* 1+Nn dot products, hinge loss; and the whole backward down to ip1.diff).     */
/* ref: eltwise_layer.cpp:67-73,119-143; normalization_layer.cpp:30-112;       */
/* sum_layer.cpp:32-82; split_layer.cpp:36-51; max_margin_loss_layer.cpp:54-214;*/
/* dropout_layer.cpp:52-68; relu_layer.cpp:23-36; slice/concat as indexing.     */
/* ------------------------------------------------------------------------- */
#define VV_MAX_CONTEXT 64
typedef struct vv_rank_cfg {
  int B, C, Nn, N;       /* items, context_size (incl. target), negatives, embedding dim */
  float coeff[VV_MAX_CONTEXT]; /* eltwise_param.coeff for the C-1 context rows */
  float margin;          /* max_margin_loss_param.margin */
  int norm;              /* 1 = L1, 2 = L2 */
  float eps;             /* 1e-10 (normalization_layer.cpp:36) */
} vv_rank_cfg_t;

/* floats per item in the `stats` buffer: s_c, then (s_x, p_x) for target, neg_1..Nn */
static inline int vv_rank_stats_stride(int Nn) { return 1 + 2 * (1 + Nn); }

/* Forward.  H [R*B, N] = ip2 (block-major).  Outputs:
 *  stats        [B, vv_rank_stats_stride(Nn)]  saved for backward
 *  target_score [B, Nn] (sum_true with num_output=Nn), neg_score [B, Nn]  (NULL ok)
 *  item_loss [B], item_viol [B] partial sums (NULL ok) ; loss[0] = mean hinge term,
 *  violations[0] = #(s+ - s- < 0)  (both device pointers, NULL ok) */
int vv_rank_loss_forward(const float* H, const vv_rank_cfg_t* cfg, float* stats,
                         float* target_score, float* neg_score,
                         float* item_loss, float* item_viol,
                         float* loss, float* violations, vv_stream_t stream);

/* Backward from saved stats to ip1_nonorm.diff:
 *  dZ = d(ip2) * dropout_scale * [ip2 > 0]  when act_fused != 0 (ReLU slope 0 and
 *  dropout folded: ip2 > 0 <=> Z > 0 and mask = 1), else dZ = d(ip2) (then the
 *  caller runs vv_dropout_backward / vv_relu_backward).
 *  Any of dZ (fp32), dZop.hi/lo (operand copies for `prec`), db_accum may be NULL.
 *  db_accum [N] is atomically accumulated with the column sums of dZ (zero it first). */
int vv_rank_loss_backward(const float* H, const vv_rank_cfg_t* cfg, const float* stats,
                          float loss_weight, int act_fused, float dropout_scale,
                          float* dZ, void* dZop_hi, void* dZop_lo, int prec,
                          float* db_accum, vv_stream_t stream);
/* same, plus dq_accum [N] += sum_m delta[m] * dZ[m,:] (zero it first; see the gather-fused wgrad) */
int vv_rank_loss_backward_ex(const float* H, const vv_rank_cfg_t* cfg, const float* stats,
                             float loss_weight, int act_fused, float dropout_scale,
                             float* dZ, void* dZop_hi, void* dZop_lo, int prec,
                             float* db_accum, const float* delta, float* dq_accum, vv_stream_t stream);

/* K2 + K3 in one pass over H (the rows of an item stay in registers between the forward reductions and the
 * gradient rows): same outputs as vv_rank_loss_forward followed by vv_rank_loss_backward_ex (to rounding); every
 * forward output may be NULL.  Supported when vv_rank_loss_fused_supported(cfg) (N <= 1024, C + Nn <= 32). */
int vv_rank_loss_fused_supported(const vv_rank_cfg_t* cfg);
int vv_rank_loss_fused(const float* H, const vv_rank_cfg_t* cfg, float loss_weight, int act_fused,
                       float dropout_scale, float* stats, float* target_score, float* neg_score,
                       float* item_loss, float* item_viol, float* loss, float* violations,
                       float* dZ, void* dZop_hi, void* dZop_lo, int prec, float* db_accum,
                       const float* delta, float* dq_accum, vv_stream_t stream);

/* ------------------------------------------------------------------------- */
/* K4: SGDSolver::ComputeUpdateValue + Net::Update + Blob::Update in one pass.  */
/* ref: solver.cpp:486-576, net.cpp:804-839, blob.cpp:113-136.                 */
/*  g = grad_scale * sum_s grad_parts[s]; g += decay*W (L2) | decay*sign(W) (L1)*/
/*  hist = momentum*hist + local_rate*g ; diff_out = hist ; W -= hist           */
/* diff_out may alias grad_parts (slab 0) or be NULL.  Wop_hi/lo: refreshed      */
/* operand copies of W for `prec` (NULL ok).                                     */
/* ------------------------------------------------------------------------- */
int vv_sgd_update(float* W, const float* grad_parts, int nparts, int64_t part_stride,
                  float* hist, float* diff_out, int64_t count,
                  float local_rate, float momentum, float local_decay, int reg_type,
                  float grad_scale, void* Wop_hi, void* Wop_lo, int prec,
                  vv_stream_t stream);
/* rate = base_lr * f(iter)  (ref: solver.cpp:441-460), evaluated in float. */
float vv_learning_rate(const char* policy, float base_lr, float gamma, float power,
                       int stepsize, int iter);

/* ------------------------------------------------------------------------- */
/* Standalone layer kernels: exact single-layer semantics for the drop-in Layer */
/* classes when a net is not fused (ref files as named).                        */
/* ------------------------------------------------------------------------- */
int vv_relu_forward(const float* x, int64_t n, float negative_slope, float* y, vv_stream_t s);      /* relu_layer.cu:9-33 */
int vv_relu_backward(const float* x, const float* dy, int64_t n, float negative_slope, float* dx, vv_stream_t s); /* :35-60 */
int vv_dropout_forward(const float* x, const uint32_t* mask, int mask_mode, int64_t n, float ratio, float* y, vv_stream_t s); /* dropout_layer.cu:14-41 */
int vv_dropout_backward(const float* dy, const uint32_t* mask, int mask_mode, int64_t n, float ratio, float* dx, vv_stream_t s); /* :44-70 */
/* 0/1 mask [rows, cols] drawn from the same Philox stream the fused fc7 epilogue uses (VV_DROPOUT_PHILOX) */
int vv_dropout_make_mask(uint32_t* mask01, int rows, int cols, float ratio, uint64_t seed, uint64_t step, vv_stream_t s);
/* same for either generated stream (mode = VV_DROPOUT_PHILOX | VV_DROPOUT_HASH) */
int vv_dropout_make_mask_mode(uint32_t* mask01, int rows, int cols, float ratio, uint64_t seed, uint64_t step, int mode, vv_stream_t s);
int vv_eltwise_sum_forward(const float* const* bottoms, const float* coeffs, int nb, int64_t n, float* top, vv_stream_t s); /* eltwise_layer.cu:48-54; host arrays of device ptrs */
int vv_eltwise_prod_forward(const float* a, const float* b, int64_t n, float* top, vv_stream_t s);  /* eltwise_layer.cu:41-47 */
int vv_axpby(int64_t n, float alpha, const float* x, float beta, float* y, vv_stream_t s);         /* y = alpha*x + beta*y */
int vv_sign_axpy(int64_t n, float alpha, const float* x, float* y, vv_stream_t s);   /* y += alpha*sign(x): L1 decay, solver.cpp:547-554 */
int vv_mul(int64_t n, const float* a, const float* b, float* y, vv_stream_t s);
int vv_l2norm_forward(const float* x, int num, int dim, float* y, vv_stream_t s);                   /* normalization_layer.cu:10-45 */
int vv_l2norm_backward(const float* x, const float* dy, int num, int dim, float* dx, vv_stream_t s);/* normalization_layer.cu:47-97 */
int vv_rowsum_forward(const float* x, int num, int dim, int num_output, float* y, vv_stream_t s);   /* sum_layer.cu:10-31 */
int vv_rowsum_backward(const float* dy, int num, int dim, int num_output, float* dx, vv_stream_t s);/* sum_layer.cu:33-55 */
int vv_copy_strided(const float* src, int64_t src_stride, float* dst, int64_t dst_stride,
                    int64_t rows, int64_t cols, vv_stream_t s);                                     /* slice/concat copies */
int vv_max_margin_forward(const float* s_true, const float* s_bogus, int count, float margin, int norm,
                          float* hinge_tmp, float* loss, float* violations, vv_stream_t s);         /* max_margin_loss_layer.cpp:54-127 */
int vv_max_margin_backward(const float* s_true, const float* s_bogus, int count, float margin, int norm,
                           float loss_weight, float* d_true, float* d_bogus, vv_stream_t s);        /* :130-214 */
/* Per-element loss weights (the layer's optional third bottom, max_margin_loss_layer.cpp:18-39,79-97,150-186):
 * `weights` [count] on the device, NULL = unweighted.  Forward term: sqrt(w)*h (L2) or w*h (L1); backward: w*h, and
 * for L1 the sign step becomes w -- the reference's own asymmetry, kept.  With use_direct_weight the blob IS `weights`;
 * otherwise it holds video ids and vv_id_to_weight maps them through the id_to_weight_file table (ids ascending; an id
 * that is not in the table weighs 0, as std::map::operator[] gives the reference). */
int vv_max_margin_forward_w(const float* s_true, const float* s_bogus, const float* weights, int count, float margin, int norm,
                            float* hinge_tmp, float* loss, float* violations, vv_stream_t s);
int vv_max_margin_backward_w(const float* s_true, const float* s_bogus, const float* weights, int count, float margin, int norm,
                             float loss_weight, float* d_true, float* d_bogus, vv_stream_t s);
int vv_id_to_weight(const float* video_ids, int count, const int* table_ids, const float* table_weights, int table_size,
                    float* weights, vv_stream_t s);

/* ------------------------------------------------------------------------- */
/* Synthetic feature bank (benchmarks): value(row, col) = relu(approx N(0,1))   */
/* from an integer hash of (seed,row,col); bit-identical on host and device.    */
/* ------------------------------------------------------------------------- */
int vv_fill_bank(float* bank, int64_t rows, int K, uint64_t seed, vv_stream_t stream);
float vv_bank_value_host(uint64_t seed, int64_t row, int col, int K);

/* ------------------------------------------------------------------------- */
/* Record reader: the DB values the reference's data layers  */
/* parse -> a dense host feature bank [rows,K] + the tables the sampler takes.  */
/*   VV_RECORD_VIDEO_SHOTS : video_shot_sentences.VideoShots                    */
/*     (video_shot_sentences.proto:15-20), as VideoSampledShotsDataLayer reads  */
/*     it (video_sampled_shots_data_layer.cpp:184-199,301-314,789-846): one     */
/*     record per video, one bank row per shot_words datum (float_data only,    */
/*     feature_size = float_data_size of the first datum), shot_ids alongside.  */
/*   VV_RECORD_TEST_WINDOWS: video_shot_sentences.TestVideoShotWindows          */
/*     (:22-30) as VideoShotWindowTestDataLayer reads it                        */
/*     (video_shot_window_test_data_layer.cpp:95-114,186-239): per record the   */
/*     context, then (include_positives) positive, then (include_negatives)     */
/*     negative datums as consecutive rows = one item of the data blob; the     */
/*     label is video_id.  Sizes are fixed by the first record, later records   */
/*     must agree (the reference's CHECK_EQs).                                  */
/* Records are added in DB key order, i.e. the order of the reference's         */
/* mdb_cursor_get(MDB_FIRST/MDB_NEXT) / leveldb iterator loop: either by the    */
/* caller from its own cursor (vv_record_set_add; the binding in INTEGRATION.md)*/
/* or from a file: a "VVRS0001" stream (per record: u32 key_len, key, u64       */
/* value_len, value; little endian) or the text `mdb_dump [-p]` prints.         */
/* Protobuf wire format decoded directly: packed and unpacked repeated scalars, */
/* unknown fields skipped, malformed input -> VV_ERR_INVALID + vv_last_error(). */
/* ------------------------------------------------------------------------- */
enum { VV_RECORD_VIDEO_SHOTS = 0, VV_RECORD_TEST_WINDOWS = 1 };
typedef struct vv_record_set vv_record_set_t;
vv_record_set_t* vv_record_set_create(int kind, int include_positives, int include_negatives);
void vv_record_set_destroy(vv_record_set_t* s);
int vv_record_set_add(vv_record_set_t* s, const void* value, size_t size);
int vv_record_set_load_file(vv_record_set_t* s, const char* path);
/* rows_per_record: rows of one TEST item (0 for VIDEO_SHOTS, where it varies: see row_off) */
int vv_record_set_info(const vv_record_set_t* s, int64_t* records, int64_t* rows, int32_t* feature_size,
                       int32_t* rows_per_record);
/* video_id [records], row_off [records+1] (= the sampler's shot_off), shot_ids [rows] (TEST: the positive /
 * negative shot ids, -1 for context rows); any may be NULL */
int vv_record_set_tables(const vv_record_set_t* s, int32_t* video_id, int32_t* row_off, int32_t* shot_ids);
const float* vv_record_set_bank(const vv_record_set_t* s);          /* host [rows, feature_size], owned by the set */
int vv_record_set_upload(const vv_record_set_t* s, float* bank_dev, vv_stream_t stream);   /* -> device bank */

/* ------------------------------------------------------------------------- */
/* Host sampler: VideoSampledShotsDataLayer's sampler as an index stream        */
/* (ref: video_sampled_shots_data_layer.cpp:25-44,245-344,372-507,769-909;     */
/* util/rng.hpp:43-54).  Self-contained glibc-compatible rand() (TYPE_3, the    */
/* reference never seeds it -> seed 1), so it is reproducible per rank and does */
/* not touch process-global state.  No CUDA inside.                             */
/* ------------------------------------------------------------------------- */
typedef struct vv_sampler vv_sampler_t;
vv_sampler_t* vv_sampler_create(int num_videos, const int32_t* video_id,
                                const int32_t* shot_off /*[V+1]*/, const int32_t* shot_ids,
                                int batch_size, int context_size, int num_negative_samples,
                                int max_buffer_size, int negative_swap_percentage,
                                int max_same_video_negs, int max_tries_for_negs,
                                unsigned int rand_seed);
/* context_type = VideoSampledShotsDataParameter.CONTEXT (caffe.proto:598-604; vv_sampler_create = WINDOW):
 *   PAIRWISE (context_size must be 2): two distinct shots of the video (:396-404)
 *   WINDOW   : sorted random window, target = its median (:425-506)
 *   PAST     : sorted random window, target = its last shot; same-video negatives from before the window (:509-583)
 *   PAST_CONTINUOUS[_FIXED]: evenly spaced frames with a random (fixed) stride, target = the last; negatives = the
 *              frames right before the window (:586-757) */
enum { VV_CONTEXT_PAIRWISE = 0, VV_CONTEXT_WINDOW = 1, VV_CONTEXT_PAST = 2, VV_CONTEXT_PAST_CONTINUOUS = 3,
       VV_CONTEXT_PAST_CONTINUOUS_FIXED = 4 };
vv_sampler_t* vv_sampler_create_ex(int num_videos, const int32_t* video_id,
                                   const int32_t* shot_off /*[V+1]*/, const int32_t* shot_ids,
                                   int batch_size, int context_size, int num_negative_samples,
                                   int max_buffer_size, int negative_swap_percentage,
                                   int max_same_video_negs, int max_tries_for_negs,
                                   unsigned int rand_seed, int context_type);
/* The data layer's two remaining options (video_sampled_shots_data_layer.cpp:104-153, 157-180, 253-284, 324-338):
 *   start_skip  = rand_skip: the record cursor starts `start_skip` records in (wrapping).  The reference draws it as
 *                 caffe_rng_rand() % rand_skip -- the first output of an mt19937 seeded with the Caffe seed -- the caller
 *                 passes the value (the compat data layer computes it the same way);
 *   neg_*       = negative_dataset: a second set of VideoShots records (same table layout) whose rows live at
 *                 [neg_row_base, ...) of the resident bank.  The negative buffer then starts with EVERY shot of these
 *                 records in order -- no rand() is drawn and the main cursor does not move -- and, as in the reference
 *                 (unchecked copy + CHECK_EQ :346), max_buffer_size must be reached exactly at a record boundary
 *                 (NULL otherwise).  Swaps during training still come from the main data.  neg_num_videos = 0: none. */
vv_sampler_t* vv_sampler_create_ex2(int num_videos, const int32_t* video_id,
                                    const int32_t* shot_off /*[V+1]*/, const int32_t* shot_ids,
                                    int batch_size, int context_size, int num_negative_samples,
                                    int max_buffer_size, int negative_swap_percentage,
                                    int max_same_video_negs, int max_tries_for_negs,
                                    unsigned int rand_seed, int context_type, int start_skip,
                                    int neg_num_videos, const int32_t* neg_video_id,
                                    const int32_t* neg_shot_off, const int32_t* neg_shot_ids, int32_t neg_row_base);
void vv_sampler_destroy(vv_sampler_t* s);
/* idx, quirk: host [B,R] int32 (see vv_gather_rows).  Returns 0 or <0. */
int vv_sampler_next(vv_sampler_t* s, int32_t* idx, int32_t* quirk);
int vv_sampler_cursor(const vv_sampler_t* s);
/* Prefetch thread, the reference's BasePrefetchingDataLayer/InternalThread (base_data_layer.cpp:53-95): a producer
 * thread draws up to `depth` batches ahead, vv_sampler_next hands them out in order (same stream as without it).
 * depth <= 0 stops the thread; batches already drawn are still served first. */
int vv_sampler_prefetch(vv_sampler_t* s, int depth);
int vv_sampler_prefetch_ready(vv_sampler_t* s);   /* batches drawn ahead and not yet handed out */
/* A sampler built over a SUB-SHARD of the videos (its shot tables start at 0) whose rows live at [row_base, ...) of a
 * larger resident bank: every emitted index (idx, and quirk where >= 0) is offset by row_base.  Lets k samplers over k
 * disjoint sub-shards -- each reference-exact on its own videos, each with its own prefetch thread -- feed one GPU. */
int vv_sampler_set_row_base(vv_sampler_t* s, int32_t row_base);
/* the generator alone, for tests against libc rand() */
typedef struct vv_glibc_rand vv_glibc_rand_t;
vv_glibc_rand_t* vv_glibc_rand_create(unsigned int seed);
int vv_glibc_rand_next(vv_glibc_rand_t* g);
void vv_glibc_rand_destroy(vv_glibc_rand_t* g);

/* ------------------------------------------------------------------------- */
/* TEST-phase evaluation (ref: mednet_embedding_train.prototxt TEST graph;      */
/* retrieval_stats_layer.cpp:98-140 ComputeStats, :143-359 Forward_cpu).        */
/* ------------------------------------------------------------------------- */
/* "average_for_test": Xbar[b,:] = sum_f coeff[f] * bank[idx[b,f],:] (coeff host array or NULL = 1/F), idx [B,F]. */
int vv_gather_mean_rows(const float* bank, int64_t bank_rows, int K, const int32_t* idx, int B, int F,
                        const float* coeff_host, float* Xbar, vv_stream_t stream);
/* RetrievalStatsLayer, shot level: distances -2 E E^T over the B L2-normalised embeddings E [B,N], every query
 * ranked against all others (the query itself excluded; with exclude_same_video_shots also its video's shots),
 * relevance = same label; queries with labels[i] < 0 are not scored.
 *   out3 (device) = {mean AP, hit@1, hit@5}; per_query (device, [B,3], optional) = per-query values (-1 if unscored).
 *   gram_given: optional precomputed E E^T [B,B] (then E may be NULL); workspace: vv_retrieval_stats_workspace_bytes(B). */
size_t vv_retrieval_stats_workspace_bytes(int B);
int vv_retrieval_stats(const float* E, int B, int N, const int32_t* video_ids, const int32_t* labels,
                       int exclude_same_video_shots, const float* gram_given, void* workspace, size_t workspace_bytes,
                       double* out3, double* per_query, vv_stream_t stream);
/* + top5 (device, [B,5] int32, optional): per query the first five ranked items of ANOTHER video -- the "ret_id_1..5" of
 * the layer's per-query CSV (stats_output_file, retrieval_stats_layer.cpp:306-333); -1 where fewer exist / unscored. */
int vv_retrieval_stats_ex(const float* E, int B, int N, const int32_t* video_ids, const int32_t* labels,
                          int exclude_same_video_shots, const float* gram_given, void* workspace, size_t workspace_bytes,
                          double* out3, double* per_query, int32_t* top5, vv_stream_t stream);
/* video_level_retrieval (:160-206): out[v,:] = mean of the batch's embeddings E[i,:] with group[i] == v (group [B] int32 in
 * [0,V)), accumulated in increasing i; the stats then run on out [V,N] with one id / label per video. */
int vv_video_mean_rows(const float* E, int B, int N, const int32_t* group, int V, float* out, vv_stream_t stream);

/* IdToWeightMapping: a per-id embedding table (ref: id_to_weight_mapping_layer.cpp:61-148; CPU loops in the reference).
 *  forward : top[i,:] = table[ids[i],:]        ids [M] floats holding integers (a Caffe blob), table [rows, N]
 *  backward: table_diff = 0, then += top_diff[i,:] into row ids[i] in increasing i (deterministic, the reference's order) */
int vv_id_lookup_forward(const float* table, int rows, int N, const float* ids, int M, float* top, vv_stream_t stream);
int vv_id_lookup_backward(const float* top_diff, const float* ids, int M, int N, int rows, float* table_diff, vv_stream_t stream);

/* ------------------------------------------------------------------------- */
/* Trainer: one data-parallel rank of the fused training step                   */
/* (gather -> fc7 fwd(+relu+dropout) -> rank loss fwd/bwd -> wgrad -> [allreduce]*/
/*  -> sgd update), i.e. Solver::Solve's loop body (ref: solver.cpp:177-220)    */
/* for the shipped net.  One process per GPU; NCCL is loaded at run time.       */
/* ------------------------------------------------------------------------- */
typedef struct vv_trainer vv_trainer_t;
typedef struct vv_trainer_cfg {
  int B;                 /* items on THIS rank per step */
  int C, Nn, K, N;
  float coeff[VV_MAX_CONTEXT]; /* zeros -> 1/(C-1) */
  float margin; int norm;
  float dropout_ratio;   /* 0 -> no dropout layer */
  int dropout_mode;      /* VV_DROPOUT_HASH (or PHILOX) for perf runs, MASK01 for parity */
  uint64_t dropout_seed;
  float loss_weight;
  float regularization;  /* inner_product_param.regularization */
  /* solver (ref: mednet_embedding_train_solver.prototxt) */
  char lr_policy[16]; float base_lr, gamma, power; int stepsize;
  float momentum, weight_decay; int reg_type;  /* 2 = L2, 1 = L1 */
  float lr_mult[2], decay_mult[2];             /* weight, bias: blobs_lr / weight_decay */
  int prec;              /* vv_precision */
  int world_size, rank;  /* data parallel */
  int compute_dgrad;     /* 0 in the shipped net (net.cpp:68-76) */
  int keep_blobs;        /* 1: also keep fp32 X / Z / dZ for inspection (parity tests) */
  int split_rank_loss;   /* 1: run K2 and K3 as two kernels even where the fused one applies (A/B timing, tests) */
} vv_trainer_cfg_t;

vv_trainer_t* vv_trainer_create(const vv_trainer_cfg_t* cfg, vv_stream_t stream);
void vv_trainer_destroy(vv_trainer_t* t);
/* device pointers owned by the trainer (valid until destroy) */
float* vv_trainer_weight(vv_trainer_t* t);       /* [N,K] */
float* vv_trainer_bias(vv_trainer_t* t);         /* [N]   */
float* vv_trainer_weight_hist(vv_trainer_t* t);
float* vv_trainer_bias_hist(vv_trainer_t* t);
float* vv_trainer_weight_diff(vv_trainer_t* t);  /* dW after the step: = hist (reference semantics) */
float* vv_trainer_bias_diff(vv_trainer_t* t);
float* vv_trainer_blob(vv_trainer_t* t, const char* name); /* "X","Z","H","dZ","stats","loss","violations","dW_raw","db_raw" */
/* Register the resident feature bank: builds its operand copies once (tf32x3: hi+lo, bf16: bf16) so that
 * vv_trainer_step on this bank runs the gather-fused GEMMs (no materialised X).  Not used in keep_blobs mode. */
int vv_trainer_set_bank(vv_trainer_t* t, const float* bank, int64_t bank_rows);
/* call after writing weights/bias through the pointers above */
int vv_trainer_sync_weights(vv_trainer_t* t);
/* One step.  bank: device feature bank; idx/quirk: DEVICE [B,R] int32 for this
 * rank's items; mask: explicit dropout mask (MASK01 mode) or NULL; iter: solver
 * iteration (0-based) used for the learning rate and the Philox stream.
 * do_update = 0 runs forward/backward only (gradients left in *_diff raw). */
int vv_trainer_step(vv_trainer_t* t, const float* bank, int64_t bank_rows,
                    const int32_t* idx, const int32_t* quirk, const uint32_t* mask,
                    int iter, int do_update);
/* kernels launched by the last step (for bench.py's gpu_launches claim) */
int vv_trainer_last_launches(const vv_trainer_t* t);
/* Per-phase device timing with CUDA events on the trainer's stream (off by default).
 * Phases: 0 gather, 1 fc7 forward, 2 rank-loss forward, 3 rank-loss backward, 4 wgrad,
 * 5 dgrad, 6 allreduce (+slab reduce; NCCL mode only), 7 sgd update (peer-memory mode: the whole exchange + update kernel).  vv_trainer_phase_ms synchronises the
 * stream, writes the summed milliseconds per phase and the number of timed steps, and resets. */
#define VV_NUM_PHASES 8
int vv_trainer_set_timing(vv_trainer_t* t, int enable);
int vv_trainer_phase_ms(vv_trainer_t* t, float* ms_out /*[VV_NUM_PHASES]*/, int* steps_out);
/* Inference / extraction: out[rows,N] = relu(F W^T + b) (ref: tools/extract_features.cpp:100-209
 * reading blob ip2 of videovec_extraction.prototxt:179-205). rows_op = operand copies of F rows. */
int vv_trainer_extract(vv_trainer_t* t, const float* F, int64_t rows, float* out);

/* Data-parallel plumbing (ref has none).  vv_dp_init creates the NCCL communicator and then maps every
 * rank's exchange buffers into every other rank (CUDA IPC over NVLink).  With the mapping in place (vv_dp_mode == 2)
 * a training step exchanges its gradients INSIDE the update kernel: split-K sum -> rows pushed to their owner rank ->
 * the owner adds the G contributions in rank order, updates its N/G rows of W and of the history, and pushes the new
 * rows (fp32 master optional, GEMM operand copy, W[:,K-1]) to every rank; the next forward GEMM waits for them in its
 * TMA producer.  Replicas are bit-identical by construction.  The optimiser state is then SHARDED: rank r holds the
 * current history (and, unless VV_DP_REPLICATE_MASTER=1 or the precision uses fp32 operands, the current fp32 master
 * weights) only for rows [r*N/G, (r+1)*N/G); vv_dp_gather_state (collective: every rank calls it) all-gathers them so
 * that vv_trainer_weight / _weight_hist / _weight_diff hold the whole blobs, e.g. before a snapshot.
 * vv_dp_mode == 1: NCCL all-reduce between wgrad and the update (VV_DP_MODE=nccl, no peer access, more than 8
 * ranks, or N not divisible by the world size; vv_dp_mode_reason says which).  In either mode the loss blob holds the
 * mean over ranks (the global-batch mean) and violations the global count, with or without an update. */
int vv_dp_unique_id(void* id128 /*128 bytes out*/);
int vv_dp_init(vv_trainer_t* t, const void* id128);
int vv_dp_mode(const vv_trainer_t* t);                 /* 0 = single rank, 1 = NCCL all-reduce, 2 = peer-memory exchange */
const char* vv_dp_mode_reason(const vv_trainer_t* t);  /* why mode 1 was chosen ("" otherwise) */
int vv_dp_gather_state(vv_trainer_t* t);
int vv_dp_allreduce_inplace(vv_trainer_t* t, float* buf, int64_t count, vv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VV_B200_H_ */
