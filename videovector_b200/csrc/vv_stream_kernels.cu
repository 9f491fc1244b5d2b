// vv_stream_kernels.cu -- the HBM-bound streaming kernels around the GEMMs:
//   K0 index gather (data layer + slice/concat/flatten), operand preparation,
//   K4 fused SGD/momentum update, split-K slab reduction, bias gradient,
//   synthetic feature bank, and the standalone single-layer kernels that give the
//   drop-in Layer classes exact per-layer semantics when a net is not fused.
// All are 128-bit vectorised, grid-stride, sized in multiples of the SM count.
#include "vv_common.cuh"
#include <string.h>

namespace vv {
namespace {

inline int stream_grid(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return int(blocks);
}

// ----------------------------------------------------------------------------
// K0 gather.  One CTA per output row (grid-stride).  Output row j*B+b <- bank[idx[b*R+j]].
// quirk: element K-1 replaced (ref: video_sampled_shots_data_layer.cpp:492 copies K-1 floats)
// ----------------------------------------------------------------------------
struct GatherOut { float* X; float* hi; float* lo; uint16_t* bf; float* blob; int prec; };

constexpr int kGatherUnroll = 4;   // 128-bit loads in flight per thread before the first store

__device__ __forceinline__ void gather_store(const GatherOut& o, size_t total, size_t off, size_t blob_off, const float4& v,
                                             float scale, float& amax) {
  if (o.X) stg_stream(reinterpret_cast<float4*>(o.X + off), v);
  if (o.blob) stg_stream(reinterpret_cast<float4*>(o.blob + blob_off), v);
  if (o.prec == VV_PREC_TF32X3) {
    store_x3(o.hi, o.lo, total, off, v);
  } else if (o.prec == VV_PREC_F16X3) {
    store_f16x3(o.hi, o.lo, off, v, scale, amax);
  } else if (o.prec == VV_PREC_BF16) {
    *reinterpret_cast<uint2*>(o.bf + off) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ bank, const int* __restrict__ idx, const int* __restrict__ quirk,
                   int B, int R, int K, const GatherOut o) {
  const int K4 = K >> 2;
  const long long M = (long long)B * R;
  const int T = blockDim.x;
  const float scale = (o.prec == VV_PREC_F16X3) ? f16_hdr(o.hi)->scale : 1.f;
  float amax = 0.f;
  // the row index of the NEXT row is fetched while this row streams (it heads a dependent chain of two DRAM trips)
  long long orow = blockIdx.x;
  int slot = 0; long long src = 0; int qk = -2;
  if (orow < M) {
    const int j = int(orow / B), b = int(orow - (long long)j * B);
    slot = b * R + j; src = idx[slot]; qk = quirk ? quirk[slot] : -2;
  }
  for (; orow < M; orow += gridDim.x) {
    const int cur_slot = slot; const int cur_qk = qk;
    const float4* s4 = reinterpret_cast<const float4*>(bank + src * K);
    const long long nrow = orow + gridDim.x;
    if (nrow < M) {
      const int j = int(nrow / B), b = int(nrow - (long long)j * B);
      slot = b * R + j; src = idx[slot]; qk = quirk ? quirk[slot] : -2;
    }
    // the K-1 copy quirk touches one element of the row: fetch its replacement once, up front
    float last = 0.f;
    if (cur_qk >= 0 && threadIdx.x == ((K4 - 1) % T)) last = bank[(long long)cur_qk * K + (K - 1)];
    for (int c0 = 0; c0 < K4; c0 += T * kGatherUnroll) {
      float4 v[kGatherUnroll];
#pragma unroll
      for (int u = 0; u < kGatherUnroll; ++u) {           // all loads of the batch are issued before any store
        const int c = c0 + u * T + threadIdx.x;
        if (c < K4) v[u] = ldg_stream(s4 + c);
      }
#pragma unroll
      for (int u = 0; u < kGatherUnroll; ++u) {
        const int c = c0 + u * T + threadIdx.x;
        if (c < K4) {
          if (c == K4 - 1 && cur_qk != -2) v[u].w = last;
          gather_store(o, size_t(M) * K, size_t(orow) * K + size_t(c) * 4, size_t(cur_slot) * K + size_t(c) * 4, v[u],
                       scale, amax);
        }
      }
    }
  }
  if (o.prec == VV_PREC_F16X3) f16_publish_absmax(o.hi, amax);
}

// Gather plan for the gather-fused GEMMs: rowmap[m] = bank row of X row m = j*B+b (0 in the padding), and
// delta[m] = (value element K-1 should have) - (value the bank row has there): the K-1 copy quirk
// (ref: video_sampled_shots_data_layer.cpp:492) expressed as a correction of the last feature only.
// bank_rows > 0: indices are checked -- a row (or quirk row) outside [0, bank_rows) is replaced by row 0 and reported
// through *bad (a mismatched bank / sampler pair must not become a silent out-of-bounds read in the GEMM producers)
__global__ void __launch_bounds__(256)
gather_plan_kernel(const float* __restrict__ bank, const int* __restrict__ idx, const int* __restrict__ quirk,
                   int B, int R, int K, int Mpad, int* __restrict__ rowmap, float* __restrict__ delta,
                   long long bank_rows, unsigned int* __restrict__ bad) {
  const int M = B * R;
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < Mpad; m += gridDim.x * blockDim.x) {
    int row = 0; float d = 0.f;
    if (m < M) {
      const int j = m / B, b = m - j * B;
      const int slot = b * R + j;
      row = idx[slot];
      int qk = quirk ? quirk[slot] : -2;
      if (bank_rows > 0 && (row < 0 || row >= bank_rows || qk >= bank_rows || qk < -2)) {
        if (bad) atomicOr(bad, 1u);
        row = 0; qk = -2;
      }
      if (qk != -2) d = (qk >= 0 ? bank[(long long)qk * K + (K - 1)] : 0.f) - bank[(long long)row * K + (K - 1)];
    }
    rowmap[m] = row;
    if (delta) delta[m] = d;
  }
}
__global__ void add_column_kernel(float* A, long long ld, int col, const float* v, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) A[(long long)i * ld + col] += v[i];
}

// fp32 -> operand copies
__global__ void __launch_bounds__(256)
prepare_operand_kernel(const float* __restrict__ src, long long n4, int prec, float* hi, float* lo, uint16_t* bf) {
  const float scale = (prec == VV_PREC_F16X3) ? f16_hdr(hi)->scale : 1.f;
  float amax = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ldg_stream(reinterpret_cast<const float4*>(src) + i);
    if (prec == VV_PREC_TF32X3) {
      store_x3(hi, lo, size_t(n4) * 4, size_t(i) * 4, v);
    } else if (prec == VV_PREC_F16X3) {
      store_f16x3(hi, lo, size_t(i) * 4, v, scale, amax);
    } else {
      reinterpret_cast<uint2*>(bf)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
  }
  if (prec == VV_PREC_F16X3) f16_publish_absmax(hi, amax);
}
// fp32 [rows, K] -> the row-interleaved F16X3 form (K % 64 == 0): element (r, k) -> h0 at r*2K + (k/64)*128 + k%64, h1 64 further
__global__ void __launch_bounds__(256)
prepare_interleaved_kernel(const float* __restrict__ src, long long rows, int K, void* hi) {
  const float scale = f16_hdr(hi)->scale;
  float amax = 0.f;
  const long long n4 = rows * (K >> 2);
  uint16_t* base = static_cast<uint16_t*>(hi);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ldg_stream(reinterpret_cast<const float4*>(src) + i);
    const long long r = i / (K >> 2);
    const int k = int(i - r * (K >> 2)) * 4;
    const size_t off = size_t(r) * 2 * K + size_t(k >> 6) * 128 + (k & 63);
    store_f16x3(base, base + 64, off, v, scale, amax);
  }
  f16_publish_absmax(hi, amax);
}
__global__ void set_layout_kernel(void* hi, unsigned layout) { f16_hdr(hi)->layout = layout; }
// F16X3 header maintenance
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ src, long long n, void* hi) {
  float amax = 0.f;
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ldg_stream(reinterpret_cast<const float4*>(src) + i);
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) amax = fmaxf(amax, fabsf(src[n4 * 4 + threadIdx.x]));
  f16_publish_absmax(hi, amax);
}
// extra[0..n_extra): more recorded maxima (bit patterns of non-negative floats) to fold in -- the data-parallel update
// leaves one per owner rank there (vv_dp_exchange.cuh), so that every rank derives the same scale for W
__global__ void operand_rescale_kernel(void* hi, int target_log2, int set_log2, int do_set, const unsigned int* extra, int n_extra) {
  F16Hdr* h = f16_hdr(hi);
  int e;
  if (do_set) {
    e = set_log2;
  } else {
    unsigned int bits = h->absmax_bits;
    for (int i = 0; i < n_extra; ++i) { const unsigned int b = extra[i]; bits = b > bits ? b : bits; }
    const float amax = __uint_as_float(bits);
    if (!(amax > 0.f) || !isfinite(amax)) {                 // nothing recorded: keep the scale (1 if never set)
      if (!(h->scale > 0.f)) { h->scale = 1.f; h->inv_scale = 1.f; }
      h->absmax_bits = 0u;
      return;
    }
    int ex; frexpf(amax, &ex);                              // amax = m * 2^ex, m in [0.5, 1)
    e = target_log2 - ex;
  }
  e = e < -60 ? -60 : (e > 60 ? 60 : e);
  h->scale = ldexpf(1.f, e); h->inv_scale = ldexpf(1.f, -e);
  h->absmax_bits = 0u;
}

// ----------------------------------------------------------------------------
// K4 fused update (ref: solver.cpp:534-568, net.cpp:837, blob.cpp:126-128):
//   g = scale * sum_s parts[s] ; g += decay * W (L2) | decay * sign(W) (L1)
//   hist = momentum * hist (scal) ; hist += rate * g (axpy) ; diff = hist ; W -= diff
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sgd_update_kernel(float* W, const float* parts, int nparts, long long stride,
                  float* hist, float* diff_out, long long n4, float rate, float momentum,
                  float decay, int reg_type, float gscale, int prec, float* hi, float* lo, uint16_t* bf) {
  const float scale = (prec == VV_PREC_F16X3 && hi) ? f16_hdr(hi)->scale : 1.f;
  float amax = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 g = reinterpret_cast<const float4*>(parts)[i];
    for (int s = 1; s < nparts; ++s) {
      const float4 t = reinterpret_cast<const float4*>(parts + s * stride)[i];
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
    if (gscale != 1.f) { g.x *= gscale; g.y *= gscale; g.z *= gscale; g.w *= gscale; }
    float4 w = reinterpret_cast<float4*>(W)[i];
    if (decay != 0.f) {
      if (reg_type == 2) {
        g.x = fmaf(decay, w.x, g.x); g.y = fmaf(decay, w.y, g.y); g.z = fmaf(decay, w.z, g.z); g.w = fmaf(decay, w.w, g.w);
      } else {
        g.x += decay * float((0.f < w.x) - (w.x < 0.f)); g.y += decay * float((0.f < w.y) - (w.y < 0.f));
        g.z += decay * float((0.f < w.z) - (w.z < 0.f)); g.w += decay * float((0.f < w.w) - (w.w < 0.f));
      }
    }
    float4 h = reinterpret_cast<float4*>(hist)[i];
    h.x = fmaf(rate, g.x, momentum * h.x); h.y = fmaf(rate, g.y, momentum * h.y);
    h.z = fmaf(rate, g.z, momentum * h.z); h.w = fmaf(rate, g.w, momentum * h.w);
    w.x -= h.x; w.y -= h.y; w.z -= h.z; w.w -= h.w;
    reinterpret_cast<float4*>(hist)[i] = h;
    reinterpret_cast<float4*>(W)[i] = w;
    if (diff_out) reinterpret_cast<float4*>(diff_out)[i] = h;
    if (prec == VV_PREC_TF32X3 && hi) {
      store_x3(hi, lo, size_t(n4) * 4, size_t(i) * 4, w);
    } else if (prec == VV_PREC_F16X3 && hi) {
      store_f16x3(hi, lo, size_t(i) * 4, w, scale, amax);
    } else if (prec == VV_PREC_BF16 && bf) {
      reinterpret_cast<uint2*>(bf)[i] = make_uint2(pack_bf16x2(w.x, w.y), pack_bf16x2(w.z, w.w));
    }
  }
  if (prec == VV_PREC_F16X3 && hi) f16_publish_absmax(hi, amax);
}
// scalar tail / unaligned version (bias blobs whose count is not a multiple of 4)
__global__ void sgd_update_scalar_kernel(float* W, const float* parts, int nparts, long long stride, float* hist,
                                         float* diff_out, long long n, float rate, float momentum, float decay,
                                         int reg_type, float gscale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = parts[i];
    for (int s = 1; s < nparts; ++s) g += parts[s * stride + i];
    if (gscale != 1.f) g *= gscale;
    float w = W[i];
    if (decay != 0.f) g = (reg_type == 2) ? fmaf(decay, w, g) : g + decay * float((0.f < w) - (w < 0.f));
    const float h = fmaf(rate, g, momentum * hist[i]);
    hist[i] = h; W[i] = w - h;
    if (diff_out) diff_out[i] = h;
  }
}

// weights + bias in one launch (see UpdateTail in vv_common.cuh); same per-element arithmetic as sgd_update_kernel
__global__ void __launch_bounds__(256)
sgd_update_tail_kernel(const UpdateTail u, const int main_blocks) {
  if (int(blockIdx.x) >= main_blocks) {            // ---- bias blob (scalar, nb elements)
    for (int i = (blockIdx.x - main_blocks) * blockDim.x + threadIdx.x; i < u.nb; i += (gridDim.x - main_blocks) * blockDim.x)
      sgd_update_bias1(u, i);
    return;
  }
  float* hi = static_cast<float*>(u.Wop_hi);
  const float scale = (u.prec == VV_PREC_F16X3 && hi) ? f16_hdr(hi)->scale : 1.f;
  const long long n4 = u.count / 4, k4 = u.K / 4;
  float amax = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)main_blocks * blockDim.x) {
    float4 g = reinterpret_cast<const float4*>(u.parts)[i];
    const bool last_col = (i % k4) == k4 - 1;                    // .w is column K-1 of row i / k4
    if (last_col && u.col_add) g.w += u.col_add[i / k4];
    for (int s = 1; s < u.nparts; ++s) {
      const float4 t = reinterpret_cast<const float4*>(u.parts + s * u.stride)[i];
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
    sgd_update4(u, i, g, scale, amax);
  }
  if (u.prec == VV_PREC_F16X3 && hi) f16_publish_absmax(hi, amax);
}

__global__ void __launch_bounds__(256)
reduce_parts_kernel(const float* parts, int nparts, long long stride, long long n, float* out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = parts[i];
    for (int s = 1; s < nparts; ++s) g += parts[s * stride + i];
    out[i] = g;
  }
}

// db[n] = sum_m dZ[m, n]  (ref: inner_product_layer.cu:48-50, gemv with the ones vector)
// grid.x covers column groups of 32, grid.y splits rows; partial sums combined with atomics on a zeroed db.
__global__ void __launch_bounds__(256)
bias_grad_kernel(const float* __restrict__ dZ, int M, int N, float* __restrict__ db) {
  __shared__ float sm[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r0 = threadIdx.x >> 5;
  float acc = 0.f;
  if (col < N)
    for (long long m = blockIdx.y * 8 + r0; m < M; m += (long long)gridDim.y * 8) acc += dZ[m * N + col];
  sm[r0][threadIdx.x & 31] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += sm[r][threadIdx.x];
    if (col < N) atomicAdd(db + col, s);
  }
}

__global__ void __launch_bounds__(256)
fill_bank_kernel(float* __restrict__ bank, long long n4, int K, unsigned long long seed) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long e = (unsigned long long)i * 4;
    float4 v = make_float4(bank_value(seed, e), bank_value(seed, e + 1), bank_value(seed, e + 2), bank_value(seed, e + 3));
    reinterpret_cast<float4*>(bank)[i] = v;
  }
}

// ----------------------------------------------------------------------------
// standalone layer kernels (scalar grid-stride; these are not on the fused path)
// ----------------------------------------------------------------------------
#define GS_LOOP(i, n) for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (n); i += (long long)gridDim.x * blockDim.x)

__global__ void relu_fwd_kernel(const float* x, long long n, float s, float* y) {
  GS_LOOP(i, n) { const float v = x[i]; y[i] = fmaxf(v, 0.f) + s * fminf(v, 0.f); }
}
__global__ void relu_bwd_kernel(const float* x, const float* dy, long long n, float s, float* dx) {
  GS_LOOP(i, n) { dx[i] = dy[i] * ((x[i] > 0.f) + s * (x[i] <= 0.f)); }
}
__global__ void dropout_kernel(const float* x, const unsigned* mask, int mode, unsigned thres, long long n, float scale, float* y) {
  GS_LOOP(i, n) {
    const unsigned m = mask[i];
    const unsigned keep = (mode == VV_DROPOUT_MASK_U32) ? (m > thres) : m;
    y[i] = x[i] * float(keep) * scale;
  }
}
// same stream as the fused fc7 epilogue: element (row, 4*col4 + j) = word j of dropout_words(seed, step, row, col4)
__global__ void dropout_make_mask_kernel(unsigned* mask, int rows, int cols4, unsigned thres, unsigned long long seed, unsigned long long step,
                                         int mode) {
  const unsigned base = dropout_hash_base(seed, step);
  GS_LOOP(i, (long long)rows * cols4) {
    const unsigned row = unsigned(i / cols4), c4 = unsigned(i - (long long)row * cols4);
    unsigned w[4];
    if (mode == VV_DROPOUT_HASH) dropout_words_hash(base, row, c4, w); else dropout_words(seed, step, row, c4, w);
    reinterpret_cast<uint4*>(mask)[i] = make_uint4(w[0] > thres, w[1] > thres, w[2] > thres, w[3] > thres);
  }
}
__global__ void axpby_kernel(long long n, float a, const float* x, float b, float* y) {
  GS_LOOP(i, n) { y[i] = (b == 0.f) ? a * x[i] : fmaf(a, x[i], b * y[i]); }
}
__global__ void sign_axpy_kernel(long long n, float a, const float* x, float* y) {
  GS_LOOP(i, n) { const float v = x[i]; y[i] += a * float((0.f < v) - (v < 0.f)); }
}
__global__ void mul_kernel(long long n, const float* a, const float* b, float* y) { GS_LOOP(i, n) { y[i] = a[i] * b[i]; } }
struct PtrPack { const float* p[VV_MAX_CONTEXT]; float c[VV_MAX_CONTEXT]; int nb; };
__global__ void eltwise_sum_kernel(const PtrPack pk, long long n, float* top) {
  GS_LOOP(i, n) {
    float acc = 0.f;
    for (int k = 0; k < pk.nb; ++k) acc = fmaf(pk.c[k], pk.p[k][i], acc);
    top[i] = acc;
  }
}
// one warp per row
__global__ void l2norm_fwd_kernel(const float* x, int num, int dim, float* y) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < num; r += nwarps) {
    const float* xr = x + (size_t)r * dim;
    float s = 0.f;
    for (int c = lane; c < dim; c += 32) s = fmaf(xr[c], xr[c], s);
    s = warp_sum(s);
    const float d = sqrtf(s) + 1e-10f;           // pow(s,.5) + eps (normalization_layer.cpp:36-50)
    for (int c = lane; c < dim; c += 32) y[(size_t)r * dim + c] = xr[c] / d;
  }
}
__global__ void l2norm_bwd_kernel(const float* x, const float* dy, int num, int dim, float* dx) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < num; r += nwarps) {
    const float* xr = x + (size_t)r * dim; const float* dr = dy + (size_t)r * dim;
    float a = 0.f, s = 0.f;
    for (int c = lane; c < dim; c += 32) { a = fmaf(xr[c], dr[c], a); s = fmaf(xr[c], xr[c], s); }
    a = warp_sum(a); s = warp_sum(s);
    const float q = powf(s, 1.5f) + 1e-10f;      // normalization_layer.cpp:101-110
    for (int c = lane; c < dim; c += 32) dx[(size_t)r * dim + c] = (s * dr[c] - xr[c] * a) / q;
  }
}
__global__ void rowsum_fwd_kernel(const float* x, int num, int dim, int nout, float* y) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < num; r += nwarps) {
    float s = 0.f;
    for (int c = lane; c < dim; c += 32) s += x[(size_t)r * dim + c];
    s = warp_sum(s);
    for (int o = lane; o < nout; o += 32) y[(size_t)r * nout + o] = s;
  }
}
__global__ void rowsum_bwd_kernel(const float* dy, int num, int dim, int nout, float* dx) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < num; r += nwarps) {
    float t = 0.f;
    for (int o = 0; o < nout; ++o) t += dy[(size_t)r * nout + o];   // gemv with ones (sum_layer.cpp:65-68)
    for (int c = lane; c < dim; c += 32) dx[(size_t)r * dim + c] = t;
  }
}
__global__ void copy_strided_kernel(const float* src, long long ss, float* dst, long long ds, long long rows, long long cols) {
  GS_LOOP(i, rows * cols) { const long long r = i / cols, c = i - r * cols; dst[r * ds + c] = src[r * ss + c]; }
}
// single block: hinge terms, loss, violations (max_margin_loss_layer.cpp:54-127).  w: optional per-element weights (the
// layer's third bottom with use_direct_weight, or video ids mapped through id_to_weight_file): the forward term is
// sqrt(w) * h for L2 and w * h for L1 (:84-97)
__global__ void __launch_bounds__(1024)
max_margin_fwd_kernel(const float* st, const float* sb, const float* w, int count, float margin, int norm, float* hinge, float* loss, float* viol) {
  __shared__ float s1[32], s2[32];
  float a = 0.f, v = 0.f;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const float d = st[i] - sb[i];
    if (d < 0.f) v += 1.f;
    float h = fmaxf(0.f, margin - d);
    if (w) h = (norm == 2 ? sqrtf(w[i]) : w[i]) * h;
    if (hinge) hinge[i] = h;
    a += (norm == 2) ? h * h : fabsf(h);
  }
  a = warp_sum(a); v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = a; s2[threadIdx.x >> 5] = v; }
  __syncthreads();
  if (threadIdx.x < 32) {
    a = s1[threadIdx.x]; v = s2[threadIdx.x];
    a = warp_sum(a); v = warp_sum(v);
    if (threadIdx.x == 0) { if (loss) *loss = a / count; if (viol) *viol = v; }
  }
}
// :130-214.  Weighted: t = w * h (NOT sqrt(w): the reference's forward / backward asymmetry, :154); L2 g = t * lw*2/count,
// L1 g = [t > 0] * w * lw/count (:178-184)
__global__ void max_margin_bwd_kernel(const float* st, const float* sb, const float* w, int count, float margin, int norm, float gs, float* dt, float* dbg) {
  GS_LOOP(i, count) {
    float h = fmaxf(0.f, margin - (st[i] - sb[i]));
    const float wi = w ? w[i] : 1.f;
    if (w) h = wi * h;
    const float g = (norm == 2) ? h * gs : (h > 0.f ? wi * gs : 0.f);
    if (dbg) dbg[i] = g;
    if (dt) dt[i] = -1.f * g;
  }
}
// weight of element i = table[video id], ids as floats (the blob's type); ids absent from the table weigh 0, as the
// reference's std::map::operator[] default-inserts (:93-95).  table_ids ascending.
__global__ void id_to_weight_kernel(const float* ids, int count, const int* table_ids, const float* table_w, int n, float* out) {
  GS_LOOP(i, count) {
    const int id = static_cast<int>(ids[i]);
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (table_ids[mid] < id) lo = mid + 1; else hi = mid; }
    out[i] = (lo < n && table_ids[lo] == id) ? table_w[lo] : 0.f;
  }
}

}  // namespace
}  // namespace vv

using namespace vv;

#define VV_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" int vv_gather_rows(const float* bank, int64_t bank_rows, int K, const int32_t* idx, const int32_t* quirk,
                              int B, int R, float* X, void* Xop_hi, void* Xop_lo, int prec, float* Xblob,
                              vv_stream_t stream) {
  VV_REQUIRE(bank && idx && B > 0 && R > 0 && bank_rows > 0, "bad gather arguments");
  VV_REQUIRE(K > 0 && K % 4 == 0, "K=%d must be a multiple of 4", K);
  VV_REQUIRE(VV_ALIGNED16(bank) && VV_ALIGNED16(X) && VV_ALIGNED16(Xop_hi) && VV_ALIGNED16(Xop_lo) && VV_ALIGNED16(Xblob),
             "gather buffers must be 16-byte aligned");
  GatherOut o; o.X = X; o.blob = Xblob; o.hi = nullptr; o.lo = nullptr; o.bf = nullptr; o.prec = VV_PREC_FP32_SIMT;
  if ((prec == VV_PREC_TF32X3 || prec == VV_PREC_F16X3) && Xop_hi) {
    VV_REQUIRE(Xop_lo, "split operand copy needs hi and lo");
    o.hi = static_cast<float*>(Xop_hi); o.lo = static_cast<float*>(Xop_lo); o.prec = prec;
  } else if (prec == VV_PREC_BF16 && Xop_hi) {
    o.bf = static_cast<uint16_t*>(Xop_hi); o.prec = prec;
  }
  VV_REQUIRE(o.X || o.blob || o.prec != VV_PREC_FP32_SIMT, "no gather output requested");
  const long long M = (long long)B * R;
  const int grid = int(M < (long long)num_sms() * 16 ? M : (long long)num_sms() * 16);
  gather_rows_kernel<<<grid, 256, 0, stream>>>(bank, idx, quirk, B, R, K, o);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}

extern "C" int vv_gather_plan(const float* bank, int K, const int32_t* idx, const int32_t* quirk, int B, int R,
                              int32_t* rowmap, float* delta, vv_stream_t stream) {
  return vv_gather_plan_checked(bank, 0, K, idx, quirk, B, R, rowmap, delta, nullptr, stream);
}
extern "C" int vv_gather_plan_checked(const float* bank, int64_t bank_rows, int K, const int32_t* idx, const int32_t* quirk, int B, int R,
                                      int32_t* rowmap, float* delta, uint32_t* bad_flag, vv_stream_t stream) {
  VV_REQUIRE(bank && idx && rowmap && B > 0 && R > 0 && K > 0 && bank_rows >= 0, "gather_plan: bad arguments");
  const int M = B * R, Mpad = ((M + 127) / 128) * 128;
  gather_plan_kernel<<<stream_grid(Mpad, 256), 256, 0, stream>>>(bank, idx, quirk, B, R, K, Mpad, rowmap, delta, bank_rows, bad_flag);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}
extern "C" int vv_add_column(float* A, int64_t ld, int col, const float* v, int n, vv_stream_t stream) {
  VV_REQUIRE(A && v && n > 0 && ld > 0 && col >= 0 && col < ld, "add_column: bad arguments");
  add_column_kernel<<<stream_grid(n, 256), 256, 0, stream>>>(A, ld, col, v, n);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}

extern "C" int vv_prepare_operand(const float* src, int64_t count, int prec, void* hi, void* lo, vv_stream_t stream) {
  if (prec == VV_PREC_FP32_SIMT || prec == VV_PREC_TF32) return VV_OK;
  VV_REQUIRE(src && hi && count > 0 && count % 4 == 0, "prepare_operand: bad arguments (count must be a multiple of 4)");
  VV_REQUIRE((prec != VV_PREC_TF32X3 && prec != VV_PREC_F16X3) || lo, "split operand formats need a lo array");
  VV_REQUIRE(VV_ALIGNED16(src) && VV_ALIGNED16(hi) && VV_ALIGNED16(lo), "operand buffers must be 16-byte aligned");
  const long long n4 = count / 4;
  if (prec == VV_PREC_F16X3) {
    int rc;
    if ((rc = vv_operand_set_scale(hi, prec, 0, stream))) return rc;
    if ((rc = vv_operand_measure(hi, prec, src, count, stream))) return rc;
    if ((rc = vv_operand_rescale(hi, prec, 10, stream))) return rc;
    set_layout_kernel<<<1, 1, 0, stream>>>(hi, 0u);          // two planes
    VV_LAUNCH_CHECK();
    count_launch();
  }
  prepare_operand_kernel<<<stream_grid(n4, 256), 256, 0, stream>>>(src, n4, prec, static_cast<float*>(hi),
                                                                  static_cast<float*>(lo), static_cast<uint16_t*>(hi));
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}

extern "C" int vv_prepare_bank_operand(const float* bank, int64_t rows, int K, int prec, void* hi, void* lo, vv_stream_t stream) {
  VV_REQUIRE(bank && hi && rows > 0 && K > 0, "prepare_bank_operand: bad arguments");
  if (prec != VV_PREC_F16X3 || (K % 64) != 0) return vv_prepare_operand(bank, rows * int64_t(K), prec, hi, lo, stream);
  VV_REQUIRE(VV_ALIGNED16(bank) && VV_ALIGNED16(hi), "operand buffers must be 16-byte aligned");
  int rc;
  if ((rc = vv_operand_set_scale(hi, prec, 0, stream))) return rc;
  if ((rc = vv_operand_measure(hi, prec, bank, rows * int64_t(K), stream))) return rc;
  if ((rc = vv_operand_rescale(hi, prec, 12, stream))) return rc;
  set_layout_kernel<<<1, 1, 0, stream>>>(hi, 1u);
  VV_LAUNCH_CHECK();
  prepare_interleaved_kernel<<<stream_grid(rows * (K / 4), 256), 256, 0, stream>>>(bank, rows, K, hi);
  VV_LAUNCH_CHECK();
  count_launch(2);
  return VV_OK;
}

extern "C" size_t vv_operand_bytes(int64_t count, int prec, size_t* hi_offset, size_t* lo_offset) {
  size_t hi_off = 0, lo_off = 0, bytes = 0;
  const size_t n = count > 0 ? size_t(count) : 0;
  switch (prec) {
    case VV_PREC_TF32X3: bytes = 8 * n; lo_off = 4 * n; break;
    case VV_PREC_BF16:   bytes = 2 * n; break;
    case VV_PREC_F16X3:  hi_off = VV_F16X3_HEADER_BYTES; lo_off = hi_off + ((2 * n + 127) & ~size_t(127));
                         bytes = lo_off + 2 * n; break;
    default: break;
  }
  if (hi_offset) *hi_offset = hi_off;
  if (lo_offset) *lo_offset = lo_off;
  return bytes;
}
extern "C" int vv_operand_set_scale(void* hi, int prec, int log2_scale, vv_stream_t stream) {
  if (prec != VV_PREC_F16X3) return VV_OK;
  VV_REQUIRE(hi, "operand_set_scale: NULL operand");
  operand_rescale_kernel<<<1, 1, 0, stream>>>(hi, 0, log2_scale, 1, nullptr, 0);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}
extern "C" int vv_operand_rescale(void* hi, int prec, int target_log2, vv_stream_t stream) {
  return vv::operand_rescale_ex(hi, prec, target_log2, nullptr, 0, stream);
}
int vv::operand_rescale_ex(void* hi, int prec, int target_log2, const unsigned int* extra_bits, int n_extra, vv_stream_t stream) {
  if (prec != VV_PREC_F16X3) return VV_OK;
  VV_REQUIRE(hi, "operand_rescale: NULL operand");
  VV_REQUIRE(target_log2 >= -14 && target_log2 <= 15, "operand_rescale: target_log2 must lie in the fp16 exponent range");
  operand_rescale_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(hi, target_log2, 0, 0, extra_bits, n_extra);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}
extern "C" int vv_operand_measure(void* hi, int prec, const float* src, int64_t count, vv_stream_t stream) {
  if (prec != VV_PREC_F16X3) return VV_OK;
  VV_REQUIRE(hi && src && count > 0 && VV_ALIGNED16(src), "operand_measure: bad arguments");
  absmax_kernel<<<stream_grid(count / 4 + 1, 256), 256, 0, stream>>>(src, count, hi);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}

extern "C" int vv_sgd_update(float* W, const float* grad_parts, int nparts, int64_t part_stride, float* hist,
                             float* diff_out, int64_t count, float local_rate, float momentum, float local_decay,
                             int reg_type, float grad_scale, void* Wop_hi, void* Wop_lo, int prec, vv_stream_t stream) {
  VV_REQUIRE(W && grad_parts && hist && count > 0 && nparts >= 1, "sgd_update: bad arguments");
  VV_REQUIRE(reg_type == 1 || reg_type == 2, "regularization type must be 1 (L1) or 2 (L2)");
  const bool vec = (count % 4 == 0) && (part_stride % 4 == 0) && VV_ALIGNED16(W) && VV_ALIGNED16(grad_parts) &&
                   VV_ALIGNED16(hist) && VV_ALIGNED16(diff_out) && VV_ALIGNED16(Wop_hi) && VV_ALIGNED16(Wop_lo);
  if (vec) {
    const long long n4 = count / 4;
    sgd_update_kernel<<<stream_grid(n4, 256), 256, 0, stream>>>(
        W, grad_parts, nparts, part_stride, hist, diff_out, n4, local_rate, momentum, local_decay, reg_type,
        grad_scale, prec, static_cast<float*>(Wop_hi), static_cast<float*>(Wop_lo), static_cast<uint16_t*>(Wop_hi));
  } else {
    VV_REQUIRE(!Wop_hi, "operand refresh needs a 16-byte aligned blob whose count is a multiple of 4");
    sgd_update_scalar_kernel<<<stream_grid(count, 256), 256, 0, stream>>>(
        W, grad_parts, nparts, part_stride, hist, diff_out, count, local_rate, momentum, local_decay, reg_type, grad_scale);
  }
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}

int vv::sgd_update_tail(const UpdateTail& u, vv_stream_t stream) {
  VV_REQUIRE(u.W && u.parts && u.hist && u.b && u.db && u.bh && u.count > 0 && u.nparts >= 1 && u.nb > 0, "sgd_update_tail: bad arguments");
  VV_REQUIRE(u.reg_type == 1 || u.reg_type == 2, "regularization type must be 1 (L1) or 2 (L2)");
  VV_REQUIRE(u.K > 0 && u.K % 4 == 0 && u.count % u.K == 0 && u.stride % 4 == 0 && VV_ALIGNED16(u.W) && VV_ALIGNED16(u.parts) &&
             VV_ALIGNED16(u.hist) && VV_ALIGNED16(u.diff_out) && VV_ALIGNED16(u.Wop_hi) && VV_ALIGNED16(u.Wop_lo),
             "sgd_update_tail: needs 16-byte aligned blobs and K a multiple of 4");
  const int main_blocks = stream_grid(u.count / 4, 256), bias_blocks = (u.nb + 255) / 256;
  sgd_update_tail_kernel<<<main_blocks + bias_blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(u, main_blocks);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}

extern "C" float vv_learning_rate(const char* policy, float base_lr, float gamma, float power, int stepsize, int iter) {
  // ref: solver.cpp:441-460, evaluated in Dtype = float
  if (!policy || !strcmp(policy, "fixed")) return base_lr;
  // pow(float, int) promotes to double in C++11, the product with base_lr is rounded to float once
  if (!strcmp(policy, "step")) { const int cur = iter / (stepsize > 0 ? stepsize : 1); return float(base_lr * pow(double(gamma), double(cur))); }
  if (!strcmp(policy, "exp")) return float(base_lr * pow(double(gamma), double(iter)));
  if (!strcmp(policy, "inv")) return base_lr * powf(float(1) + gamma * iter, -power);
  set_error("Unknown learning rate policy: %s", policy);
  return -1.f;
}

extern "C" int vv_reduce_parts(const float* parts, int nparts, int64_t stride, int64_t count, float* out, vv_stream_t stream) {
  VV_REQUIRE(parts && out && nparts >= 1 && count > 0, "reduce_parts: bad arguments");
  reduce_parts_kernel<<<stream_grid(count, 256), 256, 0, stream>>>(parts, nparts, stride, count, out);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}

extern "C" int vv_ip_bias_grad(const float* dZ, int M, int N, float* db, vv_stream_t stream) {
  VV_REQUIRE(dZ && db && M > 0 && N > 0, "bias_grad: bad arguments");
  VV_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * N, stream));
  int gy = (M + 255) / 256; if (gy > 64) gy = 64; if (gy < 1) gy = 1;
  bias_grad_kernel<<<dim3((N + 31) / 32, gy), 256, 0, stream>>>(dZ, M, N, db);
  VV_LAUNCH_CHECK();
  count_launch(2);
  return VV_OK;
}

extern "C" int vv_fill_bank(float* bank, int64_t rows, int K, uint64_t seed, vv_stream_t stream) {
  VV_REQUIRE(bank && rows > 0 && K > 0 && K % 4 == 0, "fill_bank: bad arguments");
  const long long n4 = rows * (long long)K / 4;
  fill_bank_kernel<<<stream_grid(n4, 256), 256, 0, stream>>>(bank, n4, K, seed);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}
extern "C" float vv_bank_value_host(uint64_t seed, int64_t row, int col, int K) {
  return bank_value(seed, uint64_t(row) * uint64_t(K) + uint64_t(col));
}

// ---- standalone layers -------------------------------------------------------
#define VV_SIMPLE_LAUNCH(kernel, n, ...)                                         \
  do {                                                                           \
    kernel<<<stream_grid((n), 256), 256, 0, s>>>(__VA_ARGS__);                   \
    VV_LAUNCH_CHECK();                                                           \
    count_launch();                                                              \
    return VV_OK;                                                                \
  } while (0)

extern "C" int vv_relu_forward(const float* x, int64_t n, float slope, float* y, vv_stream_t s) {
  VV_REQUIRE(x && y && n > 0, "relu_forward: bad arguments");
  VV_SIMPLE_LAUNCH(relu_fwd_kernel, n, x, n, slope, y);
}
extern "C" int vv_relu_backward(const float* x, const float* dy, int64_t n, float slope, float* dx, vv_stream_t s) {
  VV_REQUIRE(x && dy && dx && n > 0, "relu_backward: bad arguments");
  VV_SIMPLE_LAUNCH(relu_bwd_kernel, n, x, dy, n, slope, dx);
}
extern "C" int vv_dropout_forward(const float* x, const uint32_t* mask, int mode, int64_t n, float ratio, float* y, vv_stream_t s) {
  VV_REQUIRE(x && y && mask && n > 0 && (mode == VV_DROPOUT_MASK01 || mode == VV_DROPOUT_MASK_U32), "dropout_forward: bad arguments");
  VV_SIMPLE_LAUNCH(dropout_kernel, n, x, mask, mode, dropout_uint_thres(ratio), n, dropout_scale(ratio), y);
}
extern "C" int vv_dropout_backward(const float* dy, const uint32_t* mask, int mode, int64_t n, float ratio, float* dx, vv_stream_t s) {
  VV_REQUIRE(dy && dx && mask && n > 0 && (mode == VV_DROPOUT_MASK01 || mode == VV_DROPOUT_MASK_U32), "dropout_backward: bad arguments");
  VV_SIMPLE_LAUNCH(dropout_kernel, n, dy, mask, mode, dropout_uint_thres(ratio), n, dropout_scale(ratio), dx);
}
extern "C" int vv_dropout_make_mask(uint32_t* mask01, int rows, int cols, float ratio, uint64_t seed, uint64_t step, vv_stream_t s) {
  return vv_dropout_make_mask_mode(mask01, rows, cols, ratio, seed, step, VV_DROPOUT_PHILOX, s);
}
extern "C" int vv_dropout_make_mask_mode(uint32_t* mask01, int rows, int cols, float ratio, uint64_t seed, uint64_t step, int mode,
                                         vv_stream_t s) {
  VV_REQUIRE(mask01 && rows > 0 && cols > 0 && cols % 4 == 0 && VV_ALIGNED16(mask01), "dropout_make_mask: cols must be a multiple of 4, mask 16-byte aligned");
  VV_REQUIRE(mode == VV_DROPOUT_PHILOX || mode == VV_DROPOUT_HASH, "dropout_make_mask: mode must be a generated stream (PHILOX or HASH)");
  VV_SIMPLE_LAUNCH(dropout_make_mask_kernel, (long long)rows * (cols / 4), mask01, rows, cols / 4, dropout_uint_thres(ratio), seed, step, mode);
}
extern "C" int vv_eltwise_sum_forward(const float* const* bottoms, const float* coeffs, int nb, int64_t n, float* top, vv_stream_t s) {
  VV_REQUIRE(bottoms && coeffs && top && nb >= 1 && nb <= VV_MAX_CONTEXT && n > 0, "eltwise_sum: bad arguments");
  PtrPack pk; pk.nb = nb;
  for (int i = 0; i < nb; ++i) { pk.p[i] = bottoms[i]; pk.c[i] = coeffs[i]; }
  VV_SIMPLE_LAUNCH(eltwise_sum_kernel, n, pk, n, top);
}
extern "C" int vv_eltwise_prod_forward(const float* a, const float* b, int64_t n, float* top, vv_stream_t s) {
  VV_REQUIRE(a && b && top && n > 0, "eltwise_prod: bad arguments");
  VV_SIMPLE_LAUNCH(mul_kernel, n, n, a, b, top);
}
extern "C" int vv_axpby(int64_t n, float alpha, const float* x, float beta, float* y, vv_stream_t s) {
  VV_REQUIRE(x && y && n > 0, "axpby: bad arguments");
  VV_SIMPLE_LAUNCH(axpby_kernel, n, n, alpha, x, beta, y);
}
extern "C" int vv_sign_axpy(int64_t n, float alpha, const float* x, float* y, vv_stream_t s) {
  VV_REQUIRE(x && y && n > 0, "sign_axpy: bad arguments");
  VV_SIMPLE_LAUNCH(sign_axpy_kernel, n, n, alpha, x, y);
}
extern "C" int vv_mul(int64_t n, const float* a, const float* b, float* y, vv_stream_t s) {
  VV_REQUIRE(a && b && y && n > 0, "mul: bad arguments");
  VV_SIMPLE_LAUNCH(mul_kernel, n, n, a, b, y);
}
extern "C" int vv_l2norm_forward(const float* x, int num, int dim, float* y, vv_stream_t s) {
  VV_REQUIRE(x && y && num > 0 && dim > 0, "l2norm_forward: bad arguments");
  VV_SIMPLE_LAUNCH(l2norm_fwd_kernel, (long long)num * 32, x, num, dim, y);
}
extern "C" int vv_l2norm_backward(const float* x, const float* dy, int num, int dim, float* dx, vv_stream_t s) {
  VV_REQUIRE(x && dy && dx && num > 0 && dim > 0, "l2norm_backward: bad arguments");
  VV_SIMPLE_LAUNCH(l2norm_bwd_kernel, (long long)num * 32, x, dy, num, dim, dx);
}
extern "C" int vv_rowsum_forward(const float* x, int num, int dim, int nout, float* y, vv_stream_t s) {
  VV_REQUIRE(x && y && num > 0 && dim > 0 && nout > 0, "rowsum_forward: bad arguments");
  VV_SIMPLE_LAUNCH(rowsum_fwd_kernel, (long long)num * 32, x, num, dim, nout, y);
}
extern "C" int vv_rowsum_backward(const float* dy, int num, int dim, int nout, float* dx, vv_stream_t s) {
  VV_REQUIRE(dy && dx && num > 0 && dim > 0 && nout > 0, "rowsum_backward: bad arguments");
  VV_SIMPLE_LAUNCH(rowsum_bwd_kernel, (long long)num * 32, dy, num, dim, nout, dx);
}
extern "C" int vv_copy_strided(const float* src, int64_t ss, float* dst, int64_t ds, int64_t rows, int64_t cols, vv_stream_t s) {
  VV_REQUIRE(src && dst && rows > 0 && cols > 0, "copy_strided: bad arguments");
  VV_SIMPLE_LAUNCH(copy_strided_kernel, rows * cols, src, ss, dst, ds, rows, cols);
}
extern "C" int vv_max_margin_forward_w(const float* st, const float* sb, const float* weights, int count, float margin, int norm,
                                       float* hinge, float* loss, float* viol, vv_stream_t s) {
  VV_REQUIRE(st && sb && count > 0 && (norm == 1 || norm == 2), "max_margin_forward: bad arguments");
  max_margin_fwd_kernel<<<1, 1024, 0, s>>>(st, sb, weights, count, margin, norm, hinge, loss, viol);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}
extern "C" int vv_max_margin_forward(const float* st, const float* sb, int count, float margin, int norm, float* hinge,
                                     float* loss, float* viol, vv_stream_t s) {
  return vv_max_margin_forward_w(st, sb, nullptr, count, margin, norm, hinge, loss, viol, s);
}
extern "C" int vv_max_margin_backward_w(const float* st, const float* sb, const float* weights, int count, float margin, int norm,
                                        float lw, float* d_true, float* d_bogus, vv_stream_t s) {
  VV_REQUIRE(st && sb && count > 0 && (norm == 1 || norm == 2), "max_margin_backward: bad arguments");
  const float gs = (norm == 2) ? lw * 2 / count : lw / count;
  VV_SIMPLE_LAUNCH(max_margin_bwd_kernel, count, st, sb, weights, count, margin, norm, gs, d_true, d_bogus);
}
extern "C" int vv_max_margin_backward(const float* st, const float* sb, int count, float margin, int norm, float lw,
                                      float* d_true, float* d_bogus, vv_stream_t s) {
  return vv_max_margin_backward_w(st, sb, nullptr, count, margin, norm, lw, d_true, d_bogus, s);
}
extern "C" int vv_id_to_weight(const float* ids, int count, const int* table_ids, const float* table_w, int table_size,
                               float* weights, vv_stream_t s) {
  VV_REQUIRE(ids && weights && count > 0 && table_size >= 0 && (table_size == 0 || (table_ids && table_w)), "id_to_weight: bad arguments");
  VV_SIMPLE_LAUNCH(id_to_weight_kernel, count, ids, count, table_ids, table_w, table_size, weights);
}
