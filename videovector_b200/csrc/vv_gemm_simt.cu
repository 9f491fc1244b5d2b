// vv_gemm_simt.cu -- exact fp32 FMA version of the fc7 contractions (VV_PREC_FP32_SIMT).
// Any shape; used for gradient-check sized problems, odd shapes the tensor-core
// path rejects, and as an on-device cross-check of the tcgen05 kernels.
// ref: inner_product_layer.cu:12-59 (the three cublasSgemm calls it replaces).
#include "vv_gemm.cuh"

namespace vv {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtParams {
  const float* A; const float* B;
  long long sa_i, sa_r, sb_j, sb_r;   // element strides: A(i,r), B(j,r)
  int rows, cols, red;
  int nsplit, red_per_split;
  float* D; long long slab_stride; int ldd;
  int act_N; int fwd_epi;
  GemmEpilogue epi;
};

__device__ __forceinline__ float epilogue_act1(const GemmEpilogue& e, int N, int row, int col, float v, float& z) {
  if (e.bias) v += e.bias[col];
  z = v;
  if (!e.has_act) return v;
  if (e.relu) v = fmaxf(v, 0.f) + e.negative_slope * fminf(v, 0.f);
  if (e.dropout_mode == VV_DROPOUT_NONE) return v;
  uint32_t keep;
  if (e.dropout_mode == VV_DROPOUT_PHILOX || e.dropout_mode == VV_DROPOUT_HASH) {
    uint32_t w[4];
    if (e.dropout_mode == VV_DROPOUT_PHILOX) dropout_words(e.seed, e.step, uint32_t(row), uint32_t(col >> 2), w);
    else dropout_words_hash(e.hash_base, uint32_t(row), uint32_t(col >> 2), w);
    keep = (w[col & 3] > e.dropout_thres) ? 1u : 0u;
    if (e.mask_out) e.mask_out[size_t(row) * N + col] = keep;
  } else {
    const uint32_t m = e.mask[size_t(row) * N + col];
    keep = (e.dropout_mode == VV_DROPOUT_MASK_U32) ? (m > e.dropout_thres ? 1u : 0u) : m;
  }
  return v * float(keep) * e.dropout_scale;
}

__global__ void __launch_bounds__(256)
gemm_simt_kernel(const SimtParams p) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.x * TM, j0 = blockIdx.y * TN;
  const int split = blockIdx.z;
  const int r_begin = split * p.red_per_split;
  const int r_end = min(r_begin + p.red_per_split, p.red);
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  const bool a_red_contig = (p.sa_r == 1);
  const bool b_red_contig = (p.sb_r == 1);
  for (int r0 = r_begin; r0 < r_end; r0 += TK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int ii, rr;
      if (a_red_contig) { rr = tid & 15; ii = (tid >> 4) + 16 * q; }
      else              { ii = tid & 63; rr = (tid >> 6) + 4 * q; }
      const int gi = i0 + ii, gr = r0 + rr;
      As[rr][ii] = (gi < p.rows && gr < r_end) ? p.A[gi * p.sa_i + gr * p.sa_r] : 0.f;
      int jj, r2;
      if (b_red_contig) { r2 = tid & 15; jj = (tid >> 4) + 16 * q; }
      else              { jj = tid & 63; r2 = (tid >> 6) + 4 * q; }
      const int gj = j0 + jj, gr2 = r0 + r2;
      Bs[r2][jj] = (gj < p.cols && gr2 < r_end) ? p.B[gj * p.sb_j + gr2 * p.sb_r] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < TK; ++r) {
      float a[4], b[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) { a[x] = As[r][ty * 4 + x]; b[x] = Bs[r][tx * 4 + x]; }
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
  float* D = p.D + (long long)split * p.slab_stride;
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int row = i0 + ty * 4 + x;
    if (row >= p.rows) continue;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int col = j0 + tx * 4 + y;
      if (col >= p.cols) continue;
      float v = acc[x][y];
      if (p.fwd_epi) {
        float z;
        v = epilogue_act1(p.epi, p.act_N, row, col, v, z);
        if (p.epi.Z) p.epi.Z[(long long)row * p.act_N + col] = z;
      } else {
        v *= p.epi.out_scale;
      }
      D[(long long)row * p.ldd + col] = v;
    }
  }
}

}  // namespace

int gemm_simt_launch(const GemmProblem& g, cudaStream_t stream) {
  SimtParams p;
  p.A = static_cast<const float*>(g.A.hi);
  p.B = static_cast<const float*>(g.B.hi);
  switch (g.kind) {
    case GEMM_FWD:
      p.rows = g.M; p.cols = g.N; p.red = g.K;
      p.sa_i = g.K; p.sa_r = 1; p.sb_j = g.K; p.sb_r = 1; break;
    case GEMM_WGRAD:
      p.rows = g.N; p.cols = g.K; p.red = g.M;
      p.sa_i = 1; p.sa_r = g.N; p.sb_j = 1; p.sb_r = g.K; break;
    default:
      p.rows = g.M; p.cols = g.K; p.red = g.N;
      p.sa_i = g.N; p.sa_r = 1; p.sb_j = 1; p.sb_r = g.K; break;
  }
  int nsplit = g.nsplit < 1 ? 1 : g.nsplit;
  int per = (p.red + nsplit - 1) / nsplit;
  per = ((per + TK - 1) / TK) * TK;
  if ((long long)(nsplit - 1) * per >= p.red && nsplit > 1) {
    set_error("nsplit=%d leaves an empty split for reduction %d", nsplit, p.red);
    return VV_ERR_INVALID;
  }
  p.nsplit = nsplit; p.red_per_split = per;
  p.D = g.D; p.slab_stride = g.slab_stride; p.ldd = p.cols;
  p.act_N = g.N; p.fwd_epi = (g.kind == GEMM_FWD);
  p.epi = g.epi;
  dim3 grid((p.rows + TM - 1) / TM, (p.cols + TN - 1) / TN, nsplit);
  gemm_simt_kernel<<<grid, 256, 0, stream>>>(p);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}

}  // namespace vv
