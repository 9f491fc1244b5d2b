// vv_eval.cu -- TEST-phase evaluation of the embedding on the device.
//
// The shipped net's TEST graph (ref: projects/videovec_embedding/mednet_embedding_train.prototxt:29-45,75-177,
// 345-352,673-689) averages the F = 4 sampled frames of every shot window at the feature level ("average_for_test",
// ELTWISE SUM coeff 0.25), projects with the shared fc7 (+ ReLU; Dropout is a copy in TEST), L2-normalises
// ("test_norm") and feeds RetrievalStatsLayer (src/caffe/layers/retrieval_stats_layer.cpp:143-359), which runs on
// the CPU in the reference: a B x B distance matrix (-2 E E^T), one std::sort per query and ComputeStats (:98-140).
// Here: one gather+mean kernel, the fc7 GEMM and L2-norm kernels of the training path, an exact-fp32 Gram GEMM,
// and one CTA per query that bitonic-sorts the query's row in shared memory and scans it for AP / hit@1 / hit@5.
#include <math.h>
#include <string.h>
#include "vv_gemm.cuh"

namespace vv {
namespace {

constexpr int kMaxFrames = 16;
struct MeanCoeff { float c[kMaxFrames]; };

// Xbar[b,:] = sum_f coeff[f] * bank[idx[b,f],:]   (top = 0, then one axpy per bottom: eltwise_layer.cpp:67-73)
__global__ void __launch_bounds__(256)
gather_mean_kernel(const float* __restrict__ bank, const int* __restrict__ idx, int B, int F, int K, MeanCoeff coeff,
                   float* __restrict__ out) {
  const int K4 = K >> 2;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    for (int c = threadIdx.x; c < K4; c += blockDim.x) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int f = 0; f < F; ++f) {
        const float4 v = ldg_stream(reinterpret_cast<const float4*>(bank + (long long)idx[b * F + f] * K) + c);
        const float a = coeff.c[f];
        acc.x = fmaf(a, v.x, acc.x); acc.y = fmaf(a, v.y, acc.y); acc.z = fmaf(a, v.z, acc.z); acc.w = fmaf(a, v.w, acc.w);
      }
      reinterpret_cast<float4*>(out + (size_t)b * K)[c] = acc;
    }
  }
}

// One CTA per query i.  G = E E^T [B,B]; distance = -2 G (ref :226-228), the query itself forced first (-1e15, :240-241).
// Sorted by (distance, index); then ComputeStats (:98-140):  val counts the ranked candidates (all, or only other
// videos'), ret the relevant ones among them, AP = mean over relevant of ret/val, hit@1 = relevant at val 1,
// hit@5 = (#relevant with val <= 5) / 5.  per_query[i] = {ap, hit1, hit5} or -1s when labels[i] < 0 (:256-258).
__global__ void __launch_bounds__(256)
retrieval_stats_kernel(const float* __restrict__ G, int B, int n2, const int* __restrict__ video_ids,
                       const int* __restrict__ labels, int exclude_same, double* __restrict__ per_query,
                       int* __restrict__ top5) {
  extern __shared__ unsigned char sm_raw[];
  float* key = reinterpret_cast<float*>(sm_raw);
  int* id = reinterpret_cast<int*>(key + n2);
  int* cnt = id + n2;                       // [2 * T] per-thread (valid, relevant) counts
  double* red = reinterpret_cast<double*>(cnt + 2 * blockDim.x);                          // [3 * T], 8-byte aligned
  const int i = blockIdx.x, T = blockDim.x, tid = threadIdx.x;
  if (labels[i] < 0) {
    if (tid < 3) per_query[3 * i + tid] = -1.0;
    if (top5 && tid < 5) top5[5 * i + tid] = -1;
    return;
  }
  for (int j = tid; j < n2; j += T) {
    key[j] = (j < B) ? ((j == i) ? -1e15f : -2.f * G[(size_t)i * B + j]) : INFINITY;
    id[j] = j;
  }
  __syncthreads();
  // bitonic sort, ascending by (key, index)
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (n2 >> 1); t += T) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const float ka = key[lo], kb = key[hi];
        const int ia = id[lo], ib = id[hi];
        const bool a_after_b = (ka > kb) || (ka == kb && ia > ib);
        if (a_after_b == up) { key[lo] = kb; key[hi] = ka; id[lo] = ib; id[hi] = ia; }
      }
      __syncthreads();
    }
  }
  const int vi = video_ids[i], li = labels[i];
  // the CSV's "ret_id_1..5" (:309-316): the first five ranked items of ANOTHER video, whatever exclude_same says
  // (-1 where fewer exist: the reference leaves the previous query's value there, the host side mimics that)
  if (top5 && tid == 0) {
    int found = 0;
    for (int k = 0; k < B && found < 5; ++k) {
      const int j = id[k];
      if (video_ids[j] != vi) top5[5 * i + found++] = j;
    }
    for (; found < 5; ++found) top5[5 * i + found] = -1;
  }
  // scan positions 1..B-1 in chunks of consecutive positions per thread
  const int per = (B + T - 1) / T;
  const int p0 = tid * per, p1 = min(B, p0 + per);
  int nval = 0, nrel = 0;
  for (int k = max(p0, 1); k < p1; ++k) {
    const int j = id[k];
    const bool valid = (video_ids[j] != vi) || !exclude_same;
    nval += valid; nrel += valid && labels[j] == li;
  }
  cnt[2 * tid] = nval; cnt[2 * tid + 1] = nrel;
  __syncthreads();
  int val = 0, ret = 0;
  for (int t = 0; t < tid; ++t) { val += cnt[2 * t]; ret += cnt[2 * t + 1]; }     // T <= 256: a short serial prefix
  double ap = 0, a1 = 0, a5 = 0;
  for (int k = max(p0, 1); k < p1; ++k) {
    const int j = id[k];
    const bool valid = (video_ids[j] != vi) || !exclude_same;
    if (valid) {
      ++val;
      if (labels[j] == li) {
        ++ret;
        if (val <= 1) a1 += 1;
        if (val <= 5) a5 += 1;
        ap += double(ret) / double(val);
      }
    }
  }
  red[3 * tid] = ap; red[3 * tid + 1] = a1; red[3 * tid + 2] = a5;
  __syncthreads();
  if (tid == 0) {
    double sap = 0, s1 = 0, s5 = 0; int total_ret = 0;
    for (int t = 0; t < T; ++t) { sap += red[3 * t]; s1 += red[3 * t + 1]; s5 += red[3 * t + 2]; total_ret += cnt[2 * t + 1]; }
    per_query[3 * i] = total_ret > 0 ? sap / double(total_ret) : 0.0;
    per_query[3 * i + 1] = s1;
    per_query[3 * i + 2] = s5 / 5.0;
  }
}

// video_level_retrieval (:160-206): out[v,:] = sum over the batch items i of video v, in increasing i, of (1/n_v) * E[i,:]
// -- the reference's GEMM with the [V,B] matrix of 1/n_v entries.  group[i] in [0,V): the item's video.  One CTA per video.
__global__ void __launch_bounds__(256)
video_mean_kernel(const float* __restrict__ E, const int* __restrict__ group, int B, int N, float* __restrict__ out) {
  __shared__ int s_n;
  const int v = blockIdx.x;
  if (threadIdx.x == 0) { int n = 0; for (int i = 0; i < B; ++i) n += group[i] == v; s_n = n; }
  __syncthreads();
  const float w = float(1.0 / double(s_n > 0 ? s_n : 1));
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < B; ++i)
      if (group[i] == v) acc = fmaf(w, E[(size_t)i * N + c], acc);
    out[(size_t)v * N + c] = acc;
  }
}

// means over the scored queries, in index order (the reference's accumulation order, :322-327, :352-354)
__global__ void retrieval_mean_kernel(const double* __restrict__ per_query, const int* __restrict__ labels, int B,
                                      double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double ap = 0, a1 = 0, a5 = 0, n = 0;
  for (int i = 0; i < B; ++i)
    if (labels[i] >= 0) { ap += per_query[3 * i]; a1 += per_query[3 * i + 1]; a5 += per_query[3 * i + 2]; n += 1; }
  out[0] = ap / n; out[1] = a1 / n; out[2] = a5 / n;
}

// ---- IdToWeightMapping (ref: id_to_weight_mapping_layer.cpp:61-148): a per-id embedding table -----------------------
// forward: top[i,:] = table[ids[i],:]   (ids arrive as floats, like every Caffe blob)
__global__ void __launch_bounds__(256)
id_lookup_fwd_kernel(const float* __restrict__ table, const float* __restrict__ ids, int M, int N, float* __restrict__ top) {
  for (int i = blockIdx.x; i < M; i += gridDim.x) {
    const float* src = table + (size_t)static_cast<int>(ids[i]) * N;
    for (int c = threadIdx.x; c < N; c += blockDim.x) top[(size_t)i * N + c] = src[c];
  }
}
// backward: table_diff = 0; for i in order: table_diff[ids[i],:] += top_diff[i,:]  (:100-106).  Deterministic and in
// the reference's order without atomics: the CTA of the FIRST occurrence of an id sums all rows carrying that id in
// increasing i; later occurrences exit.  (table_diff must have been zeroed: rows no id names stay 0.)
__global__ void __launch_bounds__(256)
id_lookup_bwd_kernel(const float* __restrict__ top_diff, const float* __restrict__ ids, int M, int N, float* __restrict__ table_diff) {
  extern __shared__ int s_ids[];
  for (int j = threadIdx.x; j < M; j += blockDim.x) s_ids[j] = static_cast<int>(ids[j]);
  __syncthreads();
  const int i = blockIdx.x, id = s_ids[i];
  int seen = 0;
  for (int j = threadIdx.x; j < i; j += blockDim.x) seen |= (s_ids[j] == id);
  if (__syncthreads_or(seen)) return;
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    float acc = 0.f;
    for (int j = i; j < M; ++j)
      if (s_ids[j] == id) acc += top_diff[(size_t)j * N + c];
    table_diff[(size_t)id * N + c] = acc;
  }
}

}  // namespace
}  // namespace vv

using namespace vv;
#define VV_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" int vv_gather_mean_rows(const float* bank, int64_t bank_rows, int K, const int32_t* idx, int B, int F,
                                   const float* coeff_host, float* Xbar, vv_stream_t stream) {
  VV_REQUIRE(bank && idx && Xbar && B > 0 && F > 0 && bank_rows > 0, "gather_mean_rows: bad arguments");
  VV_REQUIRE(F <= kMaxFrames, "gather_mean_rows: at most %d frames per item", kMaxFrames);
  VV_REQUIRE(K > 0 && K % 4 == 0 && VV_ALIGNED16(bank) && VV_ALIGNED16(Xbar), "gather_mean_rows: K %% 4 and 16-byte alignment required");
  MeanCoeff c;
  for (int f = 0; f < kMaxFrames; ++f) c.c[f] = f < F ? (coeff_host ? coeff_host[f] : 1.f / float(F)) : 0.f;
  const int grid = B < num_sms() * 8 ? B : num_sms() * 8;
  gather_mean_kernel<<<grid, 256, 0, stream>>>(bank, idx, B, F, K, c, Xbar);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}

extern "C" size_t vv_retrieval_stats_workspace_bytes(int B) {
  return ((size_t(B) * B * sizeof(float) + 7) & ~size_t(7)) + size_t(B) * 3 * sizeof(double);
}

extern "C" int vv_video_mean_rows(const float* E, int B, int N, const int32_t* group, int V, float* out, vv_stream_t stream) {
  VV_REQUIRE(E && group && out && B > 0 && N > 0 && V > 0, "video_mean_rows: bad arguments");
  video_mean_kernel<<<V, 256, 0, stream>>>(E, group, B, N, out);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}

extern "C" int vv_retrieval_stats(const float* E, int B, int N, const int32_t* video_ids, const int32_t* labels,
                                  int exclude_same_video_shots, const float* gram_given, void* workspace,
                                  size_t workspace_bytes, double* out3, double* per_query_out, vv_stream_t stream) {
  return vv_retrieval_stats_ex(E, B, N, video_ids, labels, exclude_same_video_shots, gram_given, workspace, workspace_bytes,
                               out3, per_query_out, nullptr, stream);
}
extern "C" int vv_retrieval_stats_ex(const float* E, int B, int N, const int32_t* video_ids, const int32_t* labels,
                                     int exclude_same_video_shots, const float* gram_given, void* workspace,
                                     size_t workspace_bytes, double* out3, double* per_query_out, int32_t* top5,
                                     vv_stream_t stream) {
  VV_REQUIRE((E || gram_given) && video_ids && labels && out3 && B > 1 && N > 0, "retrieval_stats: bad arguments");
  VV_REQUIRE(workspace && workspace_bytes >= vv_retrieval_stats_workspace_bytes(B), "retrieval_stats: workspace too small");
  int n2 = 1; while (n2 < B) n2 <<= 1;
  VV_REQUIRE(n2 <= 8192, "retrieval_stats: at most 8192 items per evaluation batch (got %d)", B);
  float* G = static_cast<float*>(workspace);
  double* pq = per_query_out ? per_query_out
                             : reinterpret_cast<double*>(static_cast<char*>(workspace) + ((size_t(B) * B * sizeof(float) + 7) & ~size_t(7)));
  const float* Guse = gram_given;
  if (!Guse) {
    // G = E E^T with the exact-fp32 kernel (the ordering of near-equal distances should not hinge on tensor-core rounding)
    GemmProblem g;
    g.kind = GEMM_FWD; g.prec = VV_PREC_FP32_SIMT; g.A.hi = E; g.A.lo = nullptr; g.B.hi = E; g.B.lo = nullptr;
    g.M = B; g.N = B; g.K = N; g.D = G; g.slab_stride = 0; g.nsplit = 1; g.rowmap = nullptr; g.bank_rows = 0;
    memset(&g.epi, 0, sizeof(g.epi)); g.epi.out_scale = 1.f;
    int rc = gemm_simt_launch(g, stream);
    if (rc) return rc;
    Guse = G;
  }
  const int T = 256;
  const size_t smem = size_t(n2) * 8 + size_t(2 * T) * 4 + size_t(3 * T) * 8 + 8;
  if (smem > 48 * 1024) VV_CUDA(cudaFuncSetAttribute(retrieval_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  retrieval_stats_kernel<<<B, T, smem, stream>>>(Guse, B, n2, video_ids, labels, exclude_same_video_shots, pq, top5);
  VV_LAUNCH_CHECK();
  retrieval_mean_kernel<<<1, 32, 0, stream>>>(pq, labels, B, out3);
  VV_LAUNCH_CHECK();
  count_launch(2);
  return VV_OK;
}

extern "C" int vv_id_lookup_forward(const float* table, int rows, int N, const float* ids, int M, float* top, vv_stream_t stream) {
  VV_REQUIRE(table && ids && top && rows > 0 && N > 0 && M > 0, "id_lookup_forward: bad arguments");
  const int grid = M < num_sms() * 8 ? M : num_sms() * 8;
  id_lookup_fwd_kernel<<<grid, 256, 0, stream>>>(table, ids, M, N, top);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}
extern "C" int vv_id_lookup_backward(const float* top_diff, const float* ids, int M, int N, int rows, float* table_diff,
                                     vv_stream_t stream) {
  VV_REQUIRE(top_diff && ids && table_diff && rows > 0 && N > 0 && M > 0, "id_lookup_backward: bad arguments");
  VV_REQUIRE(size_t(M) * 4 <= 200 * 1024, "id_lookup_backward: at most 51200 ids per batch");
  VV_CUDA(cudaMemsetAsync(table_diff, 0, size_t(rows) * N * sizeof(float), stream));       // caffe_set(K_*N_, 0, diff) :98
  const size_t smem = size_t(M) * 4;
  if (smem > 48 * 1024) VV_CUDA(cudaFuncSetAttribute(id_lookup_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  id_lookup_bwd_kernel<<<M, 256, smem, stream>>>(top_diff, ids, M, N, table_diff);
  VV_LAUNCH_CHECK();
  count_launch(2);
  return VV_OK;
}
