// vv_dp_exchange.cu -- the data-parallel gradient exchange fused with the SGD update, over peer memory.
// See vv_dp_exchange.cuh for the protocol.  All remote traffic is plain (posted) stores over NVLink followed by
// st.release.sys flags; nothing is read or reduced remotely, no remote atomics.
// ref: the reference has no multi-GPU path; the update arithmetic is solver.cpp:534-568,
// net.cpp:837, blob.cpp:126-128 exactly as in sgd_update_tail_kernel (vv_stream_kernels.cu).
#include "vv_dp_exchange.cuh"
#include <stdlib.h>

namespace vv {
namespace {

__device__ __forceinline__ float4 ld_l2(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

struct DpRun { unsigned long long timeout_ns; unsigned int* err; int replicate_master; };

__global__ void __launch_bounds__(256)
dp_exchange_update_kernel(const DpExchange x, const DpRun run) {
  const int tid = threadIdx.x, T = blockDim.x;
  const long long n4 = x.count / 4, k4 = x.K / 4;
  const long long owned4 = (long long)x.rows_per * k4;
  unsigned int* myflags = x.peers.flags[x.rank];
  const long long gstride = (long long)gridDim.x * T;
  __shared__ unsigned int s_last, s_bits;

  // ---------------- phase A: local split-K sum -> push every row to its owner
  // (x.nparts == 0: the wgrad kernel's split-K finish already pushed the rows and raised dw_ready, vv_gemm.cuh)
  if (x.nparts > 0) {
    for (long long i = blockIdx.x * (long long)T + tid; i < n4; i += gstride) {
      float4 g = __ldcs(reinterpret_cast<const float4*>(x.parts) + i);
      const long long row = i / k4;
      if (x.col_add && (i - row * k4) == k4 - 1) g.w += x.col_add[row];     // .w is column K-1
      for (int s = 1; s < x.nparts; ++s) {
        const float4 t = __ldcs(reinterpret_cast<const float4*>(x.parts + s * x.stride) + i);
        g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
      }
      const int o = int(row / x.rows_per);
      const long long j = i - (long long)o * owned4;
      reinterpret_cast<float4*>(x.peers.recv_dw[o])[(long long)x.rank * owned4 + j] = g;
    }
    if (blockIdx.x == 0) {                                  // (db, loss, violations) to every rank
      for (int d = 0; d < x.G; ++d)
        for (int i = tid; i < x.nsmall; i += T) x.peers.recv_small[d][x.rank * x.small_stride + i] = x.small_src[i];
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence_system();
      const unsigned int prev = atomicAdd(&myflags[kDpFlagCtr], 1u);
      s_last = (prev + 1u == gridDim.x) ? 1u : 0u;          // every CTA of this rank has pushed: tell the owners
      if (s_last) myflags[kDpFlagCtr] = 0u;                 // (self-resetting: the next launch starts from zero)
    }
    __syncthreads();
    // one thread per destination: the G release stores (each a system-scope fence + a store over NVLink) run side by side
    if (s_last != 0u && tid < x.G) { __threadfence_system(); dp_st_release_sys(&x.peers.flags[tid][kDpFlagDwReady + x.rank], x.seq); }
  }
  // ---------------- phase B: wait for all G contributions, update the owned rows, push them to every rank
  if (tid < x.G) dp_spin_wait_flag(&myflags[kDpFlagDwReady + tid], x.seq, run.timeout_ns, run.err, 1u);
  __syncthreads();

  float* hi_me = static_cast<float*>(x.peers.wop_hi[x.rank]);
  const bool has_op = hi_me != nullptr && (x.prec == VV_PREC_TF32X3 || x.prec == VV_PREC_F16X3 || x.prec == VV_PREC_BF16);
  const float scale = (x.prec == VV_PREC_F16X3 && has_op) ? f16_hdr(hi_me)->scale : 1.f;
  const float rate = x.rate_w, momentum = x.momentum, decay = x.decay_w, gscale = x.gscale;
  const float* recv = x.peers.recv_dw[x.rank];
  float* Wme = x.peers.Wm[x.rank];
  const bool repl = run.replicate_master != 0 || !has_op;
  float amax = 0.f;
  for (long long j = blockIdx.x * (long long)T + tid; j < owned4; j += gstride) {
    float4 g = ld_l2(recv + j * 4);
    for (int s = 1; s < x.G; ++s) {                        // rank order: deterministic
      const float4 t = ld_l2(recv + ((long long)s * owned4 + j) * 4);
      g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
    }
    g.x *= gscale; g.y *= gscale; g.z *= gscale; g.w *= gscale;
    const long long i = (long long)x.rank * owned4 + j;   // float4 index in the whole [N,K] blob
    float4 w = reinterpret_cast<float4*>(Wme)[i];
    if (decay != 0.f) {
      if (x.reg_type == 2) {
        g.x = fmaf(decay, w.x, g.x); g.y = fmaf(decay, w.y, g.y); g.z = fmaf(decay, w.z, g.z); g.w = fmaf(decay, w.w, g.w);
      } else {
        g.x += decay * float((0.f < w.x) - (w.x < 0.f)); g.y += decay * float((0.f < w.y) - (w.y < 0.f));
        g.z += decay * float((0.f < w.z) - (w.z < 0.f)); g.w += decay * float((0.f < w.w) - (w.w < 0.f));
      }
    }
    float4 h = reinterpret_cast<float4*>(x.hist)[i];
    h.x = fmaf(rate, g.x, momentum * h.x); h.y = fmaf(rate, g.y, momentum * h.y);
    h.z = fmaf(rate, g.z, momentum * h.z); h.w = fmaf(rate, g.w, momentum * h.w);
    w.x -= h.x; w.y -= h.y; w.z -= h.z; w.w -= h.w;
    reinterpret_cast<float4*>(x.hist)[i] = h;
    if (x.diff_out) reinterpret_cast<float4*>(x.diff_out)[i] = h;
    reinterpret_cast<float4*>(Wme)[i] = w;
    const long long row = i / k4;
    const bool last_col = (i - row * k4) == k4 - 1;
    float dummy = 0.f;
    if (x.prec == VV_PREC_F16X3) amax = fmaxf(amax, fmaxf(fmaxf(fabsf(w.x), fabsf(w.y)), fmaxf(fabsf(w.z), fabsf(w.w))));
    for (int d = 0; d < x.G; ++d) {
      if (repl && d != x.rank) reinterpret_cast<float4*>(x.peers.Wm[d])[i] = w;
      if (last_col) x.peers.wlast[d][row] = w.w;
      if (!has_op) continue;
      if (x.prec == VV_PREC_TF32X3) {
        store_x3(static_cast<float*>(x.peers.wop_hi[d]), x.peers.wop_lo[d], size_t(n4) * 4, size_t(i) * 4, w);
      } else if (x.prec == VV_PREC_F16X3) {
        store_f16x3(x.peers.wop_hi[d], x.peers.wop_lo[d], size_t(i) * 4, w, scale, dummy);
      } else {
        reinterpret_cast<uint2*>(x.peers.wop_hi[d])[i] = make_uint2(pack_bf16x2(w.x, w.y), pack_bf16x2(w.z, w.w));
      }
    }
  }
  if (x.prec == VV_PREC_F16X3) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((tid & 31) == 0 && amax > 0.f) atomicMax(&myflags[kDpFlagCtr + 2], __float_as_uint(amax));
  }
  // bias blob, loss and violation count: every rank, redundantly, from the same G vectors in the same order
  if (blockIdx.x == gridDim.x - 1) {
    const float* rs = x.peers.recv_small[x.rank];
    for (int i = tid; i < x.nb; i += T) {
      float g = __ldcg(rs + i);
      for (int s = 1; s < x.G; ++s) g += __ldcg(rs + s * x.small_stride + i);
      g *= gscale;
      const float w = x.b[i];
      if (x.decay_b != 0.f) g = (x.reg_type == 2) ? fmaf(x.decay_b, w, g) : g + x.decay_b * float((0.f < w) - (w < 0.f));
      const float h = fmaf(x.rate_b, g, momentum * x.bh[i]);
      x.bh[i] = h; x.b[i] = w - h;
      if (x.b_diff) x.b_diff[i] = h;
    }
    if (tid == 0) {
      float l = __ldcg(rs + x.nb), v = __ldcg(rs + x.nb + 1);
      for (int s = 1; s < x.G; ++s) { l += __ldcg(rs + s * x.small_stride + x.nb); v += __ldcg(rs + s * x.small_stride + x.nb + 1); }
      if (x.loss_out) *x.loss_out = l * gscale;           // every rank normalised by its local B*Nn: mean over ranks
      if (x.viol_out) *x.viol_out = v;                    // violations: global count
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();
    const unsigned int prev = atomicAdd(&myflags[kDpFlagCtr + 1], 1u);
    s_last = (prev + 1u == gridDim.x) ? 1u : 0u;          // the owned rows are in flight to every rank
    if (s_last) { myflags[kDpFlagCtr + 1] = 0u; s_bits = atomicExch(&myflags[kDpFlagCtr + 2], 0u); }
  }
  __syncthreads();
  if (s_last != 0u && tid < x.G) {                        // one thread per destination rank, side by side
    __threadfence_system();
    if (x.prec == VV_PREC_F16X3) *reinterpret_cast<volatile unsigned int*>(&x.peers.flags[tid][kDpFlagAmax + x.rank]) = s_bits;
    dp_st_release_sys(&x.peers.flags[tid][kDpFlagWReady + x.rank], x.seq);      // release: ordered after the amax word
  }
}

__global__ void dp_wait_kernel(const unsigned int* flags, int G, unsigned int seq, unsigned long long timeout_ns, unsigned int* err) {
  if (int(threadIdx.x) < G) dp_spin_wait_flag(&flags[kDpFlagWReady + threadIdx.x], seq, timeout_ns, err, 2u);
}

unsigned long long dp_timeout_ns() {
  static const unsigned long long ns = [] {
    const char* e = getenv("VV_DP_TIMEOUT_MS");
    const long long ms = e ? atoll(e) : 10000;
    return (unsigned long long)(ms > 0 ? ms : 10000) * 1000000ull;
  }();
  return ns;
}

}  // namespace

int dp_exchange_grid() {
  static int grid = 0;
  if (grid) return grid;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dp_exchange_update_kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;                              // every CTA must be resident (in-kernel waits)
  grid = num_sms() * per_sm;
  return grid;
}

int dp_exchange_update(const DpExchange& x, unsigned int* err_word, int replicate_master, vv_stream_t stream) {
  VV_REQUIRE(x.G >= 2 && x.G <= kDpMaxRanks && x.rank >= 0 && x.rank < x.G && x.seq >= 1, "dp_exchange: bad rank / world");
  VV_REQUIRE(x.count > 0 && x.K % 4 == 0 && x.count % x.K == 0 && x.rows_per > 0 && (x.count / x.K) == (long long)x.rows_per * x.G,
             "dp_exchange: N must divide evenly over the ranks and K be a multiple of 4");
  VV_REQUIRE(x.reg_type == 1 || x.reg_type == 2, "regularization type must be 1 (L1) or 2 (L2)");
  DpRun run; run.timeout_ns = dp_timeout_ns(); run.err = err_word; run.replicate_master = replicate_master;
  dp_exchange_update_kernel<<<dp_exchange_grid(), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, run);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}

int dp_wait_w_ready(const unsigned int* flags, int G, unsigned int seq, unsigned int* err_word, vv_stream_t stream) {
  dp_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(flags, G, seq, dp_timeout_ns(), err_word);
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}

unsigned long long dp_wait_timeout_ns() { return dp_timeout_ns(); }

}  // namespace vv
