// vv_rank_loss.cu -- K2 / K3: everything between ip2 and the loss, fused.
//
// Forward (one pass over the R rows of an item, HBM-bound, R*N*4 bytes/item):
//   slice_emb -> context_average (ELTWISE SUM, coeff) -> word_embedding_norm ->
//   concat/pos_neg_normalize/slice -> prod_* + sum_* (1+Nn dot products) ->
//   concat_negative_scores -> max_margin_loss
// Backward (read R rows, write R rows: 2*R*N*4 bytes/item): the reverse graph down to
//   ip1_nonorm.diff including the Split sum, both Normalization backwards, the
//   Eltwise SUM scale, Dropout and ReLU backward.
// ref: eltwise_layer.cpp:61-73,119-143; normalization_layer.cpp:30-112;
//      sum_layer.cpp:32-82; split_layer.cpp:36-51; max_margin_loss_layer.cpp:54-214;
//      dropout_layer.cpp:52-68; relu_layer.cpp:23-36 (math summarised in SURVEY 8a).
//
// Per item the forward saves only scalars: s_c = |cbar|^2 and, for the target and every
// negative row x, s_x = |x|^2 and p_x = <cbar, x>.  Every reduction the reference's
// backward performs (a = <x, dy> in Normalization, the row sums in Sum) is a linear
// combination of these, so the backward needs no reduction at all and streams.
//
// Thread mapping: one CTA per item (grid-stride), thread t owns float4 column
// groups t, t+T, ... ; all rows of the item are read with 128-bit loads.
#include "vv_common.cuh"

namespace vv {
namespace {

constexpr int kMaxVec = 4;        // float4 groups per thread: N <= 4 * 256 * kMaxVec = 4096
constexpr int kRowBatch = 4;      // rows loaded before reducing (memory-level parallelism)

struct RankDev {
  int B, C, Nn, N, N4, nvec, stride;
  float margin, eps; int norm;
  float coeff[VV_MAX_CONTEXT];
};

__device__ __forceinline__ float4 ld4(const float* p) { return ldg_stream(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// c-bar = sum_i coeff_i * ctx_i  (zeroed output, then axpy in bottom order: eltwise_layer.cpp:67-73)
template <int NVEC>
__device__ __forceinline__ void context_mean(const float* H, const RankDev& p, int b, int tid, int T, float4 (&cbar)[NVEC]) {
#pragma unroll
  for (int v = 0; v < NVEC; ++v) cbar[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i0 = 1; i0 < p.C; i0 += kRowBatch) {
    float4 x[kRowBatch][NVEC];
#pragma unroll
    for (int r = 0; r < kRowBatch; ++r)
#pragma unroll
      for (int v = 0; v < NVEC; ++v) {
        const int c4 = v * T + tid;
        if (i0 + r < p.C && c4 < p.N4)
          x[r][v] = ld4(H + (size_t(i0 + r) * p.B + b) * p.N + c4 * 4);
        else
          x[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int r = 0; r < kRowBatch; ++r) {
      if (i0 + r < p.C) {
        const float a = p.coeff[i0 + r - 1];
#pragma unroll
        for (int v = 0; v < NVEC; ++v) {
          cbar[v].x = fmaf(a, x[r][v].x, cbar[v].x); cbar[v].y = fmaf(a, x[r][v].y, cbar[v].y);
          cbar[v].z = fmaf(a, x[r][v].z, cbar[v].z); cbar[v].w = fmaf(a, x[r][v].w, cbar[v].w);
        }
      }
    }
  }
}

template <int NVEC>
__global__ void __launch_bounds__(256)
rank_fwd_kernel(const float* __restrict__ H, const RankDev p, float* __restrict__ stats,
                float* __restrict__ tscore, float* __restrict__ nscore,
                float* __restrict__ item_loss, float* __restrict__ item_viol) {
  extern __shared__ float sm[];
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = T >> 5;
  const int J = 1 + p.Nn;                 // target + negatives
  float* part = sm;                       // [J][2][nw]
  float* part_c = part + J * 2 * nw;      // [nw]
  float* fin = part_c + nw;               // [J][2] then s_c at fin[2J]
  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    float4 cbar[NVEC];
    context_mean<NVEC>(H, p, b, tid, T, cbar);
    float sc = 0.f;
#pragma unroll
    for (int v = 0; v < NVEC; ++v) sc += dot4(cbar[v], cbar[v]);
    sc = warp_sum(sc);
    if (lane == 0) part_c[warp] = sc;
    for (int j0 = 0; j0 < J; j0 += kRowBatch) {
      float4 x[kRowBatch][NVEC];
#pragma unroll
      for (int r = 0; r < kRowBatch; ++r) {
        const int j = j0 + r;
        const int row = (j == 0) ? 0 : p.C + j - 1;
#pragma unroll
        for (int v = 0; v < NVEC; ++v) {
          const int c4 = v * T + tid;
          if (j < J && c4 < p.N4)
            x[r][v] = ld4(H + (size_t(row) * p.B + b) * p.N + c4 * 4);
          else
            x[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int r = 0; r < kRowBatch; ++r) {
        const int j = j0 + r;
        if (j < J) {
          float sx = 0.f, px = 0.f;
#pragma unroll
          for (int v = 0; v < NVEC; ++v) { sx += dot4(x[r][v], x[r][v]); px += dot4(cbar[v], x[r][v]); }
          sx = warp_sum(sx); px = warp_sum(px);
          if (lane == 0) { part[(j * 2 + 0) * nw + warp] = sx; part[(j * 2 + 1) * nw + warp] = px; }
        }
      }
    }
    __syncthreads();
    for (int e = tid; e < 2 * J + 1; e += T) {
      float s = 0.f;
      if (e < 2 * J) { for (int w = 0; w < nw; ++w) s += part[e * nw + w]; }
      else           { for (int w = 0; w < nw; ++w) s += part_c[w]; }
      fin[e] = s;
      // stats layout: [s_c, s_t, p_t, s_n1, p_n1, ...]
      stats[size_t(b) * p.stride + (e < 2 * J ? 1 + e : 0)] = s;
    }
    __syncthreads();
    if (tid == 0) {
      // normalisation_layer.cpp:36-59: r = pow(s, .5) + eps ; y = x / r.  scores = <c^, x^>
      const float nc = sqrtf(fin[2 * J]) + p.eps;
      const float st = fin[1] / (nc * (sqrtf(fin[0]) + p.eps));
      float loss = 0.f, viol = 0.f;
      for (int k = 1; k <= p.Nn; ++k) {
        const float sn = fin[2 * k + 1] / (nc * (sqrtf(fin[2 * k]) + p.eps));
        const float delta = st - sn;                       // caffe_sub :69
        if (delta < 0.f) viol += 1.f;
        const float h = fmaxf(0.f, p.margin - delta);
        loss += (p.norm == 2) ? h * h : fabsf(h);
        if (tscore) tscore[size_t(b) * p.Nn + k - 1] = st;  // sum_true replicates to Nn columns
        if (nscore) nscore[size_t(b) * p.Nn + k - 1] = sn;
      }
      if (item_loss) item_loss[b] = loss;
      if (item_viol) item_viol[b] = viol;
    }
    __syncthreads();
  }
}

// loss = sum_b item_loss[b] / (B*Nn) ; violations = sum_b item_viol[b].  Deterministic.
__global__ void __launch_bounds__(1024)
rank_loss_reduce_kernel(const float* __restrict__ item_loss, const float* __restrict__ item_viol, int B,
                        float inv_count, float* __restrict__ loss, float* __restrict__ viol) {
  __shared__ float s1[32], s2[32];
  float a = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) { a += item_loss[i]; c += item_viol[i]; }
  a = warp_sum(a); c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = a; s2[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x < 32) {
    a = (threadIdx.x < (blockDim.x >> 5)) ? s1[threadIdx.x] : 0.f;
    c = (threadIdx.x < (blockDim.x >> 5)) ? s2[threadIdx.x] : 0.f;
    a = warp_sum(a); c = warp_sum(c);
    if (threadIdx.x == 0) { if (loss) *loss = a * inv_count; if (viol) *viol = c; }
  }
}

struct BwdOut {
  float* dZ; float* hi; float* lo; uint16_t* bf; int prec; size_t count;
};

__device__ __forceinline__ void store_row4(const BwdOut& o, size_t off, const float4& v, float scale, float& amax) {
  if (o.dZ) stg_stream(reinterpret_cast<float4*>(o.dZ + off), v);
  if (o.prec == VV_PREC_TF32X3) {
    store_x3(o.hi, o.lo, o.count, off, v);
  } else if (o.prec == VV_PREC_F16X3) {
    store_f16x3(o.hi, o.lo, off, v, scale, amax);
  } else if (o.prec == VV_PREC_BF16) {
    uint2 pk = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    *reinterpret_cast<uint2*>(o.bf + off) = pk;
  }
}

// OUT bit 0: fp32 dZ; 2 = TF32X3 operand copy, 4 = BF16 operand copy, 8 = F16X3 operand copy; OUT < 0: decide at run time
template <int OUT>
__device__ __forceinline__ void store_row4_t(const BwdOut& o, size_t off, const float4& v, float scale, float& amax) {
  if (OUT < 0) { store_row4(o, off, v, scale, amax); return; }
  if (OUT & 1) stg_stream(reinterpret_cast<float4*>(o.dZ + off), v);
  if (OUT & 2) store_x3(o.hi, o.lo, o.count, off, v);
  if (OUT & 8) store_f16x3(o.hi, o.lo, off, v, scale, amax);
  if (OUT & 4) {
    uint2 pk = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    *reinterpret_cast<uint2*>(o.bf + off) = pk;
  }
}

template <int NVEC>
__global__ void __launch_bounds__(256)
rank_bwd_kernel(const float* __restrict__ H, const RankDev p, const float* __restrict__ stats,
                const float gscale /* lw*2/(B*Nn) for L2, lw/(B*Nn) for L1 */,
                const int act_fused, const float dscale, const BwdOut out, float* __restrict__ db_accum,
                const float* __restrict__ delta, float* __restrict__ dq_accum) {
  extern __shared__ float sm[];
  const int T = blockDim.x, tid = threadIdx.x;
  const int J = 1 + p.Nn;
  float* s_s = sm;            // [J] |x|^2
  float* s_p = s_s + J;       // [J] <cbar,x>
  float* s_w = s_p + J;       // [J] weight of branch j on c^ : w_0 = -sum g, w_k = g_k
  float* cA = s_w + J;        // [J] coefficient on cbar
  float* cB = cA + J;         // [J] coefficient on x
  float* cE = cB + J;         // [J] w_j / n_j   (d c^ = sum_j cE_j x_j)
  float* sc = cE + J;         // [4] s_c, nc, Fs, Fc
  // dbacc: column sums of dZ (bias gradient); dqacc: sum_m delta[m] * dZ[m,:] (the gather-fused wgrad's
  // correction of dW[:, K-1] for the K-1 copy quirk; delta is non-zero on same-video negative rows only)
  float4 dbacc[NVEC], dqacc[NVEC];
#pragma unroll
  for (int v = 0; v < NVEC; ++v) { dbacc[v] = make_float4(0.f, 0.f, 0.f, 0.f); dqacc[v] = make_float4(0.f, 0.f, 0.f, 0.f); }
  const float oscale = (out.prec == VV_PREC_F16X3) ? f16_hdr(out.hi)->scale : 1.f;
  float amax = 0.f;

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    const float* st = stats + size_t(b) * p.stride;
    // ---- per-item scalar prologue (parallel over the J branches)
    for (int j = tid; j < J; j += T) { s_s[j] = st[1 + 2 * j]; s_p[j] = st[2 + 2 * j]; }
    if (tid == 0) { const float s = st[0]; sc[0] = s; sc[1] = sqrtf(s) + p.eps; }
    __syncthreads();
    const float nc = sc[1];
    const float score_t = s_p[0] / (nc * (sqrtf(s_s[0]) + p.eps));
    for (int k = 1 + tid; k < J; k += T) {
      const float sn = s_p[k] / (nc * (sqrtf(s_s[k]) + p.eps));
      const float h = fmaxf(0.f, p.margin - (score_t - sn));
      // max_margin_loss_layer.cpp:149-192: L2 g = h * (lw*2/count); L1 g = [h>0] * lw/count
      s_w[k] = (p.norm == 2) ? h * gscale : (h > 0.f ? gscale : 0.f);
    }
    __syncthreads();
    if (tid == 0) {
      float g = 0.f;                       // sum_layer.cpp:65-68 gemv over the Nn replicated columns;
      for (int k = 1; k < J; ++k) g += s_w[k];
      s_w[0] = -g;                         // d s+ = -1 * d s- (axpby, :210-212)
    }
    __syncthreads();
    for (int j = tid; j < J; j += T) {
      const float s = s_s[j], pj = s_p[j], w = s_w[j];
      const float nj = sqrtf(s) + p.eps;
      const float q = powf(s, 1.5f) + p.eps;         // normalization_layer.cpp:101-107
      const float aj = w * pj / nc;                  // a = <x, w c^>
      cA[j] = s * w / (nc * q);                      // (s * w c^) / q, c^ = cbar / nc
      cB[j] = -aj / q;
      cE[j] = w / nj;
    }
    __syncthreads();
    if (tid == 0) {
      float ac = 0.f;                                // a_c = <cbar, d c^> = sum_j cE_j p_j (split order: true, neg_1..)
      for (int j = 0; j < J; ++j) ac += cE[j] * s_p[j];
      const float s = sc[0];
      const float q = powf(s, 1.5f) + p.eps;
      sc[2] = s / q; sc[3] = -ac / q;
    }
    // ---- context mean (needed for every output row)
    float4 cbar[NVEC], D[NVEC];
    context_mean<NVEC>(H, p, b, tid, T, cbar);
#pragma unroll
    for (int v = 0; v < NVEC; ++v) D[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    // ---- target + negative rows: dx = cA*cbar + cB*x ; D += cE*x
    for (int j0 = 0; j0 < J; j0 += kRowBatch) {
      float4 x[kRowBatch][NVEC];
#pragma unroll
      for (int r = 0; r < kRowBatch; ++r) {
        const int j = j0 + r;
        const int row = (j == 0) ? 0 : p.C + j - 1;
#pragma unroll
        for (int v = 0; v < NVEC; ++v) {
          const int c4 = v * T + tid;
          if (j < J && c4 < p.N4) x[r][v] = ld4(H + (size_t(row) * p.B + b) * p.N + c4 * 4);
          else x[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int r = 0; r < kRowBatch; ++r) {
        const int j = j0 + r;
        if (j < J) {
          const int row = (j == 0) ? 0 : p.C + j - 1;
          const float a = cA[j], bb = cB[j], e = cE[j];
          const float dl = delta ? delta[size_t(row) * p.B + b] : 0.f;
#pragma unroll
          for (int v = 0; v < NVEC; ++v) {
            const int c4 = v * T + tid;
            if (c4 < p.N4) {
              const float4 xv = x[r][v];
              float4 o;
              o.x = fmaf(a, cbar[v].x, bb * xv.x); o.y = fmaf(a, cbar[v].y, bb * xv.y);
              o.z = fmaf(a, cbar[v].z, bb * xv.z); o.w = fmaf(a, cbar[v].w, bb * xv.w);
              D[v].x = fmaf(e, xv.x, D[v].x); D[v].y = fmaf(e, xv.y, D[v].y);
              D[v].z = fmaf(e, xv.z, D[v].z); D[v].w = fmaf(e, xv.w, D[v].w);
              if (act_fused) {   // dZ = dH * mask*scale * [Z>0]  <=>  dH * scale * [H>0]
                o.x = xv.x > 0.f ? o.x * dscale : 0.f; o.y = xv.y > 0.f ? o.y * dscale : 0.f;
                o.z = xv.z > 0.f ? o.z * dscale : 0.f; o.w = xv.w > 0.f ? o.w * dscale : 0.f;
              }
              dbacc[v].x += o.x; dbacc[v].y += o.y; dbacc[v].z += o.z; dbacc[v].w += o.w;
              dqacc[v].x = fmaf(dl, o.x, dqacc[v].x); dqacc[v].y = fmaf(dl, o.y, dqacc[v].y);
              dqacc[v].z = fmaf(dl, o.z, dqacc[v].z); dqacc[v].w = fmaf(dl, o.w, dqacc[v].w);
              store_row4(out, (size_t(row) * p.B + b) * p.N + c4 * 4, o, oscale, amax);
            }
          }
        }
      }
    }
    __syncthreads();   // sc[2], sc[3] visible
    // ---- context rows: d cbar = (s_c * d c^ - cbar * a_c) / q_c ; d c_i = coeff_i * d cbar
    const float Fs = sc[2], Fc = sc[3];
    float4 dcb[NVEC];
#pragma unroll
    for (int v = 0; v < NVEC; ++v) {
      dcb[v].x = fmaf(Fs, D[v].x, Fc * cbar[v].x); dcb[v].y = fmaf(Fs, D[v].y, Fc * cbar[v].y);
      dcb[v].z = fmaf(Fs, D[v].z, Fc * cbar[v].z); dcb[v].w = fmaf(Fs, D[v].w, Fc * cbar[v].w);
    }
    for (int i = 1; i < p.C; ++i) {
      const float a = p.coeff[i - 1];
#pragma unroll
      for (int v = 0; v < NVEC; ++v) {
        const int c4 = v * T + tid;
        if (c4 < p.N4) {
          const size_t off = (size_t(i) * p.B + b) * p.N + c4 * 4;
          float4 o = make_float4(a * dcb[v].x, a * dcb[v].y, a * dcb[v].z, a * dcb[v].w);
          if (act_fused) {
            const float4 xv = ld4(H + off);     // re-read (L2 hit): the ReLU/dropout gate of this row
            o.x = xv.x > 0.f ? o.x * dscale : 0.f; o.y = xv.y > 0.f ? o.y * dscale : 0.f;
            o.z = xv.z > 0.f ? o.z * dscale : 0.f; o.w = xv.w > 0.f ? o.w * dscale : 0.f;
          }
          dbacc[v].x += o.x; dbacc[v].y += o.y; dbacc[v].z += o.z; dbacc[v].w += o.w;
          store_row4(out, off, o, oscale, amax);
        }
      }
    }
    __syncthreads();   // smem coefficients are rewritten by the next item
  }
  if (out.prec == VV_PREC_F16X3) f16_publish_absmax(out.hi, amax);
  if (db_accum) {
#pragma unroll
    for (int v = 0; v < NVEC; ++v) {
      const int c4 = v * T + tid;
      if (c4 < p.N4) {
        atomicAdd(db_accum + c4 * 4 + 0, dbacc[v].x); atomicAdd(db_accum + c4 * 4 + 1, dbacc[v].y);
        atomicAdd(db_accum + c4 * 4 + 2, dbacc[v].z); atomicAdd(db_accum + c4 * 4 + 3, dbacc[v].w);
      }
    }
  }
  if (dq_accum) {
#pragma unroll
    for (int v = 0; v < NVEC; ++v) {
      const int c4 = v * T + tid;
      if (c4 < p.N4) {
        atomicAdd(dq_accum + c4 * 4 + 0, dqacc[v].x); atomicAdd(dq_accum + c4 * 4 + 1, dqacc[v].y);
        atomicAdd(dq_accum + c4 * 4 + 2, dqacc[v].z); atomicAdd(dq_accum + c4 * 4 + 3, dqacc[v].w);
      }
    }
  }
}

// Sums NV values across the warp with NV + NV/2 + ... shuffles instead of 5 * NV: at each xor distance a lane keeps
// one half of the (remaining) values and hands the other half to its partner.  The addition tree per value is the
// xor butterfly of warp_sum, so the totals are bit-identical.  On return v[0] of lane l is the warp total of value
// `index` (valid lanes only; every value ends up in exactly one lane).
template <int NV, int N, int OFF>
__device__ __forceinline__ void warp_multi_sum_step(float (&v)[NV], int lane, int& base, int& nvalid) {
  constexpr int half = (N + 1) / 2;
  const bool up = (lane & OFF) != 0;
#pragma unroll
  for (int i = 0; i < half; ++i) {
    const float a = v[i];
    const float b = (i + half < N) ? v[i + half] : 0.f;
    const float recv = __shfl_xor_sync(0xffffffffu, up ? a : b, OFF);
    v[i] = (up ? b : a) + recv;
  }
  const int live = nvalid < N ? nvalid : N;
  if (up) { base += half; nvalid = live > half ? live - half : 0; }
  else    { nvalid = live < half ? live : half; }
  if constexpr (OFF > 1) warp_multi_sum_step<NV, half, OFF / 2>(v, lane, base, nvalid);
}
template <int NV>
__device__ __forceinline__ void warp_multi_sum(float (&v)[NV], int lane, int& index, bool& valid) {
  int base = 0, nvalid = NV;
  warp_multi_sum_step<NV, NV, 16>(v, lane, base, nvalid);
  index = base; valid = nvalid >= 1;
}

// Folded rank_loss_reduce_kernel: the CTA that takes the last ticket sums item_loss / item_viol over the batch in a
// fixed order (thread-strided partials, xor-butterfly per warp, warps in index order) -> deterministic like the
// separate launch, without it.  `counter` must be zero at launch.  Called by the T compute threads of every CTA.
template <bool NAMED_BARRIER>
__device__ __forceinline__ void last_cta_loss_reduce(unsigned int* counter, const float* item_loss, const float* item_viol,
                                                     int B, float inv_count, float* loss, float* viol, float* red /*[66]*/,
                                                     int tid, int T) {
  __threadfence();                                   // this CTA's item_loss / item_viol stores (thread 0) are visible
  if (tid == 0) red[64] = __uint_as_float(atomicAdd(counter, 1u));
  if (NAMED_BARRIER) asm volatile("bar.sync 1, %0;" :: "r"(T) : "memory"); else __syncthreads();
  if (__float_as_uint(red[64]) != gridDim.x - 1) return;
  __threadfence();
  float a = 0.f, c = 0.f;
  for (int i = tid; i < B; i += T) { a += __ldcg(item_loss + i); c += __ldcg(item_viol + i); }
  a = warp_sum(a); c = warp_sum(c);
  if ((tid & 31) == 0) { red[tid >> 5] = a; red[32 + (tid >> 5)] = c; }
  if (NAMED_BARRIER) asm volatile("bar.sync 1, %0;" :: "r"(T) : "memory"); else __syncthreads();
  if (tid == 0) {
    a = 0.f; c = 0.f;
    for (int w = 0; w < (T >> 5); ++w) { a += red[w]; c += red[32 + w]; }
    if (loss) *loss = a * inv_count;
    if (viol) *viol = c;
  }
}

// ---- K2+K3 in one pass -------------------------------------------------------------------------------------
// All R rows of an item live in registers (R*N*4 bytes are read from HBM exactly once), the 2J+1 dot products
// are reduced through shared memory, warp 0 evaluates the per-item scalar chain (scores, hinge, every backward
// coefficient) with one lane per branch, and the gradient rows are produced from the registers: algorithmic
// traffic = R*N*4 read + the dZ operand write.  The reduction trees, accumulation orders and formulas are the
// those of rank_fwd_kernel / rank_bwd_kernel above (results agree to rounding, ~1e-7 relative).
// Needs nvec == 1 (N <= 1024), R <= RMAX and J <= 32; otherwise the two-kernel path runs.
// CT / NNT > 0: context size / negatives fixed at compile time (the per-row role tests fold away); OUT: store mode.
template <int RMAX, int CT, int NNT, int OUT>
__global__ void __launch_bounds__(256)
rank_fused_kernel(const float* __restrict__ H, const RankDev p, const float gscale, const int act_fused,
                  const float dscale, const BwdOut out, float* __restrict__ db_accum,
                  const float* __restrict__ delta, float* __restrict__ dq_accum,
                  float* __restrict__ stats, float* __restrict__ tscore, float* __restrict__ nscore,
                  float* __restrict__ item_loss, float* __restrict__ item_viol,
                  unsigned int* __restrict__ done_counter, const float inv_count, float* __restrict__ loss_out,
                  float* __restrict__ viol_out) {
  extern __shared__ float sm[];
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = T >> 5;
  const int Cc = CT > 0 ? CT : p.C, Nn = NNT > 0 ? NNT : p.Nn;
  const int J = 1 + Nn, R = Cc + Nn;
  const int per = (2 * J + 1) * nw + 4 * J + 2;        // floats per smem buffer (double buffered by item parity)
  const bool col_ok = tid < p.N4;
  float4 dbacc = make_float4(0.f, 0.f, 0.f, 0.f), dqacc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float oscale = (out.prec == VV_PREC_F16X3) ? f16_hdr(out.hi)->scale : 1.f;
  float amax = 0.f;
  int parity = 0;
  for (int b = blockIdx.x; b < p.B; b += gridDim.x, parity ^= 1) {
    float* part = sm + parity * per;                   // [2J+1][nw]: (s_x, p_x) per branch, then s_c
    float* cA = part + (2 * J + 1) * nw;               // [J] coefficient on cbar
    float* cB = cA + J;                                // [J] coefficient on x
    float* cE = cB + J;                                // [J] w_j / n_j
    float* cD = cE + J;                                // [J] K-1 quirk correction of the branch's row (0 without a plan)
    float* sc = cD + J;                                // Fs, Fc
    // ---- one load phase: R independent 128-bit loads per thread
    float4 x[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r)
      x[r] = (r < R && col_ok) ? ld4(H + (size_t(r) * p.B + b) * p.N + tid * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    // ---- context mean, bottom order (eltwise_layer.cpp:67-73)
    float4 cbar = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 1; r < RMAX; ++r) {
      if (r < Cc) {
        const float a = p.coeff[r - 1];
        cbar.x = fmaf(a, x[r].x, cbar.x); cbar.y = fmaf(a, x[r].y, cbar.y);
        cbar.z = fmaf(a, x[r].z, cbar.z); cbar.w = fmaf(a, x[r].w, cbar.w);
      }
    }
    if (NNT > 0) {
      constexpr int NV = 2 * (NNT > 0 ? 1 + NNT : 1) + 1;
      float v[NV];
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        if (r < R && (r == 0 || r >= Cc)) {
          const int j = (r == 0) ? 0 : r - Cc + 1;
          v[2 * j] = dot4(x[r], x[r]); v[2 * j + 1] = dot4(cbar, x[r]);
        }
      }
      v[NV - 1] = dot4(cbar, cbar);
      int e; bool ok;
      warp_multi_sum<NV>(v, lane, e, ok);
      if (ok) part[e * nw + warp] = v[0];
    } else {
      {
        float s = warp_sum(dot4(cbar, cbar));
        if (lane == 0) part[2 * J * nw + warp] = s;
      }
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        if (r < R && (r == 0 || r >= Cc)) {             // uniform branch
          const int j = (r == 0) ? 0 : r - Cc + 1;
          float sx = warp_sum(dot4(x[r], x[r])), px = warp_sum(dot4(cbar, x[r]));
          if (lane == 0) { part[(j * 2 + 0) * nw + warp] = sx; part[(j * 2 + 1) * nw + warp] = px; }
        }
      }
    }
    __syncthreads();
    // ---- per-item scalars on warp 0, lane j = branch j
    if (warp == 0) {
      float s_c = 0.f;
      for (int w = 0; w < nw; ++w) s_c += part[2 * J * nw + w];
      float sj = 0.f, pj = 0.f;
      if (lane < J) {
        for (int w = 0; w < nw; ++w) sj += part[(lane * 2 + 0) * nw + w];
        for (int w = 0; w < nw; ++w) pj += part[(lane * 2 + 1) * nw + w];
      }
      if (stats) {
        float* st = stats + size_t(b) * p.stride;
        if (lane == 0) st[0] = s_c;
        if (lane < J) { st[1 + 2 * lane] = sj; st[2 + 2 * lane] = pj; }
      }
      // normalization_layer.cpp:36-59: r = pow(s, .5) + eps ; y = x / r.  scores = <c^, x^>
      const float nc = sqrtf(s_c) + p.eps;
      const float nj = sqrtf(sj) + p.eps;
      const float score = pj / (nc * nj);
      const float score_t = __shfl_sync(0xffffffffu, score, 0);
      const float dlt = score_t - score;                                  // caffe_sub :69
      const float h = fmaxf(0.f, p.margin - dlt);
      const bool neg = lane >= 1 && lane < J;
      // max_margin_loss_layer.cpp:149-192: L2 g = h * (lw*2/count); L1 g = [h>0] * lw/count
      float w = neg ? ((p.norm == 2) ? h * gscale : (h > 0.f ? gscale : 0.f)) : 0.f;
      const float vterm = (neg && dlt < 0.f) ? 1.f : 0.f;
      // sums over the negatives as xor-butterfly reductions (the reference sums serially; the order differs, the
      // three independent butterflies overlap instead of ~30 dependent shuffles on the item's critical path)
      const float loss = warp_sum(neg ? ((p.norm == 2) ? h * h : fabsf(h)) : 0.f);
      const float viol = warp_sum(vterm);
      const float g = warp_sum(w);
      if (lane == 0) w = -g;                                              // d s+ = -1 * d s- (axpby, :210-212)
      if (neg) {
        if (tscore) tscore[size_t(b) * Nn + lane - 1] = score_t;        // sum_true replicates to Nn columns
        if (nscore) nscore[size_t(b) * Nn + lane - 1] = score;
      }
      if (lane == 0) { if (item_loss) item_loss[b] = loss; if (item_viol) item_viol[b] = viol; }
      const float q = powf(sj, 1.5f) + p.eps;                             // normalization_layer.cpp:101-107
      const float aj = w * pj / nc;
      const float e = w / nj;
      if (lane < J) {
        cA[lane] = sj * w / (nc * q); cB[lane] = -aj / q; cE[lane] = e;
        cD[lane] = delta ? delta[size_t(lane == 0 ? 0 : Cc + lane - 1) * p.B + b] : 0.f;
      }
      const float ac = warp_sum(lane < J ? e * pj : 0.f);                 // a_c = <cbar, d c^>
      if (lane == 0) { const float qc = powf(s_c, 1.5f) + p.eps; sc[0] = s_c / qc; sc[1] = -ac / qc; }
    }
    __syncthreads();
    // ---- target + negative rows: dx = cA*cbar + cB*x ; D += cE*x
    float4 D = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      if (r < R && (r == 0 || r >= Cc)) {
        const int j = (r == 0) ? 0 : r - Cc + 1;
        const float a = cA[j], bb = cB[j], e = cE[j];
        if (col_ok) {
          const float4 xv = x[r];
          float4 o;
          o.x = fmaf(a, cbar.x, bb * xv.x); o.y = fmaf(a, cbar.y, bb * xv.y);
          o.z = fmaf(a, cbar.z, bb * xv.z); o.w = fmaf(a, cbar.w, bb * xv.w);
          D.x = fmaf(e, xv.x, D.x); D.y = fmaf(e, xv.y, D.y); D.z = fmaf(e, xv.z, D.z); D.w = fmaf(e, xv.w, D.w);
          if (act_fused) {   // dZ = dH * mask*scale * [Z>0]  <=>  dH * scale * [H>0]
            o.x = xv.x > 0.f ? o.x * dscale : 0.f; o.y = xv.y > 0.f ? o.y * dscale : 0.f;
            o.z = xv.z > 0.f ? o.z * dscale : 0.f; o.w = xv.w > 0.f ? o.w * dscale : 0.f;
          }
          dbacc.x += o.x; dbacc.y += o.y; dbacc.z += o.z; dbacc.w += o.w;
          const float dl = cD[j];
          if (dl != 0.f) {                                                // uniform; only rows hit by the K-1 copy quirk
            dqacc.x = fmaf(dl, o.x, dqacc.x); dqacc.y = fmaf(dl, o.y, dqacc.y);
            dqacc.z = fmaf(dl, o.z, dqacc.z); dqacc.w = fmaf(dl, o.w, dqacc.w);
          }
          store_row4_t<OUT>(out, (size_t(r) * p.B + b) * p.N + tid * 4, o, oscale, amax);
        }
      }
    }
    // ---- context rows: d cbar = (s_c * d c^ - cbar * a_c) / q_c ; d c_i = coeff_i * d cbar
    const float Fs = sc[0], Fc = sc[1];
    float4 dcb;
    dcb.x = fmaf(Fs, D.x, Fc * cbar.x); dcb.y = fmaf(Fs, D.y, Fc * cbar.y);
    dcb.z = fmaf(Fs, D.z, Fc * cbar.z); dcb.w = fmaf(Fs, D.w, Fc * cbar.w);
#pragma unroll
    for (int r = 1; r < RMAX; ++r) {
      if (r < Cc && col_ok) {
        const float a = p.coeff[r - 1];
        float4 o = make_float4(a * dcb.x, a * dcb.y, a * dcb.z, a * dcb.w);
        if (act_fused) {
          const float4 xv = x[r];
          o.x = xv.x > 0.f ? o.x * dscale : 0.f; o.y = xv.y > 0.f ? o.y * dscale : 0.f;
          o.z = xv.z > 0.f ? o.z * dscale : 0.f; o.w = xv.w > 0.f ? o.w * dscale : 0.f;
        }
        dbacc.x += o.x; dbacc.y += o.y; dbacc.z += o.z; dbacc.w += o.w;
        store_row4_t<OUT>(out, (size_t(r) * p.B + b) * p.N + tid * 4, o, oscale, amax);
      }
    }
    // no trailing barrier: the next item uses the other smem buffer, and the barrier after its reductions
    // orders this item's coefficient reads before the buffer is written again two items later
  }
  if (out.prec == VV_PREC_F16X3) f16_publish_absmax(out.hi, amax);
  if (col_ok) {
    if (db_accum) {
      atomicAdd(db_accum + tid * 4 + 0, dbacc.x); atomicAdd(db_accum + tid * 4 + 1, dbacc.y);
      atomicAdd(db_accum + tid * 4 + 2, dbacc.z); atomicAdd(db_accum + tid * 4 + 3, dbacc.w);
    }
    if (dq_accum) {
      atomicAdd(dq_accum + tid * 4 + 0, dqacc.x); atomicAdd(dq_accum + tid * 4 + 1, dqacc.y);
      atomicAdd(dq_accum + tid * 4 + 2, dqacc.z); atomicAdd(dq_accum + tid * 4 + 3, dqacc.w);
    }
  }
  if (done_counter) {
    __shared__ float red[66];
    __syncthreads();                                 // every item of this CTA has been written
    last_cta_loss_reduce<false>(done_counter, item_loss, item_viol, p.B, inv_count, loss_out, viol_out, red, tid, T);
  }
}

// ---- K2+K3 in one pass, second generation ------------------------------------------------------------------
// Same data flow as rank_fused_kernel (the R rows of an item in registers, one read of H, one write of the dZ operand)
// for the trainer's hot case -- operand-only output (no fp32 dZ, no score blobs), nvec == 1, R <= 16 -- with the
// instruction count cut to what the arithmetic needs (the first kernel is issue-bound at ~1 900 instructions per warp
// and item):
//   * packed fp32 pairs (FFMA2 / FMUL2 / FADD2, sm_100) for every element-wise product;
//   * the operand scale and the dropout scale are folded into the per-row coefficients by the scalar chain, so a gradient
//     element is fma2(A', cbar, B' * x) * gate and goes to the fp16 / bf16 conversion as it is; db, dq and max|dZ| are
//     accumulated in the scaled domain and unscaled once per CTA (the operand scale is a power of two: exact);
//   * the ReLU/dropout gate [H > 0] as a 0/1 float (one FSET per element) multiplied in packed form;
//   * every warp evaluates the per-item scalar chain itself from the shared partial sums (no second barrier, no warp idles
//     behind warp 0's sqrt / pow / divide chain);
//   * db / dq without float atomics: every CTA leaves its column sums in a workspace, the last CTA of each group of
//     kRankGroup CTAs adds its group's in CTA order, the last group adds the group sums in group order -> run-to-run
//     deterministic like the reference's gemv (inner_product_layer.cpp:93-96).  The same final CTA adds up the batch loss.
// Formulas, reduction trees and the order of operations that matters for rounding are those of rank_fused_kernel; the
// folded scales change the last bit of individual gradient elements (both are within 1e-6 of the two-kernel path).
constexpr int kRankGroup = 24;

__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float dot4p(const float4& a, const float4& b) {          // (a.x b.x + a.z b.z) + (a.y b.y + a.w b.w)
  const float2 t = __ffma2_rn(hi2(a), hi2(b), __fmul2_rn(lo2(a), lo2(b)));
  return t.x + t.y;
}
// 1.0f where x > 0 else 0.0f (H is post-ReLU/dropout: never negative)
__device__ __forceinline__ float gate1(float x) {
  float g;
  asm("set.gt.f32.f32 %0, %1, 0f00000000;" : "=f"(g) : "f"(x));
  return g;
}
__device__ __forceinline__ float amax4(float m, const float4& v) {
  return fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
}
// operand store of 4 consecutive elements that are ALREADY scaled (F16X3) -- OUT: 2 TF32X3, 4 BF16, 8 F16X3.
// p_hi / p_lo: this thread's byte addresses in row 0 of the two planes; eoff: the row's element offset from there.
template <int OUT>
__device__ __forceinline__ void store_op4_at(const BwdOut& o, char* p_hi, char* p_lo, unsigned int eoff, const float4& v) {
  if (OUT == 8) {
    const uint32_t p01 = pack_f16x2_sat(v.x, v.y), p23 = pack_f16x2_sat(v.z, v.w);
    const float2 f01 = unpack_f16x2(p01), f23 = unpack_f16x2(p23);
    const float2 r01 = __fadd2_rn(lo2(v), make_float2(-f01.x, -f01.y)), r23 = __fadd2_rn(hi2(v), make_float2(-f23.x, -f23.y));
    *reinterpret_cast<uint2*>(p_hi + size_t(eoff) * 2) = make_uint2(p01, p23);
    *reinterpret_cast<uint2*>(p_lo + size_t(eoff) * 2) = make_uint2(pack_f16x2_sat(r01.x, r01.y), pack_f16x2_sat(r23.x, r23.y));
  } else if (OUT == 4) {
    *reinterpret_cast<uint2*>(p_hi + size_t(eoff) * 2) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  } else {
    float4 h;
    h.x = to_tf32_rna(v.x); h.y = to_tf32_rna(v.y); h.z = to_tf32_rna(v.z); h.w = to_tf32_rna(v.w);
    *reinterpret_cast<float4*>(p_hi + size_t(eoff) * 4) = h;
    *reinterpret_cast<uint2*>(p_lo + size_t(eoff) * 2) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    *reinterpret_cast<uint2*>(p_lo + (size_t(eoff) + o.count) * 2) =
        make_uint2(pack_bf16x2(v.x - h.x, v.y - h.y), pack_bf16x2(v.z - h.z, v.w - h.w));
  }
}

struct RankWs {            // deterministic column sums (db, dq): layout of the workspace, in floats
  unsigned int* tickets;   // [groups + 1], zero at first use, self-resetting
  float* cta_part;         // [grid][2][N]
  float* grp_part;         // [groups][2][N]
};

// Deterministic column sums (db, dq) and batch loss of a rank kernel with one float4 column group per thread: every CTA
// stores its partials, the last CTA of each group of kRankGroup sums them in CTA order, the last group sums the groups
// in group order and then the item losses in a fixed order.  Tickets reset themselves.  Called by all T threads.
__device__ __forceinline__ void column_sums_finish(const RankWs& ws, const RankDev& p, const float4& dbacc, const float4& dqacc,
                                                   bool col_ok, float* db_accum, float* dq_accum, const float* item_loss,
                                                   const float* item_viol, float inv_count, float* loss_out, float* viol_out,
                                                   int tid, int T) {
  __shared__ unsigned int s_flag;
  __shared__ float red[66];
  const int N4 = p.N4;
  const int grp = blockIdx.x / kRankGroup, ngroups = (gridDim.x + kRankGroup - 1) / kRankGroup;
  const int g0 = grp * kRankGroup, gsize = min(kRankGroup, int(gridDim.x) - g0);
  if (col_ok) {
    float4* mine = reinterpret_cast<float4*>(ws.cta_part) + size_t(blockIdx.x) * 2 * N4;
    mine[tid] = dbacc; mine[N4 + tid] = dqacc;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int prev = atomicAdd(&ws.tickets[grp], 1u);
    const bool last = prev + 1u == unsigned(gsize);
    if (last) ws.tickets[grp] = 0u;
    s_flag = last ? 1u : 0u;
  }
  __syncthreads();
  if (s_flag == 0u) return;
  __threadfence();
  if (col_ok) {
    const float4* base = reinterpret_cast<const float4*>(ws.cta_part) + size_t(g0) * 2 * N4;
    float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), sq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int c = 0; c < gsize; ++c) {
      const float4 a = __ldcg(base + size_t(c) * 2 * N4 + tid), q = __ldcg(base + size_t(c) * 2 * N4 + N4 + tid);
      sb.x += a.x; sb.y += a.y; sb.z += a.z; sb.w += a.w; sq.x += q.x; sq.y += q.y; sq.z += q.z; sq.w += q.w;
    }
    float4* gp = reinterpret_cast<float4*>(ws.grp_part) + size_t(grp) * 2 * N4;
    gp[tid] = sb; gp[N4 + tid] = sq;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int prev = atomicAdd(&ws.tickets[ngroups], 1u);
    const bool last = prev + 1u == unsigned(ngroups);
    if (last) ws.tickets[ngroups] = 0u;
    s_flag = last ? 1u : 0u;
  }
  __syncthreads();
  if (s_flag == 0u) return;
  __threadfence();
  if (col_ok) {
    const float4* base = reinterpret_cast<const float4*>(ws.grp_part);
    float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), sq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int c = 0; c < ngroups; ++c) {
      const float4 a = __ldcg(base + size_t(c) * 2 * N4 + tid), q = __ldcg(base + size_t(c) * 2 * N4 + N4 + tid);
      sb.x += a.x; sb.y += a.y; sb.z += a.z; sb.w += a.w; sq.x += q.x; sq.y += q.y; sq.z += q.z; sq.w += q.w;
    }
    if (db_accum) reinterpret_cast<float4*>(db_accum)[tid] = sb;
    if (dq_accum) reinterpret_cast<float4*>(dq_accum)[tid] = sq;
  }
  if (loss_out || viol_out) {                           // the batch loss / violation count, fixed order (as last_cta_loss_reduce)
    float a = 0.f, c = 0.f;
    for (int i = tid; i < p.B; i += T) { a += __ldcg(item_loss + i); c += __ldcg(item_viol + i); }
    a = warp_sum(a); c = warp_sum(c);
    if ((tid & 31) == 0) { red[tid >> 5] = a; red[32 + (tid >> 5)] = c; }
    __syncthreads();
    if (tid == 0) {
      a = 0.f; c = 0.f;
      for (int w = 0; w < (T >> 5); ++w) { a += red[w]; c += red[32 + w]; }
      if (loss_out) *loss_out = a * inv_count;
      if (viol_out) *viol_out = c;
    }
  }
}

// FULL: blockDim.x == N/4 (every thread owns a column group: no bounds checks, no zero fill); NWT > 0: warps per CTA fixed
// at compile time (4 = the N = 512 case: the per-branch partial sums are one 128-bit shared-memory load)
// OCC: resident CTAs per SM the 128-thread specialisation is compiled for (4: 128 registers, 5: 96, 6: 80 with spills)
template <int CT, int NNT, int OUT, bool FULL, int NWT, int OCC = 4>
__global__ void __launch_bounds__(NWT == 4 ? 128 : 256, NWT == 4 ? OCC : 2)
rank_fused2_kernel(const float* __restrict__ H, const RankDev p, const float gscale, const float dscale,
                   const BwdOut out, float* __restrict__ db_accum, const float* __restrict__ delta,
                   float* __restrict__ dq_accum, float* __restrict__ stats, float* __restrict__ item_loss,
                   float* __restrict__ item_viol, const RankWs ws, unsigned int* __restrict__ done_counter,
                   const float inv_count, float* __restrict__ loss_out, float* __restrict__ viol_out) {
  constexpr int RMAX = 16;
  extern __shared__ __align__(16) float sm[];
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = NWT > 0 ? NWT : (T >> 5);
  const int Cc = CT > 0 ? CT : p.C, Nn = NNT > 0 ? NNT : p.Nn;
  const int J = 1 + Nn, R = Cc + Nn;
  const int per = (2 * J + 1) * nw;                    // floats per partial-sum buffer (double buffered by item parity)
  const bool col_ok = FULL || tid < p.N4;
  // F16X3: the operand's power-of-two scale; a header that was never set (0) measures with scale 1 (the first, measuring
  // pass of a fresh operand: what it stores is overwritten by the pass that follows vv_operand_rescale)
  const float hdr_scale = (OUT == 8) ? f16_hdr(out.hi)->scale : 1.f;
  const float oscale = hdr_scale > 0.f ? hdr_scale : 1.f;
  const float osc = dscale * oscale;                   // what every stored gradient element is multiplied with
  // row r of item b starts at element (r * B + b) * N: 32-bit element offsets from per-item bases (the launcher checks
  // that the blob has fewer than 2^31 elements), one IMAD.WIDE per row instead of a 64-bit multiply chain
  const unsigned int rs = unsigned(p.B) * unsigned(p.N);
  constexpr int kOpBytes = (OUT == 2) ? 4 : 2;         // element size of operand plane `hi`
  float4 dbacc = make_float4(0.f, 0.f, 0.f, 0.f), dqacc = make_float4(0.f, 0.f, 0.f, 0.f);
  float amax = 0.f;
  int parity = 0;
  for (int b = blockIdx.x; b < p.B; b += gridDim.x, parity ^= 1) {
    float* part = sm + parity * per;                   // [2J+1][nw]: (s_x, p_x) per branch, then s_c
    const unsigned int e0 = unsigned(b) * unsigned(p.N) + unsigned(tid) * 4u;      // element offset of this thread in row 0
    const float* hp = H + e0;
    // ---- one load phase: R independent 128-bit loads per thread
    float4 x[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      if (r < R) {
        if (FULL || col_ok) x[r] = ld4(hp + size_t(unsigned(r) * rs));
        else x[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // the NEXT item's rows on their way into L2 while this one is processed (one bulk prefetch per row, thread r takes row r)
    if (FULL && tid < R && b + int(gridDim.x) < p.B) {
      const float* nx = H + size_t(unsigned(tid) * rs) + size_t(unsigned(b + gridDim.x) * unsigned(p.N));
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(nx), "r"(unsigned(p.N) * 4u) : "memory");
    }
    // ---- context mean, bottom order (eltwise_layer.cpp:67-73)
    float2 cl = make_float2(0.f, 0.f), ch = make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 1; r < RMAX; ++r) {
      if (r < Cc) {
        const float2 a = splat2(p.coeff[r - 1]);
        cl = __ffma2_rn(a, lo2(x[r]), cl); ch = __ffma2_rn(a, hi2(x[r]), ch);
      }
    }
    const float4 cbar = make_float4(cl.x, cl.y, ch.x, ch.y);
    if (NNT > 0) {
      constexpr int NV = 2 * (NNT > 0 ? 1 + NNT : 1) + 1;
      float v[NV];
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        if (r < R && (r == 0 || r >= Cc)) {
          const int j = (r == 0) ? 0 : r - Cc + 1;
          v[2 * j] = dot4p(x[r], x[r]); v[2 * j + 1] = dot4p(cbar, x[r]);
        }
      }
      v[NV - 1] = dot4p(cbar, cbar);
      int e; bool ok;
      warp_multi_sum<NV>(v, lane, e, ok);
      if (ok) part[e * nw + warp] = v[0];
    } else {
      {
        float s = warp_sum(dot4p(cbar, cbar));
        if (lane == 0) part[2 * J * nw + warp] = s;
      }
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        if (r < R && (r == 0 || r >= Cc)) {             // uniform branch
          const int j = (r == 0) ? 0 : r - Cc + 1;
          float sx = warp_sum(dot4p(x[r], x[r])), px = warp_sum(dot4p(cbar, x[r]));
          if (lane == 0) { part[(j * 2 + 0) * nw + warp] = sx; part[(j * 2 + 1) * nw + warp] = px; }
        }
      }
    }
    __syncthreads();
    // ---- per-item scalars, lane j = branch j, in EVERY warp; the coefficients carry the output scale: rA = A * osc,
    // rB = B * osc, Fs / Fc likewise.  Same formulas as rank_fused_kernel with the divisions as correctly rounded
    // reciprocals times the numerator and pow(s, 1.5) as s * sqrt(s) (both within 1.5 ulp of the reference's libm values;
    // the chain is evaluated by every warp, so its length is paid four times over).
    float rA = 0.f, rB = 0.f, rE = 0.f, rD = 0.f, Fs, Fc;
    {
      float s_c = 0.f, sj = 0.f, pj = 0.f;
      if (NWT == 4) {                                                     // [e][4]: one 128-bit load per value, warps in order
        const float4 c4 = *reinterpret_cast<const float4*>(part + 2 * J * 4);
        s_c = ((c4.x + c4.y) + c4.z) + c4.w;
        if (lane < J) {
          const float4 a4 = *reinterpret_cast<const float4*>(part + (lane * 2 + 0) * 4);
          const float4 b4 = *reinterpret_cast<const float4*>(part + (lane * 2 + 1) * 4);
          sj = ((a4.x + a4.y) + a4.z) + a4.w; pj = ((b4.x + b4.y) + b4.z) + b4.w;
        }
      } else {
        for (int w = 0; w < nw; ++w) s_c += part[2 * J * nw + w];
        if (lane < J) {
          for (int w = 0; w < nw; ++w) sj += part[(lane * 2 + 0) * nw + w];
          for (int w = 0; w < nw; ++w) pj += part[(lane * 2 + 1) * nw + w];
        }
      }
      if (stats && warp == 0) {
        float* st = stats + size_t(b) * p.stride;
        if (lane == 0) st[0] = s_c;
        if (lane < J) { st[1 + 2 * lane] = sj; st[2 + 2 * lane] = pj; }
      }
      // normalization_layer.cpp:36-59: r = pow(s, .5) + eps ; y = x / r.  scores = <c^, x^>
      const float rt_c = sqrtf(s_c), rt_j = sqrtf(sj);
      const float inv_nc = __frcp_rn(rt_c + p.eps), inv_nj = __frcp_rn(rt_j + p.eps);
      const float inv_q = __frcp_rn(sj * rt_j + p.eps);                   // 1 / (pow(s, 1.5) + eps), normalization_layer.cpp:101-107
      const float inv_qc = __frcp_rn(s_c * rt_c + p.eps);
      const float u = pj * inv_nj;                                        // <cbar, x^_j>
      const float score = u * inv_nc;
      const float score_t = __shfl_sync(0xffffffffu, score, 0), u0 = __shfl_sync(0xffffffffu, u, 0);
      const float dlt = score_t - score;                                  // caffe_sub :69
      const float h = fmaxf(0.f, p.margin - dlt);
      const bool neg = lane >= 1 && lane < J;
      // max_margin_loss_layer.cpp:149-192: L2 g = h * (lw*2/count); L1 g = [h>0] * lw/count
      float w = neg ? ((p.norm == 2) ? h * gscale : (h > 0.f ? gscale : 0.f)) : 0.f;
      // g = sum_k w_k and sum_k w_k u_k as two interleaved butterflies (one dependent chain of 5 shuffles, not two):
      // a_c = <cbar, d c^> = sum_j (w_j / n_j) p_j with w_0 = -g  =>  a_c = sum_{k>=1} w_k u_k - g u_0
      float g = w, s1 = w * u;
      if (warp == 0) {                                                    // the item's loss terms ride along: once per item
        float ls = neg ? ((p.norm == 2) ? h * h : fabsf(h)) : 0.f, vs = (neg && dlt < 0.f) ? 1.f : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          g += __shfl_xor_sync(0xffffffffu, g, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          ls += __shfl_xor_sync(0xffffffffu, ls, o); vs += __shfl_xor_sync(0xffffffffu, vs, o);
        }
        if (lane == 0) { if (item_loss) item_loss[b] = ls; if (item_viol) item_viol[b] = vs; }
      } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { g += __shfl_xor_sync(0xffffffffu, g, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
      }
      if (lane == 0) w = -g;                                              // d s+ = -1 * d s- (axpby, :210-212)
      const float ac = s1 - g * u0;
      if (lane < J) {
        const float wn = w * inv_nc;
        rA = (sj * wn * inv_q) * osc; rB = (-(wn * pj) * inv_q) * osc; rE = w * inv_nj;
        rD = delta ? delta[size_t(lane == 0 ? 0 : Cc + lane - 1) * p.B + b] : 0.f;
      }
      Fs = (s_c * inv_qc) * osc; Fc = (-ac * inv_qc) * osc;
    }
    const unsigned int qmask = __ballot_sync(0xffffffffu, rD != 0.f);     // branches hit by the K-1 copy quirk (rare)
    // ---- target + negative rows: dx = (A cbar + B x) [H > 0]; D += E x
    float2 Dl = make_float2(0.f, 0.f), Dh = make_float2(0.f, 0.f);
    float2 dbl = lo2(dbacc), dbh = hi2(dbacc);
    char* const o_hi = reinterpret_cast<char*>(OUT == 4 ? static_cast<void*>(out.bf) : static_cast<void*>(out.hi)) + size_t(e0) * kOpBytes;
    char* const o_lo = reinterpret_cast<char*>(out.lo) + size_t(e0) * 2;  // F16X3: h1 plane; TF32X3: bf16 plane 0 (plane 1 `count` elements on)
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      if (r < R && (r == 0 || r >= Cc)) {
        const int j = (r == 0) ? 0 : r - Cc + 1;
        const float a = __shfl_sync(0xffffffffu, rA, j), bb = __shfl_sync(0xffffffffu, rB, j), e = __shfl_sync(0xffffffffu, rE, j);
        const bool hit = (qmask >> j) & 1u;                               // warp-uniform
        const float dl = hit ? __shfl_sync(0xffffffffu, rD, j) : 0.f;     // every lane takes part (not under col_ok)
        if (FULL || col_ok) {
          const float4 xv = x[r];
          const float2 a2 = splat2(a), b2 = splat2(bb), e2 = splat2(e);
          float2 ol = __ffma2_rn(a2, cl, __fmul2_rn(b2, lo2(xv))), oh = __ffma2_rn(a2, ch, __fmul2_rn(b2, hi2(xv)));
          Dl = __ffma2_rn(e2, lo2(xv), Dl); Dh = __ffma2_rn(e2, hi2(xv), Dh);
          ol = __fmul2_rn(ol, make_float2(gate1(xv.x), gate1(xv.y))); oh = __fmul2_rn(oh, make_float2(gate1(xv.z), gate1(xv.w)));
          dbl = __fadd2_rn(dbl, ol); dbh = __fadd2_rn(dbh, oh);
          const float4 o = make_float4(ol.x, ol.y, oh.x, oh.y);
          if (hit) {
            dqacc.x = fmaf(dl, o.x, dqacc.x); dqacc.y = fmaf(dl, o.y, dqacc.y);
            dqacc.z = fmaf(dl, o.z, dqacc.z); dqacc.w = fmaf(dl, o.w, dqacc.w);
          }
          if (OUT == 8) amax = amax4(amax, o);
          store_op4_at<OUT>(out, o_hi, o_lo, unsigned(r) * rs, o);
        }
      }
    }
    // ---- context rows: d cbar = (s_c * d c^ - cbar * a_c) / q_c ; d c_i = coeff_i * d cbar
    const float2 dcl = __ffma2_rn(splat2(Fs), Dl, __fmul2_rn(splat2(Fc), cl)), dch = __ffma2_rn(splat2(Fs), Dh, __fmul2_rn(splat2(Fc), ch));
#pragma unroll
    for (int r = 1; r < RMAX; ++r) {
      if (r < Cc && (FULL || col_ok)) {
        const float2 a2 = splat2(p.coeff[r - 1]);
        const float4 xv = x[r];
        const float2 ol = __fmul2_rn(__fmul2_rn(a2, dcl), make_float2(gate1(xv.x), gate1(xv.y)));
        const float2 oh = __fmul2_rn(__fmul2_rn(a2, dch), make_float2(gate1(xv.z), gate1(xv.w)));
        dbl = __fadd2_rn(dbl, ol); dbh = __fadd2_rn(dbh, oh);
        const float4 o = make_float4(ol.x, ol.y, oh.x, oh.y);
        if (OUT == 8) amax = amax4(amax, o);
        store_op4_at<OUT>(out, o_hi, o_lo, unsigned(r) * rs, o);
      }
    }
    dbacc = make_float4(dbl.x, dbl.y, dbh.x, dbh.y);
    // no trailing barrier: the next item uses the other partial-sum buffer, and the barrier after ITS reductions orders
    // this item's reads of `part` before the buffer is written again two items later
  }
  // back to the unscaled domain (oscale is a power of two: exact)
  const float inv_os = 1.f / oscale;
  dbacc.x *= inv_os; dbacc.y *= inv_os; dbacc.z *= inv_os; dbacc.w *= inv_os;
  dqacc.x *= inv_os; dqacc.y *= inv_os; dqacc.z *= inv_os; dqacc.w *= inv_os;
  if (OUT == 8) f16_publish_absmax(out.hi, amax * inv_os);
  if (ws.tickets == nullptr) {
    __shared__ float red[66];
    // no workspace: atomics (order-dependent rounding), and the ticketed loss reduction of rank_fused_kernel
    if (col_ok) {
      if (db_accum) {
        atomicAdd(db_accum + tid * 4 + 0, dbacc.x); atomicAdd(db_accum + tid * 4 + 1, dbacc.y);
        atomicAdd(db_accum + tid * 4 + 2, dbacc.z); atomicAdd(db_accum + tid * 4 + 3, dbacc.w);
      }
      if (dq_accum) {
        atomicAdd(dq_accum + tid * 4 + 0, dqacc.x); atomicAdd(dq_accum + tid * 4 + 1, dqacc.y);
        atomicAdd(dq_accum + tid * 4 + 2, dqacc.z); atomicAdd(dq_accum + tid * 4 + 3, dqacc.w);
      }
    }
    if (done_counter) {
      __syncthreads();
      last_cta_loss_reduce<false>(done_counter, item_loss, item_viol, p.B, inv_count, loss_out, viol_out, red, tid, T);
    }
    return;
  }
  column_sums_finish(ws, p, dbacc, dqacc, col_ok, db_accum, dq_accum, item_loss, item_viol, inv_count, loss_out, viol_out, tid, T);
}

// ---- K2+K3 for wide items (R > 32 rows: the large-window configuration, 16 context shots + 50 negatives) ------------
// One launch, one CTA per item, two phases over the item's rows.  Phase 1 streams the rows once from HBM (context mean,
// then |x|^2 and <cbar, x> of the target and every negative) and asks L2 to keep them (evict_last); the per-item scalar
// chain of rank_fwd_kernel / rank_bwd_kernel follows in shared memory; phase 2 reads the rows again -- 2 CTAs per SM x
// R*N*4 bytes = 81 MB in flight at R = 67, N = 1024, inside the 126 MB L2 -- and writes the gradient rows.  HBM traffic
// is R*N*4 read + the dZ operand write, as for the register-resident kernels; the two-kernel path read H twice from HBM
// and re-read the context rows a third time.  Formulas, reduction trees and summation orders are those of
// rank_fwd_kernel / rank_bwd_kernel (their results agree bit for bit except db / dq, which are summed in a fixed order
// here).  Needs nvec == 1 (N <= 1024).
// OCC: resident CTAs per SM the kernel is compiled for; rows in flight per thread follow from the register budget

__device__ __forceinline__ float4 ld4_keep(const float* p) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#if defined(__CUDA_ARCH__)
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(kEvictLast));
#endif
  return r;
}
__device__ __forceinline__ float4 ld4_last_use(const float* p) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#if defined(__CUDA_ARCH__)
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(kEvictFirst));
#endif
  return r;
}

template <int OUT, int OCC>
__global__ void __launch_bounds__(256, OCC)
rank_wide_kernel(const float* __restrict__ H, const RankDev p, const float gscale, const int act_fused,
                 const float dscale, const BwdOut out, float* __restrict__ db_accum,
                 const float* __restrict__ delta, float* __restrict__ dq_accum,
                 float* __restrict__ stats, float* __restrict__ tscore, float* __restrict__ nscore,
                 float* __restrict__ item_loss, float* __restrict__ item_viol, const RankWs ws,
                 unsigned int* __restrict__ done_counter, const float inv_count, float* __restrict__ loss_out,
                 float* __restrict__ viol_out, const int prefetch_next) {
  constexpr int kWideBatch = OCC <= 2 ? 8 : 4;
  extern __shared__ __align__(16) float sm[];
  const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = T >> 5;
  const int J = 1 + p.Nn;
  float* part = sm;                        // [2J + 1][nw]: (s_x, p_x) per branch, then s_c
  float* s_s = part + (2 * J + 1) * nw;    // [J] |x|^2
  float* s_p = s_s + J;                    // [J] <cbar, x>
  float* s_w = s_p + J;                    // [J] weight of branch j on c^ : w_0 = -sum g, w_k = g_k
  float* cA = s_w + J;                     // [J] coefficient on cbar
  float* cB = cA + J;                      // [J] coefficient on x
  float* cE = cB + J;                      // [J] w_j / n_j   (d c^ = sum_j cE_j x_j)
  float* s_l = cE + J;                     // [J] loss term of negative k
  float* s_v = s_l + J;                    // [J] violation of negative k
  float* sc = s_v + J;                     // [4] s_c, nc, Fs, Fc
  const bool col_ok = tid < p.N4;
  const size_t rs = size_t(p.B) * p.N;     // row stride of the blob, in elements
  const float hdr_scale = (out.prec == VV_PREC_F16X3) ? f16_hdr(out.hi)->scale : 1.f;
  const float oscale = hdr_scale > 0.f ? hdr_scale : 1.f;     // a header that was never set measures with scale 1
  float4 dbacc = make_float4(0.f, 0.f, 0.f, 0.f), dqacc = make_float4(0.f, 0.f, 0.f, 0.f);
  float amax = 0.f;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    const size_t e0 = size_t(b) * p.N + size_t(tid) * 4;
    const float* hp = H + e0;
    // ================= phase 1: one pass over the rows =================
    // c-bar = sum_i coeff_i * ctx_i, bottom order (eltwise_layer.cpp:67-73)
    float4 cbar = zero4;
    for (int i0 = 1; i0 < p.C; i0 += kWideBatch) {
      float4 x[kWideBatch];
#pragma unroll
      for (int r = 0; r < kWideBatch; ++r) x[r] = (i0 + r < p.C && col_ok) ? ld4_keep(hp + size_t(i0 + r) * rs) : zero4;
#pragma unroll
      for (int r = 0; r < kWideBatch; ++r) {
        if (i0 + r < p.C) {
          const float a = p.coeff[i0 + r - 1];
          cbar.x = fmaf(a, x[r].x, cbar.x); cbar.y = fmaf(a, x[r].y, cbar.y);
          cbar.z = fmaf(a, x[r].z, cbar.z); cbar.w = fmaf(a, x[r].w, cbar.w);
        }
      }
    }
    {
      const float s = warp_sum(dot4(cbar, cbar));
      if (lane == 0) part[2 * J * nw + warp] = s;
    }
    for (int j0 = 0; j0 < J; j0 += kWideBatch) {
      float v[2 * kWideBatch];
      {
        float4 x[kWideBatch];
#pragma unroll
        for (int r = 0; r < kWideBatch; ++r) {
          const int j = j0 + r;
          const int row = (j == 0) ? 0 : p.C + j - 1;
          x[r] = (j < J && col_ok) ? ld4_keep(hp + size_t(row) * rs) : zero4;
        }
#pragma unroll
        for (int r = 0; r < kWideBatch; ++r) { v[2 * r] = dot4(x[r], x[r]); v[2 * r + 1] = dot4(cbar, x[r]); }
      }
      int e; bool ok;
      warp_multi_sum<2 * kWideBatch>(v, lane, e, ok);        // same butterfly per value as warp_sum
      if (ok && 2 * j0 + e < 2 * J) part[(2 * j0 + e) * nw + warp] = v[0];
    }
    __syncthreads();
    for (int e = tid; e < 2 * J + 1; e += T) {
      float s = 0.f;
      for (int w = 0; w < nw; ++w) s += part[e * nw + w];
      if (e < 2 * J) { if (e & 1) s_p[e >> 1] = s; else s_s[e >> 1] = s; }
      else { sc[0] = s; sc[1] = sqrtf(s) + p.eps; }
      // stats layout: [s_c, s_t, p_t, s_n1, p_n1, ...]
      if (stats) stats[size_t(b) * p.stride + (e < 2 * J ? 1 + e : 0)] = s;
    }
    __syncthreads();
    // ---- scores, hinge, loss terms (normalization_layer.cpp:36-59, max_margin_loss_layer.cpp:54-214), one thread per negative
    const float nc = sc[1];
    const float score_t = s_p[0] / (nc * (sqrtf(s_s[0]) + p.eps));
    for (int k = 1 + tid; k < J; k += T) {
      const float sn = s_p[k] / (nc * (sqrtf(s_s[k]) + p.eps));
      const float dlt = score_t - sn;                        // caffe_sub :69
      const float h = fmaxf(0.f, p.margin - dlt);
      // :149-192: L2 g = h * (lw*2/count); L1 g = [h>0] * lw/count
      s_w[k] = (p.norm == 2) ? h * gscale : (h > 0.f ? gscale : 0.f);
      s_l[k] = (p.norm == 2) ? h * h : fabsf(h);
      s_v[k] = dlt < 0.f ? 1.f : 0.f;
      if (tscore) tscore[size_t(b) * p.Nn + k - 1] = score_t;  // sum_true replicates to Nn columns
      if (nscore) nscore[size_t(b) * p.Nn + k - 1] = sn;
    }
    __syncthreads();
    if (tid == 0) {
      float g = 0.f, l = 0.f, vi = 0.f;                      // sum_layer.cpp:65-68 gemv over the Nn replicated columns
      for (int k = 1; k < J; ++k) { g += s_w[k]; l += s_l[k]; vi += s_v[k]; }
      s_w[0] = -g;                                           // d s+ = -1 * d s- (axpby, :210-212)
      if (item_loss) item_loss[b] = l;
      if (item_viol) item_viol[b] = vi;
    }
    __syncthreads();
    for (int j = tid; j < J; j += T) {
      const float s = s_s[j], pj = s_p[j], w = s_w[j];
      const float nj = sqrtf(s) + p.eps;
      const float q = powf(s, 1.5f) + p.eps;                 // normalization_layer.cpp:101-107
      const float aj = w * pj / nc;                          // a = <x, w c^>
      cA[j] = s * w / (nc * q);                              // (s * w c^) / q, c^ = cbar / nc
      cB[j] = -aj / q;
      cE[j] = w / nj;
    }
    __syncthreads();
    if (tid == 0) {
      float ac = 0.f;                                        // a_c = <cbar, d c^> = sum_j cE_j p_j (split order: true, neg_1..)
      for (int j = 0; j < J; ++j) ac += cE[j] * s_p[j];
      const float s = sc[0];
      const float q = powf(s, 1.5f) + p.eps;
      sc[2] = s / q; sc[3] = -ac / q;
    }
    // ================= phase 2: the rows again (L2), gradients out =================
    // the NEXT item's rows on their way into L2 meanwhile (one bulk prefetch per row, thread r takes row r)
    if (prefetch_next && b + int(gridDim.x) < p.B && (p.N & 3) == 0) {
      for (int r = tid; r < p.C + p.Nn; r += T) {
        const float* nx = H + size_t(r) * rs + size_t(b + gridDim.x) * p.N;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(nx), "r"(unsigned(p.N) * 4u) : "memory");
      }
    }
    // target + negative rows: dx = cA*cbar + cB*x ; D += cE*x
    float4 D = zero4;
    for (int j0 = 0; j0 < J; j0 += kWideBatch) {
      float4 x[kWideBatch];
#pragma unroll
      for (int r = 0; r < kWideBatch; ++r) {
        const int j = j0 + r;
        const int row = (j == 0) ? 0 : p.C + j - 1;
        x[r] = (j < J && col_ok) ? ld4_last_use(hp + size_t(row) * rs) : zero4;
      }
#pragma unroll
      for (int r = 0; r < kWideBatch; ++r) {
        const int j = j0 + r;
        if (j < J && col_ok) {
          const int row = (j == 0) ? 0 : p.C + j - 1;
          const float a = cA[j], bb = cB[j], e = cE[j];
          const float dl = delta ? delta[size_t(row) * p.B + b] : 0.f;
          const float4 xv = x[r];
          float4 o;
          o.x = fmaf(a, cbar.x, bb * xv.x); o.y = fmaf(a, cbar.y, bb * xv.y);
          o.z = fmaf(a, cbar.z, bb * xv.z); o.w = fmaf(a, cbar.w, bb * xv.w);
          D.x = fmaf(e, xv.x, D.x); D.y = fmaf(e, xv.y, D.y); D.z = fmaf(e, xv.z, D.z); D.w = fmaf(e, xv.w, D.w);
          if (act_fused) {   // dZ = dH * mask*scale * [Z>0]  <=>  dH * scale * [H>0]
            o.x = xv.x > 0.f ? o.x * dscale : 0.f; o.y = xv.y > 0.f ? o.y * dscale : 0.f;
            o.z = xv.z > 0.f ? o.z * dscale : 0.f; o.w = xv.w > 0.f ? o.w * dscale : 0.f;
          }
          dbacc.x += o.x; dbacc.y += o.y; dbacc.z += o.z; dbacc.w += o.w;
          dqacc.x = fmaf(dl, o.x, dqacc.x); dqacc.y = fmaf(dl, o.y, dqacc.y);
          dqacc.z = fmaf(dl, o.z, dqacc.z); dqacc.w = fmaf(dl, o.w, dqacc.w);
          store_row4_t<OUT>(out, e0 + size_t(row) * rs, o, oscale, amax);
        }
      }
    }
    __syncthreads();   // sc[2], sc[3] visible
    // context rows: d cbar = (s_c * d c^ - cbar * a_c) / q_c ; d c_i = coeff_i * d cbar
    const float Fs = sc[2], Fc = sc[3];
    float4 dcb;
    dcb.x = fmaf(Fs, D.x, Fc * cbar.x); dcb.y = fmaf(Fs, D.y, Fc * cbar.y);
    dcb.z = fmaf(Fs, D.z, Fc * cbar.z); dcb.w = fmaf(Fs, D.w, Fc * cbar.w);
    for (int i0 = 1; i0 < p.C; i0 += kWideBatch) {
      float4 x[kWideBatch];
      if (act_fused) {       // the ReLU / dropout gate of these rows
#pragma unroll
        for (int r = 0; r < kWideBatch; ++r) x[r] = (i0 + r < p.C && col_ok) ? ld4_last_use(hp + size_t(i0 + r) * rs) : zero4;
      }
#pragma unroll
      for (int r = 0; r < kWideBatch; ++r) {
        const int i = i0 + r;
        if (i < p.C && col_ok) {
          const float a = p.coeff[i - 1];
          float4 o = make_float4(a * dcb.x, a * dcb.y, a * dcb.z, a * dcb.w);
          if (act_fused) {
            const float4 xv = x[r];
            o.x = xv.x > 0.f ? o.x * dscale : 0.f; o.y = xv.y > 0.f ? o.y * dscale : 0.f;
            o.z = xv.z > 0.f ? o.z * dscale : 0.f; o.w = xv.w > 0.f ? o.w * dscale : 0.f;
          }
          dbacc.x += o.x; dbacc.y += o.y; dbacc.z += o.z; dbacc.w += o.w;
          store_row4_t<OUT>(out, e0 + size_t(i) * rs, o, oscale, amax);
        }
      }
    }
    __syncthreads();   // shared-memory coefficients are rewritten by the next item
  }
  if (out.prec == VV_PREC_F16X3) f16_publish_absmax(out.hi, amax);
  if (ws.tickets == nullptr) {
    // no workspace: atomics (order-dependent rounding) into zeroed words, and the ticketed loss reduction
    __shared__ float red[66];
    if (col_ok) {
      if (db_accum) {
        atomicAdd(db_accum + tid * 4 + 0, dbacc.x); atomicAdd(db_accum + tid * 4 + 1, dbacc.y);
        atomicAdd(db_accum + tid * 4 + 2, dbacc.z); atomicAdd(db_accum + tid * 4 + 3, dbacc.w);
      }
      if (dq_accum) {
        atomicAdd(dq_accum + tid * 4 + 0, dqacc.x); atomicAdd(dq_accum + tid * 4 + 1, dqacc.y);
        atomicAdd(dq_accum + tid * 4 + 2, dqacc.z); atomicAdd(dq_accum + tid * 4 + 3, dqacc.w);
      }
    }
    if (done_counter) {
      __syncthreads();
      last_cta_loss_reduce<false>(done_counter, item_loss, item_viol, p.B, inv_count, loss_out, viol_out, red, tid, T);
    }
    return;
  }
  column_sums_finish(ws, p, dbacc, dqacc, col_ok, db_accum, dq_accum, item_loss, item_viol, inv_count, loss_out, viol_out, tid, T);
}

// ---- K2+K3 with the rows staged in shared memory by bulk async copies -----------------------------------------
// Same math, reduction trees and outputs as rank_fused_kernel (bit-identical results), different data movement: a
// producer warp streams whole items (R rows x N floats, one cp.async.bulk per row, mbarrier complete_tx) into a ring
// of `stages` shared-memory slots while the consumer warps work on earlier items, so row loads stay in flight through
// the reductions / scalar chain / stores of the current item (the register-resident kernel has no loads in flight
// during those phases and runs latency-bound at ~45 % of HBM peak).  Rows are read from shared memory twice (scores,
// then gradients) instead of living in 60+ registers.  blockDim = T consumer threads + one producer warp.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void bulk_load_row(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync(int nthreads) { asm volatile("bar.sync 1, %0;" :: "r"(nthreads) : "memory"); }
#endif

template <int RMAX, int CT, int NNT, int OUT>
__global__ void __launch_bounds__(288)
rank_ring_kernel(const float* __restrict__ H, const RankDev p, const float gscale, const int act_fused,
                 const float dscale, const BwdOut out, float* __restrict__ db_accum,
                 const float* __restrict__ delta, float* __restrict__ dq_accum,
                 float* __restrict__ stats, float* __restrict__ tscore, float* __restrict__ nscore,
                 float* __restrict__ item_loss, float* __restrict__ item_viol, const int stages,
                 unsigned int* __restrict__ done_counter, const float inv_count, float* __restrict__ loss_out,
                 float* __restrict__ viol_out) {
#if defined(__CUDA_ARCH__)      // the mbarrier / bulk-copy helpers exist in the device pass only
  extern __shared__ __align__(128) float sm[];
  const int T = blockDim.x - 32, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = T >> 5;
  const int Cc = CT > 0 ? CT : p.C, Nn = NNT > 0 ? NNT : p.Nn;
  const int J = 1 + Nn, R = Cc + Nn;
  const int per = (2 * J + 1) * nw + 4 * J + 2;        // floats per scalar buffer (double buffered by item parity)
  const size_t slot = size_t(R) * p.N;                  // floats per ring slot
  float* rows = sm;
  float* scal = rows + size_t(stages) * slot;
  uint64_t* full = reinterpret_cast<uint64_t*>(scal + 2 * per + ((2 * per) & 1));
  uint64_t* empty = full + stages;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nw); }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == nw) {                                     // ---- producer warp
    if (lane == 0) {
      const uint32_t row_bytes = uint32_t(p.N) * 4u;
      int k = 0;
      for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++k) {
        const int s = k % stages;
        if (k >= stages) mbar_wait(&empty[s], uint32_t((k / stages - 1) & 1));
        mbar_arrive_expect_tx(&full[s], row_bytes * uint32_t(R));
        float* dst = rows + size_t(s) * slot;
        for (int r = 0; r < R; ++r) bulk_load_row(dst + size_t(r) * p.N, H + (size_t(r) * p.B + b) * p.N, row_bytes, &full[s]);
      }
    }
    return;
  }
  const bool col_ok = tid < p.N4;
  float4 dbacc = make_float4(0.f, 0.f, 0.f, 0.f), dqacc = make_float4(0.f, 0.f, 0.f, 0.f);
  const float oscale = (out.prec == VV_PREC_F16X3) ? f16_hdr(out.hi)->scale : 1.f;
  float amax = 0.f;
  int parity = 0, k = 0;
  for (int b = blockIdx.x; b < p.B; b += gridDim.x, parity ^= 1, ++k) {
    float* part = scal + parity * per;                 // [2J+1][nw]: (s_x, p_x) per branch, then s_c
    const int s = k % stages;
    const float4* xs = reinterpret_cast<const float4*>(rows + size_t(s) * slot) + tid;     // row r: xs[r * N4]
    mbar_wait(&full[s], uint32_t((k / stages) & 1));
#define VV_X(r) (col_ok ? xs[size_t(r) * p.N4] : make_float4(0.f, 0.f, 0.f, 0.f))
    // ---- context mean, bottom order (eltwise_layer.cpp:67-73)
    float4 cbar = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 1; r < RMAX; ++r) {
      if (r < Cc) {
        const float a = p.coeff[r - 1];
        const float4 xv = VV_X(r);
        cbar.x = fmaf(a, xv.x, cbar.x); cbar.y = fmaf(a, xv.y, cbar.y);
        cbar.z = fmaf(a, xv.z, cbar.z); cbar.w = fmaf(a, xv.w, cbar.w);
      }
    }
    if (NNT > 0) {
      constexpr int NV = 2 * (NNT > 0 ? 1 + NNT : 1) + 1;
      float v[NV];
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        if (r < R && (r == 0 || r >= Cc)) {
          const int j = (r == 0) ? 0 : r - Cc + 1;
          const float4 xv = VV_X(r);
          v[2 * j] = dot4(xv, xv); v[2 * j + 1] = dot4(cbar, xv);
        }
      }
      v[NV - 1] = dot4(cbar, cbar);
      int e; bool ok;
      warp_multi_sum<NV>(v, lane, e, ok);
      if (ok) part[e * nw + warp] = v[0];
    } else {
      {
        float ssum = warp_sum(dot4(cbar, cbar));
        if (lane == 0) part[2 * J * nw + warp] = ssum;
      }
#pragma unroll
      for (int r = 0; r < RMAX; ++r) {
        if (r < R && (r == 0 || r >= Cc)) {             // uniform branch
          const int j = (r == 0) ? 0 : r - Cc + 1;
          const float4 xv = VV_X(r);
          float sx = warp_sum(dot4(xv, xv)), px = warp_sum(dot4(cbar, xv));
          if (lane == 0) { part[(j * 2 + 0) * nw + warp] = sx; part[(j * 2 + 1) * nw + warp] = px; }
        }
      }
    }
    consumer_sync(T);
    // ---- per-item scalars, lane j = branch j (formulas of rank_fused_kernel).  Every warp evaluates the chain
    // redundantly from the shared partial sums and keeps the coefficients in registers (broadcast by shuffle below):
    // no second barrier, and no warp idles while one of them works through the sqrt / pow / divide chain.
    float rA = 0.f, rB = 0.f, rE = 0.f, rD = 0.f, Fs, Fc;
    {
      float s_c = 0.f;
      for (int w = 0; w < nw; ++w) s_c += part[2 * J * nw + w];
      float sj = 0.f, pj = 0.f;
      if (lane < J) {
        for (int w = 0; w < nw; ++w) sj += part[(lane * 2 + 0) * nw + w];
        for (int w = 0; w < nw; ++w) pj += part[(lane * 2 + 1) * nw + w];
      }
      if (stats && warp == 0) {
        float* st = stats + size_t(b) * p.stride;
        if (lane == 0) st[0] = s_c;
        if (lane < J) { st[1 + 2 * lane] = sj; st[2 + 2 * lane] = pj; }
      }
      const float nc = sqrtf(s_c) + p.eps;
      const float nj = sqrtf(sj) + p.eps;
      const float score = pj / (nc * nj);
      const float score_t = __shfl_sync(0xffffffffu, score, 0);
      const float dlt = score_t - score;
      const float h = fmaxf(0.f, p.margin - dlt);
      const bool neg = lane >= 1 && lane < J;
      float w = neg ? ((p.norm == 2) ? h * gscale : (h > 0.f ? gscale : 0.f)) : 0.f;
      const float vterm = (neg && dlt < 0.f) ? 1.f : 0.f;
      const float loss = warp_sum(neg ? ((p.norm == 2) ? h * h : fabsf(h)) : 0.f);
      const float viol = warp_sum(vterm);
      const float g = warp_sum(w);
      if (lane == 0) w = -g;
      if (neg && warp == 0) {
        if (tscore) tscore[size_t(b) * Nn + lane - 1] = score_t;
        if (nscore) nscore[size_t(b) * Nn + lane - 1] = score;
      }
      if (lane == 0 && warp == 0) { if (item_loss) item_loss[b] = loss; if (item_viol) item_viol[b] = viol; }
      const float q = powf(sj, 1.5f) + p.eps;
      const float aj = w * pj / nc;
      const float e = w / nj;
      if (lane < J) {
        rA = sj * w / (nc * q); rB = -aj / q; rE = e;
        rD = delta ? delta[size_t(lane == 0 ? 0 : Cc + lane - 1) * p.B + b] : 0.f;
      }
      const float ac = warp_sum(lane < J ? e * pj : 0.f);
      const float qc = powf(s_c, 1.5f) + p.eps;
      Fs = s_c / qc; Fc = -ac / qc;
    }
    // ---- target + negative rows: dx = cA*cbar + cB*x ; D += cE*x
    float4 D = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      if (r < R && (r == 0 || r >= Cc)) {
        const int j = (r == 0) ? 0 : r - Cc + 1;
        const float a = __shfl_sync(0xffffffffu, rA, j), bb = __shfl_sync(0xffffffffu, rB, j), e = __shfl_sync(0xffffffffu, rE, j);
        const float dl = __shfl_sync(0xffffffffu, rD, j);
        if (col_ok) {
          const float4 xv = xs[size_t(r) * p.N4];
          float4 o;
          o.x = fmaf(a, cbar.x, bb * xv.x); o.y = fmaf(a, cbar.y, bb * xv.y);
          o.z = fmaf(a, cbar.z, bb * xv.z); o.w = fmaf(a, cbar.w, bb * xv.w);
          D.x = fmaf(e, xv.x, D.x); D.y = fmaf(e, xv.y, D.y); D.z = fmaf(e, xv.z, D.z); D.w = fmaf(e, xv.w, D.w);
          if (act_fused) {
            o.x = xv.x > 0.f ? o.x * dscale : 0.f; o.y = xv.y > 0.f ? o.y * dscale : 0.f;
            o.z = xv.z > 0.f ? o.z * dscale : 0.f; o.w = xv.w > 0.f ? o.w * dscale : 0.f;
          }
          dbacc.x += o.x; dbacc.y += o.y; dbacc.z += o.z; dbacc.w += o.w;
          if (dl != 0.f) {
            dqacc.x = fmaf(dl, o.x, dqacc.x); dqacc.y = fmaf(dl, o.y, dqacc.y);
            dqacc.z = fmaf(dl, o.z, dqacc.z); dqacc.w = fmaf(dl, o.w, dqacc.w);
          }
          store_row4_t<OUT>(out, (size_t(r) * p.B + b) * p.N + tid * 4, o, oscale, amax);
        }
      }
    }
    // ---- context rows: d cbar = (s_c * d c^ - cbar * a_c) / q_c ; d c_i = coeff_i * d cbar
    float4 dcb;
    dcb.x = fmaf(Fs, D.x, Fc * cbar.x); dcb.y = fmaf(Fs, D.y, Fc * cbar.y);
    dcb.z = fmaf(Fs, D.z, Fc * cbar.z); dcb.w = fmaf(Fs, D.w, Fc * cbar.w);
#pragma unroll
    for (int r = 1; r < RMAX; ++r) {
      if (r < Cc && col_ok) {
        const float a = p.coeff[r - 1];
        float4 o = make_float4(a * dcb.x, a * dcb.y, a * dcb.z, a * dcb.w);
        if (act_fused) {
          const float4 xv = xs[size_t(r) * p.N4];
          o.x = xv.x > 0.f ? o.x * dscale : 0.f; o.y = xv.y > 0.f ? o.y * dscale : 0.f;
          o.z = xv.z > 0.f ? o.z * dscale : 0.f; o.w = xv.w > 0.f ? o.w * dscale : 0.f;
        }
        dbacc.x += o.x; dbacc.y += o.y; dbacc.z += o.z; dbacc.w += o.w;
        store_row4_t<OUT>(out, (size_t(r) * p.B + b) * p.N + tid * 4, o, oscale, amax);
      }
    }
#undef VV_X
    // this warp is done with the slot: hand it back to the producer
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  if (out.prec == VV_PREC_F16X3) f16_publish_absmax(out.hi, amax);
  if (col_ok) {
    if (db_accum) {
      atomicAdd(db_accum + tid * 4 + 0, dbacc.x); atomicAdd(db_accum + tid * 4 + 1, dbacc.y);
      atomicAdd(db_accum + tid * 4 + 2, dbacc.z); atomicAdd(db_accum + tid * 4 + 3, dbacc.w);
    }
    if (dq_accum) {
      atomicAdd(dq_accum + tid * 4 + 0, dqacc.x); atomicAdd(dq_accum + tid * 4 + 1, dqacc.y);
      atomicAdd(dq_accum + tid * 4 + 2, dqacc.z); atomicAdd(dq_accum + tid * 4 + 3, dqacc.w);
    }
  }
  if (done_counter) {
    __shared__ float red[66];
    consumer_sync(T);                                // every item of this CTA has been written (by warp 0)
    last_cta_loss_reduce<true>(done_counter, item_loss, item_viol, p.B, inv_count, loss_out, viol_out, red, tid, T);
  }
#endif
}

int make_dev(const vv_rank_cfg_t* cfg, RankDev* d, int* threads) {
  VV_REQUIRE(cfg, "rank cfg is NULL");
  VV_REQUIRE(cfg->B > 0 && cfg->C >= 2 && cfg->Nn >= 1 && cfg->N > 0, "bad rank cfg B=%d C=%d Nn=%d N=%d", cfg->B, cfg->C, cfg->Nn, cfg->N);
  VV_REQUIRE(cfg->C - 1 <= VV_MAX_CONTEXT, "context_size-1 = %d exceeds VV_MAX_CONTEXT", cfg->C - 1);
  VV_REQUIRE(cfg->N % 4 == 0, "embedding dim N=%d must be a multiple of 4", cfg->N);
  VV_REQUIRE(cfg->N <= 4 * 256 * kMaxVec, "embedding dim N=%d exceeds %d", cfg->N, 4 * 256 * kMaxVec);
  VV_REQUIRE(cfg->norm == 1 || cfg->norm == 2, "norm must be 1 (L1) or 2 (L2)");
  d->B = cfg->B; d->C = cfg->C; d->Nn = cfg->Nn; d->N = cfg->N; d->N4 = cfg->N / 4;
  int T = ((d->N4 + 31) / 32) * 32;
  if (T > 256) T = 256;
  d->nvec = (d->N4 + T - 1) / T;
  d->stride = vv_rank_stats_stride(cfg->Nn);
  d->margin = cfg->margin; d->eps = cfg->eps; d->norm = cfg->norm;
  for (int i = 0; i < VV_MAX_CONTEXT; ++i) d->coeff[i] = cfg->coeff[i];
  *threads = T;
  return VV_OK;
}

}  // namespace
}  // namespace vv

using namespace vv;

extern "C" int vv_rank_loss_forward(const float* H, const vv_rank_cfg_t* cfg, float* stats,
                                    float* target_score, float* neg_score,
                                    float* item_loss, float* item_viol,
                                    float* loss, float* violations, vv_stream_t stream) {
  RankDev d; int T;
  int rc = make_dev(cfg, &d, &T);
  if (rc) return rc;
  VV_REQUIRE(H && stats, "H and stats must be non-NULL");
  VV_REQUIRE(!(loss || violations) || (item_loss && item_viol), "loss/violations need item_loss and item_viol scratch [B]");
  const int J = 1 + d.Nn, nw = T / 32;
  const size_t smem = sizeof(float) * (J * 2 * nw + nw + 2 * J + 1);
  const int grid = d.B < num_sms() * 8 ? d.B : num_sms() * 8;
  switch (d.nvec) {
    case 1: rank_fwd_kernel<1><<<grid, T, smem, stream>>>(H, d, stats, target_score, neg_score, item_loss, item_viol); break;
    case 2: rank_fwd_kernel<2><<<grid, T, smem, stream>>>(H, d, stats, target_score, neg_score, item_loss, item_viol); break;
    default: rank_fwd_kernel<4><<<grid, T, smem, stream>>>(H, d, stats, target_score, neg_score, item_loss, item_viol); break;
  }
  VV_LAUNCH_CHECK();
  count_launch();
  if (loss || violations) {
    rank_loss_reduce_kernel<<<1, 1024, 0, stream>>>(item_loss, item_viol, d.B, 1.f / float(d.B * d.Nn), loss, violations);
    VV_LAUNCH_CHECK();
    count_launch();
  }
  return VV_OK;
}

extern "C" int vv_rank_loss_backward(const float* H, const vv_rank_cfg_t* cfg, const float* stats,
                                     float loss_weight, int act_fused, float dropout_scale,
                                     float* dZ, void* dZop_hi, void* dZop_lo, int prec,
                                     float* db_accum, vv_stream_t stream) {
  return vv_rank_loss_backward_ex(H, cfg, stats, loss_weight, act_fused, dropout_scale, dZ, dZop_hi, dZop_lo, prec,
                                  db_accum, nullptr, nullptr, stream);
}

extern "C" int vv_rank_loss_backward_ex(const float* H, const vv_rank_cfg_t* cfg, const float* stats,
                                        float loss_weight, int act_fused, float dropout_scale,
                                        float* dZ, void* dZop_hi, void* dZop_lo, int prec,
                                        float* db_accum, const float* delta, float* dq_accum, vv_stream_t stream) {
  RankDev d; int T;
  int rc = make_dev(cfg, &d, &T);
  if (rc) return rc;
  VV_REQUIRE(H && stats, "H and stats must be non-NULL");
  BwdOut o; o.dZ = dZ; o.hi = nullptr; o.lo = nullptr; o.bf = nullptr; o.prec = VV_PREC_FP32_SIMT;
  o.count = size_t(d.B) * (d.C + d.Nn) * d.N;
  if ((prec == VV_PREC_TF32X3 || prec == VV_PREC_F16X3) && dZop_hi) {
    VV_REQUIRE(dZop_lo, "split operand copy needs hi and lo");
    o.hi = static_cast<float*>(dZop_hi); o.lo = static_cast<float*>(dZop_lo); o.prec = prec;
  } else if (prec == VV_PREC_BF16 && dZop_hi) {
    o.bf = static_cast<uint16_t*>(dZop_hi); o.prec = prec;
  }
  VV_REQUIRE(o.dZ || o.prec != VV_PREC_FP32_SIMT, "no output requested");
  VV_REQUIRE(!dq_accum || delta, "dq_accum needs delta");
  const int count = d.B * d.Nn;
  const float gscale = (d.norm == 2) ? loss_weight * 2 / count : loss_weight / count;
  const int J = 1 + d.Nn;
  const size_t smem = sizeof(float) * (6 * J + 4);
  const int grid = d.B < num_sms() * 8 ? d.B : num_sms() * 8;
  switch (d.nvec) {
    case 1: rank_bwd_kernel<1><<<grid, T, smem, stream>>>(H, d, stats, gscale, act_fused, dropout_scale, o, db_accum, delta, dq_accum); break;
    case 2: rank_bwd_kernel<2><<<grid, T, smem, stream>>>(H, d, stats, gscale, act_fused, dropout_scale, o, db_accum, delta, dq_accum); break;
    default: rank_bwd_kernel<4><<<grid, T, smem, stream>>>(H, d, stats, gscale, act_fused, dropout_scale, o, db_accum, delta, dq_accum); break;
  }
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}

// register-resident kernels: R <= 32 rows per item; wide kernel (two phases, second one out of L2): any R with Nn <= 255
static bool rank_narrow(const vv_rank_cfg_t* cfg) { return cfg->C + cfg->Nn <= 32 && 1 + cfg->Nn <= 32; }
extern "C" int vv_rank_loss_fused_supported(const vv_rank_cfg_t* cfg) {
  static const bool wide_on = [] { const char* e = getenv("VV_RANK_WIDE"); return !(e && atoi(e) == 0); }();
  return cfg && cfg->N % 4 == 0 && cfg->N <= 1024 && cfg->Nn >= 1 && cfg->C >= 2 && cfg->C - 1 <= VV_MAX_CONTEXT &&
         (rank_narrow(cfg) || (wide_on && cfg->Nn <= 255));
}

extern "C" int vv_rank_loss_fused(const float* H, const vv_rank_cfg_t* cfg, float loss_weight, int act_fused,
                                  float dropout_scale, float* stats, float* target_score, float* neg_score,
                                  float* item_loss, float* item_viol, float* loss, float* violations,
                                  float* dZ, void* dZop_hi, void* dZop_lo, int prec, float* db_accum,
                                  const float* delta, float* dq_accum, vv_stream_t stream) {
  return vv::rank_loss_fused_counted(H, cfg, loss_weight, act_fused, dropout_scale, stats, target_score, neg_score, item_loss,
                                     item_viol, loss, violations, dZ, dZop_hi, dZop_lo, prec, db_accum, delta, dq_accum,
                                     nullptr, stream, nullptr, 0);
}

// done_counter != NULL (a zeroed device word): the batch loss / violation sums are produced by the last CTA of the
// fused kernel instead of a separate rank_loss_reduce_kernel launch (the trainer's path).
// workspace of the second-generation kernel's deterministic column sums: tickets | per-CTA sums | per-group sums
size_t vv::rank_loss_workspace_bytes(int N) {
  const size_t grid = size_t(num_sms()) * 6, groups = (grid + kRankGroup - 1) / kRankGroup;     // up to 6 resident CTAs per SM
  return 1024 + (grid + groups) * 2 * size_t(N) * sizeof(float);
}
bool vv::rank_loss_fused_v2_applies(const vv_rank_cfg_t* cfg, int prec, bool want_dz, bool want_scores) {
  static const bool on = [] { const char* e = getenv("VV_RANK_V2"); return !(e && atoi(e) == 0); }();
  const char* ring_e = getenv("VV_RANK_RING");
  if (!on || (ring_e && atoi(ring_e) > 0)) return false;
  if (!cfg || want_dz || want_scores) return false;
  if (prec != VV_PREC_TF32X3 && prec != VV_PREC_BF16 && prec != VV_PREC_F16X3) return false;
  if (double(cfg->B) * (cfg->C + cfg->Nn) * cfg->N >= 2147483648.0) return false;     // the kernel uses 32-bit element offsets
  return cfg->N % 4 == 0 && cfg->N <= 1024 && cfg->C + cfg->Nn <= 16 && cfg->Nn >= 1 && cfg->C >= 2;
}

int vv::rank_loss_fused_counted(const float* H, const vv_rank_cfg_t* cfg, float loss_weight, int act_fused,
                                float dropout_scale, float* stats, float* target_score, float* neg_score,
                                float* item_loss, float* item_viol, float* loss, float* violations,
                                float* dZ, void* dZop_hi, void* dZop_lo, int prec, float* db_accum,
                                const float* delta, float* dq_accum, unsigned int* done_counter, vv_stream_t stream,
                                void* workspace, size_t workspace_bytes) {
  RankDev d; int T;
  int rc = make_dev(cfg, &d, &T);
  if (rc) return rc;
  VV_REQUIRE(vv_rank_loss_fused_supported(cfg), "rank_loss_fused: needs N <= 1024 and Nn <= 255 (use forward + backward)");
  VV_REQUIRE(H, "H must be non-NULL");
  VV_REQUIRE(!(loss || violations) || (item_loss && item_viol), "loss/violations need item_loss and item_viol scratch [B]");
  BwdOut o; o.dZ = dZ; o.hi = nullptr; o.lo = nullptr; o.bf = nullptr; o.prec = VV_PREC_FP32_SIMT;
  o.count = size_t(d.B) * (d.C + d.Nn) * d.N;
  if ((prec == VV_PREC_TF32X3 || prec == VV_PREC_F16X3) && dZop_hi) {
    VV_REQUIRE(dZop_lo, "split operand copy needs hi and lo");
    o.hi = static_cast<float*>(dZop_hi); o.lo = static_cast<float*>(dZop_lo); o.prec = prec;
  } else if (prec == VV_PREC_BF16 && dZop_hi) {
    o.bf = static_cast<uint16_t*>(dZop_hi); o.prec = prec;
  }
  VV_REQUIRE(o.dZ || o.prec != VV_PREC_FP32_SIMT, "no output requested");
  VV_REQUIRE(!dq_accum || delta, "dq_accum needs delta");
  const int count = d.B * d.Nn;
  const float gscale = (d.norm == 2) ? loss_weight * 2 / count : loss_weight / count;
  const int J = 1 + d.Nn, nw = T / 32, R = d.C + d.Nn;
  const size_t smem = sizeof(float) * 2 * ((2 * J + 1) * nw + 4 * J + 2);
  const int per_sm = (R <= 16) ? 4 : 2;         // resident CTAs per SM at the kernels' register counts x T threads
  const int grid = d.B < num_sms() * per_sm ? d.B : num_sms() * per_sm;
  const int mode = (o.dZ ? 1 : 0) | (o.prec == VV_PREC_TF32X3 ? 2 : 0) | (o.prec == VV_PREC_BF16 ? 4 : 0) |
                   (o.prec == VV_PREC_F16X3 ? 8 : 0);
  unsigned int* cnt = (loss || violations) ? done_counter : nullptr;       // fold the batch reduction into the kernel
  const float inv_count = 1.f / float(d.B * d.Nn);
  if (!rank_narrow(cfg)) {
    // wide items: one launch, two phases per item (rank_wide_kernel)
    static const int wocc = [] { const char* e = getenv("VV_RANK_WIDE_CTAS"); const int v = e ? atoi(e) : 2; return v < 2 ? 2 : (v > 4 ? 4 : v); }();
    static const int wpf = [] { const char* e = getenv("VV_RANK_WIDE_PREFETCH"); return e ? atoi(e) : 0; }();
    const int wgrid = d.B < num_sms() * wocc ? d.B : num_sms() * wocc;
    const size_t wsmem = sizeof(float) * ((2 * J + 1) * nw + 8 * J + 4);
    RankWs ws; ws.tickets = nullptr; ws.cta_part = nullptr; ws.grp_part = nullptr;
    if (workspace && (loss || violations ? (item_loss && item_viol) : true)) {
      const size_t groups = (size_t(wgrid) + kRankGroup - 1) / kRankGroup;
      const size_t need = 1024 + (size_t(wgrid) + groups) * 2 * size_t(d.N) * sizeof(float);
      VV_REQUIRE(workspace_bytes >= need && groups + 1 <= 256, "rank_loss_fused: workspace too small (%zu bytes, need %zu)", workspace_bytes, need);
      ws.tickets = static_cast<unsigned int*>(workspace);
      ws.cta_part = reinterpret_cast<float*>(static_cast<char*>(workspace) + 1024);
      ws.grp_part = ws.cta_part + size_t(wgrid) * 2 * d.N;
    }
#define VV_RANK_WIDE(OUT)                                                                                              \
    do {                                                                                                               \
      if (wocc == 2) rank_wide_kernel<OUT, 2><<<wgrid, T, wsmem, stream>>>(H, d, gscale, act_fused, dropout_scale, o, db_accum, delta, dq_accum, \
          stats, target_score, neg_score, item_loss, item_viol, ws, cnt, inv_count, loss, violations, wpf);            \
      else if (wocc == 3) rank_wide_kernel<OUT, 3><<<wgrid, T, wsmem, stream>>>(H, d, gscale, act_fused, dropout_scale, o, db_accum, delta, dq_accum, \
          stats, target_score, neg_score, item_loss, item_viol, ws, cnt, inv_count, loss, violations, wpf);            \
      else rank_wide_kernel<OUT, 4><<<wgrid, T, wsmem, stream>>>(H, d, gscale, act_fused, dropout_scale, o, db_accum, delta, dq_accum, \
          stats, target_score, neg_score, item_loss, item_viol, ws, cnt, inv_count, loss, violations, wpf);            \
    } while (0)
    if (mode == 8) VV_RANK_WIDE(8); else if (mode == 4) VV_RANK_WIDE(4); else if (mode == 2) VV_RANK_WIDE(2); else VV_RANK_WIDE(-1);
#undef VV_RANK_WIDE
    VV_LAUNCH_CHECK();
    count_launch();
    if ((loss || violations) && !cnt && !ws.tickets) {
      rank_loss_reduce_kernel<<<1, 1024, 0, stream>>>(item_loss, item_viol, d.B, inv_count, loss, violations);
      VV_LAUNCH_CHECK();
      count_launch();
    }
    return VV_OK;
  }
  if (rank_loss_fused_v2_applies(cfg, o.prec, o.dZ != nullptr, target_score || neg_score) && dZop_hi) {
    static const int occ_env = [] { const char* e = getenv("VV_RANK2_CTAS"); const int v = e ? atoi(e) : 4; return v < 4 ? 4 : (v > 6 ? 6 : v); }();
    const bool hot = (T == d.N4 && T == 128 && d.C == 5 && d.Nn == 10);       // the occupancy variants exist for the shipped shape
    const int per2 = (T <= 128) ? (hot ? occ_env : 4) : 2;
    const int grid2 = d.B < num_sms() * per2 ? d.B : num_sms() * per2;
    const size_t smem2 = sizeof(float) * 2 * (2 * J + 1) * nw;
    RankWs ws; ws.tickets = nullptr; ws.cta_part = nullptr; ws.grp_part = nullptr;
    const bool want_sums = db_accum || dq_accum || loss || violations;
    if (workspace && want_sums && (loss || violations ? (item_loss && item_viol) : true)) {
      const size_t groups = (size_t(grid2) + kRankGroup - 1) / kRankGroup;
      const size_t need = 1024 + (size_t(grid2) + groups) * 2 * size_t(d.N) * sizeof(float);
      VV_REQUIRE(workspace_bytes >= need && groups + 1 <= 256, "rank_loss_fused: workspace too small (%zu bytes, need %zu)", workspace_bytes, need);
      ws.tickets = static_cast<unsigned int*>(workspace);
      ws.cta_part = reinterpret_cast<float*>(static_cast<char*>(workspace) + 1024);
      ws.grp_part = ws.cta_part + size_t(grid2) * 2 * d.N;
    }
    const float ds = act_fused ? dropout_scale : 1.f;
    VV_REQUIRE(act_fused, "rank_loss_fused: the operand-only fast kernel is the fused-activation form");
#define VV_RANK_V2(CT, NNT, OUT)                                                                                        \
    do {                                                                                                                 \
      if (T == d.N4 && T == 128 && CT == 5 && per2 == 5) rank_fused2_kernel<CT, NNT, OUT, true, 4, (CT == 5 ? 5 : 4)><<<grid2, T, smem2, stream>>>(H, d, gscale, ds, o, db_accum,  \
          delta, dq_accum, stats, item_loss, item_viol, ws, cnt, inv_count, loss, violations);                             \
      else if (T == d.N4 && T == 128 && CT == 5 && per2 == 6) rank_fused2_kernel<CT, NNT, OUT, true, 4, (CT == 5 ? 6 : 4)><<<grid2, T, smem2, stream>>>(H, d, gscale, ds, o, db_accum,  \
          delta, dq_accum, stats, item_loss, item_viol, ws, cnt, inv_count, loss, violations);                             \
      else if (T == d.N4 && T == 128) rank_fused2_kernel<CT, NNT, OUT, true, 4><<<grid2, T, smem2, stream>>>(H, d, gscale, ds, o, db_accum,  \
          delta, dq_accum, stats, item_loss, item_viol, ws, cnt, inv_count, loss, violations);                             \
      else if (T == d.N4) rank_fused2_kernel<CT, NNT, OUT, true, 0><<<grid2, T, smem2, stream>>>(H, d, gscale, ds, o, db_accum, delta,  \
          dq_accum, stats, item_loss, item_viol, ws, cnt, inv_count, loss, violations);                                    \
      else rank_fused2_kernel<CT, NNT, OUT, false, 0><<<grid2, T, smem2, stream>>>(H, d, gscale, ds, o, db_accum, delta, dq_accum,      \
          stats, item_loss, item_viol, ws, cnt, inv_count, loss, violations);                                              \
    } while (0)
    if (d.C == 5 && d.Nn == 10) {
      if (mode == 8) VV_RANK_V2(5, 10, 8); else if (mode == 4) VV_RANK_V2(5, 10, 4); else VV_RANK_V2(5, 10, 2);
    } else {
      if (mode == 8) VV_RANK_V2(0, 0, 8); else if (mode == 4) VV_RANK_V2(0, 0, 4); else VV_RANK_V2(0, 0, 2);
    }
#undef VV_RANK_V2
    VV_LAUNCH_CHECK();
    count_launch();
    if ((loss || violations) && !cnt && !ws.tickets) {
      rank_loss_reduce_kernel<<<1, 1024, 0, stream>>>(item_loss, item_viol, d.B, inv_count, loss, violations);
      VV_LAUNCH_CHECK();
      count_launch();
    }
    return VV_OK;
  }
  // ring variant (rows staged in shared memory by a producer warp): whenever two slots of R rows fit
  const char* ring_e = getenv("VV_RANK_RING");                      // 0 = register-resident kernel, n >= 2 = ring stages
  const int ring_env = ring_e ? atoi(ring_e) : -1;
  const size_t slot_bytes = size_t(R) * d.N * sizeof(float);
  const size_t scal_bytes = sizeof(float) * (2 * ((2 * J + 1) * nw + 4 * J + 2) + 1);
  int stages = ring_env >= 0 ? ring_env : 0;     // experimental: off unless VV_RANK_RING names the stage count
  while (stages >= 1 && stages * slot_bytes + scal_bytes + 16 * stages > 200 * 1024) --stages;
  const bool ring = stages >= 1 && d.N % 4 == 0 && (d.N * 4) % 16 == 0 && T == d.N4 && R <= 32;
  if (ring) {
    const size_t rsmem = stages * slot_bytes + scal_bytes + 16 * stages;
    int rper_sm = int((224 * 1024) / (rsmem + 1024));
    if (rper_sm > 2048 / (T + 32)) rper_sm = 2048 / (T + 32);
    if (const char* e = getenv("VV_RANK_RING_PER_SM")) rper_sm = atoi(e) < rper_sm ? atoi(e) : rper_sm;
    if (rper_sm < 1) rper_sm = 1;
    const int rgrid = d.B < num_sms() * rper_sm ? d.B : num_sms() * rper_sm;
#define VV_RANK_RING(RM, CT, NNT, OUT)                                                                                 \
  do {                                                                                                                 \
    static bool attr_set = false;                                                                                      \
    if (!attr_set) {                                                                                                   \
      VV_CUDA(cudaFuncSetAttribute(rank_ring_kernel<RM, CT, NNT, OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 204 * 1024)); \
      attr_set = true;                                                                                                 \
    }                                                                                                                  \
    rank_ring_kernel<RM, CT, NNT, OUT><<<rgrid, T + 32, rsmem, stream>>>(H, d, gscale, act_fused, dropout_scale, o, db_accum, \
        delta, dq_accum, stats, target_score, neg_score, item_loss, item_viol, stages, cnt, inv_count, loss, violations); \
  } while (0)
    if (d.C == 5 && d.Nn == 10 && mode == 2) VV_RANK_RING(16, 5, 10, 2);
    else if (d.C == 5 && d.Nn == 10 && mode == 4) VV_RANK_RING(16, 5, 10, 4);
    else if (d.C == 5 && d.Nn == 10 && mode == 8) VV_RANK_RING(16, 5, 10, 8);
    else if (d.C == 5 && d.Nn == 10 && mode == 1) VV_RANK_RING(16, 5, 10, 1);
    else if (R <= 16) VV_RANK_RING(16, 0, 0, -1);
    else VV_RANK_RING(32, 0, 0, -1);
#undef VV_RANK_RING
  } else {
#define VV_RANK_FUSED(RM, CT, NNT, OUT)                                                                              \
    rank_fused_kernel<RM, CT, NNT, OUT><<<grid, T, smem, stream>>>(H, d, gscale, act_fused, dropout_scale, o, db_accum, \
        delta, dq_accum, stats, target_score, neg_score, item_loss, item_viol, cnt, inv_count, loss, violations)
    if (d.C == 5 && d.Nn == 10 && mode == 2) VV_RANK_FUSED(16, 5, 10, 2);        // the shipped net (C=5, Nn=10), training modes
    else if (d.C == 5 && d.Nn == 10 && mode == 4) VV_RANK_FUSED(16, 5, 10, 4);
    else if (d.C == 5 && d.Nn == 10 && mode == 8) VV_RANK_FUSED(16, 5, 10, 8);
    else if (d.C == 5 && d.Nn == 10 && mode == 1) VV_RANK_FUSED(16, 5, 10, 1);
    else if (R <= 16) VV_RANK_FUSED(16, 0, 0, -1);
    else VV_RANK_FUSED(32, 0, 0, -1);
  #undef VV_RANK_FUSED
  }
  VV_LAUNCH_CHECK();
  count_launch();
  if ((loss || violations) && !cnt) {
    rank_loss_reduce_kernel<<<1, 1024, 0, stream>>>(item_loss, item_viol, d.B, inv_count, loss, violations);
    VV_LAUNCH_CHECK();
    count_launch();
  }
  return VV_OK;
}
