// vv_common.cuh -- shared device helpers for the sm_100a kernels:
// PTX wrappers (mbarrier, TMA, tcgen05/TMEM), counter RNGs, small math.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "vv_b200.h"

namespace vv {

// ----------------------------------------------------------------------------
// error plumbing (host)
// ----------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
#define VV_CUDA(call)                                                        \
  do {                                                                       \
    cudaError_t _e = (call);                                                 \
    if (_e != cudaSuccess) return ::vv::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)
#define VV_LAUNCH_CHECK() VV_CUDA(cudaPeekAtLastError())
#define VV_REQUIRE(cond, ...)                                                \
  do {                                                                       \
    if (!(cond)) { ::vv::set_error(__VA_ARGS__); return VV_ERR_INVALID; }    \
  } while (0)

int num_sms();
void count_launch(int n = 1);   // per-thread launch counter (vv_trainer_last_launches)
int launches_reset();           // returns the count and zeroes it

// K4 for the trainer's two parameter blobs in ONE launch (vv_stream_kernels.cu): the weight update of vv_sgd_update
// with (a) col_add[n] added to slab 0's gradient of column K-1 first (the K-1 copy quirk's share; what vv_add_column
// did in its own launch), (b) the updated column K-1 saved to col_out (was vv_copy_strided), and (c) the bias blob
// updated by extra CTAs of the same grid (was a second vv_sgd_update).  Element-wise arithmetic identical to the
// separate launches.
struct UpdateTail {
  float* W; const float* parts; int nparts; long long stride; float* hist; float* diff_out; long long count; int K;
  float rate_w, decay_w;
  const float* col_add; float* col_out;
  void* Wop_hi; void* Wop_lo; int prec;
  float* b; const float* db; float* bh; float* b_diff; int nb; float rate_b, decay_b;
  float momentum; int reg_type; float gscale;
};
int sgd_update_tail(const UpdateTail& u, vv_stream_t stream);

// vv_operand_rescale that also folds in n_extra recorded maxima (bit patterns) kept outside the operand's header
int operand_rescale_ex(void* hi, int prec, int target_log2, const unsigned int* extra_bits, int n_extra, vv_stream_t stream);
// vv_rank_loss_fused with the batch loss / violation reduction folded into the kernel (vv_rank_loss.cu)
int rank_loss_fused_counted(const float* H, const vv_rank_cfg_t* cfg, float loss_weight, int act_fused,
                            float dropout_scale, float* stats, float* target_score, float* neg_score,
                            float* item_loss, float* item_viol, float* loss, float* violations,
                            float* dZ, void* dZop_hi, void* dZop_lo, int prec, float* db_accum,
                            const float* delta, float* dq_accum, unsigned int* done_counter, vv_stream_t stream,
                            void* workspace = nullptr, size_t workspace_bytes = 0);
// Second-generation fused kernel (operand-only output, R <= 16, N <= 1024): with a workspace of rank_loss_workspace_bytes(N)
// the bias / quirk column sums and the batch loss are reduced in a fixed order (no float atomics; db_accum / dq_accum are
// then STORED, they need no zeroing) -- see rank_fused2_kernel in vv_rank_loss.cu.
size_t rank_loss_workspace_bytes(int N);
bool rank_loss_fused_v2_applies(const vv_rank_cfg_t* cfg, int prec, bool want_dz, bool want_scores);

constexpr int kNumSMsB200 = 148;

// ----------------------------------------------------------------------------
// small device math
// ----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// round-to-nearest fp32 -> tf32 (result is an fp32 bit pattern with 13 low zero bits)
__device__ __forceinline__ float to_tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = to_tf32_rna(x);
  lo = to_tf32_rna(x - hi);   // x - hi is exact in fp32
}

// streaming 128-bit global accesses
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// The TF32X3 operand format of 4 consecutive elements: hi = tf32(x) (fp32 array), and in the `lo` array two
// bf16 planes of `count` elements each: plane 0 = bf16(x), plane 1 = bf16(x - hi)  (the cross-term operands).
__device__ __forceinline__ void store_x3(float* hi, void* lo, size_t count, size_t off, const float4& v) {
  float4 h;
  h.x = to_tf32_rna(v.x); h.y = to_tf32_rna(v.y); h.z = to_tf32_rna(v.z); h.w = to_tf32_rna(v.w);
  *reinterpret_cast<float4*>(hi + off) = h;
  uint16_t* planes = static_cast<uint16_t*>(lo);
  *reinterpret_cast<uint2*>(planes + off) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  *reinterpret_cast<uint2*>(planes + count + off) =
      make_uint2(pack_bf16x2(v.x - h.x, v.y - h.y), pack_bf16x2(v.z - h.z, v.w - h.w));
}

// The F16X3 operand: two fp16 planes h0 = fp16(s*x), h1 = fp16(s*x - h0) of a per-tensor power-of-two scale s.
// A 128-byte header sits immediately before plane 0 (so every entry point keeps its (hi, lo) pointer pair):
// the scale the producer uses, its inverse for the GEMM epilogue, and the largest |x| the last producer saw,
// from which vv_operand_rescale picks the next scale.
// layout: 0 = two planes (hi = h0, lo = h1); 1 = row-interleaved in blocks of 64 elements: a row of K elements is
// K/64 blocks [64 x h0 | 64 x h1] (256 contiguous bytes), the form the gather producers prefer for the resident bank
struct F16Hdr { float scale; float inv_scale; uint32_t absmax_bits; uint32_t layout; };
__host__ __device__ __forceinline__ F16Hdr* f16_hdr(const void* hi) {
  return reinterpret_cast<F16Hdr*>(const_cast<char*>(static_cast<const char*>(hi)) - VV_F16X3_HEADER_BYTES);
}
__device__ __forceinline__ uint32_t pack_f16x2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // low half <- a
  return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t u) {
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
}
__device__ __forceinline__ void store_f16x3(void* h0, void* h1, size_t off, const float4& v, float scale, float& amax) {
  amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  const float x0 = v.x * scale, x1 = v.y * scale, x2 = v.z * scale, x3 = v.w * scale;
  const uint32_t p01 = pack_f16x2_sat(x0, x1), p23 = pack_f16x2_sat(x2, x3);
  const float2 f01 = unpack_f16x2(p01), f23 = unpack_f16x2(p23);
  *reinterpret_cast<uint2*>(static_cast<uint16_t*>(h0) + off) = make_uint2(p01, p23);
  *reinterpret_cast<uint2*>(static_cast<uint16_t*>(h1) + off) =
      make_uint2(pack_f16x2_sat(x0 - f01.x, x1 - f01.y), pack_f16x2_sat(x2 - f23.x, x3 - f23.y));
}
// one atomicMax per warp of the largest |x| a producer wrote (positive floats order like their bit patterns)
__device__ __forceinline__ void f16_publish_absmax(const void* hi, float amax) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(&f16_hdr(hi)->absmax_bits, __float_as_uint(amax));
}

// K4's arithmetic for 4 consecutive weights (float4 index i of the [N,K] blob) given their summed gradient g: scale,
// decay (L2 / L1), momentum, update, diff := history, W[:, K-1] saved, operand copy refreshed.  One definition for the
// stand-alone update kernel and for the split-K finish inside the wgrad kernels, so both round identically
// (ref: solver.cpp:534-568, net.cpp:837, blob.cpp:126-128).
__device__ __forceinline__ void sgd_update4(const UpdateTail& u, long long i, float4 g, float scale, float& amax) {
  const long long n4 = u.count / 4, k4 = u.K / 4;
  const float rate = u.rate_w, momentum = u.momentum, decay = u.decay_w, gscale = u.gscale;
  if (gscale != 1.f) { g.x *= gscale; g.y *= gscale; g.z *= gscale; g.w *= gscale; }
  float4 w = reinterpret_cast<float4*>(u.W)[i];
  if (decay != 0.f) {
    if (u.reg_type == 2) {
      g.x = fmaf(decay, w.x, g.x); g.y = fmaf(decay, w.y, g.y); g.z = fmaf(decay, w.z, g.z); g.w = fmaf(decay, w.w, g.w);
    } else {
      g.x += decay * float((0.f < w.x) - (w.x < 0.f)); g.y += decay * float((0.f < w.y) - (w.y < 0.f));
      g.z += decay * float((0.f < w.z) - (w.z < 0.f)); g.w += decay * float((0.f < w.w) - (w.w < 0.f));
    }
  }
  float4 h = reinterpret_cast<float4*>(u.hist)[i];
  h.x = fmaf(rate, g.x, momentum * h.x); h.y = fmaf(rate, g.y, momentum * h.y);
  h.z = fmaf(rate, g.z, momentum * h.z); h.w = fmaf(rate, g.w, momentum * h.w);
  w.x -= h.x; w.y -= h.y; w.z -= h.z; w.w -= h.w;
  reinterpret_cast<float4*>(u.hist)[i] = h;
  reinterpret_cast<float4*>(u.W)[i] = w;
  if (u.diff_out) reinterpret_cast<float4*>(u.diff_out)[i] = h;
  if ((i % k4) == k4 - 1 && u.col_out) u.col_out[i / k4] = w.w;
  float* hi = static_cast<float*>(u.Wop_hi);
  if (u.prec == VV_PREC_TF32X3 && hi) {
    store_x3(hi, u.Wop_lo, size_t(n4) * 4, size_t(i) * 4, w);
  } else if (u.prec == VV_PREC_F16X3 && hi) {
    store_f16x3(hi, u.Wop_lo, size_t(i) * 4, w, scale, amax);
  } else if (u.prec == VV_PREC_BF16 && hi) {
    reinterpret_cast<uint2*>(hi)[i] = make_uint2(pack_bf16x2(w.x, w.y), pack_bf16x2(w.z, w.w));
  }
}
__device__ __forceinline__ void sgd_update_bias1(const UpdateTail& u, int i) {
  float g = u.db[i];
  if (u.gscale != 1.f) g *= u.gscale;
  const float w = u.b[i];
  if (u.decay_b != 0.f) g = (u.reg_type == 2) ? fmaf(u.decay_b, w, g) : g + u.decay_b * float((0.f < w) - (w < 0.f));
  const float h = fmaf(u.rate_b, g, u.momentum * u.bh[i]);
  u.bh[i] = h; u.b[i] = w - h;
  if (u.b_diff) u.b_diff[i] = h;
}

// ----------------------------------------------------------------------------
// counter-based RNGs
// ----------------------------------------------------------------------------
// splitmix64 finaliser: the synthetic-bank hash (bit-identical on host & device)
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// relu(approx N(0,1)): Irwin-Hall of four 16-bit uniforms, integer arithmetic, one
// exact int->float conversion and one multiply => identical on CPU and GPU.
__host__ __device__ __forceinline__ float bank_value(uint64_t seed, uint64_t elem) {
  const uint64_t h = splitmix64(seed * 0xD1342543DE82EF95ull + elem);
  const int t = int(h & 0xFFFF) + int((h >> 16) & 0xFFFF) + int((h >> 32) & 0xFFFF) +
                int((h >> 48) & 0xFFFF) - 131070;
  // var of the sum = 4 * (65536^2-1)/12 ; scale to unit variance
  const float v = float(t) * (1.0f / 37837.0f);
  return v > 0.f ? v : 0.f;
}

// Philox4x32-10, one 128-bit block per call.
__host__ __device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                    uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = uint64_t(0xD2511F53u) * c0;
    const uint64_t p1 = uint64_t(0xCD9E8D57u) * c2;
    const uint32_t n0 = uint32_t(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = uint32_t(p1);
    const uint32_t n2 = uint32_t(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = uint32_t(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// The dropout stream: element (row, col4*4 + j) of step `step` draws word j of
// philox(counter = {col4, row, step_lo, step_hi}, key = seed).  keep iff u32 > thres
// (same comparison as the reference GPU layer, dropout_layer.cu:19).
__host__ __device__ __forceinline__ void dropout_words(uint64_t seed, uint64_t step, uint32_t row,
                                                       uint32_t col4, uint32_t out[4]) {
  philox4x32(col4, row, uint32_t(step), uint32_t(step >> 32), uint32_t(seed), uint32_t(seed >> 32), out);
}
// The cheap alternative stream (VV_DROPOUT_HASH): one 32-bit integer hash per element instead of a Philox block per
// four.  h(x) = lowbias32 (two multiplies, three xor-shifts; avalanche bias < 0.2 %): word(row, col) =
// h(h(h(seed_lo ^ h(seed_hi ^ step_lo) ^ step_hi) ^ row) + col * golden), ~8 integer ops per element against ~25 --
// inside the fc7 epilogue the RNG arithmetic costs tensor-pipe clocks through the power cap.
__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t dropout_hash_base(uint64_t seed, uint64_t step) {
  return hash32(uint32_t(seed) ^ hash32(uint32_t(seed >> 32) ^ uint32_t(step)) ^ uint32_t(step >> 32));
}
__host__ __device__ __forceinline__ void dropout_words_hash(uint32_t base, uint32_t row, uint32_t col4, uint32_t out[4]) {
  const uint32_t rb = hash32(base ^ row);
  const uint32_t c = col4 * 4u;
  out[0] = hash32(rb + (c + 0u) * 0x9E3779B9u); out[1] = hash32(rb + (c + 1u) * 0x9E3779B9u);
  out[2] = hash32(rb + (c + 2u) * 0x9E3779B9u); out[3] = hash32(rb + (c + 3u) * 0x9E3779B9u);
}
__host__ __device__ __forceinline__ uint32_t dropout_uint_thres(float ratio) {
  return static_cast<unsigned int>(4294967295u * ratio);   // UINT_MAX * threshold_ in float (dropout_layer.cpp:21)
}
__host__ __device__ __forceinline__ float dropout_scale(float ratio) {
  return (float)(1. / (1. - ratio));                        // dropout_layer.cpp:20
}

#if defined(__CUDA_ARCH__)
// ----------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) { }
}

// ----------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 2D tiles
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)),
         "r"(c0), "r"(c1)
      : "memory");
}
// 4 rows (any order, by index) x one box width of columns; the tensor map's box is {cols, 1}
// 16-byte asynchronous global->shared copy (LDGSTS); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
// same, asking L2 to fetch the whole 256-byte block around the source: a K-major row gather walks along the row, so
// the neighbouring 128 bytes are the next k-block's piece -- one DRAM access per row and two k-blocks instead of two
__device__ __forceinline__ void cp_async16_l2_256(uint32_t smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16, %2;" :: "r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (pre-counted) arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const void* tmap, uint64_t* bar,
                                                 int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)),
         "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// multicast: the tile lands at the same smem offset in every CTA of `mask` and completes tx on each one's mbarrier
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)),
         "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :: "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory");
}
// L2 eviction policies (createpolicy encodings used by CUTLASS' TMA::CacheHintSm90)
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst  = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast   = 0x14F0000000000000ull;

// ----------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc]
template <bool kTF32>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                        uint32_t idesc, uint32_t accumulate) {
  if constexpr (kTF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// ---- cta_group::2: one MMA over a CTA pair (M = 256: 128 rows of A and of D per CTA, each CTA stages half of B) ----
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
// issued by ONE thread of the leader CTA (rank 0) for both CTAs; descriptors name the leader's smem offsets, the hardware
// reads the same offsets in the peer
__device__ __forceinline__ void umma2_ss_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (when the MMAs issued so far retire) on the barrier at this smem offset in every CTA of `mask`
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"(mask) : "memory");
}
// TMA tile into THIS CTA's smem, transaction bytes credited to the LEADER CTA's mbarrier at the same offset (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & 0xFEFFFFFFu),
         "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the mbarrier at this smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(ra) : "memory");
}
// wait with cluster-scope acquire (the barrier also receives arrivals from the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// arrive on an mbarrier when all previously issued MMAs of this thread retire
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(smem_u32(bar)) : "memory");
}
// same, arriving on the barrier at this smem offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"(mask) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (tcgen05), version 1.
//  start address >>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout type [61,64)
//  layout type: 2 = SWIZZLE_128B (16-byte chunks), 1 = SWIZZLE_128B_BASE32B (32-byte chunks; the only
//  layout the hardware accepts for MN-major 32-bit (tf32) operands)
constexpr uint32_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1, kLayoutSw64 = 4;
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= uint64_t((saddr >> 4) & 0x3FFF);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(layout_type) << 61;
  return d;
}
#endif  // __CUDA_ARCH__

// Instruction descriptor (kind::f16 / kind::tf32), fp32 accumulate.
//  c_format [4,6)=1 (F32) | a_format [7,10) | b_format [10,13) | a_major 15 | b_major 16 |
//  N>>3 [17,23) | M>>4 [24,29).   format: 1 = BF16, 2 = TF32.  major: 0 = K, 1 = MN.
// fmt: 0 = F16, 1 = BF16 (kind::f16); 2 = TF32 (kind::tf32)
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int a_mn_major, int b_mn_major, int M, int N) {
  return (1u << 4) | (uint32_t(fmt) << 7) | (uint32_t(fmt) << 10) | (uint32_t(a_mn_major) << 15) |
         (uint32_t(b_mn_major) << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

}  // namespace vv
