// vv_gemm_tc.cu -- K1: the fc7 projection GEMMs on 5th-gen tensor cores.
//
// One persistent, warp-specialised kernel template for sm_100a:
//   warp 0   : TMA producer  (cp.async.bulk.tensor 2D tiles, SWIZZLE_128B, mbarrier tx)
//   warp 1   : MMA issuer    (tcgen05.mma, fp32 accum in TMEM; the warp runs converged and one elected lane issues:
//              cta_group::1, 128 x BLOCK_N x (32 B of K) per CTA -- or, Cfg::two_cta, cta_group::2: 256 x BLOCK_N per CTA
//              PAIR, issued by the leader CTA, each CTA staging its own A rows and half of the B tile)
//   warp 2   : TMEM allocator; warps 2-3: cp.async row-gather producers of operand A in the gather-fused variants
//   warps 4-11: epilogue     (tcgen05.ld 32x32b -> bias / ReLU / dropout / scale -> global)
// Two TMEM accumulator buffers (2 x BLOCK_N columns) let the epilogue of tile i
// overlap the main loop of tile i+1.
//
// The three contractions of the reference layer (ref: inner_product_layer.cu:12-59)
// differ only in operand major-ness, expressed in the UMMA descriptors:
//   FWD   D[M,N] = X W^T      A = X  [M,K] K-major      B = W [N,K] K-major
//   WGRAD D[N,K] = dZ^T X     A = dZ [M,N] MN-major     B = X [M,K] MN-major   (reduce M, split-K slabs)
//   DGRAD D[M,K] = dZ W       A = dZ [M,N] K-major      B = W [N,K] MN-major   (reduce N)
// Precisions: bf16 (kind::f16), tf32 (kind::tf32), and tf32x3 = split operands x = hi + lo with three products
// per element pair: hi*hi on the tf32 pipe and the two cross terms bf16(x)*bf16(lo) on the (2x faster) bf16
// pipe, all accumulated in the same fp32 TMEM tile -> fp32-level accuracy at 4 instead of 6 MMA units;
// and f16x3 = scaled operands s*x = h0 + h1 in fp16 with the three products h1*h0' + h0*h1' + h0*h0' on the f16
// pipe (3 MMA units, 4 bytes per element), unscaled by 1/(s s') in the epilogue.
#include <cuda.h>
#include <stdlib.h>
#include "vv_gemm.cuh"

namespace vv {

namespace {

constexpr int kBlockM = 128;
constexpr int kRowBytes = 128;            // bytes of the contiguous dim per smem row (= swizzle span)
constexpr int kNumThreads = 384;          // 4 control warps + 8 epilogue warps

struct TcParams {
  int d_rows, d_cols, ldd;      // output extent and row pitch
  int tiles_m, tiles_n, nsplit;
  int num_kb, kb_per_split;     // k-blocks (of kRowBytes worth of elements)
  int chunk_kb;                 // k-blocks accumulated in TMEM before promotion to fp32 registers
  float* D; long long slab_stride;
  int act_N;                    // row pitch of mask/Z (= N of the layer) for the FWD epilogue
  const int* rowmap;            // gather variants: bank row of every X row (padded to a multiple of 128)
  const uint16_t* ga0; const uint16_t* ga1;   // gather variants: the bank's operand plane(s), row pitch ga_pitch elements
  long long ga_pitch; int ga_cols; int rowmap_len;
  const float* inv_sa; const float* inv_sb;   // f16x3: 1/scale of the two operands (device, from their headers)
  DpWait wait;                  // data parallel: owner ranks' w_ready flags to wait for before the first B (= W) tile
  WgradFinish fin;              // wgrad: fused split-K finish (fin.tickets == NULL: off)
  // FWD tail split: the units of the last, partial wave are cut in two along K (tail_first < 0: off).  Virtual unit
  // tail_first + 2 i is the first K half of unit tail_first + i, tail_first + 2 i + 1 the second; the first half leaves its
  // raw fp32 partial in tail_ws, the second adds it in front of the bias / activation epilogue.
  int tail_first, tail_count, tail_kb;      // tail_kb: k-blocks of the first half
  float* tail_ws; unsigned int* tail_flags; unsigned int tail_epoch;
  GemmEpilogue epi;
};

// what a (virtual) work unit is
struct UnitInfo { int t, split, kb0, kb1, half /*0 whole, 1 first K half, 2 second*/, tail_idx; };
__device__ __forceinline__ UnitInfo decode_unit(const TcParams& p, int u, int tiles_mn) {
  UnitInfo ui;
  if (p.tail_first < 0 || u < p.tail_first) {
    ui.split = u / tiles_mn; ui.t = u - ui.split * tiles_mn;
    ui.kb0 = ui.split * p.kb_per_split; ui.kb1 = min(ui.kb0 + p.kb_per_split, p.num_kb);
    ui.half = 0; ui.tail_idx = 0;
  } else {
    const int v = u - p.tail_first;
    ui.tail_idx = v >> 1; ui.half = 1 + (v & 1);
    ui.split = 0; ui.t = p.tail_first + ui.tail_idx;               // the tail split exists for nsplit == 1 only
    ui.kb0 = (ui.half == 1) ? 0 : p.tail_kb; ui.kb1 = (ui.half == 1) ? p.tail_kb : p.num_kb;
  }
  return ui;
}

template <bool kTF32, bool kAMN, bool kBMN, int kNProd, int kBlockN, int kStages, bool kFwdEpi, bool kGather = false,
          bool kF16 = false, bool kTransOut = false, bool kTwoCta = false>
struct Cfg {
  // kTwoCta: ONE tcgen05.mma.cta_group::2 per k-step over the CTA pair (M = 256: each CTA holds 128 rows of A and of D in
  // its own smem / TMEM) and each CTA stages only ITS half of the B tile (128 of the 256 W rows, no multicast): the smem
  // fill per CTA and k-block drops from A + B to A + B/2 and a stage shrinks by a third.  Only the leader CTA (rank 0)
  // issues MMAs; the peer's TMA credits its bytes to the leader's full barrier, its gather producers complete on a local
  // barrier that the peer's (otherwise idle) MMA warp relays to the leader.  Built for the gathered forward (K-major, 2-byte).
  static constexpr bool two_cta = kTwoCta;
  static_assert(!kTwoCta || (kGather && !kTF32 && (kNProd == 1 || kF16) &&
                             ((!kAMN && !kBMN && kFwdEpi) || (kAMN && kBMN && kTransOut && !kFwdEpi))),
                "cta_group::2 variant: the gathered forward and the gathered (transposed) weight gradient");
  static constexpr bool f16 = kF16;                    // fp16 elements (kind::f16 with F16 formats) instead of bf16
  static_assert(!(kF16 && kTF32), "f16 and tf32 exclude each other");
  // kGather: operand A (X: the rows of FWD, the k-rows of WGRAD_T) is gathered row-wise from the bank's operand copy by
  // warps 2-3 with 16-byte cp.async straight into the swizzled tile (2-byte element types only); B still arrives by TMA
  static constexpr bool gather = kGather;
  static_assert(!(kGather && kTF32), "the gather producers are written for 2-byte elements");
  static constexpr bool trans_out = kTransOut;         // store D transposed (WGRAD_T)
  // 2-CTA clusters share the B tile: each CTA loads half of it with TMA multicast, which cuts the L2->smem
  // fill per CTA from A+B to A+B/2
  static constexpr int cluster = 2;
  static constexpr bool tf32 = kTF32;
  static constexpr bool a_mn = kAMN, b_mn = kBMN;
  static constexpr int nprod = kNProd;                 // 1 or 3
  static constexpr bool mixed = (kNProd == 3) && !kF16; // tf32 hi*hi + bf16 cross terms
  static constexpr bool split16 = (kNProd == 3) && kF16; // fp16 h0*h0' + the two fp16 cross terms
  static constexpr bool promote = (kNProd == 3);        // chunked TMEM accumulation + fp32 register sums
  static constexpr int block_n = kBlockN;
  static constexpr int stages = kStages;
  static constexpr bool fwd_epi = kFwdEpi;
  static constexpr int elem_bytes = kTF32 ? 4 : 2;
  static constexpr int bk = kRowBytes / elem_bytes;    // reduction elements per k-block (32 / 64)
  static constexpr int umma_k = 32 / elem_bytes;       // 8 / 16
  static constexpr int ksteps = bk / umma_k;           // 4
  static constexpr int chunk = kRowBytes / elem_bytes; // MN elements per 128B row (MN-major)
  static constexpr int a_bytes = kBlockM * kRowBytes;  // 16 KB: the main (hi) tile
  static constexpr int b_bytes = (kBlockN / (kTwoCta ? 2 : 1)) * kRowBytes;  // 32 KB (BLOCK_N = 256); cta_group::2: this CTA's half
  // mixed mode: two extra bf16 tiles per operand (bf16(x) and bf16(lo)) covering the same bk = 32 reduction
  // elements: K-major = rows of 64 B (SWIZZLE_64B); MN-major = 64-element chunks of [32 k-rows x 128 B] (SWIZZLE_128B)
  static constexpr int a_half = a_bytes / 2, b_half = b_bytes / 2;
  // split16: a second full-size tile per operand (the h1 plane), stage = [A_h0][A_h1][B_h0][B_h1]
  static constexpr int stage_bytes = (mixed || split16) ? 2 * (a_bytes + b_bytes) : (a_bytes + b_bytes);
  static constexpr int tmem_cols = 2 * kBlockN;
  static constexpr int smem_bytes = stages * stage_bytes + 1024 /*align slack*/ + 256 /*barriers: (3 * stages + 4) * 8 + 8 bytes*/;
  static_assert((3 * kStages + 4) * 8 + 8 <= 256, "barrier area");
  static_assert(tmem_cols == 256 || tmem_cols == 512, "TMEM allocation must be a power of two");
  static_assert(smem_bytes <= 232448, "exceeds 227 KB of shared memory");
};


// ---- the split-K finish (see WgradFinish in vv_gemm.cuh), run by the 256 epilogue threads ------------------------
// (1) after a unit's partial tile is stored: publish it and count the arrival on the tile's ticket.
template <class C>
__device__ __forceinline__ void wgrad_arrive_unit(const TcParams& p, int m0, int nt0, int et) {
  if (m0 >= p.d_rows || nt0 >= p.d_cols) return;                 // the phantom tile of an odd pair (CTA-uniform)
  __threadfence();
  asm volatile("bar.sync 1, 256;" ::: "memory");
  if (et == 0) atomicAdd(&p.fin.tickets[(m0 / kBlockM) * p.tiles_n + nt0 / C::block_n], 1u);
}
__device__ __forceinline__ void st_release_gpu(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// (2) after the CTA's last unit: the drain.  Every output tile is cut into chunks of 1024 float4 (8 per full tile), chunk c
// goes to CTA c % gridDim.x: ALL CTAs finish tiles, whoever computed them.  A chunk waits until its tile has collected the
// arrivals of all nsplit units (tickets only grow: target = nsplit * epoch), then every thread takes 4 float4 elements with
// all their loads in flight together: the S partial sums in slab order (fixed order: the result does not depend on which
// split finished last), then the update (mode 1) or the push to the owner rank (mode 2).
// No deadlock: a CTA drains only after its own units, and arrivals never wait for anything.
template <class C>
__device__ __forceinline__ void wgrad_drain(const TcParams& p, int et) {
  const int ntiles = p.tiles_m * p.tiles_n;
  constexpr int kChunk4 = 1024, kPer = kChunk4 / 256;            // float4 per chunk / per thread
  constexpr int kTile4 = kBlockM * C::block_n / 4;
  constexpr int kParts = kTile4 / kChunk4;                       // 8
  const int Kdim = p.ldd;
  const unsigned int target = unsigned(p.nsplit) * p.fin.epoch;
  const float* col_add = p.fin.mode == 1 ? p.fin.u.col_add : p.fin.col_add;
  float* hi = static_cast<float*>(p.fin.u.Wop_hi);
  const float scale = (p.fin.mode == 1 && p.fin.u.prec == VV_PREC_F16X3 && hi) ? f16_hdr(hi)->scale : 1.f;
  const long long owned4 = (long long)p.fin.rows_per * (Kdim >> 2);
  float amax = 0.f;
  for (int c = blockIdx.x; c < ntiles * kParts; c += gridDim.x) {
    const int tile = c / kParts, part = c - tile * kParts;
    const int m0 = (tile / p.tiles_n) * kBlockM, nt0 = (tile % p.tiles_n) * C::block_n;
    // the tile as a region of dW [N, K] (row pitch Kdim = K): WGRAD_T stores D transposed
    int n_begin, n_count, k_begin, k_count;
    if (C::trans_out) { n_begin = nt0; n_count = min(C::block_n, p.d_cols - nt0); k_begin = m0; k_count = min(kBlockM, p.d_rows - m0); }
    else              { n_begin = m0;  n_count = min(kBlockM, p.d_rows - m0);    k_begin = nt0; k_count = min(C::block_n, p.d_cols - nt0); }
    const int k4c = k_count >> 2, total = n_count * k4c;
    if (part * kChunk4 >= total) continue;                       // ragged tile: this part is empty (uniform)
    if (et == 0) { while (ld_acquire_gpu(&p.fin.tickets[tile]) < target) __nanosleep(64); }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    long long ii[kPer]; int nn[kPer]; bool ok[kPer], lastc[kPer];
    float4 g[kPer];
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
      const int idx = part * kChunk4 + e * 256 + et;
      ok[e] = idx < total;
      const int r = ok[e] ? idx / k4c : 0, c4 = ok[e] ? idx - r * k4c : 0;
      nn[e] = n_begin + r;
      const int k = k_begin + c4 * 4;
      lastc[e] = (k + 4 == Kdim);
      ii[e] = ((long long)nn[e] * Kdim + k) >> 2;                // float4 index in [N, K]
      g[e] = ok[e] ? __ldcg(reinterpret_cast<const float4*>(p.D) + ii[e]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int e = 0; e < kPer; ++e)
      if (ok[e] && col_add && lastc[e]) g[e].w += col_add[nn[e]];                // .w is column K-1
    for (int sp = 1; sp < p.nsplit; sp += 2) {                                   // slab order; two slabs x 4 elements in flight
      const bool two = sp + 1 < p.nsplit;
      float4 t0[kPer], t1[kPer];
#pragma unroll
      for (int e = 0; e < kPer; ++e) {
        t0[e] = ok[e] ? __ldcg(reinterpret_cast<const float4*>(p.D + (long long)sp * p.slab_stride) + ii[e]) : make_float4(0.f, 0.f, 0.f, 0.f);
        t1[e] = (ok[e] && two) ? __ldcg(reinterpret_cast<const float4*>(p.D + (long long)(sp + 1) * p.slab_stride) + ii[e]) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int e = 0; e < kPer; ++e) {
        g[e].x += t0[e].x; g[e].y += t0[e].y; g[e].z += t0[e].z; g[e].w += t0[e].w;
        if (two) { g[e].x += t1[e].x; g[e].y += t1[e].y; g[e].z += t1[e].z; g[e].w += t1[e].w; }
      }
    }
    if (p.fin.mode == 1) {
#pragma unroll
      for (int e = 0; e < kPer; ++e) if (ok[e]) sgd_update4(p.fin.u, ii[e], g[e], scale, amax);
      if (k_begin == 0) {                                        // this block of the bias blob goes with feature tile 0
        // the rows whose first element lies in this chunk
        const int r_lo = (part * kChunk4 + k4c - 1) / k4c, r_hi = min(n_count, ((part + 1) * kChunk4 + k4c - 1) / k4c);
        for (int r = r_lo + et; r < r_hi; r += 256) sgd_update_bias1(p.fin.u, n_begin + r);
      }
    } else {
#pragma unroll
      for (int e = 0; e < kPer; ++e) {
        if (!ok[e]) continue;
        const int o = nn[e] / p.fin.rows_per;                                     // owner rank of row n
        reinterpret_cast<float4*>(p.fin.peers.recv_dw[o])[(long long)p.fin.rank * owned4 + (ii[e] - (long long)o * owned4)] = g[e];
      }
      // count the finished chunks; the CTA finishing the last one completes this rank's push
      __threadfence_system();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      __shared__ unsigned int s_last;
      if (et == 0) {
        int nchunks = 0;                                         // non-empty chunks of this launch
        for (int t2 = 0; t2 < ntiles; ++t2) {
          const int mm = (t2 / p.tiles_n) * kBlockM, nn0 = (t2 % p.tiles_n) * C::block_n;
          const int tot = C::trans_out ? min(C::block_n, p.d_cols - nn0) * (min(kBlockM, p.d_rows - mm) >> 2)
                                       : min(kBlockM, p.d_rows - mm) * (min(C::block_n, p.d_cols - nn0) >> 2);
          nchunks += (tot + kChunk4 - 1) / kChunk4;
        }
        const unsigned int prev = atomicAdd(&p.fin.tickets[ntiles], 1u);
        s_last = (prev + 1u == unsigned(nchunks) * p.fin.epoch) ? 1u : 0u;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (s_last != 0u) {
        __threadfence();
        for (int d = 0; d < p.fin.G; ++d)                        // (db, loss, violations) to every rank
          for (int i = et; i < p.fin.nsmall; i += 256) p.fin.peers.recv_small[d][p.fin.rank * p.fin.small_stride + i] = p.fin.small_src[i];
        __threadfence_system();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // one thread per destination rank: the G release stores (system fence + store over NVLink each) run side by side
        if (et < p.fin.G) { __threadfence_system(); dp_st_release_sys(&p.fin.peers.flags[et][kDpFlagDwReady + p.fin.rank], p.fin.seq); }
      }
    }
  }
  if (p.fin.mode == 1 && p.fin.u.prec == VV_PREC_F16X3 && hi) f16_publish_absmax(hi, amax);
}

template <class C>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_hb,
               const __grid_constant__ CUtensorMap tmA_lb, const __grid_constant__ CUtensorMap tmB_hi,
               const __grid_constant__ CUtensorMap tmB_hb, const __grid_constant__ CUtensorMap tmB_lb,
               const TcParams p) {
#if defined(__CUDA_ARCH__)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::stages * C::stage_bytes);
  uint64_t* empty_bar = full_bar + C::stages;
  uint64_t* tmem_full = empty_bar + C::stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* gfull_bar = tmem_empty + 2;            // cta_group::2, peer CTA: its gather producers' completion (relayed to the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gfull_bar + C::stages);

  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);   // warp-uniform for the compiler: role branches do not diverge
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmB_hi);
    if (C::mixed) { tma_prefetch_desc(&tmA_hb); tma_prefetch_desc(&tmA_lb); tma_prefetch_desc(&tmB_hb); tma_prefetch_desc(&tmB_lb); }
    if (C::split16) { tma_prefetch_desc(&tmA_hb); tma_prefetch_desc(&tmB_hb); }
  }
  if (warp == 1 && lane == 0) {
    // full: the TMA lane's arrive.expect_tx (+ one cp.async-completion arrival per gather lane, 2 warps)
    if (C::two_cta) {
      // leader's full: its TMA lane's arrive.expect_tx (bytes of BOTH CTAs) + its 64 gather lanes + the peer's relay;
      // empty / tmem_full: one multicast commit of the single MMA issuer; leader's tmem_empty: 8 local + 8 remote epilogue warps
      for (int s = 0; s < C::stages; ++s) { mbar_init(&full_bar[s], 1 + 64 + 1); mbar_init(&empty_bar[s], 1); mbar_init(&gfull_bar[s], 64); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 16); }
    } else {
      for (int s = 0; s < C::stages; ++s) { mbar_init(&full_bar[s], C::gather ? 1 + 64 : 1); mbar_init(&empty_bar[s], C::cluster); }
      for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 8); }
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (C::two_cta) { tmem_alloc2(tmem_slot, C::tmem_cols); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, C::tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (C::cluster > 1) cluster_sync_all(); else __syncthreads();   // barrier inits visible to the peer before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Work units.  Without clusters: unit = (split, m-tile, n-tile), CTAs stride over units.  With 2-CTA clusters:
  // unit = (split, PAIR of adjacent m-tiles, n-tile), clusters stride over units and CTA `rank` takes m-tile 2*pair+rank.
  const int cta_rank = (C::cluster > 1) ? int(cluster_ctarank()) : 0;
  const int tiles_mp = (p.tiles_m + C::cluster - 1) / C::cluster;
  const int tiles_mn = tiles_mp * p.tiles_n;
  const int total_units = p.tail_first < 0 ? tiles_mn * p.nsplit : tiles_mn + p.tail_count;
  const int unit0 = blockIdx.x / C::cluster, unit_stride = gridDim.x / C::cluster;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      constexpr int kTmaBytes = C::stage_bytes - (C::gather ? ((C::mixed || C::split16) ? 2 * C::a_bytes : C::a_bytes) : 0);
      int stage = 0; uint32_t phase = 0;
      if (p.wait.flags) {
        // the weights arrive from their owner ranks (vv_dp_exchange.cuh): everything up to here -- and the row-gather
        // producers of operand A, which do not depend on W -- has run under the tail of the exchange
        for (int r = 0; r < p.wait.G; ++r)
          dp_spin_wait_flag(&p.wait.flags[kDpFlagWReady + r], p.wait.seq, p.wait.timeout_ns, p.wait.err, 2u);
        asm volatile("fence.proxy.async;" ::: "memory");   // peer stores (generic proxy) -> TMA reads (async proxy)
      }
      for (int u = unit0; u < total_units; u += unit_stride) {
        const UnitInfo ui = decode_unit(p, u, tiles_mn);
        const int t = ui.t;
        const int m0 = ((t / p.tiles_n) * C::cluster + cta_rank) * kBlockM;
        const int n0 = (t % p.tiles_n) * C::block_n;
        const int kb0 = ui.kb0, kb1 = ui.kb1;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          // stage layout: [A_hi][A_hb][A_lb][B_hi][B_hb][B_lb]  (the extra tiles only in the split modes)
          uint8_t* sa = smem + stage * C::stage_bytes;
          uint8_t* sb = sa + ((C::mixed || C::split16) ? 2 * C::a_bytes : C::a_bytes);
          const int k0 = kb * C::bk;
          if (C::two_cta) {
            // this CTA's half of the W tile (rows n0 + 128 * rank), into its own smem; the bytes of both CTAs complete on the
            // leader's barrier, which the leader arms for the pair
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kTmaBytes);
            if (!C::b_mn) {
              tma_load_2d_2cta(sb, &tmB_hi, &full_bar[stage], k0, n0 + cta_rank * (C::block_n / 2));
              if (C::split16) tma_load_2d_2cta(sb + C::b_bytes, &tmB_hb, &full_bar[stage], k0, n0 + cta_rank * (C::block_n / 2));
            } else {
              // MN-major B (the weight gradient's dZ tile): this CTA's half = kHalf chunks of [bk k-rows x 128 B of n]
              constexpr int kHalf = C::block_n / C::chunk / 2;
#pragma unroll
              for (int c = 0; c < kHalf; ++c) {
                tma_load_2d_2cta(sb + c * (C::bk * kRowBytes), &tmB_hi, &full_bar[stage], n0 + (cta_rank * kHalf + c) * C::chunk, k0);
                if (C::split16) tma_load_2d_2cta(sb + C::b_bytes + c * (C::bk * kRowBytes), &tmB_hb, &full_bar[stage],
                                                 n0 + (cta_rank * kHalf + c) * C::chunk, k0);
              }
            }
            if (++stage == C::stages) { stage = 0; phase ^= 1u; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], kTmaBytes);
          if (!C::gather) {
            if (!C::a_mn) {
              tma_load_2d(sa, &tmA_hi, &full_bar[stage], k0, m0);
            } else {
#pragma unroll
              for (int c = 0; c < kBlockM / C::chunk; ++c)
                tma_load_2d(sa + c * (C::bk * kRowBytes), &tmA_hi, &full_bar[stage], m0 + c * C::chunk, k0);
            }
          }
          {
            // this CTA fetches half of the shared B tile and multicasts it to both CTAs of the cluster
            if (!C::b_mn) {
              tma_load_2d_mc(sb + cta_rank * (C::b_bytes / 2), &tmB_hi, &full_bar[stage], k0, n0 + cta_rank * (C::block_n / 2), 0x3);
            } else {
              constexpr int kHalf = C::block_n / C::chunk / 2;
#pragma unroll
              for (int c = 0; c < kHalf; ++c)
                tma_load_2d_mc(sb + (cta_rank * kHalf + c) * (C::bk * kRowBytes), &tmB_hi, &full_bar[stage],
                               n0 + (cta_rank * kHalf + c) * C::chunk, k0, 0x3);
            }
          }
          if (C::split16 && lane == 0) {
            // the h1 planes: same boxes as the h0 tiles, through the second pair of tensor maps
            uint8_t* da = sa + C::a_bytes;
            uint8_t* db = sb + C::b_bytes;
            if (C::gather) {
              // the h1 plane of A is gathered together with h0 by warps 2-3
            } else if (!C::a_mn) {
              tma_load_2d(da, &tmA_hb, &full_bar[stage], k0, m0);
            } else {
#pragma unroll
              for (int c = 0; c < kBlockM / C::chunk; ++c)
                tma_load_2d(da + c * (C::bk * kRowBytes), &tmA_hb, &full_bar[stage], m0 + c * C::chunk, k0);
            }
            if (C::cluster > 1) {
              if (!C::b_mn) {
                tma_load_2d_mc(db + cta_rank * (C::b_bytes / 2), &tmB_hb, &full_bar[stage], k0, n0 + cta_rank * (C::block_n / 2), 0x3);
              } else {
                constexpr int kHalf = C::block_n / C::chunk / 2;
#pragma unroll
                for (int c = 0; c < kHalf; ++c)
                  tma_load_2d_mc(db + (cta_rank * kHalf + c) * (C::bk * kRowBytes), &tmB_hb, &full_bar[stage],
                                 n0 + (cta_rank * kHalf + c) * C::chunk, k0, 0x3);
              }
            } else if (!C::b_mn) {
              tma_load_2d(db, &tmB_hb, &full_bar[stage], k0, n0);
            } else {
#pragma unroll
              for (int c = 0; c < C::block_n / C::chunk; ++c)
                tma_load_2d(db + c * (C::bk * kRowBytes), &tmB_hb, &full_bar[stage], n0 + c * C::chunk, k0);
            }
          }
          if (C::mixed && lane == 0) {
            // the two bf16 tiles of each operand for the cross terms
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const CUtensorMap* ta = h ? &tmA_lb : &tmA_hb;
              const CUtensorMap* tb = h ? &tmB_lb : &tmB_hb;
              uint8_t* da = sa + C::a_bytes + h * C::a_half;
              uint8_t* db = sb + C::b_bytes + h * C::b_half;
              if (!C::a_mn) {
                tma_load_2d(da, ta, &full_bar[stage], k0, m0);                       // [128 rows x 64 B], SWIZZLE_64B
              } else {
#pragma unroll
                for (int c = 0; c < kBlockM / 64; ++c)                               // 64 bf16 of MN per 128 B row
                  tma_load_2d(da + c * (C::bk * kRowBytes), ta, &full_bar[stage], m0 + c * 64, k0);
              }
              if (C::cluster > 1) {
                if (!C::b_mn) {
                  tma_load_2d_mc(db + cta_rank * (C::b_half / 2), tb, &full_bar[stage], k0, n0 + cta_rank * (C::block_n / 2), 0x3);
                } else {
                  constexpr int kHalf16 = C::block_n / 64 / 2;
#pragma unroll
                  for (int c = 0; c < kHalf16; ++c)
                    tma_load_2d_mc(db + (cta_rank * kHalf16 + c) * (C::bk * kRowBytes), tb, &full_bar[stage],
                                   n0 + (cta_rank * kHalf16 + c) * 64, k0, 0x3);
                }
              } else if (!C::b_mn) {
                tma_load_2d(db, tb, &full_bar[stage], k0, n0);
              } else {
#pragma unroll
                for (int c = 0; c < C::block_n / 64; ++c)
                  tma_load_2d(db + c * (C::bk * kRowBytes), tb, &full_bar[stage], n0 + c * 64, k0);
              }
            }
          }
          if (++stage == C::stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (C::two_cta && cta_rank != 0) {
      // cta_group::2, peer CTA: no MMAs to issue.  This warp relays the completion of the peer's row gathers (cp.async into
      // the peer's own smem, counted on its local barrier) to the leader's full barrier, stage by stage, in the order
      // the leader consumes them.
      // One lane per stage (lane s relays stage s, phase after phase): the waits, proxy fences and remote arrivals of
      // different stages overlap -- a single relay thread's wait + fence + NVLink-free but cluster-wide arrive per k-block
      // is slower than one bf16 k-block of MMAs.
      if (lane < C::stages) {
        long long total_kb = 0;
        for (int u = unit0; u < total_units; u += unit_stride) { const UnitInfo ui = decode_unit(p, u, tiles_mn); total_kb += ui.kb1 - ui.kb0; }
        uint32_t phase = 0;
        for (long long i = lane; i < total_kb; i += C::stages, phase ^= 1u) {
          mbar_wait(&gfull_bar[lane], phase);
          fence_proxy_async();                      // the gathered tile (generic-proxy writes) before the leader's MMA reads it
          mbar_arrive_cluster(&full_bar[lane], 0);
        }
      }
    } else {
      // The whole warp runs this loop converged (every lane polls the barriers) and ONE elected lane issues the MMAs and
      // commits: with warp-uniform control flow and operands the compiler keeps descriptors, addresses and predicates in
      // uniform registers and feeds UTCHMMA directly.  Under `if (lane == 0)` every tcgen05.mma was wrapped in a
      // vote / elect / 5 x R2UR waterfall and every descriptor was rebuilt from scratch: ~110 instructions per bf16
      // k-block on a single thread, i.e. as long as the k-block's MMAs themselves (ncu: tensor pipe 59 % active).
      constexpr uint32_t idesc = make_idesc(C::tf32 ? 2 : (C::f16 ? 0 : 1), C::a_mn ? 1 : 0, C::b_mn ? 1 : 0,
                                            C::two_cta ? 2 * kBlockM : kBlockM, C::block_n);
      // K-major : rows of 128 B, 8-row atoms 1024 B apart (SBO); LBO unused (1)
      // MN-major: 128 B of MN contiguous, k-rows 128 B apart, 8-row atoms 1024 B apart (SBO),
      //           next MN chunk bk*128 B away (LBO)
      constexpr uint32_t lbo_a = C::a_mn ? C::bk * kRowBytes : 16;
      constexpr uint32_t lbo_b = C::b_mn ? C::bk * kRowBytes : 16;
      constexpr uint32_t kadv_a = C::a_mn ? C::umma_k * kRowBytes : 32;   // bytes per k-step
      constexpr uint32_t kadv_b = C::b_mn ? C::umma_k * kRowBytes : 32;
      // MN-major tf32 must use the 32-byte-chunk swizzle: atoms of 4 k-rows (512 B) instead of 8 (1024 B)
      constexpr uint32_t lay_a = (C::tf32 && C::a_mn) ? kLayoutSw128Base32 : kLayoutSw128;
      constexpr uint32_t lay_b = (C::tf32 && C::b_mn) ? kLayoutSw128Base32 : kLayoutSw128;
      constexpr uint32_t sbo_a = (C::tf32 && C::a_mn) ? 512 : 1024;
      constexpr uint32_t sbo_b = (C::tf32 && C::b_mn) ? 512 : 1024;
      // shared-memory descriptors = constant high word | (constant LBO field + address / 16): one add per descriptor
      // (the 14-bit address field cannot carry: shared memory ends below 256 KB).  The mask matters: in a cluster the
      // shared-window address of the rank-1 CTA carries its rank in bit 24, which would land in the LBO field.
      constexpr uint32_t dhi_a = ((sbo_a >> 4) & 0x3FFFu) | (1u << 14) | (lay_a << 29);
      constexpr uint32_t dhi_b = ((sbo_b >> 4) & 0x3FFFu) | (1u << 14) | (lay_b << 29);
      const uint32_t dlo_a = (((lbo_a >> 4) & 0x3FFFu) << 16) + ((smem_u32(smem) >> 4) & 0x3FFFu);
      const uint32_t dlo_b = (((lbo_b >> 4) & 0x3FFFu) << 16) + ((smem_u32(smem) >> 4) & 0x3FFFu);
      constexpr uint32_t kBOff = ((C::mixed || C::split16) ? 2 * C::a_bytes : C::a_bytes) >> 4;   // B tiles behind the A tiles, in 16-byte units
#define VV_DESC(hi, lo) ((uint64_t(hi) << 32) | uint64_t(lo))
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int u = unit0; u < total_units; u += unit_stride) {
        const UnitInfo ui = decode_unit(p, u, tiles_mn);
        const int kb0 = ui.kb0, kb1 = ui.kb1;
        // The k-range is issued in chunks of p.chunk_kb k-blocks, each into its own TMEM buffer
        // (accumulate = 0 at the chunk start): the epilogue warps sum the chunks in fp32 registers.
        // Non-promoting configurations use one chunk per unit.
        for (int c0 = kb0; c0 < kb1; c0 += p.chunk_kb) {
          const int c1 = min(c0 + p.chunk_kb, kb1);
          if (C::two_cta) mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1u);      // both CTAs' epilogues have drained the buffer
          else mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + uint32_t(acc * C::block_n);
          uint32_t accumulate = 0;
          for (int kb = c0; kb < c1; ++kb) {
            // (cta_group::2: the peer's arrivals are release.cluster; the operands they announce are read by the tensor core
            // through the async proxy, never through this SM's L1, so the default acquire is enough -- as in CUTLASS' pipelines)
            mbar_wait(&full_bar[stage], phase);
            if (elect_one()) {
              if (C::gather) fence_proxy_async();     // A was written by cp.async (generic proxy); the MMA reads through the async proxy
              uint32_t accf = accumulate;
              const uint32_t so = uint32_t(stage) * uint32_t(C::stage_bytes >> 4);     // this stage, in 16-byte units
              if (C::mixed) {
                // cross terms first (small magnitude), on the bf16 pipe: bf16(a)*bf16(b_lo) + bf16(a_lo)*bf16(b)
                const uint32_t sa = smem_u32(smem + stage * C::stage_bytes);
                const uint32_t sb = sa + 2 * C::a_bytes;
                constexpr uint32_t idesc16 = make_idesc(1, C::a_mn ? 1 : 0, C::b_mn ? 1 : 0, kBlockM, C::block_n);
                constexpr uint32_t lay16_a = C::a_mn ? kLayoutSw128 : kLayoutSw64, lay16_b = C::b_mn ? kLayoutSw128 : kLayoutSw64;
                constexpr uint32_t sbo16_a = C::a_mn ? 1024 : 512, sbo16_b = C::b_mn ? 1024 : 512;
                constexpr uint32_t kadv16_a = C::a_mn ? 16 * kRowBytes : 32, kadv16_b = C::b_mn ? 16 * kRowBytes : 32;
                const uint32_t a_hb = sa + C::a_bytes, a_lb = a_hb + C::a_half;
                const uint32_t b_hb = sb + C::b_bytes, b_lb = b_hb + C::b_half;
#pragma unroll
                for (int k = 0; k < 2; ++k) {      // 32 reduction elements = 2 bf16 k-steps of 16
                  const uint64_t d_ahb = make_smem_desc(a_hb + k * kadv16_a, lbo_a, sbo16_a, lay16_a);
                  const uint64_t d_alb = make_smem_desc(a_lb + k * kadv16_a, lbo_a, sbo16_a, lay16_a);
                  const uint64_t d_bhb = make_smem_desc(b_hb + k * kadv16_b, lbo_b, sbo16_b, lay16_b);
                  const uint64_t d_blb = make_smem_desc(b_lb + k * kadv16_b, lbo_b, sbo16_b, lay16_b);
                  umma_ss<false>(d_tmem, d_alb, d_bhb, idesc16, accf);
                  umma_ss<false>(d_tmem, d_ahb, d_blb, idesc16, 1u);
                  accf = 1u;
                }
              }
#pragma unroll
              for (int k = 0; k < C::ksteps; ++k) {
                const uint64_t a_hi = VV_DESC(dhi_a, dlo_a + so + uint32_t(k * kadv_a >> 4));
                const uint64_t b_hi = VV_DESC(dhi_b, dlo_b + so + kBOff + uint32_t(k * kadv_b >> 4));
                if (C::split16) {
                  // cross terms first (small magnitude): h1*h0' + h0*h1', then h0*h0'
                  const uint64_t a_h1 = VV_DESC(dhi_a, dlo_a + so + uint32_t((C::a_bytes + k * kadv_a) >> 4));
                  const uint64_t b_h1 = VV_DESC(dhi_b, dlo_b + so + kBOff + uint32_t((C::b_bytes + k * kadv_b) >> 4));
                  if (C::two_cta) { umma2_ss_f16(d_tmem, a_h1, b_hi, idesc, accf); umma2_ss_f16(d_tmem, a_hi, b_h1, idesc, 1u); }
                  else { umma_ss<false>(d_tmem, a_h1, b_hi, idesc, accf); umma_ss<false>(d_tmem, a_hi, b_h1, idesc, 1u); }
                  accf = 1u;
                }
                if (C::two_cta) umma2_ss_f16(d_tmem, a_hi, b_hi, idesc, accf);
                else umma_ss<C::tf32>(d_tmem, a_hi, b_hi, idesc, accf);
                accf = 1u;
              }
              // frees the smem stage when these MMAs retire -- in BOTH CTAs of a cluster (the peer multicasts into it)
              if (C::two_cta) {
                umma2_commit_mc(&empty_bar[stage], 0x3);
                if (kb == c1 - 1) umma2_commit_mc(&tmem_full[acc], 0x3);        // the accumulator halves of both CTAs are ready
              } else {
                if (C::cluster > 1) umma_commit_mc(&empty_bar[stage], 0x3); else umma_commit(&empty_bar[stage]);
                if (kb == c1 - 1) umma_commit(&tmem_full[acc]);
              }
            }
            __syncwarp();
            accumulate = 1u;
            if (++stage == C::stages) { stage = 0; phase ^= 1u; }
          }
          acc ^= 1; if (acc == 0) acc_phase ^= 1u;
        }
      }
#undef VV_DESC
    }
  } else if (C::gather && (warp == 2 || warp == 3)) {
    // ===================== cp.async gather producers of operand A (2 warps) =====================
    // 64 lanes x 16-byte chunks: lane t copies chunk column c = t % 8 of the tile rows r0 + 8i (r0 = t / 8), so 8
    // consecutive lanes fetch one contiguous 128-byte piece of a bank row and the swizzle term (row % 8 = r0) is a
    // per-lane constant.  Completion is signalled to full_bar by cp.async.mbarrier.arrive.noinc (no waiting here:
    // as many stages in flight as the ring has); the MMA thread crosses to the async proxy with fence.proxy.async.
    const int t64 = (warp - 2) * 32 + lane;
    const int c = t64 & 7, r0 = t64 >> 3;
    const uint32_t swz = uint32_t((c ^ r0) << 4);
    constexpr int kPlanes = C::split16 ? 2 : 1;
    // operand addressing: element k of bank row r lives at r*pitch + (k/64)*blk + k%64 (+ h1d for the h1 plane);
    // two planes: pitch K, blk 64, h1d = plane distance; row-interleaved F16X3 bank: pitch 2K, blk 128, h1d 64
    const bool il = C::split16 && f16_hdr(p.ga0)->layout == 1u;
    const long long pitch = il ? 2 * p.ga_pitch : p.ga_pitch;
    const int blk = il ? 128 : 64;
    const long long h1d = il ? 64 : (kPlanes == 2 ? (long long)(p.ga1 - p.ga0) : 0);
    int stage = 0; uint32_t phase = 0;
    for (int u = unit0; u < total_units; u += unit_stride) {
      const UnitInfo ui = decode_unit(p, u, tiles_mn);
      const int t = ui.t;
      const int m0 = ((t / p.tiles_n) * C::cluster + cta_rank) * kBlockM;
      const int kb0 = ui.kb0, kb1 = ui.kb1;
      if (!C::a_mn) {
        // FWD: A tile = [128 X rows x 128 B of K]; this lane's 16 rows are fixed for the whole unit
        long long rowoff[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {     // an m-tile past the end (odd tile count in a cluster pair) reads row 0, stores nothing
          const int rr = m0 + r0 + 8 * i;
          rowoff[i] = (rr < p.rowmap_len) ? (long long)p.rowmap[rr] * pitch : 0;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          const uint32_t sa = smem_u32(smem + stage * C::stage_bytes);
          int kcol = kb * C::bk + c * 8;
          const int nb = kcol < p.ga_cols ? 16 : 0;                 // K tail: zero fill like TMA
          if (!nb) kcol = 0;                                        // keep the (unread) source address inside the bank
          const long long koff = (long long)(kcol >> 6) * blk + (kcol & 63);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t dst = sa + uint32_t((r0 + 8 * i) * kRowBytes) + swz;
            const uint16_t* src = p.ga0 + rowoff[i] + koff;
            cp_async16_l2_256(dst, src, nb);
            if (kPlanes == 2) cp_async16_l2_256(dst + C::a_bytes, src + h1d, nb);
          }
          cp_async_mbar_arrive_noinc((C::two_cta && cta_rank != 0) ? &gfull_bar[stage] : &full_bar[stage]);
          if (++stage == C::stages) { stage = 0; phase ^= 1u; }
        }
      } else {
        // WGRAD_T: A tile = 2 chunks of [bk X rows (the reduction) x 128 B of features]; rows change every k-block
        static_assert(!C::gather || C::bk == 64, "gather producers assume 64 reduction rows per k-block");
        int rws[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) rws[j] = p.rowmap[kb0 * C::bk + r0 + 8 * j];
        for (int kb = kb0; kb < kb1; ++kb) {
          long long rowoff[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) rowoff[j] = (long long)rws[j] * pitch;
          if (kb + 1 < kb1) {                                       // next k-block's row indices, ahead of the wait
#pragma unroll
            for (int j = 0; j < 8; ++j) rws[j] = p.rowmap[(kb + 1) * C::bk + r0 + 8 * j];
          }
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          const uint32_t sa = smem_u32(smem + stage * C::stage_bytes);
#pragma unroll
          for (int cf = 0; cf < kBlockM / C::chunk; ++cf) {
            int f = m0 + cf * C::chunk + c * 8;
            const int nb = f < p.ga_cols ? 16 : 0;                  // feature tail: zero fill
            if (!nb) f = 0;
            const long long foff = (long long)(f >> 6) * blk + (f & 63);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t dst = sa + uint32_t(cf * (C::bk * kRowBytes) + (r0 + 8 * j) * kRowBytes) + swz;
              const uint16_t* src = p.ga0 + rowoff[j] + foff;
              cp_async16_l2_256(dst, src, nb);
              if (kPlanes == 2) cp_async16_l2_256(dst + C::a_bytes, src + h1d, nb);
            }
          }
          cp_async_mbar_arrive_noinc((C::two_cta && cta_rank != 0) ? &gfull_bar[stage] : &full_bar[stage]);
          if (++stage == C::stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    // warp w may touch TMEM lanes [32*(w%4), +32); the two warps of a lane quarter split the columns.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    constexpr int kColsPerWarp = C::block_n / 2;
    int acc = 0; uint32_t acc_phase = 0;
    for (int u = unit0; u < total_units; u += unit_stride) {
      const UnitInfo ui = decode_unit(p, u, tiles_mn);
      const int split = ui.split, t = ui.t;
      const int m0 = ((t / p.tiles_n) * C::cluster + cta_rank) * kBlockM;
      const int n0 = (t % p.tiles_n) * C::block_n + half * kColsPerWarp;
      const int kb0 = ui.kb0, kb1 = ui.kb1;
      // FWD tail split: this thread's slice of the partial tile in the workspace is float4 j of thread et at [j * 256 + et]
      // (the pointers are formed where they are used: the promoting epilogue has no registers to spare)
#define VV_TAIL_ET (int(threadIdx.x) - 128)
#define VV_TAIL_PART (reinterpret_cast<float4*>(p.tail_ws) + (size_t(ui.tail_idx) * C::cluster + cta_rank) * (kBlockM * C::block_n / 4) + VV_TAIL_ET)
#define VV_TAIL_FLAG (p.tail_flags + ui.tail_idx * C::cluster + cta_rank)
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.d_rows;
      float* drow = p.D + (long long)split * p.slab_stride + (long long)row * p.ldd;
      if (C::promote) {
        // fp32 register accumulation of the chunk partial sums (round-to-nearest adds): the tensor
        // core's own accumulator truncates, which biases long tf32x3 chains ("tf32x3").
        float accr[kColsPerWarp];
#pragma unroll
        for (int j = 0; j < kColsPerWarp; ++j) accr[j] = 0.f;
        for (int c0 = kb0; c0 < kb1; c0 += p.chunk_kb) {
          mbar_wait(&tmem_full[acc], acc_phase);
          tc_fence_after();
          const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * C::block_n + half * kColsPerWarp);
#pragma unroll
          for (int c = 0; c < kColsPerWarp / 16; ++c) {
            uint32_t r[16];
            tmem_ld_32x16(taddr + c * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) accr[c * 16 + j] += __uint_as_float(r[j]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (C::two_cta && cta_rank != 0) mbar_arrive_cluster(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]); }
          acc ^= 1; if (acc == 0) acc_phase ^= 1u;
        }
        if (C::fwd_epi && ui.half == 1) {
          // first K half of a tail unit: the raw partial sums go to the workspace, the second half finishes the tile
#pragma unroll
          for (int j = 0; j < kColsPerWarp / 4; ++j)
            VV_TAIL_PART[j * 256] = make_float4(accr[4 * j], accr[4 * j + 1], accr[4 * j + 2], accr[4 * j + 3]);
          __threadfence();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (VV_TAIL_ET == 0) st_release_gpu(VV_TAIL_FLAG, p.tail_epoch);
          continue;
        }
        if (C::fwd_epi && ui.half == 2) {
          if (VV_TAIL_ET == 0) { while (ld_acquire_gpu(VV_TAIL_FLAG) < p.tail_epoch) __nanosleep(32); }
          asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
          for (int j = 0; j < kColsPerWarp / 4; ++j) {               // first half + second half, in that order
            const float4 pv = __ldcg(VV_TAIL_PART + j * 256);
            accr[4 * j] = pv.x + accr[4 * j]; accr[4 * j + 1] = pv.y + accr[4 * j + 1];
            accr[4 * j + 2] = pv.z + accr[4 * j + 2]; accr[4 * j + 3] = pv.w + accr[4 * j + 3];
          }
        }
        // f16x3: undo the operands' power-of-two scales (exact)
        const float unscale = C::split16 ? __ldg(p.inv_sa) * __ldg(p.inv_sb) : 1.f;
        // dropout keep bits of this thread's 128 columns, 4 bits per column group, produced by ROLLED loops (the
        // generator must not be unrolled 32 times into the epilogue below: instruction cache)
        uint32_t kbits[kColsPerWarp / 32];
#pragma unroll
        for (int q = 0; q < kColsPerWarp / 32; ++q) kbits[q] = 0u;
        if (C::fwd_epi && row_ok && epilogue_has_dropout(p.epi)) {
#pragma unroll
          for (int q = 0; q < kColsPerWarp / 32; ++q) {
            uint32_t bits = 0u;
#pragma unroll 1
            for (int jj = 0; jj < 8; ++jj) {
              const int col = n0 + (q * 8 + jj) * 4;
              if (col < p.d_cols) bits |= dropout_keep_bits(p.epi, p.act_N, row, col) << (4 * jj);
            }
            kbits[q] = bits;
          }
        }
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < kColsPerWarp / 4; ++j) {
            const int col = n0 + j * 4;
            if (col < p.d_cols) {
              float4 v = make_float4(accr[4 * j], accr[4 * j + 1], accr[4 * j + 2], accr[4 * j + 3]);
              if (C::split16) { v.x *= unscale; v.y *= unscale; v.z *= unscale; v.w *= unscale; }
              if (C::fwd_epi) {
                float4 z;
                epilogue_act4(p.epi, p.act_N, row, col, v, z, (kbits[j >> 3] >> (4 * (j & 7))) & 15u);
                if (p.epi.Z) *reinterpret_cast<float4*>(p.epi.Z + (long long)row * p.act_N + col) = z;
              } else {
                const float s = p.epi.out_scale;
                v.x *= s; v.y *= s; v.z *= s; v.w *= s;
              }
              if (C::trans_out) {
                // D^T: element (row, col) lives at D[col * ldd + row]; a warp's 32 rows are 128 contiguous bytes
                float* dt = p.D + (long long)split * p.slab_stride + row;
                dt[(long long)col * p.ldd] = v.x;
                if (col + 1 < p.d_cols) dt[(long long)(col + 1) * p.ldd] = v.y;
                if (col + 2 < p.d_cols) dt[(long long)(col + 2) * p.ldd] = v.z;
                if (col + 3 < p.d_cols) dt[(long long)(col + 3) * p.ldd] = v.w;
              } else {
                *reinterpret_cast<float4*>(drow + col) = v;
              }
            }
          }
        }
        if (!C::fwd_epi && p.fin.tickets) wgrad_arrive_unit<C>(p, m0, (t % p.tiles_n) * C::block_n, int(threadIdx.x) - 128);
      } else {
        if (C::fwd_epi && ui.half == 2) {                              // second K half of a tail unit: the first half's partial
          if (VV_TAIL_ET == 0) { while (ld_acquire_gpu(VV_TAIL_FLAG) < p.tail_epoch) __nanosleep(32); }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(acc * C::block_n + half * kColsPerWarp);
#pragma unroll 1
        for (int c = 0; c < kColsPerWarp / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          const int col0 = n0 + c * 32;
          if (C::fwd_epi && ui.half == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              VV_TAIL_PART[(c * 8 + j) * 256] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                         __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            continue;
          }
          if (C::fwd_epi && ui.half == 2) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 pv = __ldcg(VV_TAIL_PART + (c * 8 + j) * 256);
              r[4 * j] = __float_as_uint(pv.x + __uint_as_float(r[4 * j])); r[4 * j + 1] = __float_as_uint(pv.y + __uint_as_float(r[4 * j + 1]));
              r[4 * j + 2] = __float_as_uint(pv.z + __uint_as_float(r[4 * j + 2])); r[4 * j + 3] = __float_as_uint(pv.w + __uint_as_float(r[4 * j + 3]));
            }
          }
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int col = col0 + j * 4;
              if (col < p.d_cols) {
                float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                       __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                if (C::fwd_epi) {
                  float4 z;
                  const uint32_t kb = epilogue_has_dropout(p.epi) ? dropout_keep_bits(p.epi, p.act_N, row, col) : 0u;
                  epilogue_act4(p.epi, p.act_N, row, col, v, z, kb);
                  if (p.epi.Z) *reinterpret_cast<float4*>(p.epi.Z + (long long)row * p.act_N + col) = z;
                } else {
                  const float s = p.epi.out_scale;
                  v.x *= s; v.y *= s; v.z *= s; v.w *= s;
                }
                if (C::trans_out) {
                  float* dt = p.D + (long long)split * p.slab_stride + row;
                  dt[(long long)col * p.ldd] = v.x;
                  if (col + 1 < p.d_cols) dt[(long long)(col + 1) * p.ldd] = v.y;
                  if (col + 2 < p.d_cols) dt[(long long)(col + 2) * p.ldd] = v.z;
                  if (col + 3 < p.d_cols) dt[(long long)(col + 3) * p.ldd] = v.w;
                } else
                *reinterpret_cast<float4*>(drow + col) = v;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (C::two_cta && cta_rank != 0) mbar_arrive_cluster(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]); }
        acc ^= 1; if (acc == 0) acc_phase ^= 1u;
        if (C::fwd_epi && ui.half == 1) {
          __threadfence();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (VV_TAIL_ET == 0) st_release_gpu(VV_TAIL_FLAG, p.tail_epoch);
        }
        if (!C::fwd_epi && p.fin.tickets) wgrad_arrive_unit<C>(p, m0, (t % p.tiles_n) * C::block_n, int(threadIdx.x) - 128);
      }
    }
    if (!C::fwd_epi && p.fin.tickets) wgrad_drain<C>(p, int(threadIdx.x) - 128);
#undef VV_TAIL_ET
#undef VV_TAIL_PART
#undef VV_TAIL_FLAG
  }

  tc_fence_before();
  if (C::cluster > 1) cluster_sync_all(); else __syncthreads();   // the peer may still arrive on / multicast into this CTA
  if (warp == 2) { tc_fence_after(); if (C::two_cta) tmem_dealloc2(tmem_base, C::tmem_cols); else tmem_dealloc(tmem_base, C::tmem_cols); }
#endif
}

// ----------------------------------------------------------------------------
// host side: tensor maps + dispatch
// ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || !p) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2D row-major tensor [outer rows, inner contiguous elems]; box = [box_outer rows, box_bytes of inner].
int make_tmap(CUtensorMap* tm, const void* base, CUtensorMapDataType dt, int elem_bytes, uint64_t inner,
              uint64_t outer, uint32_t box_outer, CUtensorMapSwizzle swizzle, int box_bytes = kRowBytes,
              uint64_t pitch_elems = 0 /* 0 = dense rows */) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return VV_ERR_CUDA;
  const uint64_t pitch = pitch_elems ? pitch_elems : inner;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((pitch * elem_bytes) & 15) != 0) {
    set_error("GEMM operand must be 16-byte aligned with a row pitch that is a multiple of 16 bytes");
    return VV_ERR_INVALID;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch * elem_bytes};
  cuuint32_t box[2] = {uint32_t(box_bytes / elem_bytes), box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", int(r)); return VV_ERR_CUDA; }
  return VV_OK;
}

template <class C>
int launch_cfg(const GemmProblem& g, cudaStream_t stream) {
  // output extent, reduction length, operand shapes per kind
  int d_rows, d_cols, red;
  uint64_t a_inner, a_outer, b_inner, b_outer;
  switch (g.kind) {
    case GEMM_FWD:     d_rows = g.M; d_cols = g.N; red = g.K; a_inner = g.K; a_outer = g.M; b_inner = g.K; b_outer = g.N; break;
    case GEMM_WGRAD:   d_rows = g.N; d_cols = g.K; red = g.M; a_inner = g.N; a_outer = g.M; b_inner = g.K; b_outer = g.M; break;
    case GEMM_WGRAD_T: d_rows = g.K; d_cols = g.N; red = g.M; a_inner = g.K; a_outer = g.M; b_inner = g.N; b_outer = g.M; break;
    default:           d_rows = g.M; d_cols = g.K; red = g.N; a_inner = g.N; a_outer = g.M; b_inner = g.K; b_outer = g.N; break;
  }
  if (C::trans_out != (g.kind == GEMM_WGRAD_T)) { set_error("internal: transposed-output configuration mismatch"); return VV_ERR_INVALID; }
  // WGRAD_T may be asked for a slice [n0, n0 + ncols) of the outputs (the data-parallel trainer pipelines the
  // all-reduce of one slice under the GEMM of the next): a column window of dZ and a row window of dW
  const int n0 = (g.kind == GEMM_WGRAD_T) ? g.n0 : 0;
  const int ncols = (g.kind == GEMM_WGRAD_T && g.ncols > 0) ? g.ncols : int(g.N - n0);
  uint64_t b_pitch = 0;
  const char* b_hi = static_cast<const char*>(g.B.hi);
  const char* b_lo = static_cast<const char*>(g.B.lo);
  if (g.kind == GEMM_WGRAD_T) {
    if (n0 < 0 || ncols <= 0 || n0 + ncols > g.N || (n0 % 8) != 0) { set_error("wgrad slice [%d, +%d) of N=%d is invalid", n0, ncols, g.N); return VV_ERR_INVALID; }
    b_pitch = uint64_t(g.N); b_inner = uint64_t(ncols); d_cols = ncols;
    b_hi += size_t(n0) * C::elem_bytes;
    if (b_lo) b_lo += size_t(n0) * C::elem_bytes;
  }
  const CUtensorMapDataType dt = C::tf32 ? (C::nprod == 3 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32)
                                         : (C::f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUtensorMap tA_hi, tA_hb, tA_lb, tB_hi, tB_hb, tB_lb;
  // gathered operand A: no tensor map, the producer warps read the bank's operand planes through rowmap
  if (C::gather) {
    if (!g.rowmap || g.bank_rows <= 0) { set_error("gather variant without a rowmap"); return VV_ERR_INVALID; }
    if ((reinterpret_cast<uintptr_t>(g.A.hi) & 15) != 0 || (g.K % 8) != 0) { set_error("gathered operand must be 16-byte aligned with K % 8 == 0"); return VV_ERR_INVALID; }
  }
  const uint32_t a_box = C::a_mn ? C::bk : kBlockM;
  // K-major B: with 2-CTA clusters each CTA fetches (and multicasts) half of the tile's rows
  const uint32_t b_box = C::b_mn ? C::bk : C::block_n / C::cluster;
  const CUtensorMapSwizzle sw_a = (C::tf32 && C::a_mn) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  const CUtensorMapSwizzle sw_b = (C::tf32 && C::b_mn) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  int rc;
  if ((rc = make_tmap(&tB_hi, b_hi, dt, C::elem_bytes, b_inner, b_outer, b_box, sw_b, kRowBytes, b_pitch))) return rc;
  if (C::gather) tA_hi = tB_hi;
  else if ((rc = make_tmap(&tA_hi, g.A.hi, dt, C::elem_bytes, a_inner, a_outer, a_box, sw_a))) return rc;
  if (C::mixed) {
    // lo = two bf16 planes [bf16(x) | bf16(x - hi)], each the shape of the operand
    if (!g.A.lo || !g.B.lo) { set_error("TF32X3 needs the hi array and the two-plane bf16 lo array of every operand"); return VV_ERR_INVALID; }
    const uint16_t* a_planes = static_cast<const uint16_t*>(g.A.lo);
    const uint16_t* b_planes = static_cast<const uint16_t*>(g.B.lo);
    const uint64_t a_count = a_inner * a_outer, b_count = b_inner * b_outer;
    const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    // K-major: [rows x 64 B] tiles, SWIZZLE_64B; MN-major: [32 k-rows x 128 B] chunks, SWIZZLE_128B
    const CUtensorMapSwizzle s16a = C::a_mn ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const CUtensorMapSwizzle s16b = C::b_mn ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const int bba = C::a_mn ? 128 : 64, bbb = C::b_mn ? 128 : 64;
    if ((rc = make_tmap(&tA_hb, a_planes, bf, 2, a_inner, a_outer, a_box, s16a, bba))) return rc;
    if ((rc = make_tmap(&tA_lb, a_planes + a_count, bf, 2, a_inner, a_outer, a_box, s16a, bba))) return rc;
    if ((rc = make_tmap(&tB_hb, b_planes, bf, 2, b_inner, b_outer, b_box, s16b, bbb))) return rc;
    if ((rc = make_tmap(&tB_lb, b_planes + b_count, bf, 2, b_inner, b_outer, b_box, s16b, bbb))) return rc;
  } else if (C::split16) {
    if (!g.A.lo || !g.B.lo) { set_error("F16X3 needs the h0 and h1 planes of every operand"); return VV_ERR_INVALID; }
    if ((rc = make_tmap(&tB_hb, b_lo, dt, C::elem_bytes, b_inner, b_outer, b_box, sw_b, kRowBytes, b_pitch))) return rc;
    if (C::gather) tA_hb = tB_hb;
    else if ((rc = make_tmap(&tA_hb, g.A.lo, dt, C::elem_bytes, a_inner, a_outer, a_box, sw_a))) return rc;
    tA_lb = tA_hi; tB_lb = tB_hi;
  } else {
    tA_hb = tA_hi; tA_lb = tA_hi; tB_hb = tB_hi; tB_lb = tB_hi;
  }
  TcParams p;
  p.d_rows = d_rows; p.d_cols = d_cols; p.ldd = C::trans_out ? g.K : d_cols;
  p.tiles_m = (d_rows + kBlockM - 1) / kBlockM;
  p.tiles_n = (d_cols + C::block_n - 1) / C::block_n;
  p.num_kb = (red + C::bk - 1) / C::bk;
  const int nsplit = g.nsplit < 1 ? 1 : g.nsplit;
  p.kb_per_split = (p.num_kb + nsplit - 1) / nsplit;
  // every slab must receive at least one k-block, else it would stay unwritten
  if (nsplit > p.num_kb || (nsplit - 1) * p.kb_per_split >= p.num_kb) {
    set_error("nsplit=%d leaves an empty split for %d k-blocks", nsplit, p.num_kb);
    return VV_ERR_INVALID;
  }
  p.nsplit = nsplit;
  // tf32x3 / f16x3: promote every 512 reduction elements (16 k-blocks of 32 / 8 of 64) -> truncation bias < 4e-6 relative
  p.chunk_kb = C::promote ? 512 / C::bk : p.kb_per_split;
  p.inv_sa = C::split16 ? &f16_hdr(g.A.hi)->inv_scale : nullptr;
  p.inv_sb = C::split16 ? &f16_hdr(g.B.hi)->inv_scale : nullptr;
  p.D = g.D + (C::trans_out ? (long long)n0 * g.K : 0); p.slab_stride = g.slab_stride;
  p.act_N = g.N;
  p.wait = g.wait;
  // FWD tail split (see TcParams): only with a workspace, one K range per unit, and when the last wave is at most half full
  p.tail_first = -1; p.tail_count = 0; p.tail_kb = 0; p.tail_ws = nullptr; p.tail_flags = nullptr; p.tail_epoch = 0;
  if (g.kind == GEMM_FWD && g.tail_ws && nsplit == 1) {
    const int units = ((p.tiles_m + C::cluster - 1) / C::cluster) * p.tiles_n;
    const int ncl = num_sms() / C::cluster;
    const int rem = units % ncl;
    // whole promotion chunks in the first half (the non-promoting kernels accumulate a unit's whole K range in TMEM)
    const int half_kb = C::promote ? ((p.num_kb / 2 + p.chunk_kb - 1) / p.chunk_kb) * p.chunk_kb : p.num_kb / 2;
    const size_t need = size_t(rem) * C::cluster * kBlockM * C::block_n * sizeof(float);
    if (units > ncl && rem > 0 && 2 * rem <= ncl && half_kb > 0 && half_kb < p.num_kb && need <= g.tail_ws_bytes &&
        size_t(rem) * C::cluster <= g.tail_flags_count) {
      p.tail_first = units - rem; p.tail_count = rem; p.tail_kb = half_kb;
      p.tail_ws = g.tail_ws; p.tail_flags = g.tail_flags; p.tail_epoch = g.tail_epoch;
    }
  }
  p.fin = WgradFinish();
  if (g.finish && g.finish->tickets) {
    if (g.kind != GEMM_WGRAD && g.kind != GEMM_WGRAD_T) { set_error("the split-K finish belongs to the weight-gradient kernels"); return VV_ERR_INVALID; }
    if (n0 != 0 || ncols != g.N) { set_error("the split-K finish needs the whole output (no column slice)"); return VV_ERR_INVALID; }
    if ((g.K % 4) != 0) { set_error("the split-K finish needs K %% 4 == 0"); return VV_ERR_INVALID; }
    p.fin = *g.finish;
  }
  p.epi = g.epi;
  p.rowmap = g.rowmap;
  p.ga0 = static_cast<const uint16_t*>(g.A.hi); p.ga1 = static_cast<const uint16_t*>(g.A.lo);
  p.ga_pitch = g.K; p.ga_cols = g.K; p.rowmap_len = ((g.M + 127) / 128) * 128;
  static bool attr_set = false;
  if (!attr_set) {
    VV_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::smem_bytes));
    attr_set = true;
  }
  const int total = ((p.tiles_m + C::cluster - 1) / C::cluster) * p.tiles_n * p.nsplit + p.tail_count;     // (virtual) units per cluster
  int nclusters = num_sms() / C::cluster;
  if (total < nclusters) nclusters = total;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(nclusters * C::cluster); lc.blockDim = dim3(kNumThreads);
  lc.dynamicSmemBytes = C::smem_bytes; lc.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C::cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  lc.attrs = attr; lc.numAttrs = 1;
  VV_CUDA(cudaLaunchKernelEx(&lc, gemm_tc_kernel<C>, tA_hi, tA_hb, tA_lb, tB_hi, tB_hb, tB_lb, p));
  VV_LAUNCH_CHECK();
  count_launch();
  return VV_OK;
}

}  // namespace

bool gemm_tc_supported(const GemmProblem& g, const char** why) {
  static const char* w_prec = "precision is not a tensor-core mode";
  static const char* w_align = "tensor-core path needs N % 8 == 0 and K % 8 == 0";
  if (g.prec != VV_PREC_TF32X3 && g.prec != VV_PREC_TF32 && g.prec != VV_PREC_BF16 && g.prec != VV_PREC_F16X3) { if (why) *why = w_prec; return false; }
  if ((g.N % 8) != 0 || (g.K % 8) != 0) { if (why) *why = w_align; return false; }
  return true;
}

int gemm_tc_launch(const GemmProblem& g, cudaStream_t stream) {
  const char* why = nullptr;
  if (!gemm_tc_supported(g, &why)) { set_error("%s (M=%d N=%d K=%d)", why, g.M, g.N, g.K); return VV_ERR_UNSUPPORTED; }
  //                 tf32   A-MN   B-MN  nprod  BN  stages fwd-epi gather f16  transposed-out
  const bool gat = g.rowmap != nullptr;
  if (gat != (g.kind == GEMM_WGRAD_T) && g.kind != GEMM_FWD) { set_error("only FWD and WGRAD_T take a gathered operand"); return VV_ERR_INVALID; }
  if (gat && g.prec != VV_PREC_BF16 && g.prec != VV_PREC_F16X3) {
    set_error("the gather-fused variants are built for the 2-byte operand formats (bf16, f16x3)"); return VV_ERR_UNSUPPORTED;
  }
  // cta_group::2 forward (gathered, 2-byte operands). Measured (profiles/r02_2cta_forward.md): f16x3 forward 0.622 ms vs
  // 0.641 with the 1-CTA multicast kernel, bf16 0.267 vs 0.246 -- so it is the default for f16x3 only.
  // VV_GEMM_2CTA=0 turns it off, =1 also takes it for bf16.
  static const int two_cta = [] { const char* e = getenv("VV_GEMM_2CTA"); return e ? (atoi(e) != 0 ? 2 : 0) : 1; }();
  if (two_cta && gat && g.kind == GEMM_FWD && (g.N % 256) == 0) {
    if (g.prec == VV_PREC_BF16 && two_cta == 2) return launch_cfg<Cfg<false, false, false, 1, 256, 6, true, true, false, false, true>>(g, stream);
    if (g.prec == VV_PREC_F16X3)                return launch_cfg<Cfg<false, false, false, 3, 256, 3, true, true, true,  false, true>>(g, stream);
  }
  // cta_group::2 weight gradient (gathered, transposed).  Measured (profiles/r02_2cta_forward.md): f16x3 0.587-0.593 ms vs
  // 0.610-0.620, bf16 0.308 vs 0.299 -- the default for f16x3 only, as for the forward.  VV_GEMM_2CTA_WGRAD=0 off, =1 bf16 too.
  static const int two_cta_w = [] { const char* e = getenv("VV_GEMM_2CTA_WGRAD"); return e ? (atoi(e) != 0 ? 2 : 0) : 1; }();
  if (two_cta_w && gat && g.kind == GEMM_WGRAD_T && (g.N % 256) == 0 && g.n0 == 0 && (g.ncols <= 0 || g.ncols == g.N)) {
    if (g.prec == VV_PREC_BF16 && two_cta_w == 2) return launch_cfg<Cfg<false, true, true, 1, 256, 6, false, true, false, true, true>>(g, stream);
    if (g.prec == VV_PREC_F16X3)                  return launch_cfg<Cfg<false, true, true, 3, 256, 3, false, true, true,  true, true>>(g, stream);
  }
  if (g.prec == VV_PREC_BF16) {
    switch (g.kind) {
      case GEMM_FWD:     return gat ? launch_cfg<Cfg<false, false, false, 1, 256, 4, true,  true>>(g, stream)
                                    : launch_cfg<Cfg<false, false, false, 1, 256, 4, true >>(g, stream);
      case GEMM_WGRAD:   return launch_cfg<Cfg<false, true,  true,  1, 256, 4, false>>(g, stream);
      case GEMM_WGRAD_T: return launch_cfg<Cfg<false, true,  true,  1, 256, 4, false, true, false, true>>(g, stream);
      default:           return launch_cfg<Cfg<false, false, true,  1, 256, 4, false>>(g, stream);
    }
  } else if (g.prec == VV_PREC_TF32) {
    switch (g.kind) {
      case GEMM_FWD:   return launch_cfg<Cfg<true, false, false, 1, 256, 4, true >>(g, stream);
      case GEMM_WGRAD: return launch_cfg<Cfg<true, true,  true,  1, 256, 4, false>>(g, stream);
      default:         return launch_cfg<Cfg<true, false, true,  1, 256, 4, false>>(g, stream);
    }
  } else if (g.prec == VV_PREC_F16X3) {
    switch (g.kind) {
      case GEMM_FWD:     return gat ? launch_cfg<Cfg<false, false, false, 3, 256, 2, true,  true,  true>>(g, stream)
                                    : launch_cfg<Cfg<false, false, false, 3, 256, 2, true,  false, true>>(g, stream);
      case GEMM_WGRAD:   return launch_cfg<Cfg<false, true,  true,  3, 256, 2, false, false, true>>(g, stream);
      case GEMM_WGRAD_T: return launch_cfg<Cfg<false, true,  true,  3, 256, 2, false, true,  true, true>>(g, stream);
      default:           return launch_cfg<Cfg<false, false, true,  3, 256, 2, false, false, true>>(g, stream);
    }
  } else {
    switch (g.kind) {
      case GEMM_FWD:   return launch_cfg<Cfg<true, false, false, 3, 256, 2, true >>(g, stream);
      case GEMM_WGRAD: return launch_cfg<Cfg<true, true,  true,  3, 256, 2, false>>(g, stream);
      default:         return launch_cfg<Cfg<true, false, true,  3, 256, 2, false>>(g, stream);
    }
  }
}

}  // namespace vv
