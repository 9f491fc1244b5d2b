// vv_trainer.cu -- one data-parallel rank of the fused training step (host C++).
//
// The loop body of Solver::Solve for the shipped net (ref: solver.cpp:177-220:
// ForwardBackward, ComputeUpdateValue, Net::Update) as a fixed kernel sequence:
//   K0 gather -> K1 fc7 fwd (+ReLU+dropout epilogue) -> K2 rank-loss fwd ->
//   K3 rank-loss bwd (+ReLU'/dropout', db) -> K1 wgrad (split-K slabs) ->
//   [K1 dgrad] -> K4 fused SGD update
// Data parallel (one process per GPU): the gradient exchange is part of K4 -- reduce-scatter by push, owner update,
// all-gather by push over CUDA-IPC-mapped peer memory (vv_dp_exchange.cuh), consumed by the next forward GEMM's
// producer.  NCCL (dlopen'ed: the copy already loaded by the host process, e.g. torch's, is reused, so the library
// has no link-time dependency) bootstraps the IPC handles, serves forward/backward-only steps and vv_dp_gather_state,
// and remains the whole exchange when peer mapping is unavailable or VV_DP_MODE=nccl.
#include <dlfcn.h>
#include <stdlib.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "../vv_gemm.cuh"

using namespace vv;

namespace {

// ---- minimal NCCL binding --------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { set_error("cannot dlopen libnccl.so.2: %s", dlerror()); return false; }
    GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
    AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
    CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
    GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
    GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce || !AllGather || !CommDestroy) { set_error("libnccl is missing symbols"); return false; }
    return true;
  }
};
Nccl g_nccl;
constexpr int kNcclFloat = 7, kNcclSum = 0, kNcclInt8 = 0;

// ---- peer-memory exchange state (vv_dp_exchange.cuh) ------------------------
struct DpP2P {
  bool on = false;
  void* region = nullptr; size_t region_bytes = 0;       // recv_dw | recv_small | flags | wlast (this rank's)
  size_t off_small = 0, off_flags = 0, off_wlast = 0;
  int small_stride = 0, rows_per = 0;
  void* opened[3 * kDpMaxRanks] = {};                      // mappings to close (region, W operand block, W master per peer)
  int n_opened = 0;
  DpPeers peers;
  unsigned int seq = 0, waited = 0;                        // exchanged steps; the last one some consumer already waited for
  unsigned int* err_host = nullptr; unsigned int* err_dev = nullptr;   // host-mapped time-out word
  int replicate_master = 0;
  bool in_kernel_wait = true;
  std::string why_off;
};

struct DevBuf {
  void* p = nullptr; size_t bytes = 0;
  void* base = nullptr;                // the allocation (p may point inside it); NULL = p is not owned
  int alloc(size_t n) {
    bytes = n;
    if (n == 0) return VV_OK;
    VV_CUDA(cudaMalloc(&base, n));
    VV_CUDA(cudaMemset(base, 0, n));   // zero-fill on first touch like SyncedMemory (syncedmem.cpp:24-25)
    p = base;
    return VV_OK;
  }
  void release() { if (base) cudaFree(base); base = nullptr; p = nullptr; }
  template <class T> T* as() const { return static_cast<T*>(p); }
};
// operand copies of `count` elements in the layout of `prec` (vv_operand_bytes): hi owns the block, lo points into it
int alloc_operand(DevBuf& hi, DevBuf& lo, size_t count, int prec) {
  size_t ho = 0, lo_off = 0;
  const size_t bytes = vv_operand_bytes(int64_t(count), prec, &ho, &lo_off);
  if (bytes == 0) return VV_OK;
  hi.release(); lo.release();
  int rc = hi.alloc(bytes);
  if (rc) return rc;
  hi.p = static_cast<char*>(hi.base) + ho;
  if (lo_off) { lo.p = static_cast<char*>(hi.base) + lo_off; lo.base = nullptr; }
  return VV_OK;
}

}  // namespace

struct vv_trainer {
  vv_trainer_cfg_t cfg;
  cudaStream_t stream = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_grad = nullptr, ev_comm = nullptr, ev_slice[4] = {nullptr, nullptr, nullptr, nullptr};
  int R = 0, M = 0, nsplit = 1;
  vv_rank_cfg_t rank;
  // parameters
  DevBuf W, b, Wh, bh, W_hi, W_lo;
  // activations / gradients
  DevBuf Xf, X_hi, X_lo, Zf, H, stats, item_loss, item_viol, dZf, dZ_hi, dZ_lo, dW_parts, dbx, dX;
  // gather-fused path: operand copies of the registered bank, per-step gather plan, quirk corrections
  DevBuf bank_hi, bank_lo, rowmap, delta, wlast, dq, tickets, rank_ws, tail_ws, tail_flags;
  unsigned int tail_epoch = 0;
  const float* bank_reg = nullptr; int64_t bank_reg_rows = 0;
  // F16X3: the X scale is fixed from max|bank| (gathered rows are a subset), the dZ scale trails the previous step
  const float* scaled_bank = nullptr; int64_t scaled_bank_rows = 0; bool dz_scale_ready = false;
  bool x_allocated = false;
  ncclComm_t comm = nullptr;
  DpP2P p2p;
  unsigned int finish_epoch = 0;       // finishing wgrad launches so far (the tickets only grow)
  unsigned int* bad_idx_host = nullptr; unsigned int* bad_idx_dev = nullptr;   // host-mapped: a gather plan saw an index outside the bank
  int last_launches = 0;
  // optional per-phase timing
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool;      // [step][phase][2]
  int timed_steps = 0;
  cudaEvent_t* phase_events(int phase) {
    const size_t base = (size_t(timed_steps) * VV_NUM_PHASES + phase) * 2;
    while (ev_pool.size() < base + 2) { cudaEvent_t e; cudaEventCreate(&e); ev_pool.push_back(e); }
    return &ev_pool[base];
  }
  void tic(int phase) { if (timing) cudaEventRecord(phase_events(phase)[0], stream); }
  void toc(int phase) { if (timing) cudaEventRecord(phase_events(phase)[1], stream); }

  vv_operand_t opX() const { return op(Xf, X_hi, X_lo); }
  vv_operand_t opW() const { return op(W, W_hi, W_lo); }
  vv_operand_t opdZ() const { return op(dZf, dZ_hi, dZ_lo); }
  vv_operand_t op(const DevBuf& f, const DevBuf& hi, const DevBuf& lo) const {
    vv_operand_t o;
    if (cfg.prec == VV_PREC_TF32X3 || cfg.prec == VV_PREC_F16X3) { o.hi = hi.p; o.lo = lo.p; }
    else if (cfg.prec == VV_PREC_BF16) { o.hi = hi.p; o.lo = nullptr; }
    else { o.hi = f.p; o.lo = nullptr; }
    return o;
  }
  bool needs_f32_operand() const { return cfg.prec == VV_PREC_FP32_SIMT || cfg.prec == VV_PREC_TF32; }

  int init() {
    R = cfg.C + cfg.Nn; M = R * cfg.B;
    const size_t MK = size_t(M) * cfg.K, MN = size_t(M) * cfg.N, NK = size_t(cfg.N) * cfg.K;
    const bool f32op = needs_f32_operand();
    int rc;
#define A(buf, n) if ((rc = buf.alloc(n))) return rc
    A(W, NK * 4); A(b, size_t(cfg.N) * 4); A(Wh, NK * 4); A(bh, size_t(cfg.N) * 4);
    // the materialised X operand (1-2 GB at B=4096) is allocated on first use: the gather-fused path never needs it
    if ((rc = alloc_operand(W_hi, W_lo, NK, cfg.prec))) return rc;
    if ((rc = alloc_operand(dZ_hi, dZ_lo, MN, cfg.prec))) return rc;
    if (f32op || cfg.keep_blobs) { A(dZf, MN * 4); }
    A(wlast, size_t(cfg.N) * 4);
    // forward tail split: at most (clusters / 2) units x 2 CTAs x one 128 x 256 fp32 partial tile, + one flag per tail CTA
    A(tail_ws, size_t(kNumSMsB200 / 2) * 128 * 256 * 4); A(tail_flags, 1024);
    A(rank_ws, rank_loss_workspace_bytes(cfg.N));       // deterministic db / dq / loss sums of the fused rank-loss kernel
    // split-K finish: one ticket counter per wgrad output tile (+ one for the tiles finished), zeroed here, self-resetting
    A(tickets, (size_t((cfg.K + 127) / 128 + 1) * size_t((cfg.N + 127) / 128 + 1) + 1) * 4);
    A(rowmap, size_t((M + 127) / 128 * 128) * 4); A(delta, size_t((M + 127) / 128 * 128) * 4);
    if (cfg.keep_blobs) { A(Zf, MN * 4); }
    A(H, MN * 4);
    A(stats, size_t(cfg.B) * vv_rank_stats_stride(cfg.Nn) * 4);
    A(item_loss, size_t(cfg.B) * 4); A(item_viol, size_t(cfg.B) * 4);
    nsplit = vv_ip_wgrad_auto_nsplit(M, cfg.N, cfg.K, cfg.prec);
    A(dW_parts, size_t(nsplit) * NK * 4);
    // db [N] + loss + violations (one all-reduce payload) + the fused loss kernel's ticket counter + pad, then dq [N]
    // (the K-1 quirk's column sums): one memset zeroes all accumulators of a step
    A(dbx, size_t(2 * cfg.N + 4) * 4);
    dq.p = dbx.as<float>() + cfg.N + 4; dq.bytes = size_t(cfg.N) * 4; dq.base = nullptr;
    if (cfg.compute_dgrad) { A(dX, MK * 4); }
#undef A
    if (cudaHostAlloc(reinterpret_cast<void**>(&bad_idx_host), sizeof(unsigned int), cudaHostAllocMapped) == cudaSuccess) {
      *bad_idx_host = 0u;
      if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&bad_idx_dev), bad_idx_host, 0) != cudaSuccess) bad_idx_dev = nullptr;
    } else { bad_idx_host = nullptr; cudaGetLastError(); }
    memset(&rank, 0, sizeof(rank));
    rank.B = cfg.B; rank.C = cfg.C; rank.Nn = cfg.Nn; rank.N = cfg.N;
    bool any = false;
    for (int i = 0; i < cfg.C - 1 && i < VV_MAX_CONTEXT; ++i) any = any || cfg.coeff[i] != 0.f;
    for (int i = 0; i < cfg.C - 1 && i < VV_MAX_CONTEXT; ++i)
      rank.coeff[i] = any ? cfg.coeff[i] : 1.f / float(cfg.C - 1);
    rank.margin = cfg.margin; rank.norm = cfg.norm; rank.eps = 1e-10f;
    if (cfg.world_size > 1) {
      VV_CUDA(cudaStreamCreateWithFlags(&comm_stream, cudaStreamNonBlocking));
      VV_CUDA(cudaEventCreateWithFlags(&ev_grad, cudaEventDisableTiming));
      VV_CUDA(cudaEventCreateWithFlags(&ev_comm, cudaEventDisableTiming));
      for (int i = 0; i < 4; ++i) VV_CUDA(cudaEventCreateWithFlags(&ev_slice[i], cudaEventDisableTiming));
    }
    return VV_OK;
  }
  int alloc_x() {
    if (x_allocated) return VV_OK;
    const size_t MK = size_t(M) * cfg.K;
    int rc;
    if ((rc = alloc_operand(X_hi, X_lo, MK, cfg.prec))) return rc;
    if (needs_f32_operand() || cfg.keep_blobs) { if ((rc = Xf.alloc(MK * 4))) return rc; }
    x_allocated = true;
    return VV_OK;
  }
  vv_operand_t opBank() const {
    vv_operand_t o;
    if (cfg.prec == VV_PREC_F16X3) { o.hi = bank_hi.p; o.lo = bank_lo.p; }
    else { o.hi = bank_hi.p; o.lo = nullptr; }     // BF16
    return o;
  }
  // Peers write into this rank's memory only during a step's update kernel, and raise w_ready[peer] = seq here after
  // their last write: once those flags are in (and this rank's own kernels have drained, its stores acknowledged by the
  // fences in front of its flags) nothing is in flight in either direction and the mappings can go -- no collective
  // needed, so a rank that dies cannot hang the others' teardown (the wait kernel gives up after VV_DP_TIMEOUT_MS).
  void p2p_quiesce() {
    if (!p2p.on) return;
    const bool healthy = !(p2p.err_host && *p2p.err_host);
    if (healthy && p2p.waited < p2p.seq) {
      dp_wait_w_ready(p2p.peers.flags[cfg.rank], cfg.world_size, p2p.seq, p2p.err_dev, reinterpret_cast<vv_stream_t>(stream));
      p2p.waited = p2p.seq;
    }
    cudaStreamSynchronize(stream);
  }
  ~vv_trainer() {
    p2p_quiesce();
    for (int i = 0; i < p2p.n_opened; ++i) cudaIpcCloseMemHandle(p2p.opened[i]);
    if (p2p.region) { if (wlast.p && !wlast.base) wlast.p = nullptr; cudaFree(p2p.region); }
    if (p2p.err_host) cudaFreeHost(p2p.err_host);
    if (bad_idx_host) cudaFreeHost(bad_idx_host);
    if (comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm);
    DevBuf* all[] = {&W, &b, &Wh, &bh, &W_hi, &W_lo, &Xf, &X_hi, &X_lo, &Zf, &H, &stats, &item_loss, &item_viol,
                     &dZf, &dZ_hi, &dZ_lo, &dW_parts, &dbx, &dX, &bank_hi, &bank_lo, &rowmap, &delta, &wlast, &dq, &tickets, &rank_ws, &tail_ws, &tail_flags};
    for (DevBuf* d : all) d->release();
    for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
    if (ev_grad) cudaEventDestroy(ev_grad);
    if (ev_comm) cudaEventDestroy(ev_comm);
    for (int i = 0; i < 4; ++i) if (ev_slice[i]) cudaEventDestroy(ev_slice[i]);
    if (comm_stream) cudaStreamDestroy(comm_stream);
  }
  int dp_check_error() {
    if (p2p.err_host && *p2p.err_host) {
      set_error("data-parallel exchange timed out waiting for a peer (code %u: 1 = gradient contributions, 2 = updated weights); "
                "a rank died or fell more than VV_DP_TIMEOUT_MS behind", *p2p.err_host);
      return VV_ERR_NCCL;
    }
    return VV_OK;
  }
  int p2p_setup();
  float* loss_ptr() { return dbx.as<float>() + cfg.N; }
  float* viol_ptr() { return dbx.as<float>() + cfg.N + 1; }
};


// Map every rank's exchange region, W operand block and W master into this process (CUDA IPC; the handles travel
// through one ncclAllGather).  Any failure leaves the NCCL all-reduce path in place and records why.
int vv_trainer::p2p_setup() {
  const vv_trainer_cfg_t& c = cfg;
  const int G = c.world_size;
  const char* mode = getenv("VV_DP_MODE");
  if (mode && !strcmp(mode, "nccl")) { p2p.why_off = "VV_DP_MODE=nccl"; return VV_OK; }
  if (G > kDpMaxRanks) { p2p.why_off = "more than 8 ranks"; return VV_OK; }
  if (c.N % G != 0) { p2p.why_off = "N does not divide evenly over the ranks"; return VV_OK; }
  const size_t NK = size_t(c.N) * c.K;
  p2p.rows_per = c.N / G;
  p2p.small_stride = ((c.N + 2 + 31) / 32) * 32;
  size_t off = NK * 4;                                   // recv_dw: G sources x (N/G x K)
  p2p.off_small = off; off += size_t(G) * p2p.small_stride * 4;
  p2p.off_flags = (off + 255) & ~size_t(255); off = p2p.off_flags + kDpFlagWords * 4;
  p2p.off_wlast = (off + 255) & ~size_t(255); off = p2p.off_wlast + size_t(c.N) * 4;
  p2p.region_bytes = off;
  VV_CUDA(cudaMalloc(&p2p.region, p2p.region_bytes));
  VV_CUDA(cudaMemset(p2p.region, 0, p2p.region_bytes));
  VV_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&p2p.err_host), sizeof(unsigned int), cudaHostAllocMapped));
  *p2p.err_host = 0u;
  VV_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&p2p.err_dev), p2p.err_host, 0));
  // wlast moves into the region (the owners write it remotely); keep what sync_weights put there
  char* reg = static_cast<char*>(p2p.region);
  VV_CUDA(cudaMemcpy(reg + p2p.off_wlast, wlast.p, size_t(c.N) * 4, cudaMemcpyDeviceToDevice));
  // handles: region, W operand block (absent for the fp32-operand precisions), W master
  const bool has_op = W_hi.base != nullptr;
  struct Pack { cudaIpcMemHandle_t region, wblk, wm; int has_op; int pad[15]; };
  static_assert(sizeof(Pack) == 3 * 64 + 64, "handle pack layout");
  Pack mine; memset(&mine, 0, sizeof(mine));
  cudaError_t e = cudaIpcGetMemHandle(&mine.region, p2p.region);
  if (e == cudaSuccess && has_op) e = cudaIpcGetMemHandle(&mine.wblk, W_hi.base);
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&mine.wm, W.base);
  mine.has_op = has_op ? 1 : 0;
  int ok = (e == cudaSuccess) ? 1 : 0;
  if (!ok) { p2p.why_off = std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e); cudaGetLastError(); }
  // all-gather the packs (a failed rank still takes part so that nobody hangs)
  Pack* dsend = nullptr; Pack* drecv = nullptr;
  VV_CUDA(cudaMalloc(&dsend, sizeof(Pack))); VV_CUDA(cudaMalloc(&drecv, sizeof(Pack) * G));
  if (!ok) mine.has_op = -1;
  VV_CUDA(cudaMemcpy(dsend, &mine, sizeof(Pack), cudaMemcpyHostToDevice));
  ncclResult_t r = g_nccl.AllGather(dsend, drecv, sizeof(Pack), kNcclInt8, comm, stream);
  if (r != 0) { set_error("ncclAllGather of the IPC handles failed (%d)", r); return VV_ERR_NCCL; }
  VV_CUDA(cudaStreamSynchronize(stream));
  std::vector<Pack> all(G);
  VV_CUDA(cudaMemcpy(all.data(), drecv, sizeof(Pack) * G, cudaMemcpyDeviceToHost));
  cudaFree(dsend); cudaFree(drecv);
  for (int g = 0; g < G; ++g) if (all[g].has_op < 0) { ok = 0; if (p2p.why_off.empty()) p2p.why_off = "a peer could not export its memory"; }
  const size_t ho = has_op ? size_t(static_cast<char*>(W_hi.p) - static_cast<char*>(W_hi.base)) : 0;
  const size_t lo_off = (has_op && W_lo.p) ? size_t(static_cast<char*>(W_lo.p) - static_cast<char*>(W_hi.base)) : 0;
  int dev = 0; VV_CUDA(cudaGetDevice(&dev));
  memset(&p2p.peers, 0, sizeof(p2p.peers));
  for (int g = 0; g < G && ok; ++g) {
    char* preg; char* pblk = nullptr; char* pwm;
    if (g == c.rank) {
      preg = reg; pblk = static_cast<char*>(W_hi.base); pwm = static_cast<char*>(W.base);
    } else {
      void* m = nullptr;
      e = cudaIpcOpenMemHandle(&m, all[g].region, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { ok = 0; p2p.why_off = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); cudaGetLastError(); break; }
      p2p.opened[p2p.n_opened++] = m; preg = static_cast<char*>(m);
      if (has_op) {
        e = cudaIpcOpenMemHandle(&m, all[g].wblk, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { ok = 0; p2p.why_off = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); cudaGetLastError(); break; }
        p2p.opened[p2p.n_opened++] = m; pblk = static_cast<char*>(m);
      }
      e = cudaIpcOpenMemHandle(&m, all[g].wm, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { ok = 0; p2p.why_off = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e); cudaGetLastError(); break; }
      p2p.opened[p2p.n_opened++] = m; pwm = static_cast<char*>(m);
    }
    p2p.peers.recv_dw[g] = reinterpret_cast<float*>(preg);
    p2p.peers.recv_small[g] = reinterpret_cast<float*>(preg + p2p.off_small);
    p2p.peers.flags[g] = reinterpret_cast<unsigned int*>(preg + p2p.off_flags);
    p2p.peers.wlast[g] = reinterpret_cast<float*>(preg + p2p.off_wlast);
    p2p.peers.Wm[g] = reinterpret_cast<float*>(pwm);
    p2p.peers.wop_hi[g] = has_op ? pblk + ho : nullptr;
    p2p.peers.wop_lo[g] = (has_op && lo_off) ? pblk + lo_off : nullptr;
  }
  // every rank must agree (one that could not map its peers would otherwise wait for pushes that go elsewhere)
  float* agree = reinterpret_cast<float*>(reg + p2p.off_flags) + kDpFlagCtr + 3;
  const float mine_ok = ok ? 0.f : 1.f;
  VV_CUDA(cudaMemcpy(agree, &mine_ok, 4, cudaMemcpyHostToDevice));
  r = g_nccl.AllReduce(agree, agree, 1, kNcclFloat, kNcclSum, comm, stream);
  if (r != 0) { set_error("ncclAllReduce failed (%d)", r); return VV_ERR_NCCL; }
  VV_CUDA(cudaStreamSynchronize(stream));
  float bad = 0.f; VV_CUDA(cudaMemcpy(&bad, agree, 4, cudaMemcpyDeviceToHost));
  VV_CUDA(cudaMemset(agree, 0, 4));
  if (bad != 0.f) { if (p2p.why_off.empty()) p2p.why_off = "a peer could not map this rank's memory"; return VV_OK; }
  if (wlast.base) { cudaFree(wlast.base); wlast.base = nullptr; }
  wlast.p = reg + p2p.off_wlast;
  const char* rep = getenv("VV_DP_REPLICATE_MASTER");
  p2p.replicate_master = (rep && atoi(rep) != 0) ? 1 : 0;
  const char* wk = getenv("VV_DP_WAIT_KERNEL");
  p2p.in_kernel_wait = !(wk && atoi(wk) != 0);
  p2p.on = true;
  return VV_OK;
}

extern "C" vv_trainer_t* vv_trainer_create(const vv_trainer_cfg_t* cfg, vv_stream_t stream) {
  if (!cfg) { set_error("trainer cfg is NULL"); return nullptr; }
  if (cfg->B < 1 || cfg->C < 2 || cfg->Nn < 1 || cfg->K < 4 || cfg->N < 4 || (cfg->K % 4) || (cfg->N % 4)) {
    set_error("bad trainer cfg: B=%d C=%d Nn=%d K=%d N=%d", cfg->B, cfg->C, cfg->Nn, cfg->K, cfg->N); return nullptr;
  }
  if (cfg->prec < VV_PREC_FP32_SIMT || cfg->prec > VV_PREC_F16X3) { set_error("bad precision %d", cfg->prec); return nullptr; }
  if (cfg->prec != VV_PREC_FP32_SIMT && ((cfg->K % 8) || (cfg->N % 8))) {
    set_error("trainer: the tensor-core precisions need K %% 8 == 0 and N %% 8 == 0 (K=%d N=%d); use VV_PREC_FP32_SIMT", cfg->K, cfg->N);
    return nullptr;
  }
  if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size) { set_error("bad rank/world_size"); return nullptr; }
  if (vv_device_check() != VV_OK) return nullptr;
  vv_trainer* t = new vv_trainer();
  t->cfg = *cfg;
  t->stream = reinterpret_cast<cudaStream_t>(stream);
  if (t->init() != VV_OK) { delete t; return nullptr; }
  return t;
}
extern "C" void vv_trainer_destroy(vv_trainer_t* t) { delete t; }
extern "C" float* vv_trainer_weight(vv_trainer_t* t) { return t->W.as<float>(); }
extern "C" float* vv_trainer_bias(vv_trainer_t* t) { return t->b.as<float>(); }
extern "C" float* vv_trainer_weight_hist(vv_trainer_t* t) { return t->Wh.as<float>(); }
extern "C" float* vv_trainer_bias_hist(vv_trainer_t* t) { return t->bh.as<float>(); }
extern "C" float* vv_trainer_weight_diff(vv_trainer_t* t) { return t->dW_parts.as<float>(); }
extern "C" float* vv_trainer_bias_diff(vv_trainer_t* t) { return t->dbx.as<float>(); }
extern "C" float* vv_trainer_blob(vv_trainer_t* t, const char* name) {
  const std::string n(name ? name : "");
  if (n == "X") return t->Xf.as<float>();
  if (n == "Z") return t->Zf.as<float>();
  if (n == "H") return t->H.as<float>();
  if (n == "dZ") return t->dZf.as<float>();
  if (n == "stats") return t->stats.as<float>();
  if (n == "loss") return t->loss_ptr();
  if (n == "violations") return t->viol_ptr();
  if (n == "dW_raw") return t->dW_parts.as<float>();
  if (n == "db_raw") return t->dbx.as<float>();
  if (n == "dX") return t->dX.as<float>();
  if (n == "wlast") return t->wlast.as<float>();
  if (n == "dZop_hi") return t->dZ_hi.as<float>();   // raw operand planes of dZ (F16X3: the 128-byte header precedes hi)
  if (n == "dZop_lo") return t->dZ_lo.as<float>();
  if (n == "Wop_hi") return t->W_hi.as<float>();     // raw operand planes of W (layout per precision, vv_operand_bytes)
  if (n == "Wop_lo") return t->W_lo.as<float>();
  set_error("unknown trainer blob '%s'", n.c_str());
  return nullptr;
}
extern "C" int vv_trainer_last_launches(const vv_trainer_t* t) { return t->last_launches; }

extern "C" int vv_trainer_sync_weights(vv_trainer_t* t) {
  const int64_t NK = int64_t(t->cfg.N) * t->cfg.K;
  vv_stream_t s = reinterpret_cast<vv_stream_t>(t->stream);
  int rc = vv_copy_strided(t->W.as<float>() + (t->cfg.K - 1), t->cfg.K, t->wlast.as<float>(), 1, t->cfg.N, 1, s);   // wlast = W[:, K-1]
  if (rc) return rc;
  return vv_prepare_operand(t->W.as<float>(), NK, t->cfg.prec, t->W_hi.p, t->W_lo.p, s);
}

extern "C" int vv_trainer_set_bank(vv_trainer_t* t, const float* bank, int64_t bank_rows) {
  if (!t || !bank || bank_rows <= 0) { set_error("set_bank: bad arguments"); return VV_ERR_INVALID; }
  const vv_trainer_cfg_t& c = t->cfg;
  // materialised path: exact-fp32 mode, blob inspection, and the 4-byte operand formats (the gather producers copy
  // 16-byte chunks of 2-byte elements)
  if (c.prec != VV_PREC_F16X3 && c.prec != VV_PREC_BF16) { t->bank_reg = nullptr; return VV_OK; }
  if (c.keep_blobs) { t->bank_reg = nullptr; return VV_OK; }
  vv_stream_t s = reinterpret_cast<vv_stream_t>(t->stream);
  const int64_t n = bank_rows * c.K;
  int rc;
  if ((rc = alloc_operand(t->bank_hi, t->bank_lo, size_t(n), c.prec))) return rc;
  if ((rc = vv_prepare_bank_operand(bank, bank_rows, c.K, c.prec, t->bank_hi.p, t->bank_lo.p, s))) return rc;
  t->bank_reg = bank; t->bank_reg_rows = bank_rows;
  return VV_OK;
}

extern "C" int vv_trainer_step(vv_trainer_t* t, const float* bank, int64_t bank_rows, const int32_t* idx,
                               const int32_t* quirk, const uint32_t* mask, int iter, int do_update) {
  const vv_trainer_cfg_t& c = t->cfg;
  vv_stream_t s = reinterpret_cast<vv_stream_t>(t->stream);
  const int M = t->M, N = c.N, K = c.K;
  const int64_t NK = int64_t(N) * K;
  int rc;
  launches_reset();
  if (t->timing && t->timed_steps >= 256) { set_error("timing: read vv_trainer_phase_ms at least every 256 steps"); return VV_ERR_INVALID; }
  // K0: either a gather plan (fused path: the GEMMs fetch the bank rows themselves) or the materialised X
  const bool fused_gather = t->bank_reg != nullptr && bank == t->bank_reg && bank_rows == t->bank_reg_rows;
  t->tic(0);
  if (fused_gather) {
    if (t->bad_idx_host && *t->bad_idx_host) {
      set_error("an earlier step's indices named bank rows outside [0, %lld): the bank and the sampler do not belong together",
                (long long)t->bank_reg_rows);
      return VV_ERR_INVALID;
    }
    if ((rc = vv_gather_plan_checked(bank, bank_rows, K, idx, quirk, c.B, t->R, t->rowmap.as<int32_t>(), t->delta.as<float>(),
                                     t->bad_idx_dev, s))) return rc;
  } else {
    if ((rc = t->alloc_x())) return rc;
    if (c.prec == VV_PREC_F16X3 && (bank != t->scaled_bank || bank_rows != t->scaled_bank_rows)) {
      // one pass over a newly seen bank fixes the scale of X for good (not counted in the per-step timing)
      if ((rc = vv_operand_set_scale(t->X_hi.p, c.prec, 0, s))) return rc;
      if ((rc = vv_operand_measure(t->X_hi.p, c.prec, bank, bank_rows * int64_t(K), s))) return rc;
      if ((rc = vv_operand_rescale(t->X_hi.p, c.prec, 12, s))) return rc;
      t->scaled_bank = bank; t->scaled_bank_rows = bank_rows;
    }
    if ((rc = vv_gather_rows(bank, bank_rows, K, idx, quirk, c.B, t->R, t->Xf.as<float>(), t->X_hi.p, t->X_lo.p, c.prec,
                             nullptr, s))) return rc;
  }
  t->toc(0);
  // K1 forward with fused bias + ReLU + dropout
  vv_act_t act; memset(&act, 0, sizeof(act));
  act.relu = 1; act.negative_slope = 0.f;
  const bool has_dropout = c.dropout_ratio > 0.f;
  act.dropout_mode = has_dropout ? c.dropout_mode : VV_DROPOUT_NONE;
  // data parallel: every rank draws its own stream (rank 0 keeps the single-GPU one); without the fold all ranks would
  // apply the same mask to their local row m, unlike the global-batch run the ranks stand for
  act.dropout_ratio = c.dropout_ratio; act.mask = mask; act.step = uint64_t(iter);
  act.seed = c.rank > 0 ? (c.dropout_seed ^ splitmix64(uint64_t(c.rank))) : c.dropout_seed;
  if (has_dropout && (act.dropout_mode == VV_DROPOUT_MASK01 || act.dropout_mode == VV_DROPOUT_MASK_U32) && !mask) {
    set_error("trainer: dropout mask mode needs a mask"); return VV_ERR_INVALID;
  }
  t->tic(1);
  // data parallel: the rows of W written by the other ranks' last update kernels are awaited by the forward GEMM's TMA
  // producer lane (tensor-core precisions) or by a one-warp kernel in front of it
  DpWait dpw = {nullptr, 0, 0u, nullptr, 0ull};
  const DpWait* wait = nullptr;
  if ((rc = t->dp_check_error())) return rc;
  if (t->p2p.on && t->p2p.waited < t->p2p.seq) {
    if (t->p2p.in_kernel_wait && c.prec != VV_PREC_FP32_SIMT) {
      dpw.flags = t->p2p.peers.flags[c.rank]; dpw.G = c.world_size; dpw.seq = t->p2p.seq; dpw.err = t->p2p.err_dev;
      dpw.timeout_ns = dp_wait_timeout_ns();
      wait = &dpw;
    } else {
      if ((rc = dp_wait_w_ready(t->p2p.peers.flags[c.rank], c.world_size, t->p2p.seq, t->p2p.err_dev, s))) return rc;
    }
    t->p2p.waited = t->p2p.seq;
  }
  // the last, partial wave of the persistent forward runs as K halves (VV_FWD_TAIL=0: whole units only)
  static const bool want_tail = [] { const char* e = getenv("VV_FWD_TAIL"); return !(e && atoi(e) == 0); }();
  FwdTail tail = {t->tail_ws.as<float>(), t->tail_ws.bytes, t->tail_flags.as<unsigned int>(), t->tail_flags.bytes / 4, ++t->tail_epoch};
  const FwdTail* tailp = (want_tail && c.prec != VV_PREC_FP32_SIMT) ? &tail : nullptr;
  if (fused_gather) {
    if ((rc = ip_forward_gathered_ex(t->opBank(), bank_rows, t->rowmap.as<int32_t>(), t->delta.as<float>(), t->wlast.as<float>(),
                                     t->opW(), t->b.as<float>(), M, N, K, c.prec, &act, nullptr, t->H.as<float>(), wait, s, tailp))) return rc;
  } else {
    if ((rc = ip_forward_ex(t->opX(), t->opW(), t->b.as<float>(), M, N, K, c.prec, &act, t->Zf.as<float>(),
                            t->H.as<float>(), wait, s, tailp))) return rc;
  }
  t->toc(1);
  const float dscale = has_dropout ? dropout_scale(c.dropout_ratio) : 1.f;
  auto run_rank = [&]() -> int {
    int rc;
    if (vv_rank_loss_fused_supported(&t->rank) && !c.split_rank_loss) {
      // K2 + K3 in one pass over H (+ bias gradient); timed as phase 3, phase 2 stays 0
      t->tic(3);
      // the second-generation kernel STORES db, dq, loss and violations (fixed-order sums through its workspace); the
      // first-generation one accumulates with atomics into zeroed words
      if (!rank_loss_fused_v2_applies(&t->rank, c.prec, t->dZf.p != nullptr, false)) {
        VV_CUDA(cudaMemsetAsync(t->dbx.p, 0, size_t(2 * N + 4) * 4, t->stream));    // db, loss, viol, ticket, dq
        count_launch();
      }
      if ((rc = rank_loss_fused_counted(t->H.as<float>(), &t->rank, c.loss_weight, 1, dscale, t->stats.as<float>(), nullptr, nullptr,
                                        t->item_loss.as<float>(), t->item_viol.as<float>(), t->loss_ptr(), t->viol_ptr(),
                                        t->dZf.as<float>(), t->dZ_hi.p, t->dZ_lo.p, c.prec, t->dbx.as<float>(),
                                        fused_gather ? t->delta.as<float>() : nullptr, fused_gather ? t->dq.as<float>() : nullptr,
                                        reinterpret_cast<unsigned int*>(t->dbx.as<float>() + N + 2), s, t->rank_ws.p, t->rank_ws.bytes))) return rc;
    } else {
      // K2
      t->tic(2);
      if ((rc = vv_rank_loss_forward(t->H.as<float>(), &t->rank, t->stats.as<float>(), nullptr, nullptr,
                                     t->item_loss.as<float>(), t->item_viol.as<float>(), t->loss_ptr(), t->viol_ptr(), s))) return rc;
      t->toc(2);
      // K3 (+ bias gradient)
      t->tic(3);
      VV_CUDA(cudaMemsetAsync(t->dbx.p, 0, size_t(N) * 4, t->stream));          // the forward kernel above wrote loss / viol
      if (fused_gather) VV_CUDA(cudaMemsetAsync(t->dq.p, 0, size_t(N) * 4, t->stream));
      count_launch();
      if ((rc = vv_rank_loss_backward_ex(t->H.as<float>(), &t->rank, t->stats.as<float>(), c.loss_weight, 1, dscale,
                                         t->dZf.as<float>(), t->dZ_hi.p, t->dZ_lo.p, c.prec, t->dbx.as<float>(),
                                         fused_gather ? t->delta.as<float>() : nullptr, fused_gather ? t->dq.as<float>() : nullptr, s))) return rc;
    }
    return VV_OK;
  };
  if (c.prec == VV_PREC_F16X3) {
    // the dZ operand's scale trails the previous step's max|dZ| (x64 headroom, saturating conversion); the very
    // first step runs the kernel once more up front just to measure
    if (!t->dz_scale_ready) { if ((rc = run_rank())) return rc; t->dz_scale_ready = true; }
    if ((rc = vv_operand_rescale(t->dZ_hi.p, c.prec, 10, s))) return rc;
  }
  if ((rc = run_rank())) return rc;
  t->toc(3);
  // K1 wgrad into split-K slabs
  int nparts = t->nsplit;
  float gscale = 1.f;
  // Data parallel + gathered wgrad can produce the outputs in slices of whole 256-wide tiles, with the split-K
  // reduction + NCCL all-reduce of slice i on the side stream under the GEMM of slice i+1.  Measured on 2xB200 this
  // LOSES (1.50 -> 1.58 ms/step): the persistent GEMM owns every SM's register file, so the concurrent NCCL kernel and
  // the GEMM's CTAs wait for each other, and two half-size launches quantise worse.  Kept behind VV_DP_SLICES (2 / 4).
  static const int want_slices = [] { const char* e = getenv("VV_DP_SLICES"); return e ? atoi(e) : 1; }();
  const int nslices = (want_slices > 1 && want_slices <= 4 && c.world_size > 1 && fused_gather && t->comm && !t->p2p.on &&
                       N % (256 * want_slices) == 0) ? want_slices : 1;
  const bool use_p2p = t->p2p.on && do_update;           // the exchange runs inside the update kernel
  const bool fold_col = fused_gather && (c.world_size == 1 || use_p2p) && do_update;
  // The update (one GPU) / the push to the owner ranks (data parallel) as the split-K finish INSIDE the wgrad kernel
  // (WgradFinish, vv_gemm.cuh): no update launch, no second pass over the slabs.  VV_FUSED_UPDATE=0 keeps the separate
  // K4 launch (A/B timing).
  static const bool want_finish = [] { const char* e = getenv("VV_FUSED_UPDATE"); return !(e && atoi(e) == 0); }();
  const bool fuse_finish = want_finish && do_update && (c.world_size == 1 || use_p2p) && c.prec != VV_PREC_FP32_SIMT && nslices == 1;
  if (fuse_finish) {
    const float rate = vv_learning_rate(c.lr_policy, c.base_lr, c.gamma, c.power, c.stepsize, iter);
    if (rate < 0.f) return VV_ERR_INVALID;
    if (c.compute_dgrad) {             // reads W: before the wgrad kernel rewrites it
      t->tic(5);
      if ((rc = vv_ip_dgrad(t->opdZ(), t->opW(), M, N, K, c.prec, t->dX.as<float>(), s))) return rc;
      t->toc(5);
    }
    // F16X3: the scale of the new W operand copy, from max|W| as the previous update (or the initial copy) recorded it
    // (data parallel: over all owners' rows, so every rank derives the same scale)
    if (use_p2p) { if ((rc = operand_rescale_ex(t->W_hi.p, c.prec, 10, t->p2p.peers.flags[c.rank] + kDpFlagAmax, c.world_size, s))) return rc; }
    else { if ((rc = vv_operand_rescale(t->W_hi.p, c.prec, 10, s))) return rc; }
    WgradFinish f;
    f.tickets = t->tickets.as<unsigned int>(); f.epoch = ++t->finish_epoch;
    UpdateTail& u = f.u;
    u.W = t->W.as<float>(); u.parts = t->dW_parts.as<float>(); u.nparts = t->nsplit; u.stride = NK; u.hist = t->Wh.as<float>();
    u.diff_out = t->dW_parts.as<float>(); u.count = NK; u.K = K;
    u.rate_w = rate * c.lr_mult[0]; u.decay_w = c.weight_decay * c.decay_mult[0];
    u.col_add = fold_col ? t->dq.as<float>() : nullptr; u.col_out = t->wlast.as<float>();
    u.Wop_hi = t->W_hi.p; u.Wop_lo = t->W_lo.p; u.prec = c.prec;
    u.b = t->b.as<float>(); u.db = t->dbx.as<float>(); u.bh = t->bh.as<float>(); u.b_diff = t->dbx.as<float>(); u.nb = N;
    u.rate_b = rate * c.lr_mult[1]; u.decay_b = c.weight_decay * c.decay_mult[1];
    u.momentum = c.momentum; u.reg_type = c.reg_type; u.gscale = 1.f;
    if (use_p2p) {
      f.mode = 2; f.G = c.world_size; f.rank = c.rank; f.rows_per = t->p2p.rows_per; f.seq = t->p2p.seq + 1;
      f.col_add = fold_col ? t->dq.as<float>() : nullptr;
      f.small_src = t->dbx.as<float>(); f.nsmall = N + 2; f.small_stride = t->p2p.small_stride;
      f.peers = t->p2p.peers;
    } else {
      f.mode = 1;
    }
    t->tic(4);
    if (fused_gather) {
      const double reg = double(c.regularization) / 2;
      if (reg > 0) { if ((rc = vv_axpby(N, float(1.0 + reg), t->dq.as<float>(), 0.f, t->dq.as<float>(), s))) return rc; }
      if ((rc = ip_wgrad_gathered_ex(t->opdZ(), t->opBank(), bank_rows, t->rowmap.as<int32_t>(), M, N, K, c.prec, c.regularization,
                                     t->dW_parts.as<float>(), t->nsplit, 0, N, &f, s))) return rc;
    } else {
      if ((rc = ip_wgrad_ex(t->opdZ(), t->opX(), M, N, K, c.prec, c.regularization, t->dW_parts.as<float>(), t->nsplit,
                            nullptr, 0, &f, s))) return rc;
    }
    t->toc(4);
    if (use_p2p) {
      t->tic(7);
      DpExchange x;
      x.G = c.world_size; x.rank = c.rank; x.seq = t->p2p.seq + 1;
      x.parts = nullptr; x.nparts = 0; x.stride = NK; x.col_add = nullptr;       // phase A happened in the wgrad kernel
      x.small_src = t->dbx.as<float>(); x.nsmall = N + 2; x.small_stride = t->p2p.small_stride;
      x.hist = t->Wh.as<float>(); x.diff_out = t->dW_parts.as<float>(); x.count = NK; x.K = K; x.rows_per = t->p2p.rows_per;
      x.rate_w = u.rate_w; x.decay_w = u.decay_w; x.momentum = c.momentum;
      x.reg_type = c.reg_type; x.gscale = 1.f / float(c.world_size); x.prec = c.prec;
      x.b = t->b.as<float>(); x.bh = t->bh.as<float>(); x.b_diff = t->dbx.as<float>(); x.nb = N;
      x.rate_b = u.rate_b; x.decay_b = u.decay_b;
      x.loss_out = t->loss_ptr(); x.viol_out = t->viol_ptr();
      x.peers = t->p2p.peers;
      if ((rc = dp_exchange_update(x, t->p2p.err_dev, t->p2p.replicate_master, s))) return rc;
      t->p2p.seq += 1;
      t->toc(7);
    }
    t->last_launches = launches_reset();
    if (t->timing) ++t->timed_steps;
    return VV_OK;
  }
  t->tic(4);
  if (fused_gather) {
    const double reg = double(c.regularization) / 2;
    // the K-1 copy quirk's share of dW[:, K-1] (scaled like the GEMM output)
    if (reg > 0) { if ((rc = vv_axpby(N, float(1.0 + reg), t->dq.as<float>(), 0.f, t->dq.as<float>(), s))) return rc; }
    const int cols = N / nslices;
    for (int sl = 0; sl < nslices; ++sl) {
      const int n0 = sl * cols;
      float* slice = t->dW_parts.as<float>() + size_t(n0) * K;
      if ((rc = vv_ip_wgrad_gathered_part(t->opdZ(), t->opBank(), bank_rows, t->rowmap.as<int32_t>(), M, N, K, c.prec, c.regularization,
                                          t->dW_parts.as<float>(), t->nsplit, n0, cols, s))) return rc;
      // the quirk's share of dW[:, K-1]: folded into the update kernel on a single GPU; a data-parallel rank (dq is
      // rank-local and must be in before the all-reduce) and a no-update step add it here
      if (!fold_col) { if ((rc = vv_add_column(slice, K, K - 1, t->dq.as<float>() + n0, cols, s))) return rc; }
      if (nslices > 1) {
        if (nparts > 1) { if ((rc = vv_reduce_parts(slice, nparts, NK, int64_t(cols) * K, slice, s))) return rc; }
        VV_CUDA(cudaEventRecord(t->ev_slice[sl], t->stream));
        VV_CUDA(cudaStreamWaitEvent(t->comm_stream, t->ev_slice[sl], 0));
        const bool last = sl == nslices - 1;
        const bool grp = last && g_nccl.GroupStart && g_nccl.GroupEnd;
        if (grp) g_nccl.GroupStart();
        ncclResult_t r1 = g_nccl.AllReduce(slice, slice, size_t(cols) * K, kNcclFloat, kNcclSum, t->comm, t->comm_stream);
        ncclResult_t r2 = 0;
        if (last) r2 = g_nccl.AllReduce(t->dbx.p, t->dbx.p, size_t(N + 2), kNcclFloat, kNcclSum, t->comm, t->comm_stream);
        if (grp) { const ncclResult_t r3 = g_nccl.GroupEnd(); if (r1 == 0 && r2 == 0) r1 = r3; }
        if (r1 != 0 || r2 != 0) { set_error("ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r1 ? r1 : r2) : "?"); return VV_ERR_NCCL; }
        count_launch(last ? 2 : 1);
      }
    }
    if (nslices > 1) nparts = 1;
  } else {
    if ((rc = vv_ip_wgrad(t->opdZ(), t->opX(), M, N, K, c.prec, c.regularization, t->dW_parts.as<float>(), t->nsplit,
                          nullptr, 0, s))) return rc;
  }
  t->toc(4);
  if (c.compute_dgrad) {
    t->tic(5);
    if ((rc = vv_ip_dgrad(t->opdZ(), t->opW(), M, N, K, c.prec, t->dX.as<float>(), s))) return rc;
    t->toc(5);
  }
  if (c.world_size > 1 && !use_p2p) {
    t->tic(6);
    if (!t->comm) { set_error("trainer: world_size > 1 but vv_dp_init was not called"); return VV_ERR_NCCL; }

    if (nslices == 1) {
      if (nparts > 1) {
        if ((rc = vv_reduce_parts(t->dW_parts.as<float>(), nparts, NK, NK, t->dW_parts.as<float>(), s))) return rc;
        nparts = 1;
      }
      VV_CUDA(cudaEventRecord(t->ev_grad, t->stream));
      VV_CUDA(cudaStreamWaitEvent(t->comm_stream, t->ev_grad, 0));
      // dW and (db, loss, violations) as one NCCL group: a single fused launch
      const bool grp = g_nccl.GroupStart && g_nccl.GroupEnd;
      if (grp) g_nccl.GroupStart();
      ncclResult_t r1 = g_nccl.AllReduce(t->dW_parts.p, t->dW_parts.p, size_t(NK), kNcclFloat, kNcclSum, t->comm, t->comm_stream);
      ncclResult_t r2 = g_nccl.AllReduce(t->dbx.p, t->dbx.p, size_t(N + 2), kNcclFloat, kNcclSum, t->comm, t->comm_stream);
      if (grp) { const ncclResult_t r3 = g_nccl.GroupEnd(); if (r1 == 0 && r2 == 0) r1 = r3; }
      if (r1 != 0 || r2 != 0) { set_error("ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r1 ? r1 : r2) : "?"); return VV_ERR_NCCL; }
      count_launch(2);
    }
    VV_CUDA(cudaEventRecord(t->ev_comm, t->comm_stream));
    VV_CUDA(cudaStreamWaitEvent(t->stream, t->ev_comm, 0));
    gscale = 1.f / float(c.world_size);
    // loss was summed over ranks -> mean over ranks (global-batch mean), whether or not an update follows;
    // violations stay the global count
    if ((rc = vv_axpby(1, gscale, t->loss_ptr(), 0.f, t->loss_ptr(), s))) return rc;
    t->toc(6);
  }
  if (use_p2p) {
    t->tic(7);
    const float rate = vv_learning_rate(c.lr_policy, c.base_lr, c.gamma, c.power, c.stepsize, iter);
    if (rate < 0.f) return VV_ERR_INVALID;
    // F16X3: every rank derives the same new scale of W from max|W| over all owners' rows of the previous update
    if ((rc = operand_rescale_ex(t->W_hi.p, c.prec, 10, t->p2p.peers.flags[c.rank] + kDpFlagAmax, c.world_size, s))) return rc;
    DpExchange x;
    x.G = c.world_size; x.rank = c.rank; x.seq = t->p2p.seq + 1;
    x.parts = t->dW_parts.as<float>(); x.nparts = nparts; x.stride = NK;
    x.col_add = fold_col ? t->dq.as<float>() : nullptr;
    x.small_src = t->dbx.as<float>(); x.nsmall = N + 2; x.small_stride = t->p2p.small_stride;
    x.hist = t->Wh.as<float>(); x.diff_out = t->dW_parts.as<float>(); x.count = NK; x.K = K; x.rows_per = t->p2p.rows_per;
    x.rate_w = rate * c.lr_mult[0]; x.decay_w = c.weight_decay * c.decay_mult[0]; x.momentum = c.momentum;
    x.reg_type = c.reg_type; x.gscale = 1.f / float(c.world_size); x.prec = c.prec;
    x.b = t->b.as<float>(); x.bh = t->bh.as<float>(); x.b_diff = t->dbx.as<float>(); x.nb = N;
    x.rate_b = rate * c.lr_mult[1]; x.decay_b = c.weight_decay * c.decay_mult[1];
    x.loss_out = t->loss_ptr(); x.viol_out = t->viol_ptr();
    x.peers = t->p2p.peers;
    if ((rc = dp_exchange_update(x, t->p2p.err_dev, t->p2p.replicate_master, s))) return rc;
    t->p2p.seq += 1;
    t->toc(7);
  } else if (do_update) {
    t->tic(7);
    // ref: solver.cpp:486-576 + net.cpp:804-839; weight then bias (net.params() order)
    const float rate = vv_learning_rate(c.lr_policy, c.base_lr, c.gamma, c.power, c.stepsize, iter);
    if (rate < 0.f) return VV_ERR_INVALID;
    // F16X3: the new W operand copy is scaled from max|W| as the previous update (or the initial copy) recorded it
    if ((rc = vv_operand_rescale(t->W_hi.p, c.prec, 10, s))) return rc;
    // weight then bias (net.params() order), one launch: + the quirk column, + wlast = W[:, K-1] for the next plan
    UpdateTail u;
    u.W = t->W.as<float>(); u.parts = t->dW_parts.as<float>(); u.nparts = nparts; u.stride = NK; u.hist = t->Wh.as<float>();
    u.diff_out = t->dW_parts.as<float>(); u.count = NK; u.K = K;
    u.rate_w = rate * c.lr_mult[0]; u.decay_w = c.weight_decay * c.decay_mult[0];
    u.col_add = fold_col ? t->dq.as<float>() : nullptr; u.col_out = t->wlast.as<float>();
    u.Wop_hi = t->W_hi.p; u.Wop_lo = t->W_lo.p; u.prec = c.prec;
    u.b = t->b.as<float>(); u.db = t->dbx.as<float>(); u.bh = t->bh.as<float>(); u.b_diff = t->dbx.as<float>(); u.nb = N;
    u.rate_b = rate * c.lr_mult[1]; u.decay_b = c.weight_decay * c.decay_mult[1];
    u.momentum = c.momentum; u.reg_type = c.reg_type; u.gscale = gscale;
    if ((rc = sgd_update_tail(u, s))) return rc;
    t->toc(7);
  } else if (nparts > 1) {
    if ((rc = vv_reduce_parts(t->dW_parts.as<float>(), nparts, NK, NK, t->dW_parts.as<float>(), s))) return rc;
  }
  t->last_launches = launches_reset();
  if (t->timing) ++t->timed_steps;
  return VV_OK;
}

extern "C" int vv_trainer_set_timing(vv_trainer_t* t, int enable) {
  if (!t) return VV_ERR_INVALID;
  t->timing = enable != 0; t->timed_steps = 0;
  return VV_OK;
}
extern "C" int vv_trainer_phase_ms(vv_trainer_t* t, float* ms_out, int* steps_out) {
  if (!t || !ms_out) return VV_ERR_INVALID;
  VV_CUDA(cudaStreamSynchronize(t->stream));
  for (int p = 0; p < VV_NUM_PHASES; ++p) ms_out[p] = 0.f;
  const bool has_dgrad = t->cfg.compute_dgrad != 0, has_comm = t->cfg.world_size > 1;
  for (int sidx = 0; sidx < t->timed_steps; ++sidx)
    for (int p = 0; p < VV_NUM_PHASES; ++p) {
      if ((p == 5 && !has_dgrad) || (p == 6 && !has_comm)) continue;
      const size_t base = (size_t(sidx) * VV_NUM_PHASES + p) * 2;
      if (base + 1 >= t->ev_pool.size()) continue;
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, t->ev_pool[base], t->ev_pool[base + 1]) == cudaSuccess) ms_out[p] += ms;
      else cudaGetLastError();
    }
  if (steps_out) *steps_out = t->timed_steps;
  t->timed_steps = 0;
  return VV_OK;
}

extern "C" int vv_trainer_extract(vv_trainer_t* t, const float* F, int64_t rows, float* out) {
  const vv_trainer_cfg_t& c = t->cfg;
  vv_stream_t s = reinterpret_cast<vv_stream_t>(t->stream);
  if (!F || !out || rows <= 0) { set_error("extract: bad arguments"); return VV_ERR_INVALID; }
  vv_act_t act; memset(&act, 0, sizeof(act));
  act.relu = 1; act.dropout_mode = VV_DROPOUT_NONE;      // TEST phase: dropout is a copy (dropout_layer.cpp:46-48)
  int rc;
  if ((rc = t->alloc_x())) return rc;
  t->scaled_bank = nullptr;             // F16X3: prepare_operand below rescales the X operand to this input
  launches_reset();
  for (int64_t r0 = 0; r0 < rows; r0 += t->M) {
    const int m = int(rows - r0 < t->M ? rows - r0 : t->M);
    vv_operand_t x;
    if (t->needs_f32_operand()) { x.hi = F + r0 * c.K; x.lo = nullptr; }
    else {
      if ((rc = vv_prepare_operand(F + r0 * c.K, int64_t(m) * c.K, c.prec, t->X_hi.p, t->X_lo.p, s))) return rc;
      x.hi = t->X_hi.p; x.lo = t->X_lo.p;
    }
    if ((rc = vv_ip_forward(x, t->opW(), t->b.as<float>(), m, c.N, c.K, c.prec, &act, nullptr, out + r0 * c.N, s))) return rc;
  }
  t->last_launches = launches_reset();
  return VV_OK;
}

// ---- data-parallel plumbing -------------------------------------------------
extern "C" int vv_dp_unique_id(void* id128) {
  if (!id128) return VV_ERR_INVALID;
  if (!g_nccl.load()) return VV_ERR_NCCL;
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != 0) { set_error("ncclGetUniqueId failed (%d)", r); return VV_ERR_NCCL; }
  memcpy(id128, &id, 128);
  return VV_OK;
}
extern "C" int vv_dp_init(vv_trainer_t* t, const void* id128) {
  if (!t || !id128) return VV_ERR_INVALID;
  if (t->cfg.world_size == 1) return VV_OK;
  if (!g_nccl.load()) return VV_ERR_NCCL;
  ncclUniqueId id; memcpy(&id, id128, 128);
  ncclResult_t r = g_nccl.CommInitRank(&t->comm, t->cfg.world_size, id, t->cfg.rank);
  if (r != 0) { set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"); return VV_ERR_NCCL; }
  return t->p2p_setup();
}
extern "C" int vv_dp_mode(const vv_trainer_t* t) { return !t || t->cfg.world_size == 1 ? 0 : (t->p2p.on ? 2 : 1); }
extern "C" const char* vv_dp_mode_reason(const vv_trainer_t* t) { return t ? t->p2p.why_off.c_str() : ""; }
extern "C" int vv_dp_gather_state(vv_trainer_t* t) {
  if (!t) return VV_ERR_INVALID;
  if (!t->p2p.on) return VV_OK;                       // NCCL mode: every rank already holds the whole state
  int rc;
  if ((rc = t->dp_check_error())) return rc;
  vv_stream_t s = reinterpret_cast<vv_stream_t>(t->stream);
  if (t->p2p.waited < t->p2p.seq) {
    if ((rc = dp_wait_w_ready(t->p2p.peers.flags[t->cfg.rank], t->cfg.world_size, t->p2p.seq, t->p2p.err_dev, s))) return rc;
    t->p2p.waited = t->p2p.seq;
  }
  const size_t own = size_t(t->p2p.rows_per) * t->cfg.K;
  float* bufs[3] = {t->W.as<float>(), t->Wh.as<float>(), t->dW_parts.as<float>()};
  for (int i = 0; i < 3; ++i) {
    if (i == 0 && (t->p2p.replicate_master || !t->W_hi.base)) continue;      // the master rows are already everywhere
    ncclResult_t r = g_nccl.AllGather(bufs[i] + own * t->cfg.rank, bufs[i], own, kNcclFloat, t->comm, t->stream);
    if (r != 0) { set_error("ncclAllGather failed (%d)", r); return VV_ERR_NCCL; }
  }
  VV_CUDA(cudaStreamSynchronize(t->stream));
  return t->dp_check_error();
}
extern "C" int vv_dp_allreduce_inplace(vv_trainer_t* t, float* buf, int64_t count, vv_stream_t stream) {
  if (!t || !buf || count <= 0) return VV_ERR_INVALID;
  if (t->cfg.world_size == 1) return VV_OK;
  if (!t->comm) { set_error("vv_dp_init was not called"); return VV_ERR_NCCL; }
  ncclResult_t r = g_nccl.AllReduce(buf, buf, size_t(count), kNcclFloat, kNcclSum, t->comm, reinterpret_cast<cudaStream_t>(stream));
  if (r != 0) { set_error("ncclAllReduce failed (%d)", r); return VV_ERR_NCCL; }
  return VV_OK;
}
