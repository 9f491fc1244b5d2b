// vv_dp_exchange.cuh -- internal interface of the data-parallel gradient exchange over peer memory.
//
// One process per GPU.  Every rank maps a small "exchange region" and the W operand block of every other rank
// (CUDA IPC, set up once in vv_dp_init) and the whole exchange of a training step happens inside ONE kernel of
// ours, with plain stores over NVLink and release/acquire flags -- no NCCL kernel on the step's critical path:
//
//   phase A  (reduce-scatter by push)  every rank sums its split-K slabs and writes rows [o*N/G, (o+1)*N/G) of the
//            result straight into rank o's receive buffer (slot = source rank); the (db, loss, violations) vector
//            goes to every rank.  The last CTA to finish raises dw_ready[me] on every rank.
//   phase B  (owner update + all-gather by push)  after dw_ready[0..G) the owner adds the G contributions in rank
//            order (deterministic; replicas are bit-identical by construction because one rank computes every
//            element), applies decay / momentum / update to ITS rows of W and of the history (the optimiser state
//            is sharded), and writes the new rows -- fp32 master + the GEMM operand copy + W[:, K-1] -- into every
//            rank.  The last CTA raises w_ready[me] everywhere.
//   consume  the next forward GEMM's TMA producer lane waits for w_ready[0..G) right before its first W tile
//            (the gather-plan kernel, the GEMM prologue and the row-gather producers run ahead of that wait).
#pragma once
#include "vv_common.cuh"

namespace vv {

constexpr int kDpMaxRanks = 8;
// flag words in the exchange region (uint32 index)
constexpr int kDpFlagDwReady = 0;     // [kDpMaxRanks] written by the source rank
constexpr int kDpFlagWReady = 16;     // [kDpMaxRanks] written by the owner rank
constexpr int kDpFlagAmax = 32;       // [kDpMaxRanks] bits of max|W| over the owner's rows (F16X3 scale upkeep)
constexpr int kDpFlagCtr = 48;        // [2] local arrival counters (phase A, phase B), [2] local amax accumulator
constexpr int kDpFlagWords = 64;

struct DpPeers {
  float* recv_dw[kDpMaxRanks];        // [G sources][N/G * K] on that rank
  float* recv_small[kDpMaxRanks];     // [G sources][small_stride]
  unsigned int* flags[kDpMaxRanks];   // kDpFlagWords words
  float* Wm[kDpMaxRanks];             // fp32 master W [N,K]
  void* wop_hi[kDpMaxRanks];          // operand copy of W (NULL for the fp32-operand precisions)
  void* wop_lo[kDpMaxRanks];
  float* wlast[kDpMaxRanks];          // [N] = W[:, K-1]
};

struct DpExchange {
  int G, rank; unsigned int seq;      // seq: 1, 2, ... = number of exchanged steps including this one
  // phase A
  const float* parts; int nparts; long long stride;      // split-K slabs
  const float* col_add;               // [N] added to column K-1 (the K-1 copy quirk's share) or NULL
  const float* small_src; int nsmall; int small_stride;  // db[N], loss, violations
  // phase B
  float* hist; float* diff_out; long long count; int K; int rows_per;
  float rate_w, decay_w, momentum; int reg_type; float gscale; int prec;
  float* b; float* bh; float* b_diff; int nb; float rate_b, decay_b;
  float* loss_out; float* viol_out;
  DpPeers peers;
};

// err_word: host-mapped word that receives a non-zero code when an in-kernel wait times out (VV_DP_TIMEOUT_MS, default
// 10 s): a dead peer then costs an error on the next host call instead of a hung GPU.
// replicate_master: also write the fp32 master rows into every rank (always on for the fp32-operand precisions);
// off = the master weights and the history are sharded by owner (vv_dp_gather_state collects them).
int dp_exchange_update(const DpExchange& x, unsigned int* err_word, int replicate_master, vv_stream_t stream);   // one launch
int dp_exchange_grid();                                                   // co-resident grid size used by the kernel
int dp_wait_w_ready(const unsigned int* flags, int G, unsigned int seq, unsigned int* err_word, vv_stream_t stream);   // stand-alone consumer wait
unsigned long long dp_wait_timeout_ns();

// What a consumer of W (the next forward GEMM) waits for: w_ready[0..G) >= seq in this rank's flag words.
struct DpWait { const unsigned int* flags; int G; unsigned int seq; unsigned int* err; unsigned long long timeout_ns; };

__device__ __forceinline__ void dp_st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int dp_ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long dp_globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag >= seq (sequence numbers only grow); gives up after timeout_ns and reports through *err
__device__ __forceinline__ void dp_spin_wait_flag(const unsigned int* flag, unsigned int seq, unsigned long long timeout_ns,
                                                  unsigned int* err, unsigned int code) {
  if (dp_ld_acquire_sys(flag) >= seq) return;
  const unsigned long long t0 = dp_globaltimer_ns();
  while (dp_ld_acquire_sys(flag) < seq) {
    __nanosleep(40);
    if (dp_globaltimer_ns() - t0 > timeout_ns) { if (err) atomicExch_system(err, code); return; }
  }
}

}  // namespace vv
