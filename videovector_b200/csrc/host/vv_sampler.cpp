// vv_sampler.cpp -- host-side index sampler of the TRAIN data layer (row S of the scope table).
//
// Emits, per batch, the bank row every slot of the data blob [B, R, K] would be
// filled from, instead of copying K floats per slot on the host: the features stay
// resident in HBM and K0 (vv_gather_rows) materialises / the GEMM consumes them.
//
// Behaviour mirrored (same rand() consumption order, same state machine):
//   ref: src/caffe/layers/video_sampled_shots_data_layer.cpp
//        :245-344 negative-buffer initialisation, :372-507 AddSamplesToTop (WINDOW),
//        :769-909 InternalThreadEntry, :25-44 AddToBuffer / RandomShuffleTopids
//   ref: include/caffe/util/rng.hpp:43-54 random_unique
//   libstdc++ std::random_shuffle(first,last) = for i in 1..n-1: swap(i, rand() % (i+1))
// The reference calls the process-global glibc rand() and never seeds it (seed 1).
// Here the generator is a private re-implementation of glibc's TYPE_3 random_r
// (r[i] = r[i-3] + r[i-31], 310 warm-up draws), bit-identical to rand() for a given
// seed (tests/test_sampler.py compares it against libc), so ranks and threads are
// independent of global state.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <unordered_map>
#include <vector>
#include "vv_b200.h"

struct vv_glibc_rand {
  // glibc TYPE_3 (random_r): with state r[0..30], fptr = &r[3], rptr = &r[0], draw k does
  // r[f] += r[q] and returns r[f] >> 1.  Unrolled onto a linear array L with
  // L[j] = r[(j + 3) % 31] for j < 31 this is  L[i] = L[i-31] + L[i-3]  and draw k = L[31 + k].
  // srandom_r discards the first 310 draws.  Outputs are produced in blocks so next() is a load.
  static constexpr int kBlock = 2048;
  uint32_t buf[31 + kBlock];
  int pos;
  explicit vv_glibc_rand(unsigned int seed) { reseed(seed); }
  void reseed(unsigned int seed) {
    if (seed == 0) seed = 1;
    int32_t r[31];
    r[0] = int32_t(seed);
    for (int i = 1; i < 31; ++i) {
      // 16807 * r[i-1] % 2147483647 without overflow (Schrage), as glibc's srandom_r
      const long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
      long w = 16807 * lo - 2836 * hi;
      if (w < 0) w += 2147483647;
      r[i] = int32_t(w);
    }
    for (int j = 0; j < 31; ++j) buf[j] = uint32_t(r[(j + 3) % 31]);
    generate();
    pos = 31 + 310;
  }
  void generate() {
    for (int i = 31; i < 31 + kBlock; ++i) buf[i] = buf[i - 31] + buf[i - 3];
  }
  void refill() {
    for (int i = 0; i < 31; ++i) buf[i] = buf[kBlock + i];
    generate();
    pos = 31;
  }
  int next() {
    if (pos >= 31 + kBlock) refill();
    return int(buf[pos++] >> 1);
  }
  int peek() {                      // the next draw, not consumed
    if (pos >= 31 + kBlock) refill();
    return int(buf[pos] >> 1);
  }
  void skip(int n) { pos += n; }    // consume n (0 or 1) peeked draws
};

namespace {
inline uint64_t shot_key(int video_id, int shot_id) {
  return (uint64_t(uint32_t(video_id)) << 32) | uint32_t(shot_id);
}
// x % d for 32-bit x without a hardware divide (Lemire's fastmod, exact for all 32-bit x, d):
// the sampler draws ~90 rand() % n per item and the divides dominated its cost.
struct FastMod {
  std::vector<uint64_t> magic;   // magic[d] = 2^64 / d + 1 for d < size
  explicit FastMod(int max_d) : magic(size_t(max_d) + 1, 0) {
    for (int d = 1; d <= max_d; ++d) magic[d] = ~uint64_t(0) / uint64_t(d) + 1;
  }
  inline int mod(int x, int d) const {
    if (size_t(d) >= magic.size()) return x % d;
    const uint64_t low = magic[d] * uint64_t(uint32_t(x));
    return int((static_cast<unsigned __int128>(low) * uint64_t(d)) >> 64);
  }
};
}  // namespace

struct vv_sampler {
  int V, B, C, Nn, P, swap_pct, max_same;
  std::vector<int32_t> video_id, shot_off, shot_ids;
  vv_glibc_rand rng;
  int cursor;
  std::vector<float> buffer_ids;          // persistent permutation, float like the reference (:81-83)
  std::vector<int32_t> neg_row;           // buffer slot -> bank row currently held
  // "video_id:shot_id" keys (negative_id_to_key_ / negative_keys_set_) as dense ids:
  // key_of[g] = id of the key of global shot g (equal keys share an id), so the set is a bitmap.
  std::vector<int32_t> key_of;            // [num_shots]
  std::vector<int32_t> slot_key;          // buffer slot -> key id
  std::vector<uint8_t> in_set;            // [num_keys] membership
  std::vector<int32_t> last_full;         // [B,R] bank row of the last full-row write per slot, -1 = none
  FastMod fm;
  vv_sampler(unsigned seed, int max_d) : rng(seed), cursor(0), fm(max_d) {}

  // random_unique over a range (rng.hpp:43-54)
  template <class T> void random_unique(T* first, int n, int num_random) {
    int left = n;
    while (num_random--) {
      std::swap(*first, first[fm.mod(rng.next(), left)]);
      ++first; --left;
    }
  }
  bool init(int max_tries) {
    int added = 0;
    for (long nid = 0; nid < (long)max_tries * P; ++nid) {
      const int v = cursor;
      cursor = (cursor + 1) % V;
      const int n = shot_off[v + 1] - shot_off[v];
      if (n <= 0) return false;
      const int s = fm.mod(rng.next(), n);
      const int32_t key = key_of[shot_off[v] + s];
      if (!in_set[key]) {
        in_set[key] = 1;
        neg_row[added] = shot_off[v] + s;
        slot_key[added] = key;
        ++added;
      }
      if (added >= P) break;
    }
    return added == P;
  }
  int next(int32_t* idx, int32_t* quirk) {
    const int R = C + Nn;
    int item = 0; long guard = 0;
    std::vector<int> ids;
    while (item < B) {
      if (++guard > 100L * (B + V)) return VV_ERR_INVALID;
      const int v = cursor;
      cursor = (cursor + 1) % V;
      const int off = shot_off[v], n = shot_off[v + 1] - off;
      if (n < 2 || n < C) continue;                       // :387-389, :427-429 (no rand consumed)
      ids.resize(n);
      std::iota(ids.begin(), ids.end(), 0);
      random_unique(ids.data(), n, C);                    // :432
      std::sort(ids.begin(), ids.begin() + C);            // :437
      const int half = C / 2;
      int32_t* I = idx + (size_t)item * R;
      int32_t* Q = quirk + (size_t)item * R;
      int32_t* L = last_full.data() + (size_t)item * R;
      int ctx = 0;
      for (int i = 0; i < C; ++i) {
        const int slot = (i == half) ? 0 : ++ctx;         // target -> slot 0, others in temporal order
        I[slot] = off + ids[i]; Q[slot] = -2; L[slot] = off + ids[i];
      }
      int added = 0;
      if (Nn > 0 && n > C) {                              // same-video negatives :479-503
        for (int i = C + 1; i < n; ++i) {                 // std::random_shuffle(ids+C, end)
          const int j = C + fm.mod(rng.next(), i - C + 1);
          std::swap(ids[i], ids[j]);                      // (i == j is a no-op: no branch needed)
        }
        for (int nid = C; nid < n && added < max_same; ++nid) {
          if (ids[nid] < ids[half - 1] || ids[nid] > ids[half + 1]) {
            const int slot = C + added;
            I[slot] = off + ids[nid];
            Q[slot] = L[slot];                            // K-1 floats copied (:492): element K-1 keeps the old value
            ++added;
          }
        }
      }
      if (Nn > 0) {                                       // buffer negatives :852-874
        random_unique(buffer_ids.data(), P, Nn - added);
        for (int s = C + added; s < C + Nn; ++s) {
          const int neg_id = static_cast<int>(buffer_ids[s - C - added]);
          I[s] = neg_row[neg_id]; Q[s] = -2; L[s] = neg_row[neg_id];
        }
      }
      ++item;
      if (Nn > 0 && swap_pct > 0) {                       // swap this record's shots into the buffer :888-906
        // AddToBuffer (:25-37) draws rand()%100 and, only if that is below the swap percentage, a second
        // rand()%P.  The second draw is peeked and consumed conditionally (rng.pos += take) and the buffer
        // update is made branch-free through a dummy slot P / dummy key, because the 50/50 decision is
        // unpredictable and the mispredictions dominated the sampler's cost.
        const int32_t dummy_key = int32_t(in_set.size()) - 1;
        for (int j = 0; j < n; ++j) {
          const int32_t key = key_of[off + j];
          if (in_set[key]) continue;                       // rare (the buffer holds P of all shots): well predicted
          const int take = fm.mod(rng.next(), 100) < swap_pct;
          const int pos = take ? fm.mod(rng.peek(), P) : P;
          rng.skip(take);
          in_set[slot_key[pos]] = 0;                       // slot P / dummy key absorb the not-taken case
          neg_row[pos] = off + j;
          slot_key[pos] = take ? key : dummy_key;
          in_set[take ? key : dummy_key] = uint8_t(take);
        }
      }
    }
    return VV_OK;
  }
};

extern "C" vv_sampler_t* vv_sampler_create(int num_videos, const int32_t* video_id, const int32_t* shot_off,
                                           const int32_t* shot_ids, int batch_size, int context_size,
                                           int num_negative_samples, int max_buffer_size,
                                           int negative_swap_percentage, int max_same_video_negs,
                                           int max_tries_for_negs, unsigned int rand_seed) {
  if (num_videos <= 0 || !video_id || !shot_off || !shot_ids || batch_size < 1 || context_size < 2 ||
      (context_size % 2) != 1 || num_negative_samples < 0 || negative_swap_percentage < 0 ||
      negative_swap_percentage > 99 || (num_negative_samples > 0 && max_buffer_size < num_negative_samples) ||
      // the reference writes max_same_video_negs slots unconditionally (:483-499) and then draws
      // Nn - added buffer negatives; more same-video negatives than slots is undefined there, an error here
      max_same_video_negs < 0 || max_same_video_negs > num_negative_samples)
    return nullptr;
  int max_d = max_buffer_size > 100 ? max_buffer_size : 100;
  for (int v = 0; v < num_videos; ++v) max_d = std::max(max_d, shot_off[v + 1] - shot_off[v]);
  if (max_d > (1 << 20)) max_d = 1 << 20;
  vv_sampler* s = new vv_sampler(rand_seed, max_d);
  s->V = num_videos; s->B = batch_size; s->C = context_size; s->Nn = num_negative_samples;
  s->P = num_negative_samples > 0 ? max_buffer_size : 0;
  s->swap_pct = negative_swap_percentage; s->max_same = max_same_video_negs;
  s->video_id.assign(video_id, video_id + num_videos);
  s->shot_off.assign(shot_off, shot_off + num_videos + 1);
  s->shot_ids.assign(shot_ids, shot_ids + shot_off[num_videos]);
  s->buffer_ids.resize(s->P);
  for (int i = 0; i < s->P; ++i) s->buffer_ids[i] = float(i);
  s->neg_row.assign(s->P + 1, -1);           // + one dummy slot
  s->slot_key.assign(s->P + 1, 0);
  {
    const int total = shot_off[num_videos];
    std::unordered_map<uint64_t, int32_t> ids;
    ids.reserve(size_t(total) * 2);
    s->key_of.resize(total);
    for (int v = 0; v < num_videos; ++v)
      for (int g = shot_off[v]; g < shot_off[v + 1]; ++g) {
        auto it = ids.emplace(shot_key(video_id[v], shot_ids[g]), int32_t(ids.size()));
        s->key_of[g] = it.first->second;
      }
    s->in_set.assign(ids.size() + 1, 0);     // + one dummy key for the branch-free swap loop
  }
  s->last_full.assign((size_t)batch_size * (context_size + num_negative_samples), -1);
  if (s->P > 0 && !s->init(max_tries_for_negs)) { delete s; return nullptr; }
  return s;
}
extern "C" void vv_sampler_destroy(vv_sampler_t* s) { delete s; }
extern "C" int vv_sampler_next(vv_sampler_t* s, int32_t* idx, int32_t* quirk) {
  if (!s || !idx || !quirk) return VV_ERR_INVALID;
  return s->next(idx, quirk);
}
extern "C" int vv_sampler_cursor(const vv_sampler_t* s) { return s ? s->cursor : -1; }

extern "C" vv_glibc_rand_t* vv_glibc_rand_create(unsigned int seed) { return new vv_glibc_rand(seed); }
extern "C" int vv_glibc_rand_next(vv_glibc_rand_t* g) { return g->next(); }
extern "C" void vv_glibc_rand_destroy(vv_glibc_rand_t* g) { delete g; }
