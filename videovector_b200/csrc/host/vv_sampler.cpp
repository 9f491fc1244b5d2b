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
#include <unordered_set>
#include <vector>
#include "vv_b200.h"

struct vv_glibc_rand {
  int32_t state[31];
  int f, r;
  explicit vv_glibc_rand(unsigned int seed) { reseed(seed); }
  void reseed(unsigned int seed) {
    if (seed == 0) seed = 1;
    state[0] = int32_t(seed);
    int32_t word = int32_t(seed);
    for (int i = 1; i < 31; ++i) {
      // 16807 * word % 2147483647 without overflow (Schrage), as glibc's srandom_r
      const long hi = word / 127773, lo = word % 127773;
      long w = 16807 * lo - 2836 * hi;
      if (w < 0) w += 2147483647;
      word = int32_t(w);
      state[i] = word;
    }
    f = 3; r = 0;
    for (int k = 0; k < 310; ++k) next();
  }
  int next() {
    const uint32_t val = uint32_t(state[f]) + uint32_t(state[r]);
    state[f] = int32_t(val);
    const int result = int(val >> 1);
    if (++f >= 31) { f = 0; ++r; }
    else if (++r >= 31) r = 0;
    return result;
  }
};

namespace {
inline uint64_t shot_key(int video_id, int shot_id) {
  return (uint64_t(uint32_t(video_id)) << 32) | uint32_t(shot_id);
}
}  // namespace

struct vv_sampler {
  int V, B, C, Nn, P, swap_pct, max_same;
  std::vector<int32_t> video_id, shot_off, shot_ids;
  vv_glibc_rand rng;
  int cursor;
  std::vector<float> buffer_ids;          // persistent permutation, float like the reference (:81-83)
  std::vector<int32_t> neg_row;           // buffer slot -> bank row currently held
  std::vector<uint64_t> slot_key;         // negative_id_to_key_
  std::unordered_set<uint64_t> key_set;   // negative_keys_set_
  std::vector<int32_t> last_full;         // [B,R] bank row of the last full-row write per slot, -1 = none
  vv_sampler(unsigned seed) : rng(seed), cursor(0) {}

  // random_unique over a range (rng.hpp:43-54)
  template <class T> void random_unique(T* first, int n, int num_random) {
    int left = n;
    while (num_random--) {
      std::swap(*first, first[rng.next() % left]);
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
      const int s = rng.next() % n;
      const uint64_t key = shot_key(video_id[v], shot_ids[shot_off[v] + s]);
      if (key_set.insert(key).second) {
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
          const int j = C + rng.next() % (i - C + 1);
          if (i != j) std::swap(ids[i], ids[j]);
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
        for (int j = 0; j < n; ++j) {
          const uint64_t key = shot_key(video_id[v], shot_ids[off + j]);
          if (key_set.find(key) == key_set.end()) {
            if ((rng.next() % 100) < swap_pct) {          // AddToBuffer :25-37
              const int pos = rng.next() % P;
              neg_row[pos] = off + j;
              key_set.erase(slot_key[pos]);
              slot_key[pos] = key;
              key_set.insert(key);
            }
          }
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
      negative_swap_percentage > 99 || (num_negative_samples > 0 && max_buffer_size < num_negative_samples))
    return nullptr;
  vv_sampler* s = new vv_sampler(rand_seed);
  s->V = num_videos; s->B = batch_size; s->C = context_size; s->Nn = num_negative_samples;
  s->P = num_negative_samples > 0 ? max_buffer_size : 0;
  s->swap_pct = negative_swap_percentage; s->max_same = max_same_video_negs;
  s->video_id.assign(video_id, video_id + num_videos);
  s->shot_off.assign(shot_off, shot_off + num_videos + 1);
  s->shot_ids.assign(shot_ids, shot_ids + shot_off[num_videos]);
  s->buffer_ids.resize(s->P);
  for (int i = 0; i < s->P; ++i) s->buffer_ids[i] = float(i);
  s->neg_row.assign(s->P, -1);
  s->slot_key.assign(s->P, 0);
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
