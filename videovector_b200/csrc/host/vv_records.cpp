// Record reader for the step right before the hot path: the values the reference's data layers
// parse out of their LMDB/LevelDB cursors -> a dense host feature bank + the per-video tables the sampler takes.
//   * video_shot_sentences.VideoShots           (video_shot_sentences.proto:15-20), read by
//     VideoSampledShotsDataLayer (video_sampled_shots_data_layer.cpp:184-199, 789-846): video_id, shot_ids[], one
//     caffe.Datum per shot of which only float_data is used (caffe.proto:23-37, :199,:314)
//   * video_shot_sentences.TestVideoShotWindows (video_shot_sentences.proto:22-30), read by
//     VideoShotWindowTestDataLayer (video_shot_window_test_data_layer.cpp:95-114,186-239): context / positive / negative
//     datums become consecutive rows of one item, the label is video_id
// The protobuf wire format is decoded directly (varints, length-delimited fields, packed or unpacked repeated scalars,
// unknown fields skipped) - no protobuf runtime.  Record sources: vv_record_set_add() from whatever cursor the caller owns
// (the reference-side binding keeps its mdb_cursor_get loop), or vv_record_set_load_file() for a
// length-prefixed record stream ("VVRS") and for the text dump `mdb_dump` prints.
#include <cuda_runtime.h>

#include <sys/mman.h>
#include <sys/stat.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "vv_b200.h"

namespace vv { void set_error(const char* fmt, ...); }        // vv_capi.cu (vv_last_error)
#define vv_set_error(...) (::vv::set_error(__VA_ARGS__), int(VV_ERR_INVALID))

namespace {

struct Rd {
  const uint8_t* p;
  const uint8_t* end;
  bool varint(uint64_t* v) {
    uint64_t r = 0;
    for (int sh = 0; sh < 64 && p < end; sh += 7) {
      const uint8_t b = *p++;
      r |= uint64_t(b & 0x7f) << sh;
      if (!(b & 0x80)) { *v = r; return true; }
    }
    return false;
  }
  bool bytes(Rd* sub) {
    uint64_t n;
    if (!varint(&n) || n > uint64_t(end - p)) return false;
    sub->p = p; sub->end = p + n; p += n;
    return true;
  }
  bool skip(uint32_t wire_type, int depth = 0) {
    uint64_t v; Rd s;
    switch (wire_type) {
      case 0: return varint(&v);
      case 1: if (end - p < 8) return false; p += 8; return true;
      case 2: return bytes(&s);
      case 5: if (end - p < 4) return false; p += 4; return true;
      case 3:                                  // group: skip fields up to the matching end-group tag
        if (depth > 32) return false;
        while (p < end) {
          if (!varint(&v)) return false;
          if ((v & 7) == 4) return true;
          if (!skip(uint32_t(v & 7), depth + 1)) return false;
        }
        return false;
      default: return false;
    }
  }
};

// caffe.Datum: appends float_data (field 6; `repeated float` = one fixed32 per element, or packed) to out.
bool parse_datum_floats(Rd r, std::vector<float>* out) {
  uint64_t tag;
  while (r.p < r.end) {
    if (!r.varint(&tag)) return false;
    const uint32_t field = uint32_t(tag >> 3), wt = uint32_t(tag & 7);
    if (field == 6 && wt == 5) {
      if (r.end - r.p < 4) return false;
      float f; memcpy(&f, r.p, 4); r.p += 4; out->push_back(f);
    } else if (field == 6 && wt == 2) {
      Rd s;
      if (!r.bytes(&s) || ((s.end - s.p) & 3)) return false;
      const size_t n = size_t(s.end - s.p) / 4, at = out->size();
      out->resize(at + n);
      memcpy(out->data() + at, s.p, n * 4);
    } else if (!r.skip(wt)) {
      return false;
    }
  }
  return true;
}

bool parse_int32s(Rd* r, uint32_t wt, std::vector<int32_t>* out) {
  uint64_t v;
  if (wt == 0) {
    if (!r->varint(&v)) return false;
    out->push_back(int32_t(uint32_t(v)));
    return true;
  }
  if (wt != 2) return false;
  Rd s;
  if (!r->bytes(&s)) return false;
  while (s.p < s.end) {
    if (!s.varint(&v)) return false;
    out->push_back(int32_t(uint32_t(v)));
  }
  return true;
}

}  // namespace

struct vv_record_set {
  int kind = VV_RECORD_VIDEO_SHOTS;
  bool with_pos = true, with_neg = true;
  int K = 0;                       // feature_size_: float_data_size of the first datum of the first record
  int ctx = -1, pos = -1, neg = -1;// TEST kind: sizes fixed by the first record (:97-113)
  std::vector<int32_t> video_id, row_off{0}, shot_ids;
  std::vector<float> bank;
  std::vector<float> tmp;
  std::vector<int32_t> ids_tmp, neg_ids_tmp;
};

static int add_rows(vv_record_set* s, Rd datum, const char* what) {
  s->tmp.clear();
  if (!parse_datum_floats(datum, &s->tmp)) return vv_set_error("%s: malformed Datum", what);
  if (s->K == 0) {
    s->K = int(s->tmp.size());
    if (s->K < 1) return vv_set_error("%s: the first datum holds no float_data (feature_size must be >= 1)", what);
  }
  // the reference copies float_data(0..feature_size-1) of every datum (:314,:444): longer datums are cut, a shorter
  // one is an out-of-range read there and an error here
  if (int(s->tmp.size()) < s->K)
    return vv_set_error("%s: datum with %zu floats, feature_size is %d", what, s->tmp.size(), s->K);
  s->bank.insert(s->bank.end(), s->tmp.begin(), s->tmp.begin() + s->K);
  return 0;
}

static int add_video_shots(vv_record_set* s, Rd r) {
  uint64_t tag, v = 0;
  int32_t vid = 0;
  s->ids_tmp.clear();
  const size_t rows0 = s->bank.size(), K0 = size_t(s->K);
  int n = 0;
  while (r.p < r.end) {
    if (!r.varint(&tag)) return vv_set_error("VideoShots: truncated tag");
    const uint32_t field = uint32_t(tag >> 3), wt = uint32_t(tag & 7);
    bool ok = true;
    if (field == 1 && wt == 0) { ok = r.varint(&v); vid = int32_t(uint32_t(v)); }
    else if (field == 2 && (wt == 0 || wt == 2)) ok = parse_int32s(&r, wt, &s->ids_tmp);
    else if (field == 3 && wt == 2) {
      Rd d;
      ok = r.bytes(&d);
      if (ok) { if (int rc = add_rows(s, d, "VideoShots.shot_words")) { s->bank.resize(rows0); if (!K0) s->K = 0; return rc; } ++n; }
    } else ok = r.skip(wt);
    if (!ok) { s->bank.resize(rows0); if (!K0) s->K = 0; return vv_set_error("VideoShots: malformed field %u", field); }
  }
  if (s->video_id.empty() && n == 0) return vv_set_error("VideoShots: the first record has no shot_words (the reference reads shot_words(0) for the feature size)");
  if (int(s->ids_tmp.size()) < n) {
    s->bank.resize(rows0);
    return vv_set_error("VideoShots: %d shot_words but %zu shot_ids", n, s->ids_tmp.size());
  }
  s->video_id.push_back(vid);
  s->shot_ids.insert(s->shot_ids.end(), s->ids_tmp.begin(), s->ids_tmp.begin() + n);
  s->row_off.push_back(s->row_off.back() + n);
  return 0;
}

static int add_test_windows(vv_record_set* s, Rd r) {
  uint64_t tag, v = 0;
  int32_t vid = 0;
  bool has_vid = false;
  s->ids_tmp.clear(); s->neg_ids_tmp.clear();
  std::vector<Rd> ctx, pos, neg;
  while (r.p < r.end) {
    if (!r.varint(&tag)) return vv_set_error("TestVideoShotWindows: truncated tag");
    const uint32_t field = uint32_t(tag >> 3), wt = uint32_t(tag & 7);
    bool ok = true;
    Rd d;
    if (field == 1 && wt == 0) { ok = r.varint(&v); vid = int32_t(uint32_t(v)); has_vid = true; }
    else if (field == 2 && (wt == 0 || wt == 2)) ok = parse_int32s(&r, wt, &s->ids_tmp);
    else if (field == 7 && (wt == 0 || wt == 2)) ok = parse_int32s(&r, wt, &s->neg_ids_tmp);
    else if (field == 4 && wt == 2) { ok = r.bytes(&d); pos.push_back(d); }
    else if (field == 5 && wt == 2) { ok = r.bytes(&d); ctx.push_back(d); }
    else if (field == 6 && wt == 2) { ok = r.bytes(&d); neg.push_back(d); }
    else ok = r.skip(wt);
    if (!ok) return vv_set_error("TestVideoShotWindows: malformed field %u", field);
  }
  if (!has_vid) return vv_set_error("No video id found for shot window");                       // :188
  if (s->ctx < 0) {                                                                             // first record (:97-113)
    s->ctx = int(ctx.size());
    s->pos = s->with_pos ? int(pos.size()) : 0;
    s->neg = s->with_neg ? int(neg.size()) : 0;
    if (s->ctx < 1) { s->ctx = -1; return vv_set_error("TestVideoShotWindows: context_size must be >= 1"); }
  }
  const bool first = s->video_id.empty();
  int bad = 0;
  if (int(ctx.size()) != s->ctx) bad = vv_set_error("TestVideoShotWindows: %zu context words, expected %d", ctx.size(), s->ctx);
  else if (s->with_pos && (int(pos.size()) != s->pos || int(s->ids_tmp.size()) != s->pos))     // :190-197
    bad = vv_set_error("TestVideoShotWindows: %zu positive words / %zu ids, expected %d", pos.size(), s->ids_tmp.size(), s->pos);
  else if (s->with_neg && int(neg.size()) != s->neg)                                            // :199-201
    bad = vv_set_error("TestVideoShotWindows: %zu negative words, expected %d", neg.size(), s->neg);
  if (bad) { if (first) s->ctx = s->pos = s->neg = -1; return bad; }
  const size_t rows0 = s->bank.size(), ids0 = s->shot_ids.size(), K0 = size_t(s->K);
  int rc = 0;
  for (int i = 0; i < s->ctx && !rc; ++i) { rc = add_rows(s, ctx[i], "context_shot_words"); s->shot_ids.push_back(-1); }
  for (int i = 0; i < s->pos && !rc; ++i) { rc = add_rows(s, pos[i], "positive_shot_words"); s->shot_ids.push_back(s->ids_tmp[i]); }
  for (int i = 0; i < s->neg && !rc; ++i) {
    rc = add_rows(s, neg[i], "negative_shot_words");
    s->shot_ids.push_back(i < int(s->neg_ids_tmp.size()) ? s->neg_ids_tmp[i] : -1);
  }
  if (rc) {
    s->bank.resize(rows0); s->shot_ids.resize(ids0);
    if (!K0) s->K = 0;
    if (s->video_id.empty()) s->ctx = s->pos = s->neg = -1;      // sizes come from the first ACCEPTED record
    return rc;
  }
  s->video_id.push_back(vid);
  s->row_off.push_back(s->row_off.back() + s->ctx + s->pos + s->neg);
  return 0;
}

// ---- LMDB environment (data.mdb), read-only --------------------------------------------------------------------------
// Walks the main database's B+tree in key order straight from the file, following LMDB 0.9's on-disk layout (64-bit,
// little endian): 16-byte page header {pgno u64, pad u16, flags u16, lower u16 / upper u16 (or overflow page count u32)},
// u16 node offsets from byte 16; meta pages 0/1 = header + {magic 0xBEEFC0DE, version, address, mapsize, 2 x 48-byte db
// records (the free db's first u32 doubles as the page size), last_pg, txnid}, the newer txnid wins; leaf node
// {size_lo u16, size_hi u16, flags u16, ksize u16, key, data | overflow pgno u64 when F_BIGDATA}; branch node
// {pgno_lo u16, pgno_mid u16, pgno_hi u16, ksize u16, key}.  PARITY UNPINNED: liblmdb is not in this image, so this walker
// is tested only against files laid out by tests/lmdb_writer.py from the same description.
namespace {
enum { P_BRANCH = 0x01, P_LEAF = 0x02, P_OVERFLOW = 0x04, P_META = 0x08, P_LEAF2 = 0x20, F_BIGDATA = 0x01, F_SUBDATA = 0x02, F_DUPDATA = 0x04 };
struct LmdbFile {
  const uint8_t* map = nullptr;     // read-only mapping of data.mdb: pages are touched once, in key order, no copy
  size_t map_bytes = 0, psize = 0, npages = 0;
  const uint8_t* page(uint64_t pgno) const { return pgno < npages ? map + pgno * psize : nullptr; }
  ~LmdbFile() { if (map) munmap(const_cast<uint8_t*>(map), map_bytes); }
};
template <typename T> T rd(const uint8_t* p) { T v; memcpy(&v, p, sizeof(T)); return v; }

int lmdb_walk(vv_record_set* s, const LmdbFile& f, uint64_t pgno, int depth, const char* path) {
  const uint8_t* pg = f.page(pgno);
  if (!pg || depth > 64) return vv_set_error("%s: page %llu outside the file (corrupt tree)", path, (unsigned long long)pgno);
  const uint16_t flags = rd<uint16_t>(pg + 10), lower = rd<uint16_t>(pg + 12);
  if (flags & P_LEAF2) return vv_set_error("%s: MDB_DUPFIXED pages are not supported", path);
  if (!(flags & (P_BRANCH | P_LEAF)) || lower < 16 || lower > f.psize) return vv_set_error("%s: page %llu is neither branch nor leaf", path, (unsigned long long)pgno);
  const int nkeys = (lower - 16) >> 1;
  for (int i = 0; i < nkeys; ++i) {
    const uint16_t off = rd<uint16_t>(pg + 16 + 2 * i);
    if (size_t(off) + 8 > f.psize) return vv_set_error("%s: node offset outside page %llu", path, (unsigned long long)pgno);
    const uint8_t* nd = pg + off;
    const uint16_t lo = rd<uint16_t>(nd), hi = rd<uint16_t>(nd + 2), nflags = rd<uint16_t>(nd + 4), ksize = rd<uint16_t>(nd + 6);
    if (flags & P_BRANCH) {
      const uint64_t child = uint64_t(lo) | uint64_t(hi) << 16 | uint64_t(nflags) << 32;
      if (int rc = lmdb_walk(s, f, child, depth + 1, path)) return rc;
      continue;
    }
    if (nflags & (F_SUBDATA | F_DUPDATA)) return vv_set_error("%s: sub-databases / duplicate keys are not supported", path);
    const size_t dsize = size_t(lo) | size_t(hi) << 16;
    const uint8_t* data = nd + 8 + ksize;
    if (nflags & F_BIGDATA) {
      if (size_t(off) + 8 + ksize + 8 > f.psize) return vv_set_error("%s: truncated overflow reference", path);
      const uint64_t opg = rd<uint64_t>(data);
      const uint8_t* ov = f.page(opg);
      if (!ov || !(rd<uint16_t>(ov + 10) & P_OVERFLOW) || (opg * f.psize + 16 + dsize) > f.npages * f.psize)
        return vv_set_error("%s: bad overflow page %llu", path, (unsigned long long)opg);
      data = ov + 16;
    } else if (size_t(off) + 8 + ksize + dsize > f.psize) {
      return vv_set_error("%s: node data outside page %llu", path, (unsigned long long)pgno);
    }
    if (int rc = vv_record_set_add(s, data, dsize)) return rc;
  }
  return 0;
}

int load_lmdb(vv_record_set* s, FILE* fp, const char* path) {
  LmdbFile f;
  uint8_t head[2][16 + 136];
  size_t psize = 0;
  int best = -1; uint64_t best_txn = 0;
  // meta page 1 sits at the page size recorded in meta page 0
  if (fseek(fp, 0, SEEK_SET) != 0 || fread(head[0], 1, sizeof(head[0]), fp) != sizeof(head[0])) return vv_set_error("%s: too short for an LMDB meta page", path);
  if (rd<uint32_t>(head[0] + 16) != 0xBEEFC0DEu) return vv_set_error("%s: not an LMDB data file", path);
  psize = rd<uint32_t>(head[0] + 16 + 24);
  if (psize < 512 || psize > 65536 || (psize & (psize - 1))) return vv_set_error("%s: implausible LMDB page size %zu", path, psize);
  if (fseek(fp, long(psize), SEEK_SET) != 0 || fread(head[1], 1, sizeof(head[1]), fp) != sizeof(head[1])) return vv_set_error("%s: second meta page missing", path);
  for (int m = 0; m < 2; ++m) {
    if (rd<uint32_t>(head[m] + 16) != 0xBEEFC0DEu || !(rd<uint16_t>(head[m] + 10) & P_META)) continue;
    if (rd<uint32_t>(head[m] + 20) != 1) return vv_set_error("%s: LMDB data version %u, only version 1 is known", path, rd<uint32_t>(head[m] + 20));
    const uint64_t txn = rd<uint64_t>(head[m] + 16 + 24 + 96 + 8);
    if (best < 0 || txn > best_txn) { best = m; best_txn = txn; }
  }
  if (best < 0) return vv_set_error("%s: no valid LMDB meta page", path);
  const uint8_t* meta = head[best] + 16;
  const uint8_t* maindb = meta + 24 + 48;
  if (rd<uint16_t>(maindb + 4) & ~0x08) return vv_set_error("%s: main database opened with flags 0x%x (only plain byte-string keys are supported)", path, rd<uint16_t>(maindb + 4));
  const uint64_t root = rd<uint64_t>(maindb + 40), last_pg = rd<uint64_t>(meta + 24 + 96);
  if (root == ~uint64_t(0)) return 0;                       // empty database
  f.psize = psize; f.npages = size_t(last_pg) + 1;
  if (f.npages > (uint64_t(1) << 44) / psize) return vv_set_error("%s: implausible page count", path);
  struct stat st;
  if (fstat(fileno(fp), &st) != 0 || uint64_t(st.st_size) < uint64_t(f.npages) * psize)
    return vv_set_error("%s: file shorter than its last page %llu", path, (unsigned long long)last_pg);
  f.map_bytes = f.npages * psize;
  void* m = mmap(nullptr, f.map_bytes, PROT_READ, MAP_PRIVATE, fileno(fp), 0);
  if (m == MAP_FAILED) return vv_set_error("%s: mmap failed", path);
  f.map = static_cast<const uint8_t*>(m);
  madvise(m, f.map_bytes, MADV_SEQUENTIAL);
  return lmdb_walk(s, f, root, 0, path);
}
}  // namespace


extern "C" {

vv_record_set_t* vv_record_set_create(int kind, int include_positives, int include_negatives) {
  if (kind != VV_RECORD_VIDEO_SHOTS && kind != VV_RECORD_TEST_WINDOWS) { (void)vv_set_error("vv_record_set_create: unknown kind %d", kind); return nullptr; }
  vv_record_set* s = new vv_record_set;
  s->kind = kind; s->with_pos = include_positives != 0; s->with_neg = include_negatives != 0;
  return s;
}

void vv_record_set_destroy(vv_record_set_t* s) { delete s; }

int vv_record_set_add(vv_record_set_t* s, const void* value, size_t size) {
  if (!s || (!value && size)) return vv_set_error("vv_record_set_add: null argument");
  if (s->row_off.back() > (1 << 30)) return vv_set_error("vv_record_set_add: more than 2^30 rows");
  Rd r{static_cast<const uint8_t*>(value), static_cast<const uint8_t*>(value) + size};
  return s->kind == VV_RECORD_VIDEO_SHOTS ? add_video_shots(s, r) : add_test_windows(s, r);
}

int vv_record_set_info(const vv_record_set_t* s, int64_t* records, int64_t* rows, int32_t* feature_size, int32_t* rows_per_record) {
  if (!s) return vv_set_error("vv_record_set_info: null set");
  if (records) *records = int64_t(s->video_id.size());
  if (rows) *rows = s->row_off.back();
  if (feature_size) *feature_size = s->K;
  if (rows_per_record) *rows_per_record = s->kind == VV_RECORD_TEST_WINDOWS && s->ctx > 0 ? s->ctx + s->pos + s->neg : 0;
  return 0;
}

int vv_record_set_tables(const vv_record_set_t* s, int32_t* video_id, int32_t* row_off, int32_t* shot_ids) {
  if (!s) return vv_set_error("vv_record_set_tables: null set");
  if (video_id) memcpy(video_id, s->video_id.data(), s->video_id.size() * 4);
  if (row_off) memcpy(row_off, s->row_off.data(), s->row_off.size() * 4);
  if (shot_ids) memcpy(shot_ids, s->shot_ids.data(), s->shot_ids.size() * 4);
  return 0;
}

const float* vv_record_set_bank(const vv_record_set_t* s) { return s && !s->bank.empty() ? s->bank.data() : nullptr; }

int vv_record_set_upload(const vv_record_set_t* s, float* bank_dev, vv_stream_t stream) {
  if (!s || !bank_dev) return vv_set_error("vv_record_set_upload: null argument");
  if (s->bank.empty()) return vv_set_error("vv_record_set_upload: empty record set");
  cudaStream_t cs = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemcpyAsync(bank_dev, s->bank.data(), s->bank.size() * sizeof(float), cudaMemcpyHostToDevice, cs);
  if (e == cudaSuccess) e = cudaStreamSynchronize(cs);         // the source is pageable host memory owned by the set
  if (e != cudaSuccess) return vv_set_error("vv_record_set_upload: %s", cudaGetErrorString(e));
  return 0;
}

// ---- files ---------------------------------------------------------------------------------------------------------
// "VVRS" stream: 8-byte magic "VVRS0001", then per record u32 key_len, key bytes, u64 value_len, value bytes (little endian),
// in DB key order.  mdb_dump text (`mdb_dump [-p] <env>`): header lines up to "HEADER=END", then alternating key / value
// lines, each starting with one space, hex pairs (format=bytevalue) or printable text with \xx escapes (format=print),
// closed by "DATA=END".
static int hexval(int c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1; }

static int load_mdb_dump(vv_record_set* s, FILE* f, const char* path) {
  std::string line;
  bool print_fmt = false, in_data = false, is_key = true;
  std::vector<uint8_t> val;
  int c;
  long recno = 0;
  for (;;) {
    line.clear();
    while ((c = fgetc(f)) != EOF && c != '\n') line.push_back(char(c));
    if (c == EOF && line.empty()) break;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (!in_data) {
      if (line == "format=print") print_fmt = true;
      else if (line == "format=bytevalue") print_fmt = false;
      else if (line == "HEADER=END") in_data = true;
      else if (line.compare(0, 9, "database=") == 0) return vv_set_error("%s: dumps of named sub-databases are not supported", path);
      continue;
    }
    if (line == "DATA=END") { in_data = false; continue; }
    if (line.empty() || line[0] != ' ') return vv_set_error("%s: malformed dump line (no leading space) in record %ld", path, recno);
    if (is_key) { is_key = false; continue; }                 // keys only order the records; the dump is already in key order
    is_key = true;
    val.clear();
    for (size_t i = 1; i < line.size();) {
      if (!print_fmt || line[i] == '\\') {
        if (print_fmt) {
          if (i + 1 < line.size() && line[i + 1] == '\\') { val.push_back('\\'); i += 2; continue; }
          ++i;
        }
        if (i + 1 >= line.size() || hexval(line[i]) < 0 || hexval(line[i + 1]) < 0)
          return vv_set_error("%s: bad hex escape in record %ld", path, recno);
        val.push_back(uint8_t(hexval(line[i]) * 16 + hexval(line[i + 1])));
        i += 2;
      } else {
        val.push_back(uint8_t(line[i++]));
      }
    }
    if (int rc = vv_record_set_add(s, val.data(), val.size())) return rc;
    ++recno;
  }
  if (!is_key) return vv_set_error("%s: dump ends after a key without its value", path);
  return 0;
}

int vv_record_set_load_file(vv_record_set_t* s, const char* path) {
  if (!s || !path) return vv_set_error("vv_record_set_load_file: null argument");
  std::string file = path;
  struct stat st;
  if (stat(path, &st) == 0 && S_ISDIR(st.st_mode)) {         // an LMDB environment directory (what `source:` names)
    file += "/data.mdb";
    if (stat(file.c_str(), &st) != 0)
      return vv_set_error("vv_record_set_load_file: '%s' holds no data.mdb (LevelDB directories are not supported: dump the records to a VVRS stream)", path);
  }
  FILE* f = fopen(file.c_str(), "rb");
  if (!f) return vv_set_error("vv_record_set_load_file: cannot open '%s'", path);
  char magic[24] = {0};
  const size_t got = fread(magic, 1, 24, f);
  int rc = 0;
  uint32_t lm = 0;
  memcpy(&lm, magic + 16, 4);
  if (got == 24 && lm == 0xBEEFC0DEu) {
    rc = load_lmdb(s, f, path);
  } else if (got >= 8 && memcmp(magic, "VVRS0001", 8) == 0) {
    fseek(f, 8, SEEK_SET);
    std::vector<uint8_t> buf;
    for (;;) {
      uint32_t klen; uint64_t vlen;
      if (fread(&klen, 4, 1, f) != 1) break;                  // clean end of stream
      if (fseek(f, long(klen), SEEK_CUR) != 0 || fread(&vlen, 8, 1, f) != 1 || vlen > (uint64_t(1) << 34)) { rc = vv_set_error("%s: truncated record header", path); break; }
      buf.resize(size_t(vlen));
      if (vlen && fread(buf.data(), 1, size_t(vlen), f) != size_t(vlen)) { rc = vv_set_error("%s: truncated record value", path); break; }
      if ((rc = vv_record_set_add(s, buf.data(), buf.size()))) break;
    }
  } else if (got >= 8 && memcmp(magic, "VERSION=", 8) == 0) {
    rewind(f);
    rc = load_mdb_dump(s, f, path);
  } else {
    rc = vv_set_error("%s: not an LMDB environment, a VVRS record stream or an mdb_dump text dump", path);
  }
  fclose(f);
  return rc;
}

}  // extern "C"
