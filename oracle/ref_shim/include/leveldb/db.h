#pragma once
// Shim of the LevelDB API surface the reference's data layers compile against (never opened: Open fails).
#include <string>
namespace leveldb {
struct Status { bool ok_; bool ok() const { return ok_; } std::string ToString() const { return "leveldb is not available in the shim"; } };
struct Slice { std::string s; std::string ToString() const { return s; } };
struct Options { bool create_if_missing = false, error_if_exists = false; long write_buffer_size = 0; int max_open_files = 0; int block_size = 0; };
struct ReadOptions {};
struct WriteOptions {};
class WriteBatch { public: void Put(const std::string&, const std::string&) {} };
class Iterator { public: void SeekToFirst() {} bool Valid() const { return false; } void Next() {} Slice key() const { return Slice(); } Slice value() const { return Slice(); } };
class DB {
 public:
  static Status Open(const Options&, const std::string&, DB** db) { *db = nullptr; return Status{false}; }
  Iterator* NewIterator(const ReadOptions&) { return new Iterator(); }
  Status Write(const WriteOptions&, WriteBatch*) { return Status{false}; }
};
}
