#pragma once
#include <string>
namespace leveldb { class DB {}; class Iterator {}; class WriteBatch {}; struct Options { bool create_if_missing, error_if_exists; long write_buffer_size; int max_open_files; }; struct ReadOptions {}; struct Status { bool ok() const { return true; } std::string ToString() const { return ""; } }; }
