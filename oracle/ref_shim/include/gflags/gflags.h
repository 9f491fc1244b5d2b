#pragma once
#include <string>
namespace google { inline void ParseCommandLineFlags(int*, char***, bool) {} }
#define DEFINE_int32(name, val, txt) int FLAGS_##name = val
#define DEFINE_string(name, val, txt) std::string FLAGS_##name = val
#define DEFINE_bool(name, val, txt) bool FLAGS_##name = val
#define DECLARE_int32(name) extern int FLAGS_##name
