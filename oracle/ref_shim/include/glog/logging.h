// ref_shim: a minimal stand-in for glog so the reference's sources compile unmodified (test infrastructure only).
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <stdexcept>
namespace google {
struct LogMessageFatal {
  std::ostringstream os;
  LogMessageFatal(const char* f, int l) { os << f << ":" << l << " "; }
  [[noreturn]] ~LogMessageFatal() noexcept(false) { throw std::runtime_error(os.str()); }
  std::ostream& stream() { return os; }
};
struct NullStream { template <class T> NullStream& operator<<(const T&) { return *this; } NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; } };
inline void InitGoogleLogging(const char*) {}
inline void InstallFailureSignalHandler() {}
}  // namespace google
#define LOG_INFO ::google::NullStream()
#define LOG_WARNING ::google::NullStream()
#define LOG_ERROR ::google::NullStream()
#define LOG_FATAL ::google::LogMessageFatal(__FILE__, __LINE__).stream()
#define LOG(sev) LOG_##sev
#define DLOG(sev) ::google::NullStream()
#define LOG_IF(sev, cond) if (!(cond)) ; else LOG_##sev
#define VLOG(n) ::google::NullStream()
#define LOG_EVERY_N(sev, n) LOG_##sev
// (if (ok) ; else ...: safe inside an unbraced if / else, and quiet under -Wdangling-else, like glog's own macros)
#define CHECK(c) switch (0) case 0: default: if (c) ; else LOG_FATAL << "Check failed: " #c " "
#define CHECK_OP_(a, b, op) switch (0) case 0: default: if ((a) op (b)) ; else LOG_FATAL << "Check failed: " #a " " #op " " #b " (" << (a) << " vs. " << (b) << ") "
#define CHECK_EQ(a, b) CHECK_OP_(a, b, ==)
#define CHECK_NE(a, b) CHECK_OP_(a, b, !=)
#define CHECK_LE(a, b) CHECK_OP_(a, b, <=)
#define CHECK_LT(a, b) CHECK_OP_(a, b, <)
#define CHECK_GE(a, b) CHECK_OP_(a, b, >=)
#define CHECK_GT(a, b) CHECK_OP_(a, b, >)
#define CHECK_NOTNULL(p) (p)
#define DCHECK(c) CHECK(c)
#define DCHECK_EQ(a, b) CHECK_EQ(a, b)
#define DCHECK_NE(a, b) CHECK_NE(a, b)
#define DCHECK_LE(a, b) CHECK_LE(a, b)
#define DCHECK_LT(a, b) CHECK_LT(a, b)
#define DCHECK_GE(a, b) CHECK_GE(a, b)
#define DCHECK_GT(a, b) CHECK_GT(a, b)
