// caffe_compat -- a dependency-free host that mirrors the reference's operator interface
// (Blob / Layer<Dtype> / Net / SGDSolver, ref: include/caffe/{common,blob,layer,net,solver}.hpp) so that the
// drop-in layer classes of this repo compile and run without protobuf, glog, gflags or boost, and so that
// the very same class bodies can be pasted over the reference's .cu files.
// Only the GPU mode exists here: there is no CPU implementation to fall back to.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "vv_b200.h"

namespace caffe {

using std::shared_ptr;
using std::string;
using std::vector;

// glog's LOG(FATAL)/CHECK abort the process (ref: include/caffe/util/device_alternate.hpp:48-67 wraps every
// CUDA call that way).  Here a failed CHECK throws FatalError; the CLI tool lets it terminate the process,
// the C test API turns it into an error code + message.
struct FatalError : std::runtime_error {
  explicit FatalError(const string& m) : std::runtime_error(m) {}
};
struct CheckFail {
  std::ostringstream os;
  CheckFail(const char* file, int line, const char* expr) { os << file << ":" << line << " Check failed: " << expr << " "; }
  [[noreturn]] ~CheckFail() noexcept(false) { throw FatalError(os.str()); }
  template <class T> CheckFail& operator<<(const T& v) { os << v; return *this; }
};
// `if (ok) ; else fail`: safe inside an unbraced if / else, and quiet under -Wdangling-else (the shape glog's own macros have)
#if defined(__GNUC__)
#pragma GCC diagnostic ignored "-Wdangling-else"   // `if (x) CHECK(...)`: the else inside the macro is the intended binding
#endif
#define CHECK(cond) switch (0) case 0: default: if (cond) ; else ::caffe::CheckFail(__FILE__, __LINE__, #cond)
#define CHECK_OP(a, b, op) switch (0) case 0: default: if ((a) op (b)) ; else ::caffe::CheckFail(__FILE__, __LINE__, #a " " #op " " #b) << "(" << (a) << " vs " << (b) << ") "
#define CHECK_EQ(a, b) CHECK_OP(a, b, ==)
#define CHECK_NE(a, b) CHECK_OP(a, b, !=)
#define CHECK_LE(a, b) CHECK_OP(a, b, <=)
#define CHECK_LT(a, b) CHECK_OP(a, b, <)
#define CHECK_GE(a, b) CHECK_OP(a, b, >=)
#define CHECK_GT(a, b) CHECK_OP(a, b, >)
#define LOG_FATAL ::caffe::CheckFail(__FILE__, __LINE__, "FATAL")
#define NOT_IMPLEMENTED LOG_FATAL << "Not Implemented Yet"
#define NO_CPU LOG_FATAL << "CPU mode is not available in the B200 build: there is no CPU fallback"
// status of a C-ABI call, the way the reference wraps CUDA calls in CUDA_CHECK
#define VV_CHECK(call) do { int _rc = (call); CHECK_EQ(_rc, 0) << vv_last_error(); } while (0)

void LogInfo(const string& msg);   // glog-style "I..." line on stderr (quiet unless VV_CAFFE_VERBOSE=1)

// The process-global singleton of the reference (ref: include/caffe/common.hpp:70-147), reduced to what
// the path reads: mode, phase, the stream kernels are launched on, and the seed of the dropout stream.
class Caffe {
 public:
  enum Brew { CPU, GPU };
  enum Phase { TRAIN, TEST };
  static Caffe& Get();
  static Brew mode() { return Get().mode_; }
  static Phase phase() { return Get().phase_; }
  static void set_mode(Brew m) { if (m == CPU) { NO_CPU; } Get().mode_ = m; }
  static void set_phase(Phase p) { Get().phase_ = p; }
  static void set_random_seed(unsigned int seed) { Get().seed_ = seed; Get().draws_ = 0; }
  static void SetDevice(int device_id);
  static vv_stream_t stream() { return Get().stream_; }
  static void set_stream(vv_stream_t s) { Get().stream_ = s; }
  // the dropout layers draw one Philox sub-stream per Forward call
  static uint64_t rng_seed() { return Get().seed_; }
  static uint64_t next_rng_draw() { return Get().draws_++; }
  // how the fc7 GEMMs compute (vv_precision); default = the fp32-parity mode, env VV_PRECISION overrides
  static int precision() { return Get().prec_; }
  static void set_precision(int p) { Get().prec_ = p; }
 private:
  Caffe();
  Brew mode_; Phase phase_; vv_stream_t stream_; uint64_t seed_, draws_; int prec_;
};

}  // namespace caffe
