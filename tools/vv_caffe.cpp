// vv_caffe -- the `caffe train` / `caffe time` entry points for this path (ref: tools/caffe.cpp:80-122 train,
// :194-266 time) on the caffe_compat host.  Usage:
//   vv_caffe train --solver=solver.prototxt [--snapshot=x.solverstate | --weights=x.caffemodel] [--gpu=0] [--iterations=N]
//   vv_caffe time  --model=net.prototxt [--iterations=50] [--gpu=0]
// Environment: VV_PRECISION=tf32x3|f16x3|tf32|bf16|fp32_simt, VV_FUSE=0 to run layer by layer.
#include <chrono>
#include <cstdio>
#include <cstring>
#include "caffe/solver.hpp"

using namespace caffe;
extern "C" int vvc_device_synchronize();

static string flag(int argc, char** argv, const char* name, const string& dflt) {
  const string key = string("--") + name + "=";
  for (int i = 2; i < argc; ++i) if (!strncmp(argv[i], key.c_str(), key.size())) return argv[i] + key.size();
  return dflt;
}

static int train(int argc, char** argv) {
  const string solver_file = flag(argc, argv, "solver", "");
  CHECK(!solver_file.empty()) << "Need a solver definition to train.";
  SolverParameter sp = ReadSolverParamsFromTextFileOrDie(solver_file);
  Caffe::SetDevice(atoi(flag(argc, argv, "gpu", std::to_string(sp.device_id())).c_str()));
  Caffe::set_mode(Caffe::GPU);
  shared_ptr<Solver<float> > solver(GetSolver<float>(sp));
  const int iters = atoi(flag(argc, argv, "iterations", "-1").c_str());
  // ref: tools/caffe.cpp:106-118 -- resume from a .solverstate, or finetune from a .caffemodel
  const string snapshot = flag(argc, argv, "snapshot", ""), weights = flag(argc, argv, "weights", "");
  CHECK(snapshot.empty() || weights.empty()) << "Give a snapshot to resume training or weights to finetune but not both.";
  if (!weights.empty()) { fprintf(stderr, "Finetuning from %s\n", weights.c_str()); solver->net()->CopyTrainedLayersFrom(weights); }
  if (!snapshot.empty()) fprintf(stderr, "Resuming from %s\n", snapshot.c_str());
  fprintf(stderr, "Starting Optimization (%s)\n", solver->net()->fused() ? "fused kernel sequence" : "layer by layer");
  const auto t0 = std::chrono::steady_clock::now();
  solver->Solve(iters, snapshot.empty() ? nullptr : snapshot.c_str());
  vvc_device_synchronize();
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  fprintf(stderr, "Optimization Done: %d iterations in %.3f s\n", solver->iter(), sec);
  return 0;
}

static int time_net(int argc, char** argv) {
  const string model = flag(argc, argv, "model", "");
  CHECK(!model.empty()) << "Need a model definition to time.";
  Caffe::SetDevice(atoi(flag(argc, argv, "gpu", "0").c_str()));
  Net<float> net(model, Caffe::TRAIN);
  const int iters = atoi(flag(argc, argv, "iterations", "50").c_str());
  float loss = net.ForwardBackward();
  fprintf(stderr, "Initial loss: %g\n", loss);
  const vector<shared_ptr<Layer<float> > >& layers = net.layers();
  // whole-net forward-backward wall clock after a device synchronise (the reference's `caffe time` uses cudaEvent pairs)
  vvc_device_synchronize();
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < iters; ++i) net.ForwardBackward();
  vvc_device_synchronize();
  const float ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  fprintf(stderr, "%zu layers, %d iterations: forward-backward %.3f ms / iteration\n", layers.size(), iters, ms / iters);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: vv_caffe train|time --solver=... | --model=...\n"); return 1; }
  try {
    if (!strcmp(argv[1], "train")) return train(argc, argv);
    if (!strcmp(argv[1], "time")) return time_net(argc, argv);
    fprintf(stderr, "unknown action %s\n", argv[1]);
    return 1;
  } catch (const FatalError& e) {
    // glog's LOG(FATAL): message, then abort
    fprintf(stderr, "F %s\n*** Check failure stack trace: ***\n", e.what());
    abort();
  }
}
