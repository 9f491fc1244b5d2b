// Solver / SGDSolver (ref: include/caffe/solver.hpp:17-143, src/caffe/solver.cpp:160-240 Solve,
// :441-460 GetLearningRate, :464-482 PreSolve, :486-576 ComputeUpdateValue).
#pragma once
#include "caffe/net.hpp"

namespace caffe {

template <typename Dtype>
class Solver {
 public:
  explicit Solver(const SolverParameter& param) : param_(param), iter_(0) { Init(param); }
  virtual ~Solver() {}
  void Init(const SolverParameter& param);
  // resume_file: a .solverstate to continue from (ref: solver.cpp:160-172)
  virtual void Solve(int max_iter = -1, const char* resume_file = nullptr);
  // <snapshot_prefix>_iter_N.caffemodel + .solverstate (ref: solver.cpp:320-341); returns the model file name
  string Snapshot();
  void Restore(const char* state_file);                       // ref: solver.cpp:418-429
  // one iteration: ForwardBackward + ComputeUpdateValue + Update (fused into one kernel sequence when the net is fused)
  Dtype Step();
  inline shared_ptr<Net<Dtype> > net() { return net_; }
  inline const vector<shared_ptr<Net<Dtype> > >& test_nets() { return test_nets_; }
  // ref: solver.cpp:243-317: runs test_iter forward passes of a TEST-phase net sharing the trained layers, logs and
  // returns the mean of every output value
  void TestAll();
  vector<Dtype> Test(const int test_net_id = 0);
  int iter() const { return iter_; }
  const SolverParameter& param() const { return param_; }
 protected:
  virtual void PreSolve() {}
  virtual void ComputeUpdateValue() = 0;
  virtual void FillFusedSolverCfg(vv_trainer_cfg_t* cfg) = 0;
  virtual void SnapshotSolverState(PbMsg* state) = 0;
  virtual void RestoreSolverState(const PbMsg& state) = 0;
  SolverParameter param_;
  int iter_;
  shared_ptr<Net<Dtype> > net_;
  vector<shared_ptr<Net<Dtype> > > test_nets_;
  bool presolved_ = false;
};

template <typename Dtype>
class SGDSolver : public Solver<Dtype> {
 public:
  explicit SGDSolver(const SolverParameter& param) : Solver<Dtype>(param) {}
  const vector<shared_ptr<Blob<Dtype> > >& history() { return history_; }
  Dtype GetLearningRate();
 protected:
  virtual void PreSolve();
  virtual void ComputeUpdateValue();
  virtual void FillFusedSolverCfg(vv_trainer_cfg_t* cfg);
  virtual void SnapshotSolverState(PbMsg* state);             // ref: solver.cpp:576-596
  virtual void RestoreSolverState(const PbMsg& state);
  bool history_aliased_ = false;
  vector<shared_ptr<Blob<Dtype> > > history_, update_, temp_;
};

template <typename Dtype> Solver<Dtype>* GetSolver(const SolverParameter& param);

}  // namespace caffe
