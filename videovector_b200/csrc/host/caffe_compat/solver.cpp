// Solver / SGDSolver (see caffe/solver.hpp for the reference lines mirrored here).
#include <cuda_runtime_api.h>
#include <cmath>
#include <cstring>
#include "caffe/solver.hpp"

namespace caffe {

template <typename Dtype>
void Solver<Dtype>::Init(const SolverParameter& param) {
  param_ = param;
  if (param_.random_seed() >= 0) Caffe::set_random_seed(unsigned(param_.random_seed()));
  CHECK(param_.solver_mode() == "GPU") << "solver_mode: CPU is not available in the B200 build (no CPU fallback)";
  const string net_file = param_.net().empty() ? param_.train_net() : param_.net();
  CHECK(!net_file.empty()) << "SolverParameter must specify a net";
  LogInfo("Creating training net from net file: " + net_file);
  net_.reset(new Net<Dtype>(net_file, Caffe::TRAIN));
  const char* fuse = getenv("VV_FUSE");
  if (!fuse || strcmp(fuse, "0") != 0) {
    string why;
    if (!net_->EnableFusion(&why)) LogInfo("net not fused (" + why + "): running layer by layer");
  }
  // TEST nets (ref: solver.cpp:104-157 InitTestNets): one per `test_net` file, else the `net` file itself, each filtered
  // to phase TEST; every test net needs its test_iter
  if (param_.test_iter_size() > 0) {
    vector<string> files;
    for (int i = 0; i < param_.test_net_size(); ++i) files.push_back(param_.test_net(i));
    if (files.empty()) files.push_back(net_file);
    CHECK_EQ(int(files.size()), param_.test_iter_size()) << "test_iter must be specified for each test network.";
    CHECK_GT(param_.test_interval(), 0) << "test_interval must be positive when test nets are given";
    for (size_t i = 0; i < files.size(); ++i) {
      LogInfo("Creating test net (#" + std::to_string(i) + ") specified by net file: " + files[i]);
      test_nets_.push_back(shared_ptr<Net<Dtype> >(new Net<Dtype>(files[i], Caffe::TEST)));
    }
  }
  iter_ = 0;
}

template <typename Dtype>
void Solver<Dtype>::TestAll() { for (size_t i = 0; i < test_nets_.size(); ++i) Test(int(i)); }

template <typename Dtype>
vector<Dtype> Solver<Dtype>::Test(const int test_net_id) {
  fprintf(stderr, "Iteration %d, Testing net (#%d)\n", iter_, test_net_id);
  Caffe::set_phase(Caffe::TEST);                       // dropout becomes a copy
  CHECK(test_nets_[test_net_id]);
  const shared_ptr<Net<Dtype> >& test_net = test_nets_[test_net_id];
  test_net->ShareTrainedLayersWith(net_.get());
  vector<Dtype> test_score;
  vector<int> test_score_output_id;
  Dtype loss = 0;
  const int iters = param_.test_iter(test_net_id);
  for (int i = 0; i < iters; ++i) {
    Dtype iter_loss = 0;
    const vector<Blob<Dtype>*>& result = test_net->ForwardPrefilled(&iter_loss);
    if (param_.test_compute_loss()) loss += iter_loss;
    int idx = 0;
    for (size_t j = 0; j < result.size(); ++j) {
      const Dtype* v = result[j]->cpu_data();
      for (int k = 0; k < result[j]->count(); ++k) {
        if (i == 0) { test_score.push_back(v[k]); test_score_output_id.push_back(int(j)); }
        else test_score[idx++] += v[k];
      }
    }
  }
  if (param_.test_compute_loss()) fprintf(stderr, "Test loss: %g\n", double(loss / iters));
  for (size_t i = 0; i < test_score.size(); ++i) {
    const int output_blob_index = test_net->output_blob_indices()[test_score_output_id[i]];
    const string& output_name = test_net->blob_names()[output_blob_index];
    const Dtype loss_weight = test_net->blob_loss_weights()[output_blob_index];
    test_score[i] /= iters;
    if (loss_weight) fprintf(stderr, "    Test net output #%zu: %s = %g (* %g = %g loss)\n", i, output_name.c_str(), double(test_score[i]),
                             double(loss_weight), double(loss_weight * test_score[i]));
    else fprintf(stderr, "    Test net output #%zu: %s = %g\n", i, output_name.c_str(), double(test_score[i]));
  }
  Caffe::set_phase(Caffe::TRAIN);
  return test_score;
}

template <typename Dtype>
Dtype Solver<Dtype>::Step() {
  if (!presolved_) { PreSolve(); presolved_ = true; }
  Dtype loss;
  if (net_->fused()) {
    vv_trainer_cfg_t sc; memset(&sc, 0, sizeof(sc));
    FillFusedSolverCfg(&sc);
    if (param_.display() && iter_ % param_.display() == 0)        // same line ComputeUpdateValue prints (solver.cpp:492-494)
      fprintf(stderr, "Iteration %d, lr = %g\n", iter_,
              double(vv_learning_rate(sc.lr_policy, sc.base_lr, sc.gamma, sc.power, sc.stepsize, iter_)));
    net_->CreateTrainer(&sc);                        // no-op after the first step
    FillFusedSolverCfg(&sc);                         // the momentum history moves into / aliases the trainer's buffers
    loss = net_->FusedStep(iter_, true, &sc);        // forward + backward + ComputeUpdateValue + Update
    FillFusedSolverCfg(&sc);                         // re-alias: the trainer wrote the buffers behind the blobs' back
  } else {
    loss = net_->ForwardBackward();
    ComputeUpdateValue();
    net_->Update();
  }
  ++iter_;
  return loss;
}

// ref: solver.cpp:320-341
template <typename Dtype>
string Solver<Dtype>::Snapshot() {
  if (!presolved_) { PreSolve(); presolved_ = true; }
  char iter_str[32]; snprintf(iter_str, sizeof(iter_str), "_iter_%d", iter_);
  const string filename = param_.snapshot_prefix() + iter_str;
  const string model_filename = filename + ".caffemodel", state_filename = filename + ".solverstate";
  LogInfo("Snapshotting to " + model_filename);
  WriteProtoToBinaryFile(*net_->ToProto(param_.snapshot_diff()), "NetParameter", model_filename);
  PbMsg state;
  state.add_scalar("iter", std::to_string(iter_));
  state.add_scalar("learned_net", model_filename);
  SnapshotSolverState(&state);
  LogInfo("Snapshotting solver state to " + state_filename);
  WriteProtoToBinaryFile(state, "SolverState", state_filename);
  return model_filename;
}
// ref: solver.cpp:418-429
template <typename Dtype>
void Solver<Dtype>::Restore(const char* state_file) {
  if (!presolved_) { PreSolve(); presolved_ = true; }
  const shared_ptr<PbMsg> state = ReadProtoFromBinaryFile(state_file, "SolverState");
  if (state->has("learned_net")) net_->CopyTrainedLayersFrom(state->str("learned_net"));
  iter_ = int(state->num("iter", 0));
  RestoreSolverState(*state);
}

// ref: solver.cpp:160-240 (the TEST-net hooks are outside the path)
template <typename Dtype>
void Solver<Dtype>::Solve(int max_iter, const char* resume_file) {
  const int stop = max_iter >= 0 ? max_iter : param_.max_iter();
  // the reference's Solve(resume_file) starts from iteration 0 unless a state is restored (solver.cpp:165-169); the
  // explicit-max_iter form continues from wherever earlier Step()/Solve(n) calls left the solver
  if (max_iter < 0) iter_ = 0;
  if (resume_file) { LogInfo(string("Restoring previous solver status from ") + resume_file); Restore(resume_file); }
  const int start_iter = iter_;
  for (; iter_ < stop;) {
    if (param_.snapshot() && iter_ > start_iter && iter_ % param_.snapshot() == 0) Snapshot();
    if (!test_nets_.empty() && iter_ % param_.test_interval() == 0 && (iter_ > 0 || param_.test_initialization())) TestAll();
    const int it = iter_;
    const Dtype loss = Step();
    if (param_.display() && it % param_.display() == 0) {
      // the fork's log format (solver.cpp:195-217): scrapers (parse_log.sh, plot_training_stats.py) key on it
      fprintf(stderr, "Iteration %d, loss = %g\n", it, double(loss));
      const vector<Blob<Dtype>*>& result = net_->output_blobs();
      int score_index = 0;
      for (size_t j = 0; j < result.size(); ++j) {
        const Dtype* v = result[j]->cpu_data();
        const string& name = net_->blob_names()[net_->output_blob_indices()[j]];
        for (int k = 0; k < result[j]->count(); ++k)
          fprintf(stderr, "    Train net output #%d: %s = iter = %d value = %g\n", score_index++, name.c_str(), it, double(v[k]));
      }
    }
  }
  // "Always save a snapshot after optimization, unless overridden by setting snapshot_after_train := false"
  // (solver.cpp:225-227) -- <snapshot_prefix>_iter_N.caffemodel / .solverstate, with an empty prefix too, as there.
  // The explicit-max_iter form (an interactive "run n more iterations") leaves snapshots to the caller.
  if (max_iter < 0 && param_.snapshot_after_train()) Snapshot();
  // the final display-only pass: forward only, the parameters were already updated max_iter times (solver.cpp:228-236)
  if (max_iter < 0 && param_.display() && iter_ % param_.display() == 0) {
    Dtype loss = 0;
    if (net_->fused()) loss = net_->FusedStep(iter_, false, nullptr);      // the fused sequence has no forward-only form; no update
    else net_->ForwardPrefilled(&loss);
    fprintf(stderr, "Iteration %d, loss = %g\n", iter_, double(loss));
  }
  if (!test_nets_.empty() && iter_ % param_.test_interval() == 0) TestAll();          // solver.cpp:237-239
}

template <typename Dtype>
Dtype SGDSolver<Dtype>::GetLearningRate() {
  // evaluated through the C-ABI so the host-side rate is the one the fused kernels use (ref: solver.cpp:441-460)
  const float r = vv_learning_rate(this->param_.lr_policy().c_str(), this->param_.base_lr(), this->param_.gamma(),
                                   this->param_.power(), this->param_.stepsize(), this->iter_);
  CHECK_GE(r, 0.f) << "Unknown learning rate policy: " << this->param_.lr_policy();
  return r;
}
template <typename Dtype>
void SGDSolver<Dtype>::PreSolve() {
  history_.clear(); update_.clear(); temp_.clear();
  vector<shared_ptr<Blob<Dtype> > >& net_params = this->net_->params();
  for (size_t i = 0; i < net_params.size(); ++i) {
    const Blob<Dtype>* p = net_params[i].get();
    history_.push_back(shared_ptr<Blob<Dtype> >(new Blob<Dtype>(p->num(), p->channels(), p->height(), p->width())));
  }
}
template <typename Dtype>
void SGDSolver<Dtype>::FillFusedSolverCfg(vv_trainer_cfg_t* c) {
  strncpy(c->lr_policy, this->param_.lr_policy().c_str(), sizeof(c->lr_policy) - 1);
  c->base_lr = this->param_.base_lr(); c->gamma = this->param_.gamma(); c->power = this->param_.power();
  c->stepsize = this->param_.stepsize(); c->momentum = this->param_.momentum(); c->weight_decay = this->param_.weight_decay();
  const string rt = this->param_.regularization_type();
  CHECK(rt == "L2" || rt == "L1") << "Unknown regularization type: " << rt;
  c->reg_type = rt == "L2" ? 2 : 1;
  // after the trainer exists the momentum history lives in its buffers; what the solver held until then (zeros, or a
  // restored .solverstate) moves over once
  if (this->net_->trainer() && history_.size() == 2) {
    float* hist[2] = {vv_trainer_weight_hist(this->net_->trainer()), vv_trainer_bias_hist(this->net_->trainer())};
    for (int i = 0; i < 2; ++i) {
      if (!history_aliased_)
        CHECK_EQ(int(cudaMemcpy(hist[i], history_[i]->gpu_data(), sizeof(Dtype) * history_[i]->count(), cudaMemcpyDeviceToDevice)), 0);
      history_[i]->set_gpu_data(hist[i]);                                         // also moves the head to the device
    }
    history_aliased_ = true;
  }
}
template <typename Dtype>
void SGDSolver<Dtype>::SnapshotSolverState(PbMsg* state) {
  for (size_t i = 0; i < history_.size(); ++i) {
    auto bp = std::make_shared<PbMsg>();
    history_[i]->ToProto(bp.get());
    state->fields.push_back(PbField{"history", "", bp, nullptr});
  }
}
template <typename Dtype>
void SGDSolver<Dtype>::RestoreSolverState(const PbMsg& state) {
  CHECK_EQ(state.count("history"), int(history_.size())) << "Incorrect length of history blobs.";
  LogInfo("SGDSolver: restoring history");
  for (size_t i = 0; i < history_.size(); ++i) {
    const shared_ptr<PbMsg> hb = state.sub("history", int(i));
    const PbField* d = hb->nth("data", 0);
    CHECK(d && d->floats && int(d->floats->size()) == history_[i]->count()) << "history blob " << i << " has the wrong size";
    memcpy(history_[i]->mutable_cpu_data(), d->floats->data(), sizeof(Dtype) * history_[i]->count());
    history_[i]->gpu_data();       // to the device copy (the trainer's buffer once aliased)
  }
}
// Layer-by-layer mode: the reference's own sequence on the device (ref: solver.cpp:534-568):
//   diff += decay * data ; hist = rate * diff + momentum * hist ; diff = hist ; then Net::Update: data -= diff.
// (The fused net does all of it, plus the split-K slab sum, in one pass: vv_sgd_update inside vv_trainer_step.)
template <typename Dtype>
void SGDSolver<Dtype>::ComputeUpdateValue() {
  vector<shared_ptr<Blob<Dtype> > >& net_params = this->net_->params();
  vector<float>& lr = this->net_->params_lr();
  vector<float>& wd = this->net_->params_weight_decay();
  const Dtype rate = GetLearningRate();
  if (this->param_.display() && this->iter_ % this->param_.display() == 0) fprintf(stderr, "Iteration %d, lr = %g\n", this->iter_, double(rate));
  const Dtype momentum = this->param_.momentum();
  const Dtype weight_decay = this->param_.weight_decay();
  const string rt = this->param_.regularization_type();
  for (size_t i = 0; i < net_params.size(); ++i) {
    Blob<Dtype>* p = net_params[i].get();
    const Dtype local_rate = rate * lr[i], local_decay = weight_decay * wd[i];
    if (local_decay) {
      if (rt == "L2") {
        VV_CHECK(vv_axpby(p->count(), local_decay, p->gpu_data(), 1.f, p->mutable_gpu_diff(), Caffe::stream()));
      } else if (rt == "L1") {       // diff += decay * sign(data)  (caffe_gpu_sign + axpy, solver.cpp:547-554)
        VV_CHECK(vv_sign_axpy(p->count(), local_decay, p->gpu_data(), p->mutable_gpu_diff(), Caffe::stream()));
      } else {
        CHECK(false) << "Unknown regularization type: " << rt;
      }
    }
    VV_CHECK(vv_axpby(p->count(), local_rate, p->gpu_diff(), momentum, history_[i]->mutable_gpu_data(), Caffe::stream()));
    VV_CHECK(vv_axpby(p->count(), 1.f, history_[i]->gpu_data(), 0.f, p->mutable_gpu_diff(), Caffe::stream()));
  }
}

template <typename Dtype>
Solver<Dtype>* GetSolver(const SolverParameter& param) { return new SGDSolver<Dtype>(param); }
template Solver<float>* GetSolver(const SolverParameter& param);
template class Solver<float>;
template class SGDSolver<float>;

}  // namespace caffe
