// RaisimGym's `VectorizedEnvironment` (upstream raisimGymTorch/raisimGymTorch/env/VectorizedEnvironment.hpp with
// envs/rsg_anymal/Environment.hpp, [RECALL]; not in the reference snapshot -- SURVEY.md 8f row N1), two ways:
//
//  * VectorizedEnvironment<ENVIRONMENT> -- the generic drop-in.  N unmodified ENVIRONMENT objects (any observation, reward and
//    termination code written against raisim::World / ArticulatedSystem) run in LOCK STEP on one batch: each environment's step()
//    executes on its own fiber; World::integrate() only counts, and the first access that needs the result parks the fiber until all
//    N have arrived; then one upload per dirty array, ONE fused launch for every environment and every counted sub-step, and the
//    fibers go on (World.hpp `LockStep`).  Upstream runs `#pragma omp parallel for` over the environments instead, each integrating
//    its own World.  The host-side reward / observation code still runs on the CPU, so this is the compatible path, not the fast one.
//  * VectorizedAnymalTask -- the rsg_anymal task moved onto the device: step() is ONE C-ABI call (rsb_batch_gym_step): action rows in,
//    one fused launch, reward / done / observation rows out, nothing per environment on the host.
//
// Method names and argument meaning follow upstream; matrices are plain row-major float buffers (upstream: Eigen::Ref<EigenRowMajorMat>).
#pragma once
#include <ucontext.h>

#include <exception>
#include <string>
#include <vector>

#include "RaisimGymEnv.hpp"
#include "World.hpp"

namespace raisim {

template <class ChildEnvironment>
class VectorizedEnvironment {
 public:
  // upstream: VectorizedEnvironment(std::string resourceDir, std::string cfg); here the configuration object is handed through as is
  template <class Cfg>
  VectorizedEnvironment(std::string resourceDir, const Cfg& cfg, int numEnvs, int device = 0, size_t fiberStackBytes = 256 * 1024)
      : resourceDir_(std::move(resourceDir)), stackBytes_(fiberStackBytes) {
    ctx_.numEnvs = numEnvs; ctx_.device = device;
    BatchContext::current() = &ctx_;
    try {
      for (int i = 0; i < numEnvs; i++) { ctx_.envIndex = i; environments_.emplace_back(new ChildEnvironment(resourceDir_, cfg, false)); }
    } catch (...) { BatchContext::current() = nullptr; throw; }
    BatchContext::current() = nullptr;
    if (!ctx_.batch) throw std::runtime_error("VectorizedEnvironment: the environments created no robot");
    ls_ = ctx_.batch->lockStep();
    fibers_.resize(numEnvs);
    for (Fiber& f : fibers_) f.stack.resize(stackBytes_);
    ls_->suspend = [this](int env) { park(env); };
  }
  ~VectorizedEnvironment() { if (ls_) ls_->suspend = nullptr; }
  VectorizedEnvironment(const VectorizedEnvironment&) = delete;
  VectorizedEnvironment& operator=(const VectorizedEnvironment&) = delete;

  void init() {
    for (auto& e : environments_) e->init();
    obDim_ = environments_[0]->getObDim(); actionDim_ = environments_[0]->getActionDim();
    ls_->pushWrites();
  }
  void reset() { for (auto& e : environments_) e->reset(); ls_->pushWrites(); }
  // ob [num_envs][obDim] row-major
  void observe(float* ob) { for (int i = 0; i < getNumOfEnvs(); i++) environments_[i]->observe(RowRef{ob + size_t(i) * obDim_, obDim_}); }
  // action [num_envs][actionDim], reward [num_envs], done [num_envs] -- upstream perAgentStep() for every environment, in lock step
  void step(const float* action, float* reward, bool* done) {
    action_ = action; reward_ = reward; done_ = done;
    runAll();
  }
  int getObDim() const { return obDim_; }
  int getActionDim() const { return actionDim_; }
  int getNumOfEnvs() const { return int(environments_.size()); }
  void setSimulationTimeStep(double dt) { for (auto& e : environments_) e->setSimulationTimeStep(dt); }
  void setControlTimeStep(double dt) { for (auto& e : environments_) e->setControlTimeStep(dt); }
  void setSeed(int seed) { int k = seed; for (auto& e : environments_) e->setSeed(k++); }
  void curriculumUpdate() { for (auto& e : environments_) e->curriculumUpdate(); }
  void close() { for (auto& e : environments_) e->close(); }
  ChildEnvironment& environment(int i) { return *environments_[i]; }
  BatchedWorld& world() { return *ctx_.batch; }
  long launches() const { return ls_->launches; }          // batched launches issued so far (one per control step for the usual step())

 private:
  struct Fiber { ucontext_t ctx; std::vector<char> stack; bool done = true; std::exception_ptr error; };

  void perAgentStep(int i) {
    float r = environments_[i]->step(RowRef{const_cast<float*>(action_) + size_t(i) * actionDim_, actionDim_});
    float terminalReward = 0.f;
    done_[i] = environments_[i]->isTerminalState(terminalReward);
    if (done_[i]) { environments_[i]->reset(); r += terminalReward; }
    reward_[i] = r;
  }
  static void trampoline(unsigned lo, unsigned hi, unsigned idx) {
    auto* self = reinterpret_cast<VectorizedEnvironment*>((uintptr_t(hi) << 32) | uintptr_t(lo));
    Fiber& f = self->fibers_[idx];
    try { self->perAgentStep(int(idx)); } catch (...) { f.error = std::current_exception(); }
    f.done = true;
    swapcontext(&f.ctx, &self->main_);      // never returns
  }
  void park(int env) { swapcontext(&fibers_[env].ctx, &main_); }      // called on environment env's fiber: back to the scheduler
  void runAll() {
    const int n = getNumOfEnvs();
    const uintptr_t self = reinterpret_cast<uintptr_t>(this);
    for (int i = 0; i < n; i++) {
      Fiber& f = fibers_[i];
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack.data(); f.ctx.uc_stack.ss_size = f.stack.size(); f.ctx.uc_link = nullptr;
      f.done = false; f.error = nullptr;
      makecontext(&f.ctx, reinterpret_cast<void (*)()>(&VectorizedEnvironment::trampoline), 3, unsigned(self & 0xffffffffu), unsigned(self >> 32), unsigned(i));
    }
    for (;;) {
      int live = 0;
      for (int i = 0; i < n; i++) {
        Fiber& f = fibers_[i];
        if (f.done) continue;
        swapcontext(&main_, &f.ctx);                  // runs until the environment finishes or needs the batch to catch up
        if (f.error) { for (int& p : ls_->pending) p = 0; std::rethrow_exception(f.error); }
        if (!f.done) live++;
      }
      if (live == 0) break;
      if (live != n) throw std::runtime_error("VectorizedEnvironment: some environments finished step() while others wait for World::integrate()");
      ls_->flush();                                   // every environment waits at the same point: one upload, ONE launch
    }
    ls_->flush();                                     // integrate() calls nobody read back yet
    ls_->pushWrites();                                // resets written by terminated environments
  }

  std::string resourceDir_;
  size_t stackBytes_;
  BatchContext ctx_;
  LockStep* ls_ = nullptr;
  std::vector<std::unique_ptr<ChildEnvironment>> environments_;
  std::vector<Fiber> fibers_;
  ucontext_t main_;
  const float* action_ = nullptr; float* reward_ = nullptr; bool* done_ = nullptr;
  int obDim_ = 0, actionDim_ = 0;
};

struct AnymalTaskConfig {            // the fields of upstream's cfg.yaml that the environment reads
  int num_envs = 100;
  double simulation_dt = 0.0025, control_dt = 0.01;
  double p_gain = 50.0, d_gain = 0.2;
  double action_std = 0.6;
  double torque_reward_coeff = -4e-5, forward_vel_reward_coeff = 0.3, terminal_reward = -10.0;
  std::vector<double> gc_init = {0, 0, 0.57, 1, 0, 0, 0, 0.03, 0.4, -0.8, -0.03, 0.4, -0.8, 0.03, -0.4, 0.8, -0.03, -0.4, 0.8};
  std::vector<std::string> foot_links = {"LF_FOOT", "RF_FOOT", "LH_FOOT", "RH_FOOT"};   // upstream: LF/RF/LH/RH_SHANK bodies
};

class VectorizedAnymalTask {
 public:
  VectorizedAnymalTask(const std::string& urdf, const AnymalTaskConfig& cfg, int device = 0) : cfg_(cfg), world_(urdf, cfg.num_envs, device) {}

  void init() {
    rsb_params p = world_.params();
    p.dt = float(cfg_.simulation_dt);
    world_.setParams(p);
    rsbCheck(rsb_batch_set_ground(world_.batch(), 0.f), "addGround");
    const int nv = world_.nv(), nq = world_.nq(), nj = nq - 7;
    std::vector<float> kp(nv, 0.f), kd(nv, 0.f), gc0(nq), gv0(nv, 0.f), mean(nj), stdv(nj, float(cfg_.action_std));
    for (int i = 6; i < nv; i++) { kp[i] = float(cfg_.p_gain); kd[i] = float(cfg_.d_gain); }
    for (int i = 0; i < nq; i++) gc0[i] = float(cfg_.gc_init[i]);
    for (int i = 0; i < nj; i++) mean[i] = gc0[7 + i];
    rsbCheck(rsb_batch_set_control_mode(world_.batch(), RSB_PD_PLUS_FEEDFORWARD_TORQUE), "setControlMode");
    rsbCheck(rsb_batch_set_pd_gains(world_.batch(), kp.data(), kd.data()), "setPdGains");
    std::vector<int32_t> feet;
    for (const std::string& n : cfg_.foot_links) { int b = rsb_model_body_index(world_.model(), n.c_str()); rsbCheck(b, "getBodyIdx"); feet.push_back(b); }
    rsbCheck(rsb_batch_gym_configure(world_.batch(), gc0.data(), gv0.data(), mean.data(), stdv.data(), feet.data(), int(feet.size()),
                                     float(cfg_.torque_reward_coeff), float(cfg_.forward_vel_reward_coeff), float(cfg_.terminal_reward)), "gym_configure");
    reset();
  }
  void reset() { rsbCheck(rsb_batch_gym_reset(world_.batch()), "reset"); }
  // ob: [num_envs][obDim] row-major float
  void observe(float* ob) { rsbCheck(rsb_batch_observe(world_.batch(), ob, 0, world_.numEnvs(), RSB_HOST), "observe"); }
  // action [num_envs][actionDim], reward [num_envs], done [num_envs]; the observation of the NEXT state is
  // produced by the same call (pass ob = nullptr and call observe() to mirror upstream exactly)
  void step(const float* action, float* reward, bool* done, float* ob = nullptr) {
    static_assert(sizeof(bool) == 1, "done rows are bytes");
    rsbCheck(rsb_batch_gym_step(world_.batch(), action, RSB_HOST, substeps(), ob, reward, reinterpret_cast<unsigned char*>(done), RSB_HOST), "step");
  }
#ifdef RAISIM_B200_HAS_EIGEN
  using EigenRowMajorMat = Eigen::Matrix<float, -1, -1, Eigen::RowMajor>;
  using EigenVec = Eigen::Matrix<float, -1, 1>;
  using EigenBoolVec = Eigen::Matrix<bool, -1, 1>;
  void observe(Eigen::Ref<EigenRowMajorMat> ob) { observe(ob.data()); }
  void step(Eigen::Ref<EigenRowMajorMat> action, Eigen::Ref<EigenVec> reward, Eigen::Ref<EigenBoolVec> done) { step(action.data(), reward.data(), done.data()); }
#endif
  void setSimulationTimeStep(double dt) { cfg_.simulation_dt = dt; rsb_params p = world_.params(); p.dt = float(dt); world_.setParams(p); }
  void setControlTimeStep(double dt) { cfg_.control_dt = dt; }
  int getObDim() const { return rsb_batch_ob_dim(world_.batch()); }
  int getActionDim() const { return world_.nq() - 7; }
  int getNumOfEnvs() const { return world_.numEnvs(); }
  void setSeed(int) {}            // the device task draws no random numbers (reset() restores gc_init exactly): nothing to seed
  void close() {}                 // no visualisation server to shut down (out of scope)
  void curriculumUpdate() {}      // upstream's rsg_anymal has no curriculum either
  BatchedWorld& world() { return world_; }
 private:
  int substeps() const { return int(cfg_.control_dt / cfg_.simulation_dt + 1e-10); }
  AnymalTaskConfig cfg_;
  BatchedWorld world_;
};

}  // namespace raisim
