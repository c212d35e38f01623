// raisim::VectorizedEnvironment for the ANYmal locomotion task -- batched drop-in for RaisimGym's
// `VectorizedEnvironment<ENVIRONMENT>` (upstream raisimGymTorch/raisimGymTorch/env/VectorizedEnvironment.hpp
// with envs/rsg_anymal/Environment.hpp, [RECALL]; not in the reference snapshot -- SURVEY.md 8f row N1).
//
// Upstream runs `#pragma omp parallel for` over N ENVIRONMENT objects, each calling world_->integrate()
// control_dt/simulation_dt times.  Here step() is ONE C-ABI call (rsb_batch_gym_step): action rows in,
// one fused launch for every environment and sub-step, reward / done / observation rows out.
// Method names and argument meaning follow upstream; matrices are plain row-major float buffers
// (upstream: Eigen::Ref<EigenRowMajorMat>; Eigen overloads are provided when Eigen is available).
#pragma once
#include <string>
#include <vector>

#include "World.hpp"

namespace raisim {

struct AnymalTaskConfig {            // the fields of upstream's cfg.yaml that the environment reads
  int num_envs = 100;
  double simulation_dt = 0.0025, control_dt = 0.01;
  double p_gain = 50.0, d_gain = 0.2;
  double action_std = 0.6;
  double torque_reward_coeff = -4e-5, forward_vel_reward_coeff = 0.3, terminal_reward = -10.0;
  std::vector<double> gc_init = {0, 0, 0.57, 1, 0, 0, 0, 0.03, 0.4, -0.8, -0.03, 0.4, -0.8, 0.03, -0.4, 0.8, -0.03, -0.4, 0.8};
  std::vector<std::string> foot_links = {"LF_FOOT", "RF_FOOT", "LH_FOOT", "RH_FOOT"};   // upstream: LF/RF/LH/RH_SHANK bodies
};

class VectorizedEnvironment {
 public:
  VectorizedEnvironment(const std::string& urdf, const AnymalTaskConfig& cfg, int device = 0) : cfg_(cfg), world_(urdf, cfg.num_envs, device) {}

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
  void setSeed(int) {}
  void close() {}
  void curriculumUpdate() {}
  BatchedWorld& world() { return world_; }
 private:
  int substeps() const { return int(cfg_.control_dt / cfg_.simulation_dt + 1e-10); }
  AnymalTaskConfig cfg_;
  BatchedWorld world_;
};

}  // namespace raisim
