// A RaisimGym-shaped ENVIRONMENT (the structure of upstream raisimGymTorch/env/envs/rsg_anymal/Environment.hpp, [RECALL]) with its OWN
// observation, reward and termination -- none of which the library knows about.  It is written against raisim::World / ArticulatedSystem
// exactly as for upstream: the constructor builds world_ and the robot, step() sets the PD target, loops world_->integrate()
// control_dt / simulation_dt times, reads the state back and computes the reward on the host.  Run standalone it owns a batch of one;
// inside raisim::VectorizedEnvironment<ENVIRONMENT> the same object code runs in lock step with N - 1 others on one GPU batch.
#pragma once
#include <cmath>
#include <string>
#include <vector>

#include "raisim/RaisimGymEnv.hpp"

namespace raisim {

struct CustomCfg {                 // upstream passes a Yaml node; the fields are what matters
  std::string urdf = "raisimlib_b200/rsc/anymal_c_like.urdf";
  double simulation_dt = 0.0025, control_dt = 0.01;
  double action_scale = 0.3, p_gain = 80.0, d_gain = 1.5;
  double height_target = 0.52;
};

class ENVIRONMENT : public RaisimGymEnv {
 public:
  ENVIRONMENT(const std::string& resourceDir, const CustomCfg& cfg, bool /*visualizable*/) : RaisimGymEnv(resourceDir), cfg_(cfg) {
    world_ = std::make_unique<raisim::World>();
    world_->setTimeStep(cfg.simulation_dt);
    anymal_ = world_->addArticulatedSystem(cfg.urdf);
    anymal_->setName("anymal");
    world_->addGround();
    simulation_dt_ = cfg.simulation_dt; control_dt_ = cfg.control_dt;
    gcDim_ = int(anymal_->getGeneralizedCoordinateDim()); gvDim_ = int(anymal_->getDOF()); nJoints_ = gvDim_ - 6;
    gc_.setZero(gcDim_); gc_init_.setZero(gcDim_); gv_.setZero(gvDim_); gv_init_.setZero(gvDim_);
    pTarget_.setZero(gcDim_); vTarget_.setZero(gvDim_);
    const double stance[19] = {0, 0, 0.57, 1, 0, 0, 0, 0.03, 0.4, -0.8, -0.03, 0.4, -0.8, 0.03, -0.4, 0.8, -0.03, -0.4, 0.8};
    for (int i = 0; i < gcDim_; i++) gc_init_[i] = stance[i];
    VecDyn kp(gvDim_), kd(gvDim_);
    for (int i = 6; i < gvDim_; i++) { kp[i] = cfg.p_gain; kd[i] = cfg.d_gain; }
    anymal_->setControlMode(raisim::ControlMode::PD_PLUS_FEEDFORWARD_TORQUE);
    anymal_->setPdGains(kp, kd);
    obDim_ = 1 + 4 + nJoints_ + 3;          // height, base quaternion, joint angles, base linear velocity (world)
    actionDim_ = nJoints_;
    for (const char* n : {"LF_FOOT", "RF_FOOT", "LH_FOOT", "RH_FOOT"}) footIndices_.push_back(anymal_->getBodyIdx(n));
  }
  void init() final { reset(); }
  void reset() final {
    anymal_->setState(gc_init_, gv_init_);
    anymal_->setPdTarget(gc_init_, gv_init_);
    gc_ = gc_init_; gv_ = gv_init_;
  }
  float step(const RowRef& action) final {
    for (int i = 0; i < gcDim_; i++) pTarget_[i] = gc_init_[i];
    for (int i = 0; i < nJoints_; i++) pTarget_[7 + i] += cfg_.action_scale * action[i];
    anymal_->setPdTarget(pTarget_, vTarget_);
    const int loopCount = int(control_dt_ / simulation_dt_ + 1e-10);
    for (int i = 0; i < loopCount; i++) world_->integrate();
    anymal_->getState(gc_, gv_);
    const VecDyn tau = anymal_->getGeneralizedForce();
    double torque = 0;
    for (size_t i = 6; i < tau.size(); i++) torque += tau[i] * tau[i];
    // a reward of our own: hold the height, move forward, spend little torque, do not spin
    return float(-20.0 * (gc_[2] - cfg_.height_target) * (gc_[2] - cfg_.height_target) + 0.5 * gv_[0] - 2e-5 * torque - 0.05 * std::fabs(gv_[5]));
  }
  void observe(RowRef ob) final {
    ob[0] = float(gc_[2]);
    for (int i = 0; i < 4; i++) ob[1 + i] = float(gc_[3 + i]);
    for (int i = 0; i < nJoints_; i++) ob[5 + i] = float(gc_[7 + i]);
    for (int i = 0; i < 3; i++) ob[5 + nJoints_ + i] = float(gv_[i]);
  }
  bool isTerminalState(float& terminalReward) final {
    terminalReward = -5.f;
    for (auto& contact : anymal_->getContacts()) {
      bool foot = false;
      for (size_t f : footIndices_) foot |= contact.getlocalBodyIndex() == f;
      if (!foot) return true;
    }
    terminalReward = 0.f;
    return false;
  }
  const VecDyn& gc() const { return gc_; }

 private:
  CustomCfg cfg_;
  raisim::ArticulatedSystem* anymal_ = nullptr;
  int gcDim_ = 0, gvDim_ = 0, nJoints_ = 0;
  VecDyn gc_, gc_init_, gv_, gv_init_, pTarget_, vTarget_;
  std::vector<size_t> footIndices_;
};

}  // namespace raisim
