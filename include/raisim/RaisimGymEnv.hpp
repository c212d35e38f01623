// raisim::RaisimGymEnv -- the base class of a RaisimGym ENVIRONMENT (upstream raisimGymTorch/env/RaisimGymEnv.hpp, [RECALL]; not in
// the reference snapshot).  Same members and virtuals as upstream so that an Environment.hpp written for RaisimGym keeps its shape:
// the constructor builds world_ and the robot, step() sets the action, loops world_->integrate() and returns the reward.
// Observation / action rows are plain float spans here (upstream: Eigen::Ref<EigenVec>; Eigen is absent from this image, see math.hpp).
#pragma once
#include <memory>
#include <string>

#include "World.hpp"

namespace raisim {

struct RowRef {             // a row of a row-major float matrix (what Eigen::Ref<EigenVec> is to upstream environments)
  float* p; int n;
  float& operator[](int i) { return p[i]; }
  float operator[](int i) const { return p[i]; }
  int size() const { return n; }
};

class RaisimGymEnv {
 public:
  explicit RaisimGymEnv(std::string resourceDir) : resourceDir_(std::move(resourceDir)) {}
  virtual ~RaisimGymEnv() = default;
  virtual void init() = 0;
  virtual void reset() = 0;
  virtual void observe(RowRef ob) = 0;
  virtual float step(const RowRef& action) = 0;
  virtual bool isTerminalState(float& terminalReward) = 0;
  virtual void curriculumUpdate() {}
  virtual void close() {}
  virtual void setSeed(int) {}
  void setSimulationTimeStep(double dt) { simulation_dt_ = dt; world_->setTimeStep(dt); }
  void setControlTimeStep(double dt) { control_dt_ = dt; }
  int getObDim() const { return obDim_; }
  int getActionDim() const { return actionDim_; }
  double getControlTimeStep() const { return control_dt_; }
  double getSimulationTimeStep() const { return simulation_dt_; }
  raisim::World* getWorld() { return world_.get(); }
 protected:
  std::unique_ptr<raisim::World> world_;
  double simulation_dt_ = 0.0025, control_dt_ = 0.01;
  std::string resourceDir_;
  int obDim_ = 0, actionDim_ = 0;
};

}  // namespace raisim
