// BASELINE.json configs[0]: a single ANYmal-C(-like) URDF on flat ground, World::integrate() for
// 1000 steps, written against the raisim:: facade exactly as an upstream example program would be
// (the shape of raisimLib's examples/src/server/anymal.cpp, [RECALL]; compiled by the reference CI
// with -DRAISIM_EXAMPLE=ON, /root/reference/.travis.yml:11).  Runs on the GPU (batch of one).
#include <cmath>
#include <cstdio>
#include <string>
#include "raisim/World.hpp"

int main(int argc, char** argv) {
  std::string urdf = argc > 1 ? argv[1] : "raisimlib_b200/rsc/anymal_c_like.urdf";
  int steps = argc > 2 ? std::atoi(argv[2]) : 1000;
  raisim::World world;
  world.setTimeStep(0.0025);
  world.addGround(0.0);
  auto* anymal = world.addArticulatedSystem(urdf);
  raisim::VecDyn gc(anymal->getGeneralizedCoordinateDim()), gv(anymal->getDOF()), kp(anymal->getDOF()), kd(anymal->getDOF());
  const double stance[19] = {0, 0, 0.57, 1, 0, 0, 0, 0.03, 0.4, -0.8, -0.03, 0.4, -0.8, 0.03, -0.4, 0.8, -0.03, -0.4, 0.8};
  for (size_t i = 0; i < gc.size() && i < 19; i++) gc[i] = stance[i];
  for (size_t i = 6; i < kp.size(); i++) { kp[i] = 300.0; kd[i] = 8.0; }
  anymal->setState(gc, gv);
  anymal->setControlMode(raisim::ControlMode::PD_PLUS_FEEDFORWARD_TORQUE);
  anymal->setPdGains(kp, kd);
  anymal->setPdTarget(gc, gv);
  // named materials, upstream style: one foot gets a grippier sole
  anymal->getCollisionBody("LF_FOOT/0").setMaterial("rubber");
  world.setMaterialPairProp("rubber", "default", 0.95, 0.0, 0.0);
  for (int k = 0; k < steps; k++) world.integrate();
  std::printf("contact solver sweeps in the last step: %d\n", world.getContactSolver().getLoopCounter());
  anymal->getState(gc, gv);
  auto& contacts = anymal->getContacts();
  std::printf("t=%.4f s  base z=%.4f  contacts=%zu\n", world.getWorldTime(), gc[2], contacts.size());
  bool frames_ok = true;
  for (auto& c : contacts) {
    std::printf("  body %zu  depth %.5f  impulse_z %.5f\n", c.getlocalBodyIndex(), c.getDepth(), c.getImpulse()[2]);
    // upstream idiom: world-frame impulse = contactFrame^T * (impulse in the contact frame); normal part >= 0, inside the cone
    const raisim::Mat<3, 3> F = c.getContactFrame();
    const raisim::Vec<3> l = c.getImpulseInContactFrame(), w = c.getImpulse();
    for (int i = 0; i < 3; i++) {
      const double back = F(0, i) * l[0] + F(1, i) * l[1] + F(2, i) * l[2];
      if (std::fabs(back - w[i]) > 1e-6) frames_ok = false;
    }
    if (l[2] < -1e-9 || std::hypot(l[0], l[1]) > 0.95 * l[2] + 1e-6) frames_ok = false;
  }
  return (gc[2] > 0.3 && contacts.size() == 4 && frames_ok) ? 0 : 1;
}
