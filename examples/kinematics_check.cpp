// Kinematic getters of the raisim:: facade (getFramePosition / Orientation / Velocity / AngularVelocity,
// getDenseFrameJacobian, getDenseJacobian, getVelocity), checked against finite differences of the
// frame pose along a free-flight trajectory: q_{k+1} = q_k (+) dt v_{k+1}, so
// (pose_{k+1} - pose_k) / dt = J(q_k) v_{k+1} up to O(dt).
#include <cmath>
#include <cstdio>
#include <string>
#include "raisim/World.hpp"

int main(int argc, char** argv) {
  std::string urdf = argc > 1 ? argv[1] : "raisimlib_b200/rsc/anymal_c_like.urdf";
  const double dt = 0.001;
  raisim::World world;
  world.setTimeStep(dt);
  world.addGround(-50.0);                          // far below: free flight
  auto* robot = world.addArticulatedSystem(urdf);
  const size_t nq = robot->getGeneralizedCoordinateDim(), nv = robot->getDOF();
  raisim::VecDyn gc(nq), gv(nv);
  const double stance[19] = {0, 0, 0.57, 0.9689124, 0.1, -0.2, 0.1, 0.03, 0.4, -0.8, -0.03, 0.4, -0.8, 0.03, -0.4, 0.8, -0.03, -0.4, 0.8};
  for (size_t i = 0; i < nq && i < 19; i++) gc[i] = stance[i];
  { double n = std::sqrt(gc[3] * gc[3] + gc[4] * gc[4] + gc[5] * gc[5] + gc[6] * gc[6]); for (int k = 3; k < 7; k++) gc[k] /= n; }
  for (size_t i = 0; i < nv; i++) gv[i] = 0.3 * std::sin(1.7 * double(i) + 0.4) + (i >= 3 && i < 6 ? 0.8 : 0.0);
  robot->setState(gc, gv);
  robot->setControlMode(raisim::ControlMode::FORCE_AND_TORQUE);

  const size_t foot = robot->getFrameIdxByName("LH_FOOT");
  if (robot->getFrameIdxByName("LH_shank_to_foot") != foot) { std::printf("joint-name alias failed\n"); return 1; }
  const size_t shank = robot->getBodyIdx("LH_SHANK");
  double worst_v = 0, worst_w = 0, worst_orth = 0, worst_j = 0;
  raisim::Vec<3> pPrev; raisim::Mat<3, 3> RPrev; raisim::MatDyn JpPrev, JrPrev;
  for (int k = 0; k < 40; k++) {
    world.integrate();
    raisim::Vec<3> p, v, w; raisim::Mat<3, 3> R;
    robot->getFramePosition(foot, p); robot->getFrameOrientation(foot, R);
    robot->getFrameVelocity(foot, v); robot->getFrameAngularVelocity(foot, w);
    raisim::VecDyn gvNow = robot->getGeneralizedVelocity();
    // same point through the body-index API
    raisim::Vec<3> v2; robot->getVelocity(shank, p, v2);
    raisim::MatDyn J, Jf, Jr; robot->getDenseJacobian(shank, p, J);
    robot->getDenseFrameJacobian(foot, Jf); robot->getDenseFrameRotationalJacobian(foot, Jr);
    for (size_t c = 0; c < nv; c++) for (int r = 0; r < 3; r++) worst_j = std::fmax(worst_j, std::fabs(J(r, c) - Jf(r, c)));
    for (int r = 0; r < 3; r++) {
      double s = 0; for (size_t c = 0; c < nv; c++) s += Jf(r, c) * gvNow[c];
      worst_j = std::fmax(worst_j, std::fmax(std::fabs(v[r] - v2[r]), std::fabs(v[r] - s)));
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
      double s = 0; for (int c = 0; c < 3; c++) s += R(i, c) * R(j, c);
      worst_orth = std::fmax(worst_orth, std::fabs(s - (i == j ? 1.0 : 0.0)));
    }
    if (k > 0) {
      double W[3][3];   // angular velocity from R(k) R(k-1)^T ~ I + dt [w]x
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double t = 0; for (int c = 0; c < 3; c++) t += R(i, c) * RPrev(j, c); W[i][j] = t; }
      const double wfd[3] = {0.5 * (W[2][1] - W[1][2]) / dt, 0.5 * (W[0][2] - W[2][0]) / dt, 0.5 * (W[1][0] - W[0][1]) / dt};
      for (int r = 0; r < 3; r++) {   // J(q_{k-1}) v_k against the pose difference
        double s = 0, sw = 0;
        for (size_t c = 0; c < nv; c++) { s += JpPrev(r, c) * gvNow[c]; sw += JrPrev(r, c) * gvNow[c]; }
        worst_v = std::fmax(worst_v, std::fabs((p[r] - pPrev[r]) / dt - s));
        worst_w = std::fmax(worst_w, std::fabs(wfd[r] - sw));
      }
    }
    pPrev = p; RPrev = R; JpPrev = Jf; JrPrev = Jr;
  }
  // setExternalForce against the facade's own M^-1 and Jacobian: from rest in zero gravity v+ = dt M^-1 J^T F
  double worst_f = 0;
  {
    world.setGravity({0.0, 0.0, 0.0});
    raisim::VecDyn rest(nv);
    robot->setGeneralizedVelocity(rest);
    world.integrate1();
    raisim::MatDyn Minv = robot->getInverseMassMatrix();
    raisim::Vec<3> pFoot; robot->getFramePosition(foot, pFoot);
    raisim::MatDyn Jf; robot->getDenseFrameJacobian(foot, Jf);
    const raisim::Vec<3> F{4.0, -3.0, 6.0};
    raisim::Vec<3> pShank; robot->getBodyPosition(shank, pShank);
    raisim::Mat<3, 3> RShank; robot->getBodyOrientation(shank, RShank);
    raisim::Vec<3> posInBody;      // foot frame origin expressed in the shank body frame
    for (int r = 0; r < 3; r++) posInBody[r] = RShank(0, r) * (pFoot[0] - pShank[0]) + RShank(1, r) * (pFoot[1] - pShank[1]) + RShank(2, r) * (pFoot[2] - pShank[2]);
    robot->setExternalForce(shank, posInBody, F);
    world.integrate();
    raisim::VecDyn v1 = robot->getGeneralizedVelocity();
    for (size_t i = 0; i < nv; i++) {
      double s = 0;
      for (size_t c = 0; c < nv; c++) { double jtF = Jf(0, c) * F[0] + Jf(1, c) * F[1] + Jf(2, c) * F[2]; s += Minv(i, c) * jtF; }
      worst_f = std::fmax(worst_f, std::fabs(v1[i] - dt * s));
    }
    world.integrate();              // the wrench is gone: velocity stays (zero gravity, no contact), up to the O(dt |v|^2) bias
    raisim::VecDyn v2 = robot->getGeneralizedVelocity();
    for (size_t i = 0; i < nv; i++) worst_f = std::fmax(worst_f, std::fabs(v2[i] - v1[i]) * 0.1);
  }
  std::printf("external force vs dt M^-1 J^T F: %.3e\n", worst_f);
  if (!(worst_f < 5e-5)) return 1;
  // jointOrder (upstream addArticulatedSystem(urdf, resDir, jointOrder)): the same robot with its legs listed hind-first gives the same
  // motion, seen through the caller's ordering of gc / gv / targets
  {
    const char* legs[4] = {"LF", "RF", "LH", "RH"};
    const char* joints[3] = {"HAA", "HFE", "KFE"};
    std::vector<std::string> order;
    for (int l = 3; l >= 0; l--) for (int j = 0; j < 3; j++) order.push_back(std::string(legs[l]) + "_" + joints[j]);
    raisim::World wa, wb;
    for (raisim::World* w : {&wa, &wb}) { w->setTimeStep(0.0025); w->addGround(0.0); }
    auto* ra = wa.addArticulatedSystem(urdf);
    auto* rb = wb.addArticulatedSystem(urdf, "", order);
    raisim::VecDyn qa(nq), va(nv), qb(nq), vb(nv), kp(nv), kd(nv);
    const double level_stance[19] = {0, 0, 0.57, 1, 0, 0, 0, 0.03, 0.4, -0.8, -0.03, 0.4, -0.8, 0.03, -0.4, 0.8, -0.03, -0.4, 0.8};
    for (int i = 0; i < 19; i++) qa[i] = level_stance[i];
    for (int i = 0; i < 7; i++) qb[i] = qa[i];
    for (int l = 0; l < 4; l++) for (int j = 0; j < 3; j++) qb[7 + 3 * l + j] = qa[7 + 3 * (3 - l) + j];      // caller's order: RH, LH, RF, LF
    for (size_t i = 6; i < nv; i++) { kp[i] = 120.0 + double(i); kd[i] = 2.0 + 0.1 * double(i); }              // gains differ per joint: the permutation must carry them
    raisim::VecDyn kpb(nv), kdb(nv);
    for (int l = 0; l < 4; l++) for (int j = 0; j < 3; j++) { kpb[6 + 3 * l + j] = kp[6 + 3 * (3 - l) + j]; kdb[6 + 3 * l + j] = kd[6 + 3 * (3 - l) + j]; }
    ra->setState(qa, va); rb->setState(qb, vb);
    ra->setPdGains(kp, kd); rb->setPdGains(kpb, kdb);
    ra->setPdTarget(qa, va); rb->setPdTarget(qb, vb);
    for (int k = 0; k < 200; k++) { wa.integrate(); wb.integrate(); }
    ra->getState(qa, va); rb->getState(qb, vb);
    double worst_o = 0;
    for (int i = 0; i < 7; i++) worst_o = std::fmax(worst_o, std::fabs(qa[i] - qb[i]));
    for (int l = 0; l < 4; l++) for (int j = 0; j < 3; j++) worst_o = std::fmax(worst_o, std::fabs(qb[7 + 3 * l + j] - qa[7 + 3 * (3 - l) + j]) + std::fabs(vb[6 + 3 * l + j] - va[6 + 3 * (3 - l) + j]));
    std::printf("jointOrder: same motion through the caller's joint order, max difference %.3g\n", worst_o);
    if (worst_o != 0.0) return 1;
  }
  std::printf("frame velocity vs finite difference: %.3e m/s   angular: %.3e rad/s   orthonormality: %.2e   body-vs-frame API: %.2e\n",
              worst_v, worst_w, worst_orth, worst_j);
  // float32 poses differenced over dt = 1e-3: ~1e-7 / 1e-3 = 1e-4 rounding + O(dt |a|) truncation
  return (worst_v < 2e-2 && worst_w < 2e-2 && worst_orth < 1e-5 && worst_j < 1e-9) ? 0 : 1;
}
