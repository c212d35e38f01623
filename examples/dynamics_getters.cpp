// Whole-body getters of the raisim:: facade (getTotalMass, getCOM, getLinearMomentum, getGeneralizedMomentum, getKineticEnergy,
// getPotentialEnergy, getSparseJacobian, getJointLimits, setBasePos / setBaseOrientation, HeightMap::getHeight, jointOrder on M / h / J),
// checked against identities that involve the step kernel's own results:
//   * free flight without gravity conserves the linear momentum, and the centre of mass moves by dt P / m per step;
//   * the kinetic energy 1/2 gv^T M gv (M from the kernel's CRBA) equals the sum over the bodies of 1/2 m |v_c|^2 + 1/2 w^T I w
//     formed from the facade's Jacobians and the model tables (an independent route);
//   * the first three generalized momenta of a floating base are the linear momentum;
//   * free fall: kinetic + potential energy stays put up to the O(dt) drift of the semi-implicit scheme;
//   * a robot dropped on a height map comes to rest with its feet on HeightMap::getHeight().
#include <cmath>
#include <cstdio>
#include <string>
#include "raisim/World.hpp"

static double kineticEnergyFromBodies(raisim::ArticulatedSystem* robot, const rsb_model_tables& t) {
  const raisim::VecDyn gv = robot->getGeneralizedVelocity();
  const std::vector<raisim::Vec<3>> com = robot->getBodyCOM_W();
  double e = 0;
  for (int b = 0; b < t.nb; b++) {
    raisim::MatDyn Jp, Jr; raisim::Mat<3, 3> R;
    robot->getDenseJacobian(size_t(b), com[size_t(b)], Jp);
    robot->getDenseRotationalJacobian(size_t(b), Jr);
    robot->getBodyOrientation(size_t(b), R);
    double v[3] = {0, 0, 0}, w[3] = {0, 0, 0};
    for (size_t r = 0; r < 3; r++) for (size_t c = 0; c < gv.size(); c++) { v[r] += Jp(r, c) * gv[c]; w[r] += Jr(r, c) * gv[c]; }
    double wb[3];    // angular velocity in the body frame: R^T w
    for (int k = 0; k < 3; k++) wb[k] = R(0, size_t(k)) * w[0] + R(1, size_t(k)) * w[1] + R(2, size_t(k)) * w[2];
    const double* I = &t.inertia[6 * b];   // xx xy xz yy yz zz about the centre of mass, body axes
    const double Iw[3] = {I[0] * wb[0] + I[1] * wb[1] + I[2] * wb[2], I[1] * wb[0] + I[3] * wb[1] + I[4] * wb[2], I[2] * wb[0] + I[4] * wb[1] + I[5] * wb[2]};
    e += 0.5 * t.mass[b] * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) + 0.5 * (wb[0] * Iw[0] + wb[1] * Iw[1] + wb[2] * Iw[2]);
  }
  return e;
}

int main(int argc, char** argv) {
  std::string urdf = argc > 1 ? argv[1] : "raisimlib_b200/rsc/anymal_c_like.urdf";
  const double dt = 0.001;
  int bad = 0;
  {
    raisim::World world;
    world.setTimeStep(dt);
    world.setGravity({0.0, 0.0, 0.0});
    world.addGround(-50.0);
    auto* robot = world.addArticulatedSystem(urdf);
    robot->setControlMode(raisim::ControlMode::FORCE_AND_TORQUE);
    robot->setIntegrationScheme(raisim::IntegrationScheme::SEMI_IMPLICIT);
    bool threw = false;
    try { robot->setIntegrationScheme(raisim::IntegrationScheme::RUNGE_KUTTA_4); } catch (const std::exception&) { threw = true; }
    if (!threw) { std::printf("setIntegrationScheme(RUNGE_KUTTA_4) must fail loudly\n"); bad++; }
    const size_t nq = robot->getGeneralizedCoordinateDim(), nv = robot->getDOF();
    rsb_model_tables t; rsb_model_get_tables(world.batched()->model(), &t);
    raisim::VecDyn gc(nq), gv(nv);
    const double stance[19] = {0.3, -0.2, 0.57, 0.9689124, 0.1, -0.2, 0.1, 0.03, 0.4, -0.8, -0.03, 0.4, -0.8, 0.03, -0.4, 0.8, -0.03, -0.4, 0.8};
    for (size_t i = 0; i < nq && i < 19; i++) gc[i] = stance[i];
    { double n = std::sqrt(gc[3] * gc[3] + gc[4] * gc[4] + gc[5] * gc[5] + gc[6] * gc[6]); for (int k = 3; k < 7; k++) gc[size_t(k)] /= n; }
    for (size_t i = 0; i < nv; i++) gv[i] = 0.4 * std::sin(1.3 * double(i) + 0.2) + (i < 3 ? 0.5 : 0.0);
    robot->setState(gc, gv);

    // masses, names, limits
    double msum = 0; for (int b = 0; b < t.nb; b++) msum += robot->getMass(size_t(b));
    const double mtot = robot->getTotalMass();
    if (std::fabs(msum - mtot) > 1e-12 || !(mtot > 1.0)) { std::printf("total mass %g vs sum %g\n", mtot, msum); bad++; }
    if (robot->getBodyNames().size() != size_t(t.nb) || robot->getMovableJointNames().size() != nq - 7) { std::printf("name lists have the wrong size\n"); bad++; }
    const std::vector<raisim::Vec<2>> lim = robot->getJointLimits();
    if (lim.size() != nv || !(lim[6][0] < lim[6][1])) { std::printf("joint limits malformed\n"); bad++; }
    if (robot->getActuationUpperLimits().size() != nv || robot->getActuationLowerLimits()[6] > 0) { std::printf("actuation limits malformed\n"); bad++; }

    // momentum, energy, centre of mass along a force-free flight
    const raisim::Vec<3> P0 = robot->getLinearMomentum();
    raisim::Vec<3> cPrev = robot->getCOM();
    double worstP = 0, worstC = 0, worstK = 0, worstG = 0;
    for (int k = 0; k < 30; k++) {
      world.integrate();
      const raisim::Vec<3> P = robot->getLinearMomentum(), c = robot->getCOM();
      const raisim::VecDyn pg = robot->getGeneralizedMomentum();
      for (size_t r = 0; r < 3; r++) {
        worstP = std::fmax(worstP, std::fabs(P[r] - P0[r]));
        worstC = std::fmax(worstC, std::fabs((c[r] - cPrev[r]) / dt - P[r] / mtot));
        worstG = std::fmax(worstG, std::fabs(pg[r] - P[r]));
      }
      const double ke = robot->getKineticEnergy(), keb = kineticEnergyFromBodies(robot, t);
      worstK = std::fmax(worstK, std::fabs(ke - keb) / keb);
      cPrev = c;
    }
    std::printf("free flight: |P - P0| %.2e kg m/s   |dCOM/dt - P/m| %.2e m/s   |M gv - P| %.2e   KE (M) vs KE (bodies) rel %.2e\n", worstP, worstC, worstG, worstK);
    // float32 state: momenta to ~1e-5 relative of m |v| ~ 50 * 0.7; the COM difference quotient carries 1e-7 / dt
    if (!(worstP < 2e-3 && worstC < 2e-3 && worstG < 1e-3 && worstK < 1e-4)) bad++;

    // sparse Jacobian = the non-zero columns of the dense one
    const size_t foot = robot->getFrameIdxByName("LH_FOOT"), shank = robot->getBodyIdx("LH_SHANK");
    raisim::Vec<3> pf; robot->getFramePosition(foot, pf);
    raisim::MatDyn Jd; robot->getDenseJacobian(shank, pf, Jd);
    raisim::SparseJacobian Js; robot->getSparseJacobian(shank, pf, Js);
    double worstS = 0; size_t nonzero = 0;
    for (size_t c = 0; c < nv; c++) {
      bool listed = false; size_t at = 0;
      for (size_t k = 0; k < Js.size; k++) if (Js.idx[k] == c) { listed = true; at = k; }
      for (size_t r = 0; r < 3; r++) worstS = std::fmax(worstS, std::fabs(Jd(r, c) - (listed ? Js.v(r, at) : 0.0)));
      nonzero += listed ? 1 : 0;
    }
    std::printf("sparse Jacobian: %zu of %zu columns, max difference to the dense one %.1e\n", nonzero, nv, worstS);
    if (!(worstS == 0.0 && nonzero == 9)) bad++;      // 6 base dofs + the three joints of the leg

    // base pose setters
    robot->setBasePos({1.0, 2.0, 3.0});
    raisim::Mat<3, 3> Rz; Rz.setIdentity(); Rz(0, 0) = 0; Rz(0, 1) = -1; Rz(1, 0) = 1; Rz(1, 1) = 0;    // yaw 90 degrees
    robot->setBaseOrientation(Rz);
    raisim::Vec<3> bp; raisim::Mat<3, 3> bR;
    robot->getBasePosition(bp); robot->getBaseOrientation(bR);
    double worstB = std::fabs(bp[0] - 1.0) + std::fabs(bp[1] - 2.0) + std::fabs(bp[2] - 3.0);
    for (size_t i = 0; i < 3; i++) for (size_t j = 0; j < 3; j++) worstB = std::fmax(worstB, std::fabs(bR(i, j) - Rz(i, j)));
    std::printf("setBasePos / setBaseOrientation round trip: %.1e\n", worstB);
    if (!(worstB < 1e-6)) bad++;
  }
  {   // free fall: energy bookkeeping with gravity
    raisim::World world;
    world.setTimeStep(dt);
    world.addGround(-50.0);
    auto* robot = world.addArticulatedSystem(urdf);
    robot->setControlMode(raisim::ControlMode::FORCE_AND_TORQUE);
    const size_t nq = robot->getGeneralizedCoordinateDim(), nv = robot->getDOF();
    raisim::VecDyn gc(nq), gv(nv);
    const double stance[19] = {0, 0, 0.57, 1, 0, 0, 0, 0.03, 0.4, -0.8, -0.03, 0.4, -0.8, 0.03, -0.4, 0.8, -0.03, -0.4, 0.8};
    for (size_t i = 0; i < nq && i < 19; i++) gc[i] = stance[i];
    robot->setState(gc, gv);
    const raisim::Vec<3> g{0.0, 0.0, -9.81};
    const double e0 = robot->getEnergy(g);
    for (int k = 0; k < 100; k++) world.integrate();
    const double e1 = robot->getEnergy(g), ke = robot->getKineticEnergy();
    std::printf("free fall, 100 steps: kinetic %.3f J, total energy drift %.3e J (%.2e of the kinetic energy)\n", ke, e1 - e0, std::fabs(e1 - e0) / ke);
    if (!(ke > 10.0 && std::fabs(e1 - e0) < 0.03 * ke)) bad++;       // semi-implicit Euler: O(dt) drift
  }
  {   // HeightMap::getHeight against the surface the kernel collides with
    raisim::World world;
    world.setTimeStep(0.0025);
    auto* robot = world.addArticulatedSystem(urdf);
    const size_t xs = 17, ys = 13; const double sx = 8.0, sy = 6.0;
    std::vector<double> h(xs * ys);
    for (size_t iy = 0; iy < ys; iy++) for (size_t ix = 0; ix < xs; ix++) h[iy * xs + ix] = 0.05 * std::sin(0.9 * double(ix)) * std::cos(0.7 * double(iy));
    raisim::HeightMap* hm = world.addHeightMap(xs, ys, sx, sy, 0.5, -0.25, h);
    double worstH = 0;
    const double dx = sx / double(xs - 1), dy = sy / double(ys - 1), x0 = 0.5 - 0.5 * sx, y0 = -0.25 - 0.5 * sy;
    for (size_t iy = 0; iy < ys; iy++) for (size_t ix = 0; ix < xs; ix++)
      worstH = std::fmax(worstH, std::fabs(hm->getHeight(x0 + double(ix) * dx, y0 + double(iy) * dy) - double(float(h[iy * xs + ix]))));
    // on the P00-P11 diagonal of a cell the surface is the mean of those two corners
    worstH = std::fmax(worstH, std::fabs(hm->getHeight(x0 + 3.5 * dx, y0 + 4.5 * dy) - 0.5 * (double(float(h[4 * xs + 3])) + double(float(h[5 * xs + 4])))));
    const size_t nq = robot->getGeneralizedCoordinateDim(), nv = robot->getDOF();
    raisim::VecDyn gc(nq), gv(nv), kp(nv), kd(nv);
    const double stance[19] = {0.2, 0.1, 0.62, 1, 0, 0, 0, 0.03, 0.4, -0.8, -0.03, 0.4, -0.8, 0.03, -0.4, 0.8, -0.03, -0.4, 0.8};
    for (size_t i = 0; i < nq && i < 19; i++) gc[i] = stance[i];
    for (size_t i = 6; i < nv; i++) { kp[i] = 300.0; kd[i] = 8.0; }
    robot->setState(gc, gv);
    robot->setPGains(kp); robot->setDGains(kd);           // the separate setters: each keeps the other half
    robot->setPTarget(gc);
    for (int k = 0; k < 800; k++) world.integrate();
    double worstF = 0; int nc = 0;
    for (const raisim::Contact& c : robot->getContacts()) {
      const raisim::Vec<3> p = c.getPosition();
      worstF = std::fmax(worstF, std::fabs(p[2] - hm->getHeight(p[0], p[1])));
      nc++;
    }
    std::printf("HeightMap::getHeight: grid nodes / diagonal %.1e m; %d resting contacts lie within %.1e m of it\n", worstH, nc, worstF);
    // a sphere contact point sits on the sphere's surface along the triangle normal: within the penetration depth (< 1 mm) of the surface,
    // a little more where the closest terrain feature is an edge of a neighbouring cell
    if (!(worstH < 1e-12 && nc >= 3 && worstF < 5e-3)) bad++;
  }
  {   // jointOrder reaches M, h and the Jacobians: hind legs first
    const char* legs[4] = {"LF", "RF", "LH", "RH"};
    const char* joints[3] = {"HAA", "HFE", "KFE"};
    std::vector<std::string> order;
    for (int l = 3; l >= 0; l--) for (int j = 0; j < 3; j++) order.push_back(std::string(legs[l]) + "_" + joints[j]);
    raisim::World wa, wb;
    for (raisim::World* w : {&wa, &wb}) { w->setTimeStep(0.0025); w->addGround(-50.0); }
    auto* ra = wa.addArticulatedSystem(urdf);
    auto* rb = wb.addArticulatedSystem(urdf, "", order);
    const size_t nq = ra->getGeneralizedCoordinateDim(), nv = ra->getDOF();
    raisim::VecDyn qa(nq), va(nv), qb(nq), vb(nv);
    const double stance[19] = {0, 0, 0.57, 1, 0, 0, 0, 0.03, 0.4, -0.8, -0.05, 0.3, -0.7, 0.06, -0.45, 0.85, -0.02, -0.35, 0.75};
    for (int i = 0; i < 19; i++) qa[size_t(i)] = stance[i];
    for (size_t i = 0; i < nv; i++) va[i] = 0.2 * std::cos(0.9 * double(i));
    for (int i = 0; i < 7; i++) qb[size_t(i)] = qa[size_t(i)];
    for (int i = 0; i < 6; i++) vb[size_t(i)] = va[size_t(i)];
    auto perm = [](int l, int j) { return 3 * (3 - l) + j; };     // caller's joint (l, j) of the hind-first order -> joint index in URDF order
    for (int l = 0; l < 4; l++) for (int j = 0; j < 3; j++) { qb[size_t(7 + 3 * l + j)] = qa[size_t(7 + perm(l, j))]; vb[size_t(6 + 3 * l + j)] = va[size_t(6 + perm(l, j))]; }
    ra->setState(qa, va); rb->setState(qb, vb);
    const raisim::MatDyn Ma = ra->getMassMatrix(), Mb = rb->getMassMatrix();
    const raisim::VecDyn ha = ra->getNonlinearities(), hb = rb->getNonlinearities();
    auto toA = [&](size_t i) { return i < 6 ? i : size_t(6 + perm(int(i - 6) / 3, int(i - 6) % 3)); };
    double worstO = 0;
    for (size_t i = 0; i < nv; i++) {
      worstO = std::fmax(worstO, std::fabs(hb[i] - ha[toA(i)]));
      for (size_t j = 0; j < nv; j++) worstO = std::fmax(worstO, std::fabs(Mb(i, j) - Ma(toA(i), toA(j))));
    }
    const size_t foot = ra->getFrameIdxByName("RH_FOOT");
    raisim::MatDyn Ja, Jb; ra->getDenseFrameJacobian(foot, Ja); rb->getDenseFrameJacobian(foot, Jb);
    for (size_t i = 0; i < nv; i++) for (size_t r = 0; r < 3; r++) worstO = std::fmax(worstO, std::fabs(Jb(r, i) - Ja(r, toA(i))));
    const std::vector<std::string> names = rb->getMovableJointNames();
    if (names.size() != 12 || names[0] != "RH_HAA" || names[11] != "LF_KFE") { std::printf("getMovableJointNames ignores jointOrder\n"); bad++; }
    std::printf("jointOrder on M, h, J: max difference to the permuted URDF-order result %.1e; KE %.6f vs %.6f\n", worstO, ra->getKineticEnergy(), rb->getKineticEnergy());
    if (!(worstO == 0.0 && std::fabs(ra->getKineticEnergy() - rb->getKineticEnergy()) < 1e-9)) bad++;
  }
  std::printf(bad ? "FAILED (%d)\n" : "dynamics getters ok\n", bad);
  return bad ? 1 : 0;
}
