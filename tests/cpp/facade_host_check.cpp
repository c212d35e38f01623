// Host-only pieces of the raisim:: facade (no GPU): HeightMap::getHeight, Ground, the math helpers, Contact::getContactFrame.
// usage: facade_host_check <xs> <ys> <xSize> <ySize> <cx> <cy> <heights file> <queries file>
//   heights file: xs * ys doubles (text), x fastest; queries file: "x y" pairs.  Prints one height per query (tests/test_capi_cpu.py
//   compares them with the oracle's terrain query), then the self-checks.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>
#include "raisim/World.hpp"

int main(int argc, char** argv) {
  if (argc < 9) return 2;
  const size_t xs = size_t(std::atoi(argv[1])), ys = size_t(std::atoi(argv[2]));
  const double sx = std::atof(argv[3]), sy = std::atof(argv[4]), cx = std::atof(argv[5]), cy = std::atof(argv[6]);
  std::vector<double> h(xs * ys);
  { std::ifstream f(argv[7]); for (double& v : h) f >> v; }
  raisim::HeightMap hm;
  hm.set(xs, ys, sx, sy, cx, cy, h);
  { std::ifstream f(argv[8]); double x, y; while (f >> x >> y) std::printf("%.17g\n", hm.getHeight(x, y)); }
  int bad = 0;
  if (hm.getXSamples() != xs || hm.getYSamples() != ys || hm.getHeightVector().size() != xs * ys || hm.getCenterX() != cx) bad++;
  raisim::Ground g; g.setHeight(-0.25);
  if (g.getHeight(3.0, 4.0) != -0.25) bad++;
  // quaternion <-> rotation matrix round trip
  raisim::Vec<4> q{0.5, -0.5, 0.5, 0.5}, q2;
  raisim::Mat<3, 3> R;
  raisim::quatToRotMat(q, R); raisim::rotMatToQuat(R, q2);
  for (size_t k = 0; k < 4; k++) if (std::fabs(q[k] - q2[k]) > 1e-12) bad++;
  // contact frame: rows t1, t2, n orthonormal and right-handed, n = the contact normal; frame^T * local impulse = world impulse
  rsb_contact c{};
  c.normal[0] = 0.6f; c.normal[1] = 0.0f; c.normal[2] = 0.8f;
  c.impulse[0] = 0.3f; c.impulse[1] = -0.2f; c.impulse[2] = 1.1f;
  raisim::Contact ct(c);
  const raisim::Mat<3, 3> F = ct.getContactFrame();
  const raisim::Vec<3> l = ct.getImpulseInContactFrame();
  for (size_t i = 0; i < 3; i++) for (size_t j = 0; j < 3; j++) {
    double s = 0; for (size_t k = 0; k < 3; k++) s += F(i, k) * F(j, k);
    if (std::fabs(s - (i == j ? 1.0 : 0.0)) > 1e-6) bad++;
  }
  for (size_t k = 0; k < 3; k++) {
    if (std::fabs(F(2, k) - double(c.normal[k])) > 1e-12) bad++;
    const double w = F(0, k) * l[0] + F(1, k) * l[1] + F(2, k) * l[2];
    if (std::fabs(w - double(c.impulse[k])) > 1e-6) bad++;
  }
  const double det = F(0, 0) * (F(1, 1) * F(2, 2) - F(1, 2) * F(2, 1)) - F(0, 1) * (F(1, 0) * F(2, 2) - F(1, 2) * F(2, 0)) + F(0, 2) * (F(1, 0) * F(2, 1) - F(1, 1) * F(2, 0));
  if (std::fabs(det - 1.0) > 1e-6) bad++;
  // loud failure without a robot / for an unsupported integration scheme
  raisim::World w;
  bool threw = false;
  try { w.setERP(0.1); } catch (const std::exception&) { threw = true; }
  if (!threw) bad++;
  std::printf(bad ? "SELF-CHECKS FAILED %d\n" : "self-checks ok\n", bad);
  return bad ? 1 : 0;
}
