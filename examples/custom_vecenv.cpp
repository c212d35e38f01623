// VectorizedEnvironment<ENVIRONMENT>: N objects of an unmodified, upstream-shaped environment class (examples/rsg_custom/Environment.hpp:
// its own observation, reward and termination code on the host) stepped in lock step on ONE GPU batch -- one fused launch per control
// step although every environment calls world_->integrate() four times.  Checks: launches per control step, and that environment 0 of
// the batch went through exactly the states of the same class run standalone (a World of its own = a batch of one) on the same actions.
#include <cstdio>
#include <random>
#include <vector>

#include "raisim/VectorizedEnvironment.hpp"
#include "rsg_custom/Environment.hpp"

int main(int argc, char** argv) {
  raisim::CustomCfg cfg;
  if (argc > 1) cfg.urdf = argv[1];
  const int N = argc > 2 ? std::atoi(argv[2]) : 256, steps = 40;
  raisim::VectorizedEnvironment<raisim::ENVIRONMENT> vec("", cfg, N);
  vec.init();
  const int A = vec.getActionDim(), O = vec.getObDim();
  std::vector<float> action(size_t(N) * A), ob(size_t(N) * O), reward(N);
  std::vector<char> done(N);
  std::mt19937 rng(3);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<std::vector<float>> actions0;                // what environment 0 was told to do
  std::vector<std::vector<float>> obs0; std::vector<float> rew0;
  const long l0 = vec.launches();
  long terminated = 0; double rsum = 0;
  for (int k = 0; k < steps; k++) {
    for (float& a : action) a = nd(rng);
    for (int i = N / 2; i < N; i++) for (int j = 0; j < A; j++) action[size_t(i) * A + j] *= 4.f;     // the second half flails and falls: terminations and resets
    actions0.emplace_back(action.begin(), action.begin() + A);
    vec.step(action.data(), reward.data(), reinterpret_cast<bool*>(done.data()));
    vec.observe(ob.data());
    obs0.emplace_back(ob.begin(), ob.begin() + O); rew0.push_back(reward[0]);
    for (int i = 0; i < N; i++) { terminated += done[i]; rsum += reward[i]; }
  }
  const long launches = vec.launches() - l0;
  std::printf("%d environments x %d control steps: %ld batched launches (%.2f per control step), %ld terminations, mean reward %.4f\n", N, steps, launches,
              double(launches) / steps, terminated, rsum / (double(N) * steps));
  // the same class, standalone
  raisim::ENVIRONMENT solo("", cfg, false);
  solo.init();
  double worst = 0;
  std::vector<float> o(O);
  for (int k = 0; k < steps; k++) {
    float r = solo.step(raisim::RowRef{actions0[k].data(), A});
    float tr = 0.f;
    if (solo.isTerminalState(tr)) { solo.reset(); r += tr; }
    solo.observe(raisim::RowRef{o.data(), O});
    for (int i = 0; i < O; i++) worst = std::fmax(worst, std::fabs(double(o[i]) - obs0[k][i]));
    worst = std::fmax(worst, std::fabs(double(r) - rew0[k]));
  }
  std::printf("environment 0 of the batch vs the same class standalone: max |difference| over observations and rewards %.3g\n", worst);
  const bool ok = launches == steps && worst == 0.0 && terminated > 0;
  std::printf(ok ? "ok\n" : "FAILED\n");
  return ok ? 0 : 1;
}
