// RaisimGym-style use of the batched drop-in: N environments, random actions, one call per control step.
#include <cstdio>
#include <random>
#include <vector>
#include "raisim/VectorizedEnvironment.hpp"

int main(int argc, char** argv) {
  std::string urdf = argc > 1 ? argv[1] : "raisimlib_b200/rsc/anymal_c_like.urdf";
  raisim::AnymalTaskConfig cfg;
  cfg.num_envs = argc > 2 ? std::atoi(argv[2]) : 512;
  raisim::VectorizedAnymalTask env(urdf, cfg);
  env.init();
  const int N = env.getNumOfEnvs(), A = env.getActionDim(), O = env.getObDim();
  std::vector<float> action(size_t(N) * A), ob(size_t(N) * O), reward(N);
  std::vector<char> done(N);
  std::mt19937 rng(0);
  std::normal_distribution<float> nd(0.f, 0.5f);
  double rsum = 0; long terminated = 0;
  for (int k = 0; k < 100; k++) {
    for (float& a : action) a = nd(rng);
    env.step(action.data(), reward.data(), reinterpret_cast<bool*>(done.data()), ob.data());
    for (int i = 0; i < N; i++) { rsum += reward[i]; terminated += done[i]; }
  }
  std::printf("%d envs x 100 control steps: mean reward %.4f, terminations %ld, ob[0][0] (base height) %.3f\n", N, rsum / (100.0 * N), terminated, ob[0]);
  return (ob[0] > 0.2f && ob[0] < 0.8f) ? 0 : 1;
}
