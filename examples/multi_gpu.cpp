// BASELINE configs[4] shape from C++: one rsb_batch per GPU inside one process, environments sharded in contiguous
// blocks, one NCCL all-gather of the observation rows per control step -- through the C-ABI only.
#include <cuda_runtime_api.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "rsb.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ < 0) { std::fprintf(stderr, "%s failed: %s\n", #call, rsb_last_error()); return 2; } } while (0)

int main(int argc, char** argv) {
  const char* urdf = argc > 1 ? argv[1] : "raisimlib_b200/rsc/anymal_c_like.urdf";
  int ndev = argc > 2 ? std::atoi(argv[2]) : 1, per_gpu = argc > 3 ? std::atoi(argv[3]) : 1024;
  int visible = 0;
  cudaGetDeviceCount(&visible);
  if (ndev > visible) ndev = visible;
  rsb_model* model = nullptr;
  CHECK(rsb_model_create_from_urdf(urdf, &model));
  int nq, nv, nb;
  CHECK(rsb_model_dims(model, &nq, &nv, &nb, nullptr, nullptr));
  std::vector<rsb_batch*> batch(ndev);
  const float stance[19] = {0, 0, 0.57f, 1, 0, 0, 0, 0.03f, 0.4f, -0.8f, -0.03f, 0.4f, -0.8f, 0.03f, -0.4f, 0.8f, -0.03f, -0.4f, 0.8f};
  std::vector<float> gc((size_t)per_gpu * nq), gv((size_t)per_gpu * nv, 0.f), kp(nv, 0.f), kd(nv, 0.f);
  for (int i = 6; i < nv; i++) { kp[i] = 300.f; kd[i] = 8.f; }
  for (int d = 0; d < ndev; d++) {
    CHECK(rsb_batch_create(model, per_gpu, d, &batch[d]));
    CHECK(rsb_batch_set_ground(batch[d], 0.f));
    for (int e = 0; e < per_gpu; e++) for (int i = 0; i < nq && i < 19; i++) gc[(size_t)e * nq + i] = stance[i] + (i == 2 ? 0.001f * d : 0.f);
    CHECK(rsb_batch_set_state(batch[d], gc.data(), gv.data(), 0, per_gpu, RSB_HOST));
    CHECK(rsb_batch_set_pd_gains(batch[d], kp.data(), kd.data()));
    CHECK(rsb_batch_set_pd_target(batch[d], gc.data(), gv.data(), 0, per_gpu, RSB_HOST));
  }
  rsb_comm* comm = nullptr;
  CHECK(rsb_comm_init(batch.data(), ndev, &comm));
  const int od = rsb_batch_ob_dim(batch[0]);
  const size_t total = (size_t)ndev * per_gpu * od;
  std::vector<float*> obs_all(ndev);
  for (int d = 0; d < ndev; d++) { cudaSetDevice(d); cudaMalloc((void**)&obs_all[d], total * sizeof(float)); }
  for (int k = 0; k < 50; k++) {
    for (int d = 0; d < ndev; d++) CHECK(rsb_batch_integrate(batch[d], 4));      // asynchronous, one stream per GPU
    CHECK(rsb_comm_allgather_obs(comm, obs_all.data()));
  }
  for (int d = 0; d < ndev; d++) CHECK(rsb_batch_sync(batch[d]));
  std::vector<float> host(total);
  cudaSetDevice(0);
  cudaMemcpy(host.data(), obs_all[0], total * sizeof(float), cudaMemcpyDeviceToHost);
  bool ok = true;
  for (int d = 0; d < ndev; d++) {           // rank order: block d starts with base heights of GPU d's environments
    float z = host[(size_t)d * per_gpu * od];
    std::printf("gpu %d: first env base height %.4f\n", d, z);
    ok = ok && std::isfinite(z) && z > 0.3f && z < 0.8f;
  }
  rsb_comm_destroy(comm);
  for (int d = 0; d < ndev; d++) { cudaSetDevice(d); cudaFree(obs_all[d]); rsb_batch_destroy(batch[d]); }
  rsb_model_destroy(model);
  std::printf("%d GPU(s) x %d envs x 50 control steps, all-gather of %zu floats per step: %s\n", ndev, per_gpu, total, ok ? "ok" : "FAILED");
  return ok ? 0 : 1;
}
